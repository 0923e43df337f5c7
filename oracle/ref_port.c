/* CPU port of the reference's convolve_signal path.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the algorithm the reference (liebharc/basic_dsp v0.10.0) runs on the CPU for
 * `ConvolutionOps::convolve_signal` on a complex f32 vector, used (a) as a second, independent
 * checker next to oracle/dsp_oracle.py and (b) as the timed CPU baseline of bench.py
 * (`cpu_baseline.kind = "port"`, `bench.py --impl reference`).  Never linked into the product.
 *
 * The reference cannot be built here (no Rust toolchain; its FFT is the un-vendored crate
 * rustfft ^6.0.0, vector/Cargo.toml:40), so the FFT below is our own iterative radix-4/2 transform;
 * everything around it follows the reference line by line:
 *   convolve_signal dispatch        vector/src/vector_types/time_freq/convolution.rs:477-542
 *   overlap_discard                 convolution.rs:304-461  (incl. the scalar head/tail loops and the
 *                                   `remainder_len = x_len - x_len % fft_len` quirk, :341,:387-397)
 *   convolve_iteration              time_freq/mod.rs:456-473 (ReverseWrappingIterator :788-848)
 *   convolve_signal_scalar          time_freq/mod.rs:275-361
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle/ref_port.c -o oracle/_build/libref_port.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float re, im; } c32;

static inline c32 cmul(c32 a, c32 b) { c32 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline c32 cadd(c32 a, c32 b) { c32 r = { a.re + b.re, a.im + b.im }; return r; }

/* ---- stand-in for rustfft: in-place unnormalised power-of-two FFT ------------------------------- */
typedef struct { int n; int inverse; c32* tw; int* rev; } fft_plan;

static fft_plan* fft_plan_create(int n, int inverse) {
    fft_plan* p = (fft_plan*)malloc(sizeof(fft_plan));
    p->n = n; p->inverse = inverse;
    p->tw = (c32*)malloc(sizeof(c32) * (size_t)n);
    p->rev = (int*)malloc(sizeof(int) * (size_t)n);
    int bits = 0; while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        double a = (inverse ? 2.0 : -2.0) * M_PI * (double)i / (double)n;
        p->tw[i].re = (float)cos(a); p->tw[i].im = (float)sin(a);
        int r = 0; for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        p->rev[i] = r;
    }
    return p;
}
static void fft_plan_destroy(fft_plan* p) { free(p->tw); free(p->rev); free(p); }

static void fft_process(const fft_plan* p, c32* x) {
    const int n = p->n;
    for (int i = 0; i < n; i++) { int r = p->rev[i]; if (r > i) { c32 t = x[i]; x[i] = x[r]; x[r] = t; } }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < half; j++) {
                c32 w = p->tw[j * step];
                c32 u = x[i + j], v = cmul(x[i + j + half], w);
                x[i + j].re = u.re + v.re; x[i + j].im = u.im + v.im;
                x[i + j + half].re = u.re - v.re; x[i + j + half].im = u.im - v.im;
            }
        }
    }
}

static size_t next_power_of_two(size_t value) { /* convolution.rs:270-283 */
    size_t n = value; int count = 0;
    if (n != 0 && (n & (n - 1)) == 0) return n;
    while (n != 0) { n >>= 1; count++; }
    return (size_t)1 << count;
}

/* convolve_iteration (time_freq/mod.rs:456-473): sum_k data[(i + conv_len - 1 - k) mod n] * other[k] */
static inline c32 convolve_iteration(const c32* data, size_t n, const c32* other, size_t full_conv_len, long i, long conv_len) {
    long pos = (i + conv_len) % (long)n;
    if (pos < 0) pos += (long)n;
    c32 sum = { 0.f, 0.f };
    for (size_t k = 0; k < full_conv_len; k++) {
        pos = pos > 0 ? pos - 1 : (long)n - 1;      /* ReverseWrappingIterator pre-decrements */
        sum = cadd(sum, cmul(data[pos], other[k]));
    }
    return sum;
}

/* convolve_signal_scalar (time_freq/mod.rs:275-361), x.points >= h.points, out of place */
void ref_convolve_signal_scalar_c32(const c32* x, size_t n, const c32* h, size_t l, c32* y) {
    const long conv_len = (long)(l - l / 2);
    for (size_t i = 0; i < n; i++) y[i] = convolve_iteration(x, n, h, l, (long)i, conv_len);
}

/* overlap_discard (convolution.rs:304-461), in place on x.  fft_len_hint as passed by convolve_signal
 * (:536) = next_power_of_two(impulse_response.len()), len in f32 scalars = 2*l. */
int ref_overlap_discard_c32(c32* x, size_t x_len, const c32* h, size_t imp_len) {
    const size_t overlap = imp_len - 1;
    const size_t min_fft_len = next_power_of_two(4 * overlap);
    size_t fft_len = next_power_of_two(2 * imp_len);
    if (fft_len < min_fft_len) fft_len = min_fft_len;
    if (x_len < fft_len) return -1;
    fft_plan* fwd = fft_plan_create((int)fft_len, 0);
    fft_plan* inv = fft_plan_create((int)fft_len, 1);
    const size_t step_size = fft_len - overlap;
    size_t remainder_len = x_len - x_len % fft_len;          /* :341, in f32 scalars of `end` */
    c32* H = (c32*)calloc(fft_len, sizeof(c32));
    c32* signal_freq = (c32*)malloc(sizeof(c32) * fft_len);
    c32* tmp = (c32*)malloc(sizeof(c32) * fft_len);
    c32* end = (c32*)malloc(sizeof(c32) * (remainder_len / 2 + 1));
    c32* overlap_buffer = (c32*)malloc(sizeof(c32) * (overlap + 1));
    memcpy(H, h, sizeof(c32) * imp_len);
    fft_process(fwd, H);                                      /* :372 */
    const long cl = (long)((imp_len + 1) / 2);
    /* (1) scalar convolution of the beginning :376-385 */
    for (size_t p = 0; p < imp_len / 2; p++) tmp[p] = convolve_iteration(x, x_len, h, imp_len, (long)p, cl);
    /* (2) scalar convolution of the tail :387-397 */
    {
        size_t position = x_len - remainder_len / 2;
        for (size_t q = 0; q < remainder_len / 2; q++) end[q] = convolve_iteration(x, x_len, h, imp_len, (long)(position + q), cl);
    }
    const float scaling = (float)fft_len;
    size_t position = 0;
    /* (3) first block :419-432 */
    memcpy(overlap_buffer, x + position + step_size, sizeof(c32) * overlap);
    memcpy(signal_freq, x + position, sizeof(c32) * fft_len);
    fft_process(fwd, signal_freq);
    memcpy(x, tmp, sizeof(c32) * (imp_len / 2));
    for (size_t k = 0; k < fft_len; k++) { c32 v = cmul(signal_freq[k], H[k]); signal_freq[k].re = v.re / scaling; signal_freq[k].im = v.im / scaling; }
    memcpy(tmp, signal_freq, sizeof(c32) * fft_len);
    fft_process(inv, tmp);
    position += step_size;
    /* (4) :434-451 */
    while (position + fft_len < x_len) {
        memcpy(x + position, overlap_buffer, sizeof(c32) * overlap);
        memcpy(overlap_buffer, x + position + step_size, sizeof(c32) * overlap);
        memcpy(signal_freq, x + position, sizeof(c32) * fft_len);
        fft_process(fwd, signal_freq);
        memcpy(x + position - step_size + imp_len / 2, tmp + imp_len - 1, sizeof(c32) * (fft_len - imp_len + 1));
        for (size_t k = 0; k < fft_len; k++) { c32 v = cmul(signal_freq[k], H[k]); signal_freq[k].re = v.re / scaling; signal_freq[k].im = v.im / scaling; }
        memcpy(tmp, signal_freq, sizeof(c32) * fft_len);
        fft_process(inv, tmp);
        position += step_size;
    }
    /* (5) :456-458 */
    {
        size_t dst = position - step_size + imp_len / 2;
        size_t cnt = fft_len - imp_len + 1;
        if (dst + cnt > x_len) cnt = x_len - dst;
        memcpy(x + dst, tmp + imp_len - 1, sizeof(c32) * cnt);
    }
    /* (6) :460 */
    memcpy(x + x_len - remainder_len / 2, end, sizeof(c32) * (remainder_len / 2));
    free(H); free(signal_freq); free(tmp); free(end); free(overlap_buffer);
    fft_plan_destroy(fwd); fft_plan_destroy(inv);
    return 0;
}

/* convolve_signal dispatch for complex vectors (convolution.rs:477-542) without the SIMD branch
 * (impulse responses of more than 101 points): overlap_discard when len > 10000 scalars, h.len > 15,
 * len > 10*h.len; else the scalar loop.  In place on x (y is scratch of n points). */
int ref_convolve_signal_c32(c32* x, size_t n, const c32* h, size_t l, c32* scratch) {
    if (n < l) return 7;
    const size_t len = 2 * n, hlen = 2 * l;
    if (len > 10000 && hlen > 15 && len > 10 * hlen) {
        if (ref_overlap_discard_c32(x, n, h, l) == 0) return 0;
    }
    ref_convolve_signal_scalar_c32(x, n, h, l, scratch);
    memcpy(x, scratch, sizeof(c32) * n);
    return 0;
}

/* batch of `rows` independent vectors, one OpenMP thread per vector (a caller parallelising the
 * reference's sequential row loop, matrix/src/time_freq.rs:52-74) */
int ref_convolve_signal_rows_c32(c32* x, size_t n, size_t rows, const c32* h, size_t l, int threads) {
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (long r = 0; r < (long)rows; r++) {
        c32* scratch = (c32*)malloc(sizeof(c32) * n);
        int e = ref_convolve_signal_c32(x + (size_t)r * n, n, h, l, scratch);
        free(scratch);
        if (e) rc = e;
    }
    return rc;
}

/* plain batched FFT (the reference's fft() on every row: rustfft pass + swap_halves pass) */
int ref_fft_rows_c32(c32* x, size_t n, size_t rows, int shift, int threads) {
    if (n & (n - 1)) return -1;
    fft_plan* p = fft_plan_create((int)n, 0);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long r = 0; r < (long)rows; r++) {
        c32* row = x + (size_t)r * n;
        fft_process(p, row);
        if (shift) for (size_t i = 0; i < n / 2; i++) { c32 t = row[i]; row[i] = row[i + n / 2]; row[i + n / 2] = t; }
    }
    fft_plan_destroy(p);
    return 0;
}
