/* Plain-C use of the drop-in C ABI (the calls the reference's examples/basic_dsp_example.py makes through ctypes).
 *   gcc -std=c11 -Iinclude examples/c_example.c -Lbasic_dsp_b200 -lbasic_dsp_b200 -Wl,-rpath,$PWD/basic_dsp_b200 -lm -o c_example
 * Needs a CUDA device at run time (the library has no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "basic_dsp_b200.h"

int main(void) {
    const size_t points = 1 << 16, taps = 1023;
    float* x = (float*)malloc(2 * points * sizeof(float));
    float* h = (float*)malloc(2 * taps * sizeof(float));
    for (size_t i = 0; i < points; i++) { x[2 * i] = (float)cos(0.001 * (double)i); x[2 * i + 1] = 0.f; }
    for (size_t k = 0; k < taps; k++) { h[2 * k] = k == taps / 2 ? 1.f : 0.f; h[2 * k + 1] = 0.f; }   /* identity filter */

    BdspVec32* v = new32(1, 0, 0.f, 2 * points, 1.f);      /* complex, time domain */
    BdspVec32* ir = new32(1, 0, 0.f, 2 * taps, 1.f);
    if (bdsp_upload32(v, x, 2 * points) || bdsp_upload32(ir, h, 2 * taps)) { fprintf(stderr, "upload: %s\n", bdsp_last_error()); return 1; }

    BdspVecResult32 r = convolve_signal32(v, ir);           /* continue with the returned handle */
    v = r.vector;
    if (r.result_code) { fprintf(stderr, "convolve_signal32: code %d\n", r.result_code); return 1; }
    r = fft32(v); v = r.vector;
    if (r.result_code) { fprintf(stderr, "fft32: code %d\n", r.result_code); return 1; }
    r = magnitude32(v); v = r.vector;
    if (r.result_code) { fprintf(stderr, "magnitude32: code %d\n", r.result_code); return 1; }

    BdspStatistics32 st = real_statistics32(v);
    printf("spectrum peak %.1f at bin %zu of %zu (expected near the centre: fft32 shifts DC to the middle)\n",
           st.max, st.max_index, get_len32(v));
    delete_vector32(v);
    delete_vector32(ir);
    free(x);
    free(h);
    return 0;
}
