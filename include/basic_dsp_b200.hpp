// C++ host-side mirror of the reference's vector traits on top of the C ABI (basic_dsp_b200.h).
//
// The reference is Rust; no Rust toolchain exists in the build image, so the compiled-language host
// side is this header.  `GpuVec<T>` owns one device-resident vector handle and offers the trait
// methods of the hot path under the reference's names and argument meaning:
//   TimeToFrequencyDomainOperations  plain_fft, fft            (time_to_freq.rs:14-71)
//   FrequencyToTimeDomainOperations  plain_ifft, ifft          (freq_to_time.rs:16-73)
//   FrequencyDomainOperations        fft_shift, ifft_shift     (freq.rs:11-41)
//   ConvolutionOps / Convolution     convolve_signal, convolve (convolution.rs:17-62)
//   FrequencyMultiplication          multiply_frequency_response (convolution.rs:65-84)
//   InterpolationOps                 interpolatef              (interpolation.rs:18-90)
//   RealInterpolationOps             interpolate_lin           (real_interpolation.rs:10-24)
//   ScaleOps / OffsetOps / ElementaryOps   scale, offset, add, sub, mul, div (elementary.rs:13-162)
//   ComplexToRealTransformsOps / GetterOps magnitude, phase, get_mag_phase (complex_to_real.rs:17-330)
// Methods that consume `self` and change the vector's type in Rust (fft, magnitude, ...) mutate the
// object and return *this.  Errors (`VoidResult` / `TransRes` in Rust) are reported as `DspError`
// carrying the reference's numeric error code (interop/src/lib.rs:107-151).  The `Buffer` argument of
// the Rust traits has no counterpart: every handle owns its device scratch.
#pragma once
#include <complex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "basic_dsp_b200.h"

namespace basic_dsp_b200 {

enum class Domain : int { Time = 0, Frequency = 1 };
enum class Response : int { Sinc = 0, RaisedCosine = 1 };

struct DspError : std::runtime_error {
    int code;
    DspError(int c, const std::string& what) : std::runtime_error(what + ": result code " + std::to_string(c)), code(c) {}
};

namespace detail {
template <typename T> struct Api;
#define BDSP_API_STRUCT(S, T, H, R)                                                                              \
    template <> struct Api<T> {                                                                                   \
        typedef H Handle;                                                                                         \
        typedef R Result;                                                                                         \
        static Handle* create(int c, int d, T init, size_t len, T delta) { return new##S(c, d, init, len, delta); } \
        static void destroy(Handle* h) { delete_vector##S(h); }                                                   \
        static Handle* clone(Handle* h) { return clone##S(h); }                                                   \
        static int upload(Handle* h, const T* p, size_t n) { return bdsp_upload##S(h, p, n); }                    \
        static int download(const Handle* h, T* p, size_t n) { return bdsp_download##S(h, p, n); }                \
        static size_t len(const Handle* h) { return get_len##S(h); }                                              \
        static size_t points(const Handle* h) { return get_points##S(h); }                                        \
        static bool is_complex(const Handle* h) { return is_complex##S(h) != 0; }                                 \
        static int domain(const Handle* h) { return get_domain##S(h); }                                           \
        static T delta(const Handle* h) { return get_delta##S(h); }                                               \
        static Result plain_fft(Handle* h) { return plain_fft##S(h); }                                            \
        static Result fft(Handle* h) { return fft##S(h); }                                                        \
        static Result plain_ifft(Handle* h) { return plain_ifft##S(h); }                                          \
        static Result ifft(Handle* h) { return ifft##S(h); }                                                      \
        static Result fft_shift(Handle* h) { return fft_shift##S(h); }                                            \
        static Result ifft_shift(Handle* h) { return ifft_shift##S(h); }                                          \
        static Result convolve_signal(Handle* h, const Handle* ir) { return convolve_signal##S(h, ir); }          \
        static Result convolve(Handle* h, int k, T ro, T ra, size_t l) { return convolve##S(h, k, ro, ra, l); }   \
        static Result mul_freq(Handle* h, int k, T ro, T ra) { return multiply_frequency_response##S(h, k, ro, ra); } \
        static Result interpolatef(Handle* h, int k, T ro, T f, T d, size_t l) { return interpolatef##S(h, k, ro, f, d, l); } \
        static Result interpolate_lin(Handle* h, T f, T d) { return interpolate_lin##S(h, f, d); }                \
        static Result real_scale(Handle* h, T c) { return real_scale##S(h, c); }                                  \
        static Result complex_scale(Handle* h, T re, T im) { return complex_scale##S(h, re, im); }                \
        static Result real_offset(Handle* h, T c) { return real_offset##S(h, c); }                                \
        static Result add(Handle* h, const Handle* o) { return add##S(h, o); }                                    \
        static Result sub(Handle* h, const Handle* o) { return sub##S(h, o); }                                    \
        static Result mul(Handle* h, const Handle* o) { return mul##S(h, o); }                                    \
        static Result div(Handle* h, const Handle* o) { return div##S(h, o); }                                    \
        static Result magnitude(Handle* h) { return magnitude##S(h); }                                            \
        static Result phase(Handle* h) { return phase##S(h); }                                                    \
        static Result to_complex(Handle* h) { return to_complex##S(h); }                                          \
        static int get_mag_phase(Handle* h, Handle* m, Handle* p) { return get_mag_phase##S(h, m, p); }           \
        static int scale_mul_mag_phase(Handle* h, T re, T im, const Handle* w, Handle* m, Handle* p, int wb) {    \
            return bdsp_scale_mul_mag_phase##S(h, re, im, w, m, p, wb);                                           \
        }                                                                                                         \
        static Result apply_window(Handle* h, int w) { return apply_window##S(h, w); }                            \
        static Result unapply_window(Handle* h, int w) { return unapply_window##S(h, w); }                        \
        static Result windowed_fft(Handle* h, int w) { return windowed_fft##S(h, w); }                            \
        static Result windowed_ifft(Handle* h, int w) { return windowed_ifft##S(h, w); }                          \
        static Result prepare_argument(Handle* h) { return prepare_argument##S(h); }                              \
        static Result prepare_argument_padded(Handle* h) { return prepare_argument_padded##S(h); }                \
        static Result correlate(Handle* h, const Handle* o) { return correlate##S(h, o); }                        \
        static Result interpolatei(Handle* h, int k, T ro, int f) { return interpolatei##S(h, k, ro, f); }        \
        static Result interpolate(Handle* h, int k, T ro, size_t n, T d) { return interpolate##S(h, k, ro, n, d); } \
        static Result interpft(Handle* h, size_t n) { return interpft##S(h, n); }                                 \
        static Result decimatei(Handle* h, uint32_t f, uint32_t d) { return decimatei##S(h, f, d); }              \
        static Result plain_sfft(Handle* h) { return plain_sfft##S(h); }                                          \
        static Result plain_sifft(Handle* h) { return plain_sifft##S(h); }                                        \
        static Result conj(Handle* h) { return conj##S(h); }                                                      \
        static T real_sum(const Handle* h) { return real_sum##S(h); }                                             \
    };
BDSP_API_STRUCT(32, float, BdspVec32, BdspVecResult32)
BDSP_API_STRUCT(64, double, BdspVec64, BdspVecResult64)
#undef BDSP_API_STRUCT
}  // namespace detail

template <typename T> class GpuVec {
    typedef detail::Api<T> A;
    typename A::Handle* h_;

    GpuVec& take(typename A::Result r, const char* what) {
        h_ = r.vector;  // callers continue with the returned handle (interop/src/lib.rs:203-212)
        if (r.result_code != 0) throw DspError(r.result_code, what);
        return *this;
    }

public:
    GpuVec(const std::vector<T>& real, Domain d = Domain::Time, T delta = 1) : h_(A::create(0, (int)d, 0, real.size(), delta)) {
        if (int rc = A::upload(h_, real.data(), real.size())) throw DspError(rc, "upload");
    }
    GpuVec(const std::vector<std::complex<T>>& cplx, Domain d = Domain::Time, T delta = 1)
        : h_(A::create(1, (int)d, 0, 2 * cplx.size(), delta)) {
        if (int rc = A::upload(h_, reinterpret_cast<const T*>(cplx.data()), 2 * cplx.size())) throw DspError(rc, "upload");
    }
    GpuVec(size_t len, bool is_complex, Domain d = Domain::Time, T init = 0, T delta = 1) : h_(A::create(is_complex, (int)d, init, len, delta)) {}
    GpuVec(const GpuVec& o) : h_(A::clone(o.h_)) {}
    GpuVec(GpuVec&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    GpuVec& operator=(GpuVec o) { std::swap(h_, o.h_); return *this; }
    ~GpuVec() { if (h_) A::destroy(h_); }

    // MetaData / Vector<T>
    size_t len() const { return A::len(h_); }
    size_t points() const { return A::points(h_); }
    bool is_complex() const { return A::is_complex(h_); }
    Domain domain() const { return (Domain)A::domain(h_); }
    T delta() const { return A::delta(h_); }
    std::vector<T> to_vec() const {
        std::vector<T> out(len());
        if (int rc = A::download(h_, out.data(), out.size())) throw DspError(rc, "download");
        return out;
    }
    typename A::Handle* handle() { return h_; }

    // transforms
    GpuVec& plain_fft() { return take(A::plain_fft(h_), "plain_fft"); }
    GpuVec& fft() { return take(A::fft(h_), "fft"); }
    GpuVec& plain_ifft() { return take(A::plain_ifft(h_), "plain_ifft"); }
    GpuVec& ifft() { return take(A::ifft(h_), "ifft"); }
    GpuVec& fft_shift() { return take(A::fft_shift(h_), "fft_shift"); }
    GpuVec& ifft_shift() { return take(A::ifft_shift(h_), "ifft_shift"); }
    // convolution
    GpuVec& convolve_signal(const GpuVec& impulse_response) { return take(A::convolve_signal(h_, impulse_response.h_), "convolve_signal"); }
    GpuVec& convolve(Response f, T rolloff, T ratio, size_t len) { return take(A::convolve(h_, (int)f, rolloff, ratio, len), "convolve"); }
    GpuVec& multiply_frequency_response(Response f, T rolloff, T ratio) { return take(A::mul_freq(h_, (int)f, rolloff, ratio), "multiply_frequency_response"); }
    // interpolation
    GpuVec& interpolatef(Response f, T rolloff, T factor, T delay, size_t conv_len) { return take(A::interpolatef(h_, (int)f, rolloff, factor, delay, conv_len), "interpolatef"); }
    GpuVec& interpolate_lin(T factor, T delay) { return take(A::interpolate_lin(h_, factor, delay), "interpolate_lin"); }
    // elementwise
    GpuVec& scale(T c) { return take(A::real_scale(h_, c), "scale"); }
    GpuVec& scale(std::complex<T> c) { return take(A::complex_scale(h_, c.real(), c.imag()), "scale"); }
    GpuVec& offset(T c) { return take(A::real_offset(h_, c), "offset"); }
    GpuVec& add(const GpuVec& o) { return take(A::add(h_, o.h_), "add"); }
    GpuVec& sub(const GpuVec& o) { return take(A::sub(h_, o.h_), "sub"); }
    GpuVec& mul(const GpuVec& o) { return take(A::mul(h_, o.h_), "mul"); }
    GpuVec& div(const GpuVec& o) { return take(A::div(h_, o.h_), "div"); }
    GpuVec& to_complex() { return take(A::to_complex(h_), "to_complex"); }
    // complex -> real
    GpuVec& magnitude() { return take(A::magnitude(h_), "magnitude"); }
    GpuVec& phase() { return take(A::phase(h_), "phase"); }
    void get_mag_phase(GpuVec& mag, GpuVec& ph) { A::get_mag_phase(h_, mag.h_, ph.h_); }
    // TimeDomainOperations / CrossCorrelation*Ops / InterpolationOps (FFT based) / Symmetric*DomainOperations
    GpuVec& apply_window(int window) { return take(A::apply_window(h_, window), "apply_window"); }
    GpuVec& unapply_window(int window) { return take(A::unapply_window(h_, window), "unapply_window"); }
    GpuVec& windowed_fft(int window) { return take(A::windowed_fft(h_, window), "windowed_fft"); }
    GpuVec& windowed_ifft(int window) { return take(A::windowed_ifft(h_, window), "windowed_ifft"); }
    GpuVec& prepare_argument() { return take(A::prepare_argument(h_), "prepare_argument"); }
    GpuVec& prepare_argument_padded() { return take(A::prepare_argument_padded(h_), "prepare_argument_padded"); }
    GpuVec& correlate(const GpuVec& prepared) { return take(A::correlate(h_, prepared.h_), "correlate"); }
    GpuVec& interpolatei(Response f, T rolloff, int factor) { return take(A::interpolatei(h_, (int)f, rolloff, factor), "interpolatei"); }
    GpuVec& interpolate(Response f, T rolloff, size_t dest_points, T delay) { return take(A::interpolate(h_, (int)f, rolloff, dest_points, delay), "interpolate"); }
    GpuVec& interpft(size_t dest_points) { return take(A::interpft(h_, dest_points), "interpft"); }
    GpuVec& decimatei(uint32_t factor, uint32_t delay) { return take(A::decimatei(h_, factor, delay), "decimatei"); }
    GpuVec& plain_sfft() { return take(A::plain_sfft(h_), "plain_sfft"); }
    GpuVec& plain_sifft() { return take(A::plain_sifft(h_), "plain_sifft"); }
    GpuVec& conj() { return take(A::conj(h_), "conj"); }
    T sum() const { return A::real_sum(h_); }
    // fused chain of the three sequential calls scale(c); mul(&w); get_mag_phase(..) in one pass
    void scale_mul_mag_phase(std::complex<T> c, const GpuVec& w, GpuVec& mag, GpuVec& ph, bool write_back = false) {
        if (int rc = A::scale_mul_mag_phase(h_, c.real(), c.imag(), w.h_, mag.h_, ph.h_, write_back)) throw DspError(rc, "scale_mul_mag_phase");
    }
};

typedef GpuVec<float> GpuVec32;
typedef GpuVec<double> GpuVec64;

}  // namespace basic_dsp_b200
