// Compiles against include/basic_dsp_b200.hpp and links libbasic_dsp_b200.so.  With a GPU it runs a
// small fft -> ifft round trip and a convolve_signal shift KAT (convolution.rs:818-830); without one it
// only proves that the C++ host mirror builds and links (exit code 0, prints "no device").
#include <cmath>
#include <cstdio>

#include "basic_dsp_b200.hpp"

using namespace basic_dsp_b200;

int main() {
    if (bdsp_device_count() < 1) { std::printf("no device\n"); return 0; }
    std::vector<std::complex<float>> x(1024);
    for (size_t i = 0; i < x.size(); i++) x[i] = {std::cos(0.1f * i), std::sin(0.37f * i)};
    GpuVec32 v(x);
    v.fft().ifft();
    std::vector<float> back = v.to_vec();
    double err = 0;
    for (size_t i = 0; i < x.size(); i++) err = std::fmax(err, std::fabs(back[2 * i] - x[i].real()) + std::fabs(back[2 * i + 1] - x[i].imag()));
    std::vector<std::complex<float>> a(10), b(10);
    for (int i = 0; i < 10; i++) a[i] = {(float)i, 0.f};
    b[4] = {1.f, 0.f};
    GpuVec32 av(a), bv(b);
    std::vector<float> c = av.convolve_signal(bv).magnitude().to_vec();
    double err2 = 0;
    for (int i = 0; i < 10; i++) err2 = std::fmax(err2, std::fabs(c[i] - (float)i));
    bool threw = false;
    try { GpuVec32 f(x, Domain::Frequency); f.fft(); } catch (const DspError& e) { threw = e.code == -1; }
    std::printf("roundtrip %.2e shift-kat %.2e wrong-domain-throws %d\n", err, err2, (int)threw);
    return (err < 1e-4 && err2 < 1e-4 && threw) ? 0 : 1;
}
