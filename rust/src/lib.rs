//! `GpuVec32` / `GpuVec64`: vectors that live in B200 HBM behind basic_dsp's trait surface.
//!
//! NOT compiled in this repository's image (no Rust toolchain); written against basic_dsp_vector 0.10.0.
//! The reference's operations are blanket impls over every storage `S` of `DspVec<S, T, N, D>` that run CPU code on
//! `ToSlice::to_slice` (e.g. `time_to_freq.rs:126-135`, `requirements.rs:6-24`), and stable Rust cannot specialise
//! them, so the drop-in is a NEW vector type implementing the SAME traits.  Every method is one
//! call into the C ABI (`ffi.rs`, generated from `include/basic_dsp_b200.h`); data stays on the device between calls.
//!
//! * Buffers: the traits take a caller-owned `B: Buffer<S, T>`; the device keeps its own never-shrinking scratch per
//!   handle (`trade` = pointer swap, `support_std.rs:79-81`), so the argument is accepted and ignored.  `GpuStorage` is
//!   the formal `S` (it implements `ToSlice`/`ToSliceMut` through a lazily synchronised host mirror: `data32()`).
//! * Errors: result codes 1..=14 are `ErrorReason` (`interop/src/lib.rs:125-142`), -1 is the error marker of a run-time
//!   typed vector; statically typed `GpuVec`s make most of them unrepresentable, exactly as in the reference.
//! * Type changes consume `self` and return the re-typed vector around the same handle (`RededicateForceOps`,
//!   `rededicate_and_relations.rs:44-55`).
#![allow(clippy::missing_safety_doc)]

pub mod ffi;

use basic_dsp_vector::conv_types::*;
use basic_dsp_vector::meta;
use basic_dsp_vector::numbers::*;
use basic_dsp_vector::window_functions::*;
use basic_dsp_vector::*;
use num_complex::Complex;
use std::marker::PhantomData;
use std::os::raw::c_void;

/// Formal storage type (`S` of the reference's traits).  `to_slice` downloads into the handle's host mirror.
pub struct GpuStorage<T> {
    mirror: Vec<T>,
}
impl<T: RealNumber> ToSlice<T> for GpuStorage<T> {
    fn to_slice(&self) -> &[T] { &self.mirror }
    fn len(&self) -> usize { self.mirror.len() }
    fn is_empty(&self) -> bool { self.mirror.is_empty() }
    fn alloc_len(&self) -> usize { self.mirror.capacity() }
    fn try_resize(&mut self, len: usize) -> VoidResult { self.mirror.resize(len, T::zero()); Ok(()) }
}
impl<T: RealNumber> ToSliceMut<T> for GpuStorage<T> {
    fn to_slice_mut(&mut self) -> &mut [T] { &mut self.mirror }
}

fn reason(code: i32) -> ErrorReason {
    match code {   // interop/src/lib.rs:125-142
        1 => ErrorReason::InputMustHaveTheSameSize,
        2 => ErrorReason::InputMetaDataMustAgree,
        3 => ErrorReason::InputMustBeComplex,
        4 => ErrorReason::InputMustBeReal,
        5 => ErrorReason::InputMustBeInTimeDomain,
        6 => ErrorReason::InputMustBeInFrequencyDomain,
        7 => ErrorReason::InvalidArgumentLength,
        8 => ErrorReason::InputMustBeConjSymmetric,
        9 => ErrorReason::InputMustHaveAnOddLength,
        10 => ErrorReason::ArgumentFunctionMustBeSymmetric,
        11 => ErrorReason::InvalidNumberOfArgumentsForCombinedOp,
        12 => ErrorReason::InputMustNotBeEmpty,
        13 => ErrorReason::InputMustHaveAnEvenLength,
        _ => ErrorReason::TypeCanNotResize,
    }
}

// trampolines: `&dyn RealImpulseResponse<T>` etc. as the C ABI's (callback, user data) pairs (interop/src/lib.rs:279-377)
unsafe extern "C" fn real_ir32(data: *const c_void, x: f32) -> f32 { (*(data as *const &dyn RealImpulseResponse<f32>)).calc(x) }
unsafe extern "C" fn real_fr32(data: *const c_void, x: f32) -> f32 { (*(data as *const &dyn RealFrequencyResponse<f32>)).calc(x) }
unsafe extern "C" fn window32(data: *const c_void, i: usize, n: usize) -> f32 { (*(data as *const &dyn WindowFunction<f32>)).window(i, n) }
unsafe extern "C" fn real_ir64(data: *const c_void, x: f64) -> f64 { (*(data as *const &dyn RealImpulseResponse<f64>)).calc(x) }
unsafe extern "C" fn real_fr64(data: *const c_void, x: f64) -> f64 { (*(data as *const &dyn RealFrequencyResponse<f64>)).calc(x) }
unsafe extern "C" fn window64(data: *const c_void, i: usize, n: usize) -> f64 { (*(data as *const &dyn WindowFunction<f64>)).window(i, n) }

macro_rules! gpu_vec {
    ($Vec:ident, $T:ty, $Handle:ty, $Res:ty, $Cplx:ty, $real_ir:ident, $real_fr:ident, $window:ident,
     $new:ident, $delete:ident, $clone:ident, $upload:ident, $download:ident, $get_len:ident, $get_points:ident, $get_delta:ident,
     $plain_fft:ident, $fft:ident, $plain_ifft:ident, $ifft:ident, $windowed_fft:ident, $windowed_ifft:ident,
     $mirror:ident, $fft_shift:ident, $ifft_shift:ident,
     $convolve_signal:ident, $convolve_real:ident, $mfr_real:ident,
     $interpolatef:ident, $interpolatei:ident, $interpolate:ident, $interpft:ident, $decimatei:ident,
     $interpolate_lin:ident, $interpolate_hermite:ident,
     $real_scale:ident, $complex_scale:ident, $real_offset:ident, $complex_offset:ident,
     $add:ident, $sub:ident, $mul:ident, $div:ident,
     $magnitude:ident, $magnitude_squared:ident, $to_real:ident, $to_imag:ident, $phase:ident,
     $prepare_argument:ident, $prepare_argument_padded:ident, $correlate:ident) => {
        /// Device-resident vector; `N` / `D` are the reference's type-level number space and domain (`meta.rs`).
        pub struct $Vec<N: NumberSpace, D: Domain> {
            h: *mut $Handle,
            _m: PhantomData<(N, D)>,
        }
        unsafe impl<N: NumberSpace, D: Domain> Send for $Vec<N, D> {}   // one vector => one thread at a time (&mut), as in the reference

        impl<N: NumberSpace, D: Domain> Drop for $Vec<N, D> {
            fn drop(&mut self) { if !self.h.is_null() { unsafe { ffi::$delete(self.h) } } }
        }
        impl<N: NumberSpace, D: Domain> Clone for $Vec<N, D> {
            fn clone(&self) -> Self { $Vec { h: unsafe { ffi::$clone(self.h) }, _m: PhantomData } }
        }

        impl<N: NumberSpace, D: Domain> $Vec<N, D> {
            /// Uploads interleaved host data (`[re0, im0, re1, ...]` for complex vectors, `support_std.rs:366-372`).
            pub fn from_slice(data: &[$T], is_complex: bool, domain: DataDomain, delta: $T) -> Option<Self> {
                let dom = if domain == DataDomain::Time { 0 } else { 1 };
                let h = unsafe { ffi::$new(is_complex as i32, dom, 0.0, data.len(), delta) };
                if h.is_null() { return None; }   // allocation failure: bdsp_last_error() says why
                unsafe { ffi::$upload(h, data.as_ptr(), data.len()) };
                Some($Vec { h, _m: PhantomData })
            }
            /// The one unavoidable difference to a host vector: reading the data is an explicit (synchronising) download.
            pub fn to_vec(&self) -> Vec<$T> {
                let n = self.len();
                let mut out = vec![0.0; n];
                unsafe { ffi::$download(self.h, out.as_mut_ptr(), n) };
                out
            }
            pub fn len(&self) -> usize { unsafe { ffi::$get_len(self.h) } }
            pub fn is_empty(&self) -> bool { self.len() == 0 }
            pub fn points(&self) -> usize { unsafe { ffi::$get_points(self.h) } }
            pub fn delta(&self) -> $T { unsafe { ffi::$get_delta(self.h) } }

            fn take(mut self) -> *mut $Handle { std::mem::replace(&mut self.h, std::ptr::null_mut()) }
            /// The C ABI takes the handle by value and hands it back (`VectorInteropResult`, interop/src/lib.rs:203-212).
            fn inplace(&mut self, f: impl FnOnce(*mut $Handle) -> $Res) -> VoidResult {
                let r = f(self.h);
                self.h = r.vector;
                if r.result_code == 0 { Ok(()) } else { Err(reason(r.result_code)) }
            }
            fn rededicate<N2: NumberSpace, D2: Domain>(self, f: impl FnOnce(*mut $Handle) -> $Res) -> $Vec<N2, D2> {
                let r = f(self.take());
                $Vec { h: r.vector, _m: PhantomData }
            }
        }

        // ---- TimeToFrequencyDomainOperations (time_to_freq.rs:14-71) ----------------------------------------------
        impl<N: NumberSpace> ToFreqResult for $Vec<N, meta::Time> { type FreqResult = $Vec<meta::Complex, meta::Freq>; }
        impl<N: NumberSpace> TimeToFrequencyDomainOperations<GpuStorage<$T>, $T> for $Vec<N, meta::Time> {
            fn plain_fft<B>(self, _: &mut B) -> Self::FreqResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                self.rededicate(|h| unsafe { ffi::$plain_fft(h) })
            }
            fn fft<B>(self, _: &mut B) -> Self::FreqResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                self.rededicate(|h| unsafe { ffi::$fft(h) })
            }
            fn windowed_fft<B>(self, _: &mut B, window: &dyn WindowFunction<$T>) -> Self::FreqResult
            where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let w = &window as *const &dyn WindowFunction<$T> as *const c_void;
                self.rededicate(|h| unsafe { ffi::$windowed_fft(h, Some($window), w, window.is_symmetric() as u8) })
            }
        }
        // ---- FrequencyToTimeDomainOperations (freq_to_time.rs:16-73) ----------------------------------------------
        impl ToTimeResult for $Vec<meta::Complex, meta::Freq> { type TimeResult = $Vec<meta::Complex, meta::Time>; }
        impl FrequencyToTimeDomainOperations<GpuStorage<$T>, $T> for $Vec<meta::Complex, meta::Freq> {
            fn plain_ifft<B>(self, _: &mut B) -> Self::TimeResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                self.rededicate(|h| unsafe { ffi::$plain_ifft(h) })
            }
            fn ifft<B>(self, _: &mut B) -> Self::TimeResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                self.rededicate(|h| unsafe { ffi::$ifft(h) })
            }
            fn windowed_ifft<B>(self, _: &mut B, window: &dyn WindowFunction<$T>) -> Self::TimeResult
            where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let w = &window as *const &dyn WindowFunction<$T> as *const c_void;
                self.rededicate(|h| unsafe { ffi::$windowed_ifft(h, Some($window), w, window.is_symmetric() as u8) })
            }
        }
        // ---- FrequencyDomainOperations (freq.rs:11-41) -------------------------------------------------------------
        impl FrequencyDomainOperations<GpuStorage<$T>, $T> for $Vec<meta::Complex, meta::Freq> {
            fn mirror<B>(&mut self, _: &mut B) where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> { let _ = self.inplace(|h| unsafe { ffi::$mirror(h) }); }
            fn fft_shift(&mut self) { let _ = self.inplace(|h| unsafe { ffi::$fft_shift(h) }); }
            fn ifft_shift(&mut self) { let _ = self.inplace(|h| unsafe { ffi::$ifft_shift(h) }); }
        }
        // ---- ConvolutionOps / Convolution / FrequencyMultiplication (convolution.rs:17-84) ---------------------------
        impl<N: NumberSpace> ConvolutionOps<$Vec<N, meta::Time>, GpuStorage<$T>, $T, N, meta::Time> for $Vec<N, meta::Time> {
            fn convolve_signal<B>(&mut self, _: &mut B, impulse_response: &$Vec<N, meta::Time>) -> VoidResult
            where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let ir = impulse_response.h as *const $Handle;   // borrowed, read-only: its spectrum cache is thread safe
                self.inplace(|h| unsafe { ffi::$convolve_signal(h, ir) })
            }
        }
        impl<'a, N: NumberSpace> Convolution<'a, GpuStorage<$T>, $T, &'a dyn RealImpulseResponse<$T>> for $Vec<N, meta::Time> {
            fn convolve<B>(&mut self, _: &mut B, impulse_response: &'a dyn RealImpulseResponse<$T>, ratio: $T, len: usize)
            where B: for<'b> Buffer<'b, GpuStorage<$T>, $T> {
                let d = &impulse_response as *const &dyn RealImpulseResponse<$T> as *const c_void;
                let _ = self.inplace(|h| unsafe { ffi::$convolve_real(h, Some($real_ir), d, impulse_response.is_symmetric() as u8, ratio, len) });
            }
        }
        impl<'a, N: NumberSpace> FrequencyMultiplication<'a, GpuStorage<$T>, $T, &'a dyn RealFrequencyResponse<$T>> for $Vec<N, meta::Freq> {
            fn multiply_frequency_response(&mut self, frequency_response: &'a dyn RealFrequencyResponse<$T>, ratio: $T) {
                let d = &frequency_response as *const &dyn RealFrequencyResponse<$T> as *const c_void;
                let _ = self.inplace(|h| unsafe { ffi::$mfr_real(h, Some($real_fr), d, frequency_response.is_symmetric() as u8, ratio) });
            }
        }
        // ---- InterpolationOps / RealInterpolationOps (interpolation.rs:18-90, real_interpolation.rs:10-24) ----------
        impl<N: NumberSpace> InterpolationOps<GpuStorage<$T>, $T> for $Vec<N, meta::Time> {
            fn interpolatef<B>(&mut self, _: &mut B, function: &dyn RealImpulseResponse<$T>, interpolation_factor: $T, delay: $T, conv_len: usize)
            where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let d = &function as *const &dyn RealImpulseResponse<$T> as *const c_void;
                let _ = self.inplace(|h| unsafe { ffi::$interpolatef(h, Some($real_ir), d, function.is_symmetric() as u8, interpolation_factor, delay, conv_len) });
            }
            fn interpolatei<B>(&mut self, _: &mut B, function: &dyn RealFrequencyResponse<$T>, interpolation_factor: u32) -> VoidResult
            where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let d = &function as *const &dyn RealFrequencyResponse<$T> as *const c_void;
                self.inplace(|h| unsafe { ffi::$interpolatei(h, Some($real_fr), d, function.is_symmetric() as u8, interpolation_factor as i32) })
            }
            fn interpolate<B>(&mut self, _: &mut B, function: Option<&dyn RealFrequencyResponse<$T>>, target_points: usize, delay: $T) -> VoidResult
            where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                match function {
                    Some(f) => {
                        let d = &f as *const &dyn RealFrequencyResponse<$T> as *const c_void;
                        self.inplace(|h| unsafe { ffi::$interpolate(h, Some($real_fr), d, f.is_symmetric() as u8, target_points, delay) })
                    }
                    None => self.inplace(|h| unsafe { ffi::$interpft(h, target_points) }),
                }
            }
            fn interpft<B>(&mut self, _: &mut B, target_points: usize) where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let _ = self.inplace(|h| unsafe { ffi::$interpft(h, target_points) });
            }
            fn decimatei(&mut self, decimation_factor: u32, delay: u32) {
                let _ = self.inplace(|h| unsafe { ffi::$decimatei(h, decimation_factor, delay) });
            }
        }
        impl RealInterpolationOps<GpuStorage<$T>, $T> for $Vec<meta::Real, meta::Time> {
            fn interpolate_hermite<B>(&mut self, _: &mut B, interpolation_factor: $T, delay: $T) where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let _ = self.inplace(|h| unsafe { ffi::$interpolate_hermite(h, interpolation_factor, delay) });
            }
            fn interpolate_lin<B>(&mut self, _: &mut B, interpolation_factor: $T, delay: $T) where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let _ = self.inplace(|h| unsafe { ffi::$interpolate_lin(h, interpolation_factor, delay) });
            }
        }
        // ---- ScaleOps / OffsetOps / ElementaryOps (elementary.rs:13-162) ------------------------------------------------
        impl<N: NumberSpace, D: Domain> ScaleOps<$T> for $Vec<N, D> {
            fn scale(&mut self, factor: $T) { let _ = self.inplace(|h| unsafe { ffi::$real_scale(h, factor) }); }
        }
        impl<D: Domain> ScaleOps<Complex<$T>> for $Vec<meta::Complex, D> {
            fn scale(&mut self, factor: Complex<$T>) { let _ = self.inplace(|h| unsafe { ffi::$complex_scale(h, factor.re, factor.im) }); }
        }
        impl<D: Domain> OffsetOps<$T> for $Vec<meta::Real, D> {
            fn offset(&mut self, offset: $T) { let _ = self.inplace(|h| unsafe { ffi::$real_offset(h, offset) }); }
        }
        impl<D: Domain> OffsetOps<Complex<$T>> for $Vec<meta::Complex, D> {
            fn offset(&mut self, offset: Complex<$T>) { let _ = self.inplace(|h| unsafe { ffi::$complex_offset(h, offset.re, offset.im) }); }
        }
        impl<N: NumberSpace, D: Domain> ElementaryOps<$Vec<N, D>, $T, N, D> for $Vec<N, D>
        where $Vec<N, D>: GetMetaData<$T, N, D> {
            fn add(&mut self, summand: &$Vec<N, D>) -> VoidResult { let o = summand.h as *const $Handle; self.inplace(|h| unsafe { ffi::$add(h, o) }) }
            fn sub(&mut self, subtrahend: &$Vec<N, D>) -> VoidResult { let o = subtrahend.h as *const $Handle; self.inplace(|h| unsafe { ffi::$sub(h, o) }) }
            fn mul(&mut self, factor: &$Vec<N, D>) -> VoidResult { let o = factor.h as *const $Handle; self.inplace(|h| unsafe { ffi::$mul(h, o) }) }
            fn div(&mut self, divisor: &$Vec<N, D>) -> VoidResult { let o = divisor.h as *const $Handle; self.inplace(|h| unsafe { ffi::$div(h, o) }) }
        }
        // ---- ComplexToRealTransformsOps (complex_to_real.rs:17-330) -----------------------------------------------------
        impl<D: Domain> ToRealResult for $Vec<meta::Complex, D> { type RealResult = $Vec<meta::Real, D>; }
        impl<D: Domain> ComplexToRealTransformsOps<$T> for $Vec<meta::Complex, D> {
            fn magnitude(self) -> Self::RealResult { self.rededicate(|h| unsafe { ffi::$magnitude(h) }) }
            fn magnitude_squared(self) -> Self::RealResult { self.rededicate(|h| unsafe { ffi::$magnitude_squared(h) }) }
            fn to_real(self) -> Self::RealResult { self.rededicate(|h| unsafe { ffi::$to_real(h) }) }
            fn to_imag(self) -> Self::RealResult { self.rededicate(|h| unsafe { ffi::$to_imag(h) }) }
            fn phase(self) -> Self::RealResult { self.rededicate(|h| unsafe { ffi::$phase(h) }) }
        }
        // ---- CrossCorrelationArgumentOps / CrossCorrelationOps, the "preparation" API (correlation.rs:12-84) -----------
        impl CrossCorrelationArgumentOps<GpuStorage<$T>, $T> for $Vec<meta::Complex, meta::Time> {
            fn prepare_argument<B>(self, _: &mut B) -> Self::FreqResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                self.rededicate(|h| unsafe { ffi::$prepare_argument(h) })
            }
            fn prepare_argument_padded<B>(self, _: &mut B) -> Self::FreqResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                self.rededicate(|h| unsafe { ffi::$prepare_argument_padded(h) })
            }
        }
        impl CrossCorrelationOps<$Vec<meta::Complex, meta::Freq>, GpuStorage<$T>, $T, meta::Complex, meta::Freq> for $Vec<meta::Complex, meta::Time>
        where $Vec<meta::Complex, meta::Freq>: GetMetaData<$T, meta::Complex, meta::Freq> {
            fn correlate<B>(&mut self, _: &mut B, other: &$Vec<meta::Complex, meta::Freq>) -> VoidResult where B: for<'a> Buffer<'a, GpuStorage<$T>, $T> {
                let o = other.h as *const $Handle;
                self.inplace(|h| unsafe { ffi::$correlate(h, o) })
            }
        }
    };
}

gpu_vec!(GpuVec32, f32, ffi::BdspVec32, ffi::BdspVecResult32, ffi::BdspComplex32, real_ir32, real_fr32, window32,
         new32, delete_vector32, clone32, bdsp_upload32, bdsp_download32, get_len32, get_points32, get_delta32,
         plain_fft32, fft32, plain_ifft32, ifft32, windowed_custom_fft32, windowed_custom_ifft32,
         mirror32, fft_shift32, ifft_shift32,
         convolve_signal32, convolve_real32, multiply_frequency_response_real32,
         interpolatef_custom32, interpolatei_custom32, interpolate_custom32, interpft32, decimatei32,
         interpolate_lin32, interpolate_hermite32,
         real_scale32, complex_scale32, real_offset32, complex_offset32,
         add32, sub32, mul32, div32,
         magnitude32, magnitude_squared32, to_real32, to_imag32, phase32,
         prepare_argument32, prepare_argument_padded32, correlate32);
gpu_vec!(GpuVec64, f64, ffi::BdspVec64, ffi::BdspVecResult64, ffi::BdspComplex64, real_ir64, real_fr64, window64,
         new64, delete_vector64, clone64, bdsp_upload64, bdsp_download64, get_len64, get_points64, get_delta64,
         plain_fft64, fft64, plain_ifft64, ifft64, windowed_custom_fft64, windowed_custom_ifft64,
         mirror64, fft_shift64, ifft_shift64,
         convolve_signal64, convolve_real64, multiply_frequency_response_real64,
         interpolatef_custom64, interpolatei_custom64, interpolate_custom64, interpft64, decimatei64,
         interpolate_lin64, interpolate_hermite64,
         real_scale64, complex_scale64, real_offset64, complex_offset64,
         add64, sub64, mul64, div64,
         magnitude64, magnitude_squared64, to_real64, to_imag64, phase64,
         prepare_argument64, prepare_argument_padded64, correlate64);

pub type ComplexTimeGpuVec32 = GpuVec32<meta::Complex, meta::Time>;
pub type ComplexFreqGpuVec32 = GpuVec32<meta::Complex, meta::Freq>;
pub type RealTimeGpuVec32 = GpuVec32<meta::Real, meta::Time>;
pub type ComplexTimeGpuVec64 = GpuVec64<meta::Complex, meta::Time>;
pub type ComplexFreqGpuVec64 = GpuVec64<meta::Complex, meta::Freq>;
pub type RealTimeGpuVec64 = GpuVec64<meta::Real, meta::Time>;
