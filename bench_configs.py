#!/usr/bin/env python3
"""Per-config measurements for SURVEY.md section 8(d): every BASELINE config C1..C5 on ONE B200, device
resident, timed with CUDA events on the launching stream, reported as Msamples/s, GFLOP/s (5 N log2 N)
and fraction of the measured HBM roofline (algorithmic bytes from SURVEY.md 8(d)).

    python bench_configs.py [--iters 20] [--configs C1,C2a,...]   -> one JSON line per config

Inputs smaller than L2 (C1, C2a, C2b) are reported twice: L2-warm and with an L2 flush (a 256 MiB
memset) between iterations.  This script is a secondary report; the contract benchmark is bench.py."""
import argparse
import ctypes
import json
import math
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import basic_dsp_b200 as bd  # noqa: E402
from basic_dsp_b200 import DspVec  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


class Timer:
    def __init__(self, L):
        self.L = L
        self.flush_buf = L.bdsp_malloc(256 << 20)

    def run(self, fn, iters, setup=None, flush=False, warmup=3):
        import time
        L = self.L
        self.t_start = time.time()
        for _ in range(warmup):
            if setup:
                setup()
            fn()
        L.bdsp_sync()
        times = []
        for _ in range(iters):
            if setup:
                setup()
            if flush:
                L.bdsp_memset(self.flush_buf, 0, 256 << 20)
            e0, e1 = L.bdsp_event_create(), L.bdsp_event_create()
            L.bdsp_event_record(e0)
            fn()
            L.bdsp_event_record(e1)
            times.append(L.bdsp_event_elapsed_ms(e0, e1))
            L.bdsp_event_destroy(e0)
            L.bdsp_event_destroy(e1)
        return statistics.median(times), min(times)


RESULTS = []          # every report() line of this process, for callers that import the module (bench.py)
EMIT = True           # print one JSON line per config


def report(name, what, samples, alg_bytes, flops, med_ms, best_ms, extra=None):
    import time
    pk = peak()
    gbs = alg_bytes / (med_ms * 1e-3) / 1e9
    line = {"config": name, "what": what, "ms_median": med_ms, "ms_best": best_ms,
            "Msamples_per_s": samples / (med_ms * 1e-3) / 1e6, "GFLOPs_5nlog2n": flops / (med_ms * 1e-3) / 1e9 if flops else None,
            "algorithmic_bytes": alg_bytes, "achieved_GBps": gbs, "hbm_peak_GBps": pk, "roofline_frac": gbs / pk,
            "t_end": time.time()}
    if extra:
        line.update(extra)
    RESULTS.append(line)
    if EMIT:
        print(json.dumps(line), flush=True)


def dptr(v):
    return v._fn("bdsp_device_ptr")(v._h)


def rand_c(rng, n, dtype):
    ct = np.complex64 if dtype == np.float32 else np.complex128
    out = np.empty(n, dtype=ct)
    out.real = rng.uniform(-10, 10, n)
    out.imag = rng.uniform(-10, 10, n)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--configs", default="C1,C2a,C2b,C3,C4a,C4b,C5a,C5b")
    args = ap.parse_args()
    measure(set(args.configs.split(",")), args.iters)


def measure(want, iters, emit=True):
    """Runs the requested configs on the current device; returns the report lines (also kept in RESULTS)."""
    global EMIT
    EMIT = emit
    del RESULTS[:]

    class _A:
        pass
    args = _A()
    args.iters = iters
    L = bd.lib()
    bd.require_device()
    T = Timer(L)
    rng = np.random.default_rng(20260101)

    if "C1" in want:
        n = 1 << 16
        x = rand_c(rng, n, np.float32)
        v = DspVec(x)
        for flush in (False, True):
            med, best = T.run(lambda: (v.fft(), v.ifft()), args.iters, flush=flush)
            report("C1", "c32 2^16 fft -> ifft round trip" + (" (L2 flushed)" if flush else " (L2 warm)"), 2 * n, 32 * n,
                   2 * 5 * n * 16, med, best)

    if "C2a" in want or "C2b" in want:
        n = 1 << 20
        x = rand_c(rng, n, np.float32)
        k = np.arange(1023, dtype=np.float32)
        from bench import make_taps
        hv = DspVec(make_taps())
        src = DspVec(x)
        v = DspVec(x)
        nb = n * 8

        def reset():
            L.bdsp_memcpy_h2d  # noqa: B018  (device-to-device copy below keeps the input identical every iteration)
            ctypes.memmove  # noqa: B018
        if "C2a" in want:
            for flush in (False, True):
                med, best = T.run(lambda: v.convolve_signal(hv), args.iters, flush=flush)
                report("C2a", "c32 2^20 convolve_signal, 1023-tap RC" + (" (L2 flushed)" if flush else " (L2 warm)"), n, 16 * n,
                       2 * 5 * n * 20 + 6 * n, med, best)
        if "C2b" in want:
            for flush in (False, True):
                med, best = T.run(lambda: v.convolve(bd.RAISED_COSINE, 0.35, 0.25, 31), args.iters, flush=flush)
                report("C2b", "c32 2^20 convolve, 63-tap RC (len=31, ratio=0.25)" + (" (L2 flushed)" if flush else " (L2 warm)"), n,
                       16 * n, 4 * 63 * n, med, best, {"note": "tap table cached on the device after the first call"})
        del src

    if "C3" in want:
        n, rows = 1 << 14, 4096
        x = rand_c(rng, n * rows, np.float32)
        vin = DspVec(x)
        out = DspVec.zeros(n * rows, dtype=np.float32)
        pin, pout = dptr(vin), dptr(out)
        med, best = T.run(lambda: L.bdsp_fft_rows_c32(pin, pout, n, rows, bd.F_SHIFT | bd.F_MAGNITUDE), args.iters)
        report("C3", "4096 x 2^14 c32 fft + magnitude (one launch)", n * rows, 12 * n * rows, 5 * n * 14 * rows, med, best)
        del vin, out

    if "X1" in want:
        # extra: plain 2^20-point c32 FFT, 64 rows (two-pass tile kernels), and one 2^24-point transform
        n, rows = 1 << 20, 64
        x = rand_c(rng, n * rows, np.float32)
        vin = DspVec(x)
        out = DspVec.zeros(2 * n * rows, is_complex=True, dtype=np.float32)
        pin, pout = dptr(vin), dptr(out)
        med, best = T.run(lambda: L.bdsp_fft_rows_c32(pin, pout, n, rows, 0), args.iters)
        report("X1", "64 x 2^20 c32 plain fft (2-pass)", n * rows, 16 * n * rows, 5 * n * 20 * rows, med, best)
        med, best = T.run(lambda: L.bdsp_fft_rows_c32(pin, pout, n * 16, rows // 16, 0), args.iters)
        report("X1", "4 x 2^24 c32 plain fft (3-pass)", n * rows, 16 * n * rows, 5 * n * 24 * rows, med, best)
        del vin, out


    if "R1" in want:
        # extra: reductions and elementwise math over 2^27 f32 scalars (512 MiB, > L2)
        n = 1 << 27
        x = rng.uniform(-10, 10, n).astype(np.float32)
        v = DspVec(x)
        w = DspVec(x[::-1].copy())
        med, best = T.run(lambda: v.sum(), args.iters)
        report("R1", "real f32 2^27 sum (incl. result read-back)", n, 4 * n, 0, med, best)
        med, best = T.run(lambda: v.statistics(), args.iters)
        report("R1", "real f32 2^27 statistics (sum, rms, min/max + indices)", n, 4 * n, 0, med, best)
        med, best = T.run(lambda: v.dot_product(w), args.iters)
        report("R1", "real f32 2^27 dot_product", n, 8 * n, 0, med, best)
        c = DspVec(x.view(np.complex64))
        med, best = T.run(lambda: c.statistics(), args.iters)
        report("R1", "c32 2^26 statistics", n // 2, 4 * n, 0, med, best)
        del w
        med, best = T.run(lambda: v.math("abs"), args.iters)
        report("R1", "real f32 2^27 abs (in place)", n, 8 * n, 0, med, best)
        med, best = T.run(lambda: v.math("sin"), args.iters)
        report("R1", "real f32 2^27 sin (in place)", n, 8 * n, 0, med, best)
        med, best = T.run(lambda: c.math("sqrt"), args.iters)
        report("R1", "c32 2^26 complex sqrt (in place)", n // 2, 8 * n, 0, med, best)
        med, best = T.run(lambda: v.math("cum_sum"), args.iters)
        report("R1", "real f32 2^27 cum_sum (3 launches)", n, 16 * n, 0, med, best)
        del v, c

    if "C4a" in want or "C4b" in want:
        n = 1 << 24
        x = rng.uniform(-10, 10, n).astype(np.float32)
        if "C4a" in want:
            v = DspVec(x)

            def setup():
                v.set_len(n)
            med, best = T.run(lambda: v.interpolatef(bd.SINC, 0.0, 4.0, 0.0, 12), args.iters, setup=setup)
            report("C4a", "real f32 2^24 interpolatef x4 sinc conv_len=12", n, 20 * n, 2 * 25 * 4 * n, med, best,
                   {"note": "tap table cached on the device after the first call; input length reset by set_len (data differs, same cost)"})
            del v
        if "C4b" in want:
            v = DspVec(x)

            def setup2():
                v.set_len(n)
            med, best = T.run(lambda: v.interpolate_lin(4.0, 0.0), args.iters, setup=setup2)
            report("C4b", "real f32 2^24 interpolate_lin x4", n, 20 * n, 3 * 4 * n, med, best)
            del v

    if "C5a" in want or "C5b" in want:
        n = 3 * (1 << 26)
        v = DspVec.zeros(2 * n, is_complex=True, dtype=np.float64, init=0.0)
        chunk = 1 << 24
        buf = rand_c(rng, chunk, np.float64).view(np.float64)
        base = dptr(v)
        for off in range(0, 2 * n, 2 * chunk):
            L.bdsp_memcpy_h2d(base + off * 8, buf.ctypes.data, min(2 * chunk, 2 * n - off) * 8)
        L.bdsp_sync()
        if "C5a" in want:
            state = {"d": 0}

            def fwd_inv():
                # alternate fft / ifft so that the vector stays in a valid domain; both are one transform
                if state["d"] == 0:
                    v.plain_fft()
                else:
                    v.plain_ifft()
                state["d"] ^= 1
            med, best = T.run(fwd_inv, max(4, args.iters // 4), warmup=2)
            report("C5a", "c64 3*2^26 mixed-radix plain_fft / plain_ifft (alternating)", n, 32 * n, 5 * n * math.log2(n), med, best)
        if "C5b" in want:
            if v.domain() != bd.FREQ:
                v.plain_fft()
            w = v.clone()
            mag, ph = DspVec.zeros(n, dtype=np.float64), DspVec.zeros(n, dtype=np.float64)
            med, best = T.run(lambda: v.scale_mul_mag_phase(complex(0.5, 0.25), w, mag, ph), max(4, args.iters // 4), warmup=2)
            report("C5b", "c64 3*2^26 fused scale -> mul -> (magnitude, phase)", n, 48 * n, None, med, best)
    L.bdsp_free(T.flush_buf)
    return list(RESULTS)


if __name__ == "__main__":
    main()
