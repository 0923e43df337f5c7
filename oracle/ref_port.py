"""ctypes loader for oracle/ref_port.c (C port of the reference's CPU convolve_signal path).
TEST / BASELINE INFRASTRUCTURE ONLY - never imported by basic_dsp_b200."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libref_port.so")
_lib = None


def _host_tag():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return str(hash(line) & 0xFFFFFFFF)
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """(Re)builds the port with -march=native; a tag file makes a box with another CPU rebuild it."""
    src = os.path.join(_HERE, "ref_port.c")
    tag_file = os.path.join(_HERE, "_build", "host_tag")
    tag = _host_tag()
    stale = (not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)
             or not os.path.exists(tag_file) or open(tag_file).read() != tag)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
        with open(tag_file, "w") as fh:
            fh.write(tag)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        P = ctypes.c_void_p
        _lib.ref_convolve_signal_c32.restype = ctypes.c_int
        _lib.ref_convolve_signal_c32.argtypes = [P, ctypes.c_size_t, P, ctypes.c_size_t, P]
        _lib.ref_convolve_signal_rows_c32.restype = ctypes.c_int
        _lib.ref_convolve_signal_rows_c32.argtypes = [P, ctypes.c_size_t, ctypes.c_size_t, P, ctypes.c_size_t, ctypes.c_int]
        _lib.ref_convolve_signal_scalar_c32.restype = None
        _lib.ref_convolve_signal_scalar_c32.argtypes = [P, ctypes.c_size_t, P, ctypes.c_size_t, P]
        _lib.ref_overlap_discard_c32.restype = ctypes.c_int
        _lib.ref_overlap_discard_c32.argtypes = [P, ctypes.c_size_t, P, ctypes.c_size_t]
        _lib.ref_fft_rows_c32.restype = ctypes.c_int
        _lib.ref_fft_rows_c32.argtypes = [P, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    return _lib


def convolve_signal_rows(x, h, threads=1):
    """x: (rows, n) complex64, h: (l,) complex64 -> centred circular convolution of every row,
    computed the way the reference's CPU path does (overlap_discard incl. its scalar head/tail)."""
    x = np.ascontiguousarray(x, dtype=np.complex64).copy()
    h = np.ascontiguousarray(h, dtype=np.complex64)
    rows, n = (1, x.shape[0]) if x.ndim == 1 else x.shape
    rc = lib().ref_convolve_signal_rows_c32(x.ctypes.data, n, rows, h.ctypes.data, len(h), threads)
    if rc:
        raise RuntimeError("ref_convolve_signal_rows_c32 -> %d" % rc)
    return x


def convolve_signal_rows_inplace(x, h, threads=1):
    rows, n = x.shape
    return lib().ref_convolve_signal_rows_c32(x.ctypes.data, n, rows, h.ctypes.data, len(h), threads)


def fft_rows(x, shift=True, threads=1):
    x = np.ascontiguousarray(x, dtype=np.complex64).copy()
    rows, n = x.shape
    rc = lib().ref_fft_rows_c32(x.ctypes.data, n, rows, 1 if shift else 0, threads)
    if rc:
        raise RuntimeError("ref_fft_rows_c32 -> %d" % rc)
    return x
