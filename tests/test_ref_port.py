"""The C port of the reference's CPU path (oracle/ref_port.c) against the NumPy oracle.  CPU only."""
import numpy as np
import pytest

from oracle import dsp_oracle as o
from oracle import ref_port


@pytest.mark.parametrize("n,l", [(100, 6), (6000, 20), (9500, 100), (20000, 1023), (16384, 300), (12289, 513)])
def test_port_matches_oracle(n, l):
    rng = np.random.default_rng(n + l)
    x = (rng.uniform(-10, 10, n) + 1j * rng.uniform(-10, 10, n)).astype(np.complex64)
    h = (rng.uniform(-1, 1, l) + 1j * rng.uniform(-1, 1, l)).astype(np.complex64)
    got = ref_port.convolve_signal_rows(x, h)[0] if False else ref_port.convolve_signal_rows(x[None, :], h)[0]
    assert o.rel_l2(got, o.convolve_signal(x, h)) < 2e-5 * np.log2(4096)


def test_port_rows_threads():
    rng = np.random.default_rng(1)
    x = (rng.uniform(-10, 10, (4, 12000)) + 1j * rng.uniform(-10, 10, (4, 12000))).astype(np.complex64)
    h = (rng.uniform(-1, 1, 200) + 0j).astype(np.complex64)
    a = ref_port.convolve_signal_rows(x, h, threads=1)
    b = ref_port.convolve_signal_rows(x, h, threads=4)
    assert np.array_equal(a, b)


def test_port_fft_rows():
    rng = np.random.default_rng(2)
    x = (rng.uniform(-10, 10, (3, 1024)) + 1j * rng.uniform(-10, 10, (3, 1024))).astype(np.complex64)
    got = ref_port.fft_rows(x, shift=True)
    ref = np.fft.fftshift(np.fft.fft(x.astype(np.complex128), axis=1), axes=1)
    assert o.rel_l2(got, ref) < 1e-5 * 10
