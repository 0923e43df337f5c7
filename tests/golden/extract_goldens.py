#!/usr/bin/env python3
"""Extract the known-answer vectors the reference's own tests hold for the hot path.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/extract_goldens.py

It parses the `expected` arrays out of the named `#[test]` functions / doctests of the reference
source (read-only) and writes them to tests/golden/reference_kats.json together with the file:line
each one came from.  The *inputs* of each case are re-stated in tests/test_oracle_golden.py (they
are code, not data, in the reference).  Nothing here is executed from the reference; it is a text
extraction of numeric literals.
"""
import json
import os
import re
import sys

REF = os.environ.get("BASIC_DSP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")

# (key, file, test fn name, name of the array variable)
CASES = [
    ("fft_vector64", "tests/time_freq_test.rs", "fft_vector64", "expected"),
    ("windowed_fft_vector64", "tests/time_freq_test.rs", "windowed_fft_vector64", "expected"),
    ("convolve_complex_freq_and_freq32", "vector/src/vector_types/time_freq/convolution.rs", "convolve_complex_freq_and_freq32", "expected"),
    ("convolve_complex_freq_and_freq_even32", "vector/src/vector_types/time_freq/convolution.rs", "convolve_complex_freq_and_freq_even32", "expected"),
    ("convolve_real_time_and_time32", "vector/src/vector_types/time_freq/convolution.rs", "convolve_real_time_and_time32", "expected"),
    ("convolve_complex_time_and_time32", "vector/src/vector_types/time_freq/convolution.rs", "convolve_complex_time_and_time32", "expected"),
    ("convolve_complex_vectors32", "vector/src/vector_types/time_freq/convolution.rs", "convolve_complex_vectors32", "expected"),
    ("shift_left_by_1_as_conv", "vector/src/vector_types/time_freq/convolution.rs", "shift_left_by_1_as_conv", "exp"),
    ("shift_left_by_1_as_conv_shorter", "vector/src/vector_types/time_freq/convolution.rs", "shift_left_by_1_as_conv_shorter", "exp"),
    ("raised_cosine_test", "vector/src/conv_types.rs", "raised_cosine_test", "expected"),
    ("sinc_test", "vector/src/conv_types.rs", "sinc_test", "expected"),
    ("sinc_freq_test", "vector/src/conv_types.rs", "sinc_freq_test", "expected"),
    ("freq_test", "vector/src/conv_types.rs", "freq_test", "expected"),
    ("interpolatef_by_integer_sinc_even_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolatef_by_integer_sinc_even_test", "expected"),
    ("interpolatef_by_integer_sinc_odd_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolatef_by_integer_sinc_odd_test", "expected"),
    ("interpolatef_by_fractional_sinc_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolatef_by_fractional_sinc_test", "expected"),
    ("interpolatef_delayed_sinc_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolatef_delayed_sinc_test", "expected"),
    ("linear_test", "vector/src/vector_types/time_freq/real_interpolation.rs", "linear_test", "expected"),
    ("fft_swap_x_test", "vector/src/vector_types/time_freq/mod.rs", "fft_swap_x_test", "expected"),
    ("time_correlation_test_a", "vector/src/vector_types/time_freq/correlation.rs", "time_correlation_test", "mut a"),
    ("time_correlation_test_b", "vector/src/vector_types/time_freq/correlation.rs", "time_correlation_test", "b"),
    ("time_correlation_test_c", "vector/src/vector_types/time_freq/correlation.rs", "time_correlation_test", "c"),
    ("time_correlation_test2_c", "vector/src/vector_types/time_freq/correlation.rs", "time_correlation_test2", "c"),
    ("interpolatei_sinc_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolatei_sinc_test", "expected"),
    ("interpolate_sinc_even_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolate_sinc_even_test", "expected"),
    ("interpolate_sinc_odd_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolate_sinc_odd_test", "expected"),
    ("interpolate_by_fractional_sinc_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolate_by_fractional_sinc_test", "expected"),
    ("interpolate_by_fractional_sinc_real_data_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolate_by_fractional_sinc_real_data_test", "expected"),
    ("interpolate_delayed_sinc_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolate_delayed_sinc_test", "expected"),
    ("interpolate_identity", "vector/src/vector_types/time_freq/interpolation.rs", "interpolate_identity", "expected"),
    ("decimate_with_interpolate_test", "vector/src/vector_types/time_freq/interpolation.rs", "decimate_with_interpolate_test", "expected"),
    ("interpolatei_rc_test", "vector/src/vector_types/time_freq/interpolation.rs", "interpolatei_rc_test", "expected"),
    ("triangular_window32_test", "vector/src/window_functions.rs", "triangular_window32_test", "expected"),
    ("hamming_window32_test", "vector/src/window_functions.rs", "hamming_window32_test", "expected"),
    ("blackmanharris_window32_test", "vector/src/window_functions.rs", "blackmanharris_window32_test", "expected"),
]

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def extract(path, fn, var):
    src = open(os.path.join(REF, path)).read()
    m = re.search(r"fn\s+%s\s*\(" % re.escape(fn), src)
    if not m:
        raise SystemExit("test fn %s not found in %s" % (fn, path))
    line = src.count("\n", 0, m.start()) + 1
    body = src[m.end():]
    m2 = re.search(r"let\s+%s\s*(?::[^=]*)?=\s*&?(?:vec!)?\[" % re.escape(var), body)
    if not m2:
        raise SystemExit("array %s not found in %s::%s" % (var, path, fn))
    start = m2.end()
    end = body.index("]", start)
    text = re.sub(r"//[^\n]*", "", body[start:end])
    vals = [float(v) for v in re.findall(NUM, text)]
    return {"source": "%s:%d" % (path, line), "values": vals}


def doctest_vectors():
    """3-point doctest vectors (time_to_freq.rs / freq_to_time.rs); literals copied by regex."""
    out = {}
    for key, path, method in [
        ("doc_plain_fft", "vector/src/vector_types/time_freq/time_to_freq.rs", "plain_fft"),
        ("doc_fft", "vector/src/vector_types/time_freq/time_to_freq.rs", "fft"),
        ("doc_plain_ifft", "vector/src/vector_types/time_freq/freq_to_time.rs", "plain_ifft"),
        ("doc_ifft", "vector/src/vector_types/time_freq/freq_to_time.rs", "ifft"),
    ]:
        src = open(os.path.join(REF, path)).read()
        # the doc block sits directly above the trait method declaration `fn <method><B>(`
        decl = re.search(r"\n\s*fn\s+%s<B>\(" % re.escape(method), src)
        if not decl:
            raise SystemExit("trait method %s not found in %s" % (method, path))
        doc = src[:decl.start()]
        doc = doc[doc.rindex("# Example"):]
        line = src.count("\n", 0, decl.start()) + 2
        m_in = re.search(r"vec!\((.*?)\)\.to_complex_(?:time|freq)_vec\(\)", doc, re.S)
        m_ex = re.search(r"let\s+expected\s*=\s*&\[(.*?)\];", doc, re.S)
        if not (m_in and m_ex):
            raise SystemExit("doctest for %s not found in %s" % (method, path))
        out[key] = {
            "source": "%s:%d (doctest above)" % (path, line),
            "input": [float(v) for v in re.findall(NUM, m_in.group(1))],
            "values": [float(v) for v in re.findall(NUM, m_ex.group(1))],
        }
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree %s not present (this script only runs in the build container)" % REF)
    kats = {k: extract(p, f, v) for (k, p, f, v) in CASES}
    kats.update(doctest_vectors())
    with open(OUT, "w") as fh:
        json.dump(kats, fh, indent=1, sort_keys=True)
    print("wrote %s (%d cases)" % (OUT, len(kats)))


if __name__ == "__main__":
    main()
