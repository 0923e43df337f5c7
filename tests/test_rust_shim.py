"""The Rust FFI shim (rust/) cannot be compiled in this image (no Rust toolchain); these checks keep it honest:
ffi.rs is generated from the header and must be current, must cover every exported C symbol, and lib.rs / build.rs
must only refer to things that exist."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ffi_rs_is_generated_from_the_current_header():
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "rust", "gen_ffi.py"), "--check"], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stdout + rc.stderr


def _ffi():
    src = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    fns = set(re.findall(r"pub fn (\w+)\(", src))
    links = dict(re.findall(r'#\[link_name = "(\w+)"\]\s*pub fn (\w+)\(', src))
    return src, fns, links


def test_ffi_rs_covers_every_exported_symbol():
    import basic_dsp_b200.build as build
    lib = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and not l.split()[-1].startswith("_Z")}
    _, fns, links = _ffi()
    bound = (fns - set(links.values())) | set(links.keys())
    assert exported - bound == set()
    assert bound - exported == set()


def test_lib_rs_uses_only_declared_ffi_functions_with_matching_arity():
    src, fns, _ = _ffi()
    lib = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    m = re.search(r"macro_rules! gpu_vec \{\s*\((.*?)\) => \{", lib, flags=re.S)
    params = re.findall(r"\$(\w+):ident", m.group(1))
    for call in re.findall(r"gpu_vec!\((.*?)\);", lib, flags=re.S):
        args = [a.strip() for a in call.replace("\n", " ").split(",")]
        assert len(args) >= len(params)
        names = dict(zip(params, args[-len(params):] if False else args[len(args) - len(params):]))
        # every macro argument that is an FFI function must be declared in ffi.rs
        for p_, a in names.items():
            if p_ in ("Vec", "real_ir", "real_fr", "window"):
                continue
            assert a in fns, (p_, a)
    # arity of the calls written inside the macro body against the declarations (32-bit instantiation)
    decl = {n: len([x for x in args_.split(",") if x.strip()]) for n, args_ in re.findall(r"pub fn (\w+)\(([^;]*?)\)(?: ->|;)", src)}
    first = re.findall(r"gpu_vec!\((.*?)\);", lib, flags=re.S)[0].replace("\n", " ")
    actual = dict(zip(params, [a.strip() for a in first.split(",")][-len(params):]))
    body = lib[lib.index("=> {"):lib.index("gpu_vec!(GpuVec32")]
    for p_, cargs in re.findall(r"ffi::\$(\w+)\(([^;]*?)\) \}", body):
        depth, n, cur = 0, 0, ""
        for ch in cargs:
            depth += ch in "(<"
            depth -= ch in ")>"
            if ch == "," and depth == 0:
                n += 1
            cur += ch
        n = n + 1 if cur.strip() else 0
        assert decl[actual[p_]] == n, (p_, actual[p_], decl[actual[p_]], n)


def test_build_rs_lists_the_library_sources():
    import basic_dsp_b200.build as build
    rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    listed = re.findall(r'"(\w+\.cu)"', rs)
    assert listed == build.SOURCES
