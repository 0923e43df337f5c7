#!/usr/bin/env python3
"""Generates rust/src/ffi.rs from include/basic_dsp_b200.h: one `extern "C"` declaration per C declaration, so the Rust
binding cannot drift from the header (tests/test_rust_shim.py regenerates it and compares).  bindgen is not available in
this image; the header is plain C11 with one declaration per statement, which this script parses directly.

    python rust/gen_ffi.py            # rewrites rust/src/ffi.rs
    python rust/gen_ffi.py --check    # exit 1 if rust/src/ffi.rs is out of date
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "basic_dsp_b200.h")
OUT = os.path.join(ROOT, "rust", "src", "ffi.rs")

SCALARS = {"void": "c_void", "int32_t": "i32", "uint32_t": "u32", "uint8_t": "u8", "int8_t": "i8", "bool": "bool", "int64_t": "i64", "uint64_t": "u64", "size_t": "usize", "float": "f32",
           "double": "f64", "int": "c_int", "char": "c_char", "unsigned long long": "u64"}
STRUCT_FIELDS = {}   # name -> [(field, ctype)]
OPAQUE = []
FN_TYPES = {}        # typedef name -> (ret, [(arg, type)])

RUST_KEYWORDS = {"type", "ref", "fn", "in", "loop", "match", "move", "impl", "self", "super", "use", "where", "box", "as"}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return text


def rust_type(ctype):
    ctype = ctype.strip()
    if ctype.startswith("@"):
        return ctype[1:]
    const = False
    ptr = 0
    while ctype.endswith("*"):
        ptr += 1
        ctype = ctype[:-1].strip()
    if ctype.startswith("const "):
        const = True
        ctype = ctype[6:].strip()
    if ctype.endswith(" const"):
        const = True
        ctype = ctype[:-6].strip()
    base = SCALARS.get(ctype, ctype)
    if base in FN_TYPES and ptr == 0:
        return base
    out = base
    for i in range(ptr):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    return out


def split_args(args):
    args = args.strip()
    if args in ("", "void"):
        return []
    out = []
    parts, depth, cur = [], 0, ""
    for ch in args:
        depth += ch == "("
        depth -= ch == ")"
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    for i, a in enumerate(parts):
        a = a.strip()
        fp = re.match(r"^(.+?) ?\(\*(\w+)\)\((.*)\)$", a)          # inline function pointer: ret (*name)(args)
        if fp:
            inner = ", ".join("%s: %s" % (n, rust_type(t)) for n, t in split_args(fp.group(3)))
            ret = "" if fp.group(1).strip() == "void" else " -> " + rust_type(fp.group(1))
            out.append((fp.group(2), "@Option<unsafe extern \"C\" fn(%s)%s>" % (inner, ret)))
            continue
        m = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)
        if m and m.group(1).strip() and m.group(1).strip() not in ("const", "unsigned"):
            ctype, name = m.group(1).strip(), m.group(2)
        else:
            ctype, name = a, "arg%d" % i
        if name in RUST_KEYWORDS:
            name += "_"
        out.append((name, ctype))
    return out


def parse(text):
    text = strip_comments(text)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)              # preprocessor lines
    text = text.replace('extern "C" {', "").replace("}", "}")
    decls = []
    for stmt in re.split(r";", text):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        m = re.match(r"^typedef struct (\w+) (\w+)$", stmt)
        if m:
            OPAQUE.append(m.group(2))
            continue
        if stmt.startswith("typedef struct {"):
            continue                                                # handled below (fields contain ';')
        m = re.match(r"^typedef (.+?) ?\(\*(\w+)\)\((.*)\)$", stmt)
        if m:
            FN_TYPES[m.group(2)] = (m.group(1).strip(), split_args(m.group(3)))
            continue
        if stmt.startswith("typedef") or "{" in stmt or "(" not in stmt:
            continue
        head, rest = stmt.split("(", 1)
        m = re.match(r"^(.+?[ \*])(\w+) ?$", head)
        if not m:
            continue
        depth, end = 1, None
        for i, ch in enumerate(rest):
            depth += ch == "("
            depth -= ch == ")"
            if depth == 0:
                end = i
                break
        args, tail = rest[:end], rest[end + 1:].strip()
        sym = re.match(r'^BDSP_SYMBOL\("(\w+)"\)$', tail)
        if tail and not sym:
            raise SystemExit("cannot parse: " + stmt)
        decls.append((m.group(1).strip(), m.group(2), split_args(args), sym.group(1) if sym else None))
    return decls


def parse_structs(text):
    text = strip_comments(text)
    for m in re.finditer(r"typedef struct \{(.*?)\} (\w+);", text, flags=re.S):
        fields = []
        for f in m.group(1).split(";"):
            f = " ".join(f.split())
            if not f:
                continue
            parts = [x.strip() for x in f.split(",")]
            first = parts[0].split()
            ctype = " ".join(first[:-1])
            for nm in [first[-1]] + parts[1:]:
                fields.append((nm, ctype))
        STRUCT_FIELDS[m.group(2)] = fields


def generate():
    text = open(HEADER).read()
    parse_structs(text)
    decls = parse(text)
    lines = ["// GENERATED by rust/gen_ffi.py from include/basic_dsp_b200.h - do not edit.",
             "// One declaration per C declaration of the header (the reference's interop exports, interop/src/facade32.rs /",
             "// facade64.rs, plus the bdsp_* extensions).  Not compiled in this repository's image (no Rust toolchain).",
             "#![allow(non_camel_case_types, non_snake_case, dead_code, clippy::too_many_arguments)]",
             "use std::os::raw::{c_char, c_int, c_void};", ""]
    for name in OPAQUE:
        lines += ["#[repr(C)]", "pub struct %s { _private: [u8; 0] }" % name, ""]
    for name, fields in STRUCT_FIELDS.items():
        lines += ["#[repr(C)]", "#[derive(Clone, Copy)]", "pub struct %s {" % name]
        lines += ["    pub %s: %s," % (f if f not in RUST_KEYWORDS else f + "_", rust_type(t)) for f, t in fields]
        lines += ["}", ""]
    for name, (ret, args) in FN_TYPES.items():
        a = ", ".join("%s: %s" % (n, rust_type(t)) for n, t in args)
        r = "" if ret == "void" else " -> " + rust_type(ret)
        lines.append("pub type %s = Option<unsafe extern \"C\" fn(%s)%s>;" % (name, a, r))
    lines += ["", "extern \"C\" {"]
    for ret, name, args, sym in decls:
        a = ", ".join("%s: %s" % (n, rust_type(t)) for n, t in args)
        r = "" if ret == "void" else " -> " + rust_type(ret)
        if sym:
            lines.append("    #[link_name = \"%s\"]" % sym)
        lines.append("    pub fn %s(%s)%s;" % (name, a, r))
    lines += ["}", ""]
    return "\n".join(lines), [d[1] for d in decls]


def main():
    src, names = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != src:
            print("rust/src/ffi.rs is out of date: run python rust/gen_ffi.py")
            return 1
        return 0
    with open(OUT, "w") as fh:
        fh.write(src)
    print("wrote %s (%d functions)" % (OUT, len(names)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
