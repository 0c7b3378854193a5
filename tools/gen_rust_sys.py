#!/usr/bin/env python
"""Generate the Rust FFI declarations of the drop-in boundary from the C headers (a small bindgen for this ABI's C subset).

    tools/gen_rust_sys.py            writes rust/hikari-b200-sys/src/lib.rs from include/hikari_b200.h (+ hk_layout.h)
    tools/gen_rust_sys.py --check    exits 1 if the committed file differs from what the headers generate

BASELINE.json's north_star puts the host side in Rust ("cudarc + cc" over a thin C ABI).  This image has no Rust toolchain, so the
crate cannot be compiled here; what CAN be guaranteed mechanically is that the declarations a Bevy host would compile against are
exactly the header's: every struct field for field with `#[repr(C)]`, every constant, every function — and the struct sizes gcc
computes are emitted as compile-time assertions on the Rust side.  tests/test_rust_bindings.py runs --check and compares the
extern block with the symbols libhikari_b200.so exports."""
import os
import subprocess
import sys
import tempfile

from pycparser import c_ast, c_parser

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "rust", "hikari-b200-sys", "src", "lib.rs")
HEADER = "hikari_b200.h"

FAKE = {"stdint.h": "typedef unsigned char uint8_t; typedef unsigned short uint16_t; typedef unsigned int uint32_t;\n"
                    "typedef unsigned long uint64_t; typedef signed char int8_t; typedef short int16_t; typedef int int32_t;\n"
                    "typedef long int64_t;\n",
        "stddef.h": "typedef unsigned long size_t;\n#define offsetof(t, m) 0\n"}
PRIM = {"uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int8_t": "i8", "int16_t": "i16", "int32_t": "i32",
        "int64_t": "i64", "size_t": "usize", "float": "f32", "double": "f64", "int": "::core::ffi::c_int", "unsigned int": "::core::ffi::c_uint",
        "char": "::core::ffi::c_char", "void": "::core::ffi::c_void"}


def preprocess():
    with tempfile.TemporaryDirectory() as d:
        for name, text in FAKE.items():
            open(os.path.join(d, name), "w").write(text)
        r = subprocess.run(["gcc", "-E", "-P", "-nostdinc", "-I", d, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "include", HEADER)],
                           capture_output=True, text=True, check=True)
    return r.stdout


def defines():
    """#define NAME (int) of the header family"""
    out = []
    for h in ("hk_layout.h", HEADER):
        for line in open(os.path.join(ROOT, "include", h)):
            p = line.split()
            if len(p) >= 3 and p[0] == "#define" and p[1].startswith("HK_") and "(" not in p[1]:
                val = p[2].strip("()")
                try:
                    out.append((p[1], int(val, 0)))
                except ValueError:
                    pass
    return out


def c_sizes(structs):
    """sizeof of every struct, from gcc — the Rust side asserts the same numbers"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        with open(src, "w") as f:
            f.write('#include <stdio.h>\n#include "%s"\nint main(void) {\n' % HEADER)
            for s in structs:
                f.write('  printf("%s %%zu\\n", sizeof(%s));\n' % (s, s))
            f.write("  return 0;\n}\n")
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return dict((l.split()[0], int(l.split()[1])) for l in out.splitlines())


class Gen:
    def __init__(self):
        self.structs, self.opaque, self.consts, self.funcs = [], [], [], []
        self.anonymous = {}

    def rust_type(self, t, in_param=False):
        if isinstance(t, c_ast.TypeDecl):
            inner = t.type
            if isinstance(inner, c_ast.IdentifierType):
                name = " ".join(inner.names)
                return PRIM.get(name, name)
            if isinstance(inner, c_ast.Struct):
                return inner.name or self.anonymous[id(inner)]
            raise ValueError(inner)
        if isinstance(t, c_ast.PtrDecl):
            const = "const" in (t.type.quals if hasattr(t.type, "quals") else [])
            return ("*const " if const else "*mut ") + self.rust_type(t.type)
        if isinstance(t, c_ast.ArrayDecl):
            if in_param:                                   # a C array parameter is a pointer
                const = "const" in (t.type.quals if hasattr(t.type, "quals") else [])
                return ("*const " if const else "*mut ") + self.rust_type(t.type)
            return "[%s; %s]" % (self.rust_type(t.type), t.dim.value)
        raise ValueError(type(t))

    def struct(self, name, s):
        fields = []
        for d in s.decls:
            if isinstance(d.type, c_ast.TypeDecl) and isinstance(d.type.type, c_ast.Struct) and d.type.type.decls and not d.type.type.name:
                inner = "%s_%s" % (name, d.name)           # anonymous nested struct -> a named one
                self.anonymous[id(d.type.type)] = inner
                self.struct(inner, d.type.type)
            elif isinstance(d.type, c_ast.ArrayDecl) and isinstance(d.type.type, c_ast.TypeDecl) and isinstance(d.type.type.type, c_ast.Struct) \
                    and d.type.type.type.decls and not d.type.type.type.name:
                inner = "%s_%s" % (name, d.name)
                self.anonymous[id(d.type.type.type)] = inner
                self.struct(inner, d.type.type.type)
            fields.append((d.name, self.rust_type(d.type)))
        self.structs.append((name, fields))

    def visit(self, ast):
        for ext in ast.ext:
            if isinstance(ext, c_ast.Typedef) and isinstance(ext.type, c_ast.TypeDecl) and isinstance(ext.type.type, c_ast.Struct):
                s = ext.type.type
                if s.decls is None:
                    if ext.name not in self.opaque:
                        self.opaque.append(ext.name)
                else:
                    self.struct(ext.name, s)
            elif isinstance(ext, c_ast.Decl) and isinstance(ext.type, c_ast.Enum):
                value = -1
                for e in ext.type.values.enumerators:
                    value = _eval(e.value, value)
                    self.consts.append((e.name, value))
            elif isinstance(ext, c_ast.Decl) and isinstance(ext.type, c_ast.FuncDecl):
                f = ext.type
                params = []
                if f.args:
                    for p in f.args.params:
                        if isinstance(p, c_ast.Typename) or (isinstance(p.type, c_ast.TypeDecl) and isinstance(p.type.type, c_ast.IdentifierType)
                                                             and p.type.type.names == ["void"] and p.name is None):
                            continue
                        params.append((p.name, self.rust_type(p.type, in_param=True)))
                ret = self.rust_type(f.type)
                self.funcs.append((ext.name, params, None if ret == PRIM["void"] else ret))


RUST_KEYWORDS = {"in", "type", "ref", "fn", "mod", "loop", "match", "move", "self", "super", "use", "where", "box", "dyn", "impl", "trait", "as"}


def ident(name):
    return "r#" + name if name in RUST_KEYWORDS else name


def _eval(node, previous):
    if node is None:
        return previous + 1
    if isinstance(node, c_ast.Constant):
        return int(node.value, 0)
    if isinstance(node, c_ast.UnaryOp) and node.op == "-":
        return -_eval(node.expr, 0)
    raise ValueError(node)


def generate():
    ast = c_parser.CParser().parse(preprocess(), filename=HEADER)
    g = Gen()
    g.visit(ast)
    sizes = c_sizes([n for n, _ in g.structs if "_" in n and not any(n == a for a in g.anonymous.values())])
    o = []
    o.append("//! FFI declarations of libhikari_b200.so — GENERATED by tools/gen_rust_sys.py from include/hikari_b200.h and")
    o.append("//! include/hk_layout.h; do not edit.  `tools/gen_rust_sys.py --check` (run by tests/test_rust_bindings.py) keeps this file")
    o.append("//! identical to what the headers generate.  Not compiled in the build container (no Rust toolchain there); the size")
    o.append("//! assertions at the end make a mismatch with the C side a compile error wherever it IS compiled.")
    o.append("#![no_std]")
    o.append("#![allow(non_camel_case_types, non_upper_case_globals)]")
    o.append("")
    for name, value in defines():
        o.append("pub const %s: ::core::ffi::c_int = %d;" % (name, value))
    for name, value in g.consts:
        o.append("pub const %s: ::core::ffi::c_int = %d;" % (name, value))
    o.append("")
    for name in g.opaque:
        o.append("#[repr(C)] pub struct %s { _private: [u8; 0] }" % name)
    o.append("")
    for name, fields in g.structs:
        o.append("#[repr(C)]")
        o.append("#[derive(Clone, Copy)]")
        o.append("pub struct %s {" % name)
        for f, t in fields:
            o.append("    pub %s: %s," % (ident(f), t))
        o.append("}")
    o.append("")
    o.append('#[link(name = "hikari_b200")]')
    o.append('extern "C" {')
    for name, params, ret in g.funcs:
        o.append("    pub fn %s(%s)%s;" % (name, ", ".join("%s: %s" % (ident(p) if p else "_arg%d" % i, t) for i, (p, t) in enumerate(params)),
                                             " -> " + ret if ret else ""))
    o.append("}")
    o.append("")
    o.append("// sizeof() as gcc lays the structs out (x86-64 SysV; the same rules `repr(C)` follows)")
    for name, size in sizes.items():
        o.append("const _: () = assert!(::core::mem::size_of::<%s>() == %d);" % (name, size))
    return "\n".join(o) + "\n", g, sizes


def main():
    text, _, _ = generate()
    if "--check" in sys.argv:
        if not os.path.exists(OUT) or open(OUT).read() != text:
            sys.exit("rust/hikari-b200-sys/src/lib.rs is out of date: run tools/gen_rust_sys.py")
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print("wrote", OUT, len(text.splitlines()), "lines")


if __name__ == "__main__":
    main()
