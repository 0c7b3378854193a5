#!/usr/bin/env python
"""TEST INFRASTRUCTURE — executes the reference's OWN compute shaders on the CPU, to pin the oracle.

The reference's hot path is WGSL (src/shaders/{light,denoise,tone_mapping,smaa,taa}.wgsl); nothing in this container can run WGSL (no naga,
no wgpu, no Vulkan).  This script translates those files — read where they lie under /root/reference, never copied into the repo —
into C++ that g++ compiles against oracle/wgsl/wgsl_rt.h, one shared library per (file, shader-def set) exactly as the reference
specialises its pipelines (src/light.rs:134-175, src/post_process.rs:396-500).  The generated sources and libraries go to
oracle/_ref/ (git-ignored).  oracle/wgsl/run_reference.py drives them pass by pass and tools/make_wgsl_golden.py stores what they
compute as fixtures under tests/golden/; tests/test_wgsl_reference.py holds the oracle against those.

What is translated, token by token (WGSL is close enough to C++ that expressions pass through untouched):
  preprocessing   bevy's `#import`, `#define_import_path`, `#ifdef / #ifndef / #else / #endif`
  declarations    `fn f(a: T) -> R`, `let` / `var` (module and function scope), `struct`, `type A = B;`, attributes dropped
  statements      `if c {` / `else if c {` / `while c {` get their parentheses; `for (var i = 0u; ...)`
  types           `vec3<f32>`, `array<T, N>`, `mat3x3<f32>`, textures ... are templates of the same names in wgsl_rt.h;
                  `ptr<function, T>` becomes `T*`
  literals        `1.0` -> `1.0f` (WGSL abstract floats are f32 here), `1i` -> `1`
  swizzles        `.xyz` -> `.swz<0,1,2>()`, single `.r/.g/.b/.a` -> `.x/.y/.z/.w`
  struct layout   explicit padding from WGSL's alignment rules, so that storage / uniform buffers are bound as the very bytes the
                  reference's host code writes (static_asserts on every offset)
Not handled because the three files do not use them: loop / continuing, switch, atomics, overrides, pointers to anything but locals.
The two bevy modules the shaders import but that are not under /root/reference (bevy_pbr::{utils, lighting, mesh_view_types},
bevy_core_pipeline::tonemapping) are supplied from oracle/wgsl/prelude/*.wgsl, restated from bevy 0.9.1 — the one part
of the executed code that is NOT the reference's own text."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SHADERS = "/root/reference/src/shaders"
PRELUDE = os.path.join(HERE, "prelude")

CPP_KEYWORDS = {"auto", "class", "new", "delete", "this", "template", "typename", "public", "private", "protected", "signed", "unsigned",
                "short", "long", "double", "float", "int", "char", "register", "union", "namespace", "using", "operator", "friend",
                "virtual", "static", "extern", "inline", "goto", "try", "catch", "throw", "typedef", "volatile", "and", "or", "not",
                "xor", "near", "far", "export", "import", "module", "sample"}


# ------------------------------------------------------------------------------------------------ preprocessing
def index_modules():
    """`#define_import_path name` -> file, over the reference's shaders and the prelude"""
    mods = {}
    for d in (REF_SHADERS, PRELUDE):
        for f in sorted(os.listdir(d)):
            if f.endswith(".wgsl"):
                text = open(os.path.join(d, f)).read()
                m = re.search(r"^#define_import_path\s+(\S+)", text, re.M)
                if m:
                    mods[m.group(1)] = os.path.join(d, f)
    return mods


def preprocess(path, defs, mods=None, seen=None):
    mods = mods if mods is not None else index_modules()
    seen = seen if seen is not None else set()
    out, stack = [], []           # stack of (parent_active, this_branch_taken)
    active = True
    for line in open(path).read().splitlines():
        s = line.strip()
        if s.startswith("#ifdef") or s.startswith("#ifndef"):
            name = s.split()[1]
            cond = (name in defs) if s.startswith("#ifdef") else (name not in defs)
            stack.append((active, cond))
            active = active and cond
        elif s.startswith("#else"):
            parent, taken = stack[-1]
            active = parent and not taken
        elif s.startswith("#endif"):
            parent, _ = stack.pop()
            active = parent
        elif not active:
            continue
        elif s.startswith("#define_import_path"):
            continue
        elif s.startswith("#import"):
            name = s.split()[1]
            if name in seen:
                continue
            seen.add(name)
            if name not in mods:
                raise SystemExit(f"{path}: #import {name}: no such module under the reference's shaders or oracle/wgsl/prelude")
            out.append(f"// ---- #import {name} ({'reference' if mods[name].startswith(REF_SHADERS) else 'PRELUDE, restated'})")
            out.append(preprocess(mods[name], defs, mods, seen))
            out.append(f"// ---- end of {name}")
        else:
            out.append(line)
    return "\n".join(out)


# ------------------------------------------------------------------------------------------------ tokens
TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+[iu]?|(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fiuh]?)
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>->|<<=|>>=|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|\+\+|--|[-+*/%&|^~!<>=.,;:(){}\[\]@])
""", re.X | re.S)


def tokenize(text):
    toks, pos = [], 0
    while pos < len(text):
        m = TOKEN.match(text, pos)
        if not m:
            raise SystemExit(f"cannot tokenize at: {text[pos:pos + 40]!r}")
        pos = m.end()
        if m.lastgroup != "ws":
            toks.append((m.lastgroup, m.group()))
    return toks


TEMPLATE_TYPES = {"vec2", "vec3", "vec4", "mat2x2", "mat3x3", "mat4x4", "mat3x4", "mat4x3", "array", "ptr", "atomic", "texture_2d",
                  "texture_storage_2d", "binding_array", "texture_2d_array", "bitcast"}
SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


class Translator:
    def __init__(self, text):
        self.t = tokenize(text)
        self.i = 0
        self.structs = {}         # name -> [(field, type_tokens)]
        self.aliases = {}         # name -> type tokens
        self.functions = []       # (name, ret, params, body_cpp, attrs)
        self.globals = []         # (name, type_tokens, space, attrs)
        self.consts = []          # cpp lines
        self.order = []           # ('struct', name) | ('const', line) | ('alias', name) in source order

    # -------------------------------------------------------------------------------------------- helpers
    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, value):
        tok = self.next()
        if tok[1] != value:
            ctx = " ".join(v for _, v in self.t[max(0, self.i - 12):self.i + 6])
            raise SystemExit(f"expected {value!r}, got {tok[1]!r} near: {ctx}")
        return tok

    def attributes(self):
        attrs = []
        while self.peek()[1] == "@":
            self.next()
            name = self.next()[1]
            args = []
            if self.peek()[1] == "(":
                self.next()
                depth = 1
                while depth:
                    tok = self.next()
                    if tok[1] == "(":
                        depth += 1
                    elif tok[1] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                    args.append(tok[1])
            attrs.append((name, "".join(args)))
        return attrs

    def type_tokens(self):
        """consumes one type expression; returns its tokens (a `>>` that closes two brackets is split)"""
        out = [self.next()]
        if self.peek()[1] == "<":
            depth = 0
            while True:
                tok = self.next()
                if tok[1] == ">>":
                    out.extend([("op", ">"), ("op", ">")])
                    depth -= 2
                else:
                    out.append(tok)
                    if tok[1] == "<":
                        depth += 1
                    elif tok[1] == ">":
                        depth -= 1
                if depth <= 0:
                    break
        return out

    def ctype(self, toks):
        """type tokens -> C++ spelling"""
        if toks[0][1] == "ptr":          # ptr<function, T>
            inner = toks[4:-1]
            return self.ctype(inner) + "*"
        parts = []
        for kind, v in toks:
            if kind == "num":
                parts.append(self.number(v))
            else:
                parts.append(self.ident(v) if kind == "id" else v)
        return "".join(p + (" " if p in (",", ">") else "") for p in parts).strip()

    @staticmethod
    def number(v):
        if v.lower().startswith("0x"):
            return v[:-1] if v.endswith("i") else v
        if v.endswith("i"):
            return v[:-1]
        if v.endswith("h"):
            return v[:-1] + "f"
        if v.endswith("u") or v.endswith("f"):
            return v
        if "." in v or "e" in v.lower():
            return v + "f"
        return v

    @staticmethod
    def ident(v):
        return v + "_" if v in CPP_KEYWORDS else v

    def expr(self, toks):
        """expression tokens -> C++"""
        out = []
        k = 0
        while k < len(toks):
            kind, v = toks[k]
            if kind == "num":
                out.append(self.number(v))
            elif v == "." and k + 1 < len(toks) and toks[k + 1][0] == "id":
                name = toks[k + 1][1]
                is_call = k + 2 < len(toks) and toks[k + 2][1] == "("
                if not is_call and 2 <= len(name) <= 4 and (all(c in "xyzw" for c in name) or all(c in "rgba" for c in name)):
                    out.append(".swz<" + ",".join(str(SWZ[c]) for c in name) + ">()")
                elif not is_call and len(name) == 1 and name in "rgba" and not self.field_named(name):
                    out.append("." + "xyzw"[SWZ[name]])
                else:
                    out.append("." + self.ident(name))
                k += 1
            elif kind == "id" and v == "ptr" and k + 1 < len(toks) and toks[k + 1][1] == "<":
                raise SystemExit("ptr type inside an expression")
            elif kind == "id":
                out.append(self.ident(v))
            elif v == "," and k + 1 < len(toks) and toks[k + 1][1] == ")":
                pass                                  # WGSL allows a trailing comma in argument lists (smaa.wgsl:231-234)
            else:
                out.append(v)
            k += 1
        return self.join(out)

    def field_named(self, name):
        return any(name == f for fields in self.structs.values() for f, _ in fields)

    @staticmethod
    def join(parts):
        s = ""
        for p in parts:
            if s and (s[-1].isalnum() or s[-1] == "_") and (p[0].isalnum() or p[0] == "_"):
                s += " "
            s += p
        return s

    def until(self, stops, consume=True):
        """tokens up to (not including) the first token in `stops` at bracket depth 0"""
        out, depth = [], 0
        while True:
            tok = self.peek()
            if tok[0] == "eof":
                raise SystemExit("unexpected end of file")
            if depth == 0 and tok[1] in stops:
                if consume:
                    self.next()
                return out
            if tok[1] in "([":
                depth += 1
            elif tok[1] in ")]":
                depth -= 1
            out.append(self.next())

    # -------------------------------------------------------------------------------------------- module
    def module(self):
        while self.peek()[0] != "eof":
            attrs = self.attributes()
            kind, v = self.peek()
            if v == "struct":
                self.struct()
            elif v == "fn":
                self.function(attrs)
            elif v == "var":
                self.global_var(attrs)
            elif v in ("let", "const"):
                self.next()
                name = self.next()[1]
                ty = None
                if self.peek()[1] == ":":
                    self.next()
                    ty = self.ctype(self.type_tokens())
                self.expect("=")
                e = self.expr(self.until({";"}))
                line = f"static const {ty or 'auto'} {self.ident(name)} = {e};"
                self.order.append(("const", line))
            elif v in ("type", "alias"):
                self.next()
                name = self.next()[1]
                self.expect("=")
                toks = self.type_tokens()
                self.expect(";")
                self.aliases[name] = toks
                self.order.append(("alias", name))
            elif v == ";":
                self.next()
            else:
                raise SystemExit(f"unexpected token at module scope: {v!r}")

    def struct(self):
        self.expect("struct")
        name = self.next()[1]
        self.expect("{")
        fields = []
        while self.peek()[1] != "}":
            self.attributes()
            fname = self.next()[1]
            self.expect(":")
            fields.append((fname, self.type_tokens()))
            if self.peek()[1] == ",":
                self.next()
        self.expect("}")
        if self.peek()[1] == ";":
            self.next()
        self.structs[name] = fields
        self.order.append(("struct", name))

    def global_var(self, attrs):
        self.expect("var")
        space = "handle"
        if self.peek()[1] == "<":
            toks = []
            self.next()
            while self.peek()[1] != ">":
                toks.append(self.next()[1])
            self.next()
            space = toks[0]
        name = self.next()[1]
        self.expect(":")
        ty = self.type_tokens()
        init = None
        if self.peek()[1] == "=":
            self.next()
            init = self.expr(self.until({";"}, consume=False))
        self.expect(";")
        self.globals.append((name, ty, space, attrs, init))

    def function(self, attrs):
        self.expect("fn")
        name = self.next()[1]
        self.expect("(")
        params = []
        while self.peek()[1] != ")":
            pattrs = self.attributes()
            pname = self.next()[1]
            self.expect(":")
            params.append((pname, self.type_tokens(), pattrs))
            if self.peek()[1] == ",":
                self.next()
        self.expect(")")
        ret = None
        if self.peek()[1] == "->":
            self.next()
            self.attributes()
            ret = self.type_tokens()
        body = self.block(1, [p[0] for p in params])
        self.functions.append((name, ret, params, body, attrs))

    # -------------------------------------------------------------------------------------------- statements
    def block(self, indent=1, params=()):
        self.expect("{")
        lines = []
        pad = "    " * indent
        declared, extra = set(params), 0
        self.scopes = getattr(self, "scopes", [])
        self.scopes.append(declared)
        while self.peek()[1] != "}":
            if self.peek()[1] in ("let", "var") and self.peek(1)[1] in declared:
                # naga (0.10) lets a later declaration shadow an earlier one of the same scope; C++ needs a nested block for that
                lines.append("{")
                extra += 1
                declared = set()
                self.scopes[-1] = declared
            if self.peek()[1] in ("let", "var"):
                declared.add(self.peek(1)[1])
            lines.extend(self.statement(indent))
        self.expect("}")
        self.scopes.pop()
        lines.extend(["}"] * extra)
        return "{\n" + "".join(pad + l + "\n" for l in lines) + "    " * (indent - 1) + "}"

    def statement(self, indent):
        kind, v = self.peek()
        if v == "{":
            return [self.block(indent + 1)]
        if v in ("let", "var"):
            return [self.declaration() + ";"]
        if v == "if":
            self.next()
            cond = self.expr(self.until({"{"}, consume=False))
            s = f"if ({cond}) " + self.block(indent + 1)
            while self.peek()[1] == "else":
                self.next()
                if self.peek()[1] == "if":
                    self.next()
                    cond = self.expr(self.until({"{"}, consume=False))
                    s += f" else if ({cond}) " + self.block(indent + 1)
                else:
                    s += " else " + self.block(indent + 1)
                    break
            return [s]
        if v == "while":
            self.next()
            cond = self.expr(self.until({"{"}, consume=False))
            return [f"while ({cond}) " + self.block(indent + 1)]
        if v == "for":
            self.next()
            self.expect("(")
            init = ""
            if self.peek()[1] != ";":
                init = self.declaration() if self.peek()[1] in ("let", "var") else self.expr(self.until({";"}, consume=False))
            self.expect(";")
            cond = self.expr(self.until({";"}))
            update = self.expr(self.until({")"}))
            return [f"for ({init}; {cond}; {update}) " + self.block(indent + 1)]
        if v in ("loop", "switch", "continuing"):
            raise SystemExit(f"WGSL construct not translated: {v}")
        if v == "return":
            self.next()
            e = self.expr(self.until({";"}))
            return [f"return {e};" if e else "return;"]
        if v in ("break", "continue"):
            self.next()
            self.expect(";")
            return [v + ";"]
        if v == ";":
            self.next()
            return []
        e = self.expr(self.until({";"}))
        return [e + ";"]

    def declaration(self):
        kw = self.next()[1]
        name = self.ident(self.next()[1])
        ty = None
        if self.peek()[1] == ":":
            self.next()
            ty = self.ctype(self.type_tokens())
        if self.peek()[1] == "=":
            self.next()
            e = self.expr(self.until({";"}, consume=False))
            if kw == "let":
                return f"const {ty or 'auto'} {name} = {e}"
            return f"{ty} {name} = {e}" if ty else f"auto {name} = {e}"
        return f"{ty} {name}{{}}"

    # -------------------------------------------------------------------------------------------- layout
    def resolve(self, toks):
        while len(toks) == 1 and toks[0][1] in self.aliases:
            toks = self.aliases[toks[0][1]]
        return toks

    def layout(self, toks):
        """(align, size) of a host-shareable type by WGSL's rules; None for runtime-sized arrays and handles"""
        toks = self.resolve(toks)
        head = toks[0][1]
        if head in ("f32", "u32", "i32"):
            return 4, 4
        if head in ("vec2", "vec3", "vec4"):
            n = int(head[3])
            return {2: 8, 3: 16, 4: 16}[n], 4 * n
        if head.startswith("mat"):
            cols, rows = int(head[3]), int(head[5])
            al = {2: 8, 3: 16, 4: 16}[rows]
            return al, cols * al
        if head == "array":
            inner = toks[2:-1]
            depth, split = 0, None
            for j, (_, v) in enumerate(inner):
                if v == "<":
                    depth += 1
                elif v == ">":
                    depth -= 1
                elif v == ">>":
                    depth -= 2
                elif v == "," and depth == 0:
                    split = j
            if split is None:
                return None
            elem = self.layout(inner[:split])
            n = int(re.sub(r"[ui]$", "", inner[split + 1][1]), 0)
            stride = -(-elem[1] // elem[0]) * elem[0]
            return elem[0], stride * n
        if head in self.structs:
            return self.struct_layout(head)[:2]
        return None

    def struct_layout(self, name):
        off, align, members = 0, 1, []
        for fname, ty in self.structs[name]:
            l = self.layout(ty)
            if l is None and not self.is_runtime_array(ty):      # bool, handles: not host-shareable, natural C++ layout
                members.append((fname, ty, off, 0))
                continue
            if l is None:                     # runtime-sized array: last member
                elem = self.layout(self.resolve(ty)[2:-1])
                off = -(-off // elem[0]) * elem[0]
                members.append((fname, ty, off, None))
                align = max(align, elem[0])
                return align, None, members
            off = -(-off // l[0]) * l[0]
            members.append((fname, ty, off, l[1]))
            off += l[1]
            align = max(align, l[0])
        size = -(-off // align) * align
        return align, size, members

    def is_runtime_array(self, toks):
        toks = self.resolve(toks)
        return toks[0][1] == "array" and self.layout(toks) is None

    def is_handle(self, toks):
        return self.resolve(toks)[0][1] in ("texture_2d", "texture_storage_2d", "binding_array", "sampler", "texture_2d_array")

    # -------------------------------------------------------------------------------------------- output
    def emit(self, banner):
        o = [f"// GENERATED by oracle/wgsl/wgsl2cpp.py — {banner}", "// never committed (oracle/_ref/ is git-ignored); derived from the reference's shader text",
             '#include "wgsl_rt.h"', "namespace wgsl {", ""]
        for kind, item in self.order:
            if kind == "const":
                o.append(item)
            elif kind == "alias":
                o.append(f"typedef {self.ctype(self.aliases[item])} {item};")
            else:
                o.extend(self.emit_struct(item))
        o.append("")
        for name, ty, space, attrs, init in self.globals:
            cty = self.ctype(ty)
            tl = "thread_local " if space == "private" else ""
            o.append(f"static {tl}{cty} {self.ident(name)}" + (f" = {init};" if init else "{};"))
        o.append("")
        for name, ret, params, body, attrs in self.functions:
            o.append(self.signature(name, ret, params) + ";")
        o.append("")
        for name, ret, params, body, attrs in self.functions:
            o.append(self.signature(name, ret, params) + " " + body)
            o.append("")
        o.extend(self.emit_bindings())
        o.append("}  // namespace wgsl")
        return "\n".join(o) + "\n"

    def signature(self, name, ret, params):
        ps = ", ".join(f"{self.ctype(t)} {self.ident(n)}" for n, t, _ in params)
        return f"static {self.ctype(ret) if ret else 'void'} {self.ident(name)}({ps})"

    def emit_struct(self, name):
        align, size, members = self.struct_layout(name)
        lines = [f"struct {name} {{"]
        cursor, pad = 0, 0
        host_shareable = all(self.layout(t) is not None or self.is_runtime_array(t) for _, t in self.structs[name])
        if not host_shareable:
            return [f"struct {name} {{"] + [f"    {self.ctype(t)} {self.ident(f)}{{}};" for f, t in self.structs[name]] + ["};"]
        for fname, ty, off, sz in members:
            if host_shareable and off > cursor:
                lines.append(f"    char _pad{pad}[{off - cursor}];")
                pad += 1
            if sz is None:
                elem = self.ctype(self.resolve(ty)[2:-1])
                lines.append(f"    array<{elem}> {self.ident(fname)};")
                cursor = None
            else:
                lines.append(f"    {self.ctype(ty)} {self.ident(fname)};")
                cursor = off + sz
        if host_shareable and size is not None and cursor is not None and size > cursor:
            lines.append(f"    char _pad{pad}[{size - cursor}];")
        lines.append("};")
        if host_shareable and size is not None:
            lines.append(f'static_assert(sizeof({name}) == {size}, "WGSL layout of {name}");')
            for fname, ty, off, sz in members:
                lines.append(f'static_assert(offsetof({name}, {self.ident(fname)}) == {off}, "{name}.{fname}");')
        return lines

    def emit_bindings(self):
        """extern "C" entry points for the driver: bind_<global>(ptr, bytes, ...) and run_<entry>(grid)"""
        o = ['extern "C" {']
        for name, ty, space, attrs, init in self.globals:
            cname = self.ident(name)
            r = self.resolve(ty)
            head = r[0][1]
            if space in ("uniform", "storage") and self.is_runtime_array(ty):
                elem = self.ctype(r[2:-1])
                o.append(f"void bind_{name}(void* p, size_t bytes) {{ {cname}.ptr = ({elem}*)p; {cname}.len = bytes / sizeof({elem}); }}")
            elif space in ("uniform", "storage") and head in self.structs and self.struct_layout(head)[1] is None:
                _, _, members = self.struct_layout(head)
                fname, fty, off, _ = members[-1]
                elem = self.ctype(self.resolve(fty)[2:-1])
                o.append(f"void bind_{name}(void* p, size_t bytes) {{ if ({off}) memcpy((void*)&{cname}, p, {off} < bytes ? {off} : bytes); "
                         f"{cname}.{self.ident(fname)}.ptr = ({elem}*)((char*)p + {off}); {cname}.{self.ident(fname)}.len = bytes > {off} ? (bytes - {off}) / sizeof({elem}) : 0; }}")
            elif space in ("uniform", "storage"):
                o.append(f"void bind_{name}(const void* p, size_t bytes) {{ memcpy((void*)&{cname}, p, bytes < sizeof({cname}) ? bytes : sizeof({cname})); }}")
            elif head in ("texture_2d", "texture_storage_2d"):
                o.append(f"void bind_{name}(void* p, int w, int h, int format) {{ {cname}.bind(p, w, h, format); }}")
            elif head == "binding_array" and r[2][1] == "texture_2d":
                o.append(f"void bind_{name}(int index, void* p, int w, int h, int format) {{ {cname}.at(index).bind(p, w, h, format); }}")
            elif head == "binding_array" and r[2][1] == "sampler":
                o.append(f"void bind_{name}(int index, int mode_u, int mode_v, int linear) {{ {cname}.at(index).set(mode_u, mode_v, linear); }}")
            elif head == "sampler":
                o.append(f"void bind_{name}(int mode_u, int mode_v, int linear) {{ {cname}.set(mode_u, mode_v, linear); }}")
        for name, ret, params, body, attrs in self.functions:
            if not any(a == "compute" for a, _ in attrs):
                continue
            wg = [a for a in attrs if a[0] == "workgroup_size"][0][1].split(",")
            wg = [int(re.sub(r"[ui]$", "", x)) for x in wg] + [1, 1]
            args = []
            for pname, pty, pattrs in params:
                b = [a for a in pattrs if a[0] == "builtin"][0][1]
                args.append({"global_invocation_id": "wgsl_global_id()", "local_invocation_id": "wgsl_local_id()", "workgroup_id": "wgsl_group_id()",
                             "num_workgroups": "wgsl_num_groups()", "local_invocation_index": "wgsl_local_index()"}[b])
            coop = "true" if "workgroupBarrier" in body else "false"      # invocations must run concurrently only where they meet at a barrier
            o.append(f"void run_{name}(unsigned gx, unsigned gy, unsigned gz) {{ wgsl_dispatch<{coop}>(gx, gy, gz, {wg[0]}, {wg[1]}, {wg[2]}, "
                     f"[]() {{ {self.ident(name)}({', '.join(args)}); }}); }}")
        o.append("}")
        return o


def translate(path, defs, banner=None):
    text = preprocess(path, set(defs))
    tr = Translator(text)
    tr.module()
    return tr.emit(banner or f"{os.path.basename(path)} with shader defs {sorted(defs)}")


if __name__ == "__main__":
    src, out = sys.argv[1], sys.argv[2]
    open(out, "w").write(translate(src, sys.argv[3:]))
