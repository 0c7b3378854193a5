"""The Rust side of the boundary (rust/hikari-b200-sys) cannot be compiled in this container — no Rust toolchain — so what is
checked is everything that can be checked without one: the committed declarations are exactly what tools/gen_rust_sys.py
generates from include/*.h today; every function the library exports is declared and every declared function is exported; the
struct sizes written into the Rust compile-time assertions are the sizes ctypes computes for the Python mirrors of the same
structs; and the hand-written plugin crate uses only names the generated crate declares."""
import os
import re
import subprocess
import sys

from tests.conftest import ROOT

LIB_RS = os.path.join(ROOT, "rust", "hikari-b200-sys", "src", "lib.rs")


def test_generated_declarations_are_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout


def test_extern_block_matches_exported_symbols():
    from bevy_hikari_b200 import _ffi
    text = open(LIB_RS).read()
    declared = set(re.findall(r"pub fn (hk_\w+)\(", text))
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.split()[-1].startswith("hk_")}
    assert declared == exported, (sorted(declared - exported), sorted(exported - declared))
    assert len(declared) >= 38


def test_struct_sizes_in_the_rust_assertions_match_the_python_mirrors():
    import ctypes as C
    from bevy_hikari_b200 import layout as L
    sizes = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"size_of::<(\w+)>\(\) == (\d+)", open(LIB_RS).read()))
    for name, ctype in (("hk_frame_uniform", L.FrameUniform), ("hk_view", L.View), ("hk_previous_view", L.PreviousView), ("hk_lights", L.Lights),
                        ("hk_frame_inputs", L.FrameInputs), ("hk_frame_stats", L.FrameStats), ("hk_texture_desc", L.TextureDesc),
                        ("hk_scene_desc", L.SceneDesc)):
        assert sizes[name] == C.sizeof(ctype), (name, sizes[name], C.sizeof(ctype))
    for name, dtype in (("hk_node", L.NODE), ("hk_primitive", L.PRIMITIVE), ("hk_vertex", L.VERTEX), ("hk_instance", L.INSTANCE),
                        ("hk_material", L.MATERIAL), ("hk_alias_entry", L.ALIAS_ENTRY), ("hk_emissive", L.EMISSIVE),
                        ("hk_packed_reservoir", L.PACKED_RESERVOIR), ("hk_ray", L.RAY), ("hk_hit", L.HIT)):
        assert sizes[name] == dtype.itemsize, (name, sizes[name], dtype.itemsize)


def test_plugin_crate_uses_only_declared_names():
    text = open(LIB_RS).read()
    declared = set(re.findall(r"pub (?:fn|struct|const) (\w+)", text))
    plugin = open(os.path.join(ROOT, "rust", "bevy-hikari-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"(?<!:)\bffi::(\w+)", plugin))          # `ffi::name`, not `std::ffi::name`
    assert used and used <= declared, sorted(used - declared)
    fields = set(re.findall(r"^\s+pub (\w+):", text, re.M))
    for struct_literal in ("hk_frame_inputs", "hk_view", "hk_lights"):
        body = re.search(r"ffi::%s \{(.*?)\n        \}|ffi::%s \{(.*?)\n    \}" % (struct_literal, struct_literal), plugin, re.S)
        assert body, struct_literal
        names = set(re.findall(r"^\s+(\w+)(?::(?!:)|,)", (body.group(1) or body.group(2)), re.M))     # `field:` / `field,`, not `path::`
        assert names and names <= fields, (struct_literal, sorted(names - fields))
