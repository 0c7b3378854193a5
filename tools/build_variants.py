#!/usr/bin/env python
"""Build tuning variants of libhikari_b200.so here (nvcc cross-compiles without a GPU) into bevy_hikari_b200/variants/,
so that one gpurun call can time them all:  python bench.py --lib bevy_hikari_b200/variants/<name>.so ...
usage: tools/build_variants.py name='-DFLAG=.. -DFLAG2=..' [name2=...]
Variant objects live in their own scratch directories (bevy_hikari_b200/_build/variant_<name>/): the default build is untouched."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "bevy_hikari_b200", "variants")
os.makedirs(out, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    env = dict(os.environ, HK_NVCC_EXTRA=flags)
    code = ("import sys; sys.path.insert(0, %r); from bevy_hikari_b200 import build; build.build(out=%r)"
            % (ROOT, os.path.join(out, name + ".so")))
    procs.append((name, flags, subprocess.Popen([sys.executable, "-c", code], env=env)))
for name, flags, p in procs:
    if p.wait() != 0:
        raise SystemExit(f"variant {name} failed to build")
    print("built", name, flags)
