#!/usr/bin/env python
"""Build tuning variants of libhikari_b200.so here (nvcc cross-compiles without a GPU) into bevy_hikari_b200/variants/,
so that one gpurun call can time them all:  python bench.py --lib bevy_hikari_b200/variants/<name>.so ...
usage: tools/build_variants.py name='-DFLAG=.. -DFLAG2=..' [name2=...]"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "bevy_hikari_b200", "variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    env = dict(os.environ, HK_NVCC_EXTRA=flags)
    subprocess.run([sys.executable, os.path.join(ROOT, "bevy_hikari_b200", "build.py")], env=env, check=True, stdout=subprocess.DEVNULL)
    shutil.copy(os.path.join(ROOT, "bevy_hikari_b200", "libhikari_b200.so"), os.path.join(out, name + ".so"))
    print("built", name, flags)
# restore the default build
subprocess.run([sys.executable, os.path.join(ROOT, "bevy_hikari_b200", "build.py"), "--force"], check=True, stdout=subprocess.DEVNULL)
