#!/usr/bin/env python
"""TEST INFRASTRUCTURE — oracle/_ref: every library that the reference's own shader text compiles to here, built where the sources lie
(/root/reference/src/shaders/*.wgsl through wgsl2cpp.py; src/shaders/fsr/source.zip through glsl2cpp.py; prepass.wgsl + the raster
harness), into oracle/_ref/wgsl/ (git-ignored; it travels to the GPU box, where nothing reads it: the device tests use the committed
fixtures).  Called by __graft_entry__.build() when /root/reference is present; the tests build the same libraries on demand (same
content-hashed file names, so whichever comes first does the work)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def jobs():
    import run_reference as R
    import raster_prepass as RP
    out = []
    for defs in R.WgslReference.LIGHT_VARIANTS.values():
        for no_tex in ([], ["NO_TEXTURE"]):                                  # light.rs:141-143
            out.append(lambda d=no_tex + defs: R.build("light", d))
    out.append(lambda: R.build("denoise", ["DENOISE_LEVEL_0"]))
    for lvl in range(4):
        for ff in (False, True):
            out.append(lambda l=lvl, f=ff: R.build("denoise", [f"DENOISE_LEVEL_{l}"] + (["FIREFLY_FILTERING"] if f else [])))
    for shader in ("tone_mapping", "smaa", "taa"):
        out.append(lambda s=shader: R.build(s, []))
    for define in ("SAMPLE_EASU", "SAMPLE_RCAS"):
        out.append(lambda d=define: R.build_fsr(d))
    for defs in ([], ["TEMPORAL_ANTI_ALIASING"]):
        out.append(lambda d=defs: RP.build(d))
    return out


def build_all():
    import run_reference as R
    if not R.available():
        return 0
    work = jobs()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda f: f(), work))
    return len(work)


if __name__ == "__main__":
    print(f"{build_all()} libraries in oracle/_ref/wgsl")
