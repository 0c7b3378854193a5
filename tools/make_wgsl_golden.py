#!/usr/bin/env python
"""Golden vectors from the reference's OWN shaders.  Runs every sequence of tests/wgsl_cases.py through the reference's WGSL
(/root/reference/src/shaders/{light,denoise,tone_mapping,smaa,taa}.wgsl translated to C++ by oracle/wgsl/wgsl2cpp.py and executed on the CPU by
oracle/wgsl/run_reference.py, wired as src/light.rs and src/post_process.rs wire the passes) and writes per-frame, per-plane SHA-256
digests of every buffer and texture the passes produce to tests/golden/wgsl_<case>.npz.

Only in the build container (needs /root/reference + g++).  The G-buffer of each frame is the oracle's (the reference rasterises it:
a render pipeline, not part of the translated compute path); everything downstream — albedo, both direct_lit pipelines, both
spatial_reuse pipelines, indirect_lit_ambient (single / multiple bounces), demodulation, the four denoise levels with and without
firefly filtering, tone mapping, SMAA TU4x (+ extrapolation) and TAA, over free-running sequences with validation frames, camera motion and moving instances — is computed
by the reference's shader text from its own state of the previous frames."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "wgsl"))

from bevy_hikari_b200 import layout as L  # noqa: E402
from bevy_hikari_b200 import plugin  # noqa: E402
from tests import wgsl_cases as WC  # noqa: E402
import run_reference as R  # noqa: E402


def reference_planes(ref, bench):
    out = {"albedo": ref.albedo, "tone_mapped": ref.tone_mapped}
    for i in range(3):
        out[f"render{i}"], out[f"variance{i}"], out[f"denoised{i}"] = ref.render[i], ref.variance[i], ref.denoise_render[i]
    for i in range(10):
        out[f"reservoir{i}"] = ref.reservoir[i][:ref.rw * ref.rh]      # the records the passes index (render size)
    if ref.upscale_output is not None:
        out["upscaled"] = ref.upscale_output
    if ref.taa_output is not None:
        out["taa"] = ref.taa_output[ref.head]
    if ref.fsr_output is not None:
        out["fsr_easu"], out["fsr_rcas"] = ref.fsr_output
    return out


def run_case(case):
    bench = WC.make_bench(case)
    w, h = bench.width, bench.height
    textures = [(np.ascontiguousarray(t["rgba"]), t["address_mode_u"], t["address_mode_v"], t["filter_linear"], t["srgb"]) for t in bench.scene.textures]
    ref = R.WgslReference(bench.world.buffers(), textures, plugin.load_noise(), w, h, bench.settings.upscale_ratio)
    orc = bench.oracle()
    frames = WC.CASES[case][3]
    digests = {}
    for f in range(1, frames + 1):
        if WC.animate(bench, case, f):
            orc.update_instances_desc(bench.world.scene_desc())
            ref.scene = {k: np.ascontiguousarray(v) for k, v in bench.world.buffers().items()}
        inp = WC.frame_inputs(bench, case, f)
        orc.prepass(inp)                                    # the raster prepass: G-buffer only
        ref.set_gbuffer(*[np.ascontiguousarray(orc.readback(k)) for k in WC.GBUFFER])
        ref.light_node(inp)
        ref.post_process_node(inp, bool(bench.settings.denoise))
        smaa, taa = WC.upscalers_of(case, bench)
        if smaa or taa:
            ref.upscale_node(inp, smaa, taa)
        if WC.fsr_of(case, bench):
            ref.fsr_node(inp, taa, bench.settings.upscale_sharpness)
        planes = reference_planes(ref, bench)
        for name, _ in WC.planes_of(case, bench):
            digests[f"f{f}_{name}"] = WC.digest(planes[name])
    return digests, ref.tone_mapped.copy()


def run_prepass_case(case):
    """the G-buffer of the case's compared frame as prepass.wgsl rasterises it (oracle/wgsl/raster_prepass.py)"""
    import raster_prepass as RP
    taa = WC.PREPASS_CASES[case][6]
    rp = RP.RasterPrepass(taa=taa)
    out = None
    for bench, inp, previous_models, _ in WC.prepass_sequence(case):
        out = rp.render(inp, bench.width, bench.height, bench.scene.meshes, bench.scene.inst_mesh, bench.world.buffers()["instances"], previous_models)
    return out


def main():
    if not R.available():
        raise SystemExit("needs /root/reference and g++ (build container only)")
    out_dir = os.path.join(ROOT, "tests", "golden")
    if sys.argv[1:2] == ["--prepass"]:
        for case in (sys.argv[2:] or [c for c, v in WC.PREPASS_CASES.items() if v[7]]):
            r = run_prepass_case(case)
            np.savez_compressed(os.path.join(out_dir, f"wgsl_prepass_{case}.npz"), **{k: r[k] for k, _ in WC.PREPASS_PLANES})
            print(f"prepass {case}: {r['fragments']} fragments shaded, {int((r['position'][..., 3] > 0).sum())} pixels covered")
        return
    for case in (sys.argv[1:] or WC.CASES):
        digests, last = run_case(case)
        names = sorted(digests)
        np.savez_compressed(os.path.join(out_dir, f"wgsl_{case}.npz"), names=np.array(names), digests=np.array([digests[n] for n in names]),
                            last_tone_mapped=last)
        print(f"{case}: {len(names)} plane digests over {WC.CASES[case][3]} frames")


if __name__ == "__main__":
    main()
