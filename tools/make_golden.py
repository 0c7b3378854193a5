#!/usr/bin/env python
"""Generate tests/golden/cornell_48x48_cfg2_frames1-6.npz and cornell_48x40_ratio1.5_smaa_taa_frames1-6.npz FROM THE ORACLE (not from the reference: the reference cannot
run here and ships no golden vectors).  It is a regression pin of oracle/hk_oracle.cpp + include/hk_math.h."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevy_hikari_b200 import layout as L  # noqa: E402
from tests.conftest import Bench  # noqa: E402

b = Bench("cornell", 48, 48, config="cornell_1080p")
orc = b.oracle()
for f in range(1, 7):
    orc.render_frame(b.inputs(f))
out = {}
for name, which in (("tone_mapped", L.OUT_TONE_MAPPED), ("position", L.OUT_GBUFFER_POSITION), ("instance_material", L.OUT_GBUFFER_INSTANCE_MATERIAL),
                    ("reservoir9", L.OUT_RESERVOIR_0 + 9), ("render_indirect", L.OUT_RENDER_INDIRECT)):
    out[name] = np.ascontiguousarray(orc.readback(which)).view(np.uint8).reshape(-1)
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cornell_48x48_cfg2_frames1-6.npz"), **out)
print({k: v.size for k, v in out.items()})


# second fixture: render at 1/1.5 of the output size, moving camera, smaa_tu4x + taa_jasmine after tone mapping
from bevy_hikari_b200 import plugin  # noqa: E402

b = Bench("cornell", 48, 40, config="cornell_1080p", taa=plugin.TAA_JASMINE, upscale_ratio=1.5)
orc = b.oracle()
for f in range(1, 7):
    inp = b.moving_inputs(f)
    inp.temporal_upscalers = 1
    orc.render_frame(inp)
out = {}
for name, which in (("tone_mapped", L.OUT_TONE_MAPPED), ("upscaled", L.OUT_UPSCALED), ("taa", L.OUT_TAA), ("reservoir9", L.OUT_RESERVOIR_0 + 9)):
    out[name] = np.ascontiguousarray(orc.readback(which)).view(np.uint8).reshape(-1)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cornell_48x40_ratio1.5_smaa_taa_frames1-6.npz"), **out)
print({k: v.size for k, v in out.items()})
