"""The C-ABI library loads on a CPU-only box, exports every symbol the headers declare, and refuses to run without a
CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from bevy_hikari_b200 import _ffi, plugin
from tests.conftest import ROOT, has_gpu


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(hk_[a-z0-9_]+|hikari_[a-z0-9_]+)\s*\(", src))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared("hikari_b200.h") | declared("hikari_host.h")
    assert len(names) >= 40
    lib = C.CDLL(_ffi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert names == set(_ffi.SYMBOLS), names ^ set(_ffi.SYMBOLS)


def test_version_and_graph_name():
    assert b"sm_100a" in _ffi.lib().hk_version()
    assert plugin.graph_name() == "hikari"                      # src/lib.rs:44


def test_settings_defaults_are_the_reference_defaults():
    s = plugin.HikariSettings()                                  # src/lib.rs:435-455
    assert (s.direct_validate_interval, s.emissive_validate_interval) == (3, 5)
    assert (s.max_temporal_reuse_count, s.max_spatial_reuse_count) == (50, 800)
    assert s.max_reservoir_lifetime == 100.0 and abs(s.solar_angle - 0.046) < 1e-9
    assert s.indirect_bounces == 1 and s.max_indirect_luminance == 10.0
    assert [round(c, 6) for c in s.clear_color] == [0.4, 0.4, 0.4, 1.0]
    assert (s.temporal_reuse, s.emissive_spatial_reuse, s.indirect_spatial_reuse, s.denoise) == (1, 0, 1, 1)
    assert s.taa == plugin.TAA_JASMINE and s.upscale_kind == plugin.UPSCALE_SMAA_TU4X and s.upscale_ratio == 2.0
    s.upscale_ratio = 5.0
    assert _ffi.lib().hikari_upscale_ratio(C.byref(s)) == 2.0   # clamp(1, 2), src/lib.rs:501-505
    s.upscale_ratio = 0.5
    assert _ffi.lib().hikari_upscale_ratio(C.byref(s)) == 1.0


def test_frame_uniform_extraction():
    from bevy_hikari_b200 import scenes
    sc = scenes.cornell()
    view, pview, lights = sc.view_inputs(64, 64)
    s = plugin.HikariSettings(indirect_bounces=3, denoise=0, taa=plugin.TAA_NONE, upscale_ratio=1.0)
    inp = plugin.make_frame_inputs(s, 7, view, pview, lights)
    f = inp.frame
    assert f.number == 7 and f.indirect_bounces == 3 and f.upscale_ratio == 1.0 and inp.denoise == 0 and inp.taa_jitter == 0
    assert [f.kernel[1][1], f.kernel[0][0], f.kernel[0][1]] == [0.25, 0.0625, 0.125]           # view.rs:125-129
    assert abs(f.halton[1][1] - 0.666667) < 1e-7 and f.halton[7][2] == 0.9375                  # view.rs:130-139
    assert f.max_temporal_reuse_count == 50 and f.direct_validate_interval == 3


@pytest.mark.skipif(has_gpu(), reason="CPU-only behaviour")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(_ffi.HikariError, match="no CUDA device"):
        plugin.HikariPlugin(32, 32)
