"""Known-answer tests of include/hk_math.h (through the oracle build of the same header): the arithmetic both the CUDA
kernels and the oracle use.  Expected values come from numpy / the WGSL spec formulas, computed independently here."""
import ctypes as C

import numpy as np

from oracle import oracle


def ulp_diff(a, b):
    a, b = np.float32(a), np.float32(b)
    ia, ib = int(a.view(np.int32)), int(b.view(np.int32))
    return abs(ia - ib)


def test_exp2_and_exp_accuracy():
    lib = oracle.lib()
    xs = np.concatenate([np.linspace(-30, 30, 4001), np.linspace(-126, -100, 101), [0.0, 1.0, -1.0, 127.5]]).astype(np.float32)
    worst = 0
    for x in xs:
        worst = max(worst, ulp_diff(lib.hko_math_exp2(float(x)), np.exp2(np.float64(x))))
    assert worst <= 2, worst
    for x in np.linspace(-20, 5, 2001).astype(np.float32):
        # WGSL accuracy requirement for exp(x) is 3 + 2|x| ULP; exp_(x) = exp2_(x * log2(e)) stays far inside it
        assert ulp_diff(lib.hko_math_exp(float(x)), np.exp(np.float64(x))) <= 3 + 2 * abs(float(x))
    assert lib.hko_math_exp2(-200.0) == 0.0 and lib.hko_math_exp2(200.0) == float("inf")
    assert lib.hko_math_exp2(0.0) == 1.0 and lib.hko_math_exp2(10.0) == 1024.0


def test_sincos_accuracy_over_the_range_the_path_uses():
    lib = oracle.lib()
    s, c = C.c_float(), C.c_float()
    for x in np.linspace(0.0, 2 * np.pi, 5001).astype(np.float32):
        lib.hko_math_sincos(float(x), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(np.float64(x))) < 3e-7 and abs(c.value - np.cos(np.float64(x))) < 3e-7
    lib.hko_math_sincos(0.0, C.byref(s), C.byref(c))
    assert (s.value, c.value) == (0.0, 1.0)


def test_f16_pack_is_ieee_round_to_nearest_even():
    lib = oracle.lib()
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(0, 100, 2000), rng.normal(0, 1e-5, 500), [0.0, -0.0, 65504.0, 65520.0, 1e9, -1e9, 5.96e-8,
                                                                               2.98e-8, 1.0009765625, 1.00048828125]])
    for v in vals.astype(np.float32):
        with np.errstate(over="ignore"):
            want = int(np.float32(v).astype(np.float16).view(np.uint16))
        got = lib.hko_math_pack2x16float(float(v), 0.0) & 0xFFFF
        assert got == want, (v, hex(got), hex(want))
        back = lib.hko_math_f16_to_f32(got)
        wantf = float(np.uint16(want).view(np.float16).astype(np.float32))
        assert back == wantf or (np.isnan(back) and np.isnan(wantf))


def test_snorm_unorm_pack_follow_the_wgsl_formulas():
    lib = oracle.lib()
    rng = np.random.default_rng(1)
    for v in list(rng.uniform(-1.5, 1.5, 500)) + [0.0, 1.0, -1.0, 0.5, -0.5, 1 / 254.0]:
        v = np.float32(v)
        want = int(np.floor(np.float32(0.5) + np.float32(127.0) * min(np.float32(1), max(np.float32(-1), v)))) & 0xFF
        assert lib.hko_math_pack4x8snorm(float(v), 0, 0, 0) & 0xFF == want
        want = int(np.floor(np.float32(0.5) + np.float32(65535.0) * min(np.float32(1), max(np.float32(0), v))))
        assert lib.hko_math_pack2x16unorm(float(v), 0) & 0xFFFF == want
    # NaN follows IEEE minNum/maxNum: clamps to the lower bound
    assert lib.hko_math_pack4x8snorm(float("nan"), 0, 0, 0) & 0xFF == 0x81


def test_hash_matches_a_python_restatement_of_utils_wgsl():
    def h(v):  # utils.wgsl:15-24
        s = v & 0xFFFFFFFF
        s ^= 2747636419
        s = (s * 2654435769) & 0xFFFFFFFF
        s ^= s >> 16
        s = (s * 2654435769) & 0xFFFFFFFF
        s ^= s >> 16
        s = (s * 2654435769) & 0xFFFFFFFF
        return s
    lib = oracle.lib()
    for v in [0, 1, 2, 3, 64, 12345, 0xFFFFFFFF, 0x80000000]:
        assert lib.hko_math_hash(v) == h(v)


def test_normal_basis_is_orthonormal_with_n_as_third_column():
    lib = oracle.lib()
    rng = np.random.default_rng(2)
    for _ in range(200):
        n = rng.normal(size=3)
        n = (n / np.linalg.norm(n)).astype(np.float32)
        out = np.zeros(9, np.float32)
        lib.hko_math_normal_basis(n.ctypes.data, out.ctypes.data)
        m = out.reshape(3, 3)  # rows = columns t, b, n
        assert np.allclose(m[2], n)
        assert np.allclose(m @ m.T, np.eye(3), atol=2e-6)


def test_reservoir_pack_unpack_is_idempotent_after_one_round_trip():
    """pack(unpack(p)) is not the identity for arbitrary bits (normals get re-normalised) but it is a projection."""
    from bevy_hikari_b200 import layout as L
    lib = oracle.lib()
    rng = np.random.default_rng(3)
    p = np.zeros(64, L.PACKED_RESERVOIR)
    p["visible_position"] = rng.normal(size=(64, 4)).astype(np.float32)
    p["sample_position"][:, :3] = rng.normal(size=(64, 3)).astype(np.float32)
    p["sample_position"][:, 3] = rng.integers(0, 8, 64).astype(np.float32)
    for name in ("radiance", "random", "reservoir"):
        h = np.abs(rng.normal(size=(64, 4))).astype(np.float16)
        p[name] = h.view(np.uint32).reshape(64, 2)
    p["visible_normal"] = rng.integers(0, 2 ** 32, 64, dtype=np.uint64).astype(np.uint32)
    p["sample_normal"] = rng.integers(0, 2 ** 32, 64, dtype=np.uint64).astype(np.uint32)
    once, twice = np.zeros_like(p), np.zeros_like(p)
    for i in range(64):
        lib.hko_pack_reservoir_roundtrip(p[i:i + 1].ctypes.data, once[i:i + 1].ctypes.data)
        lib.hko_pack_reservoir_roundtrip(once[i:i + 1].ctypes.data, twice[i:i + 1].ctypes.data)
    # a normal quantised twice may move by one snorm8 step; everything else is stable
    for name in ("radiance", "random", "visible_position", "sample_position", "reservoir"):
        assert once[name].tobytes() == twice[name].tobytes(), name
    d = (np.ascontiguousarray(once["visible_normal"]).view(np.int8).astype(int) -
         np.ascontiguousarray(twice["visible_normal"]).view(np.int8).astype(int))
    assert np.abs(d).max() <= 1


def test_fast_division_by_constants_is_exactly_ieee_division_for_every_input():
    """unsnorm8 / unorm8 / unorm16 use a 3-instruction correctly-rounded quotient (hk::div_const); exhaustive check."""
    lib = oracle.lib()
    for b in range(256):
        i = b - 256 if b >= 128 else b
        want = max(np.float32(i) / np.float32(127.0), np.float32(-1.0))
        assert np.float32(lib.hko_math_unsnorm8(b)).view(np.uint32) == np.float32(want).view(np.uint32), b
        assert np.float32(lib.hko_math_unorm8(b)).view(np.uint32) == (np.float32(b) / np.float32(255.0)).view(np.uint32), b
    got = np.array([lib.hko_math_unorm16(u) for u in range(65536)], np.float32)
    want = np.arange(65536, dtype=np.float32) / np.float32(65535.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_brdf_matches_float64_restatement_of_the_published_formulas():
    """lit() (light.wgsl:796-818) = (Fr + Fd) * radiance * NoL with bevy_pbr 0.9's pbr_lighting functions, which are
    Filament's: D_GGX, V_SmithGGXCorrelated, F_Schlick with f90 = saturate(dot(f0, 50 * 0.33)), Fd_Burley; EnvBRDFApprox is
    Karis' mobile approximation.  Restated here in float64 from the published formulas, independent of include/hk_math.h."""
    lib = oracle.lib()
    rng = np.random.default_rng(3)
    f3 = lambda a: np.ascontiguousarray(a, np.float32)

    def unit(v):
        return v / np.linalg.norm(v)

    worst = 0.0
    for _ in range(400):
        N = unit(rng.normal(size=3))
        V = unit(N + 0.9 * rng.normal(size=3))
        Lv = unit(N + 0.9 * rng.normal(size=3))
        rough = float(np.float32(rng.uniform(0.089, 1.0) ** 2))
        F0 = rng.uniform(0.02, 0.9, 3); diffuse = rng.uniform(0, 1, 3); radiance = rng.uniform(0, 20, 3)
        args = [f3(radiance), f3(diffuse), f3(F0), f3(Lv), f3(N), f3(V)]
        radiance, diffuse, F0, Lv, N, V = [a.astype(np.float64) for a in args]
        out = np.zeros(3, np.float32)
        lib.hko_math_lit(args[0].ctypes.data, args[1].ctypes.data, C.c_float(rough), args[2].ctypes.data, args[3].ctypes.data,
                         args[4].ctypes.data, args[5].ctypes.data, out.ctypes.data)
        H = unit(Lv + V)
        sat = lambda x: min(max(x, 0.0), 1.0)
        NoL, NoH, LoH, NoV = sat(N @ Lv), sat(N @ H), sat(Lv @ H), max(N @ V, 1e-4)
        f90 = 0.5 + 2.0 * rough * LoH * LoH
        fd = (1 + (f90 - 1) * (1 - NoL) ** 5) * (1 + (f90 - 1) * (1 - NoV) ** 5) / np.pi
        a = NoH * rough
        k = rough / (1.0 - NoH * NoH + a * a)
        D = k * k / np.pi
        a2 = rough * rough
        Vis = 0.5 / (NoL * np.sqrt((NoV - a2 * NoV) * NoV + a2) + NoV * np.sqrt((NoL - a2 * NoL) * NoL + a2))
        F = F0 + (sat(F0.sum() * 50.0 * 0.33) - F0) * (1 - LoH) ** 5
        expect = (D * Vis * F + diffuse * fd) * radiance * NoL
        if np.abs(expect).max() > 1e-6:
            worst = max(worst, float(np.abs(out - expect).max() / max(np.abs(expect).max(), 1e-3)))
    assert worst < 2e-4, worst

    worst = 0.0
    for _ in range(400):
        f0 = rng.uniform(0.0, 1.0, 3); pr = float(rng.uniform(0.0, 1.0)); nov = float(rng.uniform(1e-4, 1.0))
        out = np.zeros(3, np.float32)
        f0_32 = f3(f0)
        lib.hko_math_env_brdf_approx(f0_32.ctypes.data, C.c_float(pr), C.c_float(nov), out.ctypes.data)
        pr, nov = float(np.float32(pr)), float(np.float32(nov))
        r = pr * np.array([-1.0, -0.0275, -0.572, 0.022]) + np.array([1.0, 0.0425, 1.04, -0.04])
        a004 = min(r[0] * r[0], 2.0 ** (-9.28 * nov)) * r[0] + r[1]
        ab = np.array([-1.04, 1.04]) * a004 + r[2:]
        expect = f0_32.astype(np.float64) * ab[0] + ab[1]
        worst = max(worst, float(np.abs(out - expect).max()))
    assert worst < 2e-5, worst
    assert lib.hko_math_perceptual_roughness_to_roughness(C.c_float(0.5)) == np.float32(0.25)
    assert lib.hko_math_perceptual_roughness_to_roughness(C.c_float(0.01)) == np.float32(np.float32(0.089) * np.float32(0.089))
