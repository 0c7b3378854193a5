"""FSR 1.0 (Upscale::Fsr1) in the oracle, pinned from the outside (CPU only).  The reference ships EASU / RCAS as SPIR-V
blobs; their sources (src/shaders/fsr/source.zip: FSR_Pass.glsl, ffx_fsr1.h, ffx_a.h) are what oracle/hk_oracle.cpp restates
per pixel.  Here a SECOND, independent restatement — whole-image numpy float32 arithmetic written from the same sources —
must agree with the oracle bit for bit on random images, and the filter's defining properties are checked directly:
flat images pass through, EASU never leaves the range of the 2x2 texels around the sample (de-ringing), RCAS keeps a pixel
whose ring is flat, borders read zero outside (texelFetch) and clamp inside EASU (clamp-to-edge sampler)."""
import numpy as np
import pytest

from bevy_hikari_b200 import layout as L
from bevy_hikari_b200 import plugin
from tests.conftest import Bench

F = np.float32


def f2u(a): return np.asarray(a, F).view(np.uint32)
def u2f(a): return np.asarray(a, np.uint32).view(F)
def rcp_lo(a): return u2f(np.uint32(0x7ef07ebb) - f2u(a))          # APrxLoRcpF1, ffx_a.h:1843
def rsq_lo(a): return u2f(np.uint32(0x5f347d74) - (f2u(a) >> np.uint32(1)))   # APrxLoRsqF1, :1845


def rcp_med(a):                                                    # APrxMedRcpF1, :1844
    b = u2f(np.uint32(0x7ef19fff) - f2u(a))
    return b * (-b * a + F(2.0))


def sat(a): return np.minimum(np.maximum(a, F(0.0)), F(1.0))


def easu_numpy(img, out_w, out_h):
    """img: (h, w, 3) float32 (already f16-representable).  FsrEasuF, ffx_fsr1.h:315-437, on every output pixel at once."""
    h, w, _ = img.shape
    with np.errstate(all="ignore"):
        sx = F(w) * (F(1.0) / F(out_w)); sy = F(h) * (F(1.0) / F(out_h))
        ox = F(0.5) * F(w) * (F(1.0) / F(out_w)) - F(0.5); oy = F(0.5) * F(h) * (F(1.0) / F(out_h)) - F(0.5)
        X, Y = np.meshgrid(np.arange(out_w, dtype=F), np.arange(out_h, dtype=F))
        ppx = X * sx + ox; ppy = Y * sy + oy
        fpx = np.floor(ppx); fpy = np.floor(ppy)
        ppx = ppx - fpx; ppy = ppy - fpy
        ix = fpx.astype(np.int64); iy = fpy.astype(np.int64)

        def T(dx, dy):
            return img[np.clip(iy + dy, 0, h - 1), np.clip(ix + dx, 0, w - 1)]
        t = {k: T(*o) for k, o in dict(b=(0, -1), c=(1, -1), e=(-1, 0), f=(0, 0), g=(1, 0), h=(2, 0), i=(-1, 1), j=(0, 1),
                                       k=(1, 1), l=(2, 1), n=(0, 2), o=(1, 2)).items()}
        lum = {k: v[..., 2] * F(0.5) + (v[..., 0] * F(0.5) + v[..., 1]) for k, v in t.items()}
        dirx = np.zeros_like(ppx); diry = np.zeros_like(ppx); ln = np.zeros_like(ppx)
        one = F(1.0)
        for wgt, (a, b, c, d, e) in (((one - ppx) * (one - ppy), "befgj"), (ppx * (one - ppy), "cfghk"),
                                     ((one - ppx) * ppy, "fijkn"), (ppx * ppy, "gjklo")):
            lA, lB, lC, lD, lE = lum[a], lum[b], lum[c], lum[d], lum[e]
            lenx = rcp_lo(np.maximum(np.abs(lD - lC), np.abs(lC - lB)))
            dX = lD - lB
            dirx = dirx + dX * wgt
            lenx = sat(np.abs(dX) * lenx); lenx = lenx * lenx
            ln = ln + lenx * wgt
            leny = rcp_lo(np.maximum(np.abs(lE - lC), np.abs(lC - lA)))
            dY = lE - lA
            diry = diry + dY * wgt
            leny = sat(np.abs(dY) * leny); leny = leny * leny
            ln = ln + leny * wgt
        dirr = dirx * dirx + diry * diry
        zro = dirr < F(1.0 / 32768.0)
        dirr = np.where(zro, one, rsq_lo(dirr))
        dirx = np.where(zro, one, dirx) * dirr
        diry = diry * dirr
        ln = ln * F(0.5); ln = ln * ln
        stretch = (dirx * dirx + diry * diry) * rcp_lo(np.maximum(np.abs(dirx), np.abs(diry)))
        len2x = one + (stretch - one) * ln
        len2y = one + F(-0.5) * ln
        lob = F(0.5) + F((1.0 / 4.0 - 0.04) - 0.5) * ln
        clp = rcp_lo(lob)
        aC = np.zeros(ppx.shape + (3,), F); aW = np.zeros_like(ppx)
        for k, (dx, dy) in (("b", (0, -1)), ("c", (1, -1)), ("i", (-1, 1)), ("j", (0, 1)), ("f", (0, 0)), ("e", (-1, 0)),
                            ("k", (1, 1)), ("l", (2, 1)), ("h", (2, 0)), ("g", (1, 0)), ("o", (1, 2)), ("n", (0, 2))):
            offx = F(dx) - ppx; offy = F(dy) - ppy
            vx = ((offx * dirx) + (offy * diry)) * len2x
            vy = ((offx * (-diry)) + (offy * dirx)) * len2y
            d2 = np.minimum(vx * vx + vy * vy, clp)
            wB = F(2.0 / 5.0) * d2 + F(-1.0)
            wA = lob * d2 + F(-1.0)
            wB = wB * wB; wA = wA * wA
            wB = F(25.0 / 16.0) * wB + F(-(25.0 / 16.0 - 1.0))
            wt = wB * wA
            aC = aC + t[k] * wt[..., None]
            aW = aW + wt
        mn = np.minimum(np.minimum(t["f"], np.minimum(t["g"], t["j"])), t["k"])
        mx = np.maximum(np.maximum(t["f"], np.maximum(t["g"], t["j"])), t["k"])
        pix = np.minimum(mx, np.maximum(mn, aC * (one / aW)[..., None]))
    return pix, mn, mx


def rcas_numpy(img, sharpness):
    """img: (h, w, 3) float32.  FsrRcasF, ffx_fsr1.h:684-772 (no FSR_RCAS_DENOISE); texelFetch outside the image = 0."""
    h, w, _ = img.shape
    pad = np.zeros((h + 2, w + 2, 3), F)
    pad[1:-1, 1:-1] = img
    b, d, e, f, hh = pad[:-2, 1:-1], pad[1:-1, :-2], pad[1:-1, 1:-1], pad[1:-1, 2:], pad[2:, 1:-1]
    with np.errstate(all="ignore"):
        mn4 = np.fmin(np.fmin(b, np.fmin(d, f)), hh)
        mx4 = np.fmax(np.fmax(b, np.fmax(d, f)), hh)
        hit_min = np.fmin(mn4, e) * (F(1.0) / (F(4.0) * mx4))
        hit_max = (F(1.0) - np.fmax(mx4, e)) * (F(1.0) / (F(4.0) * mn4 + F(-4.0)))
        lobe_c = np.fmax(-hit_min, hit_max)
        lobe = np.fmax(F(-(0.25 - 1.0 / 16.0)), np.fmin(np.fmax(lobe_c[..., 0], np.fmax(lobe_c[..., 1], lobe_c[..., 2])), F(0.0)))
        lobe = (lobe * F(2.0) ** F(-sharpness))[..., None]
        rcp = rcp_med(F(4.0) * lobe + F(1.0))
        return (lobe * b + lobe * d + lobe * hh + lobe * f + e) * rcp


def run_fsr(image_f16, out_w, out_h, ratio, sharpness):
    """EASU and RCAS of the oracle on a given render-size image: one frame to size everything, then the two passes alone"""
    b = Bench("cornell", out_w, out_h, config="cornell_256", taa=plugin.TAA_NONE, upscale_kind=plugin.UPSCALE_FSR1,
              upscale_ratio=ratio, upscale_sharpness=sharpness, denoise=0, indirect_bounces=0)
    orc = b.oracle()
    inp = b.inputs(1)
    inp.temporal_upscalers = 1
    orc.render_frame(inp)
    assert orc.readback(L.OUT_TONE_MAPPED).shape == image_f16.shape
    orc.upload_state(L.OUT_TONE_MAPPED, image_f16)
    orc.run_pass(inp, 8)
    easu = orc.readback(L.OUT_UPSCALED)
    orc.run_pass(inp, 9)
    return easu, orc.readback(L.OUT_FSR_SHARPENED)


def random_image(h, w, seed):
    rng = np.random.default_rng(seed)
    smooth = rng.random((h // 4 + 2, w // 4 + 2, 3)).repeat(4, 0).repeat(4, 1)[:h, :w]      # blocks -> hard edges
    img = (0.7 * smooth + 0.3 * rng.random((h, w, 3))) * rng.choice([0.05, 1.0, 4.0])      # tone-mapped range and beyond
    out = np.ones((h, w, 4), np.float16)
    out[..., :3] = img.astype(np.float16)
    return out


@pytest.mark.parametrize("out_w,out_h,ratio,sharpness,seed", [(96, 64, 1.5, 0.0, 1), (80, 60, 2.0, 0.25, 2), (75, 41, 1.3, 1.0, 3),
                                                               (64, 48, 1.0, 2.0, 4)])
def test_oracle_fsr_equals_independent_numpy_restatement(out_w, out_h, ratio, sharpness, seed):
    rw = int(np.ceil(F(1.0) / F(ratio) * F(out_w))); rh = int(np.ceil(F(1.0) / F(ratio) * F(out_h)))
    img = random_image(rh, rw, seed)
    easu, rcas = run_fsr(img, out_w, out_h, ratio, sharpness)
    want, mn, mx = easu_numpy(img[..., :3].astype(F), out_w, out_h)
    assert easu.shape == (out_h, out_w, 4) and (easu[..., 3] == 1).all()
    assert np.array_equal(easu[..., :3].view(np.uint16), want.astype(np.float16).view(np.uint16))
    got = easu[..., :3].astype(F)
    assert (got >= mn.astype(np.float16).astype(F) - 1e-3).all() and (got <= mx.astype(np.float16).astype(F) + 1e-3).all()   # de-ringing
    want_rcas = rcas_numpy(easu[..., :3].astype(F), sharpness)
    assert np.array_equal(rcas[..., :3].view(np.uint16), want_rcas.astype(np.float16).view(np.uint16))
    assert (rcas[..., 3] == 1).all()


def test_fsr_flat_image_passes_through():
    img = np.ones((40, 60, 4), np.float16)
    img[..., :3] = np.array([0.25, 0.5, 0.125], np.float16)
    easu, rcas = run_fsr(img, 90, 60, 1.5, 0.0)
    assert np.array_equal(easu[..., :3], np.broadcast_to(img[0, 0, :3], easu[..., :3].shape))   # exactly: de-ringing clamps to the texel
    inner = rcas[1:-1, 1:-1, :3].astype(F)
    assert np.allclose(inner, img[0, 0, :3].astype(F), rtol=5e-3)        # APrxMedRcpF1 in the resolve reads ~0.3 % low
    # border pixels see zeros outside the image: the lobe collapses to 0, the pixel is only scaled by APrxMedRcpF1(1) = 0.99707
    through = (easu[..., :3].astype(F) * rcp_med(F(1.0))).astype(np.float16)
    assert np.array_equal(rcas[0, :, :3], through[0]) and np.array_equal(rcas[:, 0, :3], through[:, 0])
    assert np.array_equal(rcas[-1, :, :3], through[-1]) and np.array_equal(rcas[:, -1, :3], through[:, -1])


def test_fsr_ratio_one_keeps_texels_in_flat_regions_and_sharpens_edges():
    img = np.ones((48, 64, 4), np.float16)
    img[..., :3] = 0.2
    img[:, 32:, :3] = 0.8                                                  # vertical step edge
    easu, rcas = run_fsr(img, 64, 48, 1.0, 0.0)
    assert np.array_equal(easu, img)                                       # on the texel grid the kernel is interpolating: f gets weight 1
    r = rcas[..., 0].astype(F)
    assert np.allclose(r[5:-5, 5:25], 0.2, rtol=5e-3) and np.allclose(r[5:-5, 40:-5], 0.8, rtol=5e-3)
    # RCAS is limited to the local range (no overshoot beyond min/max of the ring) but steepens nothing at a 2-level step;
    # a one-pixel ridge, however, is raised relative to a blur: check on a ridge image
    ridge = np.ones((48, 64, 4), np.float16)
    ridge[..., :3] = 0.3
    ridge[:, 31, :3] = 0.5
    _, sharp0 = run_fsr(ridge, 64, 48, 1.0, 0.0)
    _, sharp2 = run_fsr(ridge, 64, 48, 1.0, 2.0)
    c0, c2 = float(sharp0[24, 31, 0]), float(sharp2[24, 31, 0])
    assert c0 > c2 >= 0.5 - 1e-3 and c0 <= 0.75                            # sharpness 0 sharpens most; 2 stops = a quarter of the lobe
