"""GPU parity tests, kernel by kernel, through the C ABI (librsx.so) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rsx():
    import torch
    from rs_image_segmentation_b200 import _lib, device
    device.require_cuda()
    return _lib


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _hist_np(r, L):
    B = r.shape[-1]
    flat = r.reshape(-1, B)
    return np.stack([np.bincount(flat[:, b], minlength=L) for b in range(B)])


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("n_px", [1, 5, 517, 4096, 4096 * 3 + 517, 300 * 211])
@pytest.mark.parametrize("B", [7, 5, 8])
def test_hist_u8(rsx, n_px, B):
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    rng = np.random.default_rng(n_px + B)
    r = rng.integers(0, 256, size=(n_px, B), dtype=np.uint8)
    r[: n_px // 2, 0] = 7                     # heavy same-bin contention
    d = _dev(r)
    h = torch.zeros((B, 256), dtype=torch.int32, device="cuda")
    rsx.call("rsx_hist_u8", ptr(d), n_px, B, ptr(h), stream_ptr())
    assert np.array_equal(h.cpu().numpy(), _hist_np(r, 256))


@pytest.mark.parametrize("n_px", [3, 2048, 2048 * 5 + 77])
def test_hist_u16(rsx, n_px):
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    rng = np.random.default_rng(n_px)
    r = rng.integers(0, 10001, size=(n_px, 13)).astype(np.uint16)
    d = _dev(r.view(np.int16))
    h = torch.zeros((13, 65536), dtype=torch.int32, device="cuda")
    rsx.call("rsx_hist_u16", ptr(d), n_px, 13, ptr(h), stream_ptr())
    assert np.array_equal(h.cpu().numpy(), _hist_np(r, 65536))


def _random_histograms(rng, B, kind):
    h = np.zeros((B, 256), np.int64)
    for b in range(B):
        k = kind if kind != "mixed" else ["uniform", "two", "constant", "skewed", "sparse", "huge"][b % 6]
        if k == "uniform":
            h[b] = rng.integers(0, 5000, 256)
        elif k == "two":
            h[b, rng.integers(0, 128)] = rng.integers(1, 1000)
            h[b, rng.integers(128, 256)] = rng.integers(1, 1000)
        elif k == "constant":
            h[b, rng.integers(0, 256)] = rng.integers(1, 100000)
        elif k == "skewed":
            h[b] = (rng.pareto(1.2, 256) * 50).astype(np.int64)
            h[b, 0] += 1
        elif k == "sparse":
            idx = rng.choice(256, 9, replace=False)
            h[b, idx] = rng.integers(1, 40, 9)
        else:                                                  # counts of the 40k x 40k mosaic: beyond 2^31 in total
            h[b] = rng.integers(0, 2 ** 24, 256)
            h[b, 100] += 2 ** 30
    return h


@pytest.mark.parametrize("B,kind,seed", [(7, "mixed", 1), (13, "mixed", 2), (16, "mixed", 3), (1, "sparse", 4), (7, "uniform", 5), (5, "two", 6),
                                        (7, "sparse", 7), (3, "constant", 8), (7, "skewed", 9)])
@pytest.mark.parametrize("is64", [0, 1])
def test_raster_stats_on_the_device_equal_the_host_restatement(rsx, B, kind, seed, is64):
    """rsx_raster_stats_u8_device (order statistics by a kernel, so that the histograms never visit the host between two kernels)
    against rsx_raster_stats, the host restatement of numpy's percentile arithmetic that tests/test_hoststats.py holds against
    numpy itself: every output bit for bit - percentiles that fall on and between samples, constant bands, 2^31+ counts."""
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    from rs_image_segmentation_b200 import hoststats
    rng = np.random.default_rng(seed)
    for rep in range(6):
        h = _random_histograms(rng, B, kind)
        if not is64:
            h = np.minimum(h, 2 ** 31 // 300)                  # the uint32 counters of one rank
        tb = int(rng.integers(0, B))
        lower, upper = [(2, 98), (0, 100), (5.5, 94.25), (50, 50)][rep % 4]
        ref = hoststats.RasterStats(h, glcm_band=tb, lower=lower, upper=upper)
        d_h = torch.from_numpy(h if is64 else h.astype(np.uint32).view(np.int32)).cuda()
        n_bytes, off = int(rsx.load().rsx_raster_stats_device_bytes()), int(rsx.load().rsx_raster_stats_device_lut_offset())
        blk = torch.zeros(n_bytes, dtype=torch.uint8, device="cuda")
        rsx.call("rsx_raster_stats_u8_device", ptr(d_h), is64, B, tb, float(lower), float(upper), ptr(blk), stream_ptr())
        raw = blk.cpu().numpy()
        norm = raw[:16 * 12].view(np.float32).reshape(16, 3)[:B]
        qnorm = raw[192:204].view(np.float32)
        center = raw[208:208 + 64].view(np.float32)[:B]
        scale = raw[272:272 + 128].view(np.float64)[:B]
        x_lut = raw[off:off + 16 * 1024].view(np.float32).reshape(16, 256)[:B]
        assert np.array_equal(norm.view(np.uint32), np.asarray(ref.norm, np.float32).view(np.uint32)), (rep, norm, ref.norm)
        assert np.array_equal(qnorm.view(np.uint32), np.asarray(ref.qnorm, np.float32).view(np.uint32)), (rep, qnorm, ref.qnorm)
        assert np.array_equal(center.view(np.uint32), np.asarray(ref.center, np.float32).view(np.uint32)), rep
        assert np.array_equal(scale.view(np.uint64), np.asarray(ref.scale, np.float64).view(np.uint64)), rep
        assert np.array_equal(x_lut.view(np.uint32), np.ascontiguousarray(ref.x_lut, np.float32).view(np.uint32)), rep


# ------------------------------------------------------------------------------------------ planar element-wise drop-ins
def test_planar_ops_bit_exact(rsx):
    import torch
    from oracle import features as of
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    rng = np.random.default_rng(0)
    n = 4096 * 2 + 3
    a, b, c, d = (rng.random(n).astype(np.float32) for _ in range(4))
    a[:10] = 0
    b[:10] = 0
    A, Bt, Ct, Dt = map(_dev, (a, b, c, d))
    out = torch.empty(n + 1, dtype=torch.float32, device="cuda")[:n]
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    st = stream_ptr()
    rsx.call("rsx_index_ratio_f32", ptr(A), ptr(Bt), n, ptr(out), st)
    assert np.array_equal(out.cpu().numpy(), of.ndvi(a, b))
    rsx.call("rsx_index_evi_f32", ptr(A), ptr(Bt), ptr(Ct), n, 1.0, 6.0, 7.5, 2.5, ptr(out), st)
    assert np.array_equal(out.cpu().numpy(), of.evi(a, b, c))
    rsx.call("rsx_index_msavi_f32", ptr(A), ptr(Bt), n, ptr(out), st)
    assert np.array_equal(out.cpu().numpy(), of.msavi(a, b), equal_nan=True)
    rsx.call("rsx_index_bsi_f32", ptr(A), ptr(Bt), ptr(Ct), ptr(Dt), n, ptr(out), st)
    assert np.array_equal(out.cpu().numpy(), of.bsi(a, b, c, d))
    x = (rng.random(n) * 255).astype(np.float32)
    lo, hi = np.percentile(x, 2), np.percentile(x, 98)
    den = np.float32(hi - lo + 1e-10)
    rsx.call("rsx_normalize_f32", ptr(_dev(x)), n, float(lo), float(hi), float(den), ptr(out), st)
    assert np.array_equal(out.cpu().numpy(), of.robust_normalize(x))


# ------------------------------------------------------------------------------------------ GLCM
DOC_IMAGE = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [0, 2, 2, 2], [2, 2, 3, 3]], dtype=np.uint8)
DOC_COUNTS = [
    [[2, 2, 1, 0], [0, 2, 0, 0], [0, 0, 3, 1], [0, 0, 0, 1]],
    [[1, 1, 3, 0], [0, 1, 1, 0], [0, 0, 0, 2], [0, 0, 0, 0]],
    [[3, 0, 2, 0], [0, 2, 2, 0], [0, 0, 1, 2], [0, 0, 0, 0]],
    [[2, 0, 0, 0], [1, 1, 2, 0], [0, 0, 2, 1], [0, 0, 0, 0]],
]


def _counts(rsx, q, L, win, anchors):
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    H, W = q.shape
    an = _dev(np.asarray(anchors, dtype=np.int32))
    out = torch.empty((len(anchors), 4, L, L), dtype=torch.int32, device="cuda")
    rsx.call("rsx_glcm_counts", ptr(_dev(q)), H, W, L, win, ptr(an), len(anchors), ptr(out), stream_ptr())
    return out.cpu().numpy().astype(np.uint32)


def test_glcm_counts_docstring_example(rsx):
    c = _counts(rsx, DOC_IMAGE, 4, 4, [(0, 0)])
    for a in range(4):
        assert c[0, a].tolist() == DOC_COUNTS[a]


def _texture_image(H, W, L, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, L, size=(H // 6 + 2, W // 6 + 2)).repeat(6, 0).repeat(6, 1)[:H, :W]
    return np.clip(base + rng.integers(-2, 3, size=(H, W)), 0, L - 1).astype(np.uint8)


@pytest.mark.parametrize("L,win", [(32, 7), (32, 21), (16, 5), (64, 11)])
def test_glcm_counts_bit_exact(rsx, L, win):
    from oracle import glcm as og
    q = _texture_image(60, 70, L, L + win)
    ref = og.counts_map_c(q, L, win, 1)
    rng = np.random.default_rng(1)
    anchors = [(int(rng.integers(0, 60 - win + 1)), int(rng.integers(0, 70 - win + 1))) for _ in range(25)] + [(0, 0), (60 - win, 70 - win)]
    got = _counts(rsx, q, L, win, anchors)
    for k, (i, j) in enumerate(anchors):
        assert np.array_equal(got[k], ref[i, j]), (i, j)


def _props(rsx, q, L, win, step):
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    H, W = q.shape
    oh, ow = (H - win) // step + 1, (W - win) // step + 1
    stride = (oh * ow + 31) // 32 * 32
    out = torch.zeros((5, stride), dtype=torch.float32, device="cuda")
    rsx.call("rsx_glcm_props", ptr(_dev(q)), H, W, L, win, step, oh, ow, ptr(out), stride, stream_ptr())
    return out[:, : oh * ow].reshape(5, oh, ow).cpu().numpy()


@pytest.mark.parametrize("L,win,step", [(32, 7, 1), (32, 21, 21), (16, 5, 1), (64, 11, 1), (32, 11, 1), (64, 5, 1),
                                        (16, 11, 1), (32, 5, 3), (32, 7, 7), (128, 9, 4)])
def test_glcm_props_match_oracle(rsx, L, win, step):
    from oracle import glcm as og
    q = _texture_image(97, 141, L, L * win + step)
    q[:30, :40] = 5                                 # constant area: correlation == 1 branch
    ref = og.props_map_c(q, L, win, step)
    got = _props(rsx, q, L, win, step)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_glcm_props_uniform_noise_worst_case(rsx):
    from oracle import glcm as og
    q = np.random.default_rng(5).integers(0, 32, size=(64, 200)).astype(np.uint8)
    np.testing.assert_allclose(_props(rsx, q, 32, 7, 1), og.props_map_c(q, 32, 7, 1), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("sh,dh", [((28, 28), (600, 600)), ((91, 135), (97, 141)), ((5, 7), (64, 96)), ((33, 17), (33, 17))])
def test_resize_matches_cv2(rsx, sh, dh):
    import cv2
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    rng = np.random.default_rng(sh[0])
    src = rng.random((3,) + sh).astype(np.float32) * 10
    d = torch.empty((3, dh[0] * dh[1]), dtype=torch.float32, device="cuda")
    rsx.call("rsx_resize_bilinear_f32", ptr(_dev(src)), sh[0], sh[1], 0, sh[0], sh[0] * sh[1], ptr(d), dh[0], dh[1], 0, dh[0],
             dh[0] * dh[1], 3, None, stream_ptr())
    got = d.reshape(3, *dh).cpu().numpy()
    for k in range(3):
        ref = cv2.resize(src[k], (dh[1], dh[0]), interpolation=cv2.INTER_LINEAR)
        np.testing.assert_allclose(got[k], ref, rtol=1e-5, atol=1e-5)


def test_resize_strip_equals_whole(rsx):
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    from rs_image_segmentation_b200.dist import resize_src_rows
    rng = np.random.default_rng(2)
    src = rng.random((1, 40, 30)).astype(np.float32)
    H, W = 173, 61
    whole = torch.empty((1, H * W), dtype=torch.float32, device="cuda")
    rsx.call("rsx_resize_bilinear_f32", ptr(_dev(src)), 40, 30, 0, 40, 1200, ptr(whole), H, W, 0, H, H * W, 1, None, stream_ptr())
    whole = whole.reshape(H, W).cpu().numpy()
    for r0, r1 in ((0, 60), (60, 120), (120, 173)):
        a, b = resize_src_rows(r0, r1, H, 40)
        part = torch.empty((1, (r1 - r0) * W), dtype=torch.float32, device="cuda")
        rsx.call("rsx_resize_bilinear_f32", ptr(_dev(src[:, a:b])), 40, 30, a, b - a, (b - a) * 30, ptr(part), H, W, r0, r1 - r0,
                 (r1 - r0) * W, 1, None, stream_ptr())
        assert np.array_equal(part.reshape(r1 - r0, W).cpu().numpy(), whole[r0:r1])


@pytest.mark.parametrize("levels,win", [(64, 7), (64, 5), (64, 11), (48, 9), (33, 3)])
def test_glcm_dense_wide_levels(levels, win):
    """Dense GLCM with more than 32 grey levels (wide packed moments, capped windows per CTA) against the oracle."""
    import torch
    from oracle import glcm as og
    from rs_image_segmentation_b200 import indices as I
    rng = np.random.default_rng(levels + win)
    yy, xx = np.mgrid[0:83, 0:147]
    band = (0.5 + 0.35 * np.sin(yy / 9.0) * np.cos(xx / 13.0) + 0.15 * rng.random((83, 147))).astype(np.float32)
    got = I.calculate_glcm_features(band, levels=levels, window_size=win, step_size=1)
    ref = og.glcm_features(band, levels, win, 1)
    for k in ref:
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-5, atol=1e-6, err_msg=f"{k} L={levels} w={win}")


@pytest.mark.parametrize("win,step", [(7, 1), (5, 1), (3, 1), (7, 7), (11, 1)])
def test_glcm_props_analytic_known_answers(rsx, win, step):
    """Vertical stripes / constant image: closed-form textbook values (tests/test_oracle_glcm.py), dense and tiled kernels."""
    from oracle.glcm import analytic_stripes, stripes_image
    names = ("contrast", "dissimilarity", "homogeneity", "energy", "correlation")
    for a, b, L in ((0, 31, 32), (3, 9, 16), (10, 11, 64)):
        q = stripes_image(a, b, 40, 66)
        got = _props(rsx, q, L, win, step)
        # windows starting on an `a` column (even) and on a `b` column (odd) swap na and nb
        exp_even = analytic_stripes(a, b, win)
        exp_odd = analytic_stripes(b, a, win)
        for i, k in enumerate(names):
            cols = np.arange(got.shape[2]) * step
            exp = np.where(cols % 2 == 0, exp_even[k], exp_odd[k])
            np.testing.assert_allclose(got[i], np.broadcast_to(exp, got[i].shape), rtol=1e-5, atol=1e-6, err_msg=f"{k} {a} {b} {L}")
    got = _props(rsx, np.full((30, 45), 7, np.uint8), 32, win, step)
    for i, v in enumerate((0, 0, 1, 1, 1)):
        assert np.all(got[i] == v)


@pytest.mark.parametrize("B,n_px", [(13, 12_000_003), (5, 3_000_001), (13, 70_001)])
def test_hist_u16_large_and_mixed_range(rsx, B, n_px):
    """uint16 histogram: shared-memory 16-bit counters with periodic flushes (> 65535 pixels per CTA), values above the
    shared window, a constant band (one bin takes every sample) - against torch.bincount."""
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    g = torch.Generator(device="cuda").manual_seed(n_px)
    r = torch.randint(0, 9000, (n_px, B), generator=g, device="cuda", dtype=torch.int32)
    r[:, 1] = torch.randint(0, 65536, (n_px,), generator=g, device="cuda", dtype=torch.int32)      # full 16-bit range
    r[:, 2] = 4321                                                                                 # one hot bin
    r[:, 3] = torch.randint(7000, 9000, (n_px,), generator=g, device="cuda", dtype=torch.int32)   # straddles the window edge
    d = r.to(torch.int16).contiguous()                       # uint16 bit patterns
    h = torch.zeros((B, 65536), dtype=torch.int32, device="cuda")
    rsx.call("rsx_hist_u16", ptr(d), n_px, B, ptr(h), stream_ptr())
    for b in range(B):
        ref = torch.bincount(r[:, b].to(torch.int64), minlength=65536)
        assert torch.equal(h[b].to(torch.int64), ref), b


# ------------------------------------------------------------------------------------------ read-backs without a copy engine
@pytest.mark.parametrize("n,dtype", [(1, "int64"), (3, "float64"), (14, "int32"), (7 * 256, "int32"), (100_003, "uint8"), (13 * 65536, "int64")])
def test_fetch_kernel_store_matches_copy(rsx, n, dtype):
    """device.fetch / rsx_store_to_host (kernel store into page-locked memory) returns the bytes of a plain copy, for
    aligned and unaligned starts and sizes that are not a multiple of 16."""
    import torch
    from rs_image_segmentation_b200.device import fetch, stage_to_host
    g = torch.Generator(device="cuda").manual_seed(n)
    base = torch.randint(0, 200, (n + 3,), generator=g, device="cuda").to(getattr(torch, dtype))
    for off in (0, 1, 3):
        t = base[off:off + n]
        assert np.array_equal(fetch(t), t.cpu().numpy())
    h = stage_to_host(base)
    torch.cuda.current_stream().synchronize()
    assert np.array_equal(h.numpy(), base.cpu().numpy())
    assert fetch(base[:0]).size == 0


# ============================================================================================ round-2 additions
def _moments(rsx, q, L, win, step):
    """(oh, ow, 4, 8) int64 pair moments written by the PRODUCTION kernels next to the property planes (rsx_glcm_moments)."""
    import torch
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    H, W = q.shape
    oh, ow = (H - win) // step + 1, (W - win) // step + 1
    stride = (oh * ow + 31) // 32 * 32
    props = torch.zeros((5, stride), dtype=torch.float32, device="cuda")
    mom = torch.full((oh * ow, 4, 8), -1, dtype=torch.int64, device="cuda")
    rsx.call("rsx_glcm_moments", ptr(_dev(q)), H, W, L, win, step, oh, ow, ptr(props), stride, ptr(mom), stream_ptr())
    return mom.cpu().numpy().reshape(oh, ow, 4, 8), props[:, : oh * ow].reshape(5, oh, ow).cpu().numpy()


@pytest.mark.parametrize("fold", [0, 1, 8, 16])
@pytest.mark.parametrize("L", [16, 32, 64])
@pytest.mark.parametrize("win", [5, 7, 11])
def test_glcm_production_kernel_integer_stage_is_exact(rsx, L, win, fold):
    """the integer stage of the DENSE production kernel (glcm_dense_kernel - packed pair moments, sliding
    private counters, folded or not) against the same integers derived from the oracle's graycomatrix counts, bit for bit,
    for every (levels, window) of BASELINE.json config C; the float32 properties written by the same launch stay within 1e-5."""
    from oracle import glcm as og
    from rs_image_segmentation_b200 import _lib
    q = _texture_image(97, 141, L, 7 * L + win)
    q[:30, :40] = 5
    q[60:, 100:] = np.random.default_rng(L + win).integers(0, L, size=(37, 41))        # wide level spans (flagged windows when folded)
    # fold: 0 = unfolded counters, 1 = chosen on the device from the span statistics, 8 / 16 = forced (every window whose levels
    # span more than the fold goes through the energy patch kernel)
    _lib.set_option("glcm_fold", 1 if fold else 0)
    _lib.set_option("glcm_fold_force", fold if fold > 1 else -1)
    _lib.set_option("glcm_fold_cap_div", 1 if fold > 1 else 256)        # forced: every flagged window fits the patch list
    try:
        got, props = _moments(rsx, q, L, win, 1)
    finally:
        _lib.set_option("glcm_fold", 1)
        _lib.set_option("glcm_fold_force", -1)
        _lib.set_option("glcm_fold_cap_div", 256)
    ref = og.pair_moments(q, L, win, 1)
    for f, name in enumerate(og.MOMENT_FIELDS):
        assert np.array_equal(got[..., f], ref[..., f]), name
    np.testing.assert_allclose(props, og.props_map_c(q, L, win, 1), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("L,win,step", [(32, 21, 21), (32, 7, 7), (128, 9, 4), (32, 5, 3)])
def test_glcm_general_kernel_integer_stage_is_exact(rsx, L, win, step):
    from oracle import glcm as og
    q = _texture_image(97, 141, L, L + win + step)
    got, props = _moments(rsx, q, L, win, step)
    ref = og.pair_moments(q, L, win, step)
    for f, name in enumerate(og.MOMENT_FIELDS):
        assert np.array_equal(got[..., f], ref[..., f]), name


def test_glcm_skimage_property_known_answers_on_the_gpu(rsx):
    """scikit-image's published property values for its docstring image (test_texture.py), reproduced from the integers the GPU
    kernel writes for the 4x4 window, angle 0."""
    from oracle import glcm as og
    SKIMAGE_PROP_KATS = og.SKIMAGE_PROP_KATS
    got, props = _moments(rsx, DOC_IMAGE, 4, 4, 4)
    pr = og.props_from_moments(got[0, 0, 0])
    assert abs(pr["contrast"] - 14 / 24) < 1e-12 and abs(pr["dissimilarity"] - 10 / 24) < 1e-12
    np.testing.assert_almost_equal(pr["homogeneity"], SKIMAGE_PROP_KATS["homogeneity"], decimal=7)
    np.testing.assert_almost_equal(pr["energy"], SKIMAGE_PROP_KATS["energy"], decimal=7)
    np.testing.assert_almost_equal(pr["correlation"], SKIMAGE_PROP_KATS["correlation"], decimal=7)
    # and the float32 plane values = mean over the four angles of the same quantities
    for k, name in enumerate(og.PROPS):
        want = np.mean([og.props_from_moments(got[0, 0, a])[name] for a in range(4)])
        assert abs(props[k, 0, 0] - want) <= 1e-6 * max(1.0, abs(want)), name
