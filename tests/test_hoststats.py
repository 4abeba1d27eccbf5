"""Host order statistics from histograms vs numpy / sklearn on the same data (CPU only)."""
import numpy as np
import pytest

from rs_image_segmentation_b200 import hoststats as hs


def _hist(band_u, L):
    return np.bincount(band_u.ravel(), minlength=L)


@pytest.mark.parametrize("n,seed", [(7, 0), (600, 1), (6144, 2), (100003, 3), (1 << 20, 4)])
def test_percentile_matches_numpy(n, seed):
    rng = np.random.default_rng(seed)
    b = np.clip(rng.normal(120, 40, size=n), 0, 255).astype(np.uint8)
    lo = hs.LevelOrder(_hist(b, 256), np.arange(256, dtype=np.float32))
    bf = b.astype(np.float32)
    for q in (0, 2, 25, 50, 98, 100):
        got, ref = lo.percentile_scalar(q), np.percentile(bf, q)
        assert type(got) == type(ref) == np.float32 and got == ref, (q, got, ref)


def test_large_n_float32_virtual_index():
    # (n-1)*q is evaluated in float32 by numpy for float32 data; replicate its rank choice
    rng = np.random.default_rng(9)
    n = 20_000_003
    b = rng.integers(0, 256, size=n, dtype=np.uint8)
    lo = hs.LevelOrder(_hist(b, 256), np.arange(256, dtype=np.float32))
    bf = b.astype(np.float32)
    for q in (2, 98):
        assert lo.percentile_scalar(q) == np.percentile(bf, q)


@pytest.mark.parametrize("shape,seed", [((64, 96), 0), ((211, 97), 1), ((600, 600), 2)])
def test_raster_stats_match_reference_ops(shape, seed):
    from sklearn.preprocessing import RobustScaler
    from oracle import features as of
    from oracle.glcm import quantize

    rng = np.random.default_rng(seed)
    B = 7
    base = rng.normal(size=(B,) + shape).cumsum(axis=2)
    bands_u8 = np.stack([np.clip((x - x.min()) / (x.max() - x.min()) * (120 + 19 * i) + 7 * i, 0, 255).astype(np.uint8)
                         for i, x in enumerate(base)])
    hist = np.stack([_hist(b, 256) for b in bands_u8])
    st = hs.RasterStats(hist, glcm_band=3)
    bands = [b.astype(np.float32) for b in bands_u8]
    nb = [of.robust_normalize(b) for b in bands]
    for b in range(B):
        assert st.norm[b, 0] == np.percentile(bands[b], 2) and st.norm[b, 1] == np.percentile(bands[b], 98)
        assert np.array_equal(st.norm_lut[b][bands_u8[b]], nb[b])
    X = np.stack([x.ravel() for x in nb], axis=1)
    rs = RobustScaler().fit(X)
    assert np.array_equal(st.center, rs.center_)
    assert np.array_equal(st.scale, rs.scale_)
    Xt = rs.transform(X.copy())
    for b in range(B):
        assert np.array_equal(st.x_lut[b][bands_u8[b]].ravel(), Xt[:, b])
    assert np.array_equal(st.quant_lut(3, 32)[bands_u8[3]], quantize(nb[3], 32))


def test_golden_crop(aa_crop):
    u8 = aa_crop["stage1_u8"]
    st = hs.RasterStats(np.stack([_hist(b, 256) for b in u8]), glcm_band=3)
    assert np.array_equal(st.norm[:, :2], aa_crop["pct"])
    for b in range(7):
        assert np.array_equal(st.norm_lut[b][u8[b]], aa_crop["norm"][b])
    assert np.array_equal(st.quant_lut(3, 32)[u8[3]], aa_crop["q32"])


def test_pca_from_moments_matches_sklearn_f64(aa_crop):
    from sklearn.decomposition import PCA
    u8 = aa_crop["stage1_u8"]
    st = hs.RasterStats(np.stack([_hist(b, 256) for b in u8]), glcm_band=3)
    X = np.stack([st.x_lut[b][u8[b]].ravel() for b in range(7)], axis=1).astype(np.float64)
    B = 7
    mom = np.concatenate([X.sum(0), (X.T @ X)[np.triu_indices(B)]])
    got = hs.pca_from_moments(mom, X.shape[0])
    ref = PCA(svd_solver="covariance_eigh").fit(X)
    assert np.allclose(got["components"], ref.components_, rtol=0, atol=1e-9)
    assert np.allclose(got["explained_variance_ratio"], ref.explained_variance_ratio_, rtol=1e-10)
    # and against the float32 reference run recorded in the golden file
    assert np.allclose(got["components"], aa_crop["pca_components"], rtol=0, atol=2e-5)
    assert np.allclose(got["explained_variance_ratio"], aa_crop["pca_evr"], rtol=1e-4)


def test_float_band_order():
    rng = np.random.default_rng(3)
    x = rng.random((50, 40)).astype(np.float32)
    lo = hs.float_band_order(x)
    for q in (2, 98):
        assert lo.percentile_scalar(q) == np.percentile(x, q)


def test_stage1_level_tables_match_reference_chain():
    """The fused stage-1 tables against the restated chain (and, through the golden file, against the reference itself:
    tests/golden/aa_full_stats.npz records the SHA-256 of the stage-1 output the reference functions produced)."""
    from oracle import features as of
    from rs_image_segmentation_b200 import hoststats
    rng = np.random.default_rng(0)
    raw = [rng.integers(lo, hi, size=(40, 50)).astype(np.uint8) for lo, hi in ((0, 256), (10, 200), (3, 4), (100, 255), (0, 2), (7, 90), (50, 60))]
    ref = of.stage1_preprocess(raw)
    hist = np.stack([np.bincount(r.ravel(), minlength=256) for r in raw])
    remap, hist1 = hoststats.stage1_level_tables(hist)
    for b in range(7):
        if raw[b].min() == raw[b].max():
            continue                                                    # 0/0: the reference produces NaN -> undefined uint8
        assert np.array_equal(remap[b][raw[b]], ref[b]), b
        assert np.array_equal(hist1[b], np.bincount(ref[b].ravel(), minlength=256)), b


def _random_hists(rng, B, L, kind):
    if kind == "dense":
        return rng.integers(0, 5000, size=(B, L)).astype(np.int64)
    if kind == "sparse":
        h = np.zeros((B, L), np.int64)
        for b in range(B):
            idx = rng.choice(L, size=int(rng.integers(1, 12)), replace=False)
            h[b, idx] = rng.integers(1, 40, size=idx.size)
        return h
    if kind == "huge":                                   # n > 2^24: the float32 virtual index is no longer exact
        return rng.integers(0, 400000, size=(B, L)).astype(np.int64)
    if kind == "single":
        h = np.zeros((B, L), np.int64)
        h[np.arange(B), rng.integers(0, L, size=B)] = rng.integers(1, 1000, size=B)
        return h
    h = rng.integers(0, 3, size=(B, L)).astype(np.int64)  # "tiny": few samples, many ties
    h[:, 0] += 1
    return h


@pytest.mark.parametrize("L", [256, 65536])
@pytest.mark.parametrize("kind", ["dense", "sparse", "huge", "single", "tiny"])
def test_native_raster_stats_match_python_statement(L, kind):
    """rsx_raster_stats (C, in librsx.so) against the numpy-faithful Python statement, bit for bit."""
    from rs_image_segmentation_b200 import hoststats
    rng = np.random.default_rng(L + len(kind))
    for trial in range(3 if L == 256 else 1):
        B = int(rng.integers(1, 14))
        hist = _random_hists(rng, B, L, kind)
        if L == 65536:
            hist[:, 12000:] = 0                           # keep the Python statement quick
            hist[:, 0] += 1
        gb = int(rng.integers(0, B))
        for lower, upper in ((2, 98), (0, 100), (10.5, 63.25)):
            a = hoststats.RasterStats(hist, glcm_band=gb, lower=lower, upper=upper, native=True)
            b = hoststats.RasterStats(hist, glcm_band=gb, lower=lower, upper=upper, native=False)
            assert np.array_equal(a.norm.view(np.uint32), b.norm.view(np.uint32))
            assert np.array_equal(a.qnorm.view(np.uint32), b.qnorm.view(np.uint32))
            assert np.array_equal(a.norm_lut.view(np.uint32), b.norm_lut.view(np.uint32))
            assert np.array_equal(a.center.view(np.uint32), b.center.view(np.uint32))
            assert np.array_equal(a.scale.view(np.uint64), b.scale.view(np.uint64))
            assert np.array_equal(a.x_lut.view(np.uint32), b.x_lut.view(np.uint32))
