"""The oracle restatement against outputs of the reference's own functions
(tests/golden/aa_crop.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import features as of
from oracle import kmeans as ok


def _bands(g):
    return [b.astype(np.float32) for b in g["stage1_u8"]]


def test_normalize_bit_exact(aa_crop):
    for b, ref in zip(_bands(aa_crop), aa_crop["norm"]):
        out = of.robust_normalize(b)
        assert out.dtype == np.float32
        assert np.array_equal(out, ref)


def test_indices_bit_exact(aa_crop):
    nb = [of.robust_normalize(b) for b in _bands(aa_crop)]
    ix = of.all_indices(nb)
    for k in of.INDEX_ORDER:
        assert ix[k].dtype == np.float32
        assert np.array_equal(ix[k], aa_crop["ix_" + k]), k


def test_pca_matches_reference(aa_crop):
    nb = [of.robust_normalize(b) for b in _bands(aa_crop)]
    pcs, evr, model = of.perform_pca(nb)
    assert np.array_equal(np.stack(pcs), aa_crop["pca_maps"])
    assert np.array_equal(evr, aa_crop["pca_evr"])
    assert np.array_equal(model.components_, aa_crop["pca_components"])


def test_level1_and_context(aa_crop):
    nb = [of.robust_normalize(b) for b in _bands(aa_crop)]
    ix = of.all_indices(nb)
    pcs, _, _ = of.perform_pca(nb)
    l1 = of.level1_stack(ix, pcs)
    assert np.array_equal(l1, aa_crop["level1"])
    assert np.array_equal(of.spatial_context(l1), aa_crop["level1_ctx"])


def test_quantised_nir(aa_crop):
    from oracle.glcm import quantize
    nir = of.robust_normalize(_bands(aa_crop)[3])
    assert np.array_equal(quantize(nir, 32), aa_crop["q32"])


def test_kmeans_function_matches_reference(aa_crop):
    h, w = aa_crop["stage1_u8"].shape[1:]
    fd = {k: aa_crop["ix_" + k] for k in of.INDEX_ORDER}
    fd.update(height=h, width=w)
    lab = ok.kmeans_classification(fd, n_clusters=5, keys=list(of.INDEX_ORDER))
    assert lab.dtype == np.int32
    assert np.array_equal(lab, aa_crop["kmeans_labels_k5"])


def test_kmeans_error_behaviour():
    with pytest.raises(ValueError):
        ok.stack_from_dict({}, None)
    with pytest.raises(ValueError):
        ok.stack_from_dict({"height": 4, "width": 4}, None)
    with pytest.raises(ValueError):
        ok.stack_from_dict({"height": 4, "width": 4, "a": np.zeros((3, 3))}, ["a"])


def test_lloyd_restatements_agree(aa_crop):
    X = np.stack([aa_crop["ix_" + k].ravel() for k in of.INDEX_ORDER], axis=1).astype(np.float64)
    Xs = ok.minmax_scale(X)
    rng = np.random.default_rng(5)
    C0 = Xs[rng.choice(Xs.shape[0], 6, replace=False)]
    l1, c1, i1, n1 = ok.lloyd_fixed(Xs, C0, 7)
    l2, c2, i2, n2 = ok.lloyd_numpy(Xs, C0, 7)
    assert np.array_equal(l1, l2)
    assert np.allclose(c1, c2, rtol=0, atol=1e-12)
    assert abs(i1 - i2) <= 1e-9 * i1


def test_cv2_blur_is_a_correctly_rounded_double_sum(aa_crop):
    """The N2 local-std kernel (rsx_stencil.cu local_std_kernel) sums the 25 taps in double in its own order and rounds
    (float)(sum * 1/25) once.  cv2.blur, which the reference calls (indices.py:531-548), keeps running double sums; the two
    orders never reach the float32 result, so the kernel can be - and is tested to be - bit exact against cv2."""
    import cv2
    from oracle import features as of
    b = of.robust_normalize(aa_crop["stage1_u8"][3].astype(np.float32))
    rng = np.random.default_rng(3)
    for img in (b, (rng.random((301, 517), dtype=np.float32) * 3 - 1)):
        H, W = img.shape
        for ks in (3, 5, 7):
            R = ks // 2
            for a in (img, img * img):
                p = np.pad(a.astype(np.float64), R, mode="reflect")
                rows = np.zeros((H + 2 * R, W))
                for dx in range(ks):
                    rows += p[:, dx:dx + W]
                s = np.zeros((H, W))
                for dy in range(ks):
                    s += rows[dy:dy + H]
                assert np.array_equal((s * (1.0 / (ks * ks))).astype(np.float32), cv2.blur(a, (ks, ks)))


def test_full_scene_table(aa_full_stats):
    # the numbers quoted in SURVEY.md 8(c), regenerated from the reference
    assert aa_full_stats["pct"].tolist() == [[59, 90], [18, 77], [8, 59], [28, 80], [10, 97], [63, 177], [4, 59]]
    assert int(aa_full_stats["q32_max"]) == 31


def test_stage1_restatement_matches_the_reference_output(aa_full_stats):
    """oracle.features.stage1_preprocess (gain/bias -> identity warp -> min-max stretch -> uint8,
    modules/features/preprocessing.py:54-125) against the stage-1 output the UNMODIFIED reference functions produced for the
    bundled scene: the SHA-256 of that (7, 600, 600) uint8 array is in the golden file (tests/golden/make_golden.py).  Needs the
    reference's data/raw/AA.tif, i.e. runs in the build container only."""
    import hashlib
    import importlib.util
    import os

    import pytest

    from oracle import features as of
    tif = "/root/reference/data/raw/AA.tif"
    if not os.path.exists(tif):
        pytest.skip("the reference's bundled scene is not on this machine")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    raw = mg.read_tiff_planar_u8(tif)
    stage1 = np.stack(of.stage1_preprocess([raw[i] for i in range(raw.shape[0])]))
    assert stage1.dtype == np.uint8 and stage1.shape == (7, 600, 600)
    assert hashlib.sha256(stage1.tobytes()).digest() == aa_full_stats["stage1_sha"].tobytes()
    # and the level-table route the GPU path fuses into its loads (FeatureConfig.stage1) gives the same bytes
    from rs_image_segmentation_b200 import hoststats
    hist = np.stack([np.bincount(raw[b].ravel(), minlength=256) for b in range(7)]).astype(np.int64)
    remap, _ = hoststats.stage1_level_tables(hist, of.TM_GAIN, of.TM_BIAS)
    fused = np.stack([remap[b][raw[b]] for b in range(7)]).astype(np.uint8)
    assert np.array_equal(fused, stage1)
