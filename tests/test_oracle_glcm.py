"""GLCM oracle: scikit-image docstring known-answer example + self-consistency.
(Parity for this stage is unpinned by the reference: see oracle/glcm.py.)"""
import numpy as np

from oracle import glcm as og

DOC_IMAGE = np.array([[0, 0, 1, 1], [0, 0, 1, 1], [0, 2, 2, 2], [2, 2, 3, 3]], dtype=np.uint8)
DOC_COUNTS = [
    [[2, 2, 1, 0], [0, 2, 0, 0], [0, 0, 3, 1], [0, 0, 0, 1]],
    [[1, 1, 3, 0], [0, 1, 1, 0], [0, 0, 0, 2], [0, 0, 0, 0]],
    [[3, 0, 2, 0], [0, 2, 2, 0], [0, 0, 1, 2], [0, 0, 0, 0]],
    [[2, 0, 0, 0], [1, 1, 2, 0], [0, 0, 2, 1], [0, 0, 0, 0]],
]


def test_docstring_example_numpy():
    P = og.graycomatrix(DOC_IMAGE, (1,), og.DEFAULT_ANGLES, levels=4)
    for a in range(4):
        assert P[:, :, 0, a].tolist() == DOC_COUNTS[a]


def test_docstring_example_c():
    cnt = og.counts_map_c(DOC_IMAGE, 4, 4, 4)
    assert cnt.shape == (1, 1, 4, 4, 4)
    for a in range(4):
        assert cnt[0, 0, a].tolist() == DOC_COUNTS[a]


def test_offsets():
    assert [(dr, dc) for _, _, dr, dc in og.offsets_for()] == [(0, 1), (1, 1), (1, 0), (1, -1)]


def test_c_matches_numpy_props():
    rng = np.random.default_rng(0)
    base = rng.integers(0, 32, size=(8, 10)).repeat(5, 0).repeat(5, 1)
    q = np.clip(base + rng.integers(-2, 3, size=base.shape), 0, 31).astype(np.uint8)
    for win, step, L in ((7, 1, 32), (21, 21, 32), (5, 3, 16), (11, 4, 64)):
        qq = (q.astype(np.int32) * L // 32).astype(np.uint8)
        a = og.props_map_numpy(qq[:30, :34], L, win, step)
        b = og.props_map_c(qq[:30, :34], L, win, step)
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=2e-6, atol=1e-7)


def test_constant_window_correlation_is_one():
    q = np.full((9, 9), 3, np.uint8)
    m = og.props_map_numpy(q, 8, 7, 1)
    assert np.all(m[4] == 1.0) and np.all(m[0] == 0.0) and np.all(m[3] == 1.0)
    assert np.array_equal(m, og.props_map_c(q, 8, 7, 1))


def test_pair_counts_per_angle():
    q = np.zeros((7, 7), np.uint8)
    cnt = og.counts_map_c(q, 4, 7, 1)[0, 0]
    assert [int(cnt[a].sum()) for a in range(4)] == [42, 36, 42, 36]


from oracle.glcm import analytic_stripes, stripes_image  # noqa: E402


def test_analytic_known_answers():
    """Constant image, vertical stripes, horizontal stripes: properties from the definitions, not from an implementation."""
    for win in (3, 5, 7):
        for a, b, L in ((0, 31, 32), (3, 9, 16), (10, 11, 64)):
            m = og.props_map_numpy(stripes_image(a, b, win, win), L, win, 1)[:, 0, 0]
            exp = analytic_stripes(a, b, win)
            for i, k in enumerate(og.PROPS):
                assert abs(m[i] - exp[k]) < 1e-6, (win, a, b, k, m[i], exp[k])
            # the transposed pattern swaps the roles of 0 deg and 90 deg; the mean over the four angles is unchanged
            mt = og.props_map_numpy(stripes_image(a, b, win, win).T.copy(), L, win, 1)[:, 0, 0]
            assert np.allclose(mt, m, atol=1e-6)
    c = og.props_map_numpy(np.full((7, 7), 5, np.uint8), 32, 7, 1)[:, 0, 0]
    assert np.allclose(c, [0, 0, 1, 1, 1])


# ------------------------------------------------------------------------------------------ folded counters
def _asm_unordered(win, dr, dc, fold=None):
    """sum over unordered cells of weight * count^2 (weight 2 off the diagonal, 4 on it) = sum of squares of the symmetric
    GLCM, with the cells optionally indexed by the levels modulo `fold`."""
    h, w = win.shape
    a = win[:h - dr, max(0, -dc):w - max(0, dc)].ravel().astype(np.int64)
    b = win[dr:, max(0, dc):w + min(0, dc)].ravel().astype(np.int64)
    diag = a == b                                            # the weight belongs to the TRUE pair, as in the kernel's code word
    if fold:
        a, b = a % fold, b % fold
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = hi * 64 + lo
    total = 0
    for k in np.unique(key):
        sel = key == k
        u = int(sel.sum())
        # a folded slot may mix diagonal and off-diagonal true cells only when the fold is not injective
        wgt = 4 if diag[sel].all() else 2
        total += wgt * u * u
    return total


def test_folded_counters_are_exact_when_the_level_span_is_below_the_fold():
    """The planned 36-cell private counters index a cell by (a mod 8, b mod 8): exact whenever the grey levels of the window
    span fewer than 8 values (the fold is injective there); a wider span can break it, which is why such windows get flagged."""
    rng = np.random.default_rng(8)
    offs = [(0, 1), (1, 1), (1, 0), (1, -1)]
    for _ in range(300):
        base = int(rng.integers(0, 25))
        win = base + rng.integers(0, 8, size=(7, 7))         # span <= 7
        for dr, dc in offs:
            assert _asm_unordered(win, dr, dc) == _asm_unordered(win, dr, dc, fold=8)
    # span 8: levels 3 and 11 fold onto each other, the counts of two different cells merge
    win = np.full((7, 7), 3)
    win[:, 4:] = 11
    assert _asm_unordered(win, 0, 1) != _asm_unordered(win, 0, 1, fold=8)
    assert int(win.max() - win.min()) >= 8                   # ... and the span test catches it


# Property known answers published by scikit-image's own test suite (skimage/feature/tests/test_texture.py: test_contrast,
# test_dissimilarity, test_homogeneity, test_energy, test_correlation) for the docstring image, symmetric=True, normed=True,
# first (distance, angle) entry = (1, 0).  scikit-image cannot be installed here; these constants are its third-party pin.
SKIMAGE_PROP_KATS = og.SKIMAGE_PROP_KATS


def test_skimage_property_known_answers():
    P = og.graycomatrix(DOC_IMAGE, (1, 2), (0.0,), levels=4, symmetric=True, normed=True)
    # exact values (the matrix of d=1, angle 0 has 24 pair instances)
    assert abs(og.graycoprops(P, "contrast")[0, 0] - 14 / 24) < 1e-12            # 0.58333...
    assert abs(og.graycoprops(P, "dissimilarity")[0, 0] - 10 / 24) < 1e-12       # 0.41666...
    np.testing.assert_almost_equal(og.graycoprops(P, "homogeneity")[0, 0], SKIMAGE_PROP_KATS["homogeneity"], decimal=7)
    np.testing.assert_almost_equal(og.graycoprops(P, "energy")[0, 0], SKIMAGE_PROP_KATS["energy"], decimal=7)
    corr = og.graycoprops(P, "correlation")
    np.testing.assert_almost_equal(corr[0, 0], SKIMAGE_PROP_KATS["correlation"], decimal=7)
    np.testing.assert_almost_equal(corr[1, 0], SKIMAGE_PROP_KATS["correlation_d2"], decimal=7)
    # scikit-image's tests round the normalised matrix to 3 decimals first for contrast and dissimilarity
    Pr = np.round(P, 3)
    np.testing.assert_almost_equal(og.graycoprops(Pr, "contrast")[0, 0], SKIMAGE_PROP_KATS["contrast_rounded"], decimal=3)
    np.testing.assert_almost_equal(og.graycoprops(Pr, "dissimilarity")[0, 0], SKIMAGE_PROP_KATS["dissimilarity_rounded"], decimal=3)


def test_pair_moment_route_reproduces_the_known_answers():
    """The integer pair moments the dense GPU kernel keeps per angle (s1, sa, sq, sab, e, neq; rsx_glcm.cu) give the same
    property values as the histogram route - checked here on the scikit-image known answers through the numpy statement of those
    moments (oracle.glcm.pair_moments), which is also what the GPU integer outputs are compared with, exactly."""
    m = og.pair_moments(DOC_IMAGE, 4, 4)[0, 0, 0]                 # window (0, 0), angle 0
    pr = og.props_from_moments(m)
    assert abs(pr["contrast"] - 14 / 24) < 1e-12 and abs(pr["dissimilarity"] - 10 / 24) < 1e-12
    np.testing.assert_almost_equal(pr["homogeneity"], SKIMAGE_PROP_KATS["homogeneity"], decimal=7)
    np.testing.assert_almost_equal(pr["energy"], SKIMAGE_PROP_KATS["energy"], decimal=7)
    np.testing.assert_almost_equal(pr["correlation"], SKIMAGE_PROP_KATS["correlation"], decimal=7)
    # and against the histogram route on random windows, all four angles
    rng = np.random.default_rng(3)
    q = rng.integers(0, 16, size=(9, 12)).astype(np.uint8)
    mm = og.pair_moments(q, 16, 7)
    for i in range(mm.shape[0]):
        for j in range(mm.shape[1]):
            P = og.graycomatrix(q[i:i + 7, j:j + 7], (1,), og.DEFAULT_ANGLES, 16, symmetric=True, normed=True)
            for a in range(4):
                pr = og.props_from_moments(mm[i, j, a])
                for name in og.PROPS:
                    assert abs(pr[name] - og.graycoprops(P, name)[0, a]) < 1e-12, (i, j, a, name)
