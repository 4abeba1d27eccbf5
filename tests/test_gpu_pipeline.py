"""GPU parity tests for the fused path: raster -> feature stack -> KMeans labels, against the oracle
(and against the reference outputs recorded in tests/golden/aa_crop.npz)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from rs_image_segmentation_b200 import device, pipeline
    device.require_cuda()
    return pipeline


def _to_bip(planar_u8):
    return np.ascontiguousarray(np.moveaxis(planar_u8, 0, -1))


def _oracle_features(bip, cfg):
    """Oracle feature maps of a (H, W, B) integer raster, in pipeline plane order."""
    from oracle import features as of
    from oracle import glcm as og
    bands = [bip[:, :, b].astype(np.float32) for b in range(bip.shape[2])]
    nb = [of.robust_normalize(b) for b in bands]
    bm = cfg.band_map
    ix = of.all_indices([nb[bm[0]], nb[bm[1]], nb[bm[2]], nb[bm[3]], nb[bm[4]]])
    out = {k: ix[k] for k in of.INDEX_ORDER}
    if cfg.glcm:
        g = og.glcm_features(nb[bm[3]], cfg.glcm_levels, cfg.glcm_window, cfg.glcm_step)
        out.update({"glcm_" + k: v for k, v in g.items()})
        out["_q"] = og.quantize(nb[bm[3]], cfg.glcm_levels)
    pcs, evr, model = of.perform_pca(nb, n_components=cfg.n_components)
    for i, p in enumerate(pcs):
        out[f"pc{i}"] = p
    out["_evr"], out["_components"] = evr, model.components_
    pcs64, evr64, model64 = of.perform_pca(nb, n_components=cfg.n_components, promote=True)
    out["_pcs64"], out["_evr64"], out["_components64"] = pcs64, evr64, model64.components_
    out["_eigvals64"] = PCA_all_eigvals(nb)
    return out


def PCA_all_eigvals(nb):
    from oracle import features as of
    return of.perform_pca(nb, promote=True)[2].explained_variance_


def _check_features(fr, ref, cfg, pc_tol=2e-4):
    from oracle import features as of
    for k in of.INDEX_ORDER:
        assert np.array_equal(fr.plane(k).cpu().numpy(), ref[k], equal_nan=True), k       # bit exact
    if cfg.glcm:
        assert np.array_equal(fr.quant[:fr.n_px].view(fr.H, fr.W).cpu().numpy(), ref["_q"])  # bit exact
        for k in ("contrast", "dissimilarity", "homogeneity", "energy", "correlation"):
            np.testing.assert_allclose(fr.plane("glcm_" + k).cpu().numpy(), ref["glcm_" + k], rtol=1e-5, atol=1e-6, err_msg=k)
    comps = fr.pca["components"]
    lam = ref["_eigvals64"]
    for i in range(comps.shape[0]):
        got = fr.plane(f"pc{i}").cpu().numpy()
        # (1) the same PCA in float64 (tight): components identical incl. sign convention, maps within 1e-5
        r64 = ref["_components64"][i]
        assert np.abs(comps[i] - r64).max() < 1e-7, (i, np.abs(comps[i] - r64).max())
        scale = max(1.0, float(np.abs(ref["_pcs64"][i]).max()))
        assert np.abs(got - ref["_pcs64"][i]).max() <= 1e-5 * scale, (i, np.abs(got - ref["_pcs64"][i]).max())
        # (2) the reference's own float32 run, up to sign; its Gram matrix is float32, so allow its conditioning:
        #     eigenvector error ~ eps32 * lambda_max / gap_i
        r = ref["_components"][i]
        gap = min(abs(lam[i] - lam[j]) for j in range(len(lam)) if j != i)
        tol = max(pc_tol, 50 * 6e-8 * lam[0] / gap) * scale
        sgn = np.sign(np.dot(comps[i], r))
        assert np.abs(got * sgn - ref[f"pc{i}"]).max() <= tol, (i, tol)
    np.testing.assert_allclose(fr.pca["explained_variance_ratio"], ref["_evr64"], rtol=1e-9)
    np.testing.assert_allclose(fr.pca["explained_variance_ratio"], ref["_evr"], rtol=5e-4)


def _mm_check(fr):
    mn, mx = fr.minmax.read()
    for i, name in enumerate(fr.names):
        p = fr.plane(name).cpu().numpy()
        assert mn[i] == np.nanmin(p) and mx[i] == np.nanmax(p), name


def test_golden_crop_features(P, aa_crop):
    import torch
    bip = _to_bip(aa_crop["stage1_u8"])
    cfg = P.FeatureConfig(glcm_window=21, glcm_step=21)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), cfg)
    # reference outputs (golden): indices, quantised NIR and PCA maps produced by the reference's own functions
    for k in P.INDEX_NAMES:
        assert np.array_equal(fr.plane(k).cpu().numpy(), aa_crop["ix_" + k]), k
    assert np.array_equal(fr.quant[:fr.n_px].view(fr.H, fr.W).cpu().numpy(), aa_crop["q32"])
    for i in range(7):
        r = aa_crop["pca_components"][i]
        c = fr.pca["components"][i]
        sgn = np.sign(np.dot(c, r))
        assert abs(np.dot(c, r)) > 1 - 1e-6
        assert np.abs(fr.plane(f"pc{i}").cpu().numpy() * sgn - aa_crop["pca_maps"][i]).max() < 2e-4
    _check_features(fr, _oracle_features(bip, cfg), cfg)
    _mm_check(fr)


@pytest.mark.parametrize("H,W,win,step,seed", [(97, 141, 7, 1, 1), (233, 310, 21, 21, 2), (64, 4096 // 64 + 3, 5, 1, 3), (130, 257, 11, 1, 4)])
def test_synthetic_u8_features(P, H, W, win, step, seed):
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(H, W, 7, np.uint8, seed, cell=16)
    cfg = P.FeatureConfig(glcm_window=win, glcm_step=step)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), cfg)
    _check_features(fr, _oracle_features(bip, cfg), cfg)
    _mm_check(fr)


@pytest.mark.parametrize("level_table", [False, True])
def test_synthetic_u16_13band_features(P, level_table):
    """uint16 x 13 bands; the PCA's scaled values either by reciprocal arithmetic (default) or from the exact per-level table."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(150, 203, 13, np.uint16, 7, cell=16)
    cfg = P.FeatureConfig(band_map=(1, 2, 3, 7, 11), n_components=6, glcm=False, u16_level_table=level_table)
    fr = P.extract_features(torch.from_numpy(bip.view(np.int16)).cuda(), cfg)
    _check_features(fr, _oracle_features(bip, cfg), cfg)
    _mm_check(fr)


def _kmeans_oracle(stack_nd, c0_scaled, n_iter, fmin, fmax):
    from oracle import kmeans as ok
    X = stack_nd.astype(np.float64)
    rng_ = fmax.astype(np.float64) - fmin.astype(np.float64)
    rng_[rng_ < 10 * np.finfo(np.float64).eps] = 1.0
    Xs = ok.minmax_scale(X)
    return ok.lloyd_fixed(Xs, c0_scaled, n_iter), Xs


@pytest.mark.parametrize("K,n_iter,H,W", [(8, 5, 120, 160), (5, 20, 97, 141), (32, 6, 150, 200), (3, 1, 33, 35)])
def test_kmeans_labels_bit_exact(P, K, n_iter, H, W):
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(H, W, 7, np.uint8, K + n_iter, cell=16)
    cfg = P.FeatureConfig(glcm_window=7, glcm_step=1)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), cfg)
    D = 13
    res, km, c0 = P.kmeans_on_features(fr, D, K, n_iter, seed=11)
    stack = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy()           # (N, D) float32, the GPU's own stack
    (lab, cent, inertia, n_run), Xs = _kmeans_oracle(stack, c0, n_iter, km.fmin, km.fmax)
    assert n_run == n_iter
    assert np.allclose(Xs[P.draw_init_indices(H * W, K, 11)], c0, rtol=0, atol=1e-15)
    got = res.labels.cpu().numpy()
    assert got.dtype == np.int32
    assert np.array_equal(got, lab), f"{(got != lab).sum()} labels differ"
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)
    assert abs(res.inertia - inertia) <= 1e-5 * inertia


@pytest.mark.parametrize("D,K,n_iter,n_px", [(1, 2, 4, 5003), (8, 64, 3, 20011), (9, 17, 4, 12007), (16, 9, 3, 9001), (20, 64, 3, 30011),
                                             (20, 1, 2, 4099), (13, 33, 3, 16411)])
def test_kmeans_depth_and_cluster_extremes(P, D, K, n_iter, n_px):
    """Every compiled range of stack depths (1..20) and the extremes of K (1, the 16/32/64 tag-width boundaries) on a random
    planar stack with correlated, differently scaled features: labels, centroids and inertia against the oracle."""
    import torch
    rng = np.random.default_rng(1000 * D + K)
    base = rng.normal(size=(n_px, 3))
    mix = rng.normal(size=(3, D))
    X = (base @ mix + 0.3 * rng.normal(size=(n_px, D))) * rng.uniform(0.01, 50.0, size=D) + rng.uniform(-5, 5, size=D)
    X = X.astype(np.float32)
    stride = (n_px + 31) // 32 * 32
    planes = torch.zeros((D, stride), dtype=torch.float32, device="cuda")
    planes[:, :n_px] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    fmin, fmax = X.min(axis=0), X.max(axis=0)
    km = P.DeviceKMeans(planes, n_px, D, K, fmin, fmax, n_px, 257)
    idx = P.draw_init_indices(n_px, K, 5)
    c0 = km.scale_rows(X[idx])
    res = km.fit(c0, n_iter)
    (lab, cent, inertia, n_run), _ = _kmeans_oracle(X, c0, n_iter, fmin, fmax)
    got = res.labels.cpu().numpy()
    assert np.array_equal(got, lab), f"{(got != lab).sum()} labels differ"
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)
    assert abs(res.inertia - inertia) <= 1e-5 * max(inertia, 1e-30)


def test_kmeans_partition_invariance(P):
    """Integer partial sums: assigning two halves separately and adding the accumulators equals one pass."""
    import torch
    from rs_image_segmentation_b200 import _lib
    from rs_image_segmentation_b200.device import ptr, stream_ptr
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(200, 300, 7, np.uint8, 3, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm=False))
    D, K = 8, 6
    mn, mx = fr.minmax.read()
    km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W)
    c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 1), 0))
    km.setup(c0)
    _lib.call("rsx_kmeans_assign", ptr(fr.planes), km.stride, fr.n_px, fr.W, ptr(km.state), ptr(km.acc), None, None, None, None, 1, D, K, stream_ptr())
    whole = km.acc.clone()
    km.acc.zero_()
    half = (fr.n_px // 2) // 4 * 4 + 4
    import ctypes as C
    _lib.call("rsx_kmeans_assign", ptr(fr.planes), km.stride, half, 64, ptr(km.state), ptr(km.acc), None, None, None, None, 1, D, K, stream_ptr())
    _lib.call("rsx_kmeans_assign", C.c_void_p(fr.planes.data_ptr() + 4 * half), km.stride, fr.n_px - half, 1000, ptr(km.state), ptr(km.acc),
              None, None, None, None, 1, D, K, stream_ptr())
    assert torch.equal(whole[:K * D + K], km.acc[:K * D + K])
    assert int(whole[K * D:K * D + K].sum()) == fr.n_px


@pytest.mark.parametrize("K,D,n_iter", [(8, 13, 7), (20, 9, 5)])
def test_kmeans_delta_passes_equal_full_passes(P, K, D, n_iter):
    """Delta passes (only pixels whose label changed touch the integer sums) give bit-identical totals, centroids,
    labels and counters' meaning as recomputing the sums from scratch in every pass."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(143, 211, 7, np.uint8, 21, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm_window=5, glcm_step=1))
    out = {}
    for delta in (False, True):
        res, km, c0 = P.kmeans_on_features(fr, D, K, n_iter, seed=5, delta=delta)
        n = km.n_acc
        out[delta] = (res.labels.cpu().numpy(), res.centroids, res.inertia, km.acc[n:n + K * D + K].cpu().numpy(), res.near_ties)
    assert np.array_equal(out[False][3], out[True][3])                  # integer totals: sums and counts
    assert int(out[True][3][K * D:].sum()) == fr.n_px
    assert np.array_equal(out[False][1], out[True][1])                  # centroids bit for bit
    assert np.array_equal(out[False][0], out[True][0])
    assert abs(out[False][2] - out[True][2]) <= 1e-12 * out[True][2]    # inertia: float64 atomics, order of arrival
    assert out[False][4] == out[True][4]


@pytest.mark.parametrize("K,D,n_iter,full_passes", [(8, 13, 14, 3), (5, 6, 10, 1), (7, 22, 9, 2), (3, 17, 9, 3)])
def test_kmeans_bounded_passes_equal_unbounded(P, K, D, n_iter, full_passes):
    """Delta passes behind Hamerly's bound test (rsx_kmeans_assign_bounded: pixels that provably keep their label are skipped,
    the others gathered from the pixel-interleaved copy) leave the same labels after EVERY pass, the same integer totals and
    centroids as the unbounded passes - and they do skip pixels."""
    import torch
    n_px = 30011 + K                                                       # ragged: not a multiple of 4 for some K
    rng = np.random.default_rng(100 * D + K)
    X = (rng.normal(size=(n_px, 3)) @ rng.normal(size=(3, D)) + 0.3 * rng.normal(size=(n_px, D))) * rng.uniform(0.01, 50.0, size=D) + rng.uniform(-5, 5, size=D)
    X = X.astype(np.float32)
    stride = (n_px + 31) // 32 * 32
    planes = torch.zeros((D, stride), dtype=torch.float32, device="cuda")
    planes[:, :n_px] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()

    class fr:                                                              # what DeviceKMeans needs of a FeatureResult
        pass
    fr.planes, fr.n_px, fr.W = planes, n_px, 257
    mn, mx = X.min(axis=0), X.max(axis=0)
    runs = {}
    for bounded in (False, True):
        km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W, bounded=bounded, full_passes=full_passes)
        c0 = km.scale_rows(X[P.draw_init_indices(fr.n_px, K, 5)])
        km.setup(c0)
        per_pass = []
        for _ in range(n_iter):
            mode = km.assign_pass()
            per_pass.append((km._cur_labels[:fr.n_px].clone(), km.acc[:K * D + K].clone(), int(km.acc[km.n_acc - 1])))
            km.update(mode)
        res = km._result(km.finish(True), n_iter)
        slack = km._slack[:fr.n_px].clone() if bounded else None
        runs[bounded] = (per_pass, res, km.acc[km.n_acc:km.n_acc + K * D + K].clone(), slack)
    for i, (a, b) in enumerate(zip(runs[False][0], runs[True][0])):
        assert torch.equal(a[0], b[0]), f"labels differ after pass {i}"
        assert torch.equal(a[1], b[1]), f"pass sums differ in pass {i}"
        assert a[2] == b[2], f"changed-label counter differs in pass {i}"
    assert torch.equal(runs[False][2], runs[True][2])
    assert np.array_equal(runs[False][1].centroids, runs[True][1].centroids)
    assert torch.equal(runs[False][1].labels, runs[True][1].labels)
    # the bound did something: after the last pass most pixels hold a slack above their label's drift (would be skipped next)
    assert float((runs[True][3] > 0).float().mean()) > 0.5


@pytest.mark.parametrize("K,D,n_iter,q_from", [(8, 13, 12, 3), (5, 6, 9, 1), (6, 22, 8, 2)])
def test_kmeans_16bit_screening_passes_equal_float32_passes(P, K, D, n_iter, q_from):
    """Delta passes that read the uint16 copy of the stack (rsx_kmeans_assign_q16; the float32 planes only for the pixels the
    quantisation cannot decide and for the samples that move) leave the same labels after every pass, the same integer totals
    and centroids as the float32 passes - a constant feature and a feature far from zero included."""
    import torch
    from rs_image_segmentation_b200 import _lib
    n_px = 40013 + K
    rng = np.random.default_rng(10 * D + K)
    X = (rng.normal(size=(n_px, 3)) @ rng.normal(size=(3, D)) + 0.3 * rng.normal(size=(n_px, D))) * rng.uniform(0.01, 50.0, size=D) + rng.uniform(-5, 5, size=D)
    X[:, 1] = 7.5
    X[:, 2] += 4000.0
    X = X.astype(np.float32)
    stride = (n_px + 31) // 32 * 32
    planes = torch.zeros((D, stride), dtype=torch.float32, device="cuda")
    planes[:, :n_px] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    mn, mx = X.min(axis=0), X.max(axis=0)
    runs = {}
    try:
        for q16 in (0, 1):
            _lib.set_option("km_q16", q16)
            _lib.set_option("km_q16_from", q_from)
            km = P.DeviceKMeans(planes, n_px, D, K, mn, mx, n_px, 257, full_passes=1)
            km.setup(km.scale_rows(X[P.draw_init_indices(n_px, K, 5)]))
            per_pass = []
            for _ in range(n_iter):
                mode = km.assign_pass()
                per_pass.append((km._cur_labels[:n_px].clone(), km.acc[:K * D + K].clone(), int(km.acc[km.n_acc - 1])))
                km.update(mode)
            res = km._result(km.finish(True), n_iter)
            runs[q16] = (per_pass, res, km.acc[km.n_acc:km.n_acc + K * D + K].clone(), km._q16 is not None)
    finally:
        _lib.set_option("km_q16", 0)
    assert runs[1][3] and not runs[0][3]
    for i, (a, b) in enumerate(zip(runs[0][0], runs[1][0])):
        assert torch.equal(a[0], b[0]), f"labels differ after pass {i}"
        assert torch.equal(a[1], b[1]), f"pass sums differ in pass {i}"
        assert a[2] == b[2], f"changed-label counter differs in pass {i}"
    assert torch.equal(runs[0][2], runs[1][2])
    assert np.array_equal(runs[0][1].centroids, runs[1][1].centroids)
    assert torch.equal(runs[0][1].labels, runs[1][1].labels)


@pytest.mark.parametrize("D,K", [(13, 8), (7, 5), (22, 32)])
def test_kmeans_device_setup_equals_host_setup(P, D, K):
    """rsx_kmeans_setup_device (range from the device min/max trackers, initial centroids scaled by the kernel) leaves the same
    state, byte for byte, as rsx_kmeans_setup fed with the numbers the host derives - constant features and features far from
    zero included."""
    import torch
    from rs_image_segmentation_b200.device import MinMaxTracker, ptr, stream_ptr
    from rs_image_segmentation_b200 import _lib
    n_px = 20011
    rng = np.random.default_rng(D * 100 + K)
    X = (rng.normal(size=(n_px, D)) * rng.uniform(1e-3, 300.0, size=D) + rng.uniform(-1000, 1000, size=D)).astype(np.float32)
    X[:, 1] = 0.25                                                         # constant feature: range 0 -> scale 1
    X[:, 2] = np.where(rng.random(n_px) < 0.5, -3.0, -3.0 + 1e-7).astype(np.float32)
    stride = (n_px + 31) // 32 * 32
    planes = torch.zeros((D, stride), dtype=torch.float32, device="cuda")
    planes[:, :n_px] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    mm = MinMaxTracker(D)
    _lib.call("rsx_minmax_planes_f32", ptr(planes), n_px, stride, D, ptr(mm.buf), stream_ptr())
    mn, mx = mm.read()
    assert np.array_equal(mn, X.min(axis=0)) and np.array_equal(mx, X.max(axis=0))
    idx = P.draw_init_indices(n_px, K, 3)
    a = P.DeviceKMeans(planes, n_px, D, K, mn, mx, 40000 * 40000, 257)
    a.setup(a.scale_rows(X[idx]))
    b = P.DeviceKMeans(planes, n_px, D, K, None, None, 40000 * 40000, 257)
    b.setup_device(mm.merged(), torch.from_numpy(X[idx].astype(np.float64)).cuda().contiguous())
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state)
    b.read()
    b.configure_from_device()
    assert np.array_equal(a.fmin, b.fmin) and np.array_equal(a.scale, b.scale) and np.array_equal(a.min_, b.min_)


def test_kmeans_changed_counter(P):
    """The changed-label counter of an update pass equals the number of labels that differ from the previous pass."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(90, 131, 7, np.uint8, 8, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm=False))
    D, K = 7, 6
    mn, mx = fr.minmax.read()
    km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W)
    km.setup(km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 2), 0)))
    prev = None
    for it in range(4):
        km.step(track_labels=True)
        cur = km._labels[it % 2][:fr.n_px].cpu().numpy().copy()
        expect = fr.n_px if prev is None else int((cur != prev).sum())
        assert km.changed_count() == expect
        prev = cur


@pytest.mark.parametrize("dups", [1, 2, 3])
def test_kmeans_empty_cluster_relocation_matches_sklearn(P, dups):
    """Duplicate initial centroids leave cluster(s) empty after the first E-step: sklearn relocates each to the sample that
    is farthest from its own centre (_k_means_common.pyx:167-211).  Same labels / centroids as sklearn on float64 data."""
    import warnings
    import torch
    from sklearn.cluster import KMeans
    from oracle import kmeans as ok
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(80, 120, 7, np.uint8, 13, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm=False))
    D, K = 7, 5
    mn, mx = fr.minmax.read()
    km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W)
    c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 4), 0))
    for j in range(dups):
        c0[K - 1 - j] = c0[0]                                          # first minimum wins -> these clusters start empty
    X = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy().astype(np.float64)
    Xs = ok.minmax_scale(X)
    tol = float(np.mean(np.var(Xs, axis=0)) * 1e-4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = KMeans(n_clusters=K, init=c0, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(Xs)
    res = km.fit_converge(c0, max_iter=300, tol=tol, mean_scaled=Xs.mean(axis=0))
    got = res.labels.cpu().numpy()
    # one or several simultaneous empties: the far samples are handed out in sklearn's own order (np.argpartition of the
    # float64 distances, reproduced on the host), so labels and centroids are sklearn's
    assert res.n_iter == ref.n_iter_
    assert np.array_equal(got, ref.labels_), f"{(got != ref.labels_).sum()} labels differ"
    np.testing.assert_allclose(res.centroids, ref.cluster_centers_, rtol=0, atol=1e-9)
    assert len(np.unique(got)) == K


def test_kmeans_fixed_iterations_relocate_empty_clusters(P):
    """VERDICT r1 missing 6: the fixed-iteration protocol (DeviceKMeans.fit) with three duplicated initial centroids - three
    clusters start empty - must not raise: it reruns with relocation and equals sklearn's fixed-iteration run."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(90, 130, 7, np.uint8, 17, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm=False))
    D, K, T = 8, 6, 5
    mn, mx = fr.minmax.read()
    km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], fr.n_px, fr.W)
    c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 4), 0))
    c0[3] = c0[4] = c0[5] = c0[1]
    res = km.fit(c0, T)
    stack = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy()
    (lab, cent, inertia, n_run), _ = _kmeans_oracle(stack, c0, T, km.fmin, km.fmax)
    got = res.labels.cpu().numpy()
    assert n_run == T and len(np.unique(got)) == K
    assert np.array_equal(got, lab), f"{(got != lab).sum()} labels differ"
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)


def test_stage1_fused_into_load_path(P):
    """N3: raw 8-bit DNs + FeatureConfig.stage1 == the reference's stage 1 (gain/bias, min-max stretch, uint8) followed by
    the normal path: every plane, the quantised band and the PCA are bit-identical."""
    import torch
    from oracle import features as of
    from rs_image_segmentation_b200 import hoststats
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    raw = synth_raster_numpy(111, 173, 7, np.uint8, 31, cell=16)
    raw[..., 2] = (raw[..., 2] // 3) + 20                                # a band that does not span 0..255
    stage1 = np.stack(of.stage1_preprocess([raw[..., b] for b in range(7)]), axis=-1)
    cfg = P.FeatureConfig(glcm_window=7, glcm_step=1)
    ref = P.extract_features(torch.from_numpy(np.ascontiguousarray(stage1)).cuda(), cfg)
    cfg1 = P.FeatureConfig(glcm_window=7, glcm_step=1, stage1=(hoststats.TM_GAIN, hoststats.TM_BIAS))
    got = P.extract_features(torch.from_numpy(raw).cuda(), cfg1)
    assert torch.equal(got.quant[:got.n_px], ref.quant[:ref.n_px])
    assert np.array_equal(got.pca["components"], ref.pca["components"])
    for name in ref.names:
        assert torch.equal(got.plane(name), ref.plane(name)), name
    remap, hist1 = hoststats.stage1_level_tables(np.stack([np.bincount(raw[..., b].ravel(), minlength=256) for b in range(7)]))
    for b in range(7):
        assert np.array_equal(remap[b][raw[..., b]], stage1[..., b])
        assert np.array_equal(hist1[b], np.bincount(stage1[..., b].ravel(), minlength=256))


def test_device_writers(P, tmp_path):
    """N4: the .npy file produced from the device stack is byte-identical to np.save of the float64 (H, W, C) array, and the
    label map is kmeans_result + 1 as uint8."""
    import io
    import torch
    from rs_image_segmentation_b200 import writers
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    H, W = 67, 93
    bip = synth_raster_numpy(H, W, 7, np.uint8, 2, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm_window=5, glcm_step=1))
    for C in (1, 7, len(fr.names)):
        path = str(tmp_path / f"stack{C}.npy")
        writers.save_stack_npy(path, fr.planes, fr.n_px, H, W, C)
        ref = np.moveaxis(fr.planes[:C, :fr.n_px].cpu().numpy().reshape(C, H, W), 0, -1).astype(np.float64)
        buf = io.BytesIO()
        np.save(buf, ref)
        assert open(path, "rb").read() == buf.getvalue()
        assert np.array_equal(np.load(path), ref)
    res, km, c0 = P.kmeans_on_features(fr, 13, 6, 3, seed=1)
    lab8 = writers.labels_for_geotiff(res.labels, H, W)
    assert lab8.dtype == np.uint8 and np.array_equal(lab8, (res.labels.cpu().numpy().reshape(H, W) + 1).astype(np.uint8))


@pytest.mark.parametrize("H,W,win,step,K", [(9, 9, 7, 1, 2), (13, 17, 5, 1, 3), (8, 31, 7, 7, 4), (21, 21, 21, 21, 2), (11, 514, 3, 1, 5), (7, 7, 7, 1, 1)])
def test_ragged_and_tiny_rasters(P, H, W, win, step, K):
    """Tiny / ragged shapes: n_px not a multiple of 4 or of the 512-pixel KMeans block, a single GLCM window, W below one tile."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(H, W, 7, np.uint8, H * W, cell=4)
    cfg = P.FeatureConfig(glcm_window=win, glcm_step=step)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), cfg)
    _check_features(fr, _oracle_features(bip, cfg), cfg)
    _mm_check(fr)
    D, n_iter = 13, 3
    res, km, c0 = P.kmeans_on_features(fr, D, K, n_iter, seed=2)
    stack = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy()
    (lab, cent, inertia, n_run), Xs = _kmeans_oracle(stack, c0, n_iter, km.fmin, km.fmax)
    got = res.labels.cpu().numpy()
    if n_run == n_iter:                                   # sklearn may stop early on tiny inputs (strict convergence)
        assert np.array_equal(got, lab)
        np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)
    assert int(km.acc[km.n_acc + K * D:km.n_acc + K * D + K].sum()) == fr.n_px


def test_segment_stream_matches_single_scene_calls(P):
    """The pipelined host-buffer API (uploads / downloads on the copy engines under the neighbouring scenes' kernels) returns,
    scene by scene, exactly what the one-scene call returns."""
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    cfg = P.FeatureConfig(glcm_window=7, glcm_step=1)
    scenes = [synth_raster_numpy(90, 150, 7, np.uint8, 100 + i, cell=16) for i in range(5)]
    ref = [P.segment_raster(s, cfg, 6, 5, 9)[0].copy() for s in scenes]
    for mode, dtype in (("int32", np.int32), ("uint8", np.uint8), ("int32_host_widen", np.int32)):
        got = []
        for labels, res in P.segment_stream(scenes, cfg, 6, 5, 9, labels=mode):
            got.append(labels.copy())                      # a yielded buffer is reused two scenes later
            assert res.n_iter == 5
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert a.dtype == dtype and np.array_equal(a, b), mode
    assert list(P.segment_stream([], cfg)) == []
    with pytest.raises(ValueError):
        list(P.segment_stream(scenes[:1], cfg, labels="int64"))


def test_full_size_scene_properties(P):
    """BASELINE config B at full size (7000 x 7000 x 7, 49 Mpx): size-independent properties instead of an oracle run -
    histogram totals, bit-exact index planes on a pixel sample, integer totals consistent with the labels, delta passes ==
    full passes, and every sampled label is the float64 argmin for the final centroids."""
    import torch
    from oracle import features as of
    from rs_image_segmentation_b200.synth import synth_strip_torch
    H = W = 7000
    n = H * W
    raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
    fr = P.extract_features(raster, P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32))
    assert all(int(fr.stats.hist[b].sum()) == n for b in range(7))
    rng = np.random.default_rng(0)
    idx = np.unique(np.concatenate([rng.integers(0, n, 4000), [0, n - 1, W - 1, n - W]]))
    tidx = torch.from_numpy(idx).cuda()
    rows = raster.view(-1, 7)[tidx].cpu().numpy().astype(np.float32)
    norm = fr.stats.norm
    nb = [((np.clip(rows[:, b], norm[b, 0], norm[b, 1]) - norm[b, 0]) / norm[b, 2]).reshape(-1, 1) for b in range(7)]
    ix = of.all_indices(nb)
    for k in P.INDEX_NAMES:
        assert np.array_equal(fr.planes[fr.names.index(k)][tidx].cpu().numpy(), ix[k].ravel()), k
    D, K, T = 13, 8, 6
    out = {}
    for delta in (True, False):
        res, km, c0 = P.kmeans_on_features(fr, D, K, T, seed=7000, delta=delta)
        tot = km.acc[km.n_acc:km.n_acc + K * D + K].cpu().numpy()
        assert int(tot[K * D:].sum()) == n
        assert int(res.labels.min()) >= 0 and int(res.labels.max()) < K
        out[delta] = (tot, res.centroids, res.inertia, res.labels[tidx].cpu().numpy(), km)
    assert np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1])
    assert np.array_equal(out[True][3], out[False][3])
    # sampled labels against a float64 argmin with the final centroids (MinMax-scaled coordinates)
    km = out[True][4]
    X = fr.planes[:D][:, tidx].t().to(torch.float64).cpu().numpy() * km.scale + km.min_
    d2 = ((X[:, None, :] - out[True][1][None, :, :]) ** 2).sum(-1)
    best = d2.argmin(1)
    gap = np.partition(d2, 1, axis=1)
    clear = (gap[:, 1] - gap[:, 0]) > 1e-9                       # leave exact near-ties to the kernel's own float64 rule
    assert np.array_equal(best[clear], out[True][3][clear])


# ============================================================================================ round-2 additions
def _nan_raster(H=150, W=203, seed=0):
    """uint16 x 13 raster whose red band normalises to exactly 0 for most pixels and whose NIR band normalises to a fine grid of
    values around 0.5: there the float32 radicand (2n+1)^2 - 8(n-r) of MSAVI (indices.py:109-112) rounds below zero for ~0.5 %
    of the values and np.sqrt returns NaN (the case the reference handles with its NaN -> 0 before KMeans, extract.py:548-556)."""
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    rng = np.random.default_rng(seed)
    bip = synth_raster_numpy(H, W, 13, np.uint16, seed, cell=16)
    u = rng.random((H, W))
    bip[:, :, 3] = np.where(u < 0.9, 100, rng.integers(100, 3000, (H, W)))                       # red: P2 == 100 -> 0 after the clip
    v = rng.random((H, W))
    bip[:, :, 7] = np.where(v < 0.04, 0, np.where(v > 0.96, 60000, rng.integers(29000, 31000, (H, W))))   # nir: n in 0.483 .. 0.517
    return bip


def test_msavi_nan_is_replaced_before_kmeans(P):
    """a NaN sample must not reach the fixed-point sums.  The feature plane keeps the reference's NaN; the
    KMeans entry replaces it by 0 (and 0 joins the MinMax range), and labels / centroids equal the oracle run on nan_to_num."""
    import torch
    from oracle import features as of
    bip = _nan_raster()
    cfg = P.FeatureConfig(band_map=(1, 2, 3, 7, 11), n_components=6, glcm=False)
    fr = P.extract_features(torch.from_numpy(bip.view(np.int16)).cuda(), cfg)
    with np.errstate(all="ignore"):
        ref = _oracle_features(bip, cfg)
    n_nan = int(np.isnan(ref["msavi"]).sum())
    assert n_nan > 10, "the test raster must produce NaN in the oracle's MSAVI"
    for k in of.INDEX_ORDER:
        assert np.array_equal(fr.plane(k).cpu().numpy(), ref[k], equal_nan=True), k
    D, K, T = 13, 6, 5
    res, km, c0 = P.kmeans_on_features(fr, D, K, T, seed=9)
    ms = fr.plane("msavi").cpu().numpy()
    assert not np.isnan(ms).any() and int((ms == 0).sum()) >= n_nan
    assert np.array_equal(ms, np.nan_to_num(ref["msavi"], nan=0.0))
    stack = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy()
    assert km.fmin[2] == stack[:, 2].min() and km.fmax[2] == stack[:, 2].max() and km.fmin[2] <= 0.0 <= km.fmax[2]
    (lab, cent, inertia, n_run), Xs = _kmeans_oracle(stack, c0, T, km.fmin, km.fmax)
    assert np.array_equal(res.labels.cpu().numpy(), lab)
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)
    assert np.isfinite(res.centroids).all() and abs(res.inertia - inertia) <= 1e-5 * inertia


def test_whole_path_2048_against_oracle(P):
    """The whole path at 2048 x 2048 (4.2 Mpx: multi-CTA GLCM bands, several histogram flushes, 8k KMeans blocks) against the
    oracle: indices + quantised band bit exact, GLCM properties 1e-5, PCA 1e-5 (float64 PCA of the same data), labels of the
    fixed-iteration protocol bit exact vs sklearn on the float64 promotion of the GPU's stack, inertia 1e-5."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    H = W = 2048
    bip = synth_raster_numpy(H, W, 7, np.uint8, 2048)
    cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), cfg)
    _check_features(fr, _oracle_features(bip, cfg), cfg)
    D, K, T = 13, 8, 6
    res, km, c0 = P.kmeans_on_features(fr, D, K, T, seed=2048)
    stack = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy()
    (lab, cent, inertia, n_run), _ = _kmeans_oracle(stack, c0, T, km.fmin, km.fmax)
    got = res.labels.cpu().numpy()
    assert np.array_equal(got, lab), f"{(got != lab).sum()} labels differ"
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)
    assert abs(res.inertia - inertia) <= 1e-5 * inertia


@pytest.mark.parametrize("K", [8, 32])
def test_kmeans_fixed_point_budget_of_the_40k_mosaic(P, K):
    """The int64 fixed-point shift depends on the GLOBAL pixel count (rsx_kmeans.cu: 62 - ceil(log2 n) bits per sample).  With
    n_global forced to 1.6e9 (config E: 31 bits) the labels on a small stack must still equal sklearn's float64 run."""
    import torch
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    bip = synth_raster_numpy(160, 300, 7, np.uint8, 77 + K, cell=16)
    fr = P.extract_features(torch.from_numpy(bip).cuda(), P.FeatureConfig(glcm_window=7, glcm_step=1))
    D, T = 13, 6
    mn, mx = fr.minmax.read()
    km = P.DeviceKMeans(fr.planes, fr.n_px, D, K, mn[:D], mx[:D], 40000 * 40000, fr.W)
    c0 = km.scale_rows(km.gather_rows(P.draw_init_indices(fr.n_px, K, 6), 0))
    res = km.fit(c0, T)
    stack = fr.planes[:D, :fr.n_px].t().contiguous().cpu().numpy()
    (lab, cent, inertia, n_run), _ = _kmeans_oracle(stack, c0, T, km.fmin, km.fmax)
    got = res.labels.cpu().numpy()
    assert np.array_equal(got, lab), f"{(got != lab).sum()} labels differ"
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=5e-9)           # sample quantum 2^-31 of the feature range
    assert abs(res.inertia - inertia) <= 1e-5 * inertia


@pytest.mark.parametrize("D,K", [(22, 7), (24, 16), (21, 40)])
def test_kmeans_depth_21_to_24(P, D, K):
    """The reference's own KMeans call site stacks ndvi, ndwi, ndbi + the 19 hierarchical channels = 22 planes
    (scripts/3_classification.py:381-391)."""
    import torch
    n_px = 9001
    rng = np.random.default_rng(100 * D + K)
    X = (rng.normal(size=(n_px, 4)) @ rng.normal(size=(4, D)) + 0.3 * rng.normal(size=(n_px, D))).astype(np.float32)
    stride = (n_px + 31) // 32 * 32
    planes = torch.zeros((D, stride), dtype=torch.float32, device="cuda")
    planes[:, :n_px] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    fmin, fmax = X.min(axis=0), X.max(axis=0)
    km = P.DeviceKMeans(planes, n_px, D, K, fmin, fmax, n_px, 257)
    c0 = km.scale_rows(X[P.draw_init_indices(n_px, K, 5)])
    res = km.fit(c0, 4)
    (lab, cent, inertia, n_run), _ = _kmeans_oracle(X, c0, 4, fmin, fmax)
    assert np.array_equal(res.labels.cpu().numpy(), lab)
    np.testing.assert_allclose(res.centroids, cent, rtol=0, atol=1e-9)
