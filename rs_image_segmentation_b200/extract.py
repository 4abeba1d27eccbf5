"""Drop-in for `unsupervised_kmeans_classification` of the reference's modules/features/extract.py:508-581.

Same signature, same key-selection rules, same exceptions, same (H, W) int32 label image.  The stack is built on
the device as planar float32, MinMaxScaler.fit is a min/max reduction kernel, and KMeans(n_clusters,
random_state=42, n_init='auto') is reproduced with sklearn's own control flow:

  * k-means++ seeding (sklearn/cluster/_kmeans.py:180-255) driven by numpy's RandomState(42) on the host - the
    same random draws as the reference - with the distance / potential / block-sum steps in float64 kernels on the
    planar stack itself (rsx_kpp_*: 8 B of state per sample, no float64 copy of the stack);
  * Lloyd iterations through the rsx KMeans kernels with sklearn's stopping rules (strict label convergence or
    squared centre shift <= tol * mean feature variance, max_iter 300) and the final assignment pass.

Parity protocol: labels are bit-exact against the reference run on the float64 promotion of the same features
(the reference's real stack is float64).  For float32 features sklearn's own result depends on
BLAS summation order (float32 potentials and cumulative sums inside k-means++), so exact equality with a given
sklearn build is not defined there; the golden fixture of tests/ was produced in float64 for that reason.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .device import MinMaxTracker, hptr, ptr, require_cuda, stream_ptr
from .pipeline import DeviceKMeans, minmax_scale_params

__all__ = ["unsupervised_kmeans_classification"]

_META_KEYS = ("transform", "crs", "width", "height", "dimensions", "geo_transform")


def _select_planes(features_dict, feature_keys_to_use):
    """extract.py:510-566: returns the list of (H*W,) host arrays to stack, in the reference's order."""
    if not features_dict or "height" not in features_dict or "width" not in features_dict:
        raise ValueError("特征字典为空或缺少图像尺寸信息 (height/width)。")
    shape = (features_dict["height"], features_dict["width"])
    if feature_keys_to_use is None:
        keys = [k for k, v in features_dict.items()
                if isinstance(v, np.ndarray) and v.ndim == 2 and v.shape == shape and k not in _META_KEYS]
        if not keys:
            cand = ["ndvi", "ndwi", "ndbi", "texture_mean", "evi", "savi", "hierarchical_level_1", "hierarchical_level_2", "hierarchical_all"]
            keys = [k for k in cand if isinstance(features_dict.get(k), np.ndarray)
                    and ((features_dict[k].ndim == 2 and features_dict[k].shape == shape)
                         or (features_dict[k].ndim == 3 and features_dict[k].shape[:2] == shape))]
        feature_keys_to_use = keys
    if not feature_keys_to_use:
        raise ValueError("没有可用于K-Means的特征。请检查特征字典内容或手动指定 `feature_keys_to_use`。")
    print(f"用于K-Means的特征: {feature_keys_to_use}")
    planes = []
    for key in feature_keys_to_use:
        v = features_dict.get(key)
        if v is None:
            print(f"Severe Warning: Feature '{key}' does not exist or is None in the dictionary, despite being in feature_keys_to_use.")
        elif isinstance(v, np.ndarray) and v.ndim == 3 and v.shape[:2] == shape:
            planes.extend(v[:, :, i] for i in range(v.shape[2]))
        elif isinstance(v, np.ndarray) and v.ndim == 2 and v.shape == shape:
            planes.append(v)
        else:
            print(f"Warning: Feature '{key}' has unexpected shape {v.shape if isinstance(v, np.ndarray) else 'N/A'} for K-Means, skipping.")
    if not planes:
        raise ValueError("未能准备任何特征数据进行K-Means分类。")
    return planes, shape


class _Seeder:
    """sklearn's _kmeans_plusplus (sklearn/cluster/_kmeans.py:180-255) on the planar float32 stack: numpy's RandomState on the host
    (the reference's random stream), distances / potentials / block sums in float64 kernels (rsx_kpp_*).  Device state: the
    running closest squared distance, 8 B per sample - no float64 copy of the (N, D) matrix, no (trials, N) distance matrix."""

    def __init__(self, planes: torch.Tensor, n: int, D: int, scale: np.ndarray, min_: np.ndarray):
        self.planes, self.n, self.D = planes, int(n), int(D)
        self.stride = planes.stride(0)
        self.scale = np.ascontiguousarray(scale, np.float64)
        self.min_ = np.ascontiguousarray(min_, np.float64)
        lib = _lib.load()
        self.block = int(lib.rsx_kpp_block())
        self.scratch = torch.empty(int(lib.rsx_kpp_scratch_elems()), dtype=torch.float64, device="cuda")
        self.pot = torch.zeros(8, dtype=torch.float64, device="cuda")
        self.closest = torch.empty(self.n, dtype=torch.float64, device="cuda")
        self.sums = torch.empty((self.n + self.block - 1) // self.block, dtype=torch.float64, device="cuda")
        self.mean = np.zeros(D, np.float64)

    def moments(self):
        """(mean, mean of the per-feature variances) of the MinMax-scaled stack."""
        out = torch.empty(2 * self.D, dtype=torch.float64, device="cuda")
        _lib.call("rsx_kpp_feature_moments", ptr(self.planes), self.stride, self.n, self.D, hptr(self.scale), hptr(self.min_), ptr(out),
                  ptr(self.scratch), stream_ptr())
        m = out.cpu().numpy().reshape(self.D, 2)
        mean = m[:, 0] / self.n
        var = np.maximum(m[:, 1] / self.n - mean * mean, 0.0)
        self.mean = np.ascontiguousarray(mean)
        return self.mean, float(var.mean())

    def rows(self, idx) -> np.ndarray:
        """Centred, scaled float64 coordinates of the given samples."""
        sel = torch.as_tensor(np.asarray(idx, dtype=np.int64), device="cuda")
        raw = self.planes[:self.D].index_select(1, sel).t().to(torch.float64).cpu().numpy()
        return (raw * self.scale + self.min_) - self.mean

    def _distances(self, cand: np.ndarray, mode: int) -> np.ndarray:
        cand = np.ascontiguousarray(cand, np.float64)
        _lib.call("rsx_kpp_distances", ptr(self.planes), self.stride, self.n, self.D, hptr(self.scale), hptr(self.min_), hptr(self.mean),
                  hptr(cand), cand.shape[0], mode, ptr(self.closest), ptr(self.pot), ptr(self.scratch), stream_ptr())
        return self.pot.cpu().numpy()

    def _search(self, vals: np.ndarray) -> np.ndarray:
        """np.searchsorted(stable_cumsum(closest), vals): the block whose running sum reaches the value, then a sequential
        float64 cumulative sum inside that block."""
        _lib.call("rsx_kpp_block_sums", ptr(self.closest), self.n, ptr(self.sums), stream_ptr())
        prefix = np.cumsum(self.sums.cpu().numpy())
        out = np.empty(len(vals), np.int64)
        for k, v in enumerate(vals):
            b = min(int(np.searchsorted(prefix, v)), len(prefix) - 1)
            a0 = b * self.block
            seg = self.closest[a0:min(self.n, a0 + self.block)].cpu().numpy()
            run = np.cumsum(np.concatenate(([prefix[b - 1] if b else 0.0], seg)))[1:]
            out[k] = a0 + min(int(np.searchsorted(run, v)), len(seg) - 1)
        return np.minimum(out, self.n - 1)

    def seed(self, n_clusters: int, rs: np.random.RandomState) -> np.ndarray:
        n = self.n
        n_local_trials = 2 + int(np.log(n_clusters))
        # first centre: RandomState.choice(n, p=uniform) = one random_sample() searched in the cdf of p
        if n <= (1 << 27):
            first = int(rs.choice(n, p=np.full(n, 1.0) / np.full(n, 1.0).sum()))
        else:                                                              # same stream consumption, closed-form cdf
            first = min(int(rs.random_sample() * n), n - 1)
        indices = [first]
        pot = float(self._distances(self.rows([first]), 0)[0])
        for _ in range(1, n_clusters):
            rand_vals = rs.uniform(size=n_local_trials) * pot
            cand = self._search(rand_vals)
            pots = self._distances(self.rows(cand), 1)[:n_local_trials]
            best = int(np.argmin(pots))
            pot = float(pots[best])
            self._distances(self.rows([cand[best]]), 2)
            indices.append(int(cand[best]))
        return np.asarray(indices, dtype=np.int64)


def unsupervised_kmeans_classification(features_dict, n_clusters=5, feature_keys_to_use=None):
    """使用K-Means进行无监督分类 (extract.py:508-581)."""
    planes_host, shape = _select_planes(features_dict, feature_keys_to_use)
    require_cuda()
    H, W = shape
    n = H * W
    D = len(planes_host)
    if D > 24:
        raise _lib.RsxError(f"{D} feature planes; the KMeans kernels are compiled for up to 24")
    if n_clusters > 64:
        raise _lib.RsxError("n_clusters > 64 is not supported")
    stride = (n + 31) // 32 * 32
    planes = torch.zeros((D, stride), dtype=torch.float32, device="cuda")
    for i, p in enumerate(planes_host):
        planes[i, :n] = torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32).ravel()).cuda()
    st = stream_ptr()
    _lib.call("rsx_nan_to_zero_f32", ptr(planes), planes.numel(), st)          # extract.py:548-556
    mm = MinMaxTracker(D)
    _lib.call("rsx_minmax_planes_f32", ptr(planes), n, stride, D, ptr(mm.buf), st)
    fmin, fmax = mm.read()
    # MinMaxScaler + the centring and tolerance of KMeans.fit from float64 moment kernels; k-means++ on the stack itself
    scale, min_ = minmax_scale_params(fmin, fmax)
    seeder = _Seeder(planes, n, D, scale, min_)
    mean_h, mean_var = seeder.moments()
    tol = mean_var * 1e-4                                                      # _tolerance(X, 1e-4)
    print(f"正在进行K-Means聚类，目标簇数: {n_clusters}...")
    idx = seeder.seed(n_clusters, np.random.RandomState(42))
    c0 = seeder.rows(idx) + mean_h                                             # scaled, un-centred coordinates
    del seeder
    km = DeviceKMeans(planes, n, D, n_clusters, fmin, fmax, n, W)
    res = km.fit_converge(c0, max_iter=300, tol=tol, mean_scaled=mean_h)
    print("K-Means聚类完成。")
    return res.labels.cpu().numpy().reshape(shape)
