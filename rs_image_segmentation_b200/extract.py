"""Drop-in for `unsupervised_kmeans_classification` of the reference's modules/features/extract.py:508-581.

Same signature, same key-selection rules, same exceptions, same (H, W) int32 label image.  The stack is built on
the device as planar float32, MinMaxScaler.fit is a min/max reduction kernel, and KMeans(n_clusters,
random_state=42, n_init='auto') is reproduced with sklearn's own control flow:

  * k-means++ seeding (sklearn/cluster/_kmeans.py:180-255) driven by numpy's RandomState(42) on the host - the
    same random draws as the reference - with the distance / potential / cumulative-sum / search steps on the
    device in float64;
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
from .device import MinMaxTracker, ptr, require_cuda, stream_ptr
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


def _kmeans_plusplus(Xc: torch.Tensor, n_clusters: int, rs: np.random.RandomState) -> np.ndarray:
    """sklearn _kmeans_plusplus on centred float64 data Xc (N, D) living on the device; returns centre indices."""
    n = Xc.shape[0]
    n_local_trials = 2 + int(np.log(n_clusters))
    # first centre: RandomState.choice(n, p=uniform) = one random_sample() searched in the cdf of p
    if n <= (1 << 27):
        first = int(rs.choice(n, p=np.full(n, 1.0) / np.full(n, 1.0).sum()))
    else:                                                              # same stream consumption, closed-form cdf
        first = min(int(rs.random_sample() * n), n - 1)
    xsq = (Xc * Xc).sum(dim=1)

    def dist_to(idx: torch.Tensor) -> torch.Tensor:
        """_euclidean_distances(X[idx], X, squared=True): -2 X.Y^T + |x|^2 + |y|^2, clipped at 0."""
        C = Xc[idx]
        d = -2.0 * (C @ Xc.t())
        d += xsq[idx][:, None]
        d += xsq[None, :]
        return torch.clamp_(d, min=0.0)

    indices = [first]
    closest = dist_to(torch.tensor([first], device=Xc.device))[0]
    pot = float(closest.sum().item())
    for _ in range(1, n_clusters):
        rand_vals = rs.uniform(size=n_local_trials) * pot
        cum = torch.cumsum(closest, dim=0)
        cand = torch.searchsorted(cum, torch.from_numpy(rand_vals).to(Xc.device))
        cand.clamp_(max=n - 1)
        d = torch.minimum(closest[None, :], dist_to(cand))
        pots = d.sum(dim=1)
        best = int(torch.argmin(pots).item())
        pot = float(pots[best].item())
        closest = d[best]
        indices.append(int(cand[best].item()))
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
    # MinMaxScaler + the centring of KMeans.fit, float64 on the device (seeding only; the Lloyd kernels read `planes`)
    scale, min_ = minmax_scale_params(fmin, fmax)
    Xs = planes[:, :n].t().to(torch.float64) * torch.from_numpy(scale).cuda() + torch.from_numpy(min_).cuda()
    mean = Xs.mean(dim=0)
    tol = float(Xs.var(dim=0, unbiased=False).mean().item()) * 1e-4           # _tolerance(X, 1e-4)
    Xs -= mean
    print(f"正在进行K-Means聚类，目标簇数: {n_clusters}...")
    idx = _kmeans_plusplus(Xs, n_clusters, np.random.RandomState(42))
    mean_h = mean.cpu().numpy()
    c0 = Xs[torch.from_numpy(idx).cuda()].cpu().numpy() + mean_h               # scaled, un-centred coordinates
    del Xs
    km = DeviceKMeans(planes, n, D, n_clusters, fmin, fmax, n, W)
    res = km.fit_converge(c0, max_iter=300, tol=tol, mean_scaled=mean_h)
    print("K-Means聚类完成。")
    return res.labels.cpu().numpy().reshape(shape)
