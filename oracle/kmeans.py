"""Oracle (test infrastructure): the KMeans classification step.

Restates modules/features/extract.py:508-581 (key selection, flatten, NaN->0,
MinMaxScaler, sklearn KMeans(random_state=42, n_init='auto')) and, for the benchmark
protocol of SURVEY.md 8(d), the same Lloyd loop driven from given initial centroids for
a fixed number of iterations (sklearn ``_kmeans_single_lloyd`` with tol=0: T fused
assign+update passes, then the extra assignment pass of sklearn/cluster/_kmeans.py:742-754,
then the inertia of _k_means_common.pyx:94-124).
"""
from __future__ import annotations

import warnings

import numpy as np

_META_KEYS = ("transform", "crs", "width", "height", "dimensions", "geo_transform")


def stack_from_dict(features, keys=None):
    """extract.py:510-568 - returns the (N, D) matrix handed to MinMaxScaler."""
    if not features or "height" not in features or "width" not in features:
        raise ValueError("feature dict is empty or lacks height/width")
    shape = (features["height"], features["width"])
    if keys is None:
        keys = [k for k, v in features.items()
                if isinstance(v, np.ndarray) and v.ndim == 2 and v.shape == shape and k not in _META_KEYS]
        if not keys:
            cand = ["ndvi", "ndwi", "ndbi", "texture_mean", "evi", "savi",
                    "hierarchical_level_1", "hierarchical_level_2", "hierarchical_all"]
            keys = [k for k in cand if isinstance(features.get(k), np.ndarray)
                    and (features[k].shape == shape or (features[k].ndim == 3 and features[k].shape[:2] == shape))]
    if not keys:
        raise ValueError("no usable K-Means features")
    cols = []
    for k in keys:
        v = features.get(k)
        if isinstance(v, np.ndarray) and v.ndim == 3 and v.shape[:2] == shape:
            planes = [v[:, :, i].ravel() for i in range(v.shape[2])]
        elif isinstance(v, np.ndarray) and v.ndim == 2 and v.shape == shape:
            planes = [v.ravel()]
        else:
            continue
        for p in planes:
            cols.append(np.nan_to_num(p, nan=0.0) if np.isnan(p).any() else p)
    if not cols:
        raise ValueError("nothing could be stacked for K-Means")
    return np.vstack(cols).T


def kmeans_classification(features, n_clusters=5, keys=None):
    """extract.py:508-581 end to end; returns the (H, W) int32 label image."""
    from sklearn.cluster import KMeans
    from sklearn.preprocessing import MinMaxScaler

    X = MinMaxScaler().fit_transform(stack_from_dict(features, keys))
    labels = KMeans(n_clusters=n_clusters, random_state=42, n_init="auto", verbose=0).fit_predict(X)
    return labels.reshape(features["height"], features["width"])


def minmax_scale(X):
    """sklearn/preprocessing/_data.py:527-541,574-575: X*scale + min_ in X's dtype."""
    from sklearn.preprocessing import MinMaxScaler

    return MinMaxScaler().fit_transform(X)


def lloyd_fixed(X_scaled, init_centroids, n_iter, n_threads=None):
    """Benchmark protocol: Lloyd from given centroids (expressed in MinMax-scaled
    coordinates), exactly n_iter update passes unless sklearn detects strict convergence.

    Runs sklearn's own KMeans.fit so that the centring of _kmeans.py:1488-1493, the chunked
    E/M step, empty-cluster relocation and the final E-step are the library's.  Returns
    (labels int32, centroids in scaled coordinates, inertia, n_iter_run).
    """
    from sklearn.cluster import KMeans

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=init_centroids.shape[0], init=np.asarray(init_centroids, dtype=X_scaled.dtype),
                    n_init=1, max_iter=n_iter, tol=0.0, algorithm="lloyd")
        km.fit(X_scaled)
    return km.labels_.astype(np.int32), km.cluster_centers_, float(km.inertia_), int(km.n_iter_)


def lloyd_numpy(X_scaled, init_centroids, n_iter):
    """Independent float64 restatement of the same loop (no sklearn), used to cross-check
    lloyd_fixed on small inputs: _k_means_lloyd.pyx:168-218 (argmin of |c|^2 - 2 x.c, first
    minimum wins), _k_means_common.pyx:274-296 (multiply by 1/weight)."""
    X = np.asarray(X_scaled, dtype=np.float64)
    mu = X.mean(axis=0)
    Xc = X - mu
    C = np.asarray(init_centroids, dtype=np.float64) - mu
    labels_old = None
    strict = False
    it = 0
    for it in range(1, n_iter + 1):
        d = (C * C).sum(axis=1)[None, :] - 2.0 * Xc @ C.T
        labels = d.argmin(axis=1).astype(np.int32)
        Cn = np.zeros_like(C)
        np.add.at(Cn, labels, Xc)
        w = np.bincount(labels, minlength=C.shape[0]).astype(np.float64)
        if (w == 0).any():
            raise NotImplementedError("empty cluster: use lloyd_fixed")
        C = Cn * (1.0 / w)[:, None]
        if labels_old is not None and np.array_equal(labels, labels_old):
            strict = True
            break
        labels_old = labels
    if not strict:
        d = (C * C).sum(axis=1)[None, :] - 2.0 * Xc @ C.T
        labels = d.argmin(axis=1).astype(np.int32)
    inertia = float(((Xc - C[labels]) ** 2).sum())
    return labels, C + mu, inertia, it
