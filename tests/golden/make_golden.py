#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference modules import matplotlib / rasterio / osgeo / skimage at module
scope (modules/features/indices.py:18,22, extract.py:15,23-24); none of those is
installed here and none is used by the functions we call, so empty stub modules are
registered first.  What is recorded:

  aa_crop.npz      a 64x96 crop of the stage-1 output of data/raw/AA.tif (uint8, 7 bands)
                   and the reference's outputs on it: robust_normalize, the seven indices,
                   perform_pca, prepare_level_1_features, add_spatial_context,
                   unsupervised_kmeans_classification (indices as features, k=5; float32 and float64 inputs; k=7 on a 3-D stack)
  aa_full_stats.npz  P2/P98, index means, PCA evr/components of the full 600x600 scene
                   (SURVEY.md 8(c) table), for the oracle only

calculate_glcm_features cannot be run (needs scikit-image): GLCM parity is unpinned by
the reference; its fixtures come from the scikit-image docstring example instead.
"""
import os
import struct
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    mpl = _stub("matplotlib", use=lambda *a, **k: None, rcParams={}, font_manager=_Any())
    for sub in ("pyplot", "patches", "colors", "font_manager", "cm", "gridspec"):
        setattr(mpl, sub, _stub("matplotlib." + sub, __getattr__=lambda k: _Any()))
    ras = _stub("rasterio", open=None)
    ras.transform = _stub("rasterio.transform", Affine=_Any(), from_origin=_Any())
    ras.features = _stub("rasterio.features", __getattr__=lambda k: _Any())
    ras.crs = _stub("rasterio.crs", CRS=_Any())
    og = _stub("osgeo", gdal=_Any(), osr=_Any())
    _stub("osgeo.gdal", __getattr__=lambda k: _Any())
    sk = _stub("skimage")
    sk.feature = _stub("skimage.feature", graycomatrix=None, graycoprops=None, local_binary_pattern=None)
    for sub in ("filters", "morphology", "measure", "segmentation", "util", "exposure", "color"):
        setattr(sk, sub, _stub("skimage." + sub, __getattr__=lambda k: _Any()))
    _stub("skimage.filters.rank", __getattr__=lambda k: _Any())
    _stub("seaborn", __getattr__=lambda k: _Any())
    sys.path.insert(0, REF)
    import modules.features.extract as ext
    import modules.features.indices as ind
    import modules.features.preprocessing as pre
    return ind, ext, pre


def read_tiff_planar_u8(path):
    """Minimal reader for the bundled scene: uncompressed, 8-bit, PlanarConfiguration=2."""
    b = open(path, "rb").read()
    fmt = "<" if b[:2] == b"II" else ">"
    off = struct.unpack(fmt + "I", b[4:8])[0]
    n = struct.unpack(fmt + "H", b[off:off + 2])[0]
    tags = {}
    size = {1: 1, 3: 2, 4: 4, 5: 8}
    code = {1: "B", 3: "H", 4: "I"}
    for i in range(n):
        tag, typ, cnt, val = struct.unpack(fmt + "HHII", b[off + 2 + 12 * i: off + 14 + 12 * i])
        if typ in code:
            if cnt * size[typ] <= 4:
                raw = b[off + 2 + 12 * i + 8: off + 2 + 12 * i + 12]
            else:
                raw = b[val: val + cnt * size[typ]]
            tags[tag] = struct.unpack(fmt + code[typ] * cnt, raw[:cnt * size[typ]])
    W, H, spp = tags[256][0], tags[257][0], tags[277][0]
    assert tags[259][0] == 1 and tags[284][0] == 2 and set(tags[258]) == {8}
    rps = tags[278][0]
    offs, cnts = tags[273], tags[279]
    strips_per_band = (H + rps - 1) // rps
    out = np.zeros((spp, H, W), np.uint8)
    for s in range(spp):
        rows = []
        for k in range(strips_per_band):
            o, c = offs[s * strips_per_band + k], cnts[s * strips_per_band + k]
            rows.append(np.frombuffer(b, np.uint8, c, o))
        out[s] = np.concatenate(rows).reshape(H, W)
    return out


def main():
    ind, ext, pre = import_reference()
    raw = read_tiff_planar_u8(os.path.join(REF, "data/raw/AA.tif"))
    # stage 1 exactly as scripts/1_preprocessing.py drives it (the identity warp is a no-op)
    cal = pre.radiometric_calibration([raw[i] for i in range(raw.shape[0])])
    geo = pre.geometric_correction(cal, None)
    for a, c in zip(cal, geo):
        assert np.array_equal(a, c), "identity warpAffine changed the data"
    enh = pre.image_enhancement(geo)                       # list of uint8
    stage1 = np.stack(enh)                                 # (7, 600, 600) uint8

    def run(stage1_u8):
        bands = [b.astype(np.float32) for b in stage1_u8]  # scripts/2_feature_extraction.py:158
        nb = [ind.robust_normalize(b) for b in bands]
        pct = np.array([[np.percentile(b, 2), np.percentile(b, 98)] for b in bands], np.float32)
        blue, green, red, nir, swir1 = nb[0], nb[1], nb[2], nb[3], nb[4]
        ix = {
            "ndvi": ind.calculate_ndvi(nir, red),
            "evi": ind.calculate_evi(nir, red, blue),
            "msavi": ind.calculate_msavi(nir, red),
            "ndwi": ind.calculate_ndwi(green, nir),
            "mndwi": ind.calculate_mndwi(green, swir1),
            "ndbi": ind.calculate_ndbi(swir1, nir),
            "bsi": ind.calculate_bsi(blue, red, nir, swir1),
        }
        pcs, evr, model = ind.perform_pca(nb, use_robust_scaling=True)
        fd = dict(ix)
        fd["pca_result"] = pcs
        l1 = ind.prepare_level_1_features(fd)
        l1c = ind.add_spatial_context(l1)
        nir2 = ind.robust_normalize(nir)                   # indices.py:265, inside the GLCM call
        q32 = (nir2 * 31).astype(np.uint8)                 # indices.py:268
        return nb, pct, ix, pcs, evr, model, l1, l1c, q32

    # ---- crop fixture
    r0, c0, h, w = 200, 300, 64, 96
    crop = np.ascontiguousarray(stage1[:, r0:r0 + h, c0:c0 + w])
    nb, pct, ix, pcs, evr, model, l1, l1c, q32 = run(crop)
    kdict = {k: v for k, v in ix.items()}
    kdict.update(height=h, width=w)
    labels = ext.unsupervised_kmeans_classification(kdict, n_clusters=5, feature_keys_to_use=list(ix.keys()))
    # the same call on the float64 promotion of the same maps (the reference's real stack is float64):
    # this is the variant whose result does not depend on float32 BLAS summation order
    kdict64 = {k: v.astype(np.float64) for k, v in ix.items()}
    kdict64.update(height=h, width=w)
    labels64 = ext.unsupervised_kmeans_classification(kdict64, n_clusters=5, feature_keys_to_use=list(ix.keys()))
    stack3d = np.stack([ix[k] for k in ix], axis=-1).astype(np.float64)
    labels64_k7 = ext.unsupervised_kmeans_classification({"hierarchical_all": stack3d, "height": h, "width": w}, n_clusters=7,
                                                         feature_keys_to_use=["hierarchical_all"])
    np.savez_compressed(
        os.path.join(HERE, "aa_crop.npz"),
        stage1_u8=crop, pct=pct, norm=np.stack(nb),
        **{"ix_" + k: v for k, v in ix.items()},
        pca_maps=np.stack(pcs), pca_evr=evr, pca_components=model.components_, pca_mean=model.mean_,
        level1=l1, level1_ctx=l1c, q32=q32, kmeans_labels_k5=labels.astype(np.int32),
        kmeans_labels_k5_f64=labels64.astype(np.int32), kmeans_labels_k7_f64_3d=labels64_k7.astype(np.int32),
    )
    # ---- full-scene statistics
    nb, pct, ix, pcs, evr, model, l1, l1c, q32 = run(stage1)
    np.savez_compressed(
        os.path.join(HERE, "aa_full_stats.npz"),
        pct=pct, index_means=np.array([ix[k].astype(np.float64).mean() for k in
                                       ("ndvi", "evi", "msavi", "ndwi", "mndwi", "ndbi", "bsi")]),
        pca_evr=evr, pca_components=model.components_, q32_max=np.array(q32.max()),
        stage1_sha=np.frombuffer(__import__("hashlib").sha256(stage1.tobytes()).digest(), np.uint8),
    )
    np.save(os.path.join("/tmp", "aa_stage1_full.npy"), stage1)   # scratch, for the oracle check only
    print("pct", pct.tolist())
    print("evr", evr)
    print("labels", np.bincount(labels.ravel()))


if __name__ == "__main__":
    main()
