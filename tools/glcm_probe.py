"""A/B probe of the dense GLCM kernel variants on one B200: times `glcm_props` (+ `glcm_resize`) per variant on the config-B scene
and on a uniform-noise scene (every window spans all levels: worst case for the folded counters), and checks that every
variant writes the same five property planes as the baseline kernel, bit for bit.

    python tools/glcm_probe.py [size] [variant ...]      variant = name:opt=val,opt=val
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rs_image_segmentation_b200 import _lib, pipeline as P
from rs_image_segmentation_b200.device import StageTimer
from rs_image_segmentation_b200.synth import synth_strip_torch

size = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
variants = sys.argv[2:] or ["base:glcm_fold=0", "auto64:glcm_fold=1,glcm_fold_nt=64,glcm_fold_force=-1", "auto32:glcm_fold=1,glcm_fold_nt=32,glcm_fold_force=-1",
                            "force8:glcm_fold=1,glcm_fold_nt=64,glcm_fold_force=8", "force16:glcm_fold=1,glcm_fold_nt=64,glcm_fold_force=16"]


def run(raster, cfg, reps=3):
    best = None
    fr = None
    for _ in range(reps):
        fr = None
        t = StageTimer(True, only={"glcm_props", "glcm_resize"})
        fr = P.extract_features(raster, cfg, timer=t)
        torch.cuda.synchronize()
        st = {k: v[0] for k, v in t.totals_ms().items()}
        if best is None or st["glcm_props"] < best["glcm_props"]:
            best = st
    return fr, best


for scene in ("synthetic", "noise"):
    H = W = size if scene == "synthetic" else min(size, 3000)
    if scene == "synthetic":
        raster = synth_strip_torch(H, W, 7, 0, H, "uint8", seed=7000, device="cuda")
    else:
        g = torch.Generator(device="cuda")
        g.manual_seed(5)
        raster = torch.randint(0, 256, (H, W, 7), dtype=torch.uint8, device="cuda", generator=g)
    for win, levels in ((7, 32), (5, 16), (11, 32)):
        cfg = P.FeatureConfig(glcm_window=win, glcm_step=1, glcm_levels=levels)
        ref = None
        for v in variants:
            name, opts = v.split(":")
            for o in opts.split(","):
                k, val = o.split("=")
                _lib.set_option(k, int(val))
            fr, st = run(raster, cfg)
            planes = fr.planes[7:12, :fr.n_px]
            same = None
            if ref is None:
                ref = planes.clone()
            else:
                same = [bool(torch.equal(planes[i], ref[i])) for i in range(5)]
                if not all(same):
                    d = (planes - ref).abs().amax(dim=1).tolist()
                    same = {"equal": same, "max_abs_diff": d, "n_diff_energy": int((planes[3] != ref[3]).sum())}
            print(json.dumps({"scene": scene, "size": H, "window": win, "levels": levels, "variant": name,
                              "glcm_props_ms": round(st["glcm_props"], 3), "glcm_resize_ms": round(st["glcm_resize"], 3), "same_as_base": same}), flush=True)
            del fr, planes
        del ref
    del raster
