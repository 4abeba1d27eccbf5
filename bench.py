#!/usr/bin/env python
"""Benchmark of the hot path: feature stack (indices + PCA + dense GLCM) + KMeans, Mpixel/s.

  python bench.py --gpus N --steps K --warmup W            rsx (this repo), one process per GPU under torchrun
  python bench.py --impl reference ...                     the reference CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): synthetic Landsat TM scene 7000x7000x7 uint8 per GPU (row strips of a
7000*N x 7000 mosaic: weak scaling), GLCM 7x7 dense at 32 grey levels, stack-13, KMeans k=8, 20 Lloyd iterations
+ sklearn's final assignment pass.  One step = the whole path over the whole raster.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per pixel
def algorithmic_bytes(B, e, D, T, n_comp, glcm=True):
    hist = B * e
    indices = B * e + 7 * 4
    pca = 2 * B * e + 4 * n_comp
    g = 21 if glcm else 0
    km = T * 4 * D + 4
    return dict(hist=hist, indices=indices, pca=pca, glcm=g, kmeans=km, total=hist + indices + pca + g + km)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  NVML in a thread, two cheap queries every 40 ms (100 ms and 40 ms measured alike, 20 ms not): every
    NVML / nvidia-smi query takes driver locks that stall kernel launches (a polling `nvidia-smi -lms 100` child stretched
    a 20 ms step to 37 ms, NVML every 20 ms a 2-GPU step to 40 ms), so the sampling is kept sparse; nvidia-smi once as a
    fallback."""

    def __init__(self, gpu_index: int, period_s: float = 0.1):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.mx, self.reasons = [], [], set()
        self.thread, self.stop_flag, self.nvml = None, threading.Event(), None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap indices: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.idx), "pci_bus_id") else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.handle = h
                        break
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        if not self.mx:
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                          ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def _loop(self):
        self.stop_flag.wait(0.01)                               # first sample once the first timed step is in flight
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            if not self.sm:
                try:
                    self._sample()
                except Exception:
                    pass
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "how": "NVML, sampled during the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["unknown (NVML unavailable)"], "samples": 1,
                    "how": "nvidia-smi once after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------- reference (CPU) arm
def cpu_reference_sample(raster_crop: np.ndarray, glcm_crop: np.ndarray, cfg, K, T, seed):
    """The reference path (oracle port: numpy / sklearn as the reference calls them + the plain-C GLCM restatement)
    on a bounded sample.  Returns per-pixel seconds per stage and the threads used."""
    from oracle import features as of
    from oracle import glcm as og
    from oracle import kmeans as ok
    h, w, B = raster_crop.shape
    n = h * w
    t = {}
    t0 = time.perf_counter()
    bands = [raster_crop[:, :, b].astype(np.float32) for b in range(B)]
    nb = [of.robust_normalize(b) for b in bands]
    ix = of.all_indices(nb)
    t["normalize+indices"] = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    pcs, _, _ = of.perform_pca(nb)
    t["pca"] = (time.perf_counter() - t0) / n
    # dense GLCM on a smaller crop (the reference's Python window loop would take ~0.85 ms per window)
    gh, gw = glcm_crop.shape
    t0 = time.perf_counter()
    gmaps = og.glcm_features(glcm_crop, cfg.glcm_levels, cfg.glcm_window, cfg.glcm_step)
    t["glcm"] = (time.perf_counter() - t0) / (gh * gw)
    # KMeans on a 13-plane stack of the crop (GLCM planes replaced by resized copies so the stack has full depth)
    import cv2
    stack = [ix[k] for k in of.INDEX_ORDER] + [cv2.resize(gmaps[k], (w, h)) for k in og.PROPS] + [pcs[0]]
    X = np.stack([s.ravel() for s in stack], axis=1).astype(np.float64)
    t0 = time.perf_counter()
    Xs = ok.minmax_scale(X)
    rng = np.random.default_rng(seed)
    c0 = Xs[rng.choice(n, K, replace=False)]
    ok.lloyd_fixed(Xs, c0, T)
    t["kmeans"] = (time.perf_counter() - t0) / n
    return t


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from rs_image_segmentation_b200.pipeline import FeatureConfig
    from rs_image_segmentation_b200.synth import synth_raster_numpy
    cfg = FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
    S, G = args.cpu_sample, args.cpu_glcm_sample
    raster = synth_raster_numpy(S, S, 7, np.uint8, seed=7000)
    from oracle import features as of
    nir = of.robust_normalize(raster[:G, :G, 3].astype(np.float32))
    cores = os.cpu_count()
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu_reference_sample(raster, nir, cfg, args.k, args.iters, 7000)
        if i >= args.warmup:
            times.append(sum(t.values()))
    per_px = float(np.mean(times))
    value = 1e-6 / per_px
    H = W = args.size
    sample = (f"indices+PCA+KMeans(k={args.k},{args.iters} it) on a {S}x{S}x7 crop, dense 7x7/32-level GLCM on a {G}x{G} crop; "
              f"per-pixel stage times summed; numpy/sklearn default threading + OpenMP C GLCM")
    line = {
        "impl": "reference", "metric": "feature-stack+KMeans throughput", "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_px * H * W * args.gpus * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, H, W),
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


def workload_config(args, H, W):
    return {"workload": f"synthetic Landsat TM scene {H}x{W}x7 uint8 per GPU (row strips of a {H * args.gpus}x{W} mosaic), "
                        f"indices + PCA(7) + dense GLCM 7x7 @32 levels + KMeans k={args.k}, {args.iters} iterations + final assignment, stack-13",
            "l2": "inputs larger than L2 (raster 343 MB, stack 2.5 GB per GPU)", "per_gpu_pixels": H * W}


# ------------------------------------------------------------------------------------------- rsx arm
def run_rsx(args):
    import torch
    import torch.distributed as dist

    from rs_image_segmentation_b200 import _lib
    from rs_image_segmentation_b200 import pipeline as P
    from rs_image_segmentation_b200.device import StageTimer
    from rs_image_segmentation_b200.dist import Comm, strip_bounds
    from rs_image_segmentation_b200.synth import synth_strip_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    H = W = args.size
    H_total = H * world
    bounds = strip_bounds(H_total, world)
    own = bounds[rank]
    cfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)
    D, K, T = 13, args.k, args.iters

    raster = synth_strip_torch(H_total, W, 7, own[0], own[1] - own[0], "uint8", seed=7000, device="cuda")
    torch.cuda.synchronize()
    # Inside the timed region only the dominant kernel is bracketed by events (the roofline needs its launch times measured
    # live); ~100 more event records per step for the other stages cost 0.33 ms of a 19 ms step, so the stage table comes
    # from one extra instrumented step after the timed region.
    timer = StageTimer(enabled=os.environ.get("RSX_BENCH_NOTIMER", "0") != "1", only={"kmeans_assign_delta"})
    timer_all = StageTimer(enabled=os.environ.get("RSX_BENCH_NOTIMER", "0") != "1")

    def step(t):
        fr = P.extract_features(raster, cfg, comm, H_total, bounds, t)
        res, km, c0 = P.kmeans_on_features(fr, D, K, T, 7000, comm, H_total, own[0], True, t)
        return fr, res

    def timed(fn, steps):
        comm.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = None
        for _ in range(steps):
            out = None          # release the previous step's buffers first: the caching allocator then reuses the same blocks
            out = fn()          # (a second live generation would cost a multi-GB cudaMalloc inside the timed region)
        b.record()
        torch.cuda.synchronize()
        comm.barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        comm.all_reduce(ms, "max")
        return float(ms.item()), out

    keep = None
    for _ in range(args.warmup):
        keep = None
        keep = step(StageTimer(enabled=False))
    keep = None
    sampler = ClockSampler(local, float(os.environ.get("RSX_BENCH_CLOCK_PERIOD", "0.04")))
    if rank == 0 and os.environ.get("RSX_BENCH_NOCLOCKS", "0") != "1":
        sampler.start()
    launches0 = _lib.launch_count()
    timer.reset()
    ms_total, (fr, res) = timed(lambda: step(timer), args.steps)
    launches = _lib.launch_count() - launches0
    stage = timer.totals_ms()
    del fr, res
    keep = step(timer_all)                                  # the instrumented step (not part of `value`)
    torch.cuda.synchronize()
    stage_all = timer_all.totals_ms()
    fr, res = keep
    keep = None
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    n_global = H_total * W
    value = n_global / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer API: H2D of the strip + D2H of the labels inside the timed region
    e2e = None
    if not args.no_e2e:
        del fr, res
        pinned = raster.cpu().pin_memory()
        torch.cuda.synchronize()

        # the public host-buffer API, pipelined over the steps: every step's H2D (pinned raster) and D2H (int32 labels) are
        # inside the timed region; the copy engines work under the kernels of the neighbouring steps
        yields = []

        def e2e_run(steps):
            last = None
            yields.clear()
            for labels, res in P.segment_stream((pinned for _ in range(steps)), cfg, K, T, 7000, D, comm, H_total, bounds):
                last = int(labels[0, 0]) + res.n_iter          # touch the result on the host
                yields.append(time.perf_counter())
            return last

        e2e_run(2)
        e2e_steps = max(2, args.steps)
        ms_e2e, _ = timed(lambda: e2e_run(e2e_steps), 1)
        ms_e2e /= e2e_steps
        e2e = {"value": n_global / (ms_e2e * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": int(pinned.numel()) * world,
               "d2h_bytes_per_step": int(H * W * 4) * world, "ms_per_step": ms_e2e, "steps": e2e_steps,
               "api": "pipeline.segment_stream (double-buffered H2D / D2H on the copy engines); one scene alone through "
                      "pipeline.segment_raster: see single_scene_ms"}
        if len(yields) >= 4:
            # host-clock interval between consecutive label arrays in the middle of the run: what a long stream of scenes
            # costs per scene; ms_per_step above also carries the first upload and the last download of this short run
            gaps = np.diff(np.array(yields[:-1])) * 1e3
            e2e["steady_state_ms_per_step"] = float(np.median(gaps[1:])) if len(gaps) > 1 else float(gaps[0])
        for _ in range(2):
            t0 = time.perf_counter()
            P.segment_raster(None, cfg, K, T, 7000, D, comm, H_total, bounds, pinned=pinned)
            torch.cuda.synchronize()
            e2e["single_scene_ms"] = (time.perf_counter() - t0) * 1e3

    if rank != 0:
        return 0
    peak, peak_src = peaks()
    ab = algorithmic_bytes(7, 1, D, T, 7)
    n_local = H * W
    km_ms, km_n = stage.get("kmeans_assign_delta", (0.0, 0))
    km_avg_ms = km_ms / max(km_n, 1)
    km_bytes = n_local * 4 * D
    achieved = km_bytes / (km_avg_ms * 1e-3) / 1e9 if km_avg_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "kmeans_assign_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    whole = ab["total"] * n_local / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": "feature-stack+KMeans throughput", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, H, W),
        "roofline": {"bound": "hbm", "kernel": "km_stream_kernel<13,DELTA,K<=8> (19 of the 21 KMeans passes: TMA-staged assign + exact delta update)",
                     "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": km_bytes, "algorithmic_bytes_per_pixel": 4 * D,
                     "actual_bytes_per_pixel": 4 * D + 2, "avg_launch_ms": km_avg_ms, "launches_timed": km_n,
                     "share_of_step": km_ms / args.steps / ms_per_step if ms_per_step else None},
        "whole_path": {"algorithmic_bytes_per_pixel": ab["total"], "achieved_gbs_per_gpu": whole, "frac_of_peak": whole / peak,
                       "stage_ms_per_step": {k: v[0] for k, v in stage_all.items()},
                       "stage_launches_per_step": {k: v[1] for k, v in stage_all.items()},
                       "stage_note": "one extra step with every stage bracketed by CUDA events, after the timed region"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
    }
    if world == 1 and not args.no_cpu:
        from oracle import features as of
        S, G = args.cpu_sample, args.cpu_glcm_sample
        crop = raster[:S, :S].cpu().numpy()
        nir = of.robust_normalize(crop[:G, :G, 3].astype(np.float32))
        t = cpu_reference_sample(crop, nir, cfg, K, T, 7000)
        per_px = sum(t.values())
        line["cpu_baseline"] = {
            "value": 1e-6 / per_px, "unit": "Mpixel/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"indices+PCA+KMeans on a {S}x{S} crop, dense GLCM (C/OpenMP restatement) on a {G}x{G} crop, per-pixel times summed",
            "stage_us_per_pixel": {k: v * 1e6 for k, v in t.items()}}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _emit(line: dict):
    """The one JSON line on the real stdout (everything else - NCCL banners, warnings - was redirected to stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rsx", choices=["rsx", "reference"])
    ap.add_argument("--size", type=int, default=7000, help="rows = cols of the per-GPU scene")
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--cpu-glcm-sample", type=int, default=768)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_rsx(args)


if __name__ == "__main__":
    sys.exit(main())
