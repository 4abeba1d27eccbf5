#!/usr/bin/env python
"""Benchmark of the hot path: feature stack (indices + PCA + dense GLCM) + KMeans, Mpixel/s.

  python bench.py --gpus N --steps K --warmup W            rsx (this repo), one process per GPU under torchrun
  python bench.py --impl reference ...                     the reference CPU path (oracle port) on the host cores

Headline workload (BASELINE.json configs[1], "B"): synthetic Landsat TM scene 7000x7000x7 uint8 per GPU (row strips of a
7000*N x 7000 mosaic: weak scaling), GLCM 7x7 dense at 32 grey levels, stack-13, KMeans k=8, 20 Lloyd iterations
+ sklearn's final assignment pass.  One step = the whole path over the whole raster.

The other BASELINE.json configurations are measured in the same run and reported under "configs" (never as `value`):
  C  GLCM sweep on the 7000x7000 scene: windows 5/7/11 x grey levels 16/32/64, dense (N = 1 only)
  D  Sentinel-2-like tile 10980x10980x13 uint16, indices + PCA(6) + KMeans k=16, 20 iterations (row strips at N > 1: strong)
  E  40000x40000x7 uint8 mosaic, GLCM 7x7 @32 + KMeans k=32, 20 iterations: 40000/N rows per GPU with the real halo
     exchange at N >= 2; at N = 1 one 5000-row strip (the share of one of 8 GPUs)
With several ranks the run also checks that the sharded path reproduces the single-GPU result bit for bit ("mgpu_parity").
"""
from __future__ import annotations

import os
import sys

# torchrun exports OMP_NUM_THREADS=1 to its workers unless the caller set it; the reference arm is a CPU measurement that must
# use the host cores (numpy/OpenBLAS and libgomp read the variable when they are loaded, i.e. before anything else here)
if "--impl" in sys.argv and "reference" in sys.argv and os.environ.get("RSX_KEEP_OMP") != "1":
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.pop(_v, None)

import argparse
import json
import subprocess
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
def use_all_host_threads():
    """numpy's BLAS, sklearn's OpenMP loops and the oracle's OpenMP C GLCM on every host core, whatever the launcher exported
    (threadpoolctl reaches the runtimes that are already loaded; the environment was cleaned at the top of this file)."""
    cores = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def cpu_reference_sample(raster_crop: np.ndarray, glcm_crop: np.ndarray, cfg, K, T, seed):
    """The reference path (oracle port: numpy / sklearn as the reference calls them + the plain-C GLCM restatement)
    on a bounded sample.  Returns per-pixel seconds per stage."""
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


def cpu_sample_text(S, G, K, T):
    return (f"oracle port (the reference's numpy / sklearn / cv2 calls + OpenMP C restatement of skimage's GLCM loop; /root/reference "
            f"is absent on the GPU box): normalise + 7 indices + PCA + MinMax/KMeans(k={K}, {T} it, sklearn float64) on a {S}x{S}x7 crop of "
            f"the config-B scene, dense 7x7/32-level GLCM on a {G}x{G} crop; per-pixel stage times summed; one untimed warm-up sample first")


def cpu_baseline_measure(crop, nir, cfg, K, T, seed, warmups, steps):
    """Same protocol in both arms: `warmups` untimed samples (thread pools, page faults, imports), then the mean of `steps`."""
    cores = use_all_host_threads()
    for _ in range(warmups):
        cpu_reference_sample(crop, nir, cfg, K, T, seed)
    runs = [cpu_reference_sample(crop, nir, cfg, K, T, seed) for _ in range(steps)]
    stage = {k: float(np.mean([r[k] for r in runs])) for k in runs[0]}
    return stage, sum(stage.values()), cores


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
    stage, per_px, cores = cpu_baseline_measure(raster, nir, cfg, args.k, args.iters, 7000, max(1, min(args.warmup, 2)), max(1, args.steps))
    value = 1e-6 / per_px
    H = W = args.size
    sample = cpu_sample_text(S, G, args.k, args.iters)
    line = {
        "impl": "reference", "metric": "feature-stack+KMeans throughput", "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_px * H * W * args.gpus * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, H, W),
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample,
                         "stage_us_per_pixel": {k: v * 1e6 for k, v in stage.items()},
                         "omp_env": os.environ.get("OMP_NUM_THREADS", "unset")},
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
    peak, peak_src = peaks()
    timers_on = os.environ.get("RSX_BENCH_NOTIMER", "0") != "1"

    raster = synth_strip_torch(H_total, W, 7, own[0], own[1] - own[0], "uint8", seed=7000, device="cuda")
    torch.cuda.synchronize()
    # Inside the timed region only the dominant kernel is bracketed by events (the roofline needs its launch times measured
    # live); ~100 more event records per step for the other stages cost 0.33 ms of a 19 ms step, so the stage table comes
    # from one extra instrumented step after the timed region.
    timer = StageTimer(enabled=timers_on, only={"kmeans_assign_delta"})
    timer_all = StageTimer(enabled=timers_on)

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

        # the public host-buffer API, pipelined over the steps: every step's H2D (pinned raster) and D2H (labels) are
        # inside the timed region; the copy engines work under the kernels of the neighbouring steps
        yields = []
        # uint8 labels (K <= 256) are what the reference's writer stores (labels + 1 as uint8) and a quarter of the
        # bytes; at 8 GPUs the int32 image makes the run host-memory bound (8 x 539 MB per 19 ms), so it is reported
        # beside the headline as "int32_labels"
        label_mode = os.environ.get("RSX_BENCH_LABELS", "uint8")

        def e2e_run(steps, mode):
            last = None
            yields.clear()
            for labels, res in P.segment_stream((pinned for _ in range(steps)), cfg, K, T, 7000, D, comm, H_total, bounds, labels=mode):
                last = int(labels[0, 0]) + res.n_iter          # touch the result on the host
                yields.append(time.perf_counter())
            return last

        e2e_steps = max(2, args.steps)
        alt_mode = "int32" if label_mode != "int32" else "uint8"
        e2e_run(3, alt_mode)                                   # three scenes: every label slot of the stream has been used once
        ms_alt, _ = timed(lambda: e2e_run(e2e_steps, alt_mode), 1)
        ms_alt /= e2e_steps
        e2e_run(3, label_mode)
        ms_e2e, _ = timed(lambda: e2e_run(e2e_steps, label_mode), 1)
        ms_e2e /= e2e_steps
        d2h = int(P.segment_stream_d2h_bytes(H * W, label_mode))
        e2e = {"value": n_global / (ms_e2e * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": int(pinned.numel()) * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e, "steps": e2e_steps, "labels": label_mode,
               alt_mode + "_labels": {"value": n_global / (ms_alt * 1e-3) / 1e6, "ms_per_step": ms_alt,
                                      "d2h_bytes_per_step": int(P.segment_stream_d2h_bytes(H * W, alt_mode)) * world},
               "api": f"pipeline.segment_stream(labels='{label_mode}') (double-buffered H2D / D2H on the copy engines); one scene alone "
                      "through pipeline.segment_raster (int32 labels): see single_scene_ms"}
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
        del pinned
    fr = res = None
    del raster
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations (reported under "configs", never as `value`)
    which = [c for c in args.configs.split(",") if c]
    configs = {}

    def stage_table(t):
        return {k: round(v[0], 3) for k, v in t.totals_ms().items()}

    def measure(make_step, n_steps, n_warm=1):
        """-> (ms per step, max over ranks; stage table of one extra instrumented step on this rank)"""
        for _ in range(n_warm):
            make_step(StageTimer(False))
        ms, _ = timed(lambda: make_step(StageTimer(False)), n_steps)
        t = StageTimer(timers_on)
        make_step(t)
        torch.cuda.synchronize()
        return ms / n_steps, t

    def kernel_roofline(t, name, n_px, Dk):
        ms, n = t.totals_ms().get(name, (0.0, 0))
        if not n or ms <= 0:
            return None
        avg = ms / n
        ach = n_px * 4 * Dk / (avg * 1e-3) / 1e9
        return {"kernel": name, "avg_launch_ms": round(avg, 4), "launches": n, "achieved_gbs": round(ach, 1), "frac": round(ach / peak, 4)}

    if "C" in which and world == 1:
        Hc = Wc = args.size
        raster_c = synth_strip_torch(Hc, Wc, 7, 0, Hc, "uint8", seed=7000, device="cuda")
        rows = []
        for levels in (16, 32, 64):
            for win in (5, 7, 11):
                ccfg = P.FeatureConfig(glcm_window=win, glcm_step=1, glcm_levels=levels)
                best = None
                for rep in range(3):                          # rep 0 = warm-up
                    t = StageTimer(timers_on, only={"glcm_props", "glcm_resize"})
                    frc = P.extract_features(raster_c, ccfg, timer=t)
                    torch.cuda.synchronize()
                    st = stage_table(t)
                    del frc
                    g = st.get("glcm_props", 0.0) + st.get("glcm_resize", 0.0)
                    if rep and (best is None or g < best[0]):
                        best = (g, st)
                n_win = (Hc - win + 1) * (Wc - win + 1)
                g_ms, st = best
                rows.append({"window": win, "levels": levels, "glcm_props_ms": st.get("glcm_props"), "glcm_resize_ms": st.get("glcm_resize"),
                             "Mwindows_per_s": round(n_win / max(g_ms, 1e-9) / 1e3, 1),
                             "frac_of_hbm_peak_21B_per_px": round(Hc * Wc * 21 / max(g_ms, 1e-9) / 1e6 / peak, 4)})
        configs["C"] = {"workload": f"dense GLCM sweep on the {Hc}x{Wc} scene, 4 offsets, windows 5/7/11 x grey levels 16/32/64 (best of 2 after a warm-up)",
                        "note": "GLCM is instruction-issue / shared-memory bound, not HBM bound (SURVEY.md 7 hard part 1); the fraction is against the 21 B/px byte model",
                        "rows": rows}
        del raster_c
        torch.cuda.empty_cache()

    if "D" in which:
        Hd = Wd = args.d_size
        bd = strip_bounds(Hd, world)
        od = bd[rank]
        raster_d = synth_strip_torch(Hd, Wd, 13, od[0], od[1] - od[0], "uint16", seed=10980, device="cuda")
        dcfg = P.FeatureConfig(band_map=(1, 2, 3, 7, 11), n_components=6, glcm=False)

        def step_d(t):
            f = P.extract_features(raster_d, dcfg, comm, Hd, bd, t)
            return P.kmeans_on_features(f, 13, 16, T, 10980, comm, Hd, od[0], True, t)

        ms_d, t_d = measure(step_d, 2)
        n_d = Hd * Wd
        ab_d = algorithmic_bytes(13, 2, 13, T, 6, glcm=False)["total"]
        configs["D"] = {"workload": f"synthetic Sentinel-2-like tile {Hd}x{Wd}x13 uint16, indices + PCA(6) + KMeans k=16, {T} iterations + final assignment, "
                                    f"stack-13 (7 indices + 6 PCs); {'one GPU' if world == 1 else f'row strips over {world} GPUs (strong scaling)'}",
                        "ms_per_step": round(ms_d, 3), "Mpixel_per_s": round(n_d / ms_d / 1e3, 1), "n_gpus": world, "scaling": "strong",
                        "algorithmic_bytes_per_pixel": ab_d,
                        "whole_path_frac_of_peak_per_gpu": round(ab_d * (n_d / world) / (ms_d * 1e-3) / 1e9 / peak, 4),
                        "dominant_kernel": kernel_roofline(t_d, "kmeans_assign_delta", (od[1] - od[0]) * Wd, 13),
                        "stage_ms_rank0": stage_table(t_d)}
        del raster_d
        torch.cuda.empty_cache()

    if "E" in which:
        He_total, We = args.e_size, args.e_size
        if world == 1:
            rows_e = max(1, He_total // 8)
            be, oe, He_run = [(0, rows_e)], (0, rows_e), rows_e         # one GPU's share at 8 GPUs, as its own image
            what = f"ONE {rows_e}x{We} strip (the share of one of 8 GPUs; no neighbour, so no halo exchange)"
        else:
            be = strip_bounds(He_total, world)
            oe, He_run = be[rank], He_total
            what = f"{He_total}x{We} mosaic in row strips over {world} GPUs with the GLCM halo exchange and peer-memory centroid reduction (strong scaling)"
        raster_e = synth_strip_torch(He_total, We, 7, oe[0], oe[1] - oe[0], "uint8", seed=40000, device="cuda")
        ecfg = P.FeatureConfig(glcm_window=7, glcm_step=1, glcm_levels=32)

        def step_e(t):
            f = P.extract_features(raster_e, ecfg, comm, He_run, be, t)
            return P.kmeans_on_features(f, 13, 32, T, 40000, comm, He_run, oe[0], True, t)

        ms_e, t_e = measure(step_e, 2)
        n_e = He_run * We
        ab_e = algorithmic_bytes(7, 1, 13, T, 7)["total"]
        configs["E"] = {"workload": f"synthetic 40k mosaic, uint8 x 7 bands, indices + PCA(7) + dense GLCM 7x7 @32 + KMeans k=32, {T} iterations + final "
                                    f"assignment, stack-13: {what}",
                        "ms_per_step": round(ms_e, 3), "Mpixel_per_s": round(n_e / ms_e / 1e3, 1), "n_gpus": world,
                        "scaling": "strong" if world > 1 else "one strip", "pixels": n_e, "algorithmic_bytes_per_pixel": ab_e,
                        "whole_path_frac_of_peak_per_gpu": round(ab_e * (n_e / world) / (ms_e * 1e-3) / 1e9 / peak, 4),
                        "dominant_kernel": kernel_roofline(t_e, "kmeans_assign_delta", (oe[1] - oe[0]) * We, 13),
                        "stage_ms_rank0": stage_table(t_e)}
        del raster_e
        torch.cuda.empty_cache()

    # ---- several ranks: the sharded path must equal the single-GPU path bit for bit (small cases, rank 0 judges)
    mgpu_parity = None
    if world > 1 and not args.no_parity:
        from rs_image_segmentation_b200.selfcheck import sharded_equals_single
        try:
            failures = sharded_equals_single(comm)
            mgpu_parity = "ok" if not failures else "; ".join(failures)[:500]
        except Exception as ex:                                # reported, never hidden
            mgpu_parity = f"error: {type(ex).__name__}: {ex}"[:500]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
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
        "roofline": {"bound": "hbm", "kernel": "km_stream_kernel<13,DELTA,K<=8> (the delta passes of the 21 KMeans passes: TMA-staged assign + exact delta update)",
                     "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": km_bytes, "algorithmic_bytes_per_pixel": 4 * D,
                     "actual_bytes_per_pixel": 4 * D + 2, "avg_launch_ms": km_avg_ms, "launches_timed": km_n,
                     "share_of_step": km_ms / args.steps / ms_per_step if ms_per_step else None},
        "whole_path": {"algorithmic_bytes_per_pixel": ab["total"], "achieved_gbs_per_gpu": whole, "frac_of_peak": whole / peak,
                       "stage_ms_per_step": {k: v[0] for k, v in stage_all.items()},
                       "stage_launches_per_step": {k: v[1] for k, v in stage_all.items()},
                       "stage_note": "one extra step with every stage bracketed by CUDA events, after the timed region"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "configs": configs,
    }
    if mgpu_parity is not None:
        line["mgpu_parity"] = mgpu_parity
    if world == 1 and not args.no_cpu:
        from oracle import features as of
        from rs_image_segmentation_b200.synth import synth_raster_numpy
        S, G = args.cpu_sample, args.cpu_glcm_sample
        crop = synth_raster_numpy(S, S, 7, np.uint8, seed=7000)          # the same sample as `--impl reference`
        nir = of.robust_normalize(crop[:G, :G, 3].astype(np.float32))
        stage_cpu, per_px, cores = cpu_baseline_measure(crop, nir, cfg, K, T, 7000, 1, 1)
        line["cpu_baseline"] = {
            "value": 1e-6 / per_px, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": cpu_sample_text(S, G, K, T),
            "stage_us_per_pixel": {k: v * 1e6 for k, v in stage_cpu.items()}}
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
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--cpu-glcm-sample", type=int, default=768)
    ap.add_argument("--configs", default=os.environ.get("RSX_BENCH_CONFIGS", "C,D,E"), help="extra BASELINE.json configurations to measure (comma list of C, D, E; empty = none)")
    ap.add_argument("--d-size", type=int, default=10980)
    ap.add_argument("--e-size", type=int, default=40000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_rsx(args)


if __name__ == "__main__":
    sys.exit(main())
