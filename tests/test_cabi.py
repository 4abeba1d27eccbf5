"""The C-ABI library loads on a machine without a GPU and exports every symbol include/rsx.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rsx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from rs_image_segmentation_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rsx.h but not exported by librsx.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in include/rsx.h"


def test_host_only_entry_points():
    import numpy as np
    from rs_image_segmentation_b200 import _lib
    from rs_image_segmentation_b200.device import hptr
    lib = _lib.load()
    assert lib.rsx_abi_version() == 1
    assert lib.rsx_kmeans_state_bytes() > 1000
    mn = np.array([-1.5, 0.0, 3.25], np.float32)
    mx = np.array([-0.5, 0.0, 1e30], np.float32)
    enc = np.zeros(6, np.uint32)
    lib.rsx_minmax_encode(hptr(mn), hptr(mx), 3, hptr(enc))
    a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.rsx_minmax_decode(hptr(enc), 3, hptr(a), hptr(b))
    assert np.array_equal(a, mn) and np.array_equal(b, mx)
    assert enc[0] < enc[2] < enc[4]          # order preserving


def test_argument_errors_are_reported():
    from rs_image_segmentation_b200 import _lib
    lib = _lib.load()
    rc = lib.rsx_hist_u8(None, 10, 7, None, None)
    assert rc == 1 and b"bad arguments" in lib.rsx_last_error()
    with pytest.raises(_lib.RsxError):
        _lib.call("rsx_glcm_props", None, 1, 1, 32, 7, 1, 1, 1, None, 0, None)


def test_no_cpu_fallback_without_gpu():
    import torch
    from rs_image_segmentation_b200 import _lib, device
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RsxError):
        device.require_cuda()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rs_image_segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_peer_and_readback_entry_points_validate_arguments():
    """rsx_kmeans_update_peers / rsx_peer_* / rsx_store_to_host reject bad arguments before touching a device."""
    import ctypes as C
    from rs_image_segmentation_b200 import _lib
    lib = _lib.load()
    blocks = (C.c_void_p * 2)(0x1000, 0x2000)
    dummy = C.c_void_p(0x1000)
    assert lib.rsx_kmeans_update_peers(dummy, dummy, 1, 13, None, blocks, 0, 1, 1, None) == 1      # one rank is not a peer group
    assert b"ranks" in lib.rsx_last_error()
    assert lib.rsx_kmeans_update_peers(dummy, dummy, 1, 13, None, blocks, 2, 2, 1, None) == 1      # rank out of range
    assert lib.rsx_kmeans_update_peers(dummy, dummy, 1, 13, None, blocks, 0, 2, 0, None) == 1      # sequence numbers start at 1
    assert lib.rsx_kmeans_update_peers(dummy, dummy, 1, 13, None, None, 0, 2, 1, None) == 1
    assert lib.rsx_store_to_host(None, dummy, 16, None) == 1
    assert lib.rsx_store_to_host(dummy, dummy, 0, None) == 0                                       # nothing to copy
    assert lib.rsx_peer_open(None, None) == 1
    assert lib.rsx_peer_close(None) == 0 and lib.rsx_peer_free(None) == 0


def test_committed_bench_line_keeps_the_contract():
    """The bench line committed under profiles/ carries every key of the measurement contract."""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench.json")))
    assert files, "no bench line under profiles/"
    d = json.load(open(files[-1]))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] >= r["algorithmic_bytes_per_launch"]            # measured DRAM bytes cannot be below the algorithmic ones
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == []
