"""ctypes binding of librsx.so (include/rsx.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librsx.so")

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol declared in include/rsx.h
SIGNATURES = {
    "rsx_last_error": (C.c_char_p, []),
    "rsx_abi_version": (i32, []),
    "rsx_launch_count": (i64, []),
    "rsx_set_option": (i32, [C.c_char_p, i32]),
    "rsx_get_option": (i32, [C.c_char_p, i32]),
    "rsx_store_to_host": (i32, [vp, vp, i64, vp]),
    "rsx_raster_stats": (i32, [vp, i32, i32, i32, f64, f64, vp, vp, vp, vp, vp, vp]),
    "rsx_hist_u8": (i32, [vp, i64, i32, vp, vp]),
    "rsx_hist_u16": (i32, [vp, i64, i32, vp, vp]),
    "rsx_indices_fused_u8": (i32, [vp, i64, i32, vp, vp, vp, vp, i64, vp, vp, vp, i32, vp, vp]),
    "rsx_indices_fused_u8_dev": (i32, [vp, i64, i32, vp, vp, vp, vp, i64, vp, vp, i32, vp]),
    "rsx_raster_stats_device_bytes": (i64, []),
    "rsx_raster_stats_device_lut_offset": (i64, []),
    "rsx_raster_stats_u8_device": (i32, [vp, i32, i32, i32, C.c_double, C.c_double, vp, vp]),
    "rsx_indices_fused_u16": (i32, [vp, i64, i32, vp, vp, vp, vp, i64, vp, vp, vp, i32, vp]),
    "rsx_normalize_f32": (i32, [vp, i64, f32, f32, f32, vp, vp]),
    "rsx_index_ratio_f32": (i32, [vp, vp, i64, vp, vp]),
    "rsx_index_evi_f32": (i32, [vp, vp, vp, i64, f32, f32, f32, f32, vp, vp]),
    "rsx_index_msavi_f32": (i32, [vp, vp, i64, vp, vp]),
    "rsx_index_bsi_f32": (i32, [vp, vp, vp, vp, i64, vp, vp]),
    "rsx_quantize_f32": (i32, [vp, i64, f32, f32, f32, i32, vp, vp]),
    "rsx_pca_scratch_elems": (i64, [i32]),
    "rsx_pca_moments_u8": (i32, [vp, i64, i32, vp, vp, vp, vp]),
    "rsx_pca_build_lut_u16": (i32, [vp, vp, vp, i32, vp, vp]),
    "rsx_pca_moments_u16": (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]),
    "rsx_pca_project_u8": (i32, [vp, i64, i32, vp, vp, vp, i32, vp, i64, vp, vp]),
    "rsx_pca_project_u16": (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, i32, vp, i64, vp, vp]),
    "rsx_pca_planar_scratch_elems": (i64, [i32]),
    "rsx_pca_moments_planar_f32": (i32, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
    "rsx_pca_project_planar_f32": (i32, [vp, i64, i64, i32, i32, vp, vp, vp, vp, vp, i32, vp, i64, vp, vp]),
    "rsx_glcm_props": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp]),
    "rsx_glcm_props_offsets": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp, i64, vp]),
    "rsx_glcm_moments": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, i64, vp, vp]),
    "rsx_glcm_counts": (i32, [vp, i32, i32, i32, i32, vp, i32, vp, vp]),
    "rsx_resize_bilinear_f32": (i32, [vp, i32, i32, i32, i32, i64, vp, i32, i32, i32, i32, i64, i32, vp, vp]),
    "rsx_minmax_init": (i32, [vp, i32, vp]),
    "rsx_minmax_planes_f32": (i32, [vp, i64, i64, i32, vp, vp]),
    "rsx_minmax_decode": (None, [vp, i32, vp, vp]),
    "rsx_minmax_encode": (None, [vp, vp, i32, vp]),
    "rsx_nan_to_zero_f32": (i32, [vp, i64, vp]),
    "rsx_nan_to_zero_minmax_f32": (i32, [vp, i64, vp, vp]),
    "rsx_band_lut_u8": (i32, [vp, i64, i32, i32, vp, vp, vp]),
    "rsx_band_lut_f32": (i32, [vp, i64, i32, i32, vp, vp, vp]),
    "rsx_morph_gradient_u8": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    "rsx_local_std_f32": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp]),
    "rsx_sobel_mag_u8": (i32, [vp, i32, i32, i32, i32, vp, i32, i32, vp, vp]),
    "rsx_divide_f32": (i32, [vp, i64, C.c_float, vp]),
    "rsx_u8_over_255_f32": (i32, [vp, i64, vp, vp]),
    "rsx_planes_to_hwc_f64": (i32, [vp, i64, i64, i32, vp, vp]),
    "rsx_labels_plus1_u8": (i32, [vp, i64, vp, vp]),
    "rsx_widen_u8_to_i32": (i32, [vp, vp, i64, i32]),
    "rsx_box_mean_f32": (i32, [vp, i32, i32, i32, i32, i64, vp, i32, i32, i64, i32, i32, vp, vp]),
    "rsx_kmeans_state_bytes": (i64, []),
    "rsx_kmeans_setup": (i32, [vp, i32, i32, vp, vp, vp, vp, i64, vp]),
    "rsx_kmeans_assign": (i32, [vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "rsx_kmeans_setup_device": (i32, [vp, i32, i32, vp, vp, vp, i64, vp]),
    "rsx_kmeans_quantize_u16": (i32, [vp, i64, i64, vp, vp, i64, i32, vp]),
    "rsx_kmeans_assign_q16": (i32, [vp, i64, i64, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
    "rsx_kmeans_aos_stride": (i64, [i32]),
    "rsx_kmeans_assign_bounded": (i32, [vp, i64, i64, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "rsx_kmeans_update": (i32, [vp, vp, i32, i32, vp, vp]),
    "rsx_kmeans_fixed_point_scales": (i32, [vp, vp, vp]),
    "rsx_kpp_scratch_elems": (i64, []),
    "rsx_kpp_block": (i64, []),
    "rsx_kpp_feature_moments": (i32, [vp, i64, i64, i32, vp, vp, vp, vp, vp]),
    "rsx_kpp_distances": (i32, [vp, i64, i64, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]),
    "rsx_kpp_block_sums": (i32, [vp, i64, vp, vp]),
    "rsx_kmeans_read": (i32, [vp, vp, vp, vp, vp]),
    "rsx_kmeans_read_all": (i32, [vp, vp, vp, vp, vp, vp, vp]),
    "rsx_kmeans_update_peers": (i32, [vp, vp, i32, i32, vp, vp, i32, i32, i64, vp]),
    "rsx_peer_alloc": (i32, [i64, vp, vp]),
    "rsx_peer_open": (i32, [vp, vp]),
    "rsx_peer_zero": (i32, [vp, i64, vp]),
    "rsx_peer_close": (i32, [vp]),
    "rsx_peer_free": (i32, [vp]),
}


class RsxError(RuntimeError):
    pass


_lib = None


def load():
    """Load librsx.so and attach the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsxError(
            f"{LIB_PATH} is missing: build it with `python -m rs_image_segmentation_b200.build` "
            "(needs nvcc).  There is no CPU fallback for the rsx hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().rsx_last_error().decode(errors="replace")
        raise RsxError(f"{what or 'rsx call'} failed (code {rc}): {msg}")


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)


def set_option(name: str, value: int):
    """Tuning knob of the library (include/rsx.h rsx_set_option)."""
    check(load().rsx_set_option(name.encode(), int(value)), "rsx_set_option")


def get_option(name: str, default: int) -> int:
    """The knob's value: set_option, else the environment variable RSX_<NAME>, else `default`."""
    return int(load().rsx_get_option(name.encode(), int(default)))


def launch_count() -> int:
    return int(load().rsx_launch_count())
