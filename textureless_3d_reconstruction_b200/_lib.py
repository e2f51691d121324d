"""ctypes binding of libt3d.so (the C ABI declared in include/t3d.h).

There is no CPU fallback: importing works anywhere (so host-only helpers and
the symbol-export test run without a GPU), but creating a context — which every
compute path needs — raises if the library is missing or no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libt3d.so"

T3D_OK, T3D_E_INVALID, T3D_E_CUDA, T3D_E_CAPACITY, T3D_E_IO, T3D_E_NUMERIC = 0, -1, -2, -3, -4, -5
PLY_O3D_BINARY, PLY_REF_ASCII = 0, 1


class T3DError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libt3d error {code}: {msg}")
        self.code = code


class BackprojectParams(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("subsample", C.c_int32), ("depth_is_f64", C.c_int32),
        ("scale_is_f64", C.c_int32), ("has_pose", C.c_int32), ("rgb_out_f32", C.c_int32),
        ("has_color", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("scale", C.c_double), ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("R", C.c_double * 9), ("t", C.c_double * 3),
    ]


class BackprojectFrame(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("bgr", C.c_void_p), ("conf_mask", C.c_void_p),
                ("R", C.c_double * 9), ("t", C.c_double * 3)]


class TsdfParams(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_float), ("sdf_trunc", C.c_float), ("block_res", C.c_int32),
        ("pixel_round", C.c_int32), ("block_capacity", C.c_int64), ("hash_capacity", C.c_int64),
    ]


class FrameView(C.Structure):
    _fields_ = [
        ("depth", C.c_void_p), ("bgr", C.c_void_p), ("K", C.c_float * 4), ("T_cw", C.c_float * 12),
        ("conf_mask", C.c_void_p),
    ]


class IcpResult(C.Structure):
    _fields_ = [
        ("T", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double),
        ("iterations", C.c_int32), ("converged", C.c_int32), ("correspondences", C.c_int64),
    ]


_lib = None

# name -> (restype, argtypes)
_VP, _I, _I64, _D, _F = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float
SEQUENCE_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)  # t3d_sequence_hook(user, phase, event_a, event_b)

_SIGS = {
    "t3d_last_error": (C.c_char_p, []),
    "t3d_version": (_I, []),
    "t3d_create": (_VP, [_I]),
    "t3d_destroy": (None, [_VP]),
    "t3d_launch_count": (_I64, [_VP]),
    "t3d_backproject": (_I, [_VP, _VP, _VP, _VP, C.POINTER(BackprojectParams), _VP, _VP, _I64, _VP, _VP]),
    "t3d_backproject_batch": (_I, [_VP, C.POINTER(BackprojectFrame), _I, C.POINTER(BackprojectParams), _VP, _VP, _I64,
                                   _VP, _VP]),
    "t3d_voxel_downsample": (_I, [_VP, _VP, _I, _VP, _I64, _D, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    "t3d_voxel_partials": (_I, [_VP, _VP, _I, _VP, _I64, _D, _VP, _VP, _I, _VP, _I64, _VP, _VP, _VP]),
    "t3d_voxel_merge_partials": (_I, [_VP, _VP, _I64, _I, _D, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP]),
    "t3d_bounds": (_I, [_VP, _VP, _I, _I64, _VP, _VP, _VP]),
    "t3d_statistical_outlier": (_I, [_VP, _VP, _I64, _I, _D, _VP, _VP, _VP, _VP, _VP]),
    "t3d_sor_mean_distances_part": (_I, [_VP, _VP, _I64, _I, _I, _I, _VP, _VP]),
    "t3d_sor_from_mean_distances": (_I, [_VP, _VP, _I64, _D, _VP, _VP, _VP, _VP]),
    "t3d_compact_rows": (_I, [_VP, _VP, _I64, C.c_int32, _VP, _VP, _VP, _VP]),
    "t3d_tsdf_create": (_I, [_VP, C.POINTER(TsdfParams), C.POINTER(_VP)]),
    "t3d_tsdf_destroy": (None, [_VP]),
    "t3d_tsdf_reset": (_I, [_VP, _VP]),
    "t3d_tsdf_integrate": (_I, [_VP, C.POINTER(FrameView), _I, _I, _I, _I, _F, _F, _VP]),
    "t3d_tsdf_integrate_sequence": (_I, [_VP, C.POINTER(FrameView), _I, _I, _I, _I, _I, _F, _F, _VP]),
    "t3d_tsdf_touch": (_I, [_VP, C.POINTER(FrameView), _I, _I, _I, _F, _F, _VP, _I64, _VP, _VP]),
    "t3d_tsdf_num_blocks": (_I64, [_VP, _VP]),
    "t3d_tsdf_counters": (_I, [_VP, _VP, _VP]),
    "t3d_tsdf_set_profiling": (_I, [_VP, _I]),
    "t3d_tsdf_get_profile": (_I, [_VP, _VP, _VP]),
    "t3d_tsdf_export_blocks": (_I, [_VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP]),
    "t3d_tsdf_export_blocks_range": (_I, [_VP, _I, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, _I64, _VP, _VP]),
    "t3d_tsdf_export_blocks_outside": (_I, [_VP, _I, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, _I64, _VP, _VP]),
    "t3d_tsdf_route_counts": (_I, [_VP, _I, C.c_int32, _I, _I, _VP, _VP]),
    "t3d_tsdf_route_export": (_I, [_VP, _I, C.c_int32, _I, _I, _VP, _VP, _VP, _VP]),
    "t3d_tsdf_merge_records": (_I, [_VP, _VP, _I64, _VP]),
    "t3d_tsdf_route_counts_upto": (_I, [_VP, _I, C.c_int32, _I, _I, _VP, _VP, _VP]),
    "t3d_tsdf_route_export_upto": (_I, [_VP, _I, C.c_int32, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "t3d_tsdf_merge_records_multi": (_I, [_VP, _VP, _I64, _I64, _I, _I, _I64, _VP]),
    "t3d_memcpy_async": (_I, [_VP, _VP, C.c_size_t, _VP]),
    "t3d_tsdf_route_export_p2p": (_I, [_VP, _I, C.c_int32, _I, _I, _VP, _VP, _I64, _VP, _VP, _VP]),
    "t3d_tsdf_integrate_sequence_hooked": (_I, [_VP, C.POINTER(FrameView), _I, _I, _I, _I, _I, _F, _F, _VP,
                                                SEQUENCE_HOOK, _VP, _VP, _I, _VP]),
    "t3d_event_create": (_I, [C.POINTER(_VP)]),
    "t3d_event_destroy": (_I, [_VP]),
    "t3d_event_record": (_I, [_VP, _VP]),
    "t3d_stream_wait_event": (_I, [_VP, _VP]),
    "t3d_tsdf_merge_records_dev": (_I, [_VP, _VP, _VP, _I64, _VP]),
    "t3d_ipc_alloc": (_I, [_VP, C.c_size_t, C.POINTER(_VP), _VP]),
    "t3d_ipc_open": (_I, [_VP, _VP, C.POINTER(_VP)]),
    "t3d_ipc_close": (_I, [_VP, _VP]),
    "t3d_ipc_free": (_I, [_VP, _VP]),
    "t3d_memset_async": (_I, [_VP, _I, C.c_size_t, _VP]),
    "t3d_tsdf_merge_blocks": (_I, [_VP, _VP, _VP, _VP, _VP, _I64, _VP]),
    "t3d_tsdf_extract_points": (_I, [_VP, _F, _VP, _VP, _VP, _I64, _VP, _VP]),
    "t3d_tsdf_extract_points_view": (_I, [_VP, C.POINTER(FrameView), _I, _I, _F, _F, _VP, _VP, _VP, _I64, _VP, _VP, _VP]),
    "t3d_tsdf_extract_points_range": (_I, [_VP, _I, C.c_int32, C.c_int32, _F, _VP, _VP, _VP, _I64, _VP, _VP]),
    "t3d_tsdf_extract_mesh": (_I, [_VP, _F, _VP, _VP, _VP, _I64, _VP, _I64, _VP, _VP]),
    "t3d_write_ply_mesh_h": (_I, [C.c_char_p, _VP, _VP, _VP, _I64, _VP, _I64]),
    "t3d_estimate_normals": (_I, [_VP, _VP, _I64, _I, _VP, _VP, _VP]),
    "t3d_icp_point_to_plane": (_I, [_VP, _VP, _I64, _VP, _VP, _I64, _D, _VP, _I, _D, _D, C.POINTER(IcpResult), _VP]),
    "t3d_icp_correspondences": (_I, [_VP, _VP, _I64, _VP, _VP, _I64, _D, _VP, _VP, _VP, _VP, _VP]),
    "t3d_icp_point_to_plane_dev": (_I, [_VP, _VP, _I64, _VP, _VP, _VP, _I64, _VP, _I, _D, _VP, _I, _D, _D,
                                        C.POINTER(IcpResult), C.POINTER(C.c_int), _VP]),
    "t3d_icp_linearize": (_I, [_VP, _VP, _I64, _VP, _VP, _I64, _D, _VP, _VP, _VP, _VP]),
    "t3d_nearest_neighbor": (_I, [_VP, _VP, _I64, _VP, _I64, _D, _VP, _VP, _VP]),
    "t3d_depth_u16_to_f32": (_I, [_VP, _VP, _I64, _F, _VP, _VP]),
    "t3d_depth_f32_to_u16": (_I, [_VP, _VP, _I64, _F, _VP, _VP]),
    "t3d_resize_bilinear_f32": (_I, [_VP, _VP, _I, _I, _VP, _I, _I, _VP]),
    "t3d_estimate_scale": (_I, [_VP, _VP, _I, _I, _VP, _VP, _I64, _I, _I, _VP, _VP, _VP]),
    "t3d_pack_pointcloud2": (_I, [_VP, _VP, _VP, _I, _I64, _VP, _VP]),
    "t3d_write_ply_h": (_I, [C.c_char_p, _VP, _I, _VP, _VP, _I64, _I]),
    "t3d_synth_frame": (_I, [_VP, _I, _I, _I, _I, _D, _D, _D, _D, C.c_uint64, _F, _VP, _VP, _VP, _VP]),
}


def load():
    """Load libt3d.so (building it first if the sources are newer and nvcc is here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise ImportError(
                f"{LIB_PATH} is missing and could not be built ({e}). "
                "This package has no CPU fallback: run `python __graft_entry__.py build`.") from e
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().t3d_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc != T3D_OK:
        raise T3DError(rc, last_error())


def declared_symbols() -> list[str]:
    return sorted(_SIGS)
