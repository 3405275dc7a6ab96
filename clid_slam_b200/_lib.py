"""ctypes binding of libclid_sdf.so (C ABI declared in include/clid_sdf.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing or
a tensor is not on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLID_LIB_PATH", os.path.join(_PKG, "libclid_sdf.so"))  # override: kernel experiments

MAX_LEVELS = 3
MAX_KNN = 8
MAX_KC = 256

# enum ClidFlags
TRAINING_MODE = 1 << 0
QUERY_LOCALLY = 1 << 1
TIME_FILTER = 1 << 2
LAYER_NORM = 1 << 3
LEAKY_RELU = 1 << 4
USE_BRICKS = 1 << 5
TC_DECODER = 1 << 6

_f32p = C.c_void_p  # device pointers travel as plain addresses


class ClidBrickHeader(C.Structure):  # element type of ClidBricks.headers (the host only sizes the array)
    _fields_ = [("mask", C.c_uint64), ("base", C.c_int32), ("count", C.c_int32)]


class ClidBricks(C.Structure):
    _fields_ = [
        ("headers", C.c_void_p), ("hood", C.c_void_p), ("records", C.c_void_p), ("stencil", C.c_void_p),
        ("origin", C.c_int32 * 3), ("dims", C.c_int32 * 3), ("span", C.c_int32), ("reach", C.c_int32),
        ("n_records", C.c_int32), ("apron", C.c_int32),
    ]


class ClidMap(C.Structure):
    _fields_ = [
        ("buffer_pt_index", C.c_void_p), ("buffer_size", C.c_int64), ("primes", C.c_int64 * 3),
        ("neural_points", C.c_void_p), ("point_ts_create", C.c_void_p), ("n_global", C.c_int64),
        ("travel_dist", C.c_void_p), ("n_travel", C.c_int32), ("cur_ts", C.c_int32),
        ("diff_travel_dist_local", C.c_float), ("resolution", C.c_float), ("max_valid_dist2", C.c_float),
        ("kc", C.c_int32), ("neighbor_dx", C.c_void_p), ("global2local", C.c_void_p),
        ("gather_points", C.c_void_p), ("gather_features", C.c_void_p), ("gather_certainties", C.c_void_p),
        ("certainty_accum", C.c_void_p), ("gather_ts_update", C.c_void_p), ("n_gather", C.c_int64),
        ("feature_dim", C.c_int32), ("knn", C.c_int32), ("bricks", C.POINTER(ClidBricks)),
        ("work_counter", C.c_void_p),
    ]


class ClidDecoder(C.Structure):
    _fields_ = [
        ("weight", C.c_void_p * MAX_LEVELS), ("bias", C.c_void_p * MAX_LEVELS),
        ("out_weight", C.c_void_p), ("out_bias", C.c_void_p),
        ("in_dim", C.c_int32), ("hidden_dim", C.c_int32), ("levels", C.c_int32), ("sdf_scale", C.c_float),
    ]


class ClidQueryOut(C.Structure):
    _fields_ = [
        ("sdf", C.c_void_p), ("grad", C.c_void_p), ("z", C.c_void_p), ("weights", C.c_void_p),
        ("knn_idx", C.c_void_p), ("nn_count", C.c_void_p), ("certainty", C.c_void_p),
    ]


class ClidLossArgs(C.Structure):
    _fields_ = [
        ("sdf", C.c_void_p), ("grad", C.c_void_p), ("label", C.c_void_p), ("weight", C.c_void_p),
        ("dlogit", C.c_void_p), ("dgrad", C.c_void_p), ("loss", C.c_void_p),
        ("n", C.c_int64), ("nd", C.c_int64), ("n_norm", C.c_int64), ("nd_norm", C.c_int64),
        ("sdf_scale", C.c_float), ("weight_e", C.c_float), ("num_eps", C.c_float), ("weighted", C.c_int32),
    ]


class ClidTrainFusedArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ts", C.c_void_p), ("label", C.c_void_p), ("weight", C.c_void_p),
        ("n", C.c_int64), ("n_norm", C.c_int64), ("nd_norm", C.c_int64),
        ("weight_e", C.c_float), ("num_eps", C.c_float), ("weighted", C.c_int32), ("numerical", C.c_int32),
        ("gfeat", C.c_void_p), ("touched", C.c_void_p), ("dec_grad", C.c_void_p), ("loss", C.c_void_p),
        ("sdf_out", C.c_void_p), ("peer_grad", C.c_void_p * 2), ("peer_axis", C.c_int32), ("peer_band", C.c_int32 * 4),
        ("peer_row", C.c_void_p * 2),
        ("scratch", C.c_void_p), ("scratch_bytes", C.c_size_t),
    ]


MAX_DEC_TENSORS = 2 * MAX_LEVELS + 2


class ClidAdamArgs(C.Structure):
    _fields_ = [
        ("feat", C.c_void_p), ("feat_grad", C.c_void_p), ("feat_m", C.c_void_p), ("feat_v", C.c_void_p),
        ("touched", C.c_void_p), ("rows", C.c_int64),
        ("dec_param", C.c_void_p * MAX_DEC_TENSORS), ("dec_numel", C.c_int32 * MAX_DEC_TENSORS),
        ("dec_tensors", C.c_int32),
        ("dec_grad", C.c_void_p), ("dec_m", C.c_void_p), ("dec_v", C.c_void_p),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("weight_decay", C.c_float), ("step", C.c_int32), ("step_state", C.c_void_p),
    ]


class ClidPeerArgs(C.Structure):
    _fields_ = [
        ("slots_of", C.c_void_p * 8), ("flags_of", C.c_void_p * 8), ("epoch", C.c_void_p),
        ("rank", C.c_int32), ("world", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32), ("stride", C.c_int32),
        ("timeout_ms", C.c_int32), ("error", C.c_void_p),
    ]


class ClidLocalCloud(C.Structure):
    _fields_ = [
        ("table", C.c_void_p), ("buffer_size", C.c_int64), ("primes", C.c_int64 * 3), ("points", C.c_void_p),
        ("n_points", C.c_int64), ("neighbor_idx", C.c_void_p), ("kc", C.c_int32), ("resolution", C.c_float),
        ("max_valid_range", C.c_float),
    ]


class ClidReplayPool(C.Structure):
    _fields_ = [
        ("coord", C.c_void_p), ("sdf_label", C.c_void_p), ("weight", C.c_void_p), ("time", C.c_void_p),
        ("count", C.c_int64), ("new_idx", C.c_void_p), ("n_new", C.c_int64), ("bs_new", C.c_int32),
    ]


class ClidMappingArgs(C.Structure):
    _fields_ = [
        ("pool", ClidReplayPool), ("train", ClidTrainFusedArgs), ("adam", ClidAdamArgs),
        ("iters", C.c_int32), ("seed", C.c_uint64), ("offset", C.c_uint64), ("loss_history", C.c_void_p),
    ]


class ClidInsertArgs(C.Structure):
    _fields_ = [
        ("cand", C.c_void_p), ("n", C.c_int64), ("buffer_pt_index", C.c_void_p), ("buffer_size", C.c_int64),
        ("primes", C.c_int64 * 3), ("neural_points", C.c_void_p), ("ts_update", C.c_void_p), ("travel_dist", C.c_void_p),
        ("m", C.c_int64), ("n_travel", C.c_int64), ("cur_ts", C.c_int32), ("all_fresh", C.c_int32),
        ("resolution", C.c_float), ("far2", C.c_float), ("diff_travel_dist_local", C.c_float),
        ("slot", C.c_void_p), ("owner", C.c_void_p), ("fresh", C.c_void_p), ("rank", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class ClidWindowArgs(C.Structure):
    _fields_ = [
        ("neural_points", C.c_void_p), ("ts_create", C.c_void_p), ("ts_update", C.c_void_p), ("travel_dist", C.c_void_p),
        ("m", C.c_int64), ("n_travel", C.c_int64), ("sensor", C.c_double * 3), ("radius2", C.c_double),
        ("sensor_is_f64", C.c_int32), ("temporal", C.c_int32), ("use_mid_ts", C.c_int32), ("cur_ts", C.c_int32),
        ("reboot_test", C.c_int32), ("reboot_ts", C.c_int32), ("diff_ts_local", C.c_int32),
        ("diff_travel_dist_local", C.c_float),
        ("flags", C.c_void_p), ("global2local", C.c_void_p), ("local_mask", C.c_void_p), ("gids", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class ClidWindowRows(C.Structure):
    _fields_ = [
        ("gids", C.c_void_p), ("n_local", C.c_int64), ("m", C.c_int64),
        ("neural_points", C.c_void_p), ("point_orientations", C.c_void_p), ("point_certainties", C.c_void_p),
        ("point_ts_update", C.c_void_p), ("geo_features", C.c_void_p),
        ("local_points", C.c_void_p), ("local_orientations", C.c_void_p), ("local_certainties", C.c_void_p),
        ("local_ts_update", C.c_void_p), ("local_features", C.c_void_p),
    ]


class ClidRaySampleArgs(C.Structure):
    _fields_ = [
        ("points", C.c_void_p), ("depth", C.c_void_p), ("randn_surf", C.c_void_p), ("rand_front", C.c_void_p),
        ("rand_behind", C.c_void_p), ("n_points", C.c_int64), ("n_surf", C.c_int32), ("n_front", C.c_int32),
        ("n_behind", C.c_int32), ("surface_sample_range_m", C.c_float), ("margin", C.c_float),
        ("free_sample_begin_ratio", C.c_float), ("free_sample_end_dist_m", C.c_float), ("weight_top", C.c_float),
        ("dist_weight_scale", C.c_float), ("max_range", C.c_float),
        ("dist_weight_on", C.c_int32), ("coord", C.c_void_p), ("disp", C.c_void_p), ("weight", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None

# every symbol include/clid_sdf.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("clid_version", C.c_int, []),
    ("clid_last_error", C.c_char_p, []),
    ("clid_query_forward", C.c_int,
     [C.POINTER(ClidMap), C.POINTER(ClidDecoder), C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32,
      C.POINTER(ClidQueryOut), C.c_void_p]),
    ("clid_query_backward", C.c_int,
     [C.POINTER(ClidMap), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p,
      C.c_void_p]),
    ("clid_query_backward_backward", C.c_int,
     [C.POINTER(ClidMap), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p,
      C.c_void_p, C.c_void_p]),
    ("clid_sdf_loss", C.c_int, [C.POINTER(ClidLossArgs), C.c_void_p]),
    ("clid_train_backward", C.c_int,
     [C.POINTER(ClidMap), C.POINTER(ClidDecoder), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
      C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_train_fused_scratch_bytes", C.c_size_t, [C.c_int64, C.c_int32]),
    ("clid_train_fused_scratch_bytes_for", C.c_size_t, [C.POINTER(ClidDecoder), C.c_int64, C.c_int32]),
    ("clid_train_fused", C.c_int,
     [C.POINTER(ClidMap), C.POINTER(ClidDecoder), C.POINTER(ClidTrainFusedArgs), C.c_uint32, C.c_void_p]),
    ("clid_decoder_grad_reduce", C.c_int,
     [C.POINTER(ClidDecoder), C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p]),
    ("clid_adam_step", C.c_int, [C.POINTER(ClidAdamArgs), C.c_void_p]),
    ("clid_adam_advance", C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    ("clid_step_begin", C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    ("clid_enable_peer_access", C.c_int, [C.c_int32]),
    ("clid_peer_alloc", C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    ("clid_peer_free", C.c_int, [C.c_void_p]),
    ("clid_ipc_open", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    ("clid_ipc_close", C.c_int, [C.c_void_p]),
    ("clid_peer_publish", C.c_int, [C.POINTER(ClidPeerArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_peer_reduce", C.c_int, [C.POINTER(ClidPeerArgs), C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_brick_keep", C.c_int,
     [C.POINTER(ClidMap), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_brick_keys", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    ("clid_brick_fill", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_scan_workspace_bytes", C.c_size_t, [C.c_int64]),
    ("clid_voxel_keys", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_voxel_pick", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_map_insert_probe", C.c_int, [C.POINTER(ClidInsertArgs), C.c_void_p]),
    ("clid_map_insert_commit", C.c_int, [C.POINTER(ClidInsertArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_local_window_select", C.c_int, [C.POINTER(ClidWindowArgs), C.c_void_p]),
    ("clid_local_window_gather", C.c_int, [C.POINTER(ClidWindowRows), C.c_void_p]),
    ("clid_local_window_scatter", C.c_int, [C.POINTER(ClidWindowRows), C.c_void_p]),
    ("clid_pool_filter_select", C.c_int,
     [C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
      C.c_size_t, C.c_void_p]),
    ("clid_ray_samples", C.c_int, [C.POINTER(ClidRaySampleArgs), C.c_void_p]),
    ("clid_ray_labels", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_flag_ranks", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("clid_table_store", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    ("clid_compact_rows", C.c_int,
     [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.c_void_p]),
    ("clid_region_sdf", C.c_int, [C.POINTER(ClidLocalCloud), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_registration_terms", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.c_int32, C.c_float, C.c_float,
      C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_draw_batch", C.c_int,
     [C.POINTER(ClidReplayPool), C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
      C.c_void_p, C.c_void_p]),
    ("clid_mapping_run", C.c_int,
     [C.POINTER(ClidMap), C.POINTER(ClidDecoder), C.POINTER(ClidMappingArgs), C.c_uint32, C.c_void_p]),
    ("clid_radius_search", C.c_int,
     [C.POINTER(ClidMap), C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_query_certainty", C.c_int,
     [C.POINTER(ClidMap), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("clid_decoder_eval", C.c_int,
     [C.POINTER(ClidDecoder), C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
]


def exported_symbols():
    return [s[0] for s in _SIGNATURES]


def load() -> C.CDLL:
    """Load libclid_sdf.so (built in-tree by clid_slam_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m clid_slam_b200.build`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in _SIGNATURES:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().clid_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")


def ptr(t: Optional[torch.Tensor], dtype: torch.dtype, what: str) -> Optional[int]:
    """Device address of a contiguous CUDA tensor of the given dtype (None passes through)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} is on {t.device}: the neural-SDF hot path runs only on CUDA (no CPU fallback)"
        )
    if t.dtype != dtype:
        raise TypeError(f"{what}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    return t.data_ptr()


def current_stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream
