"""ctypes loader for the in-tree C-ABI library (include/aim_b200.h).

There is deliberately no fallback: if libaim_b200.so is missing the import fails loudly, and if
no sm_100 GPU is visible every alignment call returns AIM_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("AIM_B200_LIB", _HERE / "libaim_b200.so"))


class AimParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "algo", "match", "mismatch", "gap_open", "gap_ext", "max_score", "read_size",
        "backtrace", "reduce", "ngpus", "device", "arena_mb", "variant")] + [("reserved", C.c_int32 * 3)]


class AimResult(C.Structure):
    _fields_ = [("max_operations", C.c_int32), ("begin_offset", C.c_int32), ("end_offset", C.c_int32),
                ("score", C.c_int32), ("status", C.c_int32), ("idx", C.c_uint32)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"aim_b200: {LIB_PATH} not found. Build it with `make` (or __graft_entry__.build()); "
            "there is no Python/CPU fallback for the alignment path.")
    lib = C.CDLL(str(LIB_PATH))
    P = C.POINTER
    vp, i32p, cp = C.c_void_p, P(C.c_int32), C.c_char_p
    sig = {
        "aim_align_batch": (C.c_int, [P(AimParams), C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, P(C.c_double)]),
        "aim_align_device": (C.c_int, [P(AimParams), C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, vp,
                                       P(C.c_float), P(C.c_int32)]),
        "aim_align_batch_cigars": (C.c_int, [P(AimParams), C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_int32, P(C.c_double)]),
        "aim_align_packed": (C.c_int, [P(AimParams), C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_int32, P(C.c_double)]),
        "aim_pack_pairs": (C.c_int, [C.c_uint32, C.c_int32, vp, vp, vp, vp, vp, vp, C.c_int32]),
        "aim_packed_row_bytes": (C.c_int32, [C.c_int32]),
        "aim_op_runs_pitch": (C.c_int32, [C.c_int32]),
        "aim_str_rows_pitch": (C.c_int32, [C.c_int32, C.c_int32]),
        "aim_op_rows_download_bytes": (C.c_int32, [P(AimParams)]),
        "aim_expand_op_runs": (C.c_int, [vp, C.c_int32, C.c_uint32, C.c_int32, vp, vp, C.c_uint32, P(C.c_uint32)]),
        "aim_write_results_packed": (C.c_int, [cp, C.c_uint32, vp, vp, C.c_int32]),
        "aim_align_file": (C.c_int, [P(AimParams), cp, cp, C.c_uint32, C.c_uint32, P(C.c_uint64), P(C.c_uint32), P(C.c_double), P(C.c_int32)]),
        "aim_host_alloc": (vp, [C.c_size_t]),
        "aim_host_free": (None, [vp]),
        "aim_shutdown": (None, []),
        "aim_device_count": (C.c_int, []),
        "aim_measure_int_peak": (C.c_int, [C.c_int, P(C.c_double)]),
        "aim_last_error": (cp, []),
        "aim_strerror": (cp, [C.c_int]),
        "aim_abi_version": (C.c_int, []),
        "aim_derive_knobs": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_int32, i32p, i32p]),
        "aim_pairs_to_process": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint32]),
        "aim_read_pairs": (C.c_int64, [cp, C.c_uint32, C.c_int32, vp, vp, vp, vp]),
        "aim_count_pairs": (C.c_int64, [cp]),
        "aim_write_results": (C.c_int, [cp, C.c_uint32, C.c_int32, C.c_int32, vp, vp]),
        "aim_write_results_genasm": (C.c_int, [cp, C.c_uint32, C.c_int32, C.c_int32, vp, vp]),
        "aim_cigar_rle": (C.c_int, [vp, C.c_int32, C.c_int32, vp, C.c_size_t]),
        "aim_generate_pairs": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32, C.c_double, C.c_int32,
                                         vp, vp, vp, vp, C.c_int32]),
        "aim_write_pairs": (C.c_int, [cp, C.c_uint32, C.c_int32, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTED = ["aim_align_file", "aim_align_batch", "aim_align_device", "aim_align_batch_cigars", "aim_align_packed", "aim_pack_pairs", "aim_packed_row_bytes", "aim_write_results_packed", "aim_host_alloc", "aim_host_free", "aim_shutdown",
            "aim_device_count", "aim_measure_int_peak", "aim_last_error", "aim_strerror", "aim_abi_version", "aim_derive_knobs",
            "aim_pairs_to_process", "aim_read_pairs", "aim_count_pairs", "aim_write_results", "aim_write_results_genasm", "aim_cigar_rle",
            "aim_generate_pairs", "aim_write_pairs", "aim_op_runs_pitch", "aim_str_rows_pitch", "aim_op_rows_download_bytes", "aim_expand_op_runs"]
