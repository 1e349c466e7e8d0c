"""TEST INFRASTRUCTURE (oracle/): ctypes face of oracle/aim_oracle.c (libaim_oracle.so).

Used only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs as the CHECKER;
aim_b200/ never imports it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libaim_oracle.so"
_ALGO = {"nw": 0, "swg": 1, "wfa": 2, "genasm_dc": 3, "genasm_filter": 4}
GENASM_UNDEFINED, GENASM_NOALIGN = 3, 4  # statuses of pairs whose reference output is not a function of the pair

ORC_RESULT = np.dtype([("max_operations", "<i4"), ("begin_offset", "<i4"), ("end_offset", "<i4"),
                       ("score", "<i4"), ("status", "<i4")])


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("algo", "match", "mismatch", "gap_open", "gap_ext", "max_score",
                                         "read_size", "backtrace", "reduce", "variant")]


def build(force: bool = False) -> Path:
    src = HERE / "aim_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-o", str(LIB), str(src), "-lpthread"], check=True)
    return LIB


_lib = None


def _get():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_align_batch.restype = C.c_int
        _lib.orc_align_batch.argtypes = [C.POINTER(OrcParams), C.c_uint32] + [C.c_void_p] * 6 + [C.c_int]
        _lib.orc_check_batch.restype = C.c_int
        _lib.orc_check_batch.argtypes = ([C.POINTER(OrcParams), C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32, C.c_uint32, C.c_int] +
                                         [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)])
    return _lib


def align(algo: str, plen, tlen, patterns, texts, *, max_score: int, read_size: int, match: int = 0,
          mismatch: int = 3, gap_open: int = 4, gap_ext: int = 1, backtrace: bool = True, reduce: bool = False,
          nthreads: int = 1, variant: int = 0):
    """-> (results[ORC_RESULT], ops[n, 2*read_size] uint8 | None)."""
    n = len(plen)
    p = OrcParams(_ALGO[algo], match, mismatch, gap_open, gap_ext, max_score, read_size, int(backtrace), int(reduce), int(variant))
    plen = np.ascontiguousarray(plen, np.int32)
    tlen = np.ascontiguousarray(tlen, np.int32)
    patterns = np.ascontiguousarray(patterns, np.uint8)
    texts = np.ascontiguousarray(texts, np.uint8)
    assert patterns.shape == (n, read_size) and texts.shape == (n, read_size)
    res = np.zeros(n, ORC_RESULT)
    if algo == "genasm_dc":
        backtrace = True   # the ops rows carry the CIGAR string (genasmDC.c:641-676)
    elif algo == "genasm_filter":
        backtrace = False
    ops = np.zeros((n, 2 * read_size), np.uint8) if backtrace else None
    rc = _get().orc_align_batch(C.byref(p), n, plen.ctypes.data, tlen.ctypes.data, patterns.ctypes.data,
                                texts.ctypes.data, res.ctypes.data, ops.ctypes.data if backtrace else None, nthreads)
    if rc != 0:
        raise RuntimeError(f"orc_align_batch rc={rc}")
    return res, ops


def check(algo: str, plen, tlen, patterns, texts, cand_results, cand_ops, *, max_score: int, read_size: int, match: int = 0,
          mismatch: int = 3, gap_open: int = 4, gap_ext: int = 1, backtrace: bool = True, reduce: bool = False,
          nthreads: int = 1, variant: int = 0, stride: int = 1, offset: int = 0) -> dict:
    """Full-size parity: align pairs offset, offset+stride, ... here and compare each with the candidate's
    24-byte result (score, status, and with backtrace begin/end/max_operations) and the op bytes of its valid
    span, in place.  -> {"pairs_checked", "mismatches", "first_bad"}."""
    n = len(plen)
    p = OrcParams(_ALGO[algo], match, mismatch, gap_open, gap_ext, max_score, read_size, int(backtrace), int(reduce), int(variant))
    plen = np.ascontiguousarray(plen, np.int32)
    tlen = np.ascontiguousarray(tlen, np.int32)
    assert patterns.flags.c_contiguous and texts.flags.c_contiguous and cand_results.flags.c_contiguous
    assert patterns.shape == (n, read_size) and texts.shape == (n, read_size)
    assert cand_results.dtype.itemsize == 24 and len(cand_results) == n
    if cand_ops is not None:
        assert cand_ops.flags.c_contiguous and cand_ops.shape == (n, 2 * read_size)
    ck, mm, fb = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
    rc = _get().orc_check_batch(C.byref(p), n, plen.ctypes.data, tlen.ctypes.data, patterns.ctypes.data, texts.ctypes.data,
                                cand_results.ctypes.data, cand_ops.ctypes.data if cand_ops is not None else None,
                                int(stride), int(offset), int(nthreads), C.byref(ck), C.byref(mm), C.byref(fb))
    if rc != 0:
        raise RuntimeError(f"orc_check_batch rc={rc}")
    return {"pairs_checked": int(ck.value), "mismatches": int(mm.value), "first_bad": None if fb.value == 0xFFFFFFFF else int(fb.value)}
