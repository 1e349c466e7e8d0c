"""Python face of the C ABI: numpy in, numpy out.  All compute happens in libaim_b200.so on the GPU.

Mirrors the reference's surface for the offloaded path: the knobs of */run-*-pim-*.py and
*/common/common.h as an `AlignParams`, the `>pattern / <text` pair files, the `idx, score, ` +
CIGAR output (reference: WFA/DPU-MRAM/host/host.c:91-134, 332-353).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from ._lib import AimParams, AimResult, lib

ALGO_NW, ALGO_SWG, ALGO_WFA, ALGO_GENASM_DC, ALGO_GENASM_FILTER = 0, 1, 2, 3, 4
_ALGO = {"nw": ALGO_NW, "swg": ALGO_SWG, "wfa": ALGO_WFA, "genasm_dc": ALGO_GENASM_DC, "genasm_filter": ALGO_GENASM_FILTER}
STATUS_GENASM_UNDEFINED, STATUS_GENASM_NOALIGN = 3, 4

RESULT_DTYPE = np.dtype([("max_operations", "<i4"), ("begin_offset", "<i4"), ("end_offset", "<i4"),
                         ("score", "<i4"), ("status", "<i4"), ("idx", "<u4")])
assert RESULT_DTYPE.itemsize == C.sizeof(AimResult)


class AimError(RuntimeError):
    def __init__(self, code: int):
        self.code = code
        super().__init__(f"aim_b200 error {code} ({lib.aim_strerror(code).decode()}): {lib.aim_last_error().decode()}")


@dataclass
class AlignParams:
    """Runtime form of the reference's -D knobs (run-wfa-pim-mram.py:133-139)."""
    algo: str = "wfa"
    match: int = 0
    mismatch: int = 3
    gap_open: int = 4      # NW: the single linear gap (GAP_I = GAP_D)
    gap_ext: int = 1
    max_score: int = 250
    read_size: int = 112
    backtrace: bool = False
    reduce: bool = False
    ngpus: int = 1
    device: int = 0
    arena_mb: int = 0
    variant: int = 0       # GenASM-DC: 1 = DPU-MRAM-DC ('S' for substitutions, pattern 'N' is no wildcard)

    def __post_init__(self):
        # GenASM-DC always returns its CIGAR string in the ops rows; the filter has none (include/aim_b200.h)
        if self.algo == "genasm_dc":
            self.backtrace = True
        elif self.algo == "genasm_filter":
            self.backtrace = False

    def to_c(self) -> AimParams:
        p = AimParams()
        p.algo = _ALGO[self.algo]
        p.match, p.mismatch, p.gap_open, p.gap_ext = self.match, self.mismatch, self.gap_open, self.gap_ext
        p.max_score, p.read_size = self.max_score, self.read_size
        p.backtrace, p.reduce = int(self.backtrace), int(self.reduce)
        p.ngpus, p.device, p.arena_mb, p.variant = self.ngpus, self.device, self.arena_mb, self.variant
        return p


def derive_knobs(algo: str, read_length: int, error: float, mismatch: int = 3, gap_open: int = 4,
                 gap_ext: int = 1) -> tuple[int, int]:
    """(MAX_SCORE, READ_SIZE) as run-*-pim-*.py derives them (run-wfa-pim-mram.py:58-67)."""
    ms, rs = C.c_int32(), C.c_int32()
    rc = lib.aim_derive_knobs(_ALGO[algo], read_length, float(error), mismatch, gap_open, gap_ext, C.byref(ms), C.byref(rs))
    if rc != 0:
        raise AimError(rc)
    return ms.value, rs.value


def pairs_to_process(pairs_in_file: int, n_arg: int, nr_dpus: int = 1) -> int:
    return int(lib.aim_pairs_to_process(pairs_in_file, n_arg, nr_dpus))


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory (freed with the object)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self.ptr = lib.aim_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError(f"aim_host_alloc({self.nbytes}) failed: {lib.aim_last_error().decode()}")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            self.array = None
            lib.aim_host_free(ptr)


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def count_pairs(path: str | os.PathLike) -> int:
    n = lib.aim_count_pairs(os.fsencode(path))
    if n < 0:
        raise AimError(int(n))
    return int(n)


def read_pairs(path: str | os.PathLike, read_size: int, max_pairs: int | None = None):
    """get_reads (host.c:91-134) -> (plen, tlen, patterns[n, read_size], texts[n, read_size])."""
    total = count_pairs(path)
    want = total if max_pairs is None else min(total, max_pairs)
    plen = np.zeros(want, np.int32)
    tlen = np.zeros(want, np.int32)
    pats = np.zeros((want, read_size), np.uint8)
    txts = np.zeros((want, read_size), np.uint8)
    n = lib.aim_read_pairs(os.fsencode(path), want, read_size, _ptr(plen), _ptr(tlen), _ptr(pats), _ptr(txts))
    if n < 0:
        raise AimError(int(n))
    return plen[:n], tlen[:n], pats[:n], txts[:n]


def generate_pairs(seed: int, n: int, length: int, error: float, read_size: int, first_pair: int = 0,
                   nthreads: int = 0, out=None):
    """Synthetic pairs (generate_dataset semantics, Datasets/README.md:19-25)."""
    if out is None:
        plen, tlen = np.zeros(n, np.int32), np.zeros(n, np.int32)
        pats, txts = np.zeros((n, read_size), np.uint8), np.zeros((n, read_size), np.uint8)
    else:
        plen, tlen, pats, txts = out
    nthreads = nthreads or (os.cpu_count() or 1)
    rc = lib.aim_generate_pairs(seed, first_pair, n, length, float(error), read_size,
                                _ptr(plen), _ptr(tlen), _ptr(pats), _ptr(txts), nthreads)
    if rc != 0:
        raise AimError(rc)
    return plen, tlen, pats, txts


def write_pairs(path, plen, tlen, pats, txts) -> None:
    rc = lib.aim_write_pairs(os.fsencode(path), len(plen), pats.shape[1], _ptr(plen), _ptr(tlen), _ptr(pats), _ptr(txts))
    if rc != 0:
        raise AimError(rc)


def align_batch(params: AlignParams, plen, tlen, patterns, texts, idx_base: int = 0, results=None, ops=None):
    """aim_align_batch on host arrays -> (results[RESULT_DTYPE], ops[n, 2*read_size] | None, phase_ms[3])."""
    n = len(plen)
    rs = params.read_size
    plen = np.ascontiguousarray(plen, np.int32)
    tlen = np.ascontiguousarray(tlen, np.int32)
    patterns = np.ascontiguousarray(patterns, np.uint8)
    texts = np.ascontiguousarray(texts, np.uint8)
    if n and (patterns.shape != (n, rs) or texts.shape != (n, rs)):
        raise ValueError(f"patterns/texts must be [n, read_size={rs}]")
    if results is None:
        results = np.zeros(n, RESULT_DTYPE)
    if ops is None and params.backtrace:
        ops = np.zeros((n, 2 * rs), np.uint8)
    phase = (C.c_double * 3)()
    p = params.to_c()
    rc = lib.aim_align_batch(C.byref(p), n, idx_base, _ptr(plen), _ptr(tlen), _ptr(patterns), _ptr(texts),
                             _ptr(results), _ptr(ops) if params.backtrace else None, phase)
    if rc != 0:
        raise AimError(rc)
    return results, (ops if params.backtrace else None), list(phase)


def align_file(params: AlignParams, pairs_path, out_path, n_arg: int, nr_dpus: int = 1):
    """aim_align_file: `host <pairs> <out> <N>` as one streaming call, parsed and printed on the GPU
    -> (pairs_done, status_mask, phase_ms[3], launches)."""
    done, mask, nl = C.c_uint64(0), C.c_uint32(0), C.c_int32(0)
    phase = (C.c_double * 3)()
    p = params.to_c()
    rc = lib.aim_align_file(C.byref(p), os.fsencode(pairs_path), os.fsencode(out_path), n_arg, nr_dpus, C.byref(done), C.byref(mask), phase, C.byref(nl))
    if rc != 0:
        raise AimError(rc)
    return int(done.value), int(mask.value), list(phase), int(nl.value)


def align_device(params: AlignParams, n: int, d_plen: int, d_tlen: int, d_patterns: int, d_texts: int,
                 d_results: int, d_ops: int | None, stream: int = 0, device: int = 0, timed: bool = True):
    """aim_align_device on raw device addresses -> (kernel_ms | None, launches)."""
    p = params.to_c()
    ms = C.c_float(0.0)
    nl = C.c_int32(0)
    rc = lib.aim_align_device(C.byref(p), device, n, 0, d_plen, d_tlen, d_patterns, d_texts, d_results,
                              d_ops if params.backtrace else None, stream or None,
                              C.byref(ms) if timed else None, C.byref(nl))
    if rc != 0:
        raise AimError(rc)
    return (ms.value if timed else None), nl.value


def align_batch_cigars(params: AlignParams, plen, tlen, patterns, texts, cigar_pitch: int = 64, idx_base: int = 0, results=None, cigars=None):
    """aim_align_batch_cigars: reference-layout inputs, CIGAR text rows out -> (results, cigars[n, cigar_pitch] uint8, phase_ms[3])."""
    n = len(plen)
    plen = np.ascontiguousarray(plen, np.int32)
    tlen = np.ascontiguousarray(tlen, np.int32)
    patterns = np.ascontiguousarray(patterns, np.uint8)
    texts = np.ascontiguousarray(texts, np.uint8)
    if results is None:
        results = np.zeros(n, RESULT_DTYPE)
    if cigars is None:
        cigars = np.zeros((n, cigar_pitch), np.uint8)
    phase = (C.c_double * 3)()
    p = params.to_c()
    rc = lib.aim_align_batch_cigars(C.byref(p), n, idx_base, _ptr(plen), _ptr(tlen), _ptr(patterns), _ptr(texts), _ptr(results),
                                    _ptr(cigars), cigar_pitch, phase)
    if rc != 0:
        raise AimError(rc)
    return results, cigars, list(phase)


def op_runs_pitch(read_size: int) -> int:
    """Bytes of a run row (the form in which aim_align_batch downloads an op row); 0 = rows of this size travel as they are."""
    return int(lib.aim_op_runs_pitch(read_size))


def str_rows_pitch(read_size: int, max_score: int) -> int:
    """GenASM-DC: bytes of every op row (the head holding its CIGAR string) that aim_align_batch downloads; 0 = whole rows."""
    return int(lib.aim_str_rows_pitch(read_size, max_score))


def op_rows_download_bytes(params: "AlignParams") -> int:
    """Bytes per pair that cross PCIe for the op row under these parameters (0 = the 2*read_size row itself)."""
    p = params.to_c()
    return int(lib.aim_op_rows_download_bytes(C.byref(p)))


def expand_op_runs(runs: np.ndarray, read_size: int, ops: np.ndarray | None = None):
    """Host half of aim_align_batch's op-row download: run rows [n, pitch] -> (ops[n, 2*read_size], overflow pair numbers)."""
    runs = np.ascontiguousarray(runs, np.uint8)
    n, pitch = runs.shape
    if ops is None:
        ops = np.zeros((n, 2 * read_size), np.uint8)
    ov = np.zeros(max(n, 1), np.uint32)
    cnt = C.c_uint32(0)
    rc = lib.aim_expand_op_runs(_ptr(runs), pitch, n, read_size, _ptr(ops), _ptr(ov), len(ov), C.byref(cnt))
    if rc != 0:
        raise AimError(rc)
    return ops, ov[:cnt.value].copy()


def packed_row_bytes(read_size: int) -> int:
    return int(lib.aim_packed_row_bytes(read_size))


def pack_pairs(plen, tlen, patterns, texts, read_size: int, nthreads: int = 0, out=None):
    """aim_pack_pairs: ASCII rows -> (packed[n, 2, row_bytes/4] uint32, flags[ceil(n/32)] uint32)."""
    n = len(plen)
    words = packed_row_bytes(read_size) // 4
    if out is None:
        packed, flags = np.zeros((n, 2, words), np.uint32), np.zeros((n + 31) // 32, np.uint32)
    else:
        packed, flags = out
    rc = lib.aim_pack_pairs(n, read_size, _ptr(np.ascontiguousarray(plen, np.int32)), _ptr(np.ascontiguousarray(tlen, np.int32)),
                            _ptr(patterns), _ptr(texts), _ptr(packed), _ptr(flags), nthreads)
    if rc != 0:
        raise AimError(rc)
    return packed, flags


def align_packed(params: AlignParams, plen, tlen, packed, flags, cigar_pitch: int = 64, idx_base: int = 0, results=None, cigars=None):
    """aim_align_packed on host arrays -> (results[RESULT_DTYPE], cigars[n, cigar_pitch] uint8, phase_ms[3])."""
    n = len(plen)
    plen = np.ascontiguousarray(plen, np.int32)
    tlen = np.ascontiguousarray(tlen, np.int32)
    if results is None:
        results = np.zeros(n, RESULT_DTYPE)
    if cigars is None:
        cigars = np.zeros((n, cigar_pitch), np.uint8)
    phase = (C.c_double * 3)()
    p = params.to_c()
    rc = lib.aim_align_packed(C.byref(p), n, idx_base, _ptr(plen), _ptr(tlen), _ptr(packed), _ptr(flags), _ptr(results), _ptr(cigars),
                              cigar_pitch, phase)
    if rc != 0:
        raise AimError(rc)
    return results, cigars, list(phase)


def write_results_packed(path, results: np.ndarray, cigars: np.ndarray) -> None:
    rc = lib.aim_write_results_packed(os.fsencode(path), len(results), _ptr(results), _ptr(cigars), cigars.shape[1])
    if rc != 0:
        raise AimError(rc)


def cigar_strings(results: np.ndarray, ops: np.ndarray) -> list[str]:
    """RLE CIGARs as edit_cigar_print writes them (host.c:69-89)."""
    out = []
    buf = C.create_string_buffer(ops.shape[1] * 12 + 16)
    for i in range(len(results)):
        ln = lib.aim_cigar_rle(_ptr(ops[i]), int(results["begin_offset"][i]), int(results["end_offset"][i]), buf, len(buf))
        if ln < 0:
            raise AimError(-5)
        out.append(buf.raw[:ln].decode())
    return out


def write_results(path, results: np.ndarray, ops: np.ndarray | None, read_size: int, backtrace: bool) -> None:
    rc = lib.aim_write_results(os.fsencode(path), len(results), read_size, int(backtrace), _ptr(results),
                               _ptr(ops) if backtrace else None)
    if rc != 0:
        raise AimError(rc)


def write_results_genasm(path, results: np.ndarray, cigars: np.ndarray | None, read_size: int, dc: bool) -> None:
    """GenASM output lines: "idx, score, CIGAR" (DC) or "idx, score" (filter)."""
    rc = lib.aim_write_results_genasm(os.fsencode(path), len(results), read_size, int(dc), _ptr(results), _ptr(cigars) if dc else None)
    if rc != 0:
        raise AimError(rc)


def measure_int_peak(device: int = 0) -> float:
    """Measured INT32 ALU ceiling in ops/s (aim_measure_int_peak)."""
    v = C.c_double(0.0)
    rc = lib.aim_measure_int_peak(device, C.byref(v))
    if rc != 0:
        raise AimError(rc)
    return v.value


def device_count() -> int:
    return int(lib.aim_device_count())


def shutdown() -> None:
    lib.aim_shutdown()
