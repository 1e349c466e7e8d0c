"""aim_b200 — B200-native batched pairwise aligner (NW, SWG, WFA, WFA-adaptive, GenASM-DC/filter), a drop-in for the
path safaad/aim offloads to UPMEM DPUs.  Thin Python over the C ABI in include/aim_b200.h."""
from .api import (expand_op_runs, op_runs_pitch, str_rows_pitch, op_rows_download_bytes, align_batch_cigars, align_packed, pack_pairs, packed_row_bytes, write_results_packed,
                  ALGO_GENASM_DC, ALGO_GENASM_FILTER, ALGO_NW, ALGO_SWG, ALGO_WFA, RESULT_DTYPE, STATUS_GENASM_NOALIGN,
                  STATUS_GENASM_UNDEFINED, AimError, AlignParams, PinnedArray, align_batch,
                  align_device, align_file, cigar_strings, count_pairs, derive_knobs, device_count, generate_pairs,
                  measure_int_peak, pairs_to_process, read_pairs, shutdown, write_pairs, write_results,
                  write_results_genasm)

__all__ = ["expand_op_runs", "op_runs_pitch", "str_rows_pitch", "op_rows_download_bytes", "align_batch_cigars", "align_packed", "pack_pairs", "packed_row_bytes", "write_results_packed", "ALGO_GENASM_DC", "ALGO_GENASM_FILTER", "STATUS_GENASM_NOALIGN", "STATUS_GENASM_UNDEFINED", "write_results_genasm", "ALGO_NW", "ALGO_SWG", "ALGO_WFA", "RESULT_DTYPE", "AimError", "AlignParams", "PinnedArray",
           "align_batch", "align_device", "align_file", "cigar_strings", "count_pairs", "derive_knobs", "device_count",
           "generate_pairs", "measure_int_peak", "pairs_to_process", "read_pairs", "shutdown", "write_pairs", "write_results"]
