/* aim_b200.h — C ABI of the B200-native batched pairwise aligner (drop-in for the path
 * safaad/aim offloads to UPMEM DPUs).  Plain pointers and sizes only; no CUDA or torch types.
 *
 * The reference has no plugin/FFI API: its boundary is (1) the process `host <pairs> <out> <N>`
 * and (2) the six dpu_push_xfer transfers + dpu_launch inside host main()
 * (WFA/DPU-MRAM/host/host.c:186-330; the other five hosts are line-for-line equivalent).
 * aim_align_batch() replaces (2); tools/host.cpp rebuilds (1) on top of it.  All citations are
 * paths under the reference checkout.
 */
#ifndef AIM_B200_H
#define AIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIM_B200_ABI_VERSION 1

/* Which reference program is being replaced. */
#define AIM_ALGO_NW 0  /* NW/DPU-{WRAM,MRAM}: linear gap, int16 flat table (NW/DPU-WRAM/dpu/nw.c:109-153)      */
#define AIM_ALGO_SWG 1 /* SWG/DPU-MRAM: gap-affine, int16 cells, MAX_SCORE borders (SWG/DPU-MRAM/dpu/swg.c:151) */
#define AIM_ALGO_WFA 2 /* WFA/DPU-{WRAM,MRAM}: gap-affine WFA, +adaptive with reduce=1 (WFA/DPU-MRAM/dpu/wfa.c:356) */
/* aim-genasm submodule.  max_score = k error levels; results: score = the reference's
 * "bitmacScore" (DC, genasmDC.c:553) or the minimum error level (filter, genasm_filter.c:226-238; -1 = none). */
#define AIM_ALGO_GENASM_DC 3     /* GenASM/DPU-{WRAM,MRAM}-DC: Bitap + traceback (aim-genasm/GenASM/DPU-WRAM-DC/dpu/genasmDC.c:338) */
#define AIM_ALGO_GENASM_FILTER 4 /* GenASM/DPU-{WRAM,MRAM}-filter (aim-genasm/GenASM/DPU-WRAM-filter/dpu/genasm_filter.c:52)        */

/* Return codes (the reference aborts the process through DPU_ASSERT/exit instead). */
#define AIM_OK 0
#define AIM_ERR_ARG (-1)        /* NULL/invalid argument or penalty set                                  */
#define AIM_ERR_LENGTH (-2)     /* a sequence is longer than read_size (host.c:119-123 exits the process) */
#define AIM_ERR_CUDA (-3)       /* CUDA runtime failure; aim_last_error() has the text                    */
#define AIM_ERR_NO_DEVICE (-4)  /* no usable sm_100 device: there is NO CPU fallback                     */
#define AIM_ERR_IO (-5)         /* file could not be opened / written                                     */
#define AIM_ERR_NOMEM (-6)

/* Per-pair status (aim_result.status).  0 everywhere in a healthy run. */
#define AIM_STATUS_OK 0
#define AIM_STATUS_BACKTRACE 1 /* reference would print "No link found"/"No backtrace operation found" and exit(1) */
#define AIM_STATUS_ARENA 2     /* wavefront history exceeded the per-pair arena (reference: "Out of memory MRAM") */
#define AIM_STATUS_GENASM_UNDEFINED 3 /* GenASM-DC: the reference's traceback reads rows it never wrote for this pair (a text
                                       * byte outside ACGTacgt, or the walk reaches text row n): its output is not a function
                                       * of the pair; no CIGAR is returned                                                     */
#define AIM_STATUS_NEEDS_ASCII 5      /* aim_align_packed: the pair holds a byte outside {A,C,G,T} (the reference compares raw bytes, wfa.c:209):
                                       * serve it through aim_align_batch                                                          */
#define AIM_STATUS_CIGAR_OVERFLOW 6   /* aim_align_packed: the CIGAR text does not fit cigar_pitch; score is valid                      */
#define AIM_STATUS_GENASM_NOALIGN 4   /* GenASM-DC: "No alignment found!" within max_score errors (genasmDC.c:543-547), score -1 */

/* The compile-time knobs of the reference (-DMAX_SCORE -DREAD_SIZE -DMATCH -DMISMATCH -DGAP_O
 * -DGAP_E [-DGAP_I -DGAP_D] [-DBACKTRACE] [-DREDUCE]; WFA/DPU-MRAM/run-wfa-pim-mram.py:133-139)
 * as runtime parameters.  NW uses gap_open as its single linear gap (GAP_I = GAP_D) and ignores
 * gap_ext, match and max_score; WFA ignores match (assumed 0); reduce is WFA only. */
typedef struct aim_params {
    int32_t algo;       /* AIM_ALGO_*                                                  */
    int32_t match;      /* MATCH    (<= 0)                                             */
    int32_t mismatch;   /* MISMATCH (> 0)                                              */
    int32_t gap_open;   /* GAP_O, or GAP_I = GAP_D for NW (> 0)                        */
    int32_t gap_ext;    /* GAP_E (> 0)                                                 */
    int32_t max_score;  /* MAX_SCORE: WFA give-up threshold, SWG border value          */
    int32_t read_size;  /* READ_SIZE: row pitch of patterns/texts; multiple of 8       */
    int32_t backtrace;  /* BACKTRACE: produce ops / begin_offset                       */
    int32_t reduce;     /* REDUCE: WFA-adaptive (min length 10, max distance 50)       */
    int32_t ngpus;      /* <=1: one device; N: shard pairs contiguously over N devices */
    int32_t device;     /* first device ordinal                                        */
    int32_t arena_mb;   /* long-read WFA history arena per resident pair in MiB (0 = default) */
    int32_t variant;    /* GenASM-DC: 0 = DPU-WRAM-DC semantics, 1 = DPU-MRAM-DC ('S' for substitutions, pattern 'N' no wildcard);
                         * SWG: 0 = DPU-MRAM (int16 cells), 1 = DPU-WRAM (int8 cells when max_score < 127, SWG/DPU-WRAM/common/common.h:71-79) */
    int32_t reserved[3];
} aim_params;

/* result_t of the reference (WFA/DPU-MRAM/common/common.h:179-187; NW/DPU-WRAM/common/common.h:122-130)
 * with the padding word carrying the per-pair status. */
typedef struct aim_result {
    int32_t max_operations; /* pattern_len + text_len                                        */
    int32_t begin_offset;   /* first valid op; ops[begin_offset .. end_offset) is the CIGAR  */
    int32_t end_offset;     /* == max_operations                                             */
    int32_t score;
    int32_t status;         /* AIM_STATUS_*                                                  */
    uint32_t idx;           /* 0-based pair number                                           */
} aim_result;

/* GenASM through the same two entry points (aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:150-300 is the same host
 * skeleton): algo = AIM_ALGO_GENASM_DC writes pair i's CIGAR STRING - the reference's own format, run lengths with
 * REVERSED decimal digits, NUL-terminated (genasmDC.c:128-135,323-335) - at ops + i*2*read_size and sets
 * max_operations = strlen + 1 (result_t.max_operations), begin_offset = 0, end_offset = strlen; backtrace is implied.
 * algo = AIM_ALGO_GENASM_FILTER returns scores only (ops may be NULL). */

/* ---- the device boundary (replaces host.c:186-330) ----------------------------------------
 * Align n pairs held in HOST memory.  Layout is the reference host's own (host.c:126-131,
 * 203-205, 305-311): pair i's pattern at patterns + i*read_size (plen[i] bytes, compared as raw
 * bytes), its text at texts + i*read_size, its ops at ops + i*2*read_size ('M'-filled, valid span
 * [begin_offset,end_offset)), its result at results[i] with idx = idx_base + i.
 * ops may be NULL iff !backtrace.  phase_ms (may be NULL) receives the three phases the reference
 * prints: [0] "CPU-DPU" (H2D), [1] "DPU Kernel", [2] "DPU-CPU" (D2H), summed over chunks as
 * measured by CUDA events; chunks are double-buffered so the phases overlap in wall time.
 * Synchronous; the caller owns every buffer; nothing is retained.  Buffers obtained from
 * aim_host_alloc() are DMA'd directly, others are staged through pinned chunks.
 * Thread-safety: one call at a time per process (as the reference). */
int aim_align_batch(const aim_params *params, uint32_t n, uint32_t idx_base,
                    const int32_t *plen, const int32_t *tlen,
                    const char *patterns, const char *texts,
                    aim_result *results, char *ops, double phase_ms[3]);

/* Same work with every buffer already resident in the HBM of `device` (all pointers are device
 * pointers; patterns/texts/ops 16-byte aligned).  stream is a cudaStream_t passed as void*
 * (NULL = default stream).  Asynchronous w.r.t. the host unless kernel_ms != NULL, in which
 * case the kernels are bracketed by CUDA events on `stream`, synchronised, and the elapsed
 * device time is returned.  launches (may be NULL) receives the number of kernels launched. */
int aim_align_device(const aim_params *params, int device, uint32_t n, uint32_t idx_base,
                     const int32_t *d_plen, const int32_t *d_tlen,
                     const char *d_patterns, const char *d_texts,
                     aim_result *d_results, char *d_ops,
                     void *stream, float *kernel_ms, int32_t *launches);

/* ---- how aim_align_batch brings the op rows back (NW, SWG, WFA: read_size 32..1024; GenASM-DC: see below) ----------
 * The caller's `ops` buffer is filled exactly as documented above, but the rows do not cross PCIe as they are: the device-to-
 * host direction is the scarcer one of a multi-GPU host (all GPUs together: 72-94 GB/s against 110-187 GB/s host-to-device,
 * DESIGN.md 6.2) and a 2*read_size row is 'M' but for a handful of runs.  A kernel turns every op row into a RUN ROW of
 * aim_op_runs_pitch(read_size) bytes - 32-bit words: the number of runs, then per run of bytes other than 'M' over the whole
 * row: position | length (1..255) << 16 | op << 24; first word 0xffffffff = more runs than the row holds - and host threads
 * (AIM_HOST_THREADS; default: the process's share of the cores, at most 8) rebuild the rows into `ops` with cache-bypassing
 * stores while later chunks are in flight; rows that did not fit are fetched as they are.  AIM_SPARSE_OPS=0 moves the rows
 * as they are.
 * aim_expand_op_runs is that host half (exported for tests and for callers that keep run rows): rebuilds n rows, lists the
 * pairs whose run row carries the mark (ascending, at most overflow_cap of *overflow_count). */
int32_t aim_op_runs_pitch(int32_t read_size);
/* GenASM-DC rows hold the DPU's CIGAR string: the first aim_str_rows_pitch(read_size, max_score) bytes of every row - room for
 * the longest string max_score error levels can make - cross PCIe and each string is copied, with its NUL, to the head of the
 * caller's row (bytes behind the NUL are not written); 0 = the rows move as they are (the heads would exceed half a row). */
int32_t aim_str_rows_pitch(int32_t read_size, int32_t max_score);
/* Bytes per pair that cross PCIe for the op row under these parameters (0 = the row itself): aim_op_runs_pitch(read_size), for
 * WFA cut down to the 4 * (1 + max_score / min(mismatch, gap_open + gap_ext)) bytes an alignment within max_score can need, or
 * aim_str_rows_pitch for GenASM-DC. */
int32_t aim_op_rows_download_bytes(const aim_params *params);
int aim_expand_op_runs(const unsigned char *runs, int32_t pitch, uint32_t n, int32_t read_size, char *ops,
                       uint32_t *overflow, uint32_t overflow_cap, uint32_t *overflow_count);

/* ---- compact transfers (extension; WFA short reads) -------------------------------------------------
 * aim_align_batch serves the reference host's own buffers: 2*READ_SIZE ASCII bytes in and the 2*READ_SIZE op row out per
 * pair (680 B at READ_SIZE 168; the ASCII rows cross PCIe, the op rows are rebuilt on the host), which is what bounds it: one
 * PCIe Gen5 x16 link carries 1.2-1.3e8 pairs/s of that layout and the host's memory ~2.3e8 pairs/s for all GPUs of a box
 * together, while one B200 aligns 3.5e8 pairs/s.  This entry moves the
 * information instead: sequences 2 bits per base (aim_pack_pairs: what get_reads would produce, host.c:91-134), and per
 * pair the CIGAR TEXT the reference prints (edit_cigar_print, host.c:69-89: "40M1D59M", NUL-terminated) in a row of
 * cigar_pitch bytes, i.e. what the print loop (host.c:340-350) needs, ready to write.  Same scores, same CIGARs.
 *   packed : n x 2 x aim_packed_row_bytes(read_size) bytes, pair i's pattern row then its text row; 16 bases per 32-bit
 *            word, first base in the two most significant bits, codes A,C,T,G = 0,1,2,3 ((c >> 1) & 3), 'A' past the end
 *   flags  : one bit per pair (bit i & 31 of word i >> 5): set by aim_pack_pairs for a pair with a byte outside {A,C,G,T}
 *            inside its sequences; such pairs come back with AIM_STATUS_NEEDS_ASCII
 *   cigars : n x cigar_pitch bytes (cigar_pitch a multiple of 16, 16..2*read_size); AIM_STATUS_CIGAR_OVERFLOW if too small
 * params->algo must be AIM_ALGO_WFA with backtrace; configurations the short-read kernel does not serve return AIM_ERR_ARG.
 * Buffers from aim_host_alloc() are DMA'd in place. */
/* Half of that without touching the caller's input layout: aim_align_batch with the op rows kept on the device and the CIGAR
 * TEXT returned instead (any algorithm with a CIGAR; GenASM-DC: its own string).  Inputs exactly as aim_align_batch; cigars =
 * n x cigar_pitch bytes, NUL-terminated rows; AIM_STATUS_CIGAR_OVERFLOW marks a pair whose text does not fit. */
int aim_align_batch_cigars(const aim_params *params, uint32_t n, uint32_t idx_base,
                           const int32_t *plen, const int32_t *tlen, const char *patterns, const char *texts,
                           aim_result *results, char *cigars, int32_t cigar_pitch, double phase_ms[3]);
int32_t aim_packed_row_bytes(int32_t read_size);
int aim_pack_pairs(uint32_t n, int32_t read_size, const int32_t *plen, const int32_t *tlen, const char *patterns,
                   const char *texts, uint32_t *packed, uint32_t *flags, int32_t nthreads);
int aim_align_packed(const aim_params *params, uint32_t n, uint32_t idx_base, const int32_t *plen, const int32_t *tlen,
                     const uint32_t *packed, const uint32_t *flags, aim_result *results, char *cigars, int32_t cigar_pitch,
                     double phase_ms[3]);
/* "%d, %d, \n" + the CIGAR row + "\n" per pair: the reference's output bytes from aim_align_packed's results. */
int aim_write_results_packed(const char *path, uint32_t n, const aim_result *results, const char *cigars, int32_t cigar_pitch);

/* ---- the process boundary at device rate (extension; every algorithm) -----------------------------------------
 * `host <pairs-file> <out-file> <N>` (host.c:136-379) as one streaming call: the pair file is read in chunks, PARSED ON THE
 * GPU (get_reads, host.c:91-134: first and last character of every line dropped unchecked, length = line length - 2,
 * a sequence longer than read_size -> AIM_ERR_LENGTH and an empty output file, as the reference exits before printing),
 * aligned, and the output lines ("%d, %d, \n" + run-length CIGAR + "\n", host.c:332-353, 69-89) are FORMATTED ON THE GPU;
 * host threads only read() and write().  GenASM: the aim-genasm hosts' lines, "%d, %d, %s\n" with the DPU's CIGAR string (DC) /
 * "%d, %d\n" (filter) (aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:286-296).  Pairs processed = min(pairs in file, nr_dpus * roundup8(n_arg / nr_dpus))
 * (host.c:191,201-209) -> *pairs_done.  *status_mask = OR of (1 << AIM_STATUS_*) over all pairs: on AIM_STATUS_BACKTRACE /
 * AIM_STATUS_ARENA the output file is left empty (the reference's DPU program exits before anything is printed).
 * params->ngpus > 1: chunk c goes to GPU c % ngpus, the writer keeps pair order. */
int aim_align_file(const aim_params *params, const char *pairs_path, const char *out_path, uint32_t n_arg, uint32_t nr_dpus,
                   uint64_t *pairs_done, uint32_t *status_mask, double phase_ms[3], int32_t *launches);

/* Pinned host memory for zero-staging transfers (cudaHostAlloc / cudaFreeHost). */
void *aim_host_alloc(size_t bytes);
void aim_host_free(void *p);

/* Release cached device/pinned buffers and streams. */
void aim_shutdown(void);

/* Measured INT32 ALU ceiling of `device` in operations/s: a dependent-free add/logic/min-max mix on
 * every SM (the roofline denominator for the integer DP and wavefront loops). */
int aim_measure_int_peak(int device, double *ops_per_s);

int aim_device_count(void);              /* number of CUDA devices visible, 0 if none                   */
const char *aim_last_error(void);        /* text of the last error on this thread                       */
const char *aim_strerror(int code);
int aim_abi_version(void);

/* ---- host-side logic shared by the CLI and the run-*-pim-*.py wrappers (no GPU needed) ----- */

/* MAX_SCORE / READ_SIZE exactly as the run scripts derive them with Python floats
 * (WFA/DPU-MRAM/run-wfa-pim-mram.py:58-67, NW/DPU-MRAM/run-nw-pim-mram.py:51-60):
 *   w = l*e;  MAX_SCORE = ceil(max(w*x, w*(g+a)))   (NW: ceil(w*g));  READ_SIZE = ceil((l+w+7)/8)*8 */
int aim_derive_knobs(int32_t algo, int32_t read_length, double error, int32_t mismatch,
                     int32_t gap_open, int32_t gap_ext, int32_t *max_score, int32_t *read_size);

/* How many pairs the reference host aligns: min(pairs in file, nr_dpus*roundup8(N/nr_dpus))
 * (host.c:191,201-209); N <= nr_dpus is rejected by the caller as in host.c:180-184. */
uint32_t aim_pairs_to_process(uint32_t pairs_in_file, uint32_t n_arg, uint32_t nr_dpus);

/* get_reads (host.c:91-134): read up to max_pairs pairs of lines (">PATTERN\n" / "<TEXT\n"; the
 * first and last character of every line are dropped unchecked).  Buffers are n x read_size.
 * Returns the number of pairs read (>= 0), AIM_ERR_LENGTH if a sequence exceeds read_size,
 * AIM_ERR_IO if the file cannot be opened. */
int64_t aim_read_pairs(const char *path, uint32_t max_pairs, int32_t read_size,
                       int32_t *plen, int32_t *tlen, char *patterns, char *texts);
/* Count pairs (complete line pairs) in a file; AIM_ERR_IO on failure. */
int64_t aim_count_pairs(const char *path);

/* The reference's result printer (host.c:332-353 + edit_cigar_print :69-89): per pair
 * "%d, %d, \n" (idx, score) and, if backtrace, the run-length CIGAR of ops[begin..end) + "\n". */
int aim_write_results(const char *path, uint32_t n, int32_t read_size, int32_t backtrace,
                      const aim_result *results, const char *ops);
/* The GenASM printers: dc != 0 -> "%d, %d, %s\n" (idx, score, pair i's NUL-terminated CIGAR string at cigars + i*2*read_size;
 * aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:286-296); dc == 0 -> "%d, %d\n" (DPU-WRAM-filter/host/host.c:272). */
int aim_write_results_genasm(const char *path, uint32_t n, int32_t read_size, int32_t dc,
                             const aim_result *results, const char *cigars);
/* RLE one pair's ops into out (capacity cap); returns the length written (no NUL) or -1. */
int aim_cigar_rle(const char *ops, int32_t begin_offset, int32_t end_offset, char *out, size_t cap);

/* Synthetic pairs with WFA `generate_dataset` semantics (Datasets/README.md:19-25): pattern =
 * `length` i.i.d. uniform ACGT; text = copy with ceil(length*error) edits, each uniformly a
 * mismatch to a different base / a 1-base deletion / a 1-base insertion at a uniform position.
 * Deterministic in (seed, first_pair + i) whatever the thread count. */
int aim_generate_pairs(uint64_t seed, uint64_t first_pair, uint32_t n, int32_t length, double error,
                       int32_t read_size, int32_t *plen, int32_t *tlen, char *patterns, char *texts,
                       int32_t nthreads);
/* Write pairs in the Datasets file format. */
int aim_write_pairs(const char *path, uint32_t n, int32_t read_size, const int32_t *plen,
                    const int32_t *tlen, const char *patterns, const char *texts);

#ifdef __cplusplus
}
#endif
#endif /* AIM_B200_H */
