// Internal declarations shared by the host logic, the dispatcher and the kernels.
#ifndef AIM_INTERNAL_H
#define AIM_INTERNAL_H

#include <stdint.h>
#include <string>
#include <vector>

#include "aim_b200.h"

namespace aim {

void set_error(const std::string &msg);

// Deterministic (data-independent) wavefront schedule of gap-affine WFA for a penalty set:
// which scores have a wavefront, its [lo,hi] and whether it carries I / D components when no
// adaptive trimming happens.  Trimming only narrows ranges, so these are upper bounds and size
// the per-pair history exactly (WFA/DPU-MRAM/dpu/wfa.c:275-354 computes the same ranges at run time).
struct WfaSchedule {
    uint32_t hist_slots;   // int16 slots needed for the full history up to max_score
    uint32_t max_width;    // widest wavefront
    uint32_t ring_scores;  // scores that must stay live for score-only mode: max(x, o+e) + 1
};
WfaSchedule wfa_schedule(int max_score, int mismatch, int gap_open, int gap_ext);

// Everything a kernel launch needs (device pointers).
struct KernelArgs {
    aim_params p;
    uint32_t n;
    uint32_t idx_base;
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    // packed entry (aim_align_packed): sequences arrive 2-bit packed, CIGARs leave run-length encoded
    const uint32_t *packed = nullptr;  // per pair 2 x aim_packed_row_bytes()/4 words (pattern, then text); patterns/texts are NULL
    const uint32_t *pflags = nullptr;  // one bit per pair: a byte outside {A,C,G,T} - the pair cannot be served packed
    char *cigars = nullptr;            // per pair cigar_pitch bytes: the CIGAR as the reference prints it, NUL-terminated
    int32_t cigar_pitch = 0;
};

// Per-device scratch owned by the dispatcher and handed to the launchers.
struct Scratch {
    void *buf = nullptr;        // general scratch (WFA global arena / DP rows+flags)
    size_t bytes = 0;
    void *plan_cache = nullptr; // device copies of static WFA plans, keyed by content (cached_plan)
    int sm_count = 0;
    int device = 0;
    // the scratch is ONE buffer per device: a launch sequence that uses it records `busy` after its last kernel, and the next
    // sequence on ANOTHER stream waits for it first (scratch_acquire / scratch_release)
    void *busy_event = nullptr;   // cudaEvent_t
    void *busy_stream = nullptr;  // stream the last user ran on
    bool busy_valid = false;
    // GenASM-DC: traceback side stream + events (created on first use, destroyed by aim_shutdown)
    void *side_stream = nullptr;
    void *side_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

// Ensure scratch->buf holds at least `bytes` (grows, never shrinks).  Returns AIM_OK/AIM_ERR_*.
int scratch_reserve(Scratch *s, size_t bytes);
// Device copy of a static plan table: uploaded once per distinct content (synchronously, from the host vector) and reused
// by every later launch, so no per-chunk pageable cudaMemcpyAsync sits on the kernel stream.  NULL on failure (error set).
const void *cached_plan(Scratch *s, const void *host, size_t bytes);
// Order a launch sequence on `stream` after the previous user of the scratch / mark `stream` as its current user.
int scratch_acquire(Scratch *s, void *stream);
int scratch_release(Scratch *s, void *stream);
void scratch_destroy(Scratch *s);

// Launch configuration of the warp-per-pair WFA kernel (aim_wfa.cu), split from the launch so that the
// long-read kernel can size one scratch for itself plus this kernel serving its leftovers.
struct WarpPlan {
    int rc;
    int grid, block;
    bool hg;
    size_t smem_block, meta_bytes, scratch_bytes;
    alignas(8) unsigned char kernel_args[192];
};
WarpPlan wfa_warp_plan(const KernelArgs &a, int sm_count, uint32_t max_pairs);
int wfa_warp_launch(const WarpPlan &W, void *scratch, const uint32_t *list, const uint32_t *list_count, void *stream, int *launches);
// long-read score-only kernel (aim_wfa_long.cu); returns 1 when not applicable
int launch_wfa_long(const KernelArgs &a, Scratch *s, void *stream, int *launches);

// Launchers (aim_wfa.cu / aim_dp.cu).  Enqueue on `stream`; return AIM_OK or AIM_ERR_*;
// *launches is incremented by the number of kernels enqueued.
int launch_wfa(const KernelArgs &a, Scratch *s, void *stream, int *launches);
// lockstep short-read kernel; returns 1 when the configuration must go to launch_wfa's warp-per-pair kernel
int launch_wfa_sub(const KernelArgs &a, Scratch *s, void *stream, int *launches);
int launch_dp(const KernelArgs &a, Scratch *s, void *stream, int *launches);
// CIGAR text rows (a.cigars, a.cigar_pitch) from the op rows the alignment kernels wrote (aim_wfa_sub.cu)
int launch_cigar_rows(const KernelArgs &a, void *stream, int *launches);
// GenASM-DC / GenASM-filter (aim_genasm.cu)
int launch_genasm(const KernelArgs &a, Scratch *s, void *stream, int *launches);
// register-strip / shared-memory-row kernels (aim_dp_fast.cu); returns 1 when launch_dp's literal kernel must serve the batch
int launch_dp_fast(const KernelArgs &a, Scratch *s, void *stream, int *launches);

// ---- file pipeline (aim_file.cu kernels, aim_filepipe.cu host side) ----
// The dispatcher's per-device state for code outside aim_dispatch.cu: *scratch is the device's Scratch, *mu the mutex
// that serialises launch sequences on it (lock it around launch_algo + the stream work that follows).
int device_state(int device, Scratch **scratch, void **mu /* std::mutex* */);
// The alignment kernels of a.p.algo on `stream` (scratch hand-over included).
int launch_algo(const KernelArgs &a, Scratch *s, void *stream, int *launches);
size_t file_parse_scratch_bytes(size_t chunk_bytes);
size_t file_format_scratch_bytes(uint32_t max_pairs);
int launch_file_parse(const char *d_buf, size_t nbytes, uint32_t lines, int unterminated, int read_size, uint32_t *d_tiles, uint32_t *d_nl_pos,
                      uint32_t *d_counters, int32_t *d_plen, int32_t *d_tlen, char *d_pat, char *d_txt, void *stream, int *launches);
int launch_file_format(const aim_result *d_res, const char *d_ops, int read_size, int backtrace, int mode, uint32_t m, uint32_t *d_lens, uint32_t *d_offs,
                       uint32_t *d_tiles, uint32_t *d_counters, char *d_out, size_t out_cap, void *stream, int *launches);
// second half of launch_file_format alone (after the output buffer was grown)
int launch_file_format_write(const aim_result *d_res, const char *d_ops, int read_size, int backtrace, int mode, uint32_t m, const uint32_t *d_offs,
                             const uint32_t *d_counters, char *d_out, size_t out_cap, void *stream, int *launches);
bool params_valid_for_file(const aim_params *p);
aim_params params_normalized(const aim_params *p);  // GenASM-DC implies the op rows, the filter has none
size_t count_newlines(const char *p, size_t n);  // aim_host.cpp (AVX2 when available)
void file_pipeline_shutdown();                   // frees the cached chunk slots of aim_align_file (aim_shutdown)


// ---- op rows as run rows across PCIe (kernel: aim_file.cu; expansion and its thread pool: aim_host.cpp) ----
int32_t op_runs_pitch(int32_t read_size);  // bytes per pair, 0 = not served (the rows are downloaded as they are)
int launch_op_runs(const char *d_ops, int read_size, uint32_t m, unsigned char *d_runs, int pitch, void *stream, int *launches);
// Rebuilds the 2 * read_size op rows of m pairs at `ops` from their run rows on the host pool's threads; *overflow receives
// the pairs whose run row carries the "did not fit" mark (their rows are left untouched).
void expand_op_runs(const unsigned char *runs, int pitch, uint32_t m, int read_size, char *ops, std::vector<uint32_t> *overflow);
// GenASM-DC: the same for the CIGAR strings its op rows hold - the first str_rows_pitch() bytes of every row cross PCIe, the host
// copies each string (with its NUL) to the head of the caller's row; *overflow: strings that do not end inside their piece.
int32_t str_rows_pitch(int32_t read_size, int32_t max_score);  // 0 = not worth it
int32_t op_rows_download_bytes(const aim_params &p);           // run row / string head bytes of this parameter set (0 = whole rows)
int launch_str_rows(const char *d_ops, int read_size, uint32_t m, unsigned char *d_rows, int pitch, void *stream, int *launches);
void expand_str_rows(const unsigned char *rows, int pitch, uint32_t m, int read_size, char *ops, std::vector<uint32_t> *overflow);
void host_pool_shutdown();
void host_pool_want(int helpers);  // at least this many helper threads from now on (one process driving several GPUs)

}  // namespace aim

#endif
