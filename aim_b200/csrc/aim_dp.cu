// Full-table NW (linear gap) and SWG (gap-affine) for sm_100a: one pair per thread.
//
// Replaces NW/DPU-{WRAM,MRAM}/dpu/nw.c (nw_compute :109-153, nw_traceback :67-107) and
// SWG/DPU-MRAM/dpu/swg.c (swg_compute :151-217, swg_traceback :66-148).  The reference fills a
// FLAT int16 array indexed num_cols*h + v with num_cols = text_len+1 while v runs to pattern_len,
// so for pattern_len > text_len row h's tail (v >= num_cols) aliases row h+1's head and the
// traceback reads the final, overwritten state.  That behaviour is reproduced exactly without
// materialising the table, from two facts ("flat-array semantics"):
//   (1) at fill time, cell (h, v>=nc) reads "up"/"diag" from the CURRENT row's head
//       (R_h[v-nc], R_h[v-1-nc]) and row h>=2 starts from R_h[0] = R_{h-1}[nc];
//   (2) the final content of flat word i was produced by its LAST writer (row r = min(tl,(i-1)/nc),
//       column i - nc*r), and every neighbour that writer compared against already held its own
//       final value; so the traceback predicates evaluated at fill time by the last writer are the
//       ones the reference evaluates on the final table.
// Each thread therefore keeps one row of cells (global scratch, lane-interleaved so every warp
// access is one 128-byte line) and one predicate byte per cell, and walks the traceback through
// the last-writer mapping.  int16 truncation happens at the same assignments as in the reference.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "aim_internal.h"

namespace aim {

namespace {

struct DpK {
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    uint32_t n, idx_base;
    int match, x, o, e, max_score, read_size, backtrace;
    uint32_t *rows;        // [warp][v][lane]
    uint32_t *flags;       // [warp][(h-1)*wpr + (v-1)/4][lane], one predicate byte per cell
    size_t row_stride;     // words per warp
    size_t flag_stride;    // words per warp
    uint32_t wpr;          // flag words per row
};

// cell truncation: int16 (NW/*, SWG/DPU-MRAM: common.h:91-99) or int8 (SWG/DPU-WRAM with MAX_SCORE < 127: common.h:71-79)
template <int BITS>
__device__ __forceinline__ int trunc_cell(int v) { return BITS == 8 ? (int)(signed char)v : (int)(short)v; }

// NW predicate byte: 3 = 'D' (left + GAP_D), 2 = 'I' (up + GAP_I), 1 = 'X', 0 = 'M'; tested in the
// reference's order (nw.c:78-94).
// SWG predicate byte: bits 0-2 = M-layer decision in the reference's order (swg.c:106-133):
// 0 -> D layer, 1 -> I layer, 2 'M', 3 'X', 4 dead end; bit 3 = D opened here (swg.c:88);
// bit 4 = I opened here (swg.c:97).
template <int ALGO, int BITS>
__global__ void __launch_bounds__(128) dp_kernel(const DpK K)
{
    auto s16 = [](int v) { return trunc_cell<BITS>(v); };  // every "int16" assignment below is a cell-type assignment
    const int lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nthreads = gridDim.x * blockDim.x;
    uint32_t *row = K.rows + (size_t)gwarp * K.row_stride + lane;
    uint32_t *flg = K.backtrace ? K.flags + (size_t)gwarp * K.flag_stride + lane : nullptr;
    const int RS = K.read_size;
    const int X = K.x, O = K.o, E = K.e, MT = K.match, MS = K.max_score;
    const int GI = K.o, GD = K.o;  // NW: single linear gap

    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < K.n; i += nthreads) {
        const int pl = min(max(K.plen[i], 0), RS), tl = min(max(K.tlen[i], 0), RS);
        const char *gp = K.patterns + (size_t)i * RS;
        const char *gt = K.texts + (size_t)i * RS;
        const int nc = tl + 1;
        const bool alias = pl >= nc;
        int score = 0;

        // row 0 (nw.c:119-124 / swg.c:167-175).  Word = M | I << 16 for SWG, the cell for NW.
        if (ALGO == AIM_ALGO_NW) {
            for (int v = 0; v <= pl; ++v) row[(size_t)v * 32] = (uint32_t)(uint16_t)s16(v * GD);
        } else {
            row[0] = (uint32_t)(uint16_t)0 | ((uint32_t)(uint16_t)s16(MS) << 16);
            for (int v = 1; v <= pl; ++v)
                row[(size_t)v * 32] = (uint32_t)(uint16_t)s16(O + v * E) | ((uint32_t)(uint16_t)s16(MS) << 16);
        }
        int tailM = 0, tailI = 0, tailD = 0;  // cell (h-1, nc), becomes column 0 of row h when aliased

        for (int h = 1; h <= tl; ++h) {
            const unsigned char tc = (unsigned char)gt[h - 1];
            int c0M, c0I, c0D;
            if (alias && h >= 2) { c0M = tailM; c0I = tailI; c0D = tailD; }
            else if (ALGO == AIM_ALGO_NW) { c0M = s16(h * GI); c0I = c0D = 0; }
            else { c0D = s16(MS); c0I = s16(O + h * E); c0M = c0I; }
            uint32_t w0 = row[0];
            int diagM = s16((int)(w0 & 0xffffu));
            row[0] = (uint32_t)(uint16_t)c0M | ((uint32_t)(uint16_t)c0I << 16);
            int leftM = c0M, leftD = c0D;
            const int nwords = (pl + 3) >> 2;
            for (int w = 0; w < nwords; ++w) {
                const uint32_t pw = *reinterpret_cast<const uint32_t *>(gp + 4 * w);
                uint32_t fword = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int v = 4 * w + j + 1;
                    if (v > pl) break;
                    const uint32_t old = row[(size_t)v * 32];
                    uint32_t upw = old;
                    int dgM = diagM;
                    if (v >= nc) {  // aliased tail: "previous row" words are the current row's head
                        upw = row[(size_t)(v - nc) * 32];
                        if (v - 1 >= nc) dgM = s16((int)(row[(size_t)(v - 1 - nc) * 32] & 0xffffu));
                    }
                    const int upM = s16((int)(upw & 0xffffu));
                    const bool eq = ((pw >> (8 * j)) & 0xffu) == tc;
                    int valM, fl;
                    if (ALGO == AIM_ALGO_NW) {
                        const int del = s16(leftM + GD), ins = s16(upM + GI), mm = s16(dgM + (eq ? 0 : X));
                        valM = min(mm, min(ins, del));
                        fl = (valM == leftM + GD) ? 3 : (valM == upM + GI) ? 2 : (valM == dgM + X) ? 1 : 0;
                        row[(size_t)v * 32] = (uint32_t)(uint16_t)valM;
                        if (v == nc) tailM = valM;
                    } else {
                        const int upI = s16((int)(upw >> 16));
                        const int del = min(s16(leftM + O + E), s16(leftD + E));
                        const int ins = min(s16(upM + O + E), s16(upI + E));
                        const int mm = s16(dgM + (eq ? MT : X));
                        valM = min(mm, min(ins, del));
                        fl = (valM == del) ? 0 : (valM == ins) ? 1 : (valM == dgM + MT) ? 2 : (valM == dgM + X) ? 3 : 4;
                        fl |= (del == leftM + O + E) ? 8 : 0;
                        fl |= (ins == upM + O + E) ? 16 : 0;
                        row[(size_t)v * 32] = (uint32_t)(uint16_t)valM | ((uint32_t)(uint16_t)ins << 16);
                        leftD = del;
                        if (v == nc) { tailM = valM; tailI = ins; tailD = del; }
                    }
                    fword |= (uint32_t)fl << (8 * j);
                    diagM = s16((int)(old & 0xffffu));
                    leftM = valM;
                    score = valM;
                }
                if (flg) flg[((size_t)(h - 1) * K.wpr + w) * 32] = fword;
            }
        }

        int begin_offset = pl + tl - 1;
        int status = AIM_STATUS_OK;
        if (K.backtrace) {
            char *ops = K.ops + (size_t)i * 2 * RS;  // pre-filled with 'M' by the launcher
            int b = pl + tl - 1;
            int h = tl, v = pl;
            int layer = 0;  // SWG: 0 M, 1 I, 2 D
            while (h > 0 && v > 0) {
                const int fi = nc * h + v;
                const int r = min(tl, (fi - 1) / nc);
                const int c = fi - nc * r;
                const uint32_t fw = flg[((size_t)(r - 1) * K.wpr + ((c - 1) >> 2)) * 32];
                const int fl = (int)((fw >> (8 * ((c - 1) & 3))) & 0xffu);
                if (ALGO == AIM_ALGO_NW) {
                    if (fl == 3) { ops[b--] = 'D'; --v; }
                    else if (fl == 2) { ops[b--] = 'I'; --h; }
                    else { if (fl == 1) ops[b] = 'X'; --b; --h; --v; }
                } else {
                    if (b < 0) { status = AIM_STATUS_BACKTRACE; break; }
                    if (layer == 2) { ops[b--] = 'D'; if (fl & 8) layer = 0; --v; }
                    else if (layer == 1) { ops[b--] = 'I'; if (fl & 16) layer = 0; --h; }
                    else {
                        const int m = fl & 7;
                        if (m == 0) layer = 2;
                        else if (m == 1) layer = 1;
                        else if (m == 2) { --b; --h; --v; }
                        else if (m == 3) { ops[b--] = 'X'; --h; --v; }
                        else { status = AIM_STATUS_BACKTRACE; break; }
                    }
                }
            }
            if (status == AIM_STATUS_OK) {
                while (h > 0) { ops[b--] = 'I'; --h; }
                while (v > 0) { ops[b--] = 'D'; --v; }
                begin_offset = b + 1;
            }
        }
        aim_result res;
        res.max_operations = pl + tl;
        res.begin_offset = begin_offset;
        res.end_offset = pl + tl;
        res.score = (pl > 0 && tl > 0) ? score : 0;
        res.status = status;
        res.idx = K.idx_base + i;
        K.results[i] = res;
    }
}

}  // namespace

int launch_dp(const KernelArgs &a, Scratch *sc, void *stream_v, int *launches)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    const aim_params &p = a.p;
    if (a.n == 0) return AIM_OK;
    // SWG/DPU-WRAM semantics (variant 1): int8 cells when MAX_SCORE < 127 - the literal kernel with 8-bit truncation serves it
    // (the 32-bit fast kernels are exact only where no truncation can occur)
    const bool w8 = p.algo == AIM_ALGO_SWG && p.variant == 1 && p.max_score < 127;
    if (!w8) {
        const int rc = launch_dp_fast(a, sc, stream_v, launches);
        if (rc != 1) return rc;
    }
    DpK K{};
    K.plen = a.plen; K.tlen = a.tlen; K.patterns = a.patterns; K.texts = a.texts;
    K.results = a.results; K.ops = a.ops; K.n = a.n; K.idx_base = a.idx_base;
    K.match = p.match; K.x = p.mismatch; K.o = p.gap_open; K.e = p.gap_ext;
    K.max_score = p.max_score; K.read_size = p.read_size; K.backtrace = p.backtrace;
    const uint32_t RS = (uint32_t)p.read_size;
    K.wpr = (RS + 3) / 4;
    K.row_stride = (size_t)(RS + 1) * 32;
    K.flag_stride = p.backtrace ? (size_t)RS * K.wpr * 32 : 0;
    const size_t per_warp = (K.row_stride + K.flag_stride) * 4;
    const int block = 128;
    // occupancy-driven grid: enough resident warps to cover latency, bounded by the scratch budget
    uint64_t want_warps = (uint64_t)sc->sm_count * 16;
    want_warps = std::min<uint64_t>(want_warps, ((uint64_t)a.n + 31) / 32);
    const uint64_t budget = 24ull << 30;
    want_warps = std::max<uint64_t>(1, std::min<uint64_t>(want_warps, budget / per_warp));
    const int warps_per_block = block / 32;
    int grid = (int)((want_warps + warps_per_block - 1) / warps_per_block);
    const size_t total_warps = (size_t)grid * warps_per_block;
    int rc = scratch_reserve(sc, total_warps * per_warp);
    if (rc != AIM_OK) return rc;
    K.rows = reinterpret_cast<uint32_t *>(sc->buf);
    K.flags = K.rows + total_warps * K.row_stride;
    cudaError_t err = cudaSuccess;
    if (p.backtrace) {
        err = cudaMemsetAsync(a.ops, 'M', (size_t)a.n * 2 * RS, stream);
    }
    if (err == cudaSuccess) {
        if (p.algo == AIM_ALGO_NW) dp_kernel<AIM_ALGO_NW, 16><<<grid, block, 0, stream>>>(K);
        else if (w8) dp_kernel<AIM_ALGO_SWG, 8><<<grid, block, 0, stream>>>(K);
        else dp_kernel<AIM_ALGO_SWG, 16><<<grid, block, 0, stream>>>(K);
        err = cudaGetLastError();
    }
    if (err != cudaSuccess) { set_error(std::string("dp launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    if (launches) *launches += 1;
    return AIM_OK;
}

}  // namespace aim
