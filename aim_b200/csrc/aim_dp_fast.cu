// Fast NW (linear gap) / SWG (gap-affine) full-table DP for sm_100a.
//
// Same flat-array semantics as aim_dp.cu (reference: NW/DPU-WRAM/dpu/nw.c:67-153,
// SWG/DPU-MRAM/dpu/swg.c:66-217), restructured around two observations:
//
//   (1) a pair with pattern_len <= text_len never aliases rows of the flat table
//       (num_cols = text_len+1 > pattern_len), so its fill is the ordinary full DP and may be
//       evaluated in ANY dependency-respecting order.  dp_strip_kernel walks it in column strips of
//       16 cells held in registers (previous row's M and I per column), one pair per thread; only
//       the strip's right boundary column (M, D per row) goes through memory, once per 16 cells.
//   (2) a pair with pattern_len > text_len aliases: cell (h,0) IS cell (h-1,num_cols) and the tail
//       cells (h, v >= num_cols) read the current row's head as their "previous row".  The fill is
//       then one serial chain in the reference's write order; dp_row_kernel keeps that order with
//       the row in shared memory (conflict-free [column][thread] layout), one pair per thread.
//
// A classification pass splits the batch into the two index lists.  Both kernels evaluate the
// reference's traceback predicates at fill time (DESIGN.md "flat-array semantics": the last writer
// of a flat word sees the same neighbours the reference's traceback reads from the final table) and
// keep 4 bits (SWG) / 2 bits (NW) per cell in HBM:
//     p    = (del <= ins)            q   = (min(del,ins) <= diag + sub)
//     opD  = D opened at this cell   opI = I opened at this cell          (SWG only)
// from which the reference's test order (D, I, then diag+MATCH / diag+MISMATCH: swg.c:106-133,
// nw.c:78-94) is reproduced exactly: q&&p -> D, q&&!p -> I, !q -> 'M' if the writer's two bases are
// equal else 'X' (when neither gap wins the cell value IS diag+sub, so the reference's value tests
// reduce to the base comparison).
//
// Both kernels compute in 32-bit registers.  They are used only when no int16 truncation can occur
// ((plen+tlen+2) * max penalty + MAX_SCORE < 32767, checked by the launcher); otherwise, and for
// rows that do not fit shared memory, aim_dp.cu's literal kernel serves the batch.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "aim_internal.h"

namespace aim {

namespace {

constexpr int KS = 16;  // columns per register strip

struct FastK {
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    uint32_t n, idx_base;
    int match, x, o, e, max_score, read_size, backtrace;
    const uint32_t *list;   // pair indices served by this kernel
    const uint32_t *count;  // how many (device counter written by classify_kernel)
    int list_step;          // +1: list grows upwards from list[0]; -1: downwards from list[0]
    uint32_t *bound;        // strip kernel: boundary column, [row][thread]
    uint32_t *flags;        // predicate bits
    uint32_t wpr;           // row kernel: flag records per row
    uint2 *tflags;          // scan kernel: the tail cells' predicate nibbles, one 64-bit word per (row, pair)
    uint32_t lbase, llimit; // scan kernels: the list entries [lbase, min(llimit, *count)) of this launch
    int neg1;               // -1, kept opaque to the compiler (see dp_cell)
};

// ---- classification: non-aliased pairs to the front of `list`, aliased ones that fit the register-row
// kernel (tlen >= reg_cols, plen - tlen <= 16) to its back, the other aliased ones to `list2` ----
// (scan_cols > 0: the back of `list` takes the aliased pairs dp_scan_kernel serves instead: tlen <= scan_cols, plen - tlen <= min(scan_d, tlen): the tail reads head columns only)
__global__ void classify_kernel(const int32_t *plen, const int32_t *tlen, uint32_t n, int RS, int reg_cols, int scan_cols, int scan_d,
                                uint32_t *list, uint32_t *list2, uint32_t *counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool valid = i < n, alias = false, regrow = false;
    if (valid) {
        const int pl = min(max(plen[i], 0), RS), tl = min(max(tlen[i], 0), RS);
        alias = pl > tl;
        regrow = scan_cols > 0 ? (alias && tl <= scan_cols && pl - tl <= min(scan_d, tl))
                               : (alias && reg_cols > 0 && tl >= reg_cols && pl - tl <= 16);
    }
    const uint32_t m0 = __ballot_sync(0xffffffffu, valid && !alias), m1 = __ballot_sync(0xffffffffu, regrow),
                   m2 = __ballot_sync(0xffffffffu, valid && alias && !regrow);
    uint32_t b0 = 0, b1 = 0, b2 = 0;
    if (lane == 0) {
        if (m0) b0 = atomicAdd(&counters[0], (uint32_t)__popc(m0));
        if (m1) b1 = atomicAdd(&counters[1], (uint32_t)__popc(m1));
        if (m2) b2 = atomicAdd(&counters[2], (uint32_t)__popc(m2));
    }
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    b1 = __shfl_sync(0xffffffffu, b1, 0);
    b2 = __shfl_sync(0xffffffffu, b2, 0);
    const uint32_t below = (1u << lane) - 1u;
    if (valid && !alias) list[b0 + (uint32_t)__popc(m0 & below)] = i;
    if (regrow) list[n - 1 - (b1 + (uint32_t)__popc(m1 & below))] = i;
    if (valid && alias && !regrow) list2[b2 + (uint32_t)__popc(m2 & below)] = i;
}

// ---- one DP cell in registers ----
// Returns the new M; updates upI (-> ins) and leftD (-> del); the four differences carry the
// traceback predicates in their SIGN bits (set = predicate false):
//   dI = (upI+E) - (upM+O+E)     sign set  <=>  I was NOT opened here   (swg.c:97)
//   dD = (leftD+E) - (leftM+O+E) sign set  <=>  D was NOT opened here   (swg.c:88)
//   dP = ins - del               sign set  <=>  !(del <= ins)
//   dQ = mm - min(del,ins)       sign set  <=>  !(min(del,ins) <= mm)
//
// Adds/subs are left to ptxas, which splits them between the alu pipe (VIADD) and the fma pipe
// (IMAD.IADD): min / funnel-shift / logic only run on the alu pipe, and forcing more adds onto IMAD
// was measured slower (integer IMAD only issues on the heavy half of the fma pipe).
template <int ALGO>
__device__ __forceinline__ int dp_cell(int upM, int &upI, int leftM, int &leftD, int mm, int OE, int E, int neg1, int &dI, int &dD, int &dP,
                                       int &dQ)
{
    int ins, del;
    if (ALGO == AIM_ALGO_NW) {
        ins = upM + OE;    // GAP_I
        del = leftM + OE;  // GAP_D
        dI = dD = 0;
    } else {
        const int i1 = upM + OE, i2 = upI + E;
        ins = min(i1, i2);
        dI = i2 - i1;
        const int d1 = leftM + OE, d2 = leftD + E;
        del = __viaddmin_s32(leftM, OE, d2);  // VIADDMNMX: one instruction from leftM to del
        dD = d2 - d1;
        upI = ins;
        leftD = del;
    }
    // The serial chain of a row runs through leftM: with the DPX forms it is two instructions per cell (leftM -> del -> M:
    // VIADDMNMX, VIMNMX3) instead of four (add, min, min, min); the differences that carry the predicates hang off it.
    const int m1 = min(del, ins);
    dP = ins - del;
    dQ = mm - m1;
    return __vimin3_s32(del, ins, mm);
}

// append the sign bit of d to the accumulator (one SHF.L.W)
__device__ __forceinline__ uint32_t push_sign(uint32_t acc, int d) { return __funnelshift_l((uint32_t)d, acc, 1); }

// Flag record of 16 consecutive cells: word0 = notP | notQ << 16, word1 (SWG) = notOpD | notOpI << 16,
// cell j of the record at bit 15-j of each half.
__device__ __forceinline__ void decode_flags(const uint32_t *rec, int j, bool swg, bool &p, bool &q, bool &opD, bool &opI)
{
    const uint32_t w0 = rec[0];
    p = !((w0 >> (15 - j)) & 1u);
    q = !((w0 >> (31 - j)) & 1u);
    opD = opI = false;
    if (swg) {
        const uint32_t w1 = rec[1];
        opD = !((w1 >> (15 - j)) & 1u);
        opI = !((w1 >> (31 - j)) & 1u);
    }
}

// ================= non-aliased pairs: register strips =================
template <int ALGO, bool MT0>
__global__ void __launch_bounds__(128) dp_strip_kernel(const FastK K)
{
    constexpr bool SWG = (ALGO == AIM_ALGO_SWG);
    constexpr int FW = SWG ? 2 : 1;  // flag words per 16-cell record
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nth = gridDim.x * blockDim.x;
    const uint32_t count = *K.count;
    const int RS = K.read_size;
    const int X = K.x, E = K.e, MT = MT0 ? 0 : K.match, MS = K.max_score;
    const int OE = SWG ? K.o + K.e : K.o;  // NW: the single linear gap
    const int O = K.o, neg1 = K.neg1;
    uint32_t sel[4];  // IDP.4A selectors: (X - MT) in byte k
#pragma unroll
    for (int k = 0; k < 4; ++k) sel[k] = (uint32_t)(X - MT) << (8 * k);
    uint32_t *bnd = K.bound + tid;
    uint32_t *flg = K.flags + (size_t)tid * FW;
    const size_t fstep = (size_t)nth * FW;  // words between consecutive (strip,row) records

    for (uint32_t li = tid; li < count; li += nth) {
        const uint32_t i = K.list[li];
        const int pl = min(max(K.plen[i], 0), RS), tl = min(max(K.tlen[i], 0), RS);
        const char *gp = K.patterns + (size_t)i * RS;
        const char *gt = K.texts + (size_t)i * RS;
        const int nstrips = (pl + KS - 1) / KS;
        int score = 0;

        for (int s = 0; s < nstrips; ++s) {
            const int v0 = s * KS;
            uint32_t pc[KS / 4];
#pragma unroll
            for (int w = 0; w < KS / 4; ++w) pc[w] = (v0 + 4 * w < RS) ? __ldg(reinterpret_cast<const uint32_t *>(gp + v0) + w) : 0u;
            int upM[KS], upI[KS];
#pragma unroll
            for (int j = 0; j < KS; ++j) {  // row 0 (nw.c:119-124 / swg.c:167-175)
                upM[j] = SWG ? O + (v0 + 1 + j) * E : (v0 + 1 + j) * OE;
                upI[j] = MS;
            }
            // M(h-1, v0): the diagonal neighbour of the strip's first cell
            int dg0 = SWG ? (v0 == 0 ? 0 : O + v0 * E) : v0 * OE;
            const bool first = (s == 0), last = (s == nstrips - 1);
            uint32_t tw = 0;
            uint32_t *frec = flg + (size_t)s * RS * fstep;
            for (int h = 1; h <= tl; ++h) {
                if (((h - 1) & 3) == 0) tw = __ldg(reinterpret_cast<const uint32_t *>(gt) + ((h - 1) >> 2));
                const uint32_t tc4 = (tw & 0xffu) * 0x01010101u;
                tw >>= 8;
                uint32_t ne[KS / 4];  // 1 in every byte whose pattern base differs from this row's text base
#pragma unroll
                for (int w = 0; w < KS / 4; ++w) ne[w] = __vsetne4(pc[w] ^ tc4, 0u);
                int leftM, leftD;
                if (first) {  // column 0 (nw.c:114-118 / swg.c:158-166)
                    leftM = SWG ? O + h * E : h * OE;
                    leftD = MS;
                } else {
                    const uint32_t b = bnd[(size_t)(h - 1) * nth];
                    leftM = (int)(short)(b & 0xffffu);
                    leftD = (int)b >> 16;
                }
                int dg = dg0;
                dg0 = leftM;
                uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                    const int um = upM[j];
                    const int mm = (int)__dp4a(ne[j >> 2], sel[j & 3], (uint32_t)(MT0 ? dg : dg + MT));
                    int dI, dD, dP, dQ;
                    const int m = dp_cell<ALGO>(um, upI[j], leftM, leftD, mm, OE, E, neg1, dI, dD, dP, dQ);
                    aP = push_sign(aP, dP);
                    aQ = push_sign(aQ, dQ);
                    if (SWG) { aD = push_sign(aD, dD); aI = push_sign(aI, dI); }
                    dg = um;
                    upM[j] = m;
                    leftM = m;
                }
                if (!last) bnd[(size_t)(h - 1) * nth] = ((uint32_t)leftM & 0xffffu) | ((uint32_t)leftD << 16);
                if (K.backtrace) {
                    uint32_t *d = frec + (size_t)(h - 1) * fstep;
                    if (SWG) *reinterpret_cast<uint2 *>(d) = make_uint2(aP | (aQ << 16), aD | (aI << 16));
                    else d[0] = aP | (aQ << 16);
                }
            }
            if (last) {
                const int js = pl - 1 - v0;
#pragma unroll
                for (int j = 0; j < KS; ++j) if (j == js) score = upM[j];
            }
        }
        if (pl == 0 || tl == 0) score = 0;

        int begin_offset = pl + tl - 1;
        int status = AIM_STATUS_OK;
        if (K.backtrace) {
            char *ops = K.ops + (size_t)i * 2 * RS;  // pre-filled with 'M' by the launcher
            int b = pl + tl - 1;
            int h = tl, v = pl;
            int layer = 0;  // SWG: 0 M, 1 I, 2 D
            while (h > 0 && v > 0) {
                const int s = (v - 1) / KS, j = (v - 1) % KS;
                bool p, q, opD, opI;
                decode_flags(flg + ((size_t)s * RS + (h - 1)) * fstep, j, SWG, p, q, opD, opI);
                if (!SWG) {
                    if (q) {
                        if (p) { ops[b--] = 'D'; --v; }
                        else { ops[b--] = 'I'; --h; }
                    } else {
                        if (gp[v - 1] != gt[h - 1]) ops[b] = 'X';
                        --b; --h; --v;
                    }
                } else {
                    if (b < 0) { status = AIM_STATUS_BACKTRACE; break; }
                    if (layer == 2) { ops[b--] = 'D'; if (opD) layer = 0; --v; }
                    else if (layer == 1) { ops[b--] = 'I'; if (opI) layer = 0; --h; }
                    else if (q) layer = p ? 2 : 1;
                    else {
                        if (gp[v - 1] != gt[h - 1]) ops[b] = 'X';
                        --b; --h; --v;
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
        res.score = score;
        res.status = status;
        res.idx = K.idx_base + i;
        K.results[i] = res;
    }
}

#include "aim_dp_pack2.cuh"

// ================= aliased pairs: serial row-major fill, row in shared memory =================
// Shared memory per thread: (RS+1) row words [column][lane] (NW: M; SWG: M | I << 16).  Pattern bases
// come through L1 (16 bytes per record, fetched one record ahead).  Flags: one record per 16 columns.
// One warp per block: the [column][lane] stride is a compile-time 128 bytes, so every access inside a
// 16-cell record is base + immediate, and the next record's 16 "previous row" words are fetched into
// registers before the current record's dependent chain runs (each pair is one serial chain; with a
// handful of warps per SM the shared-memory latency must be hidden inside the thread).
// PSMEM: the pattern bases are staged in shared memory after the row (RS/4 words per lane).  Worth it
// when many warps still fit (short reads: L1 is too small for their patterns); for long rows the
// launcher prefers one more resident warp and reads the pattern through L1 one record ahead.
// R > 0: columns 1..R of the row live in REGISTERS (R/16 fully unrolled records); shared memory keeps column 0, a
// write-through copy of columns 1..15 (what the aliased tail reads back: plen - tlen <= 16 is a precondition of this
// variant, as is tlen >= R) and columns R+1..: the register file is idle in this kernel while shared memory bounds the
// resident warps, so moving part of the row there raises the occupancy by half.
constexpr uint32_t RT = 32;
template <int ALGO, bool PSMEM, int R>
__global__ void __launch_bounds__(RT) __maxnreg__(R > 0 ? 224 : 128) dp_row_kernel(const FastK K)
{
    constexpr bool SWG = (ALGO == AIM_ALGO_SWG);
    constexpr int FW = SWG ? 2 : 1;
    extern __shared__ uint32_t smem[];
    constexpr uint32_t T = RT;
    const uint32_t tid = blockIdx.x * T + threadIdx.x;
    const uint32_t nth = gridDim.x * T;
    const uint32_t count = *K.count;
    const int RS = K.read_size;
    const int X = K.x, E = K.e, MT = SWG ? K.match : 0, MS = K.max_score;
    const int OE = SWG ? K.o + K.e : K.o;
    const int O = K.o, neg1 = K.neg1;
    uint32_t sel[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) sel[k] = (uint32_t)(X - MT) << (8 * k);
    // shared-memory word of column v: v for v <= 15 (R > 0) or any v (R == 0); v - R + 15 for v > R (R > 0)
    uint32_t *row0 = smem + threadIdx.x;                         // row0[v * T], v <= 15 when R > 0
    uint32_t *row = R > 0 ? row0 + (15 - R) * (int)T : row0;     // row[v * T],  v > R  when R > 0
    const int smem_cols = R > 0 ? RS + 16 - R : RS + 1;
    uint32_t *pat = smem + (size_t)smem_cols * T + threadIdx.x;  // pat[w * T] (PSMEM only)
    uint32_t *flg = K.flags + (size_t)tid * FW;
    const size_t fstep = (size_t)nth * FW;
    const uint32_t rpr = K.wpr;  // records per row

    auto unpackM = [](uint32_t w) -> int { return SWG ? (int)(short)(w & 0xffffu) : (int)w; };
    auto unpackI = [](uint32_t w) -> int { return (int)w >> 16; };
    auto pack = [](int m, int i) -> uint32_t { return SWG ? (((uint32_t)m & 0xffffu) | ((uint32_t)i << 16)) : (uint32_t)m; };

    for (uint32_t li = tid; li < count; li += nth) {
        const uint32_t i = K.list[(int64_t)K.list_step * (int64_t)li];
        const int pl = min(max(K.plen[i], 0), RS), tl = min(max(K.tlen[i], 0), RS);
        const char *gp = K.patterns + (size_t)i * RS;
        const char *gt = K.texts + (size_t)i * RS;
        const int nc = tl + 1;
        const bool alias = pl >= nc;
        int score = 0;

        if (PSMEM) for (int w = 0; w * 4 < pl; ++w) pat[(size_t)w * T] = __ldg(reinterpret_cast<const uint32_t *>(gp) + w);
        // row 0 (nw.c:119-124 / swg.c:167-175)
        row0[0] = pack(0, MS);
        uint32_t reg[R > 0 ? R : 1];
        if (R > 0) {
#pragma unroll
            for (int j = 0; j < R; ++j) reg[j] = pack(SWG ? O + (j + 1) * E : (j + 1) * OE, MS);
            for (int v = 1; v <= 15; ++v) row0[(size_t)v * T] = pack(SWG ? O + v * E : v * OE, MS);
        }
        for (int v = R + 1; v <= pl; ++v) row[(size_t)v * T] = pack(SWG ? O + v * E : v * OE, MS);
        int tailM = 0, tailI = 0, tailD = 0;  // cell (h-1, nc): column 0 of row h when aliased
        const int headend = min(pl, nc - 1);  // columns 1..headend read the true previous row

        uint32_t tw = 0;
        for (int h = 1; h <= tl; ++h) {
            if (((h - 1) & 3) == 0) tw = __ldg(reinterpret_cast<const uint32_t *>(gt) + ((h - 1) >> 2));
            const uint32_t tc4 = (tw & 0xffu) * 0x01010101u;
            tw >>= 8;
            int leftM, leftD, c0I;
            if (alias && h >= 2) { leftM = tailM; c0I = tailI; leftD = tailD; }
            else if (!SWG) { leftM = h * OE; c0I = 0; leftD = 0; }
            else { leftD = MS; c0I = O + h * E; leftM = c0I; }
            int dg = unpackM(row0[0]);
            row0[0] = pack(leftM, c0I);
            uint32_t *frow = flg + (size_t)(h - 1) * rpr * fstep;

            // ---- head, full 16-cell records ----
            int v = R + 1;
            uint32_t cur[16], nxt[16];
            uint2 pcur[2], pnxt[2];  // the record's 16 pattern bases
            if (headend >= R + 16) {  // first shared-memory record: fetch early
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = row[(size_t)(R + 1 + j) * T];
                if (!PSMEM) {
                    pcur[0] = __ldg(reinterpret_cast<const uint2 *>(gp + R));
                    pcur[1] = __ldg(reinterpret_cast<const uint2 *>(gp + R) + 1);
                }
            }
            if (R > 0) {  // columns 1..R from registers
#pragma unroll
                for (int rec = 0; rec < R / 16; ++rec) {
                    uint32_t ne[4];
                    if (PSMEM) {
#pragma unroll
                        for (int w = 0; w < 4; ++w) ne[w] = __vsetne4(pat[(size_t)(rec * 4 + w) * T] ^ tc4, 0u);
                    } else {
                        const uint2 pa = __ldg(reinterpret_cast<const uint2 *>(gp + rec * 16)), pb = __ldg(reinterpret_cast<const uint2 *>(gp + rec * 16) + 1);
                        ne[0] = __vsetne4(pa.x ^ tc4, 0u); ne[1] = __vsetne4(pa.y ^ tc4, 0u);
                        ne[2] = __vsetne4(pb.x ^ tc4, 0u); ne[3] = __vsetne4(pb.y ^ tc4, 0u);
                    }
                    uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const uint32_t old = reg[rec * 16 + j];
                        const int upM = unpackM(old);
                        int upI = SWG ? unpackI(old) : 0;
                        const int mm = (int)__dp4a(ne[j >> 2], sel[j & 3], (uint32_t)(dg + MT));
                        int dI, dD, dP, dQ;
                        const int m = dp_cell<ALGO>(upM, upI, leftM, leftD, mm, OE, E, neg1, dI, dD, dP, dQ);
                        reg[rec * 16 + j] = pack(m, upI);
                        if (rec == 0 && j < 15) row0[(size_t)(1 + j) * T] = reg[j];  // what the aliased tail reads back
                        aP = push_sign(aP, dP);
                        aQ = push_sign(aQ, dQ);
                        if (SWG) { aD = push_sign(aD, dD); aI = push_sign(aI, dI); }
                        dg = upM;
                        leftM = m;
                    }
                    if (K.backtrace) {
                        uint32_t *d = frow + (size_t)rec * fstep;
                        if (SWG) *reinterpret_cast<uint2 *>(d) = make_uint2(aP | (aQ << 16), aD | (aI << 16));
                        else d[0] = aP | (aQ << 16);
                    }
                }
            }
            for (; v + 15 <= headend; v += 16) {
                uint32_t *rv = row + (size_t)v * T;
                const bool more = v + 31 <= headend;
                if (more) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) nxt[j] = rv[(size_t)(16 + j) * T];
                    if (!PSMEM) {
                        pnxt[0] = __ldg(reinterpret_cast<const uint2 *>(gp + v + 15));
                        pnxt[1] = __ldg(reinterpret_cast<const uint2 *>(gp + v + 15) + 1);
                    }
                }
                if (PSMEM) {
                    const uint32_t *pv = pat + (size_t)((v - 1) >> 2) * T;
                    pcur[0] = make_uint2(pv[0], pv[T]);
                    pcur[1] = make_uint2(pv[2 * T], pv[3 * T]);
                }
                uint32_t ne[4];
                ne[0] = __vsetne4(pcur[0].x ^ tc4, 0u);
                ne[1] = __vsetne4(pcur[0].y ^ tc4, 0u);
                ne[2] = __vsetne4(pcur[1].x ^ tc4, 0u);
                ne[3] = __vsetne4(pcur[1].y ^ tc4, 0u);
                uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t old = cur[j];
                    const int upM = unpackM(old);
                    int upI = SWG ? unpackI(old) : 0;
                    const int mm = (int)__dp4a(ne[j >> 2], sel[j & 3], (uint32_t)(dg + MT));
                    int dI, dD, dP, dQ;
                    const int m = dp_cell<ALGO>(upM, upI, leftM, leftD, mm, OE, E, neg1, dI, dD, dP, dQ);
                    cur[j] = pack(m, upI);
                    aP = push_sign(aP, dP);
                    aQ = push_sign(aQ, dQ);
                    if (SWG) { aD = push_sign(aD, dD); aI = push_sign(aI, dI); }
                    dg = upM;
                    leftM = m;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) rv[(size_t)j * T] = cur[j];
                if (K.backtrace) {
                    uint32_t *d = frow + (size_t)((v - 1) >> 4) * fstep;
                    if (SWG) *reinterpret_cast<uint2 *>(d) = make_uint2(aP | (aQ << 16), aD | (aI << 16));
                    else d[0] = aP | (aQ << 16);
                }
                if (more) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
                    if (!PSMEM) { pcur[0] = pnxt[0]; pcur[1] = pnxt[1]; }
                }
            }
            // ---- remaining head cells, then the aliased tail: columns nc..pl read the CURRENT row's
            // head as their "previous row" (flat word nc*(h-1)+v is cell (h, v-nc)) ----
            uint32_t w0 = 0, w1 = 0;  // record under construction
            for (; v <= pl; ++v) {
                const bool tail = v >= nc;
                if (!tail && v > headend) break;
                uint32_t upw;
                if (!tail) upw = row[(size_t)v * T];
                else {
                    upw = row0[(size_t)(v - nc) * T];
                    if (v - 1 >= nc) dg = unpackM(row0[(size_t)(v - 1 - nc) * T]);
                }
                const int upM = unpackM(upw);
                int upI = SWG ? unpackI(upw) : 0;
                const bool nev = (uint32_t)(unsigned char)__ldg(gp + v - 1) != (tc4 & 0xffu);  // (L1; few cells per row)
                const int mm = dg + (nev ? X : MT);
                int dI, dD, dP, dQ;
                const int m = dp_cell<ALGO>(upM, upI, leftM, leftD, mm, OE, E, neg1, dI, dD, dP, dQ);
                const uint32_t old = row[(size_t)v * T];
                row[(size_t)v * T] = pack(m, upI);
                if (v == nc) { tailM = m; tailI = upI; tailD = leftD; }
                const int j = (v - 1) & 15;
                w0 |= (((uint32_t)dP >> 31) << (15 - j)) | (((uint32_t)dQ >> 31) << (31 - j));
                if (SWG) w1 |= (((uint32_t)dD >> 31) << (15 - j)) | (((uint32_t)dI >> 31) << (31 - j));
                if (K.backtrace && (j == 15 || v == pl)) {
                    uint32_t *d = frow + (size_t)((v - 1) >> 4) * fstep;
                    d[0] = w0;
                    if (SWG) d[1] = w1;
                    w0 = w1 = 0;
                }
                dg = unpackM(old);
                leftM = m;
            }
            score = leftM;
        }
        if (pl == 0 || tl == 0) score = 0;

        int begin_offset = pl + tl - 1;
        int status = AIM_STATUS_OK;
        if (K.backtrace) {
            char *ops = K.ops + (size_t)i * 2 * RS;
            int b = pl + tl - 1;
            int h = tl, v = pl;
            int layer = 0;
            while (h > 0 && v > 0) {
                const int fi = nc * h + v;             // flat word the reference's traceback reads
                const int r = min(tl, (fi - 1) / nc);  // its last writer (row r, column c)
                const int c = fi - nc * r;
                bool p, q, opD, opI;
                decode_flags(flg + ((size_t)(r - 1) * rpr + (size_t)((c - 1) >> 4)) * fstep, (c - 1) & 15, SWG, p, q, opD, opI);
                if (!SWG) {
                    if (q) {
                        if (p) { ops[b--] = 'D'; --v; }
                        else { ops[b--] = 'I'; --h; }
                    } else {
                        if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                        --b; --h; --v;
                    }
                } else {
                    if (b < 0) { status = AIM_STATUS_BACKTRACE; break; }
                    if (layer == 2) { ops[b--] = 'D'; if (opD) layer = 0; --v; }
                    else if (layer == 1) { ops[b--] = 'I'; if (opI) layer = 0; --h; }
                    else if (q) layer = p ? 2 : 1;
                    else {
                        if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                        --b; --h; --v;
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
        res.score = score;
        res.status = status;
        res.idx = K.idx_base + i;
        K.results[i] = res;
    }
}

#include "aim_dp_scan.cuh"

// ================= aliased pairs: the row spread over G lanes, min-plus scan of the horizontal gap =================
// See aim_dp_scan.cuh for the algorithm.  32 / G pairs per warp walk their rows in lockstep (rows 1 .. the largest text_len of
// the warp; a pair that is through keeps computing rows nobody reads).  No shared memory: the row lives in 4*C (NW) / 5*C
// registers per lane, so the register file, not the row, bounds the resident pairs.  Predicates: per pair and (row, lane) one
// record {P, Q[, opD, opI]} (complements, bit scan::flag_bit(column - 1); a row of a pair = one 32..128-byte piece), + one 64-bit word per
// (row, pair) for the tail cells; the traceback walks them one pair per thread - the lanes of the warp that filled them (TBIN,
// the default) or dp_scan_tb_kernel.  The records of a pair take READ_SIZE * G * FW words: every warp owns 32 pairs' worth (TBIN),
// or the launcher serves the list in batches (fill, traceback, fill, ...) over one flag region.
// Preconditions (classify_kernel / launcher): text_len < pattern_len, text_len <= 2*C*G, pattern_len - text_len <= min(C, text_len),
// o >= 0, e >= 0, MATCH == 0, every value + the scan's "infinity" inside int16.
template <int C, bool SWG>
struct ScanFlags {
    static constexpr int FW = SWG ? (C == 16 ? 4 : 2) : (C == 16 ? 2 : 1);
    __device__ __forceinline__ static void store(uint32_t *d, uint32_t aP, uint32_t aQ, uint32_t aD, uint32_t aI)
    {
        if (SWG && C == 16) *reinterpret_cast<uint4 *>(d) = make_uint4(aP, aQ, aD, aI);
        else if (SWG) *reinterpret_cast<uint2 *>(d) = make_uint2(aP | (aQ << 16), aD | (aI << 16));
        else if (C == 16) *reinterpret_cast<uint2 *>(d) = make_uint2(aP, aQ);
        else d[0] = aP | (aQ << 16);
    }
    // the records hold the predicates' complements (scanx::push_gt); bit = scan::flag_bit<C>(column - 1)
    __device__ __forceinline__ static void load(const uint32_t *d, int bit, bool &p, bool &q, bool &opD, bool &opI)
    {
        opD = opI = false;
        if (SWG && C == 16) {
            const uint4 w = *reinterpret_cast<const uint4 *>(d);
            p = !((w.x >> bit) & 1u); q = !((w.y >> bit) & 1u); opD = !((w.z >> bit) & 1u); opI = !((w.w >> bit) & 1u);
        } else if (SWG) {
            const uint2 w = *reinterpret_cast<const uint2 *>(d);
            p = !((w.x >> bit) & 1u); q = !((w.x >> (16 + bit)) & 1u); opD = !((w.y >> bit) & 1u); opI = !((w.y >> (16 + bit)) & 1u);
        } else if (C == 16) {
            const uint2 w = *reinterpret_cast<const uint2 *>(d);
            p = !((w.x >> bit) & 1u); q = !((w.y >> bit) & 1u);
        } else {
            const uint32_t w = d[0];
            p = !((w >> bit) & 1u); q = !((w >> (16 + bit)) & 1u);
        }
    }
};

// The traceback of one pair dp_scan_kernel filled, over its predicate records (slot = where they are).  One chain of dependent
// loads per pair: always one pair per THREAD (on one lane of the fill kernel's sub-warps it kept the warp resident for a third
// of its life).
template <int ALGO, int C, int G>
__device__ __forceinline__ void scan_traceback(const FastK &K, uint32_t li, size_t slot)
{
    constexpr bool SWG = (ALGO == AIM_ALGO_SWG);
    constexpr int FW = ScanFlags<C, SWG>::FW;
    const int RS = K.read_size;
    const uint32_t i = K.list[(int64_t)K.list_step * (int64_t)li];
    const int pl = min(max(K.plen[i], 0), RS), tl = min(max(K.tlen[i], 0), RS);
    const char *gp = K.patterns + (size_t)i * RS;
    const char *gt = K.texts + (size_t)i * RS;
    const int nc = tl + 1;
    const uint32_t *flw = K.flags + slot * RS * G * FW;
    const uint2 *tf = K.tflags + slot * RS;
    char *ops = K.ops + (size_t)i * 2 * RS;  // pre-filled with 'M' by the launcher
    int status = AIM_STATUS_OK;
    int b = pl + tl - 1;
    int h = tl, v = pl;
    int layer = 0;  // SWG: 0 M, 1 I, 2 D
    while (h > 0 && v > 0) {
        int r, c;
        scan::last_writer(nc, tl, h, v, r, c);
        bool p, q, opD, opI;
        if (c >= nc) {
            const uint2 t2 = tf[r - 1];
            const uint64_t t = (uint64_t)t2.x | ((uint64_t)t2.y << 32);
            const uint32_t nib = (uint32_t)(t >> (4 * (c - nc))) & 15u;
            p = nib & 1u; q = nib & 2u; opD = nib & 4u; opI = nib & 8u;
        } else {
            const int pos = c - 1;
            const uint32_t *rec = flw + ((size_t)(r - 1) * G + (size_t)(pos / (2 * C))) * FW;
            ScanFlags<C, SWG>::load(rec, scan::flag_bit<C>(pos), p, q, opD, opI);
        }
        if (!SWG) {
            if (q) {
                if (p) { ops[b--] = 'D'; --v; }
                else { ops[b--] = 'I'; --h; }
            } else {
                if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                --b; --h; --v;
            }
        } else {
            if (b < 0) { status = AIM_STATUS_BACKTRACE; break; }
            if (layer == 2) { ops[b--] = 'D'; if (opD) layer = 0; --v; }
            else if (layer == 1) { ops[b--] = 'I'; if (opI) layer = 0; --h; }
            else if (q) layer = p ? 2 : 1;
            else {
                if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                --b; --h; --v;
            }
        }
    }
    if (status == AIM_STATUS_OK) {
        while (h > 0) { ops[b--] = 'I'; --h; }
        while (v > 0) { ops[b--] = 'D'; --v; }
        K.results[i].begin_offset = b + 1;
    } else {
        K.results[i].status = status;
    }
}

// TBIN = false: the traceback as a kernel of its own after every batch of fills.
template <int ALGO, int C, int G>
__global__ void __launch_bounds__(128) dp_scan_tb_kernel(const FastK K)
{
    const uint32_t lend = min(*K.count, K.llimit);
    const uint64_t li64 = (uint64_t)K.lbase + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li64 >= lend) return;
    scan_traceback<ALGO, C, G>(K, (uint32_t)li64, (size_t)(li64 - K.lbase));
}

// TBIN = true: the warp keeps the records of its last 32 pairs (32 / PPW groups) and then walks them itself, one pair per LANE,
// while the SM's other warps fill: the walk is latency-bound and costs the alu-bound fill next to nothing, there is no second
// kernel, no batching, and the record region is resident warps x 32 pairs (2.8 GB at config 3).
template <int ALGO, int C, int G, int MINB, bool TBIN>
__global__ void __launch_bounds__(64, MINB) dp_scan_kernel(const FastK K)
{
    constexpr bool SWG = (ALGO == AIM_ALGO_SWG);
    constexpr int PPW = 32 / G;  // pairs per warp
    constexpr int FW = ScanFlags<C, SWG>::FW;
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, sl = lane & (G - 1), sub = lane / G;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lend = min(*K.count, K.llimit);  // this launch serves list entries [lbase, lend)
    const int RS = K.read_size;
    scan::Pen P;
    P.O = K.o; P.X = K.x; P.MS = K.max_score;
    P.OE = SWG ? K.o + K.e : K.o;
    P.E = SWG ? K.e : K.o;
    P.INF = 32767 - P.E * C - P.OE - 8;
    P.OE2 = scan::both(P.OE); P.E2 = scan::both(P.E); P.INF2 = scan::both(P.INF);
    const int EC = P.E * C;

    constexpr uint32_t KG = 32 / PPW;  // groups per traceback round (TBIN)
    uint32_t it = 0;                   // groups this warp has filled
    uint64_t round_g0 = (uint64_t)K.lbase + (uint64_t)warp * PPW;  // first group of the round under way
    for (uint64_t g0 = (uint64_t)K.lbase + (uint64_t)warp * PPW; g0 < lend; g0 += (uint64_t)nwarps * PPW) {
        const uint32_t li = (uint32_t)g0 + sub;
        const bool have = li < lend;  // a sub-warp without a pair shadows the warp's first pair and writes nothing
        const uint32_t i = K.list[(int64_t)K.list_step * (int64_t)(have ? li : (uint32_t)g0)];
        const int pl = min(max(K.plen[i], 0), RS), tl = min(max(K.tlen[i], 0), RS);
        const char *gp = K.patterns + (size_t)i * RS;
        const char *gt = K.texts + (size_t)i * RS;
        const int d = pl - tl;
        const int tlmax = __reduce_max_sync(FULL, tl), dmax = __reduce_max_sync(FULL, d);
        // the pair's predicate records: [row][lane] + one tail word per row
        const size_t slot = TBIN ? (size_t)warp * 32 + (it % KG) * PPW + sub : (size_t)(li - K.lbase);
        uint32_t *fl = K.flags + (slot * RS * G + sl) * FW;
        uint2 *tf = K.tflags + slot * RS;
        const bool keep = have && K.backtrace;

        scan::Lane<C> L;
        uint32_t tp[C / 4];  // pattern bytes of the tail cells
        {
            uint32_t wlo[C / 4], whi[C / 4];
#pragma unroll
            for (int w = 0; w < C / 4; ++w) {
                const int oa = 2 * C * sl + 4 * w, ob = oa + C;
                wlo[w] = oa < RS ? __ldg(reinterpret_cast<const uint32_t *>(gp + oa)) : 0u;
                whi[w] = ob < RS ? __ldg(reinterpret_cast<const uint32_t *>(gp + ob)) : 0u;
                uint32_t t = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int o = tl + 4 * w + b;
                    t |= (o < RS ? (uint32_t)(unsigned char)__ldg(gp + o) : 0u) << (8 * b);
                }
                tp[w] = t;
            }
            scan::init_lane<C, SWG>(L, sl, P, wlo, whi);
        }
        // where column text_len lives
        const int pt = tl - 1, ot = pt / (2 * C), ht = (pt / C) & 1, rt = pt % C;
        scan::Edge ed;
        ed.bM = ed.bI = ed.bD = 0;
        ed.c0prev = 0;                                    // M(0, 0)
        ed.dgt = SWG ? P.O + tl * P.E : tl * P.OE;        // M(0, text_len)
        int tM = 0, tI = 0, tD = 0, score = 0;

        uint32_t tw = 0, twn = __ldg(reinterpret_cast<const uint32_t *>(gt));
        for (int h = 1; h <= tlmax; ++h) {
            if (((h - 1) & 3) == 0) {  // four text bytes per load, the next four fetched now
                tw = twn;
                if (h + 3 < RS) twn = __ldg(reinterpret_cast<const uint32_t *>(gt) + ((h + 3) >> 2));
            }
            const uint32_t tc = tw & 0xffu, t4 = tc * 0x01010101u;
            tw >>= 8;
            // column 0 of this row (nw.c:114-118 / swg.c:158-166; aliased from row 2 on: cell (h-1, num_cols))
            if (h >= 2) { ed.bM = tM; ed.bI = tI; ed.bD = tD; }
            else if (SWG) { ed.bD = P.MS; ed.bI = P.O + P.E; ed.bM = ed.bI; }
            else { ed.bM = P.OE; ed.bI = 0; ed.bD = 0; }

            // phase 1 + 2
            const uint32_t nb = __shfl_up_sync(FULL, L.uM[C - 1], 1, G);
            const uint32_t dg0 = scan::pack16(sl == 0 ? ed.c0prev : scan::hi16(nb), scan::lo16(L.uM[C - 1]));
            uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
            const uint32_t dl = scan::phase12<C, SWG>(L, dg0, t4, P, aI);
            // phase 3: the true D at the first column of every block
            const int a_lo = scan::lo16(dl), a_hi = scan::hi16(dl);
            int val = min(a_hi, a_lo + EC);
#pragma unroll
            for (int dlt = 1; dlt < G; dlt <<= 1) {
                const int t = __shfl_up_sync(FULL, val, dlt, G);
                if (sl >= dlt) val = min(val, t + 2 * EC * dlt);
            }
            const int din0_own = SWG ? min(ed.bM + P.OE, ed.bD + P.E) : ed.bM + P.OE;
            const int din0 = __shfl_sync(FULL, din0_own, 0, G);
            const int S = min(val, din0 + 2 * EC * (sl + 1));
            const int sp = __shfl_up_sync(FULL, S, 1, G);
            const int in_lo = sl == 0 ? din0 : sp;
            const int in_hi = min(a_lo, in_lo + EC);
            const uint32_t din = scan::pack16(in_lo, in_hi);
            // phase 4
            scan::phase4<C, SWG>(L, din, P, aP, aQ, aD);
            if (SWG) {
                const uint32_t nbn = __shfl_up_sync(FULL, L.uM[C - 1], 1, G);
                const uint32_t mleft = scan::pack16(sl == 0 ? ed.bM : scan::hi16(nbn), scan::lo16(L.uM[C - 1]));
                scan::opd_first<C>(mleft, din, P, aD);
            }
            if (keep && h <= tl)
                ScanFlags<C, SWG>::store(fl + (size_t)(h - 1) * G * FW, scan::compact<C>(aP), scan::compact<C>(aQ), scan::compact<C>(aD), scan::compact<C>(aI));
            // M and del of column text_len, then the tail cells on the first lane
            const uint32_t sm = __shfl_sync(FULL, scan::pick<C>(L.uM, rt), ot, G);
            const uint32_t sd = __shfl_sync(FULL, scan::pick<C>(L.dn, rt), ot, G);
            int lm = ht ? scan::hi16(sm) : scan::lo16(sm);
            const int ld = ht ? scan::hi16(sd) : scan::lo16(sd);
            const int mtl = lm;
            const int dlim = __any_sync(FULL, h == tl) ? dmax : 1;
            const uint64_t tword = scan::tail_cells<C, SWG>(L, ed, lm, ld, tp, tc, d, dlim, P, tM, tI, tD);
            if (keep && sl == 0 && h <= tl) tf[h - 1] = make_uint2((uint32_t)tword, (uint32_t)(tword >> 32));
            if (h == tl) score = lm;
            ed.c0prev = ed.bM;
            ed.dgt = mtl;
        }

        if (have && sl == 0) {  // with backtrace, the traceback fills in begin_offset
            aim_result res;
            res.max_operations = pl + tl;
            res.begin_offset = pl + tl - 1;
            res.end_offset = pl + tl;
            res.score = score;
            res.status = AIM_STATUS_OK;
            res.idx = K.idx_base + i;
            K.results[i] = res;
        }
        ++it;
        if (TBIN && K.backtrace) {
            const bool last = g0 + (uint64_t)nwarps * PPW >= lend;
            if (it % KG == 0 || last) {  // lane t walks the pair of group t / PPW, sub-warp t % PPW of this round: slot warp * 32 + t
                __syncwarp();
                const uint32_t filled = (it - 1) % KG + 1;
                const uint64_t li_t = round_g0 + (uint64_t)(lane / PPW) * nwarps * PPW + (uint32_t)(lane % PPW);
                if ((uint32_t)(lane / PPW) < filled && li_t < lend) scan_traceback<ALGO, C, G>(K, (uint32_t)li_t, (size_t)warp * 32 + lane);
                __syncwarp();  // the next round's rows overwrite the records
                round_g0 = g0 + (uint64_t)nwarps * PPW;
            }
        }
    }
}


inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// Returns AIM_OK after enqueueing, 1 when this parameter set must be served by aim_dp.cu's literal
// kernel (possible int16 truncation, or a row that does not fit shared memory), or an AIM_ERR_*.
int launch_dp_fast(const KernelArgs &a, Scratch *sc, void *stream_v, int *launches)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    const aim_params &p = a.p;
    if (const char *m = getenv("AIM_DP_MODE")) { if (std::string(m) == "literal") return 1; }
    const int RS = p.read_size;
    const bool nw = p.algo == AIM_ALGO_NW;
    {   // every intermediate stays inside int16: the reference's truncations are identities
        const int64_t pen = nw ? std::max(p.mismatch, p.gap_open)
                               : std::max(std::max(p.mismatch, p.gap_open + p.gap_ext), -p.match);
        const int64_t bound = (2 * (int64_t)RS + 2) * pen + (nw ? 0 : (int64_t)p.max_score + p.gap_open + p.gap_ext);
        if (bound >= 32767 || p.max_score < 0) return 1;
    }
    // aliased pairs: threads per block from the shared-memory row + pattern stage
    // aliased pairs: one-warp blocks; stage the pattern in shared memory only if that still leaves >= 8 warps per SM
    const size_t kSmemBudget = 227u * 1024u;
    auto row_cfg = [&](int reg_cols, size_t *per_thread, bool *psm, int *bps) -> bool {
        const size_t cols = reg_cols > 0 ? (size_t)RS + 16 - (size_t)reg_cols : (size_t)RS + 1;
        size_t pt = (cols + (size_t)RS / 4) * 4;
        *psm = (228u * 1024u) / (pt * RT + 1024) >= 8;
        if (!*psm) pt = cols * 4;
        *per_thread = pt;
        *bps = (int)std::min<size_t>(32, (228u * 1024u) / (pt * RT + 1024));
        return pt * RT <= kSmemBudget && *bps >= 1;
    };
    size_t pt0 = 0, pt1 = 0;
    bool psm0 = false, psm1 = false;
    int bps0 = 0, bps1 = 0;
    if (!row_cfg(0, &pt0, &psm0, &bps0)) return 1;
    // register-row variant when shared memory would hold fewer than 8 warps per SM (long rows)
    int reg_cols = (bps0 < 8 && RS >= 160) ? 96 : 0;  // (112 spills: measured slower)
    if (const char *rc_s = getenv("AIM_DP_REGCOLS")) { const int v = atoi(rc_s); if (v == 0 || v == 96 || v == 112) reg_cols = v; }
    if (reg_cols > 0 && !row_cfg(reg_cols, &pt1, &psm1, &bps1)) reg_cols = 0;
    const int row_threads = (int)RT;

    // aliased pairs with the row spread over the lanes of a sub-warp (dp_scan_kernel): geometry by READ_SIZE
    struct ScanCfg { int C, G; void (*fn)(const FastK); void (*tb)(const FastK); };
    ScanCfg scn{0, 0, nullptr, nullptr};
    // the traceback inside the fill kernel (one pair per lane every 32 pairs of a warp) or as a kernel of its own after every batch
    bool tbin = true;
    if (const char *e = getenv("AIM_DP_SCAN_TB")) tbin = std::string(e) != "kernel";
    {
        // 0 off, 1 default geometry, 2 the narrower blocks.  Default: SWG with long rows (where dp_row_kernel's shared-memory row
        // leaves 6-8 warps per SM); NW and short rows stay with dp_row_kernel (config 2: 11.7 against 13.5 ms).
        int mode = (!nw && RS > 144) ? 1 : 0;
        if (const char *e = getenv("AIM_DP_SCAN")) mode = atoi(e);
        const bool pen_ok = p.gap_open >= 0 && p.mismatch >= 0 && (nw || (p.gap_ext >= 0 && p.match == 0));
        if (mode > 0 && pen_ok && RS >= 16 && RS <= 528) {
            // resident blocks per SM the register allocation aims at: 6 (168 registers), 8 (128) or 10 (96; 16-column blocks spill there).
            // Config 3, 16-column blocks: 43.6 / 43.7 / 44.1 ms with 6 / 8 / 10 (the kernel is bound by the alu pipe, not by resident warps).
            int minb = 0;
            if (const char *e = getenv("AIM_DP_SCAN_MINB")) minb = atoi(e);
#define AIM_SCAN_FN(C, G, B)                                                                                                        \
    (tbin ? (nw ? (void (*)(const FastK))dp_scan_kernel<AIM_ALGO_NW, C, G, B, true> : (void (*)(const FastK))dp_scan_kernel<AIM_ALGO_SWG, C, G, B, true>) \
          : (nw ? (void (*)(const FastK))dp_scan_kernel<AIM_ALGO_NW, C, G, B, false> : (void (*)(const FastK))dp_scan_kernel<AIM_ALGO_SWG, C, G, B, false>))
#define AIM_SCAN_TB(C, G) (nw ? (void (*)(const FastK))dp_scan_tb_kernel<AIM_ALGO_NW, C, G> : (void (*)(const FastK))dp_scan_tb_kernel<AIM_ALGO_SWG, C, G>)
#define AIM_SCAN_CFG(C, G, BDEF)                                                                     \
    ((minb ? minb : BDEF) >= 10  ? ScanCfg{C, G, AIM_SCAN_FN(C, G, 10), AIM_SCAN_TB(C, G)}           \
     : (minb ? minb : BDEF) >= 8 ? ScanCfg{C, G, AIM_SCAN_FN(C, G, 8), AIM_SCAN_TB(C, G)}            \
                                 : ScanCfg{C, G, AIM_SCAN_FN(C, G, 6), AIM_SCAN_TB(C, G)})
            if (RS <= 144) scn = mode == 2 ? AIM_SCAN_CFG(4, 16, 10) : AIM_SCAN_CFG(8, 8, 10);
            else if (RS <= 288) scn = mode == 2 ? AIM_SCAN_CFG(8, 16, 10) : AIM_SCAN_CFG(16, 8, 8);
            else scn = AIM_SCAN_CFG(16, 16, 8);
#undef AIM_SCAN_CFG
#undef AIM_SCAN_TB
#undef AIM_SCAN_FN
            const int64_t oe = nw ? p.gap_open : p.gap_open + p.gap_ext, ee = nw ? p.gap_open : p.gap_ext;
            const int64_t pen = std::max<int64_t>(p.mismatch, oe);
            const int64_t top = (2 * (int64_t)RS + 2 + 2 * scn.C * scn.G) * pen + (nw ? 0 : (int64_t)p.max_score) + oe + ee * scn.C + 16;
            if (top >= 32767) scn = ScanCfg{0, 0, nullptr, nullptr};
        }
    }
    int scan_grid = 0;
    uint32_t scan_batch = 0;  // list entries per (fill, traceback) launch pair: their predicate records share one region
    size_t scan_flag_bytes = 0, scan_tail_off = 0;
    if (scn.fn) {
        int bps = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, scn.fn, 64, 0);
        if (e != cudaSuccess || bps < 1) { cudaGetLastError(); scn = ScanCfg{0, 0, nullptr, nullptr}; }
        else {
            const int ppw = 32 / scn.G;
            const int fw = nw ? (scn.C == 16 ? 2 : 1) : (scn.C == 16 ? 4 : 2);
            const size_t per_pair = (size_t)RS * scn.G * fw * 4, per_pair_tail = (size_t)RS * 8;
            size_t budget = (size_t)4 << 30;  // of predicate records in flight
            if (const char *e2 = getenv("AIM_DP_SCAN_BATCH_MB")) { const long v = atol(e2); if (v >= 16 && v <= 65536) budget = (size_t)v << 20; }
            scan_grid = (int)std::min<uint64_t>((uint64_t)sc->sm_count * (uint64_t)bps, ((uint64_t)a.n / ppw + 2) / 2 + 1);
            if (const char *e3 = getenv("AIM_DP_SCAN_GRID")) { const int v = atoi(e3); if (v >= 1 && v < scan_grid) scan_grid = v; }  // (tests: many groups per warp)
            const uint64_t resident = (uint64_t)scan_grid * 2 * ppw;
            uint64_t batch = p.backtrace ? std::max<uint64_t>(budget / (per_pair + per_pair_tail), 1) : a.n;
            if (batch > resident) batch = batch / resident * resident;  // whole waves
            scan_batch = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(batch, 1), a.n);
            size_t slots = scan_batch;
            if (tbin) { scan_batch = a.n; slots = (size_t)scan_grid * 2 * 32; }  // one launch; every warp owns 32 record slots
            scan_tail_off = align_up(p.backtrace ? slots * per_pair : 0, 256);
            scan_flag_bytes = scan_tail_off + (p.backtrace ? slots * per_pair_tail : 0);
            reg_cols = 0;  // its class of the list is the scan kernel's
        }
    }

    const int FW = nw ? 1 : 2;                         // flag words per 16-cell record
    const uint32_t wpr = ((uint32_t)RS + 15) / 16;     // records per row (row kernels)
    const int nstrips_max = (RS + KS - 1) / KS;

    // two pairs per thread in s16x2 lanes (aim_dp_pack2.cuh): every value must be a NON-NEGATIVE int16, i.e. MATCH == 0 (NW ignores MATCH)
    bool pack = nw || p.match == 0;
    if (const char *e = getenv("AIM_DP_PACK")) pack = pack && atoi(e) != 0;
    // the packed ROW kernel (aliased pairs) is selectable: at READ_SIZE 272 its shared-memory row holds 3 warps of 64 pairs per
    // SM against the one-pair-per-thread kernels' 8 warps of 32, and measures slower (54.9 against 39 ms per 420 K pairs)
    bool pack_row = false;
    if (const char *e = getenv("AIM_DP_PACK_ROW")) pack_row = atoi(e) != 0;
    pack_row = pack_row && pack;
    // grids: persistent threads striding over the lists, sized by what is actually resident
    const int strip_block = 128;
    int strip_bps = 0;
    {
        cudaError_t oe;
        if (pack) oe = nw ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&strip_bps, dp2_strip_kernel<AIM_ALGO_NW>, strip_block, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&strip_bps, dp2_strip_kernel<AIM_ALGO_SWG>, strip_block, 0);
        else if (nw) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&strip_bps, dp_strip_kernel<AIM_ALGO_NW, true>, strip_block, 0);
        else if (p.match == 0) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&strip_bps, dp_strip_kernel<AIM_ALGO_SWG, true>, strip_block, 0);
        else oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&strip_bps, dp_strip_kernel<AIM_ALGO_SWG, false>, strip_block, 0);
        if (oe != cudaSuccess || strip_bps < 1) { cudaGetLastError(); strip_bps = 4; }
    }
    if (pack) {  // a thread carries two pairs: half the threads keep the same pairs in flight (and the same scratch)
        int cap = 4;
        if (const char *e = getenv("AIM_DP_PACK_BPS")) { const int v = atoi(e); if (v >= 1 && v <= 16) cap = v; }
        strip_bps = std::min(strip_bps, cap);
    }
    int strip_grid = sc->sm_count * strip_bps;
    strip_grid = (int)std::min<uint64_t>((uint64_t)strip_grid, ((uint64_t)a.n / (pack ? 2 : 1) + strip_block - 1) / strip_block + 1);
    const size_t strip_threads = (size_t)strip_grid * strip_block;
    const size_t strip_mul = pack ? 2 : 1;  // scratch per thread: boundary words and flag words per record double

    // row kernels: launch helper (sets the shared-memory attribute, asks the occupancy, returns the grid)
    struct RowLaunch { void (*fn)(const FastK); size_t smem; int grid; };
    auto pick_row = [&](int rcols, bool psm) -> void (*)(const FastK) {
#define AIM_ROW_PICK(A)                                                                                    \
    (rcols == 112 ? (psm ? dp_row_kernel<A, true, 112> : dp_row_kernel<A, false, 112>)                     \
     : rcols == 96 ? (psm ? dp_row_kernel<A, true, 96> : dp_row_kernel<A, false, 96>)                      \
                   : (psm ? dp_row_kernel<A, true, 0> : dp_row_kernel<A, false, 0>))
        return nw ? AIM_ROW_PICK(AIM_ALGO_NW) : AIM_ROW_PICK(AIM_ALGO_SWG);
#undef AIM_ROW_PICK
    };
    auto prep_row = [&](int rcols, bool psm, size_t pt, int bps_smem, RowLaunch *out) -> cudaError_t {
        out->fn = pick_row(rcols, psm);
        out->smem = pt * RT;
        cudaError_t e = cudaFuncSetAttribute(out->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->smem);
        int bps = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, out->fn, (int)RT, out->smem);
        if (e != cudaSuccess) return e;
        bps = std::max(1, std::min(bps, bps_smem));
        out->grid = (int)std::min<uint64_t>((uint64_t)sc->sm_count * (uint64_t)bps, ((uint64_t)a.n + RT - 1) / RT);
        return cudaSuccess;
    };
    // packed row kernel (two aliased pairs of equal text_len per thread): the row of both pairs in shared memory
    const size_t row2_smem = ((size_t)RS + 1) * (nw ? 1 : 2) * 4 * RT2;
    int row2_grid = 0;
    if (pack_row && row2_smem > kSmemBudget) pack_row = false;
    if (pack_row) {
        void (*fn)(const FastK) = nw ? dp2_row_kernel<AIM_ALGO_NW> : dp2_row_kernel<AIM_ALGO_SWG>;
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row2_smem);
        int bps = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, (int)RT2, row2_smem);
        if (e != cudaSuccess || bps < 1) { set_error(std::string("dp2_row setup: ") + cudaGetErrorString(e)); cudaGetLastError(); return AIM_ERR_CUDA; }
        row2_grid = (int)std::min<uint64_t>((uint64_t)sc->sm_count * (uint64_t)bps, ((uint64_t)a.n / 2 + RT2) / RT2 + 1);
    }
    RowLaunch L0{}, L1{};
    cudaError_t err = prep_row(0, psm0, pt0, bps0, &L0);
    if (err == cudaSuccess && reg_cols > 0) err = prep_row(reg_cols, psm1, pt1, bps1, &L1);
    if (err != cudaSuccess) { set_error(std::string("dp_fast setup: ") + cudaGetErrorString(err)); cudaGetLastError(); return AIM_ERR_CUDA; }
    const size_t rowk_threads = pack_row ? (size_t)row2_grid * RT2 : (size_t)std::max(L0.grid, L1.grid) * row_threads;

    // scratch: counters | list | list2 | strip boundary | strip flags | row flags
    // (packed: list2 = the aliased pairs bucketed by text_len, every bucket starting on an even slot; hist = the buckets)
    const size_t off_list = 256;
    const size_t off_list2 = align_up(off_list + (size_t)a.n * 4, 256);
    const size_t off_hist = align_up(off_list2 + ((size_t)a.n + (size_t)RS + 4) * 4, 256);
    const size_t off_bound = align_up(off_hist + ((size_t)RS + 2) * 4, 256);
    const size_t off_sflags = align_up(off_bound + strip_threads * (size_t)RS * 4 * strip_mul, 256);
    const size_t sflag_bytes = p.backtrace ? strip_threads * (size_t)nstrips_max * RS * FW * 4 * strip_mul : 0;
    const size_t off_rflags = align_up(off_sflags + sflag_bytes, 256);
    const size_t rflag_bytes = std::max(p.backtrace ? rowk_threads * (size_t)RS * wpr * FW * 4 * (pack_row ? 2 : 1) : 0, scan_flag_bytes);
    int rc = scratch_reserve(sc, off_rflags + rflag_bytes);
    if (rc != AIM_OK) return rc;
    unsigned char *base = reinterpret_cast<unsigned char *>(sc->buf);
    uint32_t *counters = reinterpret_cast<uint32_t *>(base);
    uint32_t *list = reinterpret_cast<uint32_t *>(base + off_list);
    uint32_t *list2 = reinterpret_cast<uint32_t *>(base + off_list2);

    FastK K{};
    K.plen = a.plen; K.tlen = a.tlen; K.patterns = a.patterns; K.texts = a.texts;
    K.results = a.results; K.ops = a.ops; K.n = a.n; K.idx_base = a.idx_base;
    K.match = p.match; K.x = p.mismatch; K.o = p.gap_open; K.e = p.gap_ext;
    K.max_score = p.max_score; K.read_size = RS; K.backtrace = p.backtrace;
    K.wpr = wpr;
    K.neg1 = -1;

    uint32_t *hist = reinterpret_cast<uint32_t *>(base + off_hist);
    err = cudaMemsetAsync(counters, 0, 16, stream);
    if (err == cudaSuccess && p.backtrace) err = cudaMemsetAsync(a.ops, 'M', (size_t)a.n * 2 * RS, stream);
    int nlaunch = 0;
    if (err == cudaSuccess && pack_row) {
        err = cudaMemsetAsync(hist, 0, ((size_t)RS + 2) * 4, stream);
        if (err == cudaSuccess) err = cudaMemsetAsync(list2, 0xff, ((size_t)a.n + (size_t)RS + 4) * 4, stream);  // kNoPartner in the pad slots
        if (err == cudaSuccess) {
            classify2_kernel<<<(a.n + 255) / 256, 256, 0, stream>>>(a.plen, a.tlen, a.n, RS, list, counters, hist);
            bucket_scan_kernel<<<1, 1024, 0, stream>>>(hist, (uint32_t)RS + 1, counters);
            bucket_scatter_kernel<<<(a.n + 255) / 256, 256, 0, stream>>>(a.plen, a.tlen, a.n, RS, hist, list2);
            err = cudaGetLastError();
            nlaunch += 3;
        }
    } else if (err == cudaSuccess) {
        classify_kernel<<<(a.n + 255) / 256, 256, 0, stream>>>(a.plen, a.tlen, a.n, RS, reg_cols, scn.fn ? 2 * scn.C * scn.G : 0, scn.C, list, list2,
                                                               counters);
        err = cudaGetLastError();
        ++nlaunch;
    }
    if (err == cudaSuccess) {
        FastK S = K;
        S.list = list; S.count = counters; S.list_step = 1;
        S.bound = reinterpret_cast<uint32_t *>(base + off_bound);
        S.flags = reinterpret_cast<uint32_t *>(base + off_sflags);
        if (pack && nw) dp2_strip_kernel<AIM_ALGO_NW><<<strip_grid, strip_block, 0, stream>>>(S);
        else if (pack) dp2_strip_kernel<AIM_ALGO_SWG><<<strip_grid, strip_block, 0, stream>>>(S);
        else if (nw) dp_strip_kernel<AIM_ALGO_NW, true><<<strip_grid, strip_block, 0, stream>>>(S);
        else if (p.match == 0) dp_strip_kernel<AIM_ALGO_SWG, true><<<strip_grid, strip_block, 0, stream>>>(S);
        else dp_strip_kernel<AIM_ALGO_SWG, false><<<strip_grid, strip_block, 0, stream>>>(S);
        err = cudaGetLastError();
    }
    ++nlaunch;
    if (err == cudaSuccess && pack_row) {  // all aliased pairs, two of equal text_len per thread
        FastK R2 = K;
        R2.list = list2; R2.count = counters + 1; R2.list_step = 1;
        R2.flags = reinterpret_cast<uint32_t *>(base + off_rflags);
        if (nw) dp2_row_kernel<AIM_ALGO_NW><<<row2_grid, RT2, row2_smem, stream>>>(R2);
        else dp2_row_kernel<AIM_ALGO_SWG><<<row2_grid, RT2, row2_smem, stream>>>(R2);
        err = cudaGetLastError();
        ++nlaunch;
    }
    if (err == cudaSuccess && !pack_row && scn.fn) {  // aliased pairs: the row over the lanes of a sub-warp
        FastK Sc = K;
        Sc.list = list + (a.n - 1); Sc.count = counters + 1; Sc.list_step = -1;
        Sc.flags = reinterpret_cast<uint32_t *>(base + off_rflags);
        Sc.tflags = reinterpret_cast<uint2 *>(base + off_rflags + scan_tail_off);
        // Traceback inside the kernel: one launch.  Otherwise: how many of the n pairs are in this class is known on the device only:
        // batches over the whole range, empty ones return at once; fill and traceback alternate on the one stream.  Measured and rejected: the traceback of batch b on a side stream
        // under the fill of batch b + 1 (two halves of the record region, the fill kernel one block per SM short so that the traceback
        // blocks fit): 44.6 against 43.7 ms at config 3 - the fill kernel loses more with 14 warps per SM than the 2.9 ms it hides.
        for (uint64_t b0 = 0; b0 < a.n && err == cudaSuccess; b0 += scan_batch) {
            Sc.lbase = (uint32_t)b0;
            Sc.llimit = (uint32_t)std::min<uint64_t>(b0 + scan_batch, a.n);
            scn.fn<<<scan_grid, 64, 0, stream>>>(Sc);
            ++nlaunch;
            if (p.backtrace && !tbin) {
                scn.tb<<<(Sc.llimit - Sc.lbase + 127) / 128, 128, 0, stream>>>(Sc);
                ++nlaunch;
            }
            err = cudaGetLastError();
        }
    }
    if (err == cudaSuccess && !pack_row && reg_cols > 0) {  // aliased pairs that fit the register-row variant
        FastK Rr = K;
        Rr.list = list + (a.n - 1); Rr.count = counters + 1; Rr.list_step = -1;
        Rr.flags = reinterpret_cast<uint32_t *>(base + off_rflags);
        L1.fn<<<L1.grid, row_threads, L1.smem, stream>>>(Rr);
        err = cudaGetLastError();
        ++nlaunch;
    }
    if (err == cudaSuccess && !pack_row) {  // the other aliased pairs (all of them when reg_cols == 0)
        FastK Rs = K;
        Rs.list = list2; Rs.count = counters + 2; Rs.list_step = 1;
        Rs.flags = reinterpret_cast<uint32_t *>(base + off_rflags);
        L0.fn<<<L0.grid, row_threads, L0.smem, stream>>>(Rs);
        err = cudaGetLastError();
        ++nlaunch;
    }
    if (err != cudaSuccess) { set_error(std::string("dp_fast launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    if (launches) *launches += nlaunch;
    return AIM_OK;
}

}  // namespace aim
