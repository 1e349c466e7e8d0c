// Short-read WFA / WFA-adaptive for sm_100a: G lanes per pair, 32/G pairs per warp in LOCKSTEP.
//
// Same algorithm and literal semantics as aim_wfa.cu (reference: WFA/DPU-MRAM/dpu/wfa.c:70-407,
// wfa_backtracing.c:219-375), restructured around one observation: without adaptive trimming the
// wavefront SCHEDULE of gap-affine WFA - which scores have a wavefront, its [lo,hi], whether it has
// I/D components - depends only on the penalties, not on the sequences (wfa.c:275-354 derives it from
// the ranges of scores s-x, s-o-e, s-e).  So
//   * the schedule is computed once on the host into a per-score PLAN (8 words) shared by every pair;
//     adaptive trimming (wfa.c:70-141) only narrows [lo,hi] per pair (one word per score and pair);
//   * all pairs of a warp walk the scores together: the loop structure is warp-uniform, the lanes of a
//     sub-warp split the diagonals of their pair, and the only divergence left is the data-dependent
//     extend length, the trimmed widths and the per-pair finishing score;
//   * compute_offsets only ever reads M of scores s-x and s-o-e and I/D of score s-e, so shared memory
//     holds two small RINGS (M: max(x,o+e)+1 wavefronts, I/D: e+1 wavefronts) instead of the history;
//   * the full history the backtrace needs (the reference's MRAM component store,
//     dpu_allocator_mram.c) is streamed to a per-pair-slot HBM arena, one 8-byte {M,I,D} cell per
//     store, and read back only along the backtrace path; the arena is reused pair after pair and
//     stays L2-resident;
//   * compute_offsets and extend are fused per diagonal (extending diagonal k touches only M[k]);
//   * every ring row is NULL-FRAMED: cells outside the occupant wavefront's (trimmed) range hold exactly the NULL
//     the reference substitutes for an out-of-range read (wfa.c:243-266), and missing components point at a
//     block-wide all-NULL row, so the inner loop has no range or existence tests.  The frame is kept by three small
//     clean-ups: rows are NULL-filled when a pair starts, cells an adaptive trim cuts off are NULLed at once, and
//     cells a slot's previous occupant left outside the new wavefront's range are NULLed before the slot is reused.
//     The reference's -10 "component missing" sentinel is one max() with a per-score floor (-10 when any of I / D /
//     sub is missing, 0 for score 0);
//   * the ASCII -> 2-bit packing (and the non-ACGT check) is a separate HBM-bound kernel, wfa_prep_kernel, that
//     writes every 16-base window as one aligned 8-byte entry {word j, word j+1}; extend is one 8-byte shared load
//     + one funnel shift per sequence; the 'M' pre-fill of the op rows (wfa.c:499-501) is a cudaMemsetAsync;
//   * backtraces of the 32/G pairs run concurrently on the sub-warps' first lanes and write the few
//     non-'M' ops straight into the 'M'-filled global op rows.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "aim_internal.h"
#include "aim_wfa_common.cuh"

namespace aim {

namespace {

// plan flags
constexpr uint32_t P_PRESENT = 1, P_SUB_NULL = 2, P_O_NULL = 4, P_IE_NULL = 8, P_DE_NULL = 16, P_HAS_I = 32, P_HAS_D = 64, P_NULL_ROW = 128;
constexpr int PLAN_WORDS = 8;
constexpr uint32_t kNull2 = 0xc000c000u;    // kNull in both halfwords
constexpr uint32_t kNoRange = 0x7fff7fffu;  // {lo, -hi} of an empty range: neutral for the packed min
constexpr uint32_t OFF_NULL = 0xffffu;  // row offset meaning "the block-wide all-NULL row"; also "no score"
// per-score plan: w0 flags | floor (int16) << 16
//                 w1 lo_s (int16) | (-hi_s) << 16        static range of the wavefront (ranges are kept as {lo, -hi}
//                                                        so that one packed min() merges them)
//                 w2 row offset of this score's M row   | of score s-x's M row << 16
//                 w3 row offset of score s-o-e's M row  | of score s-e's {I,D} row << 16
//                 w4 row offset of this score's {I,D} row
//                 w5 arena cell index of diagonal lo_s
//                 w6 previous occupant (score) of this M row << 16
//                 w7 previous occupant (score) of this {I,D} row
// rows: byte offsets from the pair's first row; M ring (cw int16 each, cw % 8 == 0), then the {I,D} ring (cw 32-bit cells
// each: I in the low half, D in the high half).  A present score writes all of M, I, D over its range - a component the
// reference does not have comes out as NULL by itself because its sources are NULL rows.

struct SubK {
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    const uint4 *packed;    // wfa_prep_kernel's output: per pair seq_words 2-bit packed words of the pattern, then of the text
    const uint32_t *flags;  // one bit per pair: a byte outside {A,C,G,T} inside the sequences -> the pair is on the hand-over list
    const uint32_t *plan;   // device copy of the plan
    uint2 *arena;           // history arena (BT only): {M | I << 16, D} per (score, diagonal)
    size_t arena_stride;    // cells per pair slot
    uint32_t n, idx_base;
    int x, o, e;
    int max_score, read_size;
    int koff;               // ring cell of diagonal k is k + koff
    uint32_t plan_words;    // PLAN_WORDS * (max_score + 1)
    uint32_t seq_words;     // 32-bit words per packed sequence (multiple of 4, one spare word past the last base)
    uint32_t dyn_words;     // trimmed-range words per pair (multiple of 4; 0 when !reduce)
    uint32_t cw;            // row width in halfwords (multiple of 8)
    uint32_t rows_v4;       // 16-byte units of all rows of a pair
    uint32_t pair_words;    // shared-memory words per pair slot
    int bias;               // NARROW rows: stored byte = offset + bias; the NULL family (NULL + d) is stored as d
    int null_max;           // NARROW rows: bytes <= null_max are the NULL family
    uint32_t own_bytes;     // POOL: bytes of a warp's cell table (16-bit entries, PPW * cw of them), after the pair slots
};

// ---- shared-memory accessors on 32-bit shared-window addresses ----
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_v2(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_u16(uint32_t a) { int v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_u16s(uint32_t a) { return lds_u16(a); }
__device__ __forceinline__ int lds_u8(uint32_t a) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(uint32_t a, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((short)v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One thread per (pair, sequence, 16-base word): word j = bases 16j..16j+15, first base in the top two bits.  HBM-bound:
// a block covers 256 / SW consecutive rows per step (a thread keeps its word index j, so there is no division in the loop)
// and every thread has two rows' loads in flight.
// A pair holding a byte outside {A,C,G,T} (the reference compares raw bytes, wfa.c:209) cannot be 2-bit packed: it is
// flagged and listed once for the warp-per-pair kernel of aim_wfa.cu, which compares bytes.
__global__ void __launch_bounds__(256) wfa_prep_kernel(const char *patterns, const char *texts, const int32_t *plen, const int32_t *tlen,
                                                       uint32_t n, int RS, uint32_t SW, uint32_t *packed, uint32_t *flags,
                                                       uint32_t *list, uint32_t *list_count)
{
    const uint32_t rpb = 256u / SW;  // rows (sequences) per block and step
    const uint32_t jr = threadIdx.x / SW, j = threadIdx.x - jr * SW;
    if (jr >= rpb) return;
    const uint64_t rows = 2ull * n, step = (uint64_t)gridDim.x * rpb;
    const int b = (int)j * 16;
    struct Raw { uint2 lo, hi; int len; };
    auto fetch = [&](uint64_t r) {
        Raw w{make_uint2(0, 0), make_uint2(0, 0), 0};
        const uint32_t i = (uint32_t)(r >> 1);
        const bool q = r & 1u;
        w.len = min(max(q ? tlen[i] : plen[i], 0), RS);
        const char *row = (q ? texts : patterns) + (size_t)i * RS;
        if (b < w.len && b < RS) w.lo = __ldg(reinterpret_cast<const uint2 *>(row + b));
        if (b + 8 < w.len && b + 8 < RS) w.hi = __ldg(reinterpret_cast<const uint2 *>(row + b + 8));
        return w;
    };
    auto finish = [&](uint64_t r, const Raw &w) {
        bool ok = true;
        uint32_t h0 = 0, h1 = 0;
        if (b < w.len && b < RS) h0 = pack8(w.lo, w.len - b, &ok);
        if (b + 8 < w.len && b + 8 < RS) h1 = pack8(w.hi, w.len - b - 8, &ok);
        packed[r * SW + j] = (h0 << 16) | h1;
        const uint32_t i = (uint32_t)(r >> 1);
        if (!ok && !(atomicOr(&flags[i >> 5], 1u << (i & 31)) & (1u << (i & 31)))) list[atomicAdd(list_count, 1u)] = i;  // first to flag it
    };
    for (uint64_t r = (uint64_t)blockIdx.x * rpb + jr; r < rows; r += 2 * step) {
        const bool two = r + step < rows;
        const Raw w0 = fetch(r);
        Raw w1{make_uint2(0, 0), make_uint2(0, 0), 0};
        if (two) w1 = fetch(r + step);
        finish(r, w0);
        if (two) finish(r + step, w1);
    }
}

// Extend (wfa.c:193-215) in two steps so that several diagonals' first steps can be scheduled together.
// extend_first: the number of equal bases in the first 16-base window from pattern[v = m - k], text[m] (16 = all equal) and
// *lim = how far the run may go (<= 0: nothing to extend, which includes negative m or v).  aP/aT = shared addresses of
// the packed words (two 4-byte loads + one funnel shift per sequence).  Nearly every run ends inside the first window; extend_more continues the others.
__device__ __forceinline__ int extend_first(uint32_t aP, uint32_t aT, int k, int m, int pl, int tl, int *lim)
{
    const int v = m - k;
    *lim = (m | v) >= 0 ? min(pl - v, tl - m) : 0;
    const int vc = max(v, 0), hc = max(m, 0);  // any in-buffer window will do when there is nothing to extend
    const uint32_t ax = aP + ((uint32_t)(vc >> 4) << 2), ay = aT + ((uint32_t)(hc >> 4) << 2);
    const uint32_t x0 = lds_u32(ax), x1 = lds_u32(ax + 4), y0 = lds_u32(ay), y1 = lds_u32(ay + 4);
    const uint32_t d = __funnelshift_l(x1, x0, 2 * vc) ^ __funnelshift_l(y1, y0, 2 * hc);
    return __clz(d) >> 1;
}
__device__ __noinline__ int extend_more(uint32_t aP, uint32_t aT, int v, int h, int lim)
{
    int cnt = 16;
    for (;;) {
        const int pv = v + cnt, ph = h + cnt;
        const uint32_t ax = aP + ((uint32_t)(pv >> 4) << 2), ay = aT + ((uint32_t)(ph >> 4) << 2);
        const uint32_t d2 = __funnelshift_l(lds_u32(ax + 4), lds_u32(ax), 2 * pv) ^ __funnelshift_l(lds_u32(ay + 4), lds_u32(ay), 2 * ph);
        if (d2) { cnt += __clz(d2) >> 1; break; }
        cnt += 16;
        if (cnt >= lim) break;
    }
    return cnt;
}

template <int G>
__device__ __forceinline__ int group_max(int v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(kFull, v, d));
    return v;
}
template <int G>
__device__ __forceinline__ int group_min(int v)
{
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16s(uint32_t w) { return (int)(short)(w >> 16); }
// lo <= k <= hi
__device__ __forceinline__ bool in_range(int k, int lo, int hi) { return (unsigned)(k - lo) <= (unsigned)(hi - lo) && lo <= hi; }

// NB (narrow rows): the ring rows hold ONE BYTE per offset instead of an int16, which fits 32 warps per SM where the int16 rows fit
// 20 (config 4: 0.8 KB of shared memory per pair instead of 1.36 KB).  A byte holds offset + bias for every value the reference
// can produce from -bias up, and d for the NULL family NULL + d (the reference's NULL = -16384 drifts by +1 per wavefront it passes
// through: wfa.c:249-266); with d <= MAX_SCORE < bias - 10 the two families keep their order and their ties, every max()
// and +1 of compute_offsets is the same operation on the bytes, and the literal value is recovered wherever a DIFFERENCE to a
// true offset is taken (adaptive distances, backtrace).  Used when READ_SIZE + 2*MAX_SCORE + 12 <= 255; the history cell shrinks to
// four bytes {M, I, D}.
// POOL (narrow rows, G = 4): the cells of a wavefront are dealt to the 32 lanes of the WARP instead of the four lanes of their pair.
// The pairs of a warp walk the scores together, but adaptive trimming gives them different widths and they finish at different
// scores, so with four fixed lanes per pair the warp runs the widest pair's trips (18 of 32 lanes busy, ncu).  Per score the
// sub-warps write their pair's cells (pair, diagonal) into a warp-wide table in shared memory, in pair order; lane l then takes
// cells l, l + 32, ... whatever pair they belong to (row addresses from the pair number, lengths by shuffle from the pair's
// lanes) and leaves each cell's distance to the end in the cell's table entry, from which the pair's lanes take the minimum the
// reduction needs.  Range computation, frame clean-up, end test, reduction and backtrace stay with the pair's own lanes.
template <int G, bool REDUCE, bool BT, int MAXT = 128, bool NB = false, bool POOL = false>
__global__ void __launch_bounds__(MAXT) wfa_sub_kernel(const SubK K)
{
    static_assert(!POOL || (NB && G == 4), "the pooled cell loop is written for narrow rows with four lanes per pair");
    constexpr int PPW = 32 / G;
    constexpr uint32_t ES = NB ? 1u : 2u;  // bytes per M cell; an {I,D} cell is 2 * ES bytes (I first)
    const int BZ = NB ? K.bias : 0;
    constexpr int NULLV = NB ? 0 : kNull;
    auto ldM = [](uint32_t a) -> int { return NB ? lds_u8(a) : lds_s16(a); };
    auto stM = [](uint32_t a, int v) { if (NB) sts_u8(a, v); else sts_u16(a, v); };
    auto stID_null = [](uint32_t a) { if (NB) sts_u16(a, 0); else sts_u32(a, kNull2); };
    // literal (reference) value of a stored offset
    auto lit = [&](int x) -> int { return NB ? (x <= K.null_max ? kNull + x : x - BZ) : x; };
    extern __shared__ __align__(16) uint32_t smem_w[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / G, sl = lane % G;
    const uint32_t wpb = blockDim.x >> 5;
    const int RS = K.read_size, MS = K.max_score;
    const int X = K.x, OE = K.o + K.e, E = K.e;
    const uint32_t nullw = NB ? 0u : 0xc000c000u;  // NULL in every cell of a word
    const uint4 null4 = make_uint4(nullw, nullw, nullw, nullw);

    // block-wide: the plan and the all-NULL row
    for (uint32_t j = threadIdx.x; j < K.plan_words; j += blockDim.x) smem_w[j] = K.plan[j];
    for (uint32_t j = threadIdx.x; j < K.cw; j += blockDim.x) smem_w[K.plan_words + j] = nullw;  // wide enough for an {I,D} row
    __syncthreads();

    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_w);
    const uint32_t aPlan = sbase;
    const uint32_t aSlot = sbase + (K.plan_words + K.cw + (uint32_t)(wib * PPW + sub) * K.pair_words) * 4u;
    const uint32_t aP = aSlot;
    const uint32_t aT = aP + K.seq_words * 4u;
    const uint32_t aDyn = aT + K.seq_words * 4u;
    const uint32_t aRows = aDyn + K.dyn_words * 4u;
    // M rows hold one int16 per diagonal, {I,D} rows one 32-bit cell (I low, D high); diagonal 0 of the row at plan
    // offset `off` is aK0 + off (M rows) / aK0D + off ({I,D} rows); OFF_NULL = the block-wide all-NULL row
    const uint32_t aK0 = aRows + (uint32_t)K.koff * ES, aK0D = aRows + (uint32_t)K.koff * 2u * ES;
    const uint32_t nullrel = (sbase + K.plan_words * 4u + (uint32_t)K.koff * ES) - aK0;
    const uint32_t nullrelD = (sbase + K.plan_words * 4u + (uint32_t)K.koff * 2u * ES) - aK0D;
#define AIM_ROW(off) (aK0 + ((off) == OFF_NULL ? nullrel : (off)))
#define AIM_ROWD(off) (aK0D + ((off) == OFF_NULL ? nullrelD : (off)))

    const uint32_t aOwn = sbase + (K.plan_words + K.cw + wpb * PPW * K.pair_words) * 4u + (uint32_t)wib * K.own_bytes;  // POOL
    const uint32_t aSlot0 = aSlot - (uint32_t)sub * K.pair_words * 4u;                                                // slot of the warp's first pair
    const uint32_t slot_global = (blockIdx.x * wpb + wib) * PPW + sub;
    const uint32_t nslots = gridDim.x * wpb * PPW;
    uint2 *arena = (BT && !NB) ? K.arena + (size_t)slot_global * K.arena_stride : nullptr;
    uint32_t *arena1 = (BT && NB) ? reinterpret_cast<uint32_t *>(K.arena) + (size_t)slot_global * K.arena_stride : nullptr;  // {M, I, D} bytes

    for (uint32_t base_i = 0; base_i < K.n; base_i += nslots) {  // warp-uniform trip count
        const uint32_t i = base_i + slot_global;
        const bool active = i < K.n && !((K.flags[i >> 5] >> (i & 31)) & 1u);  // flagged pairs go to the warp-per-pair kernel
        const int pl = active ? min(max(K.plen[i], 0), RS) : 0;
        const int tl = active ? min(max(K.tlen[i], 0), RS) : 0;
        char *gops = BT ? K.ops + (size_t)(active ? i : 0) * 2 * RS : nullptr;
        const int ak = tl - pl;

        // ---- stage the packed windows (16-byte copies), NULL-fill the rows ----
        if (active) {
            const uint4 *src = K.packed + (size_t)i * (K.seq_words / 2);
            for (uint32_t c = sl; c < K.seq_words / 2; c += G) sts_v4(aP + c * 16u, __ldg(src + c));
        }
        for (uint32_t c = sl; c < K.rows_v4; c += G) sts_v4(aRows + c * 16u, null4);
        __syncwarp();

        bool done = !active;
        bool reached = false;
        int fscore = MS + 1;  // give-up value (wfa.c:399-404)

        for (int s = 0; s <= MS; ++s) {
            const uint4 p0 = lds_v4(aPlan + (uint32_t)s * (PLAN_WORDS * 4));  // warp-uniform
            const uint32_t fl = p0.x;
            if (!(fl & P_PRESENT)) continue;
            const uint4 p1 = lds_v4(aPlan + (uint32_t)s * (PLAN_WORDS * 4) + 16);
            // the reference's -10 for a missing candidate is a floor under the max (none: -32768, which no stored value is below)
            const int floor_m = NB ? (hi16s(fl) < -16000 ? 0 : hi16s(fl) + BZ) : hi16s(fl);
            const int lo_s = lo16(p0.y);

            // this pair's range (wfa.c:318-343): from the (trimmed) ranges of the source wavefronts; ranges are {lo, -hi}
            uint32_t rng = p0.y;
            if (REDUCE && s > 0) {
                uint32_t ra = kNoRange, rb = kNoRange, re = kNoRange;
                if (!(fl & P_SUB_NULL)) ra = lds_u32(aDyn + (uint32_t)(s - X) * 4u);
                if (!(fl & P_O_NULL)) rb = lds_u32(aDyn + (uint32_t)(s - OE) * 4u);
                if ((fl & (P_IE_NULL | P_DE_NULL)) != (P_IE_NULL | P_DE_NULL)) re = lds_u32(aDyn + (uint32_t)(s - E) * 4u);
                rng = __vadd2(__vimin3_s16x2(ra, rb, re), 0xffffffffu);  // {min lo - 1, -(max hi + 1)}
            }
            const int lo = lo16(rng), hi = -hi16s(rng);
            // row addresses of diagonal 0
            const uint32_t aNM = aK0 + (p0.z & 0xffffu), aN = aK0D + (p1.x & 0xffffu);
            uint32_t aAM = aK0 + (p0.z >> 16), aBM = aK0 + (p0.w & 0xffffu), aE = aK0D + (p0.w >> 16);
            if (fl & P_NULL_ROW) {  // early scores only: a missing source reads the all-NULL row
                aAM = AIM_ROW(p0.z >> 16); aBM = AIM_ROW(p0.w & 0xffffu); aE = AIM_ROWD(p0.w >> 16);
            }

            // ---- keep the frame: NULL what the rows' previous occupants left outside [lo,hi] ----
            {
                const uint32_t prevM = p1.z >> 16, prevID = p1.w & 0xffffu;
                uint32_t pr = kNoRange, qr = kNoRange;
                if (prevM != OFF_NULL) pr = REDUCE ? lds_u32(aDyn + prevM * 4u) : lds_u32(aPlan + prevM * (PLAN_WORDS * 4) + 4);
                if (prevID != OFF_NULL) qr = REDUCE ? lds_u32(aDyn + prevID * 4u) : lds_u32(aPlan + prevID * (PLAN_WORDS * 4) + 4);
                // {lo,-hi} inside this range <=> elementwise >= ; an empty previous range is kNoRange
                const bool left = !done && (__vmins2(pr, rng) != rng || __vmins2(qr, rng) != rng);
                if (__any_sync(kFull, left)) {
                    if (left) {
                        const int plo = lo16(pr), phi = -hi16s(pr), qlo = lo16(qr), qhi = -hi16s(qr);
                        // only the (few) cells of the previous range that stick out below lo or above hi are visited
                        for (int k = plo + sl; k <= min(phi, lo - 1); k += G) stM(aNM + (uint32_t)k * ES, NULLV);
                        for (int k = max(plo, hi + 1) + sl; k <= phi; k += G) stM(aNM + (uint32_t)k * ES, NULLV);
                        for (int k = qlo + sl; k <= min(qhi, lo - 1); k += G) stID_null(aN + (uint32_t)k * 2u * ES);
                        for (int k = max(qlo, hi + 1) + sl; k <= qhi; k += G) stID_null(aN + (uint32_t)k * 2u * ES);
                    }
                }
            }

            // ---- compute_offsets (wfa.c:238-273) fused with extend (wfa.c:193-215); no range tests: see the header ----
            int md = max(pl, tl);
            bool pooled = false;
            if constexpr (POOL) {
                if (!(fl & P_NULL_ROW)) {  // (early scores read the block-wide NULL row for a missing source: they keep the per-pair loop)
                    pooled = true;
                    // ---- the warp's cell table: pair p's cells at [pre_p, pre_p + w_p) ----
                    const int w = (!done && hi >= lo) ? hi - lo + 1 : 0;
                    int incl = w;
                    {
                        int t = __shfl_up_sync(kFull, incl, G);
                        if (lane >= G) incl += t;
                        t = __shfl_up_sync(kFull, incl, 2 * G);
                        if (lane >= 2 * G) incl += t;
                        t = __shfl_up_sync(kFull, incl, 4 * G);
                        if (lane >= 4 * G) incl += t;
                    }
                    const int pre = incl - w;
                    const int W = __shfl_sync(kFull, incl, 31);
                    for (int j = sl; j < w; j += G) sts_u16(aOwn + 2u * (uint32_t)(pre + j), (sub << 8) | (lo + j + 128));
                    __syncwarp();
                    // row offsets from a pair's slot (the layout is the same for every pair)
                    const uint32_t uNM = aNM - aSlot, uAM = aAM - aSlot, uBM = aBM - aSlot, uE = aE - aSlot, uN = aN - aSlot;
                    const uint32_t pairb = K.pair_words * 4u, seqb = K.seq_words * 4u;
                    const uint32_t mypt = (uint32_t)pl | ((uint32_t)tl << 16);
                    const uint32_t hist0 = p1.y - (uint32_t)lo_s;  // arena cell of diagonal k: hist0 + k
                    uint32_t *arena_w0 = BT ? arena1 - (size_t)sub * K.arena_stride : nullptr;
                    const uint32_t astride = (uint32_t)K.arena_stride;
                    for (int c0 = 0; c0 < W; c0 += 64) {  // warp-uniform trips, two cells per lane
                        constexpr int N = 2;
                        bool ok[N];
                        int kk[N], cpl[N], ctl[N], g1[N] = {}, g2[N] = {}, ii[N] = {}, dd[N] = {}, sb[N] = {}, mm[N], cc[N], ll[N];
                        uint32_t ps[N], slot[N], idd[N];
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const int c = c0 + 32 * j + lane;
                            ok[j] = c < W;
                            const uint32_t en = ok[j] ? (uint32_t)lds_u16(aOwn + 2u * (uint32_t)c) : 128u;
                            ps[j] = en >> 8;
                            kk[j] = (int)(en & 0xffu) - 128;
                            const uint32_t pt = __shfl_sync(kFull, mypt, (int)ps[j] * G);
                            cpl[j] = (int)(pt & 0xffffu);
                            ctl[j] = (int)(pt >> 16);
                            slot[j] = aSlot0 + ps[j] * pairb;
                        }
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            if (ok[j]) {
                                const uint32_t a1 = slot[j] + (uint32_t)kk[j], a2 = a1 + (uint32_t)kk[j];
                                g1[j] = lds_u8(a1 + uBM - 1u); g2[j] = lds_u8(a1 + uBM + 1u);
                                ii[j] = lds_u8(a2 + uE - 2u); dd[j] = lds_u8(a2 + uE + 3u);  // I of cell k-1, D of cell k+1
                                sb[j] = lds_u8(a1 + uAM);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const int t = max(g1[j], ii[j]) + 1;
                            const int ins = t == 1 ? 0 : t;  // both NULL -> NULL (wfa.c:249-252)
                            const int del = max(g2[j], dd[j]);
                            mm[j] = max(max(del, sb[j] + 1), max(ins, floor_m));
                            idd[j] = (uint32_t)ins | ((uint32_t)del << 8);
                        }
#pragma unroll
                        for (int j = 0; j < N; ++j)
                            if (ok[j]) sts_u16(slot[j] + 2u * (uint32_t)kk[j] + uN, (int)idd[j]);
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            cc[j] = 0; ll[j] = 0;
                            if (ok[j]) cc[j] = extend_first(slot[j], slot[j] + seqb, kk[j], mm[j] - BZ, cpl[j], ctl[j], &ll[j]);
                        }
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            if (ok[j]) {
                                int m = mm[j], cnt = cc[j];
                                if (cnt == 16 && ll[j] > 16) cnt = extend_more(slot[j], slot[j] + seqb, m - BZ - kk[j], m - BZ, ll[j]);  // rare
                                m += max(min(cnt, ll[j]), 0);
                                sts_u8(slot[j] + (uint32_t)kk[j] + uNM, m);
                                if (BT) arena_w0[ps[j] * astride + (hist0 + (uint32_t)kk[j])] = (uint32_t)m | (idd[j] << 8);
                                if (REDUCE) {
                                    // the cell's distance to the end (a NULL-family cell is -16384 + d in the reference: never the minimum)
                                    const int d = m > K.null_max ? max(cpl[j] + kk[j], ctl[j]) + BZ - m : 0x7fff;
                                    sts_u16(aOwn + 2u * (uint32_t)(c0 + 32 * j + lane), d);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (REDUCE)
                        for (int j = sl; j < w; j += G) md = min(md, lds_u16s(aOwn + 2u * (uint32_t)(pre + j)));
                }
            }
            if (!done && !pooled) {
                // this lane's cells are k0, k0 + G, ...; row pointers at k0, advanced once per trip
                const int k0 = lo + sl;
                uint32_t rB = aBM + (uint32_t)k0 * ES, rA = aAM + (uint32_t)k0 * ES, rNM = aNM + (uint32_t)k0 * ES;
                uint32_t rE = aE + (uint32_t)k0 * 2u * ES, rN = aN + (uint32_t)k0 * 2u * ES;
                uint2 *hp = (BT && !NB) ? arena + (p1.y + (uint32_t)(k0 - lo_s)) : nullptr;  // arena cell of (s, k0)
                uint32_t *hp1 = (BT && NB) ? arena1 + (p1.y + (uint32_t)(k0 - lo_s)) : nullptr;
                // back half of a cell at j*G past the pointers: finish the extend, store M, history cell, distance for the reduction
                auto back = [&](const int k, const int j, int m, const uint32_t id, int cnt, const int lim) {
                    if (cnt == 16 && lim > 16) cnt = extend_more(aP, aT, m - BZ - k, m - BZ, lim);  // rare
                    m += max(min(cnt, lim), 0);
                    stM(rNM + (uint32_t)j * ES * G, m);
                    if (BT) {
                        if (NB) hp1[j * G] = (uint32_t)m | (id << 8);
                        else hp[j * G] = make_uint2((uint32_t)m & 0xffffu, id);
                    }
                    if (REDUCE) {
                        // (a NULL-family cell is -16384 + d in the reference: its distance never is the minimum)
                        if (NB) { if (m > K.null_max) md = min(md, max(pl + k, tl) + BZ - m); }
                        else md = min(md, max(pl + k, tl) - m);
                    }
                };
                int k = k0;
                // N cells per trip in PHASES - all source loads, the recurrences, the {I,D} stores, the first extend windows,
                // the back halves - so that the N cells' shared-memory loads are in flight together (a store between two
                // cells' loads would order them: the rows may alias as far as the compiler knows)
                auto trip = [&](auto nc) {
                    constexpr int N = decltype(nc)::value;
                    int g1[N], g2[N], ii[N], dd[N], sb[N], mm[N], cc[N], ll[N];
                    uint32_t idd[N];
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const uint32_t o2 = (uint32_t)j * ES * G, o4 = (uint32_t)j * 2u * ES * G;
                        g1[j] = ldM(rB + o2 - ES); g2[j] = ldM(rB + o2 + ES);
                        ii[j] = ldM(rE + o4 - 2u * ES); dd[j] = ldM(rE + o4 + 3u * ES);  // I of cell k-1, D of cell k+1
                        sb[j] = ldM(rA + o2);
                    }
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        const int t = max(g1[j], ii[j]) + 1;
                        const int ins = t == NULLV + 1 ? NULLV : t;  // both NULL -> NULL (wfa.c:249-252)
                        const int del = max(g2[j], dd[j]);
                        mm[j] = max(max(del, sb[j] + 1), max(ins, floor_m));
                        idd[j] = NB ? ((uint32_t)ins | ((uint32_t)del << 8)) : (((uint32_t)ins & 0xffffu) | ((uint32_t)del << 16));
                    }
#pragma unroll
                    for (int j = 0; j < N; ++j) { if (NB) sts_u16(rN + (uint32_t)j * 2u * G, (int)idd[j]); else sts_u32(rN + (uint32_t)(j * 4 * G), idd[j]); }
#pragma unroll
                    for (int j = 0; j < N; ++j) cc[j] = extend_first(aP, aT, k + j * G, mm[j] - BZ, pl, tl, &ll[j]);
#pragma unroll
                    for (int j = 0; j < N; ++j) back(k + j * G, j, mm[j], idd[j], cc[j], ll[j]);
                    rB += ES * N * G; rA += ES * N * G; rNM += ES * N * G; rE += 2 * ES * N * G; rN += 2 * ES * N * G;
                    if (BT) { if (NB) hp1 += N * G; else hp += N * G; }
                    k += N * G;
                };
#ifndef AIM_CPT
#define AIM_CPT 2
#endif
                // (measured and rejected, round 2: warp-wide two-cell trips with the cell a lane does not have masked off, so that lanes
                // with an odd count do not cost the warp an extra one-cell trip - the votes, selects and predicates cost more than the
                // trips they save: 3.41e8 against 3.48e8 pairs/s)
                while (k + (AIM_CPT - 1) * G <= hi) trip(std::integral_constant<int, AIM_CPT>{});
                if (AIM_CPT > 2 && k + G <= hi) trip(std::integral_constant<int, 2>{});
                if (k <= hi) trip(std::integral_constant<int, 1>{});
            }
            __syncwarp();
            // ---- end reached (wfa.c:217-237).  Trimming never removes diagonal ak, so testing before the
            // reduction is equivalent, and the finishing wavefront's trimmed range is never read again. ----
            if (!done && in_range(ak, lo, hi) && ldM(aNM + (uint32_t)ak * ES) >= tl + BZ) { done = true; reached = true; fscore = s; }
            if (__all_sync(kFull, done)) break;

            // ---- adaptive reduction (wfa.c:70-141) on the pairs still running ----
            // The reference scans up from lo and down from hi and stops at the first diagonal within 50 of the minimum
            // distance (or at its limit).  Here every lane scans its OWN cells (stride G) from both ends; the first
            // stopping diagonal overall is the min (max) of the lanes' first stopping cells.
            if (REDUCE) {
                const bool wide = !done && (hi - lo + 1) >= 10;
                int newlo = lo, newhi = hi;
                if (__any_sync(kFull, wide)) {
                    md = group_min<G>(md) + 50;  // keep while distance <= md
                    const int top_limit = min(ak - 1, hi);
                    const int kf = lo + sl;
                    const int kl = hi - ((hi - kf) & (G - 1));  // this lane's last cell (when kf <= hi)
                    int kb = kf;
                    if (wide)
                        while (kb < top_limit && (max(pl + kb, tl) - lit(ldM(aNM + (uint32_t)kb * ES))) > md) kb += G;
                    kb = group_min<G>(kb);
                    if (wide) newlo = max(lo, min(kb, top_limit));
                    const int bottom_limit = max(ak + 1, newlo);
                    int kt = kf <= hi ? kl : INT_MIN;
                    if (wide)
                        while (kt > bottom_limit && (max(pl + kt, tl) - lit(ldM(aNM + (uint32_t)kt * ES))) > md) kt -= G;
                    kt = group_max<G>(kt);
                    if (wide) {
                        newhi = min(hi, max(kt, bottom_limit));
                        // keep the frame: the cells the trim cut off read as NULL from now on
                        for (int k = kf; k < newlo; k += G) { stM(aNM + (uint32_t)k * ES, NULLV); stID_null(aN + (uint32_t)k * 2u * ES); }
                        if (kf <= hi)
                            for (int k = kl; k > newhi; k -= G) { stM(aNM + (uint32_t)k * ES, NULLV); stID_null(aN + (uint32_t)k * 2u * ES); }
                    }
                }
                if (sl == 0 && !done) sts_u32(aDyn + (uint32_t)s * 4u, ((uint32_t)newlo & 0xffffu) | ((uint32_t)(-newhi) << 16));
            }
            __syncwarp();
        }
        __syncwarp();  // arena stores of the last wavefront must be visible to the backtracing lane

        // ---- backtrace (wfa_backtracing.c:219-375): first lane of every sub-warp, concurrently ----
        const int max_ops = pl + tl;
        int begin_offset = max_ops - 1;
        int status = AIM_STATUS_OK;
        // history cell c of this pair slot, as the reference's int16 values
        auto hM = [&](uint32_t c) -> int { return NB ? lit((int)(arena1[c] & 0xffu)) : lo16(arena[c].x); };
        auto hI = [&](uint32_t c) -> int { return NB ? lit((int)((arena1[c] >> 8) & 0xffu)) : lo16(arena[c].y); };
        auto hD = [&](uint32_t c) -> int { return NB ? lit((int)((arena1[c] >> 16) & 0xffu)) : hi16s(arena[c].y); };
        if (BT && reached && sl == 0) {
            const int ops_cap = 2 * RS;
            int b = begin_offset;
            int score = fscore, k = ak;
            int offset;
            {
                const uint4 q1 = lds_v4(aPlan + (uint32_t)fscore * (PLAN_WORDS * 4) + 16);
                const uint32_t r = lds_u32(aPlan + (uint32_t)fscore * (PLAN_WORDS * 4) + 4);
                offset = hM(q1.y + (uint32_t)(k - lo16(r)));
            }
            int v = offset - k, h = offset;
            bool valid = (v > 0 && v <= pl && h > 0 && h <= tl);
            int type = 0;  // 0 M, 1 I, 2 D
            bool bad = false;
#define AIM_PUT(ch) do { if (b < 0 || b >= ops_cap) { bad = true; } else { gops[b] = (ch); } --b; } while (0)
            while (v > 0 && h > 0 && score > 0 && !bad) {
                if (!valid) {
                    valid = (v > 0 && v <= pl && h > 0 && h <= tl);
                    if (valid) {
                        if (k < ak) { for (int j = k; j < ak; ++j) AIM_PUT('I'); }
                        else if (k > ak) { for (int j = ak; j < k; ++j) AIM_PUT('D'); }
                    }
                }
                const int s_open = score - OE, s_ext = score - E, s_mis = score - X;
                // records: static layout from the plan, range = trimmed range (adaptive) or static range
                uint32_t go_f = 0, ge_f = 0, mm_f = 0, go_base = 0, ge_base = 0, mm_base = 0;
                int go_lo = 1, go_hi = -1, ge_lo = 1, ge_hi = -1, mm_lo = 1, mm_hi = -1, go_l0 = 0, ge_l0 = 0, mm_l0 = 0;
                if (s_open >= 0) {
                    const uint32_t a = aPlan + (uint32_t)s_open * (PLAN_WORDS * 4);
                    go_f = lds_u32(a);
                    const uint32_t r = lds_u32(a + 4);
                    go_base = lds_u32(a + 20); go_l0 = lo16(r); go_lo = go_l0; go_hi = -hi16s(r);
                    if (REDUCE && (go_f & P_PRESENT)) { const uint32_t w = lds_u32(aDyn + (uint32_t)s_open * 4u); go_lo = lo16(w); go_hi = -hi16s(w); }
                }
                if (s_ext >= 0) {
                    const uint32_t a = aPlan + (uint32_t)s_ext * (PLAN_WORDS * 4);
                    ge_f = lds_u32(a);
                    const uint32_t r = lds_u32(a + 4);
                    ge_base = lds_u32(a + 20); ge_l0 = lo16(r); ge_lo = ge_l0; ge_hi = -hi16s(r);
                    if (REDUCE && (ge_f & P_PRESENT)) { const uint32_t w = lds_u32(aDyn + (uint32_t)s_ext * 4u); ge_lo = lo16(w); ge_hi = -hi16s(w); }
                }
                if (s_mis >= 0) {
                    const uint32_t a = aPlan + (uint32_t)s_mis * (PLAN_WORDS * 4);
                    mm_f = lds_u32(a);
                    const uint32_t r = lds_u32(a + 4);
                    mm_base = lds_u32(a + 20); mm_l0 = lo16(r); mm_lo = mm_l0; mm_hi = -hi16s(r);
                    if (REDUCE && (mm_f & P_PRESENT)) { const uint32_t w = lds_u32(aDyn + (uint32_t)s_mis * 4u); mm_lo = lo16(w); mm_hi = -hi16s(w); }
                }
                int del_ext = kNull, del_open = kNull, ins_ext = kNull, ins_open = kNull, misms = kNull;
                if (type != 1) {
                    if ((ge_f & P_PRESENT) && (ge_f & P_HAS_D) && ge_lo <= k + 1 && k + 1 <= ge_hi)
                        del_ext = hD(ge_base + (uint32_t)(k + 1 - ge_l0));
                    if ((go_f & P_PRESENT) && go_lo <= k + 1 && k + 1 <= go_hi) del_open = hM(go_base + (uint32_t)(k + 1 - go_l0));
                }
                if (type != 2) {
                    if ((ge_f & P_PRESENT) && (ge_f & P_HAS_I) && ge_lo <= k - 1 && k - 1 <= ge_hi)
                        ins_ext = (int16_t)(hI(ge_base + (uint32_t)(k - 1 - ge_l0)) + 1);
                    if ((go_f & P_PRESENT) && go_lo <= k - 1 && k - 1 <= go_hi)
                        ins_open = (int16_t)(hM(go_base + (uint32_t)(k - 1 - go_l0)) + 1);
                }
                if (type == 0) {
                    if ((mm_f & P_PRESENT) && mm_lo <= k && k <= mm_hi) misms = (int16_t)(hM(mm_base + (uint32_t)(k - mm_l0)) + 1);
                }
                const int max_all = max(misms, max(max(ins_ext, ins_open), max(del_ext, del_open)));
                if (type == 0) {
                    const int num_matches = offset - max_all;  // ops are 'M' already
                    if (num_matches > 0) {
                        if (num_matches > b + 1) { bad = true; break; }
                        b -= num_matches;
                    }
                    offset = max_all;
                    v = offset - k;
                    h = offset;
                    if (v <= 0 || h <= 0) break;
                }
                if (max_all == del_ext) { if (valid) AIM_PUT('D'); score = s_ext; ++k; type = 2; }
                else if (max_all == del_open) { if (valid) AIM_PUT('D'); score = s_open; ++k; type = 0; }
                else if (max_all == ins_ext) { if (valid) AIM_PUT('I'); score = s_ext; --k; --offset; type = 1; }
                else if (max_all == ins_open) { if (valid) AIM_PUT('I'); score = s_open; --k; --offset; type = 0; }
                else if (max_all == misms) { if (valid) AIM_PUT('X'); score = s_mis; --offset; }
                else { bad = true; break; }
                v = offset - k;
                h = offset;
            }
            if (!bad) {
                if (score == 0) {
                    if (offset > 0) { if (offset > b + 1) bad = true; else b -= offset; }
                } else {
                    while (v > 0 && !bad) { AIM_PUT('D'); --v; }
                    while (h > 0 && !bad) { AIM_PUT('I'); --h; }
                }
            }
#undef AIM_PUT
            if (bad) status = AIM_STATUS_BACKTRACE;
            begin_offset = b + 1;
        }
        if (active && sl == 0) {
            aim_result r;
            r.max_operations = max_ops;
            r.begin_offset = begin_offset;
            r.end_offset = max_ops;
            r.score = fscore;
            r.status = status;
            r.idx = K.idx_base + i;
            K.results[i] = r;
        }
        __syncwarp();
    }
#undef AIM_ROW
#undef AIM_ROWD
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// Packed entry only: the CIGAR of every pair as edit_cigar_print writes it (WFA/DPU-MRAM/host/host.c:69-89: "%d%c" per run of
// ops[begin_offset .. end_offset), the op at begin_offset always printed), NUL-terminated, in a fixed-pitch row, so that
// ~30 bytes instead of the 2 * READ_SIZE op row cross PCIe.  One pair per thread.  A pair the lockstep kernel skipped
// (non-ACGT byte: no ASCII is on the device in this mode) gets AIM_STATUS_NEEDS_ASCII; a CIGAR longer than the row gets
// AIM_STATUS_CIGAR_OVERFLOW; the caller serves both through aim_align_batch.
// mode 1 (GenASM-DC: the op row already holds the reference's CIGAR string): copy the string into the row.
__global__ void __launch_bounds__(128) cigar_rle_kernel(const int32_t *plen, const int32_t *tlen, const uint32_t *flags,
                                                        aim_result *results, const char *ops, char *cigars, int pitch,
                                                        uint32_t n, uint32_t idx_base, int RS, int mode)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char *out = cigars + (size_t)i * pitch;
    if (mode == 1) {
        const int len = results[i].end_offset;
        const char *row = ops + (size_t)i * 2 * RS;
        if (len + 1 > pitch) { results[i].status = AIM_STATUS_CIGAR_OVERFLOW; out[0] = '\0'; return; }
        for (int j = 0; j < len; ++j) out[j] = row[j];
        out[len] = '\0';
        return;
    }
    if (flags && ((flags[i >> 5] >> (i & 31)) & 1u)) {
        aim_result r;
        r.max_operations = min(max(plen[i], 0), RS) + min(max(tlen[i], 0), RS);
        r.begin_offset = r.max_operations - 1;
        r.end_offset = r.max_operations;
        r.score = 0;
        r.status = AIM_STATUS_NEEDS_ASCII;
        r.idx = idx_base + i;
        results[i] = r;
        out[0] = '\0';
        return;
    }
    const aim_result r = results[i];
    if (r.status != AIM_STATUS_OK) { out[0] = '\0'; return; }
    const char *row = ops + (size_t)i * 2 * RS;
    const int b = r.begin_offset, e = r.end_offset > b ? r.end_offset : b + 1;
    int pos = 0;
    bool over = false;
    auto emit = [&](int len, char op) {
        char tmp[10];
        int k = 0;
        do { tmp[k++] = (char)('0' + len % 10); len /= 10; } while (len);
        if (pos + k + 1 >= pitch) { over = true; return; }
        while (k) out[pos++] = tmp[--k];
        out[pos++] = op;
    };
    if (b < 0) {  // an empty alignment has b = -1 and the reference prints the byte before its span: an 'M' of the memset (wfa.c:499-501)
        emit(1, 'M');
    } else {
        // run boundaries eight ops at a time: byte k of (x ^ x shifted down one op) is non-zero where op 8w+k differs from op 8w+k+1
        const unsigned long long *row64 = reinterpret_cast<const unsigned long long *>(row);
        const int w0 = b >> 3, w1 = (e - 1) >> 3;
        int run_start = b;
        unsigned long long x = row64[w0];
        for (int w = w0; w <= w1 && !over; ++w) {
            const unsigned long long nx = w < w1 ? row64[w + 1] : x;
            unsigned long long diff = x ^ ((x >> 8) | (nx << 56));
            // only boundaries between ops j and j + 1 with b <= j and j + 1 <= e - 1 count
            const int lo_k = max(b - 8 * w, 0), hi_k = min(e - 1 - 8 * w, 8);  // bytes [lo_k, hi_k)
            if (lo_k > 0) diff &= ~0ull << (8 * lo_k);
            if (hi_k < 8) diff &= hi_k > 0 ? ~(~0ull << (8 * hi_k)) : 0ull;
            while (diff && !over) {
                const int k = (__ffsll((long long)diff) - 1) >> 3;
                const int j = 8 * w + k;
                emit(j - run_start + 1, (char)((x >> (8 * k)) & 0xffull));
                run_start = j + 1;
                diff &= ~(0xffull << (8 * k));
            }
            x = nx;
        }
        if (!over) emit(e - run_start, row[e - 1]);
    }
    if (over) { results[i].status = AIM_STATUS_CIGAR_OVERFLOW; out[0] = '\0'; }
    else out[pos] = '\0';
}

template <int G>
cudaError_t launch_g(const SubK &K, bool reduce, bool bt, bool narrow, bool pool, int grid, int block, size_t smem, cudaStream_t st)
{
    cudaError_t e;
    // blocks of more than 4 warps (G <= 4 only: one big block fills an SM's shared memory with less per-block overhead)
    // use a second instantiation whose register budget allows them; the narrow-row variant exists for G = 4
#define AIM_LAUNCH(R, B)                                                                                                  \
    do {                                                                                                                  \
        if (narrow) {                                                                                                     \
            if constexpr (G == 4) {                                                                                       \
                if (pool) {                                                                                               \
                    e = cudaFuncSetAttribute(wfa_sub_kernel<G, R, B, 896, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                    if (e == cudaSuccess) wfa_sub_kernel<G, R, B, 896, true, true><<<grid, block, smem, st>>>(K);         \
                } else if (block <= 128) {                                                                                       \
                    e = cudaFuncSetAttribute(wfa_sub_kernel<G, R, B, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                    if (e == cudaSuccess) wfa_sub_kernel<G, R, B, 128, true><<<grid, block, smem, st>>>(K);               \
                } else {                                                                                                  \
                    e = cudaFuncSetAttribute(wfa_sub_kernel<G, R, B, 1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                    if (e == cudaSuccess) wfa_sub_kernel<G, R, B, 1024, true><<<grid, block, smem, st>>>(K);              \
                }                                                                                                         \
            } else {                                                                                                      \
                e = cudaErrorInvalidConfiguration;                                                                        \
            }                                                                                                             \
        } else if (block <= 128) {                                                                                        \
            e = cudaFuncSetAttribute(wfa_sub_kernel<G, R, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (e == cudaSuccess) wfa_sub_kernel<G, R, B><<<grid, block, smem, st>>>(K);                                   \
        } else if constexpr (G <= 4) {                                                                                    \
            e = cudaFuncSetAttribute(wfa_sub_kernel<G, R, B, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e == cudaSuccess) wfa_sub_kernel<G, R, B, 768><<<grid, block, smem, st>>>(K);                              \
        } else {                                                                                                          \
            e = cudaErrorInvalidConfiguration;                                                                            \
        }                                                                                                                 \
    } while (0)
    if (reduce && bt) AIM_LAUNCH(true, true);
    else if (reduce) AIM_LAUNCH(true, false);
    else if (bt) AIM_LAUNCH(false, true);
    else AIM_LAUNCH(false, false);
#undef AIM_LAUNCH
    return e;
}

}  // namespace

// CIGAR rows from the op rows of ANY algorithm (aim_align_batch_cigars): the text edit_cigar_print would print, or a copy of
// GenASM-DC's string; enqueued after the alignment kernels on the same stream.
int launch_cigar_rows(const KernelArgs &a, void *stream_v, int *launches)
{
    if (a.n == 0) return AIM_OK;
    cigar_rle_kernel<<<(a.n + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(a.plen, a.tlen, nullptr, a.results, a.ops, a.cigars, a.cigar_pitch, a.n,
                                                                          a.idx_base, a.p.read_size, a.p.algo == AIM_ALGO_GENASM_DC ? 1 : 0);
    if (cudaGetLastError() != cudaSuccess) { set_error("cigar rows launch failed"); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    return AIM_OK;
}

// Returns AIM_OK after enqueueing, 1 if this configuration is not served by the lockstep kernel
// (long reads / very large MAX_SCORE: the rings or the packed sequences do not fit), or an AIM_ERR_*.
int launch_wfa_sub(const KernelArgs &a, Scratch *sc, void *stream_v, int *launches)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    const aim_params &p = a.p;
    const int MS = p.max_score;
    if (MS > 2000) return 1;
    const char *mode = getenv("AIM_WFA_MODE");
    if (mode && std::string(mode) == "warp") return 1;

    // ---- static schedule (the ranges wfa.c:275-354 would derive without trimming) ----
    const int x = p.mismatch, o = p.gap_open, e = p.gap_ext;
    struct S { bool present, has_i, has_d; int lo, hi; };
    std::vector<S> w((size_t)MS + 1);
    w[0] = {true, false, false, 0, 0};
    int kmin = 0, kmax = 0;
    for (int s = 1; s <= MS; ++s) {
        const S *A = (s - x >= 0 && w[s - x].present) ? &w[s - x] : nullptr;
        const S *B = (s - o - e >= 0 && w[s - o - e].present) ? &w[s - o - e] : nullptr;
        const S *E = (s - e >= 0 && w[s - e].present) ? &w[s - e] : nullptr;
        const bool ie_null = !(E && E->has_i), de_null = !(E && E->has_d);
        const bool i_out_null = !B && ie_null, d_out_null = !B && de_null;
        if (!A && i_out_null && d_out_null) { w[s] = {false, false, false, 1, -1}; continue; }
        const int lo = std::min(std::min(A ? A->lo : 1, B ? B->lo : 1), (ie_null && de_null) ? 1 : E->lo) - 1;
        const int hi = std::max(std::max(A ? A->hi : -1, B ? B->hi : -1), (ie_null && de_null) ? -1 : E->hi) + 1;
        w[s] = {true, !i_out_null, !d_out_null, lo, hi};
        kmin = std::min(kmin, lo);
        kmax = std::max(kmax, hi);
    }
    const uint32_t ring_m = (uint32_t)std::max(x, o + e) + 1, ring_e = (uint32_t)e + 1;
    SubK K{};
    // narrow rows (one byte per offset, see the kernel): every offset the reference can produce, plus the bias, fits a byte
    bool narrow = p.read_size + 2 * MS + 12 <= 255 && MS <= 40 && !getenv("AIM_WFA_G");  // (MS <= 40: G = 4, whose instantiations carry the narrow variant)
    if (const char *e8 = getenv("AIM_WFA_NARROW")) narrow = narrow && atoi(e8) != 0;
    K.bias = narrow ? MS + 12 : 0;
    K.null_max = narrow ? MS : 0;
    const uint32_t es = narrow ? 1u : 2u;
    K.cw = round_up((uint32_t)(kmax - kmin + 3), narrow ? 16 : 8);  // one frame cell on both sides; rows are 16-byte multiples
    K.koff = 1 - kmin;
    const uint32_t row_bytes = K.cw * es;
    const uint32_t rows_bytes = (ring_m + 2 * ring_e) * row_bytes;
    if (rows_bytes > 0xfff0u) return 1;
    K.rows_v4 = rows_bytes / 16;
    K.plan_words = (uint32_t)PLAN_WORDS * ((uint32_t)MS + 1);
    std::vector<uint32_t> plan(K.plan_words, 0u);
    uint64_t arena_cells = 0;
    // rows: M ring, then the {I,D} ring
    auto m_off = [&](int s) -> uint32_t { return ((uint32_t)s % ring_m) * row_bytes; };
    auto id_off = [&](int s) -> uint32_t { return (ring_m + 2 * ((uint32_t)s % ring_e)) * row_bytes; };
    std::vector<uint32_t> last_m(ring_m, OFF_NULL), last_id(ring_e, OFF_NULL);  // score whose cells sit in each row
    for (int s = 0; s <= MS; ++s) {
        if (!w[s].present) continue;
        uint32_t *q = &plan[(size_t)s * PLAN_WORDS];
        const bool A = s - x >= 0 && w[s - x].present, B = s - o - e >= 0 && w[s - o - e].present, E = s - e >= 0 && w[s - e].present;
        const bool ie_null = !(E && w[s - e].has_i), de_null = !(E && w[s - e].has_d);
        const uint32_t width = (uint32_t)(w[s].hi - w[s].lo + 1);
        // the reference's -10 for a missing I / D / sub candidate (wfa.c:243,255,266) is a floor under the max; score 0 starts at offset 0
        const int floor_m = s == 0 ? 0 : ((!A || !w[s].has_i || !w[s].has_d) ? -10 : -32768);
        q[0] = P_PRESENT | (A ? 0u : P_SUB_NULL) | (B ? 0u : P_O_NULL) | (ie_null ? P_IE_NULL : 0u) | (de_null ? P_DE_NULL : 0u) |
               (w[s].has_i ? P_HAS_I : 0u) | (w[s].has_d ? P_HAS_D : 0u) | ((!A || !B || !E) ? P_NULL_ROW : 0u) |
               (((uint32_t)floor_m & 0xffffu) << 16);
        q[1] = ((uint32_t)w[s].lo & 0xffffu) | ((uint32_t)(-w[s].hi) << 16);
        q[2] = m_off(s) | ((A ? m_off(s - x) : OFF_NULL) << 16);
        q[3] = (B ? m_off(s - o - e) : OFF_NULL) | ((E ? id_off(s - e) : OFF_NULL) << 16);
        q[4] = id_off(s);
        q[5] = (uint32_t)arena_cells;
        q[6] = last_m[(uint32_t)s % ring_m] << 16;
        q[7] = last_id[(uint32_t)s % ring_e];
        last_m[(uint32_t)s % ring_m] = (uint32_t)s;
        last_id[(uint32_t)s % ring_e] = (uint32_t)s;
        arena_cells += width;
    }
    if (arena_cells > 0x0fffffffu) return 1;

    const bool prepacked = a.packed != nullptr;
    K.plen = a.plen; K.tlen = a.tlen; K.patterns = a.patterns; K.texts = a.texts;
    K.results = a.results; K.ops = a.ops; K.n = a.n; K.idx_base = a.idx_base;
    K.x = x; K.o = o; K.e = e; K.max_score = MS; K.read_size = p.read_size;
    K.seq_words = round_up((uint32_t)p.read_size / 16 + 2, 4);
    if (K.seq_words > 256) return 1;  // (wfa_prep_kernel: a block covers whole rows; such reads are the long-read kernel's anyway)
    K.dyn_words = p.reduce ? round_up((uint32_t)MS + 1, 4) : 0;
    const uint32_t pair_words_raw = 2 * K.seq_words + K.dyn_words + rows_bytes / 4;

    // lanes per pair: wavefronts are about MAX_SCORE diagonals wide on average; a few iterations per score keeps the
    // lanes busy while the per-score bookkeeping is shared by 32/G pairs; env override for tuning
    int G = MS <= 40 ? 4 : MS <= 100 ? 8 : MS <= 300 ? 16 : 32;
    if (const char *gs = getenv("AIM_WFA_G")) { int g = atoi(gs); if (g == 2 || g == 4 || g == 8 || g == 16 || g == 32) G = g; }
    const int PPW = 32 / G;
    {   // stagger the pair slots of one warp over the banks with the least padding: a slot stride that is an ODD multiple of
        // 32/PPW words puts the PPW slots on distinct bank groups; slots stay 16-byte aligned
        K.pair_words = (pair_words_raw + 3u) & ~3u;
        if (PPW > 1 && !getenv("AIM_WFA_NOSTAGGER")) {
            const uint32_t unit = std::max(4u, 32u / (uint32_t)PPW);
            while (K.pair_words % (2 * unit) != unit) K.pair_words += 4;
        }
    }
    const uint32_t kSmemBudget = 227u * 1024u, kSmemPerSm = 228u * 1024u, kBlockReserve = 1024u;
    const size_t pair_bytes = (size_t)K.pair_words * 4;
    const size_t fixed_bytes = (size_t)K.plan_words * 4 + (size_t)K.cw * 4;  // plan + the all-NULL row (K.cw words: an int16 {I,D} row)
    // G = 4 and a large per-pair footprint: ONE block per SM with as many warps as its shared memory holds (20 at config 4
    // against 16 as two-warp blocks) - the per-block plan copy and reserve are paid once, and the kernel, latency-bound at
    // these occupancies, gains ~10 %.  Small footprints keep the two-warp blocks (registers cap those at 32 warps/SM).
    const int default_wpb = G <= 4 ? 2 : 4;
    int warps_per_block = default_wpb;
    // POOL (see the kernel): a warp's cells dealt to all its lanes; one table of PPW * cw 16-bit entries per warp.  Bit-exact, but
    // measured SLOWER at config 4 (2.83e8 against 3.48e8 pairs/s: 23 instead of 18 lanes busy and 37 % fewer cell rounds, but a
    // round costs ~108 warp instructions instead of ~50 - the per-pair loop's running pointers and register-resident pair state
    // are rebuilt per cell): off unless AIM_WFA_POOL=1
    bool pool = false;
    if (const char *ps = getenv("AIM_WFA_POOL")) pool = narrow && G == 4 && atoi(ps) != 0;
    K.own_bytes = pool ? round_up((uint32_t)PPW * K.cw * 2u, 16) : 0u;
    const size_t warp_bytes = (size_t)PPW * pair_bytes + K.own_bytes;
    if (G <= 4) {
        const size_t small_block = fixed_bytes + (size_t)default_wpb * warp_bytes;
        const int small_warps = (int)std::min<size_t>(32, kSmemPerSm / (small_block + kBlockReserve) * default_wpb);
        int big_warps = (int)std::min<size_t>(pool ? 28 : narrow ? 32 : 24, (kSmemBudget - fixed_bytes) / warp_bytes);
        if (big_warps >= 8 && G == 4) big_warps &= ~3;  // the same number of warps on each of the SM's four schedulers (20 beats 21: 300 vs 290 M pairs/s)
        if (big_warps > small_warps || pool) warps_per_block = big_warps;
    }
    if (const char *ws = getenv("AIM_WFA_WPB")) { int v = atoi(ws); if (v >= 1 && v <= (G <= 4 ? (pool ? 28 : narrow ? 32 : 24) : 4)) warps_per_block = v; }
    if (warps_per_block < 1) return 1;
    size_t smem_block = fixed_bytes + (size_t)warps_per_block * warp_bytes;
    if (smem_block > kSmemBudget) return 1;
    if (fixed_bytes + (size_t)default_wpb * warp_bytes > kSmemBudget / 3) return 1;  // too few warps/SM would fit: leave it to the long-read kernel
    int blocks_per_sm = (int)std::min<uint32_t>(kSmemPerSm / ((uint32_t)smem_block + kBlockReserve), 32u);
    int max_warps = 48;
    if (const char *ws = getenv("AIM_WFA_MAXWARPS")) { int v = atoi(ws); if (v >= 4 && v <= 64) max_warps = v; }
    blocks_per_sm = std::max(1, std::min(blocks_per_sm, max_warps / warps_per_block));
    int grid = sc->sm_count * blocks_per_sm;
    {
        const uint64_t per_block = (uint64_t)warps_per_block * PPW;
        if ((uint64_t)grid * per_block > a.n) grid = (int)((a.n + per_block - 1) / per_block);
    }
    const uint64_t total_slots = (uint64_t)grid * warps_per_block * PPW;

    // pairs the prep kernel flags (non-ACGT bytes) are served by the warp-per-pair kernel from a device-side list
    const WarpPlan W = wfa_warp_plan(a, sc->sm_count, std::min<uint32_t>(a.n, (uint32_t)sc->sm_count * 8u));
    if (W.rc != AIM_OK) return 1;

    // device scratch: plan | list count + non-ACGT flag bits | hand-over list | packed windows | (BT) history arena | warp-kernel scratch
    auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t plan_bytes = (size_t)K.plan_words * 4;
    const size_t plan_dev = up256(plan_bytes);
    const size_t flags_dev = up256(256 + ((size_t)a.n + 31) / 32 * 4);
    const size_t list_dev = up256((size_t)a.n * 4);
    const size_t packed_dev = up256((size_t)a.n * 2 * K.seq_words * 4);
    K.arena_stride = p.backtrace ? (size_t)round_up((uint32_t)arena_cells, 16) : 0;
    const size_t arena_bytes = up256((size_t)total_slots * K.arena_stride * (narrow ? 4 : 8));
    int rc = scratch_reserve(sc, plan_dev + flags_dev + list_dev + packed_dev + arena_bytes + W.scratch_bytes);
    if (rc != AIM_OK) return rc;
    unsigned char *base = reinterpret_cast<unsigned char *>(sc->buf);
    uint32_t *list_count = reinterpret_cast<uint32_t *>(base + plan_dev);
    uint32_t *flags = list_count + 64;
    uint32_t *list = reinterpret_cast<uint32_t *>(base + plan_dev + flags_dev);
    uint32_t *packed = reinterpret_cast<uint32_t *>(base + plan_dev + flags_dev + list_dev);
    K.plan = reinterpret_cast<const uint32_t *>(cached_plan(sc, plan.data(), plan_bytes));  // uploaded once per penalty set
    if (!K.plan) return AIM_ERR_CUDA;
    K.flags = prepacked ? a.pflags : flags;
    K.packed = reinterpret_cast<const uint4 *>(prepacked ? a.packed : packed);
    if (prepacked && aim_packed_row_bytes(p.read_size) != (int32_t)(K.seq_words * 4)) { set_error("packed row pitch mismatch"); return AIM_ERR_ARG; }
    K.arena = reinterpret_cast<uint2 *>(base + plan_dev + flags_dev + list_dev + packed_dev);
    void *warp_scratch = base + plan_dev + flags_dev + list_dev + packed_dev + arena_bytes;
    cudaError_t err = cudaMemsetAsync(list_count, 0, flags_dev, stream);
    if (err == cudaSuccess && p.backtrace)  // op rows: 'M' everywhere (wfa.c:499-501); the backtrace overwrites the few edits
        err = cudaMemsetAsync(a.ops, 'M', (size_t)a.n * 2 * (size_t)p.read_size, stream);
    if (err == cudaSuccess && !prepacked) {
        const uint64_t rpb = 256u / K.seq_words;  // (seq_words <= 72: read_size is below the long-read threshold here)
        const int pgrid = (int)std::min<uint64_t>((2ull * a.n + 2 * rpb - 1) / (2 * rpb), (uint64_t)sc->sm_count * 64);
        wfa_prep_kernel<<<pgrid, 256, 0, stream>>>(a.patterns, a.texts, a.plen, a.tlen, a.n, p.read_size, K.seq_words, packed, flags, list, list_count);
        err = cudaGetLastError();
        if (launches) ++*launches;
    }
    if (err == cudaSuccess) {
        const int block = warps_per_block * 32;
        if (G == 2) err = launch_g<2>(K, p.reduce != 0, p.backtrace != 0, false, false, grid, block, smem_block, stream);
        else if (G == 4) err = launch_g<4>(K, p.reduce != 0, p.backtrace != 0, narrow, pool, grid, block, smem_block, stream);
        else if (G == 8) err = launch_g<8>(K, p.reduce != 0, p.backtrace != 0, false, false, grid, block, smem_block, stream);
        else if (G == 16) err = launch_g<16>(K, p.reduce != 0, p.backtrace != 0, false, false, grid, block, smem_block, stream);
        else err = launch_g<32>(K, p.reduce != 0, p.backtrace != 0, false, false, grid, block, smem_block, stream);
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("wfa_sub launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    if (prepacked) {  // no ASCII on the device: flagged pairs are reported, and the CIGARs leave run-length encoded
        if (a.cigars) {
            cigar_rle_kernel<<<(a.n + 127) / 128, 128, 0, stream>>>(a.plen, a.tlen, a.pflags, a.results, a.ops, a.cigars, a.cigar_pitch,
                                                                  a.n, a.idx_base, p.read_size, 0);
            if (cudaGetLastError() != cudaSuccess) { set_error("cigar_rle launch failed"); return AIM_ERR_CUDA; }
            if (launches) ++*launches;
        }
        return AIM_OK;
    }
    return wfa_warp_launch(W, warp_scratch, list, list_count, stream_v, launches);
}

}  // namespace aim
