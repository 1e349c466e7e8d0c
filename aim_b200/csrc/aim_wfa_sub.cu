// Short-read WFA / WFA-adaptive for sm_100a: G lanes per pair, 32/G pairs per warp in LOCKSTEP.
//
// Same algorithm and literal semantics as aim_wfa.cu (reference: WFA/DPU-MRAM/dpu/wfa.c:70-407,
// wfa_backtracing.c:219-375), restructured around one observation: without adaptive trimming the
// wavefront SCHEDULE of gap-affine WFA - which scores have a wavefront, its [lo,hi], whether it has
// I/D components - depends only on the penalties, not on the sequences (wfa.c:275-354 derives it from
// the ranges of scores s-x, s-o-e, s-e).  So
//   * the schedule is computed once on the host into a per-score PLAN (8 words) shared by every pair;
//     adaptive trimming (wfa.c:70-141) only narrows [lo,hi] per pair;
//   * all pairs of a warp walk the scores together: the loop structure is warp-uniform, the lanes of a
//     sub-warp split the diagonals of their pair, and the only divergence left is the data-dependent
//     extend length and the per-pair finishing score;
//   * only the last max(x, o+e)+1 wavefronts are ever read by compute_offsets, so they live in a
//     shared-memory RING whose arrays share one origin and hold NULL outside the occupant's [lo,hi]:
//     the reference's range-checked fetch (AFFINE_WAVEFRONT_COND_FETCH, common.h:121-124) becomes a
//     plain load.  (`sub` keeps its check: the reference adds 1 to an in-range NULL but not to an
//     out-of-range one, wfa.c:266.)
//   * the full history the backtrace needs (the reference's MRAM component store) is streamed to a
//     per-pair-slot HBM arena with coalesced 16-bit stores and read back only along the backtrace
//     path; the arena is reused pair after pair and stays L2-resident;
//   * compute_offsets and extend are fused per diagonal (extending diagonal k touches only M[k]);
//   * backtraces of the 32/G pairs run concurrently on the sub-warps' first lanes.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "aim_internal.h"
#include "aim_wfa_common.cuh"

namespace aim {

namespace {

// plan flags
constexpr uint32_t P_PRESENT = 1, P_SUB_NULL = 2, P_O_NULL = 4, P_IE_NULL = 8, P_DE_NULL = 16, P_HAS_I = 32, P_HAS_D = 64;
constexpr int PLAN_WORDS = 8;
// per-score plan: w0 span_lo(int16) | span_len<<16   loop range: union of this wavefront's static range and the
//                                                     static range of the ring slot's previous occupant
//                 w1 flags
//                 w2 ring halfword offset of this score's slot | of score s-x's slot << 16
//                 w3 ring halfword offset of s-o-e's slot      | of s-e's slot << 16
//                 w4 arena slot of M[lo_s] (history for the backtrace)
//                 w5 lo_s (int16) | width_s << 16              static range
//                 w6 static a_lo (int16) | a_hi << 16          range of score s-x (non-adaptive runs)
//                 w7 unused

struct SubK {
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    const uint32_t *plan;  // device copy of the plan
    int16_t *arena;        // history arena (BT only)
    size_t arena_stride;   // int16 slots per pair slot
    uint32_t n, idx_base;
    int x, o, e;
    int max_score, read_size;
    int koff;              // ring cell of diagonal k is k + koff
    uint32_t plan_words;   // PLAN_WORDS * (max_score + 1)
    uint32_t seq_words;    // words per packed sequence (multiple of 4)
    uint32_t dyn_words;    // trimmed-range words per pair (multiple of 4; 0 when !reduce)
    uint32_t cw;           // ring array width in halfwords (even)
    uint32_t ring_halfs;   // ring size in halfwords (multiple of 8)
    uint32_t pair_words;   // shared-memory words per pair slot
};

// ---- shared-memory accessors on 32-bit shared-window addresses (no generic-pointer conversion) ----
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((short)v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "h"((short)v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// equal bases from pattern[v], text[h], at most lim (> 0); sequences 2-bit packed at shared addresses aP/aT
__device__ __forceinline__ int match_packed_s(uint32_t aP, uint32_t aT, int v, int h, int lim)
{
    int cnt = 0;
    for (;;) {
        const int pv = v + cnt, ph = h + cnt;
        const uint32_t pa = aP + ((pv >> 4) << 2), ta = aT + ((ph >> 4) << 2);
        const uint32_t a = __funnelshift_l(lds_u32(pa + 4), lds_u32(pa), (pv & 15) * 2);
        const uint32_t b = __funnelshift_l(lds_u32(ta + 4), lds_u32(ta), (ph & 15) * 2);
        const uint32_t d = a ^ b;
        if (d) { cnt += __clz(d) >> 1; break; }
        cnt += 16;
        if (cnt >= lim) break;
    }
    return min(cnt, lim);
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

template <int G, bool REDUCE, bool BT>
__global__ void __launch_bounds__(128) wfa_sub_kernel(const SubK K)
{
    constexpr int PPW = 32 / G;
    constexpr uint32_t GM = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    extern __shared__ __align__(16) uint32_t smem_w[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / G, sl = lane % G, subshift = sub * G;
    const uint32_t wpb = blockDim.x >> 5;
    const int RS = K.read_size, MS = K.max_score;
    const int X = K.x, OE = K.o + K.e, E = K.e;

    // block-wide: the plan
    for (uint32_t j = threadIdx.x; j < K.plan_words; j += blockDim.x) smem_w[j] = K.plan[j];
    __syncthreads();

    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_w);
    const uint32_t aPlan = sbase;
    const uint32_t aSlot = sbase + (K.plan_words + (uint32_t)(wib * PPW + sub) * K.pair_words) * 4u;
    const uint32_t aP = aSlot;
    const uint32_t aT = aP + K.seq_words * 4u;
    const uint32_t aDyn = aT + K.seq_words * 4u;
    const uint32_t aRing = aDyn + K.dyn_words * 4u;
    const uint32_t aOps = aRing + K.ring_halfs * 2u;
    const uint32_t CW2 = K.cw * 2u;  // bytes between the M, I and D arrays of a ring slot

    const uint32_t slot_global = (blockIdx.x * wpb + wib) * PPW + sub;
    const uint32_t nslots = gridDim.x * wpb * PPW;
    int16_t *arena = BT ? K.arena + (size_t)slot_global * K.arena_stride : nullptr;

    for (uint32_t base_i = 0; base_i < K.n; base_i += nslots) {  // warp-uniform trip count
        const uint32_t i = base_i + slot_global;
        const bool active = i < K.n;
        const int pl = active ? min(max(K.plen[i], 0), RS) : 0;
        const int tl = active ? min(max(K.tlen[i], 0), RS) : 0;
        const char *gp = K.patterns + (size_t)(active ? i : 0) * RS;
        const char *gt = K.texts + (size_t)(active ? i : 0) * RS;
        const int ak = tl - pl;

        // ---- stage + 2-bit pack (8-byte loads, contiguous inside a sub-warp) ----
        bool ok = true;
        for (int c = sl; c * 8 < pl; c += G) {
            const uint2 w = __ldg(reinterpret_cast<const uint2 *>(gp) + c);
            sts_u16(aP + (uint32_t)(c ^ 1) * 2u, (int)pack8(w, pl - c * 8, &ok));
        }
        for (int c = sl; c * 8 < tl; c += G) {
            const uint2 w = __ldg(reinterpret_cast<const uint2 *>(gt) + c);
            sts_u16(aT + (uint32_t)(c ^ 1) * 2u, (int)pack8(w, tl - c * 8, &ok));
        }
        const bool packed = ((__ballot_sync(kFull, ok) >> subshift) & GM) == GM;
        // ring: NULL everywhere (a cell is either inside its occupant's [lo,hi] or NULL)
        for (uint32_t c = sl; c < K.ring_halfs / 2; c += G) sts_u32(aRing + c * 4u, 0xc000c000u);
        if (BT) {
            for (int c = sl; c < (2 * RS) / 4; c += G) sts_u32(aOps + (uint32_t)c * 4u, 0x4d4d4d4du);  // 'M' (wfa.c:499-501)
        }
        __syncwarp();

        bool done = !active;
        bool reached = false;
        int fscore = MS + 1;  // give-up value (wfa.c:399-404)

        for (int s = 0; s <= MS; ++s) {
            const uint4 p0 = lds_v4(aPlan + (uint32_t)s * (PLAN_WORDS * 4));  // warp-uniform
            const uint32_t fl = p0.y;
            if (!(fl & P_PRESENT)) continue;
            const uint4 p1 = lds_v4(aPlan + (uint32_t)s * (PLAN_WORDS * 4) + 16);
            const int span_lo = lo16(p0.x), span_len = (int)(p0.x >> 16);
            const bool sub_null = fl & P_SUB_NULL, o_null = fl & P_O_NULL, ie_null = fl & P_IE_NULL, de_null = fl & P_DE_NULL;
            const bool has_i = fl & P_HAS_I, has_d = fl & P_HAS_D;
            const int lo_s = lo16(p1.y), wlen_s = (int)(p1.y >> 16);

            // this pair's range (wfa.c:318-343) and the range of score s-x for the `sub` check
            int lo = lo_s, hi = lo_s + wlen_s - 1, a_lo = lo16(p1.z), a_hi = hi16s(p1.z);
            if (REDUCE && s > 0) {
                int b_lo = 1, b_hi = -1, e_lo = 1, e_hi = -1;
                a_lo = 1; a_hi = -1;
                if (!sub_null) { const uint32_t w = lds_u32(aDyn + (uint32_t)(s - X) * 4u); a_lo = lo16(w); a_hi = hi16s(w); }
                if (!o_null) { const uint32_t w = lds_u32(aDyn + (uint32_t)(s - OE) * 4u); b_lo = lo16(w); b_hi = hi16s(w); }
                if (!(ie_null && de_null)) { const uint32_t w = lds_u32(aDyn + (uint32_t)(s - E) * 4u); e_lo = lo16(w); e_hi = hi16s(w); }
                lo = min(min(a_lo, b_lo), e_lo) - 1;
                hi = max(max(a_hi, b_hi), e_hi) + 1;
            }
            // ring addresses of cell k = base + 2*k  (koff folded in)
            const uint32_t rbase = aRing + (uint32_t)(K.koff * 2);
            const uint32_t aNM = rbase + (p0.z & 0xffffu) * 2u;
            const uint32_t aAM = rbase + (p0.z >> 16) * 2u;
            const uint32_t aBM = rbase + (p0.w & 0xffffu) * 2u;
            const uint32_t aEI = rbase + (p0.w >> 16) * 2u + CW2;
            const uint32_t aED = aEI + CW2;
            int16_t *hM = BT ? arena + p1.x - lo_s : nullptr;  // arena copy, indexed by k
            int16_t *hI = BT ? hM + wlen_s : nullptr;
            int16_t *hD = BT ? hM + (has_i ? 2 * wlen_s : wlen_s) : nullptr;

            // ---- compute_offsets (wfa.c:238-273) fused with extend (wfa.c:193-215) ----
            int md = max(pl, tl);
            bool hit_end = false;
            for (int k = span_lo + sl; k < span_lo + span_len; k += G) {
                if (done) continue;
                const uint32_t k2 = (uint32_t)(k * 2);
                if (k < lo || k > hi) {  // outside this pair's wavefront: keep the ring NULL there
                    sts_u16(aNM + k2, kNull);
                    if (has_i) sts_u16(aNM + CW2 + k2, kNull);
                    if (has_d) sts_u16(aNM + 2 * CW2 + k2, kNull);
                    continue;
                }
                int m = 0;
                if (s > 0) {
                    int ins = -10, del = -10, sb = -10;
                    if (has_i) {
                        const int g = o_null ? kNull : lds_s16(aBM + k2 - 2);
                        const int ii = ie_null ? kNull : lds_s16(aEI + k2 - 2);
                        ins = (g == kNull && ii == kNull) ? kNull : (int)(short)(max(g, ii) + 1);
                        sts_u16(aNM + CW2 + k2, ins);
                        if (BT) hI[k] = (int16_t)ins;
                    }
                    if (has_d) {
                        const int g = o_null ? kNull : lds_s16(aBM + k2 + 2);
                        const int dd = de_null ? kNull : lds_s16(aED + k2 + 2);
                        del = max(g, dd);
                        sts_u16(aNM + 2 * CW2 + k2, del);
                        if (BT) hD[k] = (int16_t)del;
                    }
                    if (!sub_null) sb = (a_lo <= k && k <= a_hi) ? (int)(short)(lds_s16(aAM + k2) + 1) : kNull;
                    m = max(del, max(sb, ins));
                }
                const int v = m - k;
                if ((m | v) >= 0) {
                    const int lim = min(pl - v, tl - m);
                    if (lim > 0) m += packed ? match_packed_s(aP, aT, v, m, lim) : match_bytes(gp, gt, v, m, lim);
                }
                sts_u16(aNM + k2, m);
                if (BT) hM[k] = (int16_t)m;
                if (REDUCE) md = min(md, max(pl - (m - k), tl - m));
                if (k == ak && m >= tl) hit_end = true;
            }
            // ---- end reached (wfa.c:217-237).  Trimming never removes diagonal ak, so testing before the
            // reduction is equivalent, and the finishing wavefront's trimmed range is never read again. ----
            const uint32_t eb = __ballot_sync(kFull, hit_end);
            if (!done && ((eb >> subshift) & GM)) { done = true; reached = true; fscore = s; }
            if (__all_sync(kFull, done)) break;

            // ---- adaptive reduction (wfa.c:70-141) on the pairs still running ----
            if (REDUCE) {
                __syncwarp();
                const bool wide = !done && (hi - lo + 1) >= 10;
                md = group_min<G>(md);
                const int top_limit = min(ak - 1, hi);
                int newlo = lo;
                bool pend = wide && lo < top_limit;
                if (pend) newlo = top_limit;
                for (int c = 0; __any_sync(kFull, pend && (lo + c < top_limit)); c += G) {
                    const int k = lo + c + sl;
                    bool hit = false;
                    if (pend && k < top_limit) {
                        const int off = lds_s16(aNM + (uint32_t)(k * 2));
                        hit = (max(pl - (off - k), tl - off) - md) <= 50;
                    }
                    const uint32_t mine = (__ballot_sync(kFull, hit) >> subshift) & GM;
                    if (pend && mine) { newlo = lo + c + __ffs(mine) - 1; pend = false; }
                    if (lo + c + G >= top_limit) pend = false;
                }
                const int bottom_limit = max(ak + 1, newlo);
                int newhi = hi;
                pend = wide && hi > bottom_limit;
                if (pend) newhi = bottom_limit;
                for (int c = 0; __any_sync(kFull, pend && (hi - c > bottom_limit)); c += G) {
                    const int k = hi - c - sl;
                    bool hit = false;
                    if (pend && k > bottom_limit) {
                        const int off = lds_s16(aNM + (uint32_t)(k * 2));
                        hit = (max(pl - (off - k), tl - off) - md) <= 50;
                    }
                    const uint32_t mine = (__ballot_sync(kFull, hit) >> subshift) & GM;
                    if (pend && mine) { newhi = hi - c - (__ffs(mine) - 1); pend = false; }
                    if (hi - c - G <= bottom_limit) pend = false;
                }
                if (sl == 0 && !done) sts_u32(aDyn + (uint32_t)s * 4u, ((uint32_t)newlo & 0xffffu) | ((uint32_t)newhi << 16));
                // trimmed cells leave the wavefront: back to NULL in the ring (the arena keeps the values; the
                // backtrace range-checks against the trimmed range like the reference's getters)
                const bool trimmed = !done && (newlo != lo || newhi != hi);
                if (__any_sync(kFull, trimmed)) {
                    if (trimmed) {
                        for (int k = lo + sl; k <= hi; k += G) {
                            if (k >= newlo && k <= newhi) continue;
                            const uint32_t k2 = (uint32_t)(k * 2);
                            sts_u16(aNM + k2, kNull);
                            if (has_i) sts_u16(aNM + CW2 + k2, kNull);
                            if (has_d) sts_u16(aNM + 2 * CW2 + k2, kNull);
                        }
                    }
                }
            }
            __syncwarp();
        }

        __syncwarp();  // arena stores of the last wavefront must be visible to the backtracing lane
        // ---- backtrace (wfa_backtracing.c:219-375): first lane of every sub-warp, concurrently ----
        const int max_ops = pl + tl;
        int begin_offset = max_ops - 1;
        int status = AIM_STATUS_OK;
        if (BT && reached && sl == 0) {
            const int ops_cap = 2 * RS;
            int b = begin_offset;
            int score = fscore, k = ak;
            int offset;
            {
                const uint4 q1 = lds_v4(aPlan + (uint32_t)fscore * (PLAN_WORDS * 4) + 16);
                offset = arena[q1.x + (uint32_t)(k - lo16(q1.y))];
            }
            int v = offset - k, h = offset;
            bool valid = (v > 0 && v <= pl && h > 0 && h <= tl);
            int type = 0;  // 0 M, 1 I, 2 D
            bool bad = false;
#define AIM_PUT(ch) do { if (b < 0 || b >= ops_cap) { bad = true; } else { sts_u8(aOps + (uint32_t)b, (ch)); } --b; } while (0)
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
                int go_lo = 1, go_hi = -1, ge_lo = 1, ge_hi = -1, mm_lo = 1, mm_hi = -1, go_l0 = 0, ge_l0 = 0, mm_l0 = 0, ge_w = 0;
                if (s_open >= 0) {
                    const uint32_t a = aPlan + (uint32_t)s_open * (PLAN_WORDS * 4);
                    go_f = lds_u32(a + 4);
                    const uint32_t r = lds_u32(a + 20);
                    go_base = lds_u32(a + 16); go_l0 = lo16(r); go_lo = go_l0; go_hi = go_l0 + (int)(r >> 16) - 1;
                    if (REDUCE && (go_f & P_PRESENT)) { const uint32_t w = lds_u32(aDyn + (uint32_t)s_open * 4u); go_lo = lo16(w); go_hi = hi16s(w); }
                }
                if (s_ext >= 0) {
                    const uint32_t a = aPlan + (uint32_t)s_ext * (PLAN_WORDS * 4);
                    ge_f = lds_u32(a + 4);
                    const uint32_t r = lds_u32(a + 20);
                    ge_base = lds_u32(a + 16); ge_l0 = lo16(r); ge_w = (int)(r >> 16); ge_lo = ge_l0; ge_hi = ge_l0 + ge_w - 1;
                    if (REDUCE && (ge_f & P_PRESENT)) { const uint32_t w = lds_u32(aDyn + (uint32_t)s_ext * 4u); ge_lo = lo16(w); ge_hi = hi16s(w); }
                }
                if (s_mis >= 0) {
                    const uint32_t a = aPlan + (uint32_t)s_mis * (PLAN_WORDS * 4);
                    mm_f = lds_u32(a + 4);
                    const uint32_t r = lds_u32(a + 20);
                    mm_base = lds_u32(a + 16); mm_l0 = lo16(r); mm_lo = mm_l0; mm_hi = mm_l0 + (int)(r >> 16) - 1;
                    if (REDUCE && (mm_f & P_PRESENT)) { const uint32_t w = lds_u32(aDyn + (uint32_t)s_mis * 4u); mm_lo = lo16(w); mm_hi = hi16s(w); }
                }
                int del_ext = kNull, del_open = kNull, ins_ext = kNull, ins_open = kNull, misms = kNull;
                if (type != 1) {
                    if ((ge_f & P_PRESENT) && (ge_f & P_HAS_D) && ge_lo <= k + 1 && k + 1 <= ge_hi)
                        del_ext = arena[ge_base + (uint32_t)(((ge_f & P_HAS_I) ? 2 * ge_w : ge_w) + (k + 1 - ge_l0))];
                    if ((go_f & P_PRESENT) && go_lo <= k + 1 && k + 1 <= go_hi) del_open = arena[go_base + (uint32_t)(k + 1 - go_l0)];
                }
                if (type != 2) {
                    if ((ge_f & P_PRESENT) && (ge_f & P_HAS_I) && ge_lo <= k - 1 && k - 1 <= ge_hi)
                        ins_ext = (int16_t)(arena[ge_base + (uint32_t)(ge_w + (k - 1 - ge_l0))] + 1);
                    if ((go_f & P_PRESENT) && go_lo <= k - 1 && k - 1 <= go_hi)
                        ins_open = (int16_t)(arena[go_base + (uint32_t)(k - 1 - go_l0)] + 1);
                }
                if (type == 0) {
                    if ((mm_f & P_PRESENT) && mm_lo <= k && k <= mm_hi) misms = (int16_t)(arena[mm_base + (uint32_t)(k - mm_l0)] + 1);
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
        if (BT && active) {
            uint4 *dst = reinterpret_cast<uint4 *>(K.ops + (size_t)i * 2 * RS);
            for (int c = sl; c < (2 * RS) / 16; c += G) dst[c] = lds_v4(aOps + (uint32_t)c * 16u);
        }
        __syncwarp();
    }
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

template <int G>
cudaError_t launch_g(const SubK &K, bool reduce, bool bt, int grid, int block, size_t smem, cudaStream_t st)
{
    cudaError_t e;
#define AIM_LAUNCH(R, B)                                                                                             \
    do {                                                                                                             \
        e = cudaFuncSetAttribute(wfa_sub_kernel<G, R, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) wfa_sub_kernel<G, R, B><<<grid, block, smem, st>>>(K);                                  \
    } while (0)
    if (reduce && bt) AIM_LAUNCH(true, true);
    else if (reduce) AIM_LAUNCH(true, false);
    else if (bt) AIM_LAUNCH(false, true);
    else AIM_LAUNCH(false, false);
#undef AIM_LAUNCH
    return e;
}

}  // namespace

// Returns AIM_OK after enqueueing, 1 if this configuration is not served by the lockstep kernel
// (long reads / very large MAX_SCORE: the ring or the packed sequences do not fit), or an AIM_ERR_*.
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
    const uint32_t ring = (uint32_t)std::max(x, o + e) + 1;
    SubK K{};
    K.cw = round_up((uint32_t)(kmax - kmin + 3), 2);  // one NULL guard cell each side
    K.koff = 1 - kmin;
    K.ring_halfs = round_up(ring * 3 * K.cw, 8);
    if (K.cw > 4000 || ring * 3 * K.cw > 0xffffu) return 1;
    K.plan_words = (uint32_t)PLAN_WORDS * ((uint32_t)MS + 1);
    std::vector<uint32_t> plan(K.plan_words, 0u);
    std::vector<int> occ_lo(ring, 1), occ_hi(ring, -1);  // union of the static ranges a ring slot has held so far
    uint64_t arena_slots = 0;
    auto slot_off = [&](int s) -> uint32_t { return s < 0 ? 0u : ((uint32_t)s % ring) * 3u * K.cw; };
    for (int s = 0; s <= MS; ++s) {
        if (!w[s].present) continue;
        uint32_t *q = &plan[(size_t)s * PLAN_WORDS];
        const bool A = s - x >= 0 && w[s - x].present, B = s - o - e >= 0 && w[s - o - e].present, E = s - e >= 0 && w[s - e].present;
        const bool ie_null = !(E && w[s - e].has_i), de_null = !(E && w[s - e].has_d);
        int span_lo = w[s].lo, span_hi = w[s].hi;
        {
            int &ol = occ_lo[(size_t)s % ring], &oh = occ_hi[(size_t)s % ring];
            if (ol <= oh) { span_lo = std::min(span_lo, ol); span_hi = std::max(span_hi, oh); }
            ol = span_lo;
            oh = span_hi;
        }
        const uint32_t width = (uint32_t)(w[s].hi - w[s].lo + 1);
        q[0] = ((uint32_t)span_lo & 0xffffu) | ((uint32_t)(span_hi - span_lo + 1) << 16);
        q[1] = P_PRESENT | (A ? 0u : P_SUB_NULL) | (B ? 0u : P_O_NULL) | (ie_null ? P_IE_NULL : 0u) | (de_null ? P_DE_NULL : 0u) |
               (w[s].has_i ? P_HAS_I : 0u) | (w[s].has_d ? P_HAS_D : 0u);
        q[2] = slot_off(s) | (slot_off(s - x) << 16);
        q[3] = slot_off(s - o - e) | (slot_off(s - e) << 16);
        q[4] = (uint32_t)arena_slots;
        q[5] = ((uint32_t)w[s].lo & 0xffffu) | (width << 16);
        q[6] = A ? (((uint32_t)w[s - x].lo & 0xffffu) | ((uint32_t)w[s - x].hi << 16)) : (1u | (0xffffu << 16));
        arena_slots += (uint64_t)width * (1u + w[s].has_i + w[s].has_d);
    }
    if (arena_slots > 0x0fffffffu) return 1;

    K.plen = a.plen; K.tlen = a.tlen; K.patterns = a.patterns; K.texts = a.texts;
    K.results = a.results; K.ops = a.ops; K.n = a.n; K.idx_base = a.idx_base;
    K.x = x; K.o = o; K.e = e; K.max_score = MS; K.read_size = p.read_size;
    K.seq_words = round_up((uint32_t)p.read_size / 16 + 2, 4);
    K.dyn_words = p.reduce ? round_up((uint32_t)MS + 1, 4) : 0;
    const uint32_t ops_words = p.backtrace ? (2u * (uint32_t)p.read_size) / 4 : 0;
    const uint32_t pair_words_raw = 2 * K.seq_words + K.dyn_words + K.ring_halfs / 2 + ops_words;

    // lanes per pair; env override for tuning
    int G = 16;
    if (const char *gs = getenv("AIM_WFA_G")) { int g = atoi(gs); if (g == 8 || g == 16 || g == 32) G = g; }
    const int PPW = 32 / G;
    {   // stagger the pair slots of one warp over the banks: slot stride == 32/PPW words (mod 32)
        const uint32_t want = PPW > 1 ? 32u / (uint32_t)PPW : 0u;
        K.pair_words = pair_words_raw + ((want + 32u - pair_words_raw % 32u) % 32u);
    }
    const uint32_t kSmemBudget = 227u * 1024u, kSmemPerSm = 228u * 1024u, kBlockReserve = 1024u;
    const size_t pair_bytes = (size_t)K.pair_words * 4;
    const size_t plan_bytes = (size_t)K.plan_words * 4;
    int warps_per_block = 4;
    size_t smem_block = plan_bytes + (size_t)warps_per_block * PPW * pair_bytes;
    if (smem_block > kSmemBudget / 4) return 1;  // fewer than 16 warps/SM would fit: leave it to the long-read kernel
    int blocks_per_sm = (int)std::min<uint32_t>(kSmemPerSm / ((uint32_t)smem_block + kBlockReserve), 32u);
    blocks_per_sm = std::max(1, std::min(blocks_per_sm, 48 / warps_per_block));
    int grid = sc->sm_count * blocks_per_sm;
    {
        const uint64_t per_block = (uint64_t)warps_per_block * PPW;
        if ((uint64_t)grid * per_block > a.n) grid = (int)((a.n + per_block - 1) / per_block);
    }
    const uint64_t total_slots = (uint64_t)grid * warps_per_block * PPW;

    // device copies: plan + (BT) history arena
    const size_t plan_dev = (plan_bytes + 255) / 256 * 256;
    K.arena_stride = p.backtrace ? (size_t)round_up((uint32_t)arena_slots, 64) : 0;
    const size_t arena_bytes = (size_t)total_slots * K.arena_stride * 2;
    int rc = scratch_reserve(sc, plan_dev + arena_bytes);
    if (rc != AIM_OK) return rc;
    cudaError_t err = cudaMemcpyAsync(sc->buf, plan.data(), plan_bytes, cudaMemcpyHostToDevice, stream);
    K.plan = reinterpret_cast<const uint32_t *>(sc->buf);
    K.arena = reinterpret_cast<int16_t *>(reinterpret_cast<unsigned char *>(sc->buf) + plan_dev);
    if (err == cudaSuccess) {
        const int block = warps_per_block * 32;
        if (G == 8) err = launch_g<8>(K, p.reduce != 0, p.backtrace != 0, grid, block, smem_block, stream);
        else if (G == 16) err = launch_g<16>(K, p.reduce != 0, p.backtrace != 0, grid, block, smem_block, stream);
        else err = launch_g<32>(K, p.reduce != 0, p.backtrace != 0, grid, block, smem_block, stream);
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("wfa_sub launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    return AIM_OK;
}

}  // namespace aim
