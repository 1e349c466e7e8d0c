// Long-read WFA / WFA-adaptive (score, optionally + CIGAR) for sm_100a: G lanes per pair, 32/G pairs per warp in
// LOCKSTEP over the scores, wavefronts in WINDOWED shared-memory rings.
//
// Same algorithm and literal semantics as aim_wfa.cu / aim_wfa_sub.cu (reference:
// WFA/DPU-MRAM/dpu/wfa.c:70-407).  What changes for reads of thousands of bases:
//   * MAX_SCORE is in the thousands, so the static wavefront range (2s+1 diagonals) is useless; with
//     adaptive trimming (wfa.c:70-141) the live range stays around 50-130 diagonals.  Every ring slot
//     is therefore a 256-diagonal window addressed modulo 256 (diagonal k lives in cell k & 255): no
//     per-slot origin to track while the window drifts with the alignment's diagonal.  A pair whose
//     wavefront outgrows the window (224 diagonals) is handed to the warp-per-pair kernel of
//     aim_wfa.cu (global-memory wavefronts) through a device-side list; so is a pair holding a byte
//     outside {A,C,G,T} (the reference compares raw bytes, wfa.c:209).
//   * every stored wavefront is framed by PAD cells of NULL on both sides.  From score ~o+e+x on every
//     score carries all three components; when in addition the three source ranges start and end within
//     PAD-2 diagonals of each other (the normal case: trimming moves the ends by a diagonal or two),
//     out-of-range reads land on the frame and return exactly the NULL the reference substitutes
//     (wfa.c:243-266), so the inner loop needs no range tests - only the "M[k]+1 on an out-of-range
//     diagonal stays NULL" rule of the substitution term keeps one.  Otherwise the literal loop runs.
//   * compute_offsets only reads M of scores s-x, s-o-e and I/D of score s-e: the rings hold
//     max(x,o+e)+1 M wavefronts and e+1 I/D wavefronts plus one (lo|hi) word per live score - about
//     5 KB per pair at x3 o4 e1, i.e. ~44 pairs resident per SM.
//   * the sequences are packed 2 bits/base by a pre-pass into HBM in a DUPLICATED layout (entry i =
//     {word i, word i+1}) so that the 16-base window at any offset is one 8-byte load + one funnel
//     shift; the extend loop reads them through L1 (only the ~100 bases around the current offsets
//     are hot at any time).
//   * the per-score schedule (which scores exist, which components they carry, ring slot offsets)
//     depends only on the penalties: a 16-byte record per score, read warp-uniformly from global.
//   * pairs are handed out by an atomic counter (cost per pair varies with its score).
//   * with BACKTRACE every computed (score, diagonal) cell is also streamed as one 8-byte {M | I << 16, D} record
//     to a per-pair-slot HBM arena, plus one 16-byte record per score (trimmed range, array origin, arena base):
//     the reference's wfa_component store (common.h:126-138, dpu_allocator_mram.c).  The backtrace
//     (wfa_backtracing.c:219-375) is walked by the first lane of every sub-warp with the reference's candidate
//     order; a pair whose history outgrows the slot is handed to the warp-per-pair kernel as well.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "aim_internal.h"
#include "aim_wfa_common.cuh"

namespace aim {

namespace {

constexpr uint32_t L_PRESENT = 1, L_SUB_NULL = 2, L_O_NULL = 4, L_IE_NULL = 8, L_DE_NULL = 16, L_HAS_I = 32, L_HAS_D = 64;
// Window geometry (template parameters of the kernel): WC cells (diagonals) per ring slot, a power of two; PAD NULL cells kept on
// both sides of every stored wavefront; W_CAP = WC - 2 * PAD = the widest wavefront a window serves; an M slot is WC int16, an
// I/D slot the I array then the D array.  Two geometries are instantiated: 128 / 8 (2.6 KB of shared memory per pair: 36 instead
// of 20 resident warps per SM) for the first pass and 256 / 16 for the pairs that outgrow it - at l = 10 K, e = 10 % the widest
// wavefront of a pair is 91 diagonals in the median, 111 at the 99th percentile and 118 at most (1 500 pairs, counted on
// the CPU), so 112 serves all but ~1 % of them.
__host__ __device__ constexpr uint32_t m_slot_bytes(int wc) { return (uint32_t)wc * 2u; }
__host__ __device__ constexpr uint32_t id_slot_bytes(int wc) { return (uint32_t)wc * 4u; }
__host__ __device__ constexpr int w_cap(int wc, int pad) { return wc - 2 * pad; }
// per-score plan (4 words): w0 flags
//                           w1 M slot offset of s   | M slot offset of s-x << 16      (bytes)
//                           w2 M slot offset of s-o-e | I/D slot offset of s-e << 16
//                           w3 I/D slot offset of s | range-ring index of s-e << 16

struct LongK {
    const int32_t *plen;
    const int32_t *tlen;
    const uint2 *packed;        // [pair][2][pk_words] duplicated 2-bit words
    const unsigned char *dirty; // pair holds a non-ACGT byte
    aim_result *results;
    const uint4 *plan;
    uint32_t *work_ctr;
    uint32_t *fail_list;
    uint32_t *fail_count;
    const uint32_t *in_list;   // second pass: the pairs the first pass handed back (NULL: pairs 0 .. n-1)
    const uint32_t *in_count;
    uint32_t n, idx_base;
    int x, o, e;
    int max_score, read_size;
    uint32_t pk_words;    // entries per packed sequence
    uint32_t dyn_words;   // range ring words per pair (multiple of 4, >= ring_m)
    uint32_t mring_bytes; // ring_m * M_SLOT_BYTES
    uint32_t pair_words;  // shared-memory words per pair slot
    // backtrace only
    char *ops;
    uint2 *arena;         // history cells, arena_cells per pair slot
    uint4 *meta;          // per pair slot and score: {lo | hi << 16 (trimmed), array origin lo, arena base, -}
    uint32_t arena_cells;
};

// ---- pre-pass: ASCII rows -> duplicated 2-bit words; one warp per sequence ----
__global__ void __launch_bounds__(256) pack_kernel(const int32_t *plen, const int32_t *tlen, const char *patterns, const char *texts, uint32_t n,
                                                   int RS, uint32_t pk_words, uint2 *packed, unsigned char *dirty)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint64_t u = gw; u < 2ull * n; u += nw) {
        const uint32_t i = (uint32_t)(u >> 1);
        const bool is_text = u & 1;
        const int len = min(max(is_text ? tlen[i] : plen[i], 0), RS);
        const char *g = (is_text ? texts : patterns) + (size_t)i * RS;
        uint2 *out = packed + (size_t)u * pk_words;
        bool ok = true;
        for (uint32_t e = lane; e < pk_words; e += 32) {
            uint32_t w[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int b0 = 16 * (int)(e + q);  // first base of word e+q
                uint32_t v = 0;
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const int off = b0 + 8 * hlf;
                    uint2 raw = make_uint2(0u, 0u);
                    if (off < len) raw = __ldg(reinterpret_cast<const uint2 *>(g + off));
                    v = (v << 16) | pack8(raw, len - off, &ok);
                }
                w[q] = v;
            }
            out[e] = make_uint2(w[0], w[1]);
        }
        if (!__all_sync(kFull, ok) && lane == 0) dirty[i] = 1;
    }
}

__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, int v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((short)v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// equal bases from pattern[v], text[h], at most lim (> 0); duplicated packed words through L1
__device__ __forceinline__ int match_packed_g(const uint2 *P2, const uint2 *T2, int v, int h, int lim)
{
    int cnt = 0;
    for (int pv = v, ph = h;; pv += 16, ph += 16) {
        const uint2 a2 = __ldg(P2 + ((uint32_t)pv >> 4)), b2 = __ldg(T2 + ((uint32_t)ph >> 4));
        const uint32_t a = __funnelshift_l(a2.y, a2.x, 2 * pv);  // (the funnel shift takes its amount modulo 32)
        const uint32_t b = __funnelshift_l(b2.y, b2.x, 2 * ph);
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
__device__ __forceinline__ bool in_range(int k, int lo, int hi) { return (unsigned)(k - lo) <= (unsigned)(hi - lo) && lo <= hi; }

template <int G, bool REDUCE, bool BT, int WC, int PAD>
__global__ void __launch_bounds__(128) wfa_long_kernel(const LongK K)
{
    constexpr int W_CAP = w_cap(WC, PAD);
    constexpr uint32_t M_SLOT_BYTES = m_slot_bytes(WC);
    auto cell = [](int k) -> uint32_t { return ((uint32_t)k & (uint32_t)(WC - 1)) << 1; };  // byte offset of diagonal k inside a ring array
    constexpr int PPW = 32 / G;
    constexpr uint32_t GM = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    extern __shared__ __align__(16) uint32_t smem_w[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sub = lane / G, sl = lane % G, subshift = sub * G;
    const int RS = K.read_size, MS = K.max_score;

    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_w);
    const uint32_t aDyn = sbase + (uint32_t)(wib * PPW + sub) * K.pair_words * 4u;  // (lo | hi << 16) per live score
    const uint32_t aMR = aDyn + K.dyn_words * 4u;                                  // M ring
    const uint32_t aIDR = aMR + K.mring_bytes;                                     // I/D ring
    const uint32_t slot_global = (blockIdx.x * (blockDim.x >> 5) + (uint32_t)wib) * PPW + (uint32_t)sub;
    uint2 *arena = BT ? K.arena + (size_t)slot_global * K.arena_cells : nullptr;
    uint4 *meta = BT ? K.meta + (size_t)slot_global * (size_t)(MS + 1) : nullptr;
    const int X = K.x, OE = K.o + K.e, E = K.e;

    const uint32_t n_work = K.in_count ? min(__ldg(K.in_count), K.n) : K.n;
    for (;;) {
        uint32_t base_i = 0;
        if (lane == 0) base_i = atomicAdd(K.work_ctr, (uint32_t)PPW);
        base_i = __shfl_sync(kFull, base_i, 0);
        if (base_i >= n_work) break;  // warp-uniform
        const uint32_t wi = base_i + (uint32_t)sub;
        bool active = wi < n_work;
        const uint32_t i = active ? (K.in_list ? K.in_list[wi] : wi) : 0u;
        bool failed = false;
        if (active && K.dirty[i]) { failed = true; active = false; }
        const int pl = active ? min(max(K.plen[i], 0), RS) : 0;
        const int tl = active ? min(max(K.tlen[i], 0), RS) : 0;
        const uint2 *P2 = K.packed + (size_t)(active ? i : 0) * 2 * K.pk_words;
        const uint2 *T2 = P2 + K.pk_words;
        const int ak = tl - pl;

        bool done = !active;
        int fscore = MS + 1;  // give-up value (wfa.c:399-404)
        bool reached = false;
        uint32_t abase = 0;   // arena cells used so far
        char *gops = BT ? K.ops + (size_t)(active ? i : 0) * 2 * RS : nullptr;
        if (BT && active) {   // op row: 'M' everywhere (wfa.c:499-501); the backtrace overwrites the edits
            uint4 *dst = reinterpret_cast<uint4 *>(gops);
            for (int c = sl; c < (2 * RS) / 16; c += G) dst[c] = make_uint4(0x4d4d4d4du, 0x4d4d4d4du, 0x4d4d4d4du, 0x4d4d4d4du);
        }

        for (int s = 0; s <= MS; ++s) {
            const uint4 pw = __ldg(K.plan + s);  // warp-uniform
            const uint32_t fl = pw.x;
            if (!(fl & L_PRESENT)) continue;
            const bool sub_null = fl & L_SUB_NULL, o_null = fl & L_O_NULL, ie_null = fl & L_IE_NULL, de_null = fl & L_DE_NULL;
            const bool has_i = fl & L_HAS_I, has_d = fl & L_HAS_D;
            const uint32_t offN = pw.y & 0xffffu, offA = pw.y >> 16, offB = pw.z & 0xffffu;

            // this pair's range (wfa.c:318-343) from the (trimmed) ranges of the three source wavefronts
            int lo = 0, hi = 0;
            int a_lo = 1, a_hi = -1, b_lo = 1, b_hi = -1, e_lo = 1, e_hi = -1;
            if (s > 0) {
                if (!sub_null) { const uint32_t w = lds_u32(aDyn + (offA / M_SLOT_BYTES) * 4u); a_lo = lo16(w); a_hi = hi16s(w); }
                if (!o_null) { const uint32_t w = lds_u32(aDyn + (offB / M_SLOT_BYTES) * 4u); b_lo = lo16(w); b_hi = hi16s(w); }
                if (!(ie_null && de_null)) { const uint32_t w = lds_u32(aDyn + (pw.w >> 16) * 4u); e_lo = lo16(w); e_hi = hi16s(w); }
                lo = min(min(a_lo, b_lo), e_lo) - 1;
                hi = max(max(a_hi, b_hi), e_hi) + 1;
            }
            if (!done && hi - lo + 1 > W_CAP) { done = true; failed = true; }  // window outgrown: leave it to the warp-per-pair kernel
            if (BT && !done && abase + (uint32_t)(hi - lo + 1) > K.arena_cells) { done = true; failed = true; }  // history outgrown
            const bool ran = !done;
            uint2 *hC = BT ? arena + abase - lo : nullptr;  // this score's history cells, indexed by k
            const uint32_t aNM = aMR + offN, aAM = aMR + offA, aBM = aMR + offB;
            const uint32_t aEI = aIDR + (pw.z >> 16), aED = aEI + M_SLOT_BYTES;
            const uint32_t aNI = aIDR + (pw.w & 0xffffu), aND = aNI + M_SLOT_BYTES;

            // ---- compute_offsets (wfa.c:238-273) fused with extend (wfa.c:193-215) ----
            int md = max(pl, tl);
            // frame-based loop: all components present and the source ranges nearly aligned, for every running pair of the warp
            bool framed = (fl & (L_SUB_NULL | L_O_NULL | L_IE_NULL | L_DE_NULL | L_HAS_I | L_HAS_D)) == (L_HAS_I | L_HAS_D);
            if (framed) {
                const int lo_max = max(max(a_lo, b_lo), e_lo), hi_min = min(min(a_hi, b_hi), e_hi);
                const bool ok = done || ((lo_max - (lo + 1) + 2 <= PAD) && ((hi - 1) - hi_min + 2 <= PAD));
                framed = __all_sync(kFull, ok);
            }
            if (!done && framed) {
                const uint32_t a_w = (uint32_t)(a_hi - a_lo);
                int plk = pl + lo + sl;                      // pl + k
                uint32_t ka = (uint32_t)(lo + sl - a_lo);    // k - a_lo
                for (int k = lo + sl; k <= hi; k += G, plk += G, ka += G) {
                    const uint32_t k2 = (uint32_t)k << 1;
                    const uint32_t ck = k2 & (2 * WC - 2), ckm = (k2 - 2u) & (2 * WC - 2), ckp = (k2 + 2u) & (2 * WC - 2);
                    const int mx = max(lds_s16(aBM + ckm), lds_s16(aEI + ckm));
                    const int ins = (mx == kNull) ? kNull : mx + 1;
                    const int del = max(lds_s16(aBM + ckp), lds_s16(aED + ckp));
                    const int sa = lds_s16(aAM + ck) + 1;
                    const int sb = (ka <= a_w) ? sa : kNull;
                    sts_u16(aNI + ck, ins);
                    sts_u16(aND + ck, del);
                    int m = max(del, max(sb, ins));
                    const int v = m - k;
                    if ((m | v) >= 0) {
                        const int lim = min(pl - v, tl - m);
                        if (lim > 0) m += match_packed_g(P2, T2, v, m, lim);
                    }
                    sts_u16(aNM + ck, m);
                    if (BT) hC[k] = make_uint2(((uint32_t)m & 0xffffu) | ((uint32_t)ins << 16), (uint32_t)del & 0xffffu);
                    if (REDUCE) md = min(md, max(plk, tl) - m);  // max(pl - (m - k), tl - m)
                }
            } else if (!done) {
                for (int k = lo + sl; k <= hi; k += G) {
                    const uint32_t ck = cell(k), ckm = cell(k - 1), ckp = cell(k + 1);
                    int m = 0, ins = -10, del = -10;
                    if (s > 0) {
                        int sb = -10;
                        if (has_i) {
                            const int g = in_range(k - 1, b_lo, b_hi) ? lds_s16(aBM + ckm) : kNull;
                            const int ii = (!ie_null && in_range(k - 1, e_lo, e_hi)) ? lds_s16(aEI + ckm) : kNull;
                            ins = (g == kNull && ii == kNull) ? kNull : (int)(short)(max(g, ii) + 1);
                            sts_u16(aNI + ck, ins);
                        }
                        if (has_d) {
                            const int g = in_range(k + 1, b_lo, b_hi) ? lds_s16(aBM + ckp) : kNull;
                            const int dd = (!de_null && in_range(k + 1, e_lo, e_hi)) ? lds_s16(aED + ckp) : kNull;
                            del = max(g, dd);
                            sts_u16(aND + ck, del);
                        }
                        if (!sub_null) sb = in_range(k, a_lo, a_hi) ? (int)(short)(lds_s16(aAM + ck) + 1) : kNull;
                        m = max(del, max(sb, ins));
                    }
                    const int v = m - k;
                    if ((m | v) >= 0) {
                        const int lim = min(pl - v, tl - m);
                        if (lim > 0) m += match_packed_g(P2, T2, v, m, lim);
                    }
                    sts_u16(aNM + ck, m);
                    if (BT) hC[k] = make_uint2(((uint32_t)m & 0xffffu) | ((uint32_t)ins << 16), (uint32_t)del & 0xffffu);
                    if (REDUCE) md = min(md, max(pl - (m - k), tl - m));
                }
            }
            // ---- end reached (wfa.c:217-237); trimming never removes diagonal ak, so testing before the
            // reduction is equivalent and the finishing wavefront's trimmed range is never read again.  The offset of
            // diagonal ak is read back from the row (one broadcast load per score instead of a test per cell) ----
            __syncwarp();
            if (!done && in_range(ak, lo, hi) && lds_s16(aNM + cell(ak)) >= tl) { done = true; reached = true; fscore = s; }
            if (BT && ran && sl == 0 && done) meta[s] = make_uint4(((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16), (uint32_t)lo, abase, 0u);
            if (__all_sync(kFull, done)) break;

            // ---- adaptive reduction (wfa.c:70-141) on the pairs still running ----
            int newlo = lo, newhi = hi;
            if (REDUCE) {
                __syncwarp();
                const bool wide = !done && (hi - lo + 1) >= 10;
                if (__any_sync(kFull, wide)) {
                    md = group_min<G>(md);
                    const int top_limit = min(ak - 1, hi);
                    bool pend = wide && lo < top_limit;
                    if (pend) newlo = top_limit;
                    for (int c = 0; __any_sync(kFull, pend && (lo + c < top_limit)); c += G) {
                        const int k = lo + c + sl;
                        bool hit = false;
                        if (pend && k < top_limit) {
                            const int off = lds_s16(aNM + cell(k));
                            hit = (max(pl - (off - k), tl - off) - md) <= 50;
                        }
                        const uint32_t mine = (__ballot_sync(kFull, hit) >> subshift) & GM;
                        if (pend && mine) { newlo = lo + c + __ffs(mine) - 1; pend = false; }
                        if (lo + c + G >= top_limit) pend = false;
                    }
                    const int bottom_limit = max(ak + 1, newlo);
                    pend = wide && hi > bottom_limit;
                    if (pend) newhi = bottom_limit;
                    for (int c = 0; __any_sync(kFull, pend && (hi - c > bottom_limit)); c += G) {
                        const int k = hi - c - sl;
                        bool hit = false;
                        if (pend && k > bottom_limit) {
                            const int off = lds_s16(aNM + cell(k));
                            hit = (max(pl - (off - k), tl - off) - md) <= 50;
                        }
                        const uint32_t mine = (__ballot_sync(kFull, hit) >> subshift) & GM;
                        if (pend && mine) { newhi = hi - c - (__ffs(mine) - 1); pend = false; }
                        if (hi - c - G <= bottom_limit) pend = false;
                    }
                }
            }
            __syncwarp();  // the trim scans above read cells of other lanes; the frame below overwrites cells they may have read
            if (!done) {
                if (sl == 0) sts_u32(aDyn + (offN / M_SLOT_BYTES) * 4u, ((uint32_t)newlo & 0xffffu) | ((uint32_t)newhi << 16));
                if (BT) {
                    if (sl == 0) meta[s] = make_uint4(((uint32_t)newlo & 0xffffu) | ((uint32_t)newhi << 16), (uint32_t)lo, abase, 0u);
                    abase += (uint32_t)(hi - lo + 1);
                }
                // NULL frame around the (trimmed) wavefront: PAD cells below newlo and above newhi, every component
                for (int j = sl; j < 2 * PAD; j += G) {
                    const uint32_t cf = cell(j < PAD ? newlo - 1 - j : newhi + 1 + (j - PAD));
                    sts_u16(aNM + cf, kNull);
                    if (has_i) sts_u16(aNI + cf, kNull);
                    if (has_d) sts_u16(aND + cf, kNull);
                }
            }
            __syncwarp();
        }
        __syncwarp();

        // ---- backtrace (wfa_backtracing.c:219-375): first lane of every sub-warp, concurrently ----
        const int max_ops = pl + tl;
        int begin_offset = max_ops - 1;
        int status = AIM_STATUS_OK;
        if (BT && reached && !failed && sl == 0) {
            __threadfence_block();
            const int ops_cap = 2 * RS;
            int b = begin_offset;
            int score = fscore, k = ak;
            int offset;
            {
                const uint4 mf = meta[fscore];
                offset = lo16(arena[mf.z + (uint32_t)(k - (int)mf.y)].x);
            }
            int v = offset - k, h = offset;
            bool valid = (v > 0 && v <= pl && h > 0 && h <= tl);
            int type = 0;  // 0 M, 1 I, 2 D
            bool bad = false;
#define AIM_PUT(ch) do { if (b < 0 || b >= ops_cap) { bad = true; } else { gops[b] = (ch); } --b; } while (0)
            while (v > 0 && h > 0 && score > 0 && !bad) {
                if (!valid) {
                    valid = (v > 0 && v <= pl && h > 0 && h <= tl);
                    if (valid) {  // add_trailing_gap (wfa_backtracing.c:48-69)
                        if (k < ak) { for (int j = k; j < ak; ++j) AIM_PUT('I'); }
                        else if (k > ak) { for (int j = ak; j < k; ++j) AIM_PUT('D'); }
                    }
                }
                const int s_open = score - OE, s_ext = score - E, s_mis = score - X;
                // records: flags from the plan, trimmed range / origin / arena base from the per-score meta
                uint32_t go_f = 0, ge_f = 0, mm_f = 0, go_base = 0, ge_base = 0, mm_base = 0;
                int go_lo = 1, go_hi = -1, ge_lo = 1, ge_hi = -1, mm_lo = 1, mm_hi = -1, go_l0 = 0, ge_l0 = 0, mm_l0 = 0;
                if (s_open >= 0) {
                    go_f = __ldg(&K.plan[s_open].x);
                    if (go_f & L_PRESENT) { const uint4 q = meta[s_open]; go_lo = lo16(q.x); go_hi = hi16s(q.x); go_l0 = (int)q.y; go_base = q.z; }
                }
                if (s_ext >= 0) {
                    ge_f = __ldg(&K.plan[s_ext].x);
                    if (ge_f & L_PRESENT) { const uint4 q = meta[s_ext]; ge_lo = lo16(q.x); ge_hi = hi16s(q.x); ge_l0 = (int)q.y; ge_base = q.z; }
                }
                if (s_mis >= 0) {
                    mm_f = __ldg(&K.plan[s_mis].x);
                    if (mm_f & L_PRESENT) { const uint4 q = meta[s_mis]; mm_lo = lo16(q.x); mm_hi = hi16s(q.x); mm_l0 = (int)q.y; mm_base = q.z; }
                }
                int del_ext = kNull, del_open = kNull, ins_ext = kNull, ins_open = kNull, misms = kNull;
                if (type != 1) {
                    if ((ge_f & L_PRESENT) && (ge_f & L_HAS_D) && ge_lo <= k + 1 && k + 1 <= ge_hi)
                        del_ext = lo16(arena[ge_base + (uint32_t)(k + 1 - ge_l0)].y);
                    if ((go_f & L_PRESENT) && go_lo <= k + 1 && k + 1 <= go_hi) del_open = lo16(arena[go_base + (uint32_t)(k + 1 - go_l0)].x);
                }
                if (type != 2) {
                    if ((ge_f & L_PRESENT) && (ge_f & L_HAS_I) && ge_lo <= k - 1 && k - 1 <= ge_hi)
                        ins_ext = (int16_t)(hi16s(arena[ge_base + (uint32_t)(k - 1 - ge_l0)].x) + 1);
                    if ((go_f & L_PRESENT) && go_lo <= k - 1 && k - 1 <= go_hi)
                        ins_open = (int16_t)(lo16(arena[go_base + (uint32_t)(k - 1 - go_l0)].x) + 1);
                }
                if (type == 0) {
                    if ((mm_f & L_PRESENT) && mm_lo <= k && k <= mm_hi) misms = (int16_t)(lo16(arena[mm_base + (uint32_t)(k - mm_l0)].x) + 1);
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
        if (sl == 0) {
            if (failed) {
                K.fail_list[atomicAdd(K.fail_count, 1u)] = i;
            } else if (active) {
                aim_result r;
                r.max_operations = max_ops;
                r.begin_offset = begin_offset;
                r.end_offset = max_ops;
                r.score = fscore;
                r.status = status;
                r.idx = K.idx_base + i;
                K.results[i] = r;
            }
        }
        __syncwarp();
    }
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// Picks the instantiation, sets its shared-memory attribute and returns its resident grid.
typedef void (*LongKernel)(const LongK);
template <int G, int WC, int PAD>
LongKernel pick_g(bool reduce, bool bt)
{
    if (reduce) return bt ? wfa_long_kernel<G, true, true, WC, PAD> : wfa_long_kernel<G, true, false, WC, PAD>;
    return bt ? wfa_long_kernel<G, false, true, WC, PAD> : wfa_long_kernel<G, false, false, WC, PAD>;
}
LongKernel pick(int G, int wc, bool reduce, bool bt)
{
    if (wc == 128) return G == 8 ? pick_g<8, 128, 8>(reduce, bt) : G == 16 ? pick_g<16, 128, 8>(reduce, bt) : pick_g<32, 128, 8>(reduce, bt);
    return G == 8 ? pick_g<8, 256, 16>(reduce, bt) : G == 16 ? pick_g<16, 256, 16>(reduce, bt) : pick_g<32, 256, 16>(reduce, bt);
}

}  // namespace

// Returns AIM_OK after enqueueing, 1 if this configuration is not served here, or an AIM_ERR_*.
int launch_wfa_long(const KernelArgs &a, Scratch *sc, void *stream_v, int *launches)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    const aim_params &p = a.p;
    if (const char *mode = getenv("AIM_WFA_MODE")) { if (std::string(mode) == "warp") return 1; }
    const int MS = p.max_score, x = p.mismatch, o = p.gap_open, e = p.gap_ext;
    if (!p.reduce && MS > 2 * w_cap(256, 16)) return 1;  // untrimmed wavefronts (2s+1 wide) would mostly outgrow the window
    const uint32_t ring_m = (uint32_t)std::max(x, o + e) + 1, ring_e = (uint32_t)e + 1;
    if (ring_m * m_slot_bytes(256) > 0xffffu || ring_e * id_slot_bytes(256) > 0xffffu) return 1;
    // first pass in the narrow window when trimming keeps the wavefronts inside it (see the geometry note above); AIM_WFA_LONG_WC=256: one pass
    int wc1 = p.reduce ? 128 : 256;
    if (const char *ws = getenv("AIM_WFA_LONG_WC")) { const int v = atoi(ws); if (v == 128 || v == 256) wc1 = v; }
    const int npass = wc1 == 128 ? 2 : 1;

    // ---- static schedule (presence and components only; ranges are dynamic here); slot offsets depend on the window ----
    struct S { bool present, has_i, has_d; };
    std::vector<S> w((size_t)MS + 1);
    w[0] = {true, false, false};
    for (int s = 1; s <= MS; ++s) {
        const bool A = s - x >= 0 && w[s - x].present, B = s - o - e >= 0 && w[s - o - e].present, E = s - e >= 0 && w[s - e].present;
        const bool ie_null = !(E && w[s - e].has_i), de_null = !(E && w[s - e].has_d);
        const bool i_out_null = !B && ie_null, d_out_null = !B && de_null;
        if (!A && i_out_null && d_out_null) { w[s] = {false, false, false}; continue; }
        w[s] = {true, !i_out_null, !d_out_null};
    }
    auto make_plan = [&](int wc) {
        std::vector<uint4> plan((size_t)MS + 1, make_uint4(0, 0, 0, 0));
        auto m_off = [&](int s) -> uint32_t { return s < 0 ? 0u : ((uint32_t)s % ring_m) * m_slot_bytes(wc); };
        auto id_off = [&](int s) -> uint32_t { return s < 0 ? 0u : ((uint32_t)s % ring_e) * id_slot_bytes(wc); };
        for (int s = 0; s <= MS; ++s) {
            if (!w[s].present) continue;
            const bool A = s - x >= 0 && w[s - x].present, B = s - o - e >= 0 && w[s - o - e].present, E = s - e >= 0 && w[s - e].present;
            const bool ie_null = !(E && w[s - e].has_i), de_null = !(E && w[s - e].has_d);
            uint4 q;
            q.x = L_PRESENT | (A ? 0u : L_SUB_NULL) | (B ? 0u : L_O_NULL) | (ie_null ? L_IE_NULL : 0u) | (de_null ? L_DE_NULL : 0u) |
                  (w[s].has_i ? L_HAS_I : 0u) | (w[s].has_d ? L_HAS_D : 0u);
            q.y = m_off(s) | (m_off(s - x) << 16);
            q.z = m_off(s - o - e) | (id_off(s - e) << 16);
            q.w = id_off(s) | ((s - e < 0 ? 0u : (uint32_t)(s - e) % ring_m) << 16);
            plan[(size_t)s] = q;
        }
        return plan;
    };

    int G1 = 16;
    if (const char *gs = getenv("AIM_WFA_LONG_G")) { int g = atoi(gs); if (g == 8 || g == 16 || g == 32) G1 = g; }
    const bool bt = p.backtrace != 0;
    const int block = 128;

    // per pass: kernel arguments that depend on the window, the instantiation and its resident grid
    struct Pass { LongK K; LongKernel fn; size_t smem; int grid; std::vector<uint4> plan; };
    Pass ps[2];
    size_t slots = 0;
    for (int q = 0; q < npass; ++q) {
        const int wc = q == 0 ? wc1 : 256;
        // the second pass is a tail - a few hundred pairs of milliseconds each on an otherwise empty GPU - so every pair gets a whole warp
        const int G = q == 0 ? G1 : 32;
        const uint32_t PPW = 32u / (uint32_t)G;
        LongK &K = ps[q].K;
        K = LongK{};
        K.plen = a.plen; K.tlen = a.tlen; K.results = a.results; K.n = a.n; K.idx_base = a.idx_base;
        K.x = x; K.o = o; K.e = e; K.max_score = MS; K.read_size = p.read_size;
        K.pk_words = (uint32_t)p.read_size / 16 + 1;
        K.dyn_words = round_up(ring_m, 4);
        K.mring_bytes = ring_m * m_slot_bytes(wc);
        {   // stagger the pair slots of one warp over the banks
            const uint32_t raw = K.dyn_words + (K.mring_bytes + ring_e * id_slot_bytes(wc)) / 4;
            const uint32_t want = PPW > 1 ? std::max(4u, 32u / PPW) : 0u;
            K.pair_words = raw + ((want + 32u - raw % 32u) % 32u);
        }
        const size_t pair_bytes = (size_t)K.pair_words * 4;
        if (pair_bytes * 4 * PPW > 227u * 1024u / 2) return 1;  // fewer than two blocks per SM: not worth it
        ps[q].fn = pick(G, wc, p.reduce != 0, bt);
        ps[q].smem = (size_t)(block / 32) * PPW * pair_bytes;
        int bps = 0;
        cudaError_t err = cudaFuncSetAttribute(ps[q].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps[q].smem);
        if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ps[q].fn, block, ps[q].smem);
        if (err != cudaSuccess) { set_error(std::string("wfa_long setup: ") + cudaGetErrorString(err)); cudaGetLastError(); return AIM_ERR_CUDA; }
        int grid = sc->sm_count * std::max(1, bps);
        const uint64_t per_block = (uint64_t)(block / 32) * PPW;
        grid = (int)std::min<uint64_t>((uint64_t)grid, (a.n + per_block - 1) / per_block);
        ps[q].grid = grid;
        slots = std::max(slots, (size_t)grid * (block / 32) * PPW);
        ps[q].plan = make_plan(wc);
    }

    // the warp-per-pair kernel serves what the last pass hands back (rare: a few resident warps are enough)
    // (without trimming the window is outgrown at score ~W_CAP/2: hand-overs are then common, keep the full grid)
    const WarpPlan W = wfa_warp_plan(a, sc->sm_count, p.reduce ? std::min<uint32_t>(a.n, (uint32_t)sc->sm_count * 4u) : a.n);
    if (W.rc != AIM_OK) { set_error("READ_SIZE too large for the shared-memory sequence stage"); return W.rc; }

    // history arena per resident pair slot (backtrace): 4 MiB by default = 512 K cells, ~3x what a 10 Kbp / 10 % pair
    // writes; a pair that needs more is handed to the warp-per-pair kernel with its own (arena_mb) arena
    const uint32_t arena_cells = bt ? (uint32_t)(((size_t)(p.arena_mb > 0 ? p.arena_mb : 4) << 20) / sizeof(uint2)) : 0;

    // scratch: counters | dirty | fail lists (one per pass) | packed sequences | meta | arena | warp-per-pair scratch
    const size_t off_dirty = 256;
    const size_t off_fail = align256(off_dirty + a.n);
    const size_t fail_bytes = align256((size_t)a.n * 4);
    const size_t off_packed = off_fail + 2 * fail_bytes;
    const size_t packed_bytes = (size_t)a.n * 2 * ps[0].K.pk_words * sizeof(uint2);
    const size_t off_meta = align256(off_packed + packed_bytes);
    const size_t meta_bytes = bt ? slots * ((size_t)MS + 1) * sizeof(uint4) : 0;
    const size_t off_arena = align256(off_meta + meta_bytes);
    const size_t arena_bytes = bt ? slots * (size_t)arena_cells * sizeof(uint2) : 0;
    const size_t off_warp = align256(off_arena + arena_bytes);
    int rc = scratch_reserve(sc, off_warp + W.scratch_bytes);
    if (rc != AIM_OK) return rc;
    unsigned char *base = reinterpret_cast<unsigned char *>(sc->buf);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(base);  // per pass: work counter, fail count
    for (int q = 0; q < npass; ++q) {
        LongK &K = ps[q].K;
        K.work_ctr = ctr + 2 * q;
        K.fail_count = ctr + 2 * q + 1;
        K.fail_list = reinterpret_cast<uint32_t *>(base + off_fail + (size_t)q * fail_bytes);
        K.in_list = q == 0 ? nullptr : ps[q - 1].K.fail_list;
        K.in_count = q == 0 ? nullptr : ps[q - 1].K.fail_count;
        K.plan = reinterpret_cast<const uint4 *>(cached_plan(sc, ps[q].plan.data(), ps[q].plan.size() * sizeof(uint4)));  // uploaded once per penalty set
        if (!K.plan) return AIM_ERR_CUDA;
        K.dirty = base + off_dirty;
        K.packed = reinterpret_cast<const uint2 *>(base + off_packed);
        K.ops = a.ops;
        K.meta = reinterpret_cast<uint4 *>(base + off_meta);
        K.arena = reinterpret_cast<uint2 *>(base + off_arena);
        K.arena_cells = arena_cells;
    }

    cudaError_t err = cudaMemsetAsync(base, 0, 256, stream);
    if (err == cudaSuccess) err = cudaMemsetAsync(base + off_dirty, 0, a.n, stream);
    if (err == cudaSuccess) {
        const uint64_t warps = std::min<uint64_t>(2ull * a.n, (uint64_t)sc->sm_count * 64);
        const int pgrid = (int)((warps * 32 + 255) / 256);
        pack_kernel<<<pgrid, 256, 0, stream>>>(a.plen, a.tlen, a.patterns, a.texts, a.n, p.read_size, ps[0].K.pk_words,
                                               reinterpret_cast<uint2 *>(base + off_packed), base + off_dirty);
        err = cudaGetLastError();
        if (launches) ++*launches;
    }
    for (int q = 0; q < npass && err == cudaSuccess; ++q) {
        ps[q].fn<<<ps[q].grid, block, ps[q].smem, stream>>>(ps[q].K);
        err = cudaGetLastError();
        if (launches) ++*launches;
    }
    if (err != cudaSuccess) { set_error(std::string("wfa_long launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    // leftovers (window or history outgrown, non-ACGT bytes)
    return wfa_warp_launch(W, base + off_warp, ps[npass - 1].K.fail_list, ps[npass - 1].K.fail_count, stream_v, launches);
}

}  // namespace aim
