// GenASM-DC / GenASM-filter for sm_100a (the aim-genasm submodule's DPU kernels).
//
// Reference: aim-genasm/GenASM/DPU-WRAM-DC/dpu/genasmDC.c (pattern bitmasks :40-88, genasmDC :338-556,
// genasmTB :90-336) and aim-genasm/GenASM/DPU-WRAM-filter/dpu/genasm_filter.c:52-239.  Bitap with k = MAX_SCORE
// error levels over the WHOLE pattern (m-bit vectors of count = (m + 64) / 64 words), text walked from its last
// character to its first; R[d] = del & sub & ins & match with del = oldR[d-1], sub = oldR[d-1] << 1,
// ins = R[d-1] << 1 (the NEW R of the level below), match = (oldR[d] << 1) | mask[c].
//
// B200 mapping.  `ins` chains the levels inside one text step, so the levels run as a SYSTOLIC array over the
// lanes of a sub-warp: lane j owns levels j*LPL .. j*LPL+LPL-1 in registers and works one text step behind lane
// j-1, from which it pulls (two shuffles per word) that lane's top level after steps u and u-1.  G = 4..32 lanes
// per pair, 32/G pairs per warp, persistent warps striding over the pairs.  The pattern bitmasks and the text's
// 2-bit codes sit in shared memory.  The reference's traceback matrix (4 vectors per text step and level,
// genasmDC.c:386,461-470,515-525) is not stored: all four are shifts of R, so only R[level][text step] streams to
// an HBM arena, batch by batch - and of R only the WINDOW the traceback can touch: its pattern bit cp and text row ct
// obey |cp + ct - (m - 1)| <= k, so row ti needs bits [m - 2 - k - ti, m + k - ti] (2k + 3 bits: one 32-bit word per
// level and text step at k = 5 instead of the reference's four 128-bit vectors) -, and the traceback is a SECOND kernel with one pair per
// THREAD (the walk is serial per pair: inside the fill kernel it ran on 1 lane in G and took 60 % of the time) that
// reads the four BITS it tests per step (genasmDC.c:107-322) straight from the arena with branch-free, independent
// loads; row n (the state before any text) is known in closed form and the pattern-mask bit is one byte compare.
// The CIGAR leaves in the reference's own format: run lengths with their decimal digits REVERSED
// (genasmDC.c:128-135), NUL-terminated, max_operations = strlen + 1.
//
// Pairs whose reference output depends on memory the reference never wrote for them get a status instead
// (AIM_STATUS_GENASM_UNDEFINED: a text byte outside ACGTacgt skips the traceback rows of that step; the traceback
// walks to text row n); DESIGN.md lists the cases.  "No alignment found" (genasmDC.c:543-547): score -1,
// AIM_STATUS_GENASM_NOALIGN.  variant 1 = DPU-MRAM-DC: substitutions print as 'S' and a pattern 'N' is no wildcard.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "aim_internal.h"

namespace aim {

namespace {

typedef unsigned long long u64;
constexpr unsigned kFullMask = 0xffffffffu;

struct GenK {
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    uint32_t *hist;     // DC only: per pair of the batch [level][text index][ww] 32-bit words: the WINDOW of R the traceback can touch
    size_t hist_stride; // 32-bit words per pair
    int ww;             // window width in 32-bit words: 32 * ww >= 2k + 3 (banded fill: the band, 32 * ww >= 4k + 3)
    int tmode;          // banded fill: a level's history row is indexed by the TICK (t = n - 1 - ti + level) so that a warp stores eight ticks of
                        // every lane as one aligned 32-byte piece; 0: indexed by the text index ti
    int row_len;        // entries per level row
    int lpl;            // levels per lane of the fill kernel (tick of (ti, level) = n - 1 - ti + level / lpl)
    int kk;             // row ti of the history starts at vector bit m - 2 - kk - ti (kk = k: the traceback's window)
    int32_t *meta;      // DC only: per pair of the batch, min_error | text_ok << 16 (fill kernel -> traceback kernel)
    uint32_t first, n;  // pairs [first, first + n) of the launch's batch
    uint32_t idx_base;
    int k, read_size, G, variant;
    int match, mismatch, gap_oe, gap_e;
    uint32_t slot_bytes; // shared memory per pair slot: 4 * W bitmask words, then read_size text codes
    int pm_words;        // banded fill: 32-bit words per base row (ones padding on both sides of the bitmask)
};

__device__ __forceinline__ int base_code(int c)
{   // genasmDC.c:57-74: upper or lower case A, C, G, T; anything else is 4
    c &= ~0x20;
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4;
}

template <int W>
__device__ __forceinline__ void shl1(u64 (&dst)[W], const u64 (&src)[W])
{
#pragma unroll
    for (int w = 0; w < W; ++w) dst[w] = (src[w] << 1) | (w ? src[w - 1] >> 63 : 0ull);
}

// 32-bit half j of a W-word vector (half 0 = bits 0..31); 0 outside
template <int W>
__device__ __forceinline__ uint32_t half_at(const u64 (&v)[W], int j)
{
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 2 * W; ++q) {
        const uint32_t hq = (q & 1) ? (uint32_t)(v[q >> 1] >> 32) : (uint32_t)v[q >> 1];
        r = j == q ? hq : r;
    }
    return r;
}

template <int W, int LPL, bool DC>
__global__ void __launch_bounds__(128) genasm_kernel(const GenK K)
{
    extern __shared__ __align__(16) unsigned char smem_g[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int G = K.G, PPW = 32 / G;
    const int sub = lane / G, sl = lane - sub * G;
    unsigned char *slot = smem_g + (size_t)(wib * PPW + sub) * K.slot_bytes;
    u64 *pm = reinterpret_cast<u64 *>(slot);
    unsigned char *codes = slot + 4 * W * 8;
    const uint32_t wpb = blockDim.x >> 5;
    const uint32_t slot_global = (blockIdx.x * wpb + wib) * PPW + sub;
    const uint32_t nslots = gridDim.x * wpb * PPW;
    const int k = K.k, RS = K.read_size;

    for (uint32_t base_i = 0; base_i < K.n; base_i += nslots) {  // warp-uniform trip count
        const uint32_t li = base_i + slot_global;  // pair of the batch
        const bool active = li < K.n;
        const uint32_t i = K.first + (active ? li : 0);
        const int m = active ? min(max(K.plen[i], 0), RS) : 0;
        const int n = active ? min(max(K.tlen[i], 0), RS) : 0;
        const int count = (m + 64) / 64;
        const char *gp = K.patterns + (size_t)i * RS, *gt = K.texts + (size_t)i * RS;

        // ---- stage: text codes, all-ones bitmasks, then clear the pattern's bits (genasmDC.c:40-88) ----
        for (int q = sl; q < 4 * W; q += G) pm[q] = ~0ull;
        bool text_ok = true;
        for (int j8 = sl; j8 * 8 < n; j8 += G) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(gt + j8 * 8));
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c0 = base_code((int)((v.x >> (8 * b)) & 0xffu)), c1 = base_code((int)((v.y >> (8 * b)) & 0xffu));
                if (j8 * 8 + b < n && c0 > 3) text_ok = false;
                if (j8 * 8 + 4 + b < n && c1 > 3) text_ok = false;
                lo |= (uint32_t)c0 << (8 * b);
                hi |= (uint32_t)c1 << (8 * b);
            }
            *reinterpret_cast<uint2 *>(codes + j8 * 8) = make_uint2(lo, hi);
        }
        __syncwarp();
        for (int j = sl; j < m; j += G) {
            const int ch = gp[j], c = base_code(ch), b = m - 1 - j;
            unsigned *w32 = reinterpret_cast<unsigned *>(pm) + (b >> 5);  // 32-bit half of word b / 64 (little-endian words)
            const unsigned clr = ~(1u << (b & 31));
            if (c < 4) atomicAnd(w32 + c * W * 2, clr);
            else if ((ch & ~0x20) == 'N' && K.variant == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) atomicAnd(w32 + q * W * 2, clr);
            }
        }
        // the whole sub-warp must agree on text_ok
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int o = __shfl_xor_sync(kFullMask, (int)text_ok, off);
            if (off < G) text_ok = text_ok && o;
        }
        __syncwarp();

        // ---- initial state (genasmDC.c:401-425): level d = all ones shifted left by d ----
        u64 cur[LPL][W], lo_old[W];
#pragma unroll
        for (int l = 0; l < LPL; ++l) {
            const int d = sl * LPL + l;
#pragma unroll
            for (int w = 0; w < W; ++w) cur[l][w] = d >= 64 * (w + 1) ? 0ull : (d <= 64 * w ? ~0ull : (~0ull << (d - 64 * w)));
        }
        {   // what the lane below holds before any step: its top level's initial state
            const int d = sl * LPL - 1;
#pragma unroll
            for (int w = 0; w < W; ++w) lo_old[w] = d >= 64 * (w + 1) ? 0ull : (d <= 64 * w ? ~0ull : (~0ull << (d - 64 * w)));
        }
        // this lane's history rows: level d at hist + (d*RS + ti) * ww; the pointer walks down with the text index
        const int WWN = K.ww;
        uint32_t *hrow = DC ? K.hist + (size_t)(active ? li : 0) * K.hist_stride + ((size_t)(sl * LPL) * RS + (n - 1)) * WWN : nullptr;
        const bool store = DC && active;
        int wbase = m - 2 - k - (n - 1);  // first bit of the window of text index ti = n - 1 - u: m - 2 - k - ti

        // ---- the fill: tick t, lane j is on text step u = t - j, i.e. text index n - 1 - u (genasmDC.c:428-527).
        // The lane below is one step ahead: its top level after step u arrives by shuffle, and its state after step
        // u - 1 is what arrived one tick earlier. ----
        const int tmax = __reduce_max_sync(kFullMask, n) + G - 1;
        int u = -sl;
        // the text code and the bitmask words of a step are fetched one and two ticks ahead (shared-memory latency off the
        // critical path): c1/pm1 belong to step u, c2 to step u + 1.  Code 4 = skip the step (also outside the text).
        auto code_at = [&](int uu) -> int { return (uu >= 0 && uu < n) ? (int)codes[n - 1 - uu] : 4; };
        int c1 = code_at(u), c2 = code_at(u + 1);
        u64 pm1[W];
#pragma unroll
        for (int w = 0; w < W; ++w) pm1[w] = pm[(c1 & 3) * W + w];
        for (int t = 0; t < tmax; ++t, ++u) {
            u64 lo_new[W];
#pragma unroll
            for (int w = 0; w < W; ++w) lo_new[w] = __shfl_up_sync(kFullMask, cur[LPL - 1][w], 1, G);
            u64 keep[W], pmw[W];
#pragma unroll
            for (int w = 0; w < W; ++w) { keep[w] = lo_new[w]; pmw[w] = pm1[w]; }
            const int c = c1;
            c1 = c2;
#pragma unroll
            for (int w = 0; w < W; ++w) pm1[w] = pm[(c1 & 3) * W + w];
            c2 = code_at(u + 2);
            if (u >= 0 && u < n) {
                if (c < 4) {  // else: the reference skips the step, R is unchanged
#pragma unroll
                    for (int l = 0; l < LPL; ++l) {
                        const int d = sl * LPL + l;
                        u64 old_d[W], mat[W], nw[W];
#pragma unroll
                        for (int w = 0; w < W; ++w) old_d[w] = cur[l][w];
                        shl1<W>(mat, old_d);
#pragma unroll
                        for (int w = 0; w < W; ++w) mat[w] |= pmw[w];
                        if (d == 0) {
#pragma unroll
                            for (int w = 0; w < W; ++w) nw[w] = mat[w];
                        } else {
                            u64 sb[W], in[W];
                            shl1<W>(sb, lo_old);
                            shl1<W>(in, lo_new);
#pragma unroll
                            for (int w = 0; w < W; ++w) nw[w] = lo_old[w] & sb[w] & in[w] & mat[w];
                        }
#pragma unroll
                        for (int w = 0; w < W; ++w) { lo_old[w] = old_d[w]; lo_new[w] = nw[w]; cur[l][w] = nw[w]; }
                        if (store && d <= k) {
                            uint32_t *h = hrow + (size_t)l * RS * WWN;
                            const int idx = wbase >> 5, sh = wbase & 31;  // floor / non-negative remainder also below bit 0
                            uint32_t lo = half_at<W>(nw, idx);
#pragma unroll 1
                            for (int q = 0; q < WWN; ++q) {
                                const uint32_t hi = half_at<W>(nw, idx + q + 1);
                                h[q] = __funnelshift_r(lo, hi, sh);
                                lo = hi;
                            }
                        }
                    }
                }
                hrow -= WWN;
                ++wbase;
            }
#pragma unroll
            for (int w = 0; w < W; ++w) lo_old[w] = keep[w];
        }

        // ---- lowest level whose end bit is clear (genasmDC.c:389-399,530-541: word 0 of the reference = word count-1) ----
        const int end_bit = (count - 1) * 64 + ((m & 63) ? (m & 63) - 1 : 63);
        int lvl = 1 << 20;
#pragma unroll
        for (int l = LPL - 1; l >= 0; --l) {
            const int d = sl * LPL + l;
            u64 word = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) if (w == (end_bit >> 6)) word = cur[l][w];
            if (d <= k && !((word >> (end_bit & 63)) & 1ull)) lvl = d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int o = __shfl_xor_sync(kFullMask, lvl, off);
            if (off < G) lvl = min(lvl, o);
        }
        const int min_error = lvl > k ? -1 : lvl;

        if (active && sl == 0) {
            if (DC) K.meta[li] = (min_error & 0xffff) | ((int)text_ok << 16);
            else {
                aim_result r;
                r.max_operations = 0;
                r.begin_offset = 0;
                r.end_offset = 0;
                r.score = min_error;  // genasm_filter.c:226-238
                r.status = AIM_STATUS_OK;
                r.idx = K.idx_base + i;
                K.results[i] = r;
            }
        }
        __syncwarp();
    }
}

// ---- banded fill (k <= 31) ----------------------------------------------------------------------------------------
// Every bit of R moves along a DIAGONAL: bit b after tau steps depends on bit b-1 after tau-1 steps of the same level (match)
// and of the level below (substitution), on bit b-1 after tau steps of the level below (insertion) and on bit b after tau-1
// steps of the level below (deletion).  With q = b - tau: same q, same q, q-1, q+1.  So a level-d bit depends only on the
// level-0 bits within d diagonals of its own, and a level-0 bit only on its own diagonal (its start is known: the initial
// ones << d, or the zero a shift brings in at bit 0).  The bits the reference's OUTPUT depends on - the end test at bit m-1
// after n steps and the traceback's 2k+3-wide window - lie within k+1 diagonals of q = m-1-n; hence a band of 4k+3
// diagonals, bits outside it taken as anything, reproduces them EXACTLY at every level (an error entering at the band's edge
// moves in by one diagonal per level and needs k+1 levels to reach the window), with one 32-bit word per level at k <= 7
// instead of count 64-bit words with carries.  Band bit j <-> diagonal Q0 + j, Q0 = m - n - 2 - 2k; at time tau it is
// vector bit b = Q0 + j + tau; b < 0 reads as 0 (what `<< 1` shifts in), the bitmask window of a step is cut from the
// pattern bitmasks in shared memory (padded with all-ones words) by one funnel shift per word.
// A text byte outside ACGTacgt is skipped by the reference without advancing tau (genasmDC.c:430-433), so such texts are
// compacted first (filter only: DC reports those pairs as undefined).
template <int BW, int LPL, bool DC>
__global__ void __launch_bounds__(128) genasm_band_kernel(const GenK K)
{
    constexpr int WS = BW >= 4 ? BW / 2 : 1;  // history words per level and tick: the traceback's 2k+3-bit window, band bits k .. k + 32 WS - 1
    extern __shared__ __align__(16) unsigned char smem_g[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int G = K.G, PPW = 32 / G;
    const int sub = lane / G, sl = lane - sub * G;
    unsigned char *slot = smem_g + (size_t)(wib * PPW + sub) * K.slot_bytes;
    const int PW = K.pm_words;  // 32-bit words per base: BW ones words, the bitmask, BW + 1 ones words
    uint32_t *pm = reinterpret_cast<uint32_t *>(slot);
    unsigned char *codes = slot + 4 * PW * 4;
    const uint32_t wpb = blockDim.x >> 5;
    const uint32_t slot_global = (blockIdx.x * wpb + wib) * PPW + sub;
    const uint32_t nslots = gridDim.x * wpb * PPW;
    const int k = K.k, RS = K.read_size;

    auto ones_from = [](int pos, int w) -> uint32_t {  // word w of "bits >= pos set"
        const int p = pos - 32 * w;
        return p <= 0 ? ~0u : (p >= 32 ? 0u : (~0u << p));
    };

    for (uint32_t base_i = 0; base_i < K.n; base_i += nslots) {  // warp-uniform trip count
        const uint32_t li = base_i + slot_global;
        const bool active = li < K.n;
        const uint32_t i = K.first + (active ? li : 0);
        const int m = active ? min(max(K.plen[i], 0), RS) : 0;
        int n = active ? min(max(K.tlen[i], 0), RS) : 0;
        const char *gp = K.patterns + (size_t)i * RS, *gt = K.texts + (size_t)i * RS;

        // ---- stage: text codes, all-ones bitmask rows, then clear the pattern's bits (genasmDC.c:40-88) ----
        for (int q = sl; q < 4 * PW; q += G) pm[q] = ~0u;
        bool text_ok = true;
        for (int j8 = sl; j8 * 8 < n; j8 += G) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(gt + j8 * 8));
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c0 = base_code((int)((v.x >> (8 * b)) & 0xffu)), c1 = base_code((int)((v.y >> (8 * b)) & 0xffu));
                if (j8 * 8 + b < n && c0 > 3) text_ok = false;
                if (j8 * 8 + 4 + b < n && c1 > 3) text_ok = false;
                lo |= (uint32_t)c0 << (8 * b);
                hi |= (uint32_t)c1 << (8 * b);
            }
            *reinterpret_cast<uint2 *>(codes + j8 * 8) = make_uint2(lo, hi);
        }
        __syncwarp();
        for (int j = sl; j < m; j += G) {
            const int ch = gp[j], c = base_code(ch), b = m - 1 - j;
            uint32_t *w32 = pm + BW + (b >> 5);
            const unsigned clr = ~(1u << (b & 31));
            if (c < 4) atomicAnd(w32 + c * PW, clr);
            else if ((ch & ~0x20) == 'N' && K.variant == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) atomicAnd(w32 + q * PW, clr);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int o = __shfl_xor_sync(kFullMask, (int)text_ok, off);
            if (off < G) text_ok = text_ok && o;
        }
        __syncwarp();
        if (__any_sync(kFullMask, !text_ok)) {  // (warp-uniform branch) the reference skips those steps: drop the bytes, tau counts the steps that happen
            int kept = n;
            if (!text_ok && sl == 0) {
                kept = 0;
                for (int j = 0; j < n; ++j) { const int c = codes[j]; if (c < 4) codes[kept++] = (unsigned char)c; }
            }
            kept = __shfl_sync(kFullMask, kept, sub * G);
            if (!text_ok) n = kept;
            __syncwarp();
        }

        const int Q0 = m - n - 2 - 2 * k;
        // ---- initial state (genasmDC.c:401-425): vector bit b of level d is (b >= d) ----
        uint32_t cur[LPL][BW], lo_old[BW];
#pragma unroll
        for (int l = 0; l < LPL; ++l)
#pragma unroll
            for (int w = 0; w < BW; ++w) cur[l][w] = ones_from(sl * LPL + l - Q0, w);
#pragma unroll
        for (int w = 0; w < BW; ++w) lo_old[w] = ones_from(sl * LPL - 1 - Q0, w);
        // history rows are indexed by the tick: all lanes store the same eight ticks as one aligned 32 BW-byte piece (a 4-byte store
        // per lane and tick costs a whole 32-byte sector write between L1 and L2, which is what bounded the fill before)
        uint32_t *hrow = DC ? K.hist + (size_t)(active ? li : 0) * K.hist_stride + (size_t)(sl * LPL) * K.row_len * WS : nullptr;
        const bool store = DC && active && text_ok;

        const int tmax = (__reduce_max_sync(kFullMask, n) + G - 1 + 7) & ~7;
        int u = -sl;
        const int b0_min = -32 * BW, b0_max = (PW - 2 * BW - 1) * 32;  // window starts the padded rows can serve
        for (int t0 = 0; t0 < tmax; t0 += 8) {
            uint32_t buf[LPL][8][WS];
#pragma unroll
            for (int tj = 0; tj < 8; ++tj, ++u) {
                uint32_t lo_new[BW], keep[BW];
#pragma unroll
                for (int w = 0; w < BW; ++w) { lo_new[w] = __shfl_up_sync(kFullMask, cur[LPL - 1][w], 1, G); keep[w] = lo_new[w]; }
                if (u >= 0 && u < n) {
                    const int tau = u + 1;
                    const int c = codes[n - 1 - u];
                    // bitmask bits b0 .. b0 + 32 BW - 1 of the step's character; below b0_min everything is virtual, above b0_max all ones
                    const int b0 = min(max(Q0 + tau, b0_min), b0_max);
                    const uint32_t *pw = pm + c * PW + BW + (b0 >> 5);
                    const int sh = b0 & 31;
                    uint32_t mw[BW];
                    uint32_t plo = pw[0];
#pragma unroll
                    for (int w = 0; w < BW; ++w) { const uint32_t phi = pw[w + 1]; mw[w] = __funnelshift_r(plo, phi, sh); plo = phi; }
#pragma unroll
                    for (int l = 0; l < LPL; ++l) {
                        const int d = sl * LPL + l;
                        uint32_t old_d[BW], nw[BW];
#pragma unroll
                        for (int w = 0; w < BW; ++w) { old_d[w] = cur[l][w]; nw[w] = old_d[w] | mw[w]; }  // match: same diagonal
                        if (d > 0) {
#pragma unroll
                            for (int w = 0; w < BW; ++w) {
                                const uint32_t ins = (lo_new[w] << 1) | (w ? lo_new[w - 1] >> 31 : 1u);           // diagonal q - 1, new
                                const uint32_t del = (lo_old[w] >> 1) | (w + 1 < BW ? lo_old[w + 1] << 31 : 0x80000000u);  // diagonal q + 1, old
                                nw[w] &= lo_old[w] & ins & del;                                                     // substitution: same diagonal, old
                            }
                        }
#pragma unroll
                        for (int w = 0; w < BW; ++w) {
                            nw[w] &= ones_from(-Q0 - tau, w);  // vector bits below 0 do not exist: they read as the 0 a shift brings in
                            lo_old[w] = old_d[w]; lo_new[w] = nw[w]; cur[l][w] = nw[w];
                        }
                    }
                }
#pragma unroll
                for (int l = 0; l < LPL; ++l)
#pragma unroll
                    for (int w = 0; w < WS; ++w) buf[l][tj][w] = __funnelshift_r(cur[l][w], w + 1 < BW ? cur[l][w + 1] : 0u, k);  // k < 32
#pragma unroll
                for (int w = 0; w < BW; ++w) lo_old[w] = keep[w];
            }
            if (store) {
#pragma unroll
                for (int l = 0; l < LPL; ++l) {
                    if (sl * LPL + l <= k) {
                        uint4 *h4 = reinterpret_cast<uint4 *>(hrow + ((size_t)l * K.row_len + t0) * WS);
                        const uint32_t *bf = &buf[l][0][0];
#pragma unroll
                        for (int q = 0; q < 2 * WS; ++q) h4[q] = make_uint4(bf[4 * q], bf[4 * q + 1], bf[4 * q + 2], bf[4 * q + 3]);
                    }
                }
            }
        }

        // ---- lowest level whose end bit is clear (genasmDC.c:389-399,530-541).  The reference tests bit m-1 of the vector, except
        // when m % 64 == 0, where it tests bit m+63, which no step ever clears: no alignment is found then. ----
        const int jend = 2 * k + 1;  // (m - 1) - n - Q0
        int lvl = 1 << 20;
        if ((m & 63) != 0) {
#pragma unroll
            for (int l = LPL - 1; l >= 0; --l) {
                const int d = sl * LPL + l;
                uint32_t word = 0;
#pragma unroll
                for (int w = 0; w < BW; ++w) if (w == (jend >> 5)) word = cur[l][w];
                if (d <= k && !((word >> (jend & 31)) & 1u)) lvl = d;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int o = __shfl_xor_sync(kFullMask, lvl, off);
            if (off < G) lvl = min(lvl, o);
        }
        const int min_error = lvl > k ? -1 : lvl;

        if (active && sl == 0) {
            if (DC) K.meta[li] = (min_error & 0xffff) | ((int)text_ok << 16);
            else {
                aim_result r;
                r.max_operations = 0;
                r.begin_offset = 0;
                r.end_offset = 0;
                r.score = min_error;  // genasm_filter.c:226-238
                r.status = AIM_STATUS_OK;
                r.idx = K.idx_base + i;
                K.results[i] = r;
            }
        }
        __syncwarp();
    }
}

// genasmTB (genasmDC.c:90-336), one pair per thread.  Bit b of R[d] after text index ti comes from the fill kernel's
// arena; ti == n is the initial state (closed form), b < 0 is the zero a left shift brings in.
__global__ void __launch_bounds__(128, 12) genasm_tb_kernel(const GenK K)
{
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= K.n) return;
    const uint32_t i = K.first + li;
    const int RS = K.read_size;
    const int m = min(max(K.plen[i], 0), RS), n = min(max(K.tlen[i], 0), RS);
    const int count = (m + 64) / 64;
    const int meta = K.meta[li];
    const int min_error = (int)(short)(meta & 0xffff);
    const bool text_ok = (meta >> 16) & 1;
    const uint32_t *hist = K.hist + (size_t)li * K.hist_stride;
    const int WWN = K.ww, k = K.k;
    const char *gp = K.patterns + (size_t)i * RS, *gt = K.texts + (size_t)i * RS;
    char *cig = K.ops + (size_t)i * 2 * RS;
    const int cap = 2 * RS;

    aim_result r;
    r.max_operations = m + n;
    r.begin_offset = 0;
    r.end_offset = 0;
    r.score = -1;
    r.status = AIM_STATUS_OK;
    r.idx = K.idx_base + i;
    if (min_error < 0) { r.status = AIM_STATUS_GENASM_NOALIGN; cig[0] = '\0'; }
    else if (!text_ok) { r.status = AIM_STATUS_GENASM_UNDEFINED; cig[0] = '\0'; }
    else {
        // branch-free: the load always happens (indices clamped into the pair's rows), the special cases are selects
        auto hbit = [&](int ti, int d, int b) -> unsigned {
            const int tc = min(ti, n - 1);
            const int wb = min(max(b - (m - 2 - K.kk - tc), 0), 32 * WWN - 1);  // bit of row tc's window
            const int row = K.tmode ? n - 1 - tc + d / K.lpl : tc;  // banded fill: rows are indexed by the tick
            const uint32_t wv = hist[((size_t)d * K.row_len + row) * WWN + (wb >> 5)];
            unsigned v = (wv >> (wb & 31)) & 1u;
            v = ti >= n ? (b >= d ? 1u : 0u) : v;
            return b < 0 ? 0u : v;
        };
        int cp = m - 1, ct = 0, ce = min_error, c = 0;
        int nM = 0, nS = 0, nOpen = 0, nExt = 0, run = 0;
        char last = '0';
        bool first = true, undefined = false;
        const char sub_ch = K.variant ? 'S' : 'X';
        auto flush = [&]() {
            if (first) return;
            int num = run;
            while (num != 0 && c < cap - 2) { cig[c++] = (char)('0' + num % 10); num /= 10; }
            if (c < cap - 1) cig[c++] = last;
        };
        // the pattern-mask bit of (pattern position, text character) is clear: same base in either case (the text is ACGTacgt
        // here), or a pattern 'N' (genasmDC.c:75-81)
        auto mask_hit = [&](int pch, int tch) -> bool { return (((pch ^ tch) & ~0x20) == 0) || ((pch & ~0x20) == 'N' && K.variant == 0); };
        while (cp >= 0 && ce >= 0) {
            if (ct >= n) { undefined = true; break; }
            // ---- a RUN of matches.  After a match, a substitution or at the start the reference's first two tests (affine insertion /
            // deletion, :107,:142) cannot fire (they need last == I / D), so the match bit alone decides: one history bit and two
            // bytes per step, in a loop of its own - the threads of a warp reach their edits at different steps, and inside one
            // common loop every step paid for the (rare) edit path of some other thread ----
            if (last != 'I' && last != 'D') {
                int cnt = 0;
                while (cp >= 0 && ct < n) {
                    const unsigned a = ce == 0 ? hbit(ct, 0, cp) : hbit(ct + 1, ce, cp - 1);
                    const unsigned t0r = ce == 0 ? a : (a | (mask_hit(gp[m - 1 - cp], gt[ct]) ? 0u : 1u));
                    if (t0r != 0) break;
                    ++ct; --cp; ++cnt;
                }
                if (cnt) {
                    if (last == 'M') run += cnt; else { flush(); run = cnt; last = 'M'; }
                    nM += cnt;
                    first = false;
                }
                if (cp < 0) break;
                if (ct >= n) { undefined = true; break; }
            }
            // ---- one general step: an edit, or anything after an insertion / deletion ----
            const bool z = ce == 0;
            const int dl = max(ce - 1, 0);
            // level 0 keeps only R itself (genasmDC.c:461-470): match = its bit, the other three read as set
            const unsigned a = z ? hbit(ct, 0, cp) : hbit(ct + 1, ce, cp - 1);
            const bool hit = mask_hit(gp[m - 1 - cp], gt[ct]);
            const unsigned t0 = z ? a : (a | (hit ? 0u : 1u));
            unsigned t1 = 1u, t2 = 1u, t3 = 1u;
            if (!z && (t0 != 0 || last == 'I' || last == 'D')) {
                t1 = hbit(ct + 1, dl, cp - 1);
                t2 = hbit(ct, dl, cp - 1);
                t3 = hbit(ct + 1, dl, cp);
            }
            if (last == 'I' && t2 == 0) { --cp; --ce; ++run; ++nExt; }
            else if (last == 'D' && t3 == 0) { ++ct; --ce; ++run; ++nExt; }
            else if (t0 == 0) {
                ++ct; --cp;
                if (last == 'M') ++run; else { flush(); run = 1; last = 'M'; }
                ++nM;
            } else if (t1 == 0) {
                ++ct; --cp; --ce;
                if (last == sub_ch) ++run; else { flush(); run = 1; last = sub_ch; }
                ++nS;
            } else if (t3 == 0) { ++ct; --ce; flush(); run = 1; last = 'D'; ++nOpen; }
            else if (t2 == 0) { --cp; --ce; flush(); run = 1; last = 'I'; ++nOpen; }
            else { undefined = true; break; }  // the reference would spin forever
            first = false;
        }
        if (undefined) { r.status = AIM_STATUS_GENASM_UNDEFINED; cig[0] = '\0'; }
        else {
            int num = run;
            while (num != 0 && c < cap - 2) { cig[c++] = (char)('0' + num % 10); num /= 10; }
            if (c < cap - 1) cig[c++] = last;
            cig[c] = '\0';
            r.max_operations = c + 1;
            r.end_offset = c;
            r.score = nM * K.match + nS * K.mismatch + nOpen * K.gap_oe + nExt * K.gap_e;
        }
    }
    K.results[i] = r;
}

template <int W, int LPL>
cudaError_t launch_wl(const GenK &K, bool dc, int grid, size_t smem, cudaStream_t st)
{
    if (dc) genasm_kernel<W, LPL, true><<<grid, 128, smem, st>>>(K);
    else genasm_kernel<W, LPL, false><<<grid, 128, smem, st>>>(K);
    return cudaGetLastError();
}
template <int W>
cudaError_t launch_w(const GenK &K, int lpl, bool dc, int grid, size_t smem, cudaStream_t st)
{
    if (lpl == 1) return launch_wl<W, 1>(K, dc, grid, smem, st);
    if (lpl == 2) return launch_wl<W, 2>(K, dc, grid, smem, st);
    return launch_wl<W, 4>(K, dc, grid, smem, st);
}

template <int BW, int LPL>
cudaError_t launch_bl(const GenK &K, bool dc, int grid, size_t smem, cudaStream_t st)
{
    if (dc) genasm_band_kernel<BW, LPL, true><<<grid, 128, smem, st>>>(K);
    else genasm_band_kernel<BW, LPL, false><<<grid, 128, smem, st>>>(K);
    return cudaGetLastError();
}
cudaError_t launch_band(const GenK &K, int bw, int lpl, bool dc, int grid, size_t smem, cudaStream_t st)
{   // k <= 30: one, two or (one-word bands) four levels per lane
    if (lpl == 4) return launch_bl<1, 4>(K, dc, grid, smem, st);
    if (lpl == 2) {
        if (bw == 1) return launch_bl<1, 2>(K, dc, grid, smem, st);
        if (bw == 2) return launch_bl<2, 2>(K, dc, grid, smem, st);
        return launch_bl<4, 2>(K, dc, grid, smem, st);
    }
    if (bw == 1) return launch_bl<1, 1>(K, dc, grid, smem, st);
    if (bw == 2) return launch_bl<2, 1>(K, dc, grid, smem, st);
    return launch_bl<4, 1>(K, dc, grid, smem, st);
}

}  // namespace

int launch_genasm(const KernelArgs &a, Scratch *sc, void *stream_v, int *launches)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    const aim_params &p = a.p;
    const bool dc = p.algo == AIM_ALGO_GENASM_DC;
    const int k = p.max_score;
    const int count_max = (p.read_size + 64) / 64;
    if (k > 126) { set_error("GenASM: MAX_SCORE above 126 is not served by the B200 kernel"); return AIM_ERR_ARG; }
    if (count_max > 8) { set_error("GenASM: READ_SIZE above 448 is not served by the B200 kernel"); return AIM_ERR_ARG; }
    if (a.n == 0) return AIM_OK;
    // banded fill (k <= 30) unless switched off; it takes two levels per lane when that halves the lanes per pair (a lane's second
    // level costs a fifth of a tick: the shuffle, the bitmask window and the loop are shared)
    const bool band = k <= 30 && !getenv("AIM_GENASM_FULL");
    int lpl = k + 1 <= 32 ? 1 : (k + 1 <= 64 ? 2 : 4);
    int G = 4;
    while (G * lpl < k + 1) G *= 2;
    if (band && !getenv("AIM_GENASM_LPL1")) {
        // two levels per lane, two lanes per pair at least (measured at k = 5: 2 levels x 4 lanes 482 M pairs/s, 4 levels x 2 lanes 470 M,
        // 1 level x 8 lanes 420 M; at k = 1: 1 level x 2 lanes 1.73 G against 1.13 G with 4 lanes)
        lpl = k + 1 <= 2 ? 1 : 2;
        G = 2;
        while (G * lpl < k + 1) G *= 2;
        if (k > 7 && G < 4) G = 4;
    }
    const int W = count_max <= 2 ? 2 : (count_max <= 4 ? 4 : 8);

    GenK K{};
    K.plen = a.plen; K.tlen = a.tlen; K.patterns = a.patterns; K.texts = a.texts; K.results = a.results; K.ops = a.ops;
    K.idx_base = a.idx_base; K.k = k; K.read_size = p.read_size; K.G = G; K.variant = p.variant;
    K.match = p.match; K.mismatch = p.mismatch; K.gap_oe = p.gap_open + p.gap_ext; K.gap_e = p.gap_ext;
    K.slot_bytes = (uint32_t)(4 * W * 8 + p.read_size);
    const int PPW = 32 / G;
    const size_t smem = (size_t)4 * PPW * K.slot_bytes;
    const int blocks_per_sm = band ? 16 : (W == 8 || lpl == 4 ? 4 : (W == 4 || lpl == 2 ? 8 : 16));
    // banded fill unless switched off: band of 4k + 3 bits in BW words, of which the traceback's 2k + 3-bit window (band bits k ..)
    // goes to the history in WS words: k <= 7: 1 / 1, k <= 14: 2 / 1, k <= 30: 4 / 2; beyond, full vectors with the window cut out
    const int bw = k <= 7 ? 1 : (k <= 14 ? 2 : 4);
    K.kk = k;
    K.ww = 1;
    if (band) K.ww = bw >= 4 ? bw / 2 : 1;
    else while (32 * K.ww < 2 * k + 3) K.ww *= 2;
    if (band) {
        K.pm_words = 2 * bw + 1 + 2 * count_max;
        K.slot_bytes = (uint32_t)(4 * K.pm_words * 4 + p.read_size);
    }
    K.tmode = band ? 1 : 0;
    K.lpl = lpl;
    K.row_len = band ? ((p.read_size + G + 7) & ~7) + 8 : p.read_size;
    K.hist_stride = dc ? (size_t)(k + 1) * (size_t)K.row_len * (size_t)K.ww : 0;

    // DC: the history of a whole batch lives in the arena between the fill and the traceback kernels.  Two arena halves:
    // the traceback of batch b (latency-bound, one thread per pair) runs on a side stream under the fill of batch b + 1
    // (issue-bound), so a launch is cut into at least four batches when it is large enough.
    size_t budget = (size_t)4 << 30;
    if (const char *bm = getenv("AIM_GENASM_ARENA_MB")) { const long v = atol(bm); if (v >= 16 && v <= 65536) budget = (size_t)v << 20; }
    uint32_t batch = a.n;
    if (dc) {
        const size_t fit = std::max<size_t>(1024, budget / 2 / (K.hist_stride * 4));
        batch = (uint32_t)std::min<size_t>(fit, std::max<size_t>(32768, ((size_t)a.n + 3) / 4));
        batch = std::min(batch, a.n);
    }
    const bool overlap = dc && batch < a.n && !getenv("AIM_GENASM_NO_OVERLAP");
    const int halves = overlap ? 2 : 1;
    const size_t meta_bytes = dc ? ((size_t)batch * 4 + 255) / 256 * 256 : 0;
    const size_t hist_bytes = ((size_t)batch * K.hist_stride * 4 + 255) / 256 * 256;
    int rc = scratch_reserve(sc, std::max<size_t>((size_t)halves * (meta_bytes + hist_bytes), 256));
    if (rc != AIM_OK) return rc;
    // traceback side stream + its events live in the device's Scratch (aim_shutdown destroys them)
    if (overlap && !sc->side_stream) {
        cudaStream_t st = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        sc->side_stream = st;
        for (int h = 0; h < 4 && e == cudaSuccess; ++h) {
            cudaEvent_t ev = nullptr;
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            sc->side_ev[h] = ev;
        }
        if (e != cudaSuccess) { set_error(std::string("genasm side stream: ") + cudaGetErrorString(e)); return AIM_ERR_CUDA; }
    }
    cudaStream_t const side = (cudaStream_t)sc->side_stream;
    cudaEvent_t const ev_fill[2] = {(cudaEvent_t)sc->side_ev[0], (cudaEvent_t)sc->side_ev[1]};
    cudaEvent_t const ev_tb[2] = {(cudaEvent_t)sc->side_ev[2], (cudaEvent_t)sc->side_ev[3]};
    // (tuning knob: dummy dynamic shared memory caps the traceback kernel's blocks per SM.  Measured at config 7: 0 KB 293 M pairs/s,
    // 24 KB 157 M, 48 KB 212 M, 100 KB 136 M - the shared-memory carve-out shrinks L1 and the walk needs the threads.)
    size_t tb_smem = 0;
    if (const char *e = getenv("AIM_GENASM_TB_SMEM_KB")) { const long v = atol(e); if (v >= 0 && v <= 200) tb_smem = (size_t)v << 10; }
    if (dc && tb_smem > (48u << 10)) {
        if (cudaFuncSetAttribute(genasm_tb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem) != cudaSuccess) { set_error("genasm tb smem attribute"); return AIM_ERR_CUDA; }
    }
    uint32_t nb = 0;
    for (uint32_t first = 0; first < a.n; first += batch, ++nb) {
        const int h = overlap ? (int)(nb & 1u) : 0;
        unsigned char *base = reinterpret_cast<unsigned char *>(sc->buf) + (size_t)h * (meta_bytes + hist_bytes);
        K.meta = reinterpret_cast<int32_t *>(base);
        K.hist = reinterpret_cast<uint32_t *>(base + meta_bytes);
        K.first = first;
        K.n = std::min(batch, a.n - first);
        int grid = sc->sm_count * blocks_per_sm;
        const uint64_t per_block = (uint64_t)4 * PPW;
        if ((uint64_t)grid * per_block > K.n) grid = (int)std::max<uint64_t>(1, (K.n + per_block - 1) / per_block);
        cudaError_t err = cudaSuccess;
        if (overlap && nb >= 2) err = cudaStreamWaitEvent(stream, ev_tb[h], 0);  // the traceback that read this half is done
        if (err == cudaSuccess && band) err = launch_band(K, bw, lpl, dc, grid, (size_t)4 * PPW * K.slot_bytes, stream);
        else if (err == cudaSuccess)
            err = W == 2 ? launch_w<2>(K, lpl, dc, grid, smem, stream)
                         : (W == 4 ? launch_w<4>(K, lpl, dc, grid, smem, stream) : launch_w<8>(K, lpl, dc, grid, smem, stream));
        if (err == cudaSuccess && launches) ++*launches;
        if (err == cudaSuccess && dc) {
            cudaStream_t ts = stream;
            if (overlap) {
                ts = side;
                err = cudaEventRecord(ev_fill[h], stream);
                if (err == cudaSuccess) err = cudaStreamWaitEvent(ts, ev_fill[h], 0);
            }
            if (err == cudaSuccess) {
                genasm_tb_kernel<<<(K.n + 127) / 128, 128, tb_smem, ts>>>(K);
                err = cudaGetLastError();
            }
            if (err == cudaSuccess && overlap) err = cudaEventRecord(ev_tb[h], ts);
            if (err == cudaSuccess && launches) ++*launches;
        }
        if (err != cudaSuccess) { set_error(std::string("genasm launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    }
    if (overlap) {  // the caller's stream continues only after the last tracebacks
        for (int h = 0; h < 2; ++h)
            if (nb > (uint32_t)h && cudaStreamWaitEvent(stream, ev_tb[h], 0) != cudaSuccess) { set_error("genasm: stream join failed"); return AIM_ERR_CUDA; }
    }
    return AIM_OK;
}

}  // namespace aim
