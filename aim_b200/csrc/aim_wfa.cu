// Gap-affine WFA / WFA-adaptive for sm_100a: one pair per warp, lanes over diagonals.
//
// Replaces the DPU tasklet code of WFA/DPU-{WRAM,MRAM}/dpu/wfa.c + wfa_backtracing.c (reference
// checkout paths).  Per pair:
//   * both sequences are read once from HBM as coalesced 8-byte words, packed 2 bits/base
//     (big-endian inside each 32-bit word) into shared memory; a pair holding any byte outside
//     {A,C,G,T} is compared byte-wise straight from global memory instead (the reference compares
//     raw bytes, wfa.c:209);
//   * extend (wfa.c:193-215) compares 16 bases per step with XOR + CLZ on funnel-shifted words;
//   * compute_next/compute_offsets (wfa.c:238-354) run one diagonal per lane with the reference's
//     sentinels (-10, NULL=-16384, "+1 on NULL") kept literally;
//   * adaptive reduction (wfa.c:70-141) is a warp min-reduction plus two ballot scans;
//   * the wavefront history (the reference's WRAM/MRAM component store, dpu_allocator_mram.c) is a
//     bump-allocated int16 arena: in shared memory in short-read mode, in a per-warp, L2-resident
//     HBM arena in long-read mode (HG = true); score-only runs keep a ring of the last
//     max(x, o+e)+1 wavefronts instead;
//   * backtrace (wfa_backtracing.c:219-375) is walked by lane 0 with the reference's tie-break order;
//     the op string is staged in shared memory ('M'-filled, so match runs cost O(1)) and written
//     back with 16-byte stores.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "aim_internal.h"
#include "aim_wfa_common.cuh"

namespace aim {

namespace {

constexpr uint32_t F_PRESENT = 1, F_HAS_I = 2, F_HAS_D = 4;

struct WfaK {
    const int32_t *plen;
    const int32_t *tlen;
    const char *patterns;
    const char *texts;
    aim_result *results;
    char *ops;
    uint32_t n, idx_base;
    int x, o, e;
    int max_score, read_size, backtrace, reduce;
    uint32_t seq_words;        // 32-bit words reserved per packed sequence (multiple of 4)
    uint32_t meta_words;       // 3 * (max_score + 1), rounded to a multiple of 4
    uint32_t hist_cap;         // int16 slots available per pair
    uint32_t ring;             // 0: keep the whole history (bump allocation); else live scores
    uint32_t ring_stride;      // slots per ring entry
    uint32_t warp_smem_bytes;  // shared memory per warp
    uint32_t *g_meta;          // long-read mode: per-warp meta tables
    int16_t *g_hist;           // long-read mode: per-warp history arenas
    size_t g_meta_stride;      // words
    size_t g_hist_stride;      // slots
    const uint32_t *list;      // optional: pair indices to serve (the long-read kernel's leftovers) ...
    const uint32_t *list_count;  // ... and how many (device counter)
};

struct Rec {
    int lo, hi;      // current (possibly trimmed) diagonal range
    int lob;         // lo at allocation time: array origin
    int wlen;        // allocated width
    uint32_t base;   // slot index of M[lob]; I follows at +wlen (if any), then D
    uint32_t flags;  // F_*
};

__device__ __forceinline__ Rec load_rec(const uint32_t *meta, int s)
{
    Rec r;
    uint32_t w0 = meta[3 * s], w1 = meta[3 * s + 1], w2 = meta[3 * s + 2];
    r.lo = (int)(short)(w0 & 0xffffu);
    r.hi = (int)(short)(w0 >> 16);
    r.lob = (int)(short)(w1 & 0xffffu);
    r.wlen = (int)(w1 >> 16);
    r.base = w2 & 0x0fffffffu;
    r.flags = w2 >> 28;
    return r;
}
__device__ __forceinline__ Rec load_rec_or_absent(const uint32_t *meta, int s)
{
    if (s >= 0) return load_rec(meta, s);
    Rec r;
    r.lo = 1; r.hi = -1; r.lob = 0; r.wlen = 0; r.base = 0; r.flags = 0;
    return r;
}
__device__ __forceinline__ void store_rec(uint32_t *meta, int s, const Rec &r)
{
    meta[3 * s] = ((uint32_t)r.lo & 0xffffu) | ((uint32_t)r.hi << 16);
    meta[3 * s + 1] = ((uint32_t)r.lob & 0xffffu) | ((uint32_t)r.wlen << 16);
    meta[3 * s + 2] = r.base | (r.flags << 28);
}

// One pair, executed by a full warp.  HG: history/meta live in global memory (long-read mode).
template <bool HG>
__device__ void wfa_pair(const WfaK &K, const int lane, const uint32_t i, uint32_t *sP, uint32_t *sT,
                         uint32_t *meta, int16_t *hist, char *sOps)
{
    const int RS = K.read_size;
    const int pl = min(max(K.plen[i], 0), RS), tl = min(max(K.tlen[i], 0), RS);
    const char *gp = K.patterns + (size_t)i * RS;
    const char *gt = K.texts + (size_t)i * RS;
    const int ak = tl - pl;
    const int max_ops = pl + tl;

    // ---- stage + pack the sequences (coalesced 8-byte loads) ----
    bool ok = true;
    for (int c = lane; c * 8 < pl; c += 32) {
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(gp) + c);
        uint32_t h16 = pack8(w, pl - c * 8, &ok);
        reinterpret_cast<uint16_t *>(sP)[c ^ 1] = (uint16_t)h16;
    }
    for (int c = lane; c * 8 < tl; c += 32) {
        uint2 w = __ldg(reinterpret_cast<const uint2 *>(gt) + c);
        uint32_t h16 = pack8(w, tl - c * 8, &ok);
        reinterpret_cast<uint16_t *>(sT)[c ^ 1] = (uint16_t)h16;
    }
    const bool packed = __all_sync(kFull, ok);

    // ---- ops staging: 'M' everywhere (wfa.c:499-501) ----
    if (K.backtrace) {
        if (HG) {
            uint4 *g = reinterpret_cast<uint4 *>(K.ops + (size_t)i * 2 * RS);
            for (int c = lane; c < (2 * RS) / 16; c += 32) g[c] = make_uint4(0x4d4d4d4du, 0x4d4d4d4du, 0x4d4d4d4du, 0x4d4d4d4du);
        } else {
            uint32_t *s = reinterpret_cast<uint32_t *>(sOps);
            for (int c = lane; c < (2 * RS) / 4; c += 32) s[c] = 0x4d4d4d4du;
        }
    }

    // ---- score 0 (wfa.c:363-365) ----
    Rec cur;
    cur.lo = cur.hi = cur.lob = 0;
    cur.wlen = 1;
    cur.base = 0;
    cur.flags = F_PRESENT;
    uint32_t top = 1;  // bump pointer (slots)
    if (lane == 0) {
        store_rec(meta, 0, cur);
        hist[0] = 0;
    }
    __syncwarp();

    int s = 0;
    int status = AIM_STATUS_OK;
    bool reached = false;
    for (;;) {
        if (cur.flags & F_PRESENT) {
            int16_t *M = hist + cur.base - cur.lob;
            // extend (wfa.c:193-215)
            for (int k = cur.lo + lane; k <= cur.hi; k += 32) {
                int off = M[k];
                int v = off - k;
                if (off >= 0 && v >= 0) {
                    int lim = min(pl - v, tl - off);
                    if (lim > 0) {
                        int cnt = packed ? match_packed(sP, sT, v, off, lim) : match_bytes(gp, gt, v, off, lim);
                        M[k] = (int16_t)(off + cnt);
                    }
                }
            }
            __syncwarp();
            // adaptive reduction (wfa.c:70-141): min_wavefront_length 10, max_distance_threshold 50
            if (K.reduce && (cur.hi - cur.lo + 1) >= 10) {
                int md = max(pl, tl);
                for (int k = cur.lo + lane; k <= cur.hi; k += 32) {
                    int off = M[k];
                    md = min(md, max(pl - (off - k), tl - off));
                }
                md = warp_min(md);
                const int lo0 = cur.lo, hi0 = cur.hi;
                const int top_limit = min(ak - 1, hi0);
                int newlo = lo0;
                if (lo0 < top_limit) {
                    newlo = top_limit;
                    for (int b = lo0; b < top_limit; b += 32) {
                        int k = b + lane;
                        bool hit = false;
                        if (k < top_limit) {
                            int off = M[k];
                            hit = (max(pl - (off - k), tl - off) - md) <= 50;
                        }
                        unsigned m = __ballot_sync(kFull, hit);
                        if (m) { newlo = b + __ffs(m) - 1; break; }
                    }
                }
                const int bottom_limit = max(ak + 1, newlo);
                int newhi = hi0;
                if (hi0 > bottom_limit) {
                    newhi = bottom_limit;
                    for (int t = hi0; t > bottom_limit; t -= 32) {
                        int k = t - lane;
                        bool hit = false;
                        if (k > bottom_limit) {
                            int off = M[k];
                            hit = (max(pl - (off - k), tl - off) - md) <= 50;
                        }
                        unsigned m = __ballot_sync(kFull, hit);
                        if (m) { newhi = t - (__ffs(m) - 1); break; }
                    }
                }
                // newlo <= newhi always holds (the reference's "klo > khi" branch, wfa.c:132-140, is unreachable)
                if (newlo != lo0 || newhi != hi0) {
                    cur.lo = newlo;
                    cur.hi = newhi;
                    if (lane == 0) meta[3 * s] = ((uint32_t)newlo & 0xffffu) | ((uint32_t)newhi << 16);
                }
            }
            // end reached (wfa.c:217-237)
            if (cur.lo <= ak && ak <= cur.hi && (int)M[ak] >= tl) { reached = true; break; }
        }
        ++s;
        if (s > K.max_score) break;  // give up (wfa.c:399-404): score = MAX_SCORE + 1, ops untouched
        __syncwarp();

        // ---- compute_next (wfa.c:275-354) ----
        const Rec A = load_rec_or_absent(meta, s - K.x);
        const Rec B = load_rec_or_absent(meta, s - K.o - K.e);
        const Rec E = load_rec_or_absent(meta, s - K.e);
        const bool sub_null = !(A.flags & F_PRESENT);
        const bool o_null = !(B.flags & F_PRESENT);
        const bool ie_null = !((E.flags & F_PRESENT) && (E.flags & F_HAS_I));
        const bool de_null = !((E.flags & F_PRESENT) && (E.flags & F_HAS_D));
        const bool i_out_null = o_null && ie_null;
        const bool d_out_null = o_null && de_null;
        if (sub_null && i_out_null && d_out_null) {
            cur.flags = 0;
            if (lane == 0) meta[3 * s + 2] = 0;
            continue;
        }
        const int a_lo = sub_null ? 1 : A.lo, a_hi = sub_null ? -1 : A.hi;
        const int b_lo = o_null ? 1 : B.lo, b_hi = o_null ? -1 : B.hi;
        const bool e_none = ie_null && de_null;
        const int e_lo = e_none ? 1 : E.lo, e_hi = e_none ? -1 : E.hi;
        cur.lo = cur.lob = min(min(a_lo, b_lo), e_lo) - 1;
        cur.hi = max(max(a_hi, b_hi), e_hi) + 1;
        cur.wlen = cur.hi - cur.lo + 1;
        cur.flags = F_PRESENT | (i_out_null ? 0u : F_HAS_I) | (d_out_null ? 0u : F_HAS_D);
        const uint32_t need = (uint32_t)cur.wlen * (1u + !i_out_null + !d_out_null);
        if (K.ring) {
            cur.base = (uint32_t)(s % (int)K.ring) * K.ring_stride;
            if (need > K.ring_stride) { status = AIM_STATUS_ARENA; break; }
        } else {
            cur.base = top;
            top += need;
            if (top > K.hist_cap) { status = AIM_STATUS_ARENA; break; }
        }
        if (lane == 0) store_rec(meta, s, cur);

        // ---- compute_offsets (wfa.c:238-273), one diagonal per lane ----
        {
            const int16_t *AM = hist + A.base - A.lob;
            const int16_t *BM = hist + B.base - B.lob;
            const int16_t *EI = hist + E.base + E.wlen - E.lob;
            const int16_t *ED = hist + E.base + ((E.flags & F_HAS_I) ? 2 * E.wlen : E.wlen) - E.lob;
            int16_t *NM = hist + cur.base - cur.lob;
            int16_t *NI = NM + cur.wlen;
            int16_t *ND = NM + (i_out_null ? cur.wlen : 2 * cur.wlen);
            for (int k = cur.lo + lane; k <= cur.hi; k += 32) {
                int ins = -10;
                if (!i_out_null) {
                    int g = (b_lo <= k - 1 && k - 1 <= b_hi) ? (int)BM[k - 1] : kNull;
                    int ii = (!ie_null && e_lo <= k - 1 && k - 1 <= e_hi) ? (int)EI[k - 1] : kNull;
                    ins = (g == kNull && ii == kNull) ? kNull : (int)(int16_t)(max(g, ii) + 1);
                    NI[k] = (int16_t)ins;
                }
                int del = -10;
                if (!d_out_null) {
                    int g = (b_lo <= k + 1 && k + 1 <= b_hi) ? (int)BM[k + 1] : kNull;
                    int dd = (!de_null && e_lo <= k + 1 && k + 1 <= e_hi) ? (int)ED[k + 1] : kNull;
                    del = max(g, dd);
                    ND[k] = (int16_t)del;
                }
                int sub = -10;
                if (!sub_null) sub = (a_lo <= k && k <= a_hi) ? (int)(int16_t)(AM[k] + 1) : kNull;
                NM[k] = (int16_t)max(del, max(sub, ins));
            }
        }
        __syncwarp();
    }

    // ---- result + backtrace (wfa_backtracing.c:219-375), lane 0 ----
    int begin_offset = max_ops - 1;
    if (reached && K.backtrace && lane == 0) {
        char *ops = HG ? (K.ops + (size_t)i * 2 * RS) : sOps;
        const int ops_cap = 2 * RS;
        int b = begin_offset;
        int score = s, k = ak;
        int offset = hist[cur.base + (k - cur.lob)];
        int v = offset - k, h = offset;
        bool valid = (v > 0 && v <= pl && h > 0 && h <= tl);
        int type = 0;  // 0 M, 1 I, 2 D
        bool bad = false;
#define AIM_PUT(ch) do { if (b < 0 || b >= ops_cap) { bad = true; } else { ops[b] = (ch); } --b; } while (0)
        while (v > 0 && h > 0 && score > 0 && !bad) {
            if (!valid) {
                valid = (v > 0 && v <= pl && h > 0 && h <= tl);
                if (valid) {  // add_trailing_gap (wfa_backtracing.c:48-69)
                    if (k < ak) { for (int j = k; j < ak; ++j) AIM_PUT('I'); }
                    else if (k > ak) { for (int j = ak; j < k; ++j) AIM_PUT('D'); }
                }
            }
            const int s_open = score - K.o - K.e, s_ext = score - K.e, s_mis = score - K.x;
            const Rec GO = load_rec_or_absent(meta, s_open);
            const Rec GE = load_rec_or_absent(meta, s_ext);
            const Rec MM = load_rec_or_absent(meta, s_mis);
            int del_ext = kNull, del_open = kNull, ins_ext = kNull, ins_open = kNull, misms = kNull;
            if (type != 1) {
                if ((GE.flags & F_PRESENT) && (GE.flags & F_HAS_D) && GE.lo <= k + 1 && k + 1 <= GE.hi)
                    del_ext = hist[GE.base + ((GE.flags & F_HAS_I) ? 2 * GE.wlen : GE.wlen) + (k + 1 - GE.lob)];
                if ((GO.flags & F_PRESENT) && GO.lo <= k + 1 && k + 1 <= GO.hi)
                    del_open = hist[GO.base + (k + 1 - GO.lob)];
            }
            if (type != 2) {
                if ((GE.flags & F_PRESENT) && (GE.flags & F_HAS_I) && GE.lo <= k - 1 && k - 1 <= GE.hi)
                    ins_ext = (int16_t)(hist[GE.base + GE.wlen + (k - 1 - GE.lob)] + 1);
                if ((GO.flags & F_PRESENT) && GO.lo <= k - 1 && k - 1 <= GO.hi)
                    ins_open = (int16_t)(hist[GO.base + (k - 1 - GO.lob)] + 1);
            }
            if (type == 0) {
                if ((MM.flags & F_PRESENT) && MM.lo <= k && k <= MM.hi)
                    misms = (int16_t)(hist[MM.base + (k - MM.lob)] + 1);
            }
            const int max_all = max(misms, max(max(ins_ext, ins_open), max(del_ext, del_open)));
            if (type == 0) {
                int num_matches = offset - max_all;  // ops are 'M' already: a match run is a pointer move
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
    if (lane == 0) {
        aim_result r;
        r.max_operations = max_ops;
        r.begin_offset = begin_offset;
        r.end_offset = max_ops;
        r.score = s;  // == MAX_SCORE + 1 on give-up
        r.status = status;
        r.idx = K.idx_base + i;
        K.results[i] = r;
    }
    __syncwarp();
    if (!HG && K.backtrace) {
        const uint4 *src = reinterpret_cast<const uint4 *>(sOps);
        uint4 *dst = reinterpret_cast<uint4 *>(K.ops + (size_t)i * 2 * RS);
        for (int c = lane; c < (2 * RS) / 16; c += 32) dst[c] = src[c];
    }
    __syncwarp();
}

template <bool HG>
__global__ void __launch_bounds__(256) wfa_kernel(const WfaK K)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t wpb = blockDim.x >> 5;
    const uint32_t gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
    unsigned char *ws = smem + (size_t)wib * K.warp_smem_bytes;
    uint32_t *sP = reinterpret_cast<uint32_t *>(ws);
    uint32_t *sT = sP + K.seq_words;
    uint32_t *meta;
    int16_t *hist;
    char *sOps = nullptr;
    if (HG) {
        meta = K.g_meta + (size_t)gw * K.g_meta_stride;
        hist = K.g_hist + (size_t)gw * K.g_hist_stride;
    } else {
        meta = sT + K.seq_words;
        hist = reinterpret_cast<int16_t *>(meta + K.meta_words);
        sOps = reinterpret_cast<char *>(hist + K.hist_cap);
    }
    if (K.list) {
        const uint32_t cnt = *K.list_count;
        for (uint32_t j = gw; j < cnt; j += nw) wfa_pair<HG>(K, lane, K.list[j], sP, sT, meta, hist, sOps);
        return;
    }
    for (uint32_t i = gw; i < K.n; i += nw) wfa_pair<HG>(K, lane, i, sP, sT, meta, hist, sOps);
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace

// Simulate the lo/hi/I/D schedule of wfa.c:275-354 without data (no trimming).
WfaSchedule wfa_schedule(int max_score, int x, int o, int e)
{
    struct S { bool present, has_i, has_d; int lo, hi; };
    std::vector<S> w((size_t)max_score + 1);
    WfaSchedule out{0, 1, (uint32_t)std::max(x, o + e) + 1};
    w[0] = {true, false, false, 0, 0};
    uint64_t slots = 1;
    for (int s = 1; s <= max_score; ++s) {
        const S *A = (s - x >= 0 && w[s - x].present) ? &w[s - x] : nullptr;
        const S *B = (s - o - e >= 0 && w[s - o - e].present) ? &w[s - o - e] : nullptr;
        const S *E = (s - e >= 0 && w[s - e].present) ? &w[s - e] : nullptr;
        bool ie_null = !(E && E->has_i), de_null = !(E && E->has_d);
        bool i_out_null = !B && ie_null, d_out_null = !B && de_null;
        if (!A && i_out_null && d_out_null) { w[s] = {false, false, false, 1, -1}; continue; }
        int lo = std::min(std::min(A ? A->lo : 1, B ? B->lo : 1), (ie_null && de_null) ? 1 : E->lo) - 1;
        int hi = std::max(std::max(A ? A->hi : -1, B ? B->hi : -1), (ie_null && de_null) ? -1 : E->hi) + 1;
        w[s] = {true, !i_out_null, !d_out_null, lo, hi};
        uint32_t width = (uint32_t)(hi - lo + 1);
        out.max_width = std::max(out.max_width, width);
        slots += (uint64_t)width * (1u + !i_out_null + !d_out_null);
    }
    out.hist_slots = (uint32_t)std::min<uint64_t>(slots, 0xffffffffu);
    return out;
}

// Launch configuration of the warp-per-pair kernel for one batch (sizes only, nothing enqueued).
WarpPlan wfa_warp_plan(const KernelArgs &a, int sm_count, uint32_t max_pairs)
{
    const aim_params &p = a.p;
    WarpPlan W{};
    WfaK K{};
    K.plen = a.plen; K.tlen = a.tlen; K.patterns = a.patterns; K.texts = a.texts;
    K.results = a.results; K.ops = a.ops; K.n = a.n; K.idx_base = a.idx_base;
    K.x = p.mismatch; K.o = p.gap_open; K.e = p.gap_ext;
    K.max_score = p.max_score; K.read_size = p.read_size; K.backtrace = p.backtrace; K.reduce = p.reduce;
    K.seq_words = round_up((uint32_t)p.read_size / 16 + 2, 4);
    K.meta_words = round_up(3u * ((uint32_t)p.max_score + 1), 4);
    const WfaSchedule sch = wfa_schedule(p.max_score, p.mismatch, p.gap_open, p.gap_ext);
    uint32_t width_cap = sch.max_width;
    if (p.backtrace) {
        K.ring = 0;
        K.hist_cap = round_up(sch.hist_slots, 8);
    } else {
        K.ring = sch.ring_scores;
        K.ring_stride = round_up(3 * width_cap, 8);
        K.hist_cap = K.ring * K.ring_stride;
    }
    const uint32_t ops_bytes = p.backtrace ? 2u * (uint32_t)p.read_size : 0u;
    const uint32_t seq_bytes = 2 * K.seq_words * 4;
    const uint64_t short_bytes = (uint64_t)seq_bytes + K.meta_words * 4ull + (uint64_t)K.hist_cap * 2 + ops_bytes;
    const uint32_t kSmemBudget = 227u * 1024u;
    const bool hg = short_bytes > kSmemBudget / 8;  // fewer than 8 warps/SM would fit: long-read mode

    int warps_per_block, blocks_per_sm;
    const uint32_t kSmemPerSm = 228u * 1024u, kBlockReserve = 1024u;
    if (!hg) {
        K.warp_smem_bytes = round_up((uint32_t)short_bytes, 16);
        warps_per_block = 8;
    } else {
        // long-read mode: only the packed sequences stay in shared memory; history + meta go to a
        // per-warp HBM arena that is reused pair after pair (stays L2-resident when small).
        // One pair per (one-warp) block.
        K.warp_smem_bytes = round_up(seq_bytes, 16);
        if (p.backtrace) {
            // unbounded in theory (adaptive trimming is a heuristic); the arena is a tunable.
            uint64_t arena = (uint64_t)(p.arena_mb > 0 ? p.arena_mb : 8) << 20;
            uint64_t want = (uint64_t)sch.hist_slots * 2;
            K.hist_cap = (uint32_t)(std::min<uint64_t>(std::min(arena, want), 0x0ffffff0ull * 2) / 2);
        }
        warps_per_block = 1;
    }
    W.smem_block = (size_t)warps_per_block * K.warp_smem_bytes;
    if (W.smem_block > kSmemBudget) { W.rc = AIM_ERR_ARG; return W; }
    blocks_per_sm = (int)std::min<uint32_t>(kSmemPerSm / ((uint32_t)W.smem_block + kBlockReserve), 32u);
    blocks_per_sm = std::max(1, std::min(blocks_per_sm, (hg ? 32 : 64) / warps_per_block));
    int grid = sm_count * blocks_per_sm;
    uint32_t total_warps = (uint32_t)grid * (uint32_t)warps_per_block;
    if (total_warps > max_pairs) {
        grid = (int)((max_pairs + warps_per_block - 1) / warps_per_block);
        total_warps = (uint32_t)grid * (uint32_t)warps_per_block;
    }
    W.grid = grid;
    W.block = warps_per_block * 32;
    W.hg = hg;
    if (hg) {
        K.g_meta_stride = K.meta_words;
        K.g_hist_stride = round_up(K.hist_cap, 8);
        W.meta_bytes = ((size_t)total_warps * K.g_meta_stride * 4 + 255) / 256 * 256;
        W.scratch_bytes = W.meta_bytes + (size_t)total_warps * K.g_hist_stride * 2;
    }
    static_assert(sizeof(WfaK) <= sizeof(W.kernel_args), "WarpPlan::kernel_args too small");
    memcpy(W.kernel_args, &K, sizeof(K));
    W.rc = AIM_OK;
    return W;
}

// Enqueue the warp-per-pair kernel; `scratch` holds W.scratch_bytes.  With list != nullptr only the
// pairs list[0 .. *list_count) are served.
int wfa_warp_launch(const WarpPlan &W, void *scratch, const uint32_t *list, const uint32_t *list_count, void *stream_v, int *launches)
{
    cudaStream_t stream = (cudaStream_t)stream_v;
    WfaK K;
    memcpy(&K, W.kernel_args, sizeof(K));
    if (W.hg) {
        K.g_meta = reinterpret_cast<uint32_t *>(scratch);
        K.g_hist = reinterpret_cast<int16_t *>(reinterpret_cast<unsigned char *>(scratch) + W.meta_bytes);
    }
    K.list = list;
    K.list_count = list_count;
    cudaError_t err;
    if (W.hg) {
        err = cudaFuncSetAttribute(wfa_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W.smem_block);
        if (err == cudaSuccess) wfa_kernel<true><<<W.grid, W.block, W.smem_block, stream>>>(K);
    } else {
        err = cudaFuncSetAttribute(wfa_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W.smem_block);
        if (err == cudaSuccess) wfa_kernel<false><<<W.grid, W.block, W.smem_block, stream>>>(K);
    }
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { set_error(std::string("wfa launch: ") + cudaGetErrorString(err)); return AIM_ERR_CUDA; }
    if (launches) ++*launches;
    return AIM_OK;
}

int launch_wfa(const KernelArgs &a, Scratch *sc, void *stream_v, int *launches)
{
    const aim_params &p = a.p;
    if (p.max_score < 0 || p.max_score > 32000) { set_error("MAX_SCORE out of range for int16 offsets"); return AIM_ERR_ARG; }
    if (a.n == 0) return AIM_OK;
    {   // short reads: several pairs per warp in lockstep (aim_wfa_sub.cu); falls through when not applicable
        const int rc = launch_wfa_sub(a, sc, stream_v, launches);
        if (rc != 1) return rc;
        if (a.packed) { set_error("the packed entry serves the short-read WFA kernel only (MAX_SCORE / READ_SIZE too large)"); return AIM_ERR_ARG; }
    }
    {   // long reads, score only: windowed rings in shared memory (aim_wfa_long.cu); its leftovers come back
        // through wfa_warp_launch with a list
        const int rc = launch_wfa_long(a, sc, stream_v, launches);
        if (rc != 1) return rc;
    }
    const WarpPlan W = wfa_warp_plan(a, sc->sm_count, a.n);
    if (W.rc != AIM_OK) { set_error("READ_SIZE too large for the shared-memory sequence stage"); return W.rc; }
    if (W.scratch_bytes) {
        const int rc = scratch_reserve(sc, W.scratch_bytes);
        if (rc != AIM_OK) return rc;
    }
    return wfa_warp_launch(W, sc->buf, nullptr, nullptr, stream_v, launches);
}

}  // namespace aim
