// TEST INFRASTRUCTURE: lane-by-lane CPU model of dp_scan_kernel (aim_b200/csrc/aim_dp_fast.cu).
// The per-lane code is the product header aim_dp_scan.cuh itself (AIM_SCAN_HOST_MODEL); only the cross-lane exchange is
// restated here - every __shfl_*_sync(…, width G) of the kernel is a loop over the G lanes of one sub-warp, in the same order
// and with the same arithmetic - and the predicate records keep the kernel's layout, so that tests/test_dp_scan_model.py
// checks the ALGORITHM (min-plus scan of the horizontal gap, aliased tail, last-writer traceback) against the oracle on the
// CPU, where there is no GPU.  Never linked into the product.
#define AIM_SCAN_HOST_MODEL 1
#include "aim_dp_scan.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

template <int C, bool SWG>
struct Flags {  // ScanFlags of the kernel
    static constexpr int FW = SWG ? (C == 16 ? 4 : 2) : (C == 16 ? 2 : 1);
    static void store(uint32_t *d, uint32_t aP, uint32_t aQ, uint32_t aD, uint32_t aI)
    {
        if (SWG && C == 16) { d[0] = aP; d[1] = aQ; d[2] = aD; d[3] = aI; }
        else if (SWG) { d[0] = aP | (aQ << 16); d[1] = aD | (aI << 16); }
        else if (C == 16) { d[0] = aP; d[1] = aQ; }
        else d[0] = aP | (aQ << 16);
    }
    static void load(const uint32_t *d, int bit, bool &p, bool &q, bool &opD, bool &opI)
    {   // complements
        opD = opI = false;
        if (SWG && C == 16) { p = !((d[0] >> bit) & 1u); q = !((d[1] >> bit) & 1u); opD = !((d[2] >> bit) & 1u); opI = !((d[3] >> bit) & 1u); }
        else if (SWG) { p = !((d[0] >> bit) & 1u); q = !((d[0] >> (16 + bit)) & 1u); opD = !((d[1] >> bit) & 1u); opI = !((d[1] >> (16 + bit)) & 1u); }
        else if (C == 16) { p = !((d[0] >> bit) & 1u); q = !((d[1] >> bit) & 1u); }
        else { p = !((d[0] >> bit) & 1u); q = !((d[0] >> (16 + bit)) & 1u); }
    }
};

struct Res { int32_t max_operations, begin_offset, end_offset, score, status; };

template <int C, int G, bool SWG>
void model_pair(int X, int O, int E_, int MS, int RS, int backtrace, int pl, int tl, const uint8_t *gp, const uint8_t *gt, Res *res, uint8_t *ops,
                int tlmax_extra)
{
    constexpr int FW = Flags<C, SWG>::FW;
    scan::Pen P;
    P.O = O; P.X = X; P.MS = MS;
    P.OE = SWG ? O + E_ : O;
    P.E = SWG ? E_ : O;
    P.INF = 32767 - P.E * C - P.OE - 8;
    P.OE2 = scan::both(P.OE); P.E2 = scan::both(P.E); P.INF2 = scan::both(P.INF);
    const int EC = P.E * C;
    const int nc = tl + 1, d = pl - tl;
    const int tlmax = std::min(RS, tl + tlmax_extra), dmax = std::min(C, d + (tlmax_extra ? 1 : 0));  // as if a longer pair shared the warp
    auto ldw = [&](const uint8_t *p, int off) -> uint32_t { uint32_t w; memcpy(&w, p + off, 4); return w; };

    std::vector<scan::Lane<C>> L(G);
    uint32_t tp[C / 4];
    for (int sl = 0; sl < G; ++sl) {
        uint32_t wlo[C / 4], whi[C / 4];
        for (int w = 0; w < C / 4; ++w) {
            const int oa = 2 * C * sl + 4 * w, ob = oa + C;
            wlo[w] = oa < RS ? ldw(gp, oa) : 0u;
            whi[w] = ob < RS ? ldw(gp, ob) : 0u;
        }
        scan::init_lane<C, SWG>(L[sl], sl, P, wlo, whi);
    }
    for (int w = 0; w < C / 4; ++w) {
        uint32_t t = 0;
        for (int b = 0; b < 4; ++b) {
            const int o = tl + 4 * w + b;
            t |= (o < RS ? (uint32_t)gp[o] : 0u) << (8 * b);
        }
        tp[w] = t;
    }
    const int pt = tl - 1, ot = pt / (2 * C), ht = (pt / C) & 1, rt = pt % C;
    scan::Edge ed;
    ed.bM = ed.bI = ed.bD = 0;
    ed.c0prev = 0;
    ed.dgt = SWG ? P.O + tl * P.E : tl * P.OE;
    int tM = 0, tI = 0, tD = 0, score = 0;
    std::vector<uint32_t> fl((size_t)RS * G * FW, 0u);
    std::vector<uint64_t> tf((size_t)RS, 0u);

    for (int h = 1; h <= tlmax; ++h) {
        const uint32_t tc = h - 1 < RS ? gt[h - 1] : 0u, t4 = tc * 0x01010101u;
        if (h >= 2) { ed.bM = tM; ed.bI = tI; ed.bD = tD; }
        else if (SWG) { ed.bD = P.MS; ed.bI = P.O + P.E; ed.bM = ed.bI; }
        else { ed.bM = P.OE; ed.bI = 0; ed.bD = 0; }

        uint32_t dg0[G], dl[G], aP[G], aQ[G], aD[G], aI[G];
        for (int sl = 0; sl < G; ++sl) {  // nb = shfl_up(uM[C-1], 1)
            const uint32_t nb = sl >= 1 ? L[sl - 1].uM[C - 1] : L[sl].uM[C - 1];
            dg0[sl] = scan::pack16(sl == 0 ? ed.c0prev : scan::hi16(nb), scan::lo16(L[sl].uM[C - 1]));
        }
        for (int sl = 0; sl < G; ++sl) {
            aP[sl] = aQ[sl] = aD[sl] = aI[sl] = 0;
            dl[sl] = scan::phase12<C, SWG>(L[sl], dg0[sl], t4, P, aI[sl]);
        }
        int a_lo[G], val[G];
        for (int sl = 0; sl < G; ++sl) {
            a_lo[sl] = scan::lo16(dl[sl]);
            val[sl] = std::min(scan::hi16(dl[sl]), a_lo[sl] + EC);
        }
        for (int dlt = 1; dlt < G; dlt <<= 1) {
            int t[G];
            for (int sl = 0; sl < G; ++sl) t[sl] = sl >= dlt ? val[sl - dlt] : val[sl];
            for (int sl = 0; sl < G; ++sl) if (sl >= dlt) val[sl] = std::min(val[sl], t[sl] + 2 * EC * dlt);
        }
        const int din0 = SWG ? std::min(ed.bM + P.OE, ed.bD + P.E) : ed.bM + P.OE;
        int S[G];
        uint32_t din[G];
        for (int sl = 0; sl < G; ++sl) S[sl] = std::min(val[sl], din0 + 2 * EC * (sl + 1));
        for (int sl = 0; sl < G; ++sl) {
            const int in_lo = sl == 0 ? din0 : S[sl - 1];
            const int in_hi = std::min(a_lo[sl], in_lo + EC);
            din[sl] = scan::pack16(in_lo, in_hi);
        }
        for (int sl = 0; sl < G; ++sl) scan::phase4<C, SWG>(L[sl], din[sl], P, aP[sl], aQ[sl], aD[sl]);
        if (SWG) {
            for (int sl = 0; sl < G; ++sl) {
                const uint32_t nbn = sl >= 1 ? L[sl - 1].uM[C - 1] : L[sl].uM[C - 1];
                const uint32_t mleft = scan::pack16(sl == 0 ? ed.bM : scan::hi16(nbn), scan::lo16(L[sl].uM[C - 1]));
                scan::opd_first<C>(mleft, din[sl], P, aD[sl]);
            }
        }
        for (int sl = 0; sl < G; ++sl)
            Flags<C, SWG>::store(&fl[((size_t)(h - 1) * G + sl) * FW], scan::compact<C>(aP[sl]), scan::compact<C>(aQ[sl]), scan::compact<C>(aD[sl]),
                                 scan::compact<C>(aI[sl]));
        const uint32_t sm = scan::pick<C>(L[ot].uM, rt), sd = scan::pick<C>(L[ot].dn, rt);
        int lm = ht ? scan::hi16(sm) : scan::lo16(sm);
        const int ld = ht ? scan::hi16(sd) : scan::lo16(sd);
        const int mtl = lm;
        const int dlim = (h == tl || (tlmax_extra && h == tlmax)) ? dmax : 1;  // __any_sync(h == tl) over the warp's pairs
        tf[h - 1] = scan::tail_cells<C, SWG>(L[0], ed, lm, ld, tp, tc, d, dlim, P, tM, tI, tD);
        if (h == tl) score = lm;
        ed.c0prev = ed.bM;
        ed.dgt = mtl;
    }

    int begin_offset = pl + tl - 1, status = 0;
    if (backtrace) {
        int b = pl + tl - 1, h = tl, v = pl, layer = 0;
        while (h > 0 && v > 0) {
            int r, c;
            scan::last_writer(nc, tl, h, v, r, c);
            bool p, q, opD, opI;
            if (c >= nc) {
                const uint32_t nib = (uint32_t)(tf[r - 1] >> (4 * (c - nc))) & 15u;
                p = nib & 1u; q = nib & 2u; opD = nib & 4u; opI = nib & 8u;
            } else {
                const int pos = c - 1;
                Flags<C, SWG>::load(&fl[((size_t)(r - 1) * G + pos / (2 * C)) * FW], scan::flag_bit<C>(pos), p, q, opD, opI);
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
                if (b < 0) { status = 1; break; }
                if (layer == 2) { ops[b--] = 'D'; if (opD) layer = 0; --v; }
                else if (layer == 1) { ops[b--] = 'I'; if (opI) layer = 0; --h; }
                else if (q) layer = p ? 2 : 1;
                else {
                    if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                    --b; --h; --v;
                }
            }
        }
        if (status == 0) {
            while (h > 0) { ops[b--] = 'I'; --h; }
            while (v > 0) { ops[b--] = 'D'; --v; }
            begin_offset = b + 1;
        }
    }
    res->max_operations = pl + tl;
    res->begin_offset = begin_offset;
    res->end_offset = pl + tl;
    res->score = score;
    res->status = status;
}

template <int C, int G>
int run(int algo, int X, int O, int E, int MS, int RS, int backtrace, uint32_t n, const int32_t *plen, const int32_t *tlen, const uint8_t *pats,
        const uint8_t *txts, Res *results, uint8_t *ops, uint8_t *served, int extra)
{
    for (uint32_t i = 0; i < n; ++i) {
        const int pl = std::min(std::max(plen[i], 0), RS), tl = std::min(std::max(tlen[i], 0), RS);
        served[i] = pl > tl && tl <= 2 * C * G && pl - tl <= std::min(C, tl);  // classify_kernel
        if (!served[i]) continue;
        uint8_t *o = ops + (size_t)i * 2 * RS;
        memset(o, 'M', (size_t)2 * RS);
        if (algo == 0) model_pair<C, G, false>(X, O, E, MS, RS, backtrace, pl, tl, pats + (size_t)i * RS, txts + (size_t)i * RS, results + i, o, extra);
        else model_pair<C, G, true>(X, O, E, MS, RS, backtrace, pl, tl, pats + (size_t)i * RS, txts + (size_t)i * RS, results + i, o, extra);
    }
    return 0;
}

}  // namespace

// algo: 0 NW, 1 SWG.  served[i] = 1 where the kernel's classification would hand pair i to dp_scan_kernel (results / ops written).
// extra > 0: the pair walks `extra` more rows and one more tail cell than it has, as when a longer pair shares its warp.
extern "C" int scan_model_align(int algo, int C, int G, int X, int O, int E, int MS, int RS, int backtrace, uint32_t n, const int32_t *plen,
                                const int32_t *tlen, const uint8_t *pats, const uint8_t *txts, int32_t *results, uint8_t *ops, uint8_t *served,
                                int extra)
{
    Res *r = reinterpret_cast<Res *>(results);
#define AIM_RUN(c, g) if (C == c && G == g) return run<c, g>(algo, X, O, E, MS, RS, backtrace, n, plen, tlen, pats, txts, r, ops, served, extra);
    AIM_RUN(4, 16) AIM_RUN(8, 8) AIM_RUN(8, 16) AIM_RUN(16, 8) AIM_RUN(16, 16)
#undef AIM_RUN
    return -1;
}
