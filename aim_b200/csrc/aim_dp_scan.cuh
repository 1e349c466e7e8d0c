// Aliased NW / SWG pairs (pattern_len > text_len) with a ROW SPREAD OVER THE LANES of a sub-warp: the per-lane half of
// dp_scan_kernel (aim_dp_fast.cu).  Same flat-array semantics as dp_row_kernel (reference: NW/DPU-WRAM/dpu/nw.c:67-153,
// SWG/DPU-MRAM/dpu/swg.c:66-217 with num_cols = text_len + 1 <= pattern_len).
//
// Why a row can be split although the pair is one serial chain.  With the flat table aliased, row h+1 starts from the first
// TAIL cell of row h (flat word num_cols*(h+1) is both), which needs row h's whole head: rows cannot overlap.  But INSIDE a
// row the only left-to-right dependency is the horizontal gap,
//     D[v] = min(M[v-1] + o + e, D[v-1] + e),   M[v] = min(X[v], D[v]),   X[v] = min(I[v], diag + sub)   (swg.c:199-211)
// and with o >= 0 this is D[v] = min(X[v-1] + o + e, D[v-1] + e): a min-plus prefix scan over X, which depends on the
// previous row only (NW: the same with o + e = e = GAP_D).  Min and + on integers are exact, so every M, I, D - and every
// comparison the traceback makes - is the value the serial order produces.  G lanes hold the row: lane l the columns
// 1 + 2*C*l .. 2*C*(l+1) as two BLOCKS of C columns, block A in the low halves and block B in the high halves of C packed
// s16x2 registers (every value is a non-negative int16, launcher guard), so one instruction works on two columns.
// Per row:  (1) X of all columns and the block's local scan (carry-in "infinite");  (2) a G-lane min-plus scan of the
// block aggregates by shuffle gives every block its true carry-in D;  (3) del / M / the four traceback predicates per
// column with the true values;  (4) the 1..C tail cells (columns num_cols..pattern_len: they read the CURRENT row's
// columns 0..C as their "previous row") serially on the sub-warp's first lane, whose low block holds exactly those columns.
//
// This file is also compiled by g++ as a lane-by-lane model of the kernel (tests/model/dp_scan_model.cpp, AIM_SCAN_HOST_MODEL):
// the functions below are the same source in both, only the cross-lane exchange (shuffles there, loops here) differs.
#ifndef AIM_DP_SCAN_CUH
#define AIM_DP_SCAN_CUH

// build-time variants (measured on config 3)
#ifndef AIM_SCAN_MIN3
#define AIM_SCAN_MIN3 0       // 1: the block's local scan as one VIMNMX3 per column, 0: VIMNMX + VIADDMNMX (config 3: 43.7 against 43.6 ms)
#endif

#include <stdint.h>

#ifdef AIM_SCAN_HOST_MODEL
#define AIM_SD static inline
namespace scanx {
static inline uint32_t vmin(uint32_t a, uint32_t b)
{   // per-half signed minimum (VIMNMX.S16x2)
    const int16_t al = (int16_t)(a & 0xffffu), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffffu), bh = (int16_t)(b >> 16);
    return (uint32_t)(uint16_t)(al <= bl ? al : bl) | ((uint32_t)(uint16_t)(ah <= bh ? ah : bh) << 16);
}
static inline uint32_t vminu(uint32_t a, uint32_t b)
{
    const uint32_t al = a & 0xffffu, ah = a >> 16, bl = b & 0xffffu, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
static inline uint32_t vadd(uint32_t a, uint32_t b) { return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16); }
static inline uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) { return vmin(vmin(a, b), c); }
static inline uint32_t viaddmin(uint32_t a, uint32_t b, uint32_t c) { return vmin(vadd(a, b), c); }
static inline uint32_t vmin_relu(uint32_t a, uint32_t b)
{   // per-half max(min(a, b), 0), signed (VIMNMX.S16x2.RELU)
    const uint32_t m = vmin(a, b);
    const int16_t ml = (int16_t)(m & 0xffffu), mh = (int16_t)(m >> 16);
    return (uint32_t)(uint16_t)(ml < 0 ? 0 : ml) | ((uint32_t)(uint16_t)(mh < 0 ? 0 : mh) << 16);
}
static inline uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
}  // namespace scanx
#else
#define AIM_SD __device__ __forceinline__
namespace scanx {
AIM_SD uint32_t vmin(uint32_t a, uint32_t b) { return __vmins2(a, b); }
AIM_SD uint32_t vminu(uint32_t a, uint32_t b) { return __vminu2(a, b); }
AIM_SD uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
AIM_SD uint32_t viaddmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }
AIM_SD uint32_t vmin_relu(uint32_t a, uint32_t b) { return __vimin_s16x2_relu(a, b); }
AIM_SD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) { return __byte_perm(x, y, s); }
}  // namespace scanx
#endif

namespace scanx {
// The two predicates of a packed comparison pushed into acc, in the DATA path: acc = 2 * acc + [a > b] per half, i.e. the
// COMPLEMENT of the traceback's "a <= b", the first push ending highest (C pushes: column r at bit C-1-r, block B 16 higher).
// Both halves are non-negative int16, so (b + 0x8000) - a never borrows from the neighbour and, read as a signed half, is
// positive exactly where a > b; VIMNMX.S16x2.RELU clamps it to {0, 1}, one IMAD appends it: two alu-pipe instructions (the
// three-input add, the clamp) and one on the fma pipe.  The kernel is bound by the alu pipe (85 % busy at 66 % issue), and a
// third of its instructions push predicates.  Measured against this: shift + and-or merge (SHF, LOP3: three alu instructions per
// push, 43.7 ms at config 3); the shift as IMAD.HI (more instructions, 48.5 ms); the predicate outputs of VIMNMX.S16x2 with
// predicated adds (the columns of a block are independent, ptxas issues their minima back to back, runs out of predicate
// registers, parks them in general registers (P2R) and re-tests them at the end of the row: 1100 instead of 500 instructions).
AIM_SD void push_gt(uint32_t a, uint32_t b, uint32_t &acc)
{
    const uint32_t t = (b + 0x80008000u) - a;  // per half b - a + 0x8000, in [1, 0xffff]: 1..0x7fff (a positive int16) where a > b
    acc = acc * 2u + vmin_relu(t, 0x00010001u);
}
}  // namespace scanx

namespace scan {

AIM_SD uint32_t both(int v) { return (uint32_t)v * 0x00010001u; }
AIM_SD int lo16(uint32_t w) { return (int)(w & 0xffffu); }
AIM_SD int hi16(uint32_t w) { return (int)(w >> 16); }
AIM_SD uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
AIM_SD int imin(int a, int b) { return a < b ? a : b; }

// Penalties of a launch.  NW runs the same code with OE = E = the linear gap and no I / D state.
struct Pen {
    int O, E, OE, X, MS;       // E: SWG gap_ext, NW the gap (so that E * C is the carry of a block in both); MATCH == 0 (launcher)
    int INF;                   // "no carry-in": above every table value, INF + E * C still an int16
    uint32_t OE2, E2, INF2;
};

// The lane's share of one row: C packed registers, block A (columns base+1 .. base+C) in the low halves, block B
// (base+C+1 .. base+2C) in the high ones, base = 2*C*lane.
template <int C>
struct Lane {
    uint32_t uM[C];   // M of the previous row, after phase4 of this row
    uint32_t uI[C];   // I likewise (SWG); after phase12 already this row's I
    uint32_t pat[C];  // the columns' pattern bytes, each doubled into its 16-bit half
    uint32_t mm[C];   // diag + substitution of this row
    uint32_t dn[C];   // del (= D of this row) after phase4
};

// Row 0 (nw.c:119-124 / swg.c:167-175) and the pattern lanes.  wlo / whi: the 2 x C/4 pattern words of the lane's blocks.
template <int C, bool SWG>
AIM_SD void init_lane(Lane<C> &L, int lane, const Pen &P, const uint32_t *wlo, const uint32_t *whi)
{
#pragma unroll
    for (int r = 0; r < C; ++r) {
        const int va = 2 * C * lane + r + 1, vb = va + C;
        L.uM[r] = SWG ? pack16(P.O + va * P.E, P.O + vb * P.E) : pack16(va * P.OE, vb * P.OE);
        L.uI[r] = both(P.MS);
        const uint32_t ca = (wlo[r >> 2] >> (8 * (r & 3))) & 0xffu, cb = (whi[r >> 2] >> (8 * (r & 3))) & 0xffu;
        L.pat[r] = ca * 0x0101u | cb * 0x01010000u;
        L.mm[r] = 0;
        L.dn[r] = 0;
    }
}

// Phase 1 + 2: I and diag + sub of every column, X = min(I, diag + sub), and the blocks' local scans.
// dg0 = M of the previous row at the column left of each block; t4 = this row's text byte in all four bytes.
// Returns, per half, D at the first column AFTER the block if nothing came in from the left.
// Predicate bits (push_gt: complements): column r of block A -> bit C-1-r, of block B -> bit 16+C-1-r; compact() closes the gap when C < 16.
template <int C, bool SWG>
AIM_SD uint32_t phase12(Lane<C> &L, uint32_t dg0, uint32_t t4, const Pen &P, uint32_t &aI)
{
    uint32_t dg = dg0, dl = P.INF2;
#pragma unroll
    for (int r = 0; r < C; ++r) {
        const uint32_t um = L.uM[r];
        uint32_t ins;
        if (SWG) {
            const uint32_t i1 = um + P.OE2;
            const uint32_t i2 = L.uI[r] + P.E2;
            scanx::push_gt(i1, i2, aI);  // opI = (upM+o+e <= upI+e)  (swg.c:97)
            ins = scanx::vmin(i1, i2);
        } else {
            ins = um + P.OE2;  // GAP_I
        }
        L.uI[r] = ins;
        const uint32_t mm = scanx::vminu(L.pat[r] ^ t4, 0x00010001u) * (uint32_t)P.X + dg;
        L.mm[r] = mm;
        dg = um;
        // D[v+1] = min(D[v] + e, X[v] + o + e), X = min(ins, diag + sub)
#if AIM_SCAN_MIN3
        dl = scanx::vmin3(dl + P.E2, ins + P.OE2, mm + P.OE2);
#else
        dl = scanx::viaddmin(dl, P.E2, scanx::vmin(ins, mm) + P.OE2);
#endif
    }
    return dl;
}

// Phase 4: with the true D at the first column of each block (din), del / M and the predicates of every column.
// The opD bit of the blocks' FIRST columns is set by opd_first() once the neighbour's last M is known.
template <int C, bool SWG>
AIM_SD void phase4(Lane<C> &L, uint32_t din, const Pen &P, uint32_t &aP, uint32_t &aQ, uint32_t &aD)
{
    uint32_t mprev = 0, dprev = 0;
#pragma unroll
    for (int r = 0; r < C; ++r) {
        uint32_t del;
        if (r == 0) del = din;
        else if (SWG) {
            const uint32_t d1 = mprev + P.OE2;
            const uint32_t d2 = dprev + P.E2;
            scanx::push_gt(d1, d2, aD);  // opD = (leftM+o+e <= leftD+e)  (swg.c:88)
            del = scanx::vmin(d1, d2);
        } else {
            del = mprev + P.OE2;  // GAP_D
        }
        scanx::push_gt(del, L.uI[r], aP);  // p = (del <= ins)
        const uint32_t m1 = scanx::vmin(del, L.uI[r]);
        scanx::push_gt(m1, L.mm[r], aQ);   // q = (min(del, ins) <= diag + sub)
        const uint32_t m = scanx::vmin(m1, L.mm[r]);
        L.uM[r] = m;
        L.dn[r] = del;
        mprev = m;
        dprev = del;
    }
}

// opD of the blocks' first columns: del there is min(Mleft + o + e, Dleft + e), so "opened" <=> del == Mleft + o + e.
// mleft = this row's M at the column left of each block.
template <int C>
AIM_SD void opd_first(uint32_t mleft, uint32_t din, const Pen &P, uint32_t &aD)
{
    const uint32_t t = (din + 0x80008000u) - (mleft + P.OE2);  // (as push_gt)
    aD |= scanx::vmin_relu(t, 0x00010001u) << (C - 1);
}

// Bit of column position pos (= column - 1) in a compacted accumulator; the stored bit is the predicate's complement
template <int C>
AIM_SD int flag_bit(int pos)
{
    const int b = pos % (2 * C);
    return (b / C) * C + (C - 1 - b % C);
}

// The 2*C predicate bits of an accumulator as bits 0 .. 2C-1 (block A, then block B)
template <int C>
AIM_SD uint32_t compact(uint32_t acc)
{
    if (C == 16) return acc;
    if (C == 8) return scanx::byte_perm(acc, 0u, 0x4420u);  // bytes 0 and 2
    return (acc & ((1u << C) - 1u)) | (((acc >> 16) & ((1u << C) - 1u)) << C);
}

// a[r] for a run-time r (the registers cannot be indexed): a binary tree of selects
template <int C>
AIM_SD uint32_t pick(const uint32_t (&a)[C], int r)
{
    uint32_t t[C];
#pragma unroll
    for (int i = 0; i < C; ++i) t[i] = a[i];
#pragma unroll
    for (int s = 1; s < C; s <<= 1) {
#pragma unroll
        for (int i = 0; i + s < C; i += 2 * s) t[i] = (r & s) ? t[i + s] : t[i];
    }
    return t[0];
}

// The border of a row and what the tail leaves for the next one (first lane of the sub-warp).
struct Edge {
    int bM, bI, bD;  // column 0 of this row: M, I, and the D the first cell extends  (aliased: cell (h-1, num_cols))
    int c0prev;      // M of column 0 of the previous row (diag of column 1)
    int dgt;         // M(h-1, text_len): diag of the first tail cell
};

// The tail cells of a row: columns num_cols + j, j < d (d = pattern_len - text_len, 1 <= d <= C), in the reference's order.
// L = the FIRST lane's registers after phase4 (columns 1..C in the low halves); (lm, ld) = M and del of column text_len.
// tp = pattern bytes text_len + j; tc = the row's text byte; dlim = how many cells to walk (uniform over the warp): the largest d
// of the warp in a row that is some pair's LAST, else 1 - the fill reads nothing but the first tail cell of a row (it is column 0
// of the next row; the flat words of the others are rewritten by the next row's columns 1..d-1 before anything reads them), so the
// others only matter where they stay: in row text_len, for the score and the traceback.
// Returns the flag nibbles (bit 4j: p, 4j+1: q, 4j+2: opD, 4j+3: opI); next = border of the next row; lm = last cell's M.
template <int C, bool SWG>
AIM_SD uint64_t tail_cells(const Lane<C> &L, const Edge &ed, int &lm, int ld, const uint32_t *tp, uint32_t tc, int d, int dlim, const Pen &P,
                           int &tM, int &tI, int &tD)
{
    uint64_t tw = 0;
#pragma unroll
    for (int j = 0; j < C; ++j) {
        if (j >= dlim) break;
        const int upM = j == 0 ? ed.bM : lo16(L.uM[j == 0 ? 0 : j - 1]);
        const int upI = j == 0 ? ed.bI : lo16(L.uI[j == 0 ? 0 : j - 1]);
        const int dg = j == 0 ? ed.dgt : (j == 1 ? ed.bM : lo16(L.uM[j < 2 ? 0 : j - 2]));
        const uint32_t pb = (tp[j >> 2] >> (8 * (j & 3))) & 0xffu;
        const int mm = dg + (pb != tc ? P.X : 0);
        int ins, del;
        bool opI = false, opD = false;
        if (SWG) {
            const int i1 = upM + P.OE, i2 = upI + P.E;
            opI = i1 <= i2;
            ins = imin(i1, i2);
            const int d1 = lm + P.OE, d2 = ld + P.E;
            opD = d1 <= d2;
            del = imin(d1, d2);
        } else {
            ins = upM + P.OE;
            del = lm + P.OE;
        }
        const bool p = del <= ins;
        const int m1 = imin(del, ins);
        const bool q = m1 <= mm;
        const int m = imin(m1, mm);
        if (j < d) {
            tw |= (uint64_t)((p ? 1u : 0u) | (q ? 2u : 0u) | (opD ? 4u : 0u) | (opI ? 8u : 0u)) << (4 * j);
            lm = m;
            ld = del;
            if (j == 0) { tM = m; tI = ins; tD = del; }
        }
    }
    return tw;
}

// Where the traceback finds the predicates of flat word num_cols*h + v: its LAST WRITER (row r, column c)
// (nw.c:78-94 / swg.c:106-133 read the final table; see the header of aim_dp.cu).
AIM_SD void last_writer(int nc, int tl, int h, int v, int &r, int &c)
{
    const int fi = nc * h + v;
    r = imin(tl, (fi - 1) / nc);
    c = fi - nc * r;
}

}  // namespace scan

#endif  // AIM_DP_SCAN_CUH
