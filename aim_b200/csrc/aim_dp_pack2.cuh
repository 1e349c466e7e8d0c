// Two pairs per thread in s16x2 lanes: the NW / SWG cell of aim_dp_fast.cu with every value of pair A in the low half and
// of pair B in the high half of one 32-bit register (included by aim_dp_fast.cu; same FastK, same flat-array semantics).
//
// sm_100a executes 16x2 integer SIMD in ONE instruction each: VIADD.16x2, VIMNMX.S16x2 (with one predicate output per
// half), VIADDMNMX.S16x2, VIMNMX3.S16x2 (checked with cuobjdump; __vcmp* / __vset* are 5-6 instruction emulations and are
// not used).  Because every DP value here is a non-negative number below 32768 (launcher guard: MATCH == 0 and
// (2*READ_SIZE+2)*max penalty + MAX_SCORE + o + e < 32767), a plain 32-bit add of two packed registers never carries
// from the low half into the high one, so the adds are ordinary IADD/IMAD (either pipe), the minima are VIMNMX.S16x2 /
// VIADDMNMX.S16x2 / VIMNMX3.S16x2, and the four traceback predicates of both pairs
//     opI = (upM+o+e <= upI+e)   opD = (leftM+o+e <= leftD+e)   p = (del <= ins)   q = (min(del,ins) <= diag+sub)
// (tie -> first operand, the same "<=" the sign-bit formulation of dp_cell() encodes) are bit 15 of (b + 0x8000 - a) in each
// half, one IADD3 for both pairs (see cell2).  A cell PAIR costs about 24 instructions where two dp_cell()s cost about 44.
//
// The substitution term compares the raw bytes of both pairs in 16-bit lanes (PRMT, XOR, unsigned 16x2 min with 1) and adds
// MISMATCH with one IMAD on the fma pipe: exact for any byte, as the reference compares bytes (nw.c:143, swg.c:206).
#ifndef AIM_DP_PACK2_CUH
#define AIM_DP_PACK2_CUH

namespace pack2 {

__device__ __forceinline__ uint32_t both(int v) { return (uint32_t)v * 0x00010001u; }
__device__ __forceinline__ int lo_half(uint32_t w) { return (int)(w & 0xffffu); }
__device__ __forceinline__ int hi_half(uint32_t w) { return (int)(w >> 16); }

// Substitution term of column j (0..15) of a 16-column record for both pairs: the raw bytes are compared, as the reference does
// (nw.c:143, swg.c:206).  pa / pb = the word holding column j's pattern byte of pair A / B, t2 = text byte of A in bytes 0-1 and
// of B in bytes 2-3.  PRMT puts A's byte into both bytes of the low half and B's into the high half, XOR leaves a zero half where
// the bases are equal, an unsigned 16x2 min with 1 turns "non-zero" into 1, and one IMAD adds MISMATCH where they differ.
template <int J>
__device__ __forceinline__ uint32_t sub2(uint32_t pa, uint32_t pb, uint32_t t2, uint32_t X, uint32_t dg)
{
    constexpr uint32_t k = J & 3;
    constexpr uint32_t sel = k | (k << 4) | ((4 + k) << 8) | ((4 + k) << 12);
    const uint32_t x = __byte_perm(pa, pb, sel) ^ t2;
    return __vminu2(x, 0x00010001u) * X + dg;
}
__device__ __forceinline__ uint32_t sub2_dyn(uint32_t ca, uint32_t cb, uint32_t t2, uint32_t X, uint32_t dg)
{   // ca / cb: the pattern BYTES of the two pairs
    const uint32_t x = (ca * 0x0101u | cb * 0x01010000u) ^ t2;
    return __vminu2(x, 0x00010001u) * X + dg;
}
__device__ __forceinline__ uint32_t text2(uint32_t ta, uint32_t tb) { return (ta & 0xffu) * 0x0101u | (tb & 0xffu) * 0x01010000u; }

// One cell of both pairs.  mm = diag + substitution (packed).  Returns the packed M; updates upI -> ins, leftD -> del.
// VIMNMX.S16x2 delivers the predicate of each half together with the minimum; each predicate ORs one bit into its accumulator
// (bitA for pair A, bitB for pair B: compile-time constants inside the unrolled records, where the compiler turns the ORs into
// selects merged by three-input adds: 1.5 instructions per bit).  Measured against taking the predicates in the data path
// (bit 15 of b + 0x8000 - a per half, shifted into the accumulators: no predicate registers, VIADDMNMX / VIMNMX3 for the M
// chain): 24.8 against 30.7 instructions per cell pair in the strip kernel, and faster in both kernels.  Also measured and rejected:
// the M chain through VIADDMNMX / VIMNMX3 (two dependent instructions per cell) with the predicate-delivering VIMNMX hanging off it:
// two more instructions per cell pair and 2 ms slower at config 3 in either kernel - the chain depth is not what limits them.
template <int ALGO>
__device__ __forceinline__ uint32_t cell2(uint32_t upM, uint32_t &upI, uint32_t leftM, uint32_t &leftD, uint32_t mm, uint32_t OE2, uint32_t E2,
                                          uint32_t bitA, uint32_t bitB, uint32_t &aP, uint32_t &aQ, uint32_t &aD, uint32_t &aI)
{
    uint32_t ins, del;
    bool h, l;
    if (ALGO == AIM_ALGO_NW) {
        ins = upM + OE2;    // GAP_I
        del = leftM + OE2;  // GAP_D
    } else {
        ins = __vibmin_s16x2(upM + OE2, upI + E2, &h, &l);
        if (l) aI |= bitA;
        if (h) aI |= bitB;
        del = __vibmin_s16x2(leftM + OE2, leftD + E2, &h, &l);
        if (l) aD |= bitA;
        if (h) aD |= bitB;
        upI = ins;
        leftD = del;
    }
    const uint32_t m1 = __vibmin_s16x2(del, ins, &h, &l);
    if (l) aP |= bitA;
    if (h) aP |= bitB;
    const uint32_t m = __vibmin_s16x2(m1, mm, &h, &l);
    if (l) aQ |= bitA;
    if (h) aQ |= bitB;
    return m;
}

// predicates of cell j of a packed record for pair `half` (0 = A, 1 = B): rec = {P, Q[, opD, opI]}
__device__ __forceinline__ void decode2(const uint32_t *rec, int j, int half, bool swg, bool &p, bool &q, bool &opD, bool &opI)
{
    const int b = j + 16 * half;
    p = (rec[0] >> b) & 1u;
    q = (rec[1] >> b) & 1u;
    opD = opI = false;
    if (swg) {
        opD = (rec[2] >> b) & 1u;
        opI = (rec[3] >> b) & 1u;
    }
}

}  // namespace pack2

// ================= non-aliased pairs, two per thread: register strips =================
// Thread t takes list entries 2t and 2t+1 (an odd last entry is paired with itself).  Both pairs walk max(tlen) rows and
// max(plen) columns in lockstep; the cells a pair does not have are computed on its zero-padded row bytes and never read.
template <int ALGO>
__global__ void __launch_bounds__(128) dp2_strip_kernel(const FastK K)
{
    constexpr bool SWG = (ALGO == AIM_ALGO_SWG);
    constexpr int FW = SWG ? 4 : 2;  // flag words per 16-cell record of a pair of pairs
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nth = gridDim.x * blockDim.x;
    const uint32_t count = *K.count;
    const uint32_t nduo = (count + 1) / 2;
    const int RS = K.read_size;
    const int X = K.x, E = K.e, MS = K.max_score, O = K.o;
    const int OE = SWG ? K.o + K.e : K.o;  // NW: the single linear gap
    const uint32_t OE2 = pack2::both(OE), E2 = pack2::both(E), MS2 = pack2::both(MS);
    uint2 *bnd = reinterpret_cast<uint2 *>(K.bound) + tid;
    uint32_t *flg = K.flags + (size_t)tid * FW;
    const size_t fstep = (size_t)nth * FW;  // words between consecutive (strip,row) records

    for (uint32_t d = tid; d < nduo; d += nth) {
        const uint32_t iA = K.list[2 * d], iB = (2 * d + 1 < count) ? K.list[2 * d + 1] : iA;
        const int plA = min(max(K.plen[iA], 0), RS), tlA = min(max(K.tlen[iA], 0), RS);
        const int plB = min(max(K.plen[iB], 0), RS), tlB = min(max(K.tlen[iB], 0), RS);
        const char *gpA = K.patterns + (size_t)iA * RS, *gtA = K.texts + (size_t)iA * RS;
        const char *gpB = K.patterns + (size_t)iB * RS, *gtB = K.texts + (size_t)iB * RS;
        const int tlm = max(tlA, tlB), plm = max(plA, plB);
        const int nstrips = (plm + KS - 1) / KS;
        const int lastA = (plA - 1) / KS, lastB = (plB - 1) / KS;  // strip holding column plen
        int scoreA = 0, scoreB = 0;

        for (int s = 0; s < nstrips; ++s) {
            const int v0 = s * KS;
            uint32_t pcA[KS / 4], pcB[KS / 4];
#pragma unroll
            for (int w = 0; w < KS / 4; ++w) {
                const bool in = v0 + 4 * w < RS;
                pcA[w] = in ? __ldg(reinterpret_cast<const uint32_t *>(gpA + v0) + w) : 0u;
                pcB[w] = in ? __ldg(reinterpret_cast<const uint32_t *>(gpB + v0) + w) : 0u;
            }
            uint32_t upM[KS], upI[KS];
#pragma unroll
            for (int j = 0; j < KS; ++j) {  // row 0 (nw.c:119-124 / swg.c:167-175)
                upM[j] = pack2::both(SWG ? O + (v0 + 1 + j) * E : (v0 + 1 + j) * OE);
                upI[j] = MS2;
            }
            // M(h-1, v0): the diagonal neighbour of the strip's first cell
            uint32_t dg0 = pack2::both(SWG ? (v0 == 0 ? 0 : O + v0 * E) : v0 * OE);
            const bool first = (s == 0), last = (s == nstrips - 1);
            uint32_t twA = 0, twB = 0;
            uint32_t *frec = flg + (size_t)s * RS * fstep;
            uint2 bnext = make_uint2(0u, 0u);  // the boundary of the NEXT row is fetched one row ahead
            if (!first && tlm >= 1) bnext = bnd[0];
            uint32_t twAn = __ldg(reinterpret_cast<const uint32_t *>(gtA)), twBn = __ldg(reinterpret_cast<const uint32_t *>(gtB));
            for (int h = 1; h <= tlm; ++h) {
                if (((h - 1) & 3) == 0) {  // four text bytes per load, the next four fetched now
                    twA = twAn;
                    twB = twBn;
                    if (h + 3 < RS) {
                        twAn = __ldg(reinterpret_cast<const uint32_t *>(gtA) + ((h + 3) >> 2));
                        twBn = __ldg(reinterpret_cast<const uint32_t *>(gtB) + ((h + 3) >> 2));
                    }
                }
                const uint32_t t2 = pack2::text2(twA, twB);
                twA >>= 8;
                twB >>= 8;
                uint32_t leftM, leftD;
                if (first) {  // column 0 (nw.c:114-118 / swg.c:158-166)
                    leftM = pack2::both(SWG ? O + h * E : h * OE);
                    leftD = MS2;
                } else {
                    leftM = bnext.x;
                    leftD = bnext.y;
                    if (h < tlm) bnext = bnd[(size_t)h * nth];
                }
                uint32_t dg = dg0;
                dg0 = leftM;
                uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
#define AIM_CELL2(j)                                                                                                              \
    {                                                                                                                             \
        const uint32_t um = upM[j];                                                                                               \
        const uint32_t mm = pack2::sub2<j>(pcA[(j) >> 2], pcB[(j) >> 2], t2, (uint32_t)X, dg);                                   \
        const uint32_t m = pack2::cell2<ALGO>(um, upI[j], leftM, leftD, mm, OE2, E2, 1u << (j), 1u << (16 + (j)), aP, aQ, aD, aI); \
        dg = um;                                                                                                                  \
        upM[j] = m;                                                                                                               \
        leftM = m;                                                                                                                \
    }
                AIM_CELL2(0) AIM_CELL2(1) AIM_CELL2(2) AIM_CELL2(3) AIM_CELL2(4) AIM_CELL2(5) AIM_CELL2(6) AIM_CELL2(7)
                AIM_CELL2(8) AIM_CELL2(9) AIM_CELL2(10) AIM_CELL2(11) AIM_CELL2(12) AIM_CELL2(13) AIM_CELL2(14) AIM_CELL2(15)
#undef AIM_CELL2
                if (!last) bnd[(size_t)(h - 1) * nth] = make_uint2(leftM, leftD);
                if (K.backtrace) {
                    uint32_t *dst = frec + (size_t)(h - 1) * fstep;
                    if (SWG) *reinterpret_cast<uint4 *>(dst) = make_uint4(aP, aQ, aD, aI);
                    else *reinterpret_cast<uint2 *>(dst) = make_uint2(aP, aQ);
                }
                // the score is the cell (tlen, plen) of each pair
                if (h == tlA && s == lastA) {
                    const int js = plA - 1 - v0;
#pragma unroll
                    for (int j = 0; j < KS; ++j) if (j == js) scoreA = pack2::lo_half(upM[j]);
                }
                if (h == tlB && s == lastB) {
                    const int js = plB - 1 - v0;
#pragma unroll
                    for (int j = 0; j < KS; ++j) if (j == js) scoreB = pack2::hi_half(upM[j]);
                }
            }
        }

#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const uint32_t i = half ? iB : iA;
            if (half && 2 * d + 1 >= count) break;  // the odd last entry was paired with itself
            const int pl = half ? plB : plA, tl = half ? tlB : tlA;
            const char *gp = half ? gpB : gpA, *gt = half ? gtB : gtA;
            int score = half ? scoreB : scoreA;
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
                    pack2::decode2(flg + ((size_t)s * RS + (h - 1)) * fstep, j, half, SWG, p, q, opD, opI);
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
}


// ================= aliased pairs (pattern_len > text_len), two per thread =================
// Such a pair is ONE serial chain in the reference's write order (header of aim_dp_fast.cu): row h+1 starts from cell
// (h, num_cols), the first TAIL cell of row h, which needs the whole head of row h.  What can be shared is the instruction
// stream: two pairs with the SAME text_len (the aliased pairs are bucketed by text_len first, an odd bucket's last pair runs
// alone) have the same head columns 1..text_len and the same aliasing geometry, so their cells run in the two halves of one
// register.  The row (M and, for SWG, I of columns 0..text_len, both pairs) lives in shared memory as [column][lane]; the tail
// cells (columns num_cols..pattern_len) read the CURRENT row's cells num_cols columns to their left as their previous row.
constexpr uint32_t RT2 = 32;
constexpr uint32_t kNoPartner = 0xffffffffu;

template <int ALGO>
__global__ void __launch_bounds__(RT2) dp2_row_kernel(const FastK K)
{
    constexpr bool SWG = (ALGO == AIM_ALGO_SWG);
    constexpr int FW = SWG ? 4 : 2;
    constexpr int RW = SWG ? 2 : 1;  // row words per column: M2 (, I2)
    extern __shared__ uint32_t smem2[];
    constexpr uint32_t T = RT2;
    const uint32_t tid = blockIdx.x * T + threadIdx.x;
    const uint32_t nth = gridDim.x * T;
    const uint32_t nduo = (*K.count + 1) / 2;  // slots / 2 (buckets start on even slots)
    const int RS = K.read_size;
    const int X = K.x, E = K.e, MS = K.max_score, O = K.o;
    const int OE = SWG ? K.o + K.e : K.o;
    const uint32_t OE2 = pack2::both(OE), E2 = pack2::both(E), MS2 = pack2::both(MS);
    uint32_t *row = smem2 + threadIdx.x;  // column c: row[(c * RW + k) * T], k = 0 M2, 1 I2
    uint32_t *flg = K.flags + (size_t)tid * FW;
    const size_t fstep = (size_t)nth * FW;
    const uint32_t rpr = K.wpr;  // records per row

    auto ldM = [&](int c) -> uint32_t { return row[(size_t)(c * RW) * T]; };
    auto ldI = [&](int c) -> uint32_t { return SWG ? row[(size_t)(c * RW + 1) * T] : 0u; };
    auto st = [&](int c, uint32_t m, uint32_t i) { row[(size_t)(c * RW) * T] = m; if (SWG) row[(size_t)(c * RW + 1) * T] = i; };

    for (uint32_t d = tid; d < nduo; d += nth) {
        const uint32_t iA = K.list[2 * d];
        if (iA == kNoPartner) continue;  // (an empty slot pair cannot occur; defensive)
        uint32_t iB = K.list[2 * d + 1];
        const bool single = iB == kNoPartner;
        if (single) iB = iA;
        const int tl = min(max(K.tlen[iA], 0), RS);  // == tlen[iB]: same bucket
        const int plA = min(max(K.plen[iA], 0), RS), plB = min(max(K.plen[iB], 0), RS);
        const char *gpA = K.patterns + (size_t)iA * RS, *gtA = K.texts + (size_t)iA * RS;
        const char *gpB = K.patterns + (size_t)iB * RS, *gtB = K.texts + (size_t)iB * RS;
        const int nc = tl + 1, plm = max(plA, plB);
        int scoreA = 0, scoreB = 0;

        // Pattern bytes do not change from row to row: record 0's and those of the first two records after the full head records
        // (remaining head cells + aliased tail) stay in registers for the whole pair; the other records' are fetched one record ahead.
        auto ld16 = [&](const char *gp, int off, uint2 &a, uint2 &b) {  // 16 bytes at gp + off (off a multiple of 16), zero past the row
            a = off + 8 <= RS ? __ldg(reinterpret_cast<const uint2 *>(gp + off)) : make_uint2(0u, 0u);
            b = off + 16 <= RS ? __ldg(reinterpret_cast<const uint2 *>(gp + off) + 1) : make_uint2(0u, 0u);
        };
        const int nfull = tl >> 4;       // full head records
        const bool fast = nc >= 32;      // tail cells read cells at least 32 columns back: never inside the record being computed
        uint2 p0[4], pg0[4], pg1[4];
        ld16(gpA, 0, p0[0], p0[1]); ld16(gpB, 0, p0[2], p0[3]);
        ld16(gpA, 16 * nfull, pg0[0], pg0[1]); ld16(gpB, 16 * nfull, pg0[2], pg0[3]);
        ld16(gpA, 16 * nfull + 16, pg1[0], pg1[1]); ld16(gpB, 16 * nfull + 16, pg1[2], pg1[3]);

        // row 0 (nw.c:119-124 / swg.c:167-175); only the head columns of row 0 are ever read
        st(0, 0u, MS2);
        for (int v = 1; v <= tl; ++v) st(v, pack2::both(SWG ? O + v * E : v * OE), MS2);
        uint32_t tailM = 0, tailI = 0, tailD = 0;  // cell (h-1, nc): column 0 of row h

        uint32_t twA = 0, twB = 0;
        for (int h = 1; h <= tl; ++h) {
            if (((h - 1) & 3) == 0) {
                twA = __ldg(reinterpret_cast<const uint32_t *>(gtA) + ((h - 1) >> 2));
                twB = __ldg(reinterpret_cast<const uint32_t *>(gtB) + ((h - 1) >> 2));
            }
            const uint32_t t2 = pack2::text2(twA, twB);
            twA >>= 8;
            twB >>= 8;
            uint32_t leftM, leftD, c0I;
            if (h >= 2) { leftM = tailM; c0I = tailI; leftD = tailD; }
            else if (!SWG) { leftM = pack2::both(h * OE); c0I = 0; leftD = 0; }
            else { leftD = MS2; c0I = pack2::both(O + h * E); leftM = c0I; }
            uint32_t dg = ldM(0);
            st(0, leftM, c0I);
            uint32_t *frow = flg + (size_t)(h - 1) * rpr * fstep;

            // ---- head, full 16-cell records (columns 1..tl); the next record's row words and pattern bytes are fetched before the
            // current record's chain runs (one warp per scheduler: nothing else hides a load) ----
            int v = 1;
            uint32_t curM[16], curI[16], nxtM[16], nxtI[16];
            uint2 pcur[4], pnxt[4];  // 16 pattern bytes of pair A ([0],[1]) and of pair B ([2],[3]); rows are READ_SIZE (a multiple of 8) apart
            if (tl >= 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { curM[j] = ldM(1 + j); curI[j] = ldI(1 + j); }
#pragma unroll
                for (int k = 0; k < 4; ++k) pcur[k] = p0[k];
            }
            for (; v + 15 <= tl; v += 16) {
                const bool more = v + 31 <= tl;
                if (more) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { nxtM[j] = ldM(v + 16 + j); nxtI[j] = ldI(v + 16 + j); }
                    pnxt[0] = __ldg(reinterpret_cast<const uint2 *>(gpA + v + 15)); pnxt[1] = __ldg(reinterpret_cast<const uint2 *>(gpA + v + 15) + 1);
                    pnxt[2] = __ldg(reinterpret_cast<const uint2 *>(gpB + v + 15)); pnxt[3] = __ldg(reinterpret_cast<const uint2 *>(gpB + v + 15) + 1);
                }
                const uint32_t pcA[4] = {pcur[0].x, pcur[0].y, pcur[1].x, pcur[1].y}, pcB[4] = {pcur[2].x, pcur[2].y, pcur[3].x, pcur[3].y};
                uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
#define AIM_RCELL2(j)                                                                                               \
    {                                                                                                               \
        const uint32_t um = curM[j];                                                                                \
        const uint32_t mm = pack2::sub2<j>(pcA[(j) >> 2], pcB[(j) >> 2], t2, (uint32_t)X, dg);                     \
        const uint32_t m = pack2::cell2<ALGO>(um, curI[j], leftM, leftD, mm, OE2, E2, 1u << (j), 1u << (16 + (j)), aP, aQ, aD, aI);\
        dg = um;                                                                                                    \
        curM[j] = m;                                                                                                \
        leftM = m;                                                                                                  \
    }
                AIM_RCELL2(0) AIM_RCELL2(1) AIM_RCELL2(2) AIM_RCELL2(3) AIM_RCELL2(4) AIM_RCELL2(5) AIM_RCELL2(6) AIM_RCELL2(7)
                AIM_RCELL2(8) AIM_RCELL2(9) AIM_RCELL2(10) AIM_RCELL2(11) AIM_RCELL2(12) AIM_RCELL2(13) AIM_RCELL2(14) AIM_RCELL2(15)
#undef AIM_RCELL2
#pragma unroll
                for (int j = 0; j < 16; ++j) st(v + j, curM[j], curI[j]);
                if (K.backtrace) {
                    uint32_t *dst = frow + (size_t)((v - 1) >> 4) * fstep;
                    if (SWG) *reinterpret_cast<uint4 *>(dst) = make_uint4(aP, aQ, aD, aI);
                    else *reinterpret_cast<uint2 *>(dst) = make_uint2(aP, aQ);
                }
                if (more) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { curM[j] = nxtM[j]; curI[j] = nxtI[j]; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) pcur[k] = pnxt[k];
                }
            }
            // ---- the remaining head cells, then the aliased tail (columns nc..pl: "previous row" = the CURRENT row's cells nc columns
            // to the left, flat word nc*(h-1)+v is cell (h, v-nc)).  In flat terms nothing changes at column nc: the left chain runs
            // on, up = word i-nc, diag = word i-nc-1.  So these columns are records like the others whose sixteen "up" words are
            // GATHERED from column v (head) or v-nc (tail); the diagonal follows the up sequence one cell behind, as everywhere. ----
            if (fast) {
                for (int g = 0; v <= plm; v += 16, ++g) {
                    uint32_t gM[16], gI[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int vj = v + j;
                        const int c = vj > plm ? 0 : (vj <= tl ? vj : vj - nc);
                        gM[j] = ldM(c);
                        gI[j] = ldI(c);
                    }
                    uint2 pq[4];
                    if (g == 0) { pq[0] = pg0[0]; pq[1] = pg0[1]; pq[2] = pg0[2]; pq[3] = pg0[3]; }
                    else if (g == 1) { pq[0] = pg1[0]; pq[1] = pg1[1]; pq[2] = pg1[2]; pq[3] = pg1[3]; }
                    else { ld16(gpA, v - 1, pq[0], pq[1]); ld16(gpB, v - 1, pq[2], pq[3]); }
                    const uint32_t pcA[4] = {pq[0].x, pq[0].y, pq[1].x, pq[1].y}, pcB[4] = {pq[2].x, pq[2].y, pq[3].x, pq[3].y};
                    const bool lastrow = h == tl;
                    uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
#define AIM_GCELL2(j)                                                                                                                   \
    {                                                                                                                                   \
        const uint32_t um = gM[j];                                                                                                      \
        const uint32_t mm = pack2::sub2<j>(pcA[(j) >> 2], pcB[(j) >> 2], t2, (uint32_t)X, dg);                                         \
        const uint32_t m = pack2::cell2<ALGO>(um, gI[j], leftM, leftD, mm, OE2, E2, 1u << (j), 1u << (16 + (j)), aP, aQ, aD, aI);       \
        dg = um;                                                                                                                        \
        gM[j] = m;                                                                                                                      \
        leftM = m;                                                                                                                      \
        if (v + (j) == nc) { tailM = m; tailI = gI[j]; tailD = leftD; }                                                                 \
        if (lastrow) {                                                                                                                  \
            if (v + (j) == plA) scoreA = pack2::lo_half(m);                                                                             \
            if (v + (j) == plB) scoreB = pack2::hi_half(m);                                                                             \
        }                                                                                                                               \
    }
                    AIM_GCELL2(0) AIM_GCELL2(1) AIM_GCELL2(2) AIM_GCELL2(3) AIM_GCELL2(4) AIM_GCELL2(5) AIM_GCELL2(6) AIM_GCELL2(7)
                    AIM_GCELL2(8) AIM_GCELL2(9) AIM_GCELL2(10) AIM_GCELL2(11) AIM_GCELL2(12) AIM_GCELL2(13) AIM_GCELL2(14) AIM_GCELL2(15)
#undef AIM_GCELL2
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (v + j <= plm) st(v + j, gM[j], gI[j]);
                    if (K.backtrace) {
                        uint32_t *dst = frow + (size_t)((v - 1) >> 4) * fstep;
                        if (SWG) *reinterpret_cast<uint4 *>(dst) = make_uint4(aP, aQ, aD, aI);
                        else *reinterpret_cast<uint2 *>(dst) = make_uint2(aP, aQ);
                    }
                }
            }
            // (short texts, num_cols < 32: one cell at a time)
            uint32_t aP = 0, aQ = 0, aD = 0, aI = 0;
            for (; v <= plm; ++v) {
                const bool tail = v >= nc;
                uint32_t upM, upI;
                if (!tail) { upM = ldM(v); upI = ldI(v); }
                else {
                    upM = ldM(v - nc);
                    upI = ldI(v - nc);
                    if (v - 1 >= nc) dg = ldM(v - 1 - nc);
                }
                const uint32_t ca = (uint32_t)(unsigned char)__ldg(gpA + min(v, RS) - 1), cb = (uint32_t)(unsigned char)__ldg(gpB + min(v, RS) - 1);
                const uint32_t mm = pack2::sub2_dyn(ca, cb, t2, (uint32_t)X, dg);
                const int j = (v - 1) & 15;
                const uint32_t oldM = upM;
                const uint32_t m = pack2::cell2<ALGO>(upM, upI, leftM, leftD, mm, OE2, E2, 1u << j, 1u << (16 + j), aP, aQ, aD, aI);
                st(v, m, upI);  // (a tail cell is read back by the tail cell num_cols columns further on, when the tail is that long)
                if (!tail) dg = oldM;
                if (v == nc) { tailM = m; tailI = upI; tailD = leftD; }
                leftM = m;
                if (h == tl) {  // the score is the last cell written of each pair: (tl, pl)
                    if (v == plA) scoreA = pack2::lo_half(m);
                    if (v == plB) scoreB = pack2::hi_half(m);
                }
                if (K.backtrace && (j == 15 || v == plm)) {
                    uint32_t *dst = frow + (size_t)((v - 1) >> 4) * fstep;
                    if (SWG) *reinterpret_cast<uint4 *>(dst) = make_uint4(aP, aQ, aD, aI);
                    else *reinterpret_cast<uint2 *>(dst) = make_uint2(aP, aQ);
                    aP = aQ = aD = aI = 0;
                }
            }
        }

        // ---- traceback of both pairs, one step of each per iteration: every step is a dependent load of a flag record from
        // global memory (L2 latency), and the two walks are independent ----
        int tb_h[2] = {tl, tl}, tb_v[2] = {plA, plB}, tb_b[2] = {plA + tl - 1, plB + tl - 1}, tb_layer[2] = {0, 0};
        int tb_status[2] = {AIM_STATUS_OK, AIM_STATUS_OK};
        bool tb_live[2] = {K.backtrace != 0, K.backtrace != 0 && !single};
        while (tb_live[0] || tb_live[1]) {
            uint32_t rec[2][4];
            int rr[2] = {1, 1}, cc[2] = {1, 1};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                rec[half][0] = rec[half][1] = rec[half][2] = rec[half][3] = 0u;
                if (tb_live[half] && tb_h[half] > 0 && tb_v[half] > 0) {
                    const int fi = nc * tb_h[half] + tb_v[half];  // flat word the reference's traceback reads
                    rr[half] = min(tl, (fi - 1) / nc);            // its last writer (row r, column c)
                    cc[half] = fi - nc * rr[half];
                    const uint32_t *src = flg + ((size_t)(rr[half] - 1) * rpr + (size_t)((cc[half] - 1) >> 4)) * fstep;
                    if (SWG) { const uint4 w = *reinterpret_cast<const uint4 *>(src); rec[half][0] = w.x; rec[half][1] = w.y; rec[half][2] = w.z; rec[half][3] = w.w; }
                    else { const uint2 w = *reinterpret_cast<const uint2 *>(src); rec[half][0] = w.x; rec[half][1] = w.y; }
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (!tb_live[half]) continue;
                const uint32_t i = half ? iB : iA;
                char *ops = K.ops + (size_t)i * 2 * RS;
                int &h = tb_h[half], &v = tb_v[half], &b = tb_b[half], &layer = tb_layer[half];
                if (!(h > 0 && v > 0)) {  // the walk left the table: the rest is one gap
                    while (h > 0) { ops[b--] = 'I'; --h; }
                    while (v > 0) { ops[b--] = 'D'; --v; }
                    tb_live[half] = false;
                    continue;
                }
                const char *gp = half ? gpB : gpA, *gt = half ? gtB : gtA;
                const int r = rr[half], c = cc[half];
                bool p, q, opD, opI;
                pack2::decode2(rec[half], (c - 1) & 15, half, SWG, p, q, opD, opI);
                if (!SWG) {
                    if (q) {
                        if (p) { ops[b--] = 'D'; --v; }
                        else { ops[b--] = 'I'; --h; }
                    } else {
                        if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                        --b; --h; --v;
                    }
                } else {
                    if (b < 0) { tb_status[half] = AIM_STATUS_BACKTRACE; tb_live[half] = false; continue; }
                    if (layer == 2) { ops[b--] = 'D'; if (opD) layer = 0; --v; }
                    else if (layer == 1) { ops[b--] = 'I'; if (opI) layer = 0; --h; }
                    else if (q) layer = p ? 2 : 1;
                    else {
                        if (gp[c - 1] != gt[r - 1]) ops[b] = 'X';
                        --b; --h; --v;
                    }
                }
            }
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            if (half && single) break;
            const int pl = half ? plB : plA;
            int score = half ? scoreB : scoreA;
            if (pl == 0 || tl == 0) score = 0;
            aim_result res;
            res.max_operations = pl + tl;
            res.begin_offset = (K.backtrace && tb_status[half] == AIM_STATUS_OK) ? tb_b[half] + 1 : pl + tl - 1;
            res.end_offset = pl + tl;
            res.score = score;
            res.status = tb_status[half];
            res.idx = K.idx_base + (half ? iB : iA);
            K.results[half ? iB : iA] = res;
        }
    }
}

// ---- bucketing of the aliased pairs by text_len ----
// classify2: non-aliased pairs -> list (front, any order); aliased pairs -> histogram over text_len.
__global__ void classify2_kernel(const int32_t *plen, const int32_t *tlen, uint32_t n, int RS, uint32_t *list, uint32_t *counters, uint32_t *hist)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool valid = i < n, alias = false;
    int tl = 0;
    if (valid) {
        const int pl = min(max(plen[i], 0), RS);
        tl = min(max(tlen[i], 0), RS);
        alias = pl > tl;
    }
    const uint32_t m0 = __ballot_sync(0xffffffffu, valid && !alias);
    uint32_t b0 = 0;
    if (lane == 0 && m0) b0 = atomicAdd(&counters[0], (uint32_t)__popc(m0));
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (valid && !alias) list[b0 + (uint32_t)__popc(m0 & ((1u << lane) - 1u))] = i;
    if (valid && alias) atomicAdd(&hist[tl], 1u);
}
// one block: bucket b starts at the even slot start[b]; counters[1] = number of slots; hist becomes the per-bucket cursor
__global__ void __launch_bounds__(1024) bucket_scan_kernel(uint32_t *hist, uint32_t nb, uint32_t *counters)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? ((hist[i] + 1u) & ~1u) : 0u;  // slots of the bucket, rounded up to even
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= (uint32_t)d) x += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = ws[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= (uint32_t)d) w += y; }
            ws[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (threadIdx.x >= 32 ? ws[(threadIdx.x >> 5) - 1] : 0u) + carry_s;
        if (i < nb) hist[i] = incl - v;  // first slot of the bucket = its cursor
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counters[1] = carry_s;
}
__global__ void bucket_scatter_kernel(const int32_t *plen, const int32_t *tlen, uint32_t n, int RS, uint32_t *cursor, uint32_t *rowlist)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pl = min(max(plen[i], 0), RS), tl = min(max(tlen[i], 0), RS);
    if (pl > tl) rowlist[atomicAdd(&cursor[tl], 1u)] = i;
}

#endif
