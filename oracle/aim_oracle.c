/* TEST INFRASTRUCTURE — CPU oracle for the aim_b200 parity tests.  NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library; aim_b200/ never links, imports or falls back to it.
 *
 * It restates, with RUNTIME parameters, the algorithms the reference (safaad/aim) compiles into
 * its DPU binaries with -D knobs.  Every function cites the reference file:line it follows
 * (paths under /root/reference).  Parity pinning: the reference has no tests or golden vectors
 * (SURVEY.md section 4); this restatement is pinned instead against the reference ITSELF,
 * compiled natively by oracle/refbuild.py (tests/test_oracle_vs_reference.py, run wherever
 * /root/reference exists) and against the committed fixtures in tests/golden/ that were
 * generated from those reference builds (tests/golden/make_golden.py).
 *
 * Semantics kept literally: int16 offsets/cells with truncation at the same assignments, the
 * -10 / NULL(-16384) sentinels, the flat DP array indexed num_cols*h+v with num_cols=text_len+1
 * (rows alias when pattern_len > text_len), MAX_SCORE as SWG border "infinity", the WFA give-up
 * at score > MAX_SCORE, and the backtrace tie-break order.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_ALGO_NW 0
#define ORC_ALGO_SWG 1
#define ORC_ALGO_WFA 2
#define ORC_ALGO_GENASM_DC 3
#define ORC_ALGO_GENASM_FILTER 4

#define ORC_OK 0
#define ORC_ERR_BACKTRACE 1 /* reference would print "No link found"/"No backtrace operation found" and exit(1) */

typedef struct {
    int32_t algo, match, mismatch, gap_open, gap_ext, max_score, read_size, backtrace, reduce;
    int32_t variant; /* GenASM-DC: 1 = DPU-MRAM-DC ('S' for substitutions, pattern 'N' is no wildcard: genasmDC.c of that directory) */
} orc_params;

typedef struct {
    int32_t max_operations, begin_offset, end_offset, score, status;
} orc_result;

#define OFFSET_NULL (INT16_MIN / 2) /* WFA/DPU-MRAM/common/common.h:95 */
#define MINI(a, b) (((a) <= (b)) ? (a) : (b))
#define MAXI(a, b) (((a) >= (b)) ? (a) : (b))

/* ------------------------------------------------------------------------------------------
 * WFA / WFA-adaptive.  One history record per score (WFA/DPU-MRAM/common/common.h:126-138).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int present; /* mramIdx[score] != 0 */
    int klo, khi, lo_base, hi_base;
    int m_null, i_null, d_null;
    int16_t *m, *i, *d; /* indexed by k - lo_base */
} wf_rec;

static void wf_free(wf_rec *r)
{
    free(r->m); free(r->i); free(r->d);
    memset(r, 0, sizeof(*r));
}

/* WFA/DPU-MRAM/dpu/wfa.c:143-190 allocate_new_score */
static void wf_alloc(wf_rec *r, int lo, int hi, int kernel)
{
    int len = hi - lo + 1;
    r->present = 1;
    r->m = (int16_t *)malloc(sizeof(int16_t) * len);
    r->d = (kernel == 3 || kernel == 1) ? (int16_t *)malloc(sizeof(int16_t) * len) : NULL;
    r->i = (kernel == 3 || kernel == 2) ? (int16_t *)malloc(sizeof(int16_t) * len) : NULL;
    r->d_null = r->d == NULL;
    r->i_null = r->i == NULL;
    r->m_null = 0;
    r->klo = r->lo_base = lo;
    r->khi = r->hi_base = hi;
}

#define WM(r, k) ((r)->m[(k) - (r)->lo_base])
#define WI(r, k) ((r)->i[(k) - (r)->lo_base])
#define WD(r, k) ((r)->d[(k) - (r)->lo_base])

/* WFA/DPU-MRAM/dpu/wfa.c:193-215 affine_wfa_extend */
static void wfa_extend(wf_rec *w, const char *pattern, const char *text, int plen, int tlen)
{
    if (!w->present || w->m_null) return;
    for (int k = w->klo; k <= w->khi; ++k) {
        int moffset = WM(w, k);
        if (moffset < 0) continue;
        int v = moffset - k, h = moffset, count = 0;
        while ((v < plen && h < tlen && v >= 0 && h >= 0) && pattern[v++] == text[h++]) ++count;
        WM(w, k) = (int16_t)(WM(w, k) + count);
    }
}

/* WFA/DPU-MRAM/dpu/wfa.c:70-141 affine_wfa_reduce_wvs (min_wavefront_length 10, max_distance_threshold 50) */
static void wfa_reduce(wf_rec *w, int plen, int tlen)
{
    const int min_wavefront_length = 10, max_distance_threshold = 50;
    int alignment_k = tlen - plen;
    if (!w->present || w->m_null) return;
    if ((w->khi - w->klo + 1) < min_wavefront_length) return;
    int min_distance = MAXI(plen, tlen);
    int klo = w->klo, khi = w->khi;
    for (int k = klo; k <= khi; ++k) {
        int16_t offset = WM(w, k);
        int distance = MAXI(plen - (offset - k), tlen - offset);
        min_distance = MINI(distance, min_distance);
    }
    int top_limit = MINI(alignment_k - 1, khi);
    for (int k = w->klo; k < top_limit; ++k) {
        int16_t offset = WM(w, k);
        int distance = MAXI(plen - (offset - k), tlen - offset);
        if ((distance - min_distance) <= max_distance_threshold) break;
        w->klo = w->klo + 1;
    }
    int bottom_limit = MAXI(alignment_k + 1, w->klo);
    for (int k = khi; k > bottom_limit; --k) {
        int16_t offset = WM(w, k);
        int distance = MAXI(plen - (offset - k), tlen - offset);
        if (distance - min_distance <= max_distance_threshold) break;
        w->khi = w->khi - 1;
    }
    if (w->klo > w->khi) {
        w->m_null = w->i_null = w->d_null = 1;
        w->khi = khi;
        w->klo = klo;
    }
}

/* WFA/DPU-MRAM/dpu/wfa.c:217-237 affine_wfa_end_reached */
static int wfa_end_reached(const wf_rec *w, int plen, int tlen)
{
    if (!w->present || w->m_null) return 0;
    int alignment_k = tlen - plen;
    if (w->klo <= alignment_k && w->khi >= alignment_k) {
        int offset = WM(w, alignment_k);
        if (offset >= tlen) return 1;
    }
    return 0;
}

/* WFA/DPU-MRAM/dpu/wfa.c:275-354 affine_wfa_compute_next + :238-273 affine_wfa_compute_offsets */
static void wfa_compute_next(wf_rec *hist, int score, const orc_params *p)
{
    int mismatch_score = score - p->mismatch;
    int o_score = score - p->gap_open - p->gap_ext;
    int e_score = score - p->gap_ext;
    const wf_rec *A = (mismatch_score < 0 || !hist[mismatch_score].present) ? NULL : &hist[mismatch_score];
    const wf_rec *B = (o_score < 0 || !hist[o_score].present) ? NULL : &hist[o_score];
    const wf_rec *E = (e_score < 0 || !hist[e_score].present) ? NULL : &hist[e_score];

    int m_sub_null = (A == NULL) || A->m_null;
    int m_o_null = (B == NULL) || B->m_null;
    int i_e_null = (E == NULL) || E->i_null || E->i == NULL;
    int d_e_null = (E == NULL) || E->d_null || E->d == NULL;
    int i_out_null = m_o_null && i_e_null;
    int d_out_null = m_o_null && d_e_null;

    wf_rec *w = &hist[score];
    if (m_sub_null && i_out_null && d_out_null) { w->present = 0; return; }

    int m_sub_lo = 1, m_sub_hi = -1, m_o_lo = 1, m_o_hi = -1, e_lo = 1, e_hi = -1;
    if (!m_sub_null) { m_sub_lo = A->klo; m_sub_hi = A->khi; }
    if (!m_o_null) { m_o_lo = B->klo; m_o_hi = B->khi; }
    if (!(i_e_null && d_e_null)) { e_lo = E->klo; e_hi = E->khi; }
    int lo = MINI(MINI(m_sub_lo, m_o_lo), e_lo) - 1;
    int hi = MAXI(MAXI(m_sub_hi, m_o_hi), e_hi) + 1;
    int kernel = ((!i_out_null) << 1) | (!d_out_null);
    wf_alloc(w, lo, hi, kernel);

    for (int k = lo; k <= hi; ++k) {
        int16_t ins = -10;
        if (!m_o_null || !i_e_null) {
            int16_t ins_g = (!m_o_null && m_o_lo <= k - 1 && k - 1 <= m_o_hi) ? WM(B, k - 1) : OFFSET_NULL;
            int16_t ins_i = (!i_e_null && e_lo <= k - 1 && k - 1 <= e_hi) ? WI(E, k - 1) : OFFSET_NULL;
            if (ins_g == OFFSET_NULL && ins_i == OFFSET_NULL) ins = OFFSET_NULL;
            else ins = (int16_t)(MAXI(ins_g, ins_i) + 1);
            WI(w, k) = ins;
        }
        int16_t del = -10;
        if (!m_o_null || !d_e_null) {
            int16_t del_g = (!m_o_null && m_o_lo <= k + 1 && k + 1 <= m_o_hi) ? WM(B, k + 1) : OFFSET_NULL;
            int16_t del_d = (!d_e_null && e_lo <= k + 1 && k + 1 <= e_hi) ? WD(E, k + 1) : OFFSET_NULL;
            del = MAXI(del_g, del_d);
            WD(w, k) = del;
        }
        int16_t sub = -10;
        if (!m_sub_null)
            sub = (int16_t)((m_sub_lo <= k && k <= m_sub_hi) ? WM(A, k) + 1 : OFFSET_NULL);
        int16_t nw = MAXI(sub, ins);
        WM(w, k) = MAXI(del, nw);
    }
}

/* WFA/DPU-MRAM/dpu/wfa_backtracing.c:219-375 affine_wavefronts_backtrace (+ getters :73-166).
 * ops is the 2*READ_SIZE 'M'-filled buffer; returns ORC_ERR_BACKTRACE on "No link found". */
static int wfa_backtrace(const wf_rec *hist, int alignment_score, int plen, int tlen, const orc_params *p,
                         char *ops, int ops_cap, int *begin_offset_io)
{
    int begin_offset = *begin_offset_io;
    int alignment_k = tlen - plen;
    int score = alignment_score, k = alignment_k;
    int16_t offset = WM(&hist[alignment_score], k);
    int v = offset - k, h = offset;
    int valid_location = (v > 0 && v <= plen && h > 0 && h <= tlen);
    enum { T_M, T_I, T_D } type = T_M;
#define PUT(c) do { if (begin_offset < 0 || begin_offset >= ops_cap) return ORC_ERR_BACKTRACE; ops[begin_offset--] = (c); } while (0)

    while (v > 0 && h > 0 && score > 0) {
        if (!valid_location) {
            valid_location = (v > 0 && v <= plen && h > 0 && h <= tlen);
            if (valid_location) { /* wfa_backtracing.c:48-69 add_trailing_gap */
                if (k < alignment_k) { for (int i = k; i < alignment_k; ++i) PUT('I'); }
                else if (k > alignment_k) { for (int i = alignment_k; i < k; ++i) PUT('D'); }
            }
        }
        int gap_open_score = score - p->gap_open - p->gap_ext;
        int gap_extend_score = score - p->gap_ext;
        int mismatch_score = score - p->mismatch;
        const wf_rec *go = (gap_open_score < 0 || !hist[gap_open_score].present) ? NULL : &hist[gap_open_score];
        const wf_rec *ge = (gap_extend_score < 0 || !hist[gap_extend_score].present) ? NULL : &hist[gap_extend_score];
        const wf_rec *mm = (mismatch_score < 0 || !hist[mismatch_score].present) ? NULL : &hist[mismatch_score];

        int16_t del_ext = OFFSET_NULL, del_open = OFFSET_NULL, ins_ext = OFFSET_NULL, ins_open = OFFSET_NULL, misms = OFFSET_NULL;
        if (type != T_I) {
            if (ge && !ge->d_null && ge->klo <= k + 1 && k + 1 <= ge->khi) del_ext = WD(ge, k + 1);
            if (gap_open_score >= 0 && go && go->klo <= k + 1 && k + 1 <= go->khi) del_open = WM(go, k + 1);
        }
        if (type != T_D) {
            if (gap_extend_score >= 0 && ge && ge->i != NULL && !ge->i_null && ge->klo <= k - 1 && k - 1 <= ge->khi)
                ins_ext = (int16_t)(WI(ge, k - 1) + 1);
            if (gap_open_score >= 0 && go && go->klo <= k - 1 && k - 1 <= go->khi) ins_open = (int16_t)(WM(go, k - 1) + 1);
        }
        if (type == T_M) {
            if (mismatch_score >= 0 && mm && mm->klo <= k && k <= mm->khi) misms = (int16_t)(WM(mm, k) + 1);
        }
        int16_t max_del = MAXI(del_ext, del_open);
        int16_t max_ins = MAXI(ins_ext, ins_open);
        int16_t max_all = MAXI(misms, MAXI(max_ins, max_del));

        if (type == T_M) {
            int num_matches = offset - max_all;
            for (int i = 0; i < num_matches; ++i) PUT('M');
            offset = max_all;
            v = offset - k; h = offset;
            if (v <= 0 || h <= 0) break;
        }
        if (max_all == del_ext) { if (valid_location) PUT('D'); score = gap_extend_score; ++k; type = T_D; }
        else if (max_all == del_open) { if (valid_location) PUT('D'); score = gap_open_score; ++k; type = T_M; }
        else if (max_all == ins_ext) { if (valid_location) PUT('I'); score = gap_extend_score; --k; --offset; type = T_I; }
        else if (max_all == ins_open) { if (valid_location) PUT('I'); score = gap_open_score; --k; --offset; type = T_M; }
        else if (max_all == misms) { if (valid_location) PUT('X'); score = mismatch_score; --offset; }
        else return ORC_ERR_BACKTRACE;
        v = offset - k; h = offset;
    }
    if (score == 0) {
        for (int i = 0; i < offset; ++i) PUT('M');
    } else {
        while (v > 0) { PUT('D'); --v; }
        while (h > 0) { PUT('I'); --h; }
    }
#undef PUT
    *begin_offset_io = begin_offset + 1;
    return ORC_OK;
}

/* Note on i_null in the ins_ext getter: load_idwavefront_cmpnt_from_mram (dpu_allocator_mram.c:268-346)
 * leaves iwavefront NULL exactly when the stored i_null flag is set, so "iwavefront != NULL"
 * (wfa_backtracing.c:124) == !i_null. */

/* WFA/DPU-MRAM/dpu/wfa.c:356-407 affine_wfa_compute + :499-501 ops init + :58-68 cigar init */
static void wfa_align(const orc_params *p, const char *pattern, const char *text, int plen, int tlen,
                      orc_result *res, char *ops)
{
    res->max_operations = plen + tlen;
    res->begin_offset = res->max_operations - 1;
    res->end_offset = res->max_operations;
    res->score = INT32_MIN;
    res->status = ORC_OK;
    if (p->backtrace && ops) memset(ops, 'M', 2 * (size_t)p->read_size);

    wf_rec *hist = (wf_rec *)calloc((size_t)p->max_score + 2, sizeof(wf_rec));
    wf_alloc(&hist[0], 0, 0, 0);
    hist[0].m[0] = 0;
    int score = 0;
    for (;;) {
        wf_rec *w = &hist[score];
        wfa_extend(w, pattern, text, plen, tlen);
        if (p->reduce) wfa_reduce(w, plen, tlen);
        if (wfa_end_reached(w, plen, tlen)) {
            if (p->backtrace && ops)
                res->status = wfa_backtrace(hist, score, plen, tlen, p, ops, 2 * p->read_size, &res->begin_offset);
            res->score = score;
            break;
        }
        ++score;
        if (score > p->max_score) { res->score = score; break; }
        wfa_compute_next(hist, score, p);
    }
    for (int s = 0; s <= p->max_score + 1; ++s) wf_free(&hist[s]);
    free(hist);
}

/* ------------------------------------------------------------------------------------------
 * NW, linear gap (GAP_I = GAP_D = gap_open).  NW/DPU-WRAM/dpu/nw.c:109-153 nw_compute and
 * :67-107 nw_traceback on the flat int16 table; the NW/DPU-MRAM variant (nw.c:151-236, 91-149)
 * computes the same values through one-cell MRAM caches.
 * ------------------------------------------------------------------------------------------ */
static void nw_align(const orc_params *p, const char *pattern, const char *text, int plen, int tlen,
                     orc_result *res, char *ops, int16_t *tab)
{
    const int GAP_I = p->gap_open, GAP_D = p->gap_open, MISMATCH = p->mismatch;
    res->max_operations = plen + tlen;
    res->begin_offset = res->max_operations - 1;
    res->end_offset = res->max_operations;
    res->score = 0;
    res->status = ORC_OK;
    int num_cols = tlen + 1;
    int cell = 0;
    tab[0] = (int16_t)cell;
    for (int v = 1; v <= plen; ++v) { cell += GAP_D; tab[v] = (int16_t)cell; }
    cell = 0;
    for (int h = 1; h <= tlen; ++h) { cell += GAP_I; tab[num_cols * h] = (int16_t)cell; }
    int16_t score = 0;
    for (int h = 1; h <= tlen; ++h) {
        for (int v = 1; v <= plen; ++v) {
            int16_t del = (int16_t)(tab[num_cols * h + v - 1] + GAP_D);
            int16_t ins = (int16_t)(tab[num_cols * (h - 1) + v] + GAP_I);
            int16_t m_match = (int16_t)(tab[num_cols * (h - 1) + v - 1] + ((pattern[v - 1] == text[h - 1]) ? 0 : MISMATCH));
            score = tab[num_cols * h + v] = (int16_t)MINI(m_match, MINI(ins, del));
        }
    }
    res->score = (int)score;
    if (!(p->backtrace && ops)) return;
    memset(ops, 'M', 2 * (size_t)p->read_size);
    int op_sentinel = res->end_offset - 1;
    int h = num_cols - 1, v = plen;
    while (h > 0 && v > 0) {
        if (tab[num_cols * h + v] == tab[num_cols * h + v - 1] + GAP_D) { ops[op_sentinel--] = 'D'; --v; }
        else if (tab[num_cols * h + v] == tab[num_cols * (h - 1) + v] + GAP_I) { ops[op_sentinel--] = 'I'; --h; }
        else {
            ops[op_sentinel--] = (tab[num_cols * h + v] == tab[num_cols * (h - 1) + v - 1] + MISMATCH) ? 'X' : 'M';
            --h; --v;
        }
    }
    while (h > 0) { ops[op_sentinel--] = 'I'; --h; }
    while (v > 0) { ops[op_sentinel--] = 'D'; --v; }
    res->begin_offset = op_sentinel + 1;
}

/* ------------------------------------------------------------------------------------------
 * SWG, gap-affine, int16 cells, MAX_SCORE borders.  SWG/DPU-MRAM/dpu/swg.c:151-217 swg_compute
 * and :66-148 swg_traceback (the MRAM variant is canonical: SWG/DPU-WRAM switches to int8 cells
 * when MAX_SCORE < 127, SWG/DPU-WRAM/common/common.h:71-79).
 * ------------------------------------------------------------------------------------------ */
typedef struct { int16_t M, I, D, pad; } swg_cell;

static void swg_align(const orc_params *p, const char *pattern, const char *text, int plen, int tlen,
                      orc_result *res, char *ops, swg_cell *tab)
{
    /* SWG/DPU-WRAM (variant 1): cell_size_t is int8 when MAX_SCORE < 127 (SWG/DPU-WRAM/common/common.h:71-79), truncation at the
     * same assignments (SWG/DPU-WRAM/dpu/swg.c:128-165) */
    const int w8 = p->variant == 1 && p->max_score < 127;
#define CT(x) ((int16_t)(w8 ? (int)(int8_t)(x) : (int)(int16_t)(x)))
    const int GAP_O = p->gap_open, GAP_E = p->gap_ext, MATCH = p->match, MISMATCH = p->mismatch, MAX_SCORE = p->max_score;
    res->max_operations = plen + tlen;
    res->begin_offset = res->max_operations - 1;
    res->end_offset = res->max_operations;
    res->score = INT32_MIN;
    res->status = ORC_OK;
    if (p->backtrace && ops) memset(ops, 'M', 2 * (size_t)p->read_size);
    int num_cols = tlen + 1;
    tab[0].D = CT(MAX_SCORE); tab[0].I = CT(MAX_SCORE); tab[0].M = 0;
    for (int v = 1; v <= plen; ++v) {
        tab[v].D = CT(GAP_O + v * GAP_E);
        tab[v].I = CT(MAX_SCORE);
        tab[v].M = tab[v].D;
    }
    for (int h = 1; h <= tlen; ++h) {
        swg_cell *c = &tab[num_cols * h];
        c->D = CT(MAX_SCORE);
        c->I = CT(GAP_O + h * GAP_E);
        c->M = c->I;
    }
    int score = 0;
    for (int h = 1; h <= tlen; ++h) {
        for (int v = 1; v <= plen; ++v) {
            swg_cell upper = tab[num_cols * h + v - 1];
            swg_cell diag = tab[num_cols * (h - 1) + v - 1];
            swg_cell left = tab[num_cols * (h - 1) + v];
            swg_cell cur;
            int16_t del_new = CT(upper.M + GAP_O + GAP_E);
            int16_t del_ext = CT(upper.D + GAP_E);
            int16_t del = MINI(del_new, del_ext);
            cur.D = del;
            int16_t ins_new = CT(left.M + GAP_O + GAP_E);
            int16_t ins_ext = CT(left.I + GAP_E);
            int16_t ins = MINI(ins_new, ins_ext);
            cur.I = ins;
            int16_t m_match = CT(diag.M + ((pattern[v - 1] == text[h - 1]) ? MATCH : MISMATCH));
            cur.M = MINI(m_match, MINI(ins, del));
            cur.pad = 0;
            score = cur.M;
            tab[num_cols * h + v] = cur;
        }
    }
    res->score = score;
    if (!(p->backtrace && ops)) return;
    int op_sentinel = res->end_offset - 1;
    int h = num_cols - 1, v = plen;
    enum { L_M, L_I, L_D } layer = L_M;
    while (h > 0 && v > 0) {
        swg_cell cell = tab[num_cols * h + v];
        swg_cell upper = tab[num_cols * h + v - 1];
        swg_cell diag = tab[num_cols * (h - 1) + v - 1];
        swg_cell left = tab[num_cols * (h - 1) + v];
        if (op_sentinel < 0) { res->status = ORC_ERR_BACKTRACE; return; }
        switch (layer) {
        case L_D:
            ops[op_sentinel--] = 'D';
            if (cell.D == upper.M + GAP_O + GAP_E) layer = L_M;
            --v;
            break;
        case L_I:
            ops[op_sentinel--] = 'I';
            if (cell.I == left.M + GAP_O + GAP_E) layer = L_M;
            --h;
            break;
        case L_M:
            if (cell.M == cell.D) layer = L_D;
            else if (cell.M == cell.I) layer = L_I;
            else if (cell.M == diag.M + MATCH) { ops[op_sentinel--] = 'M'; --h; --v; }
            else if (cell.M == diag.M + MISMATCH) { ops[op_sentinel--] = 'X'; --h; --v; }
            else { res->status = ORC_ERR_BACKTRACE; return; }
            break;
        }
    }
    while (h > 0) { ops[op_sentinel--] = 'I'; --h; }
    while (v > 0) { ops[op_sentinel--] = 'D'; --v; }
    res->begin_offset = op_sentinel + 1;
}

/* ------------------------------------------------------------------------------------------
 * GenASM-DC / GenASM-filter (aim-genasm submodule).
 * Follows aim-genasm/GenASM/DPU-WRAM-DC/dpu/genasmDC.c: pattern bitmasks :40-88, genasmDC :338-556
 * (bit-vector fill + traceback matrix), genasmTB :90-336 (traceback, CIGAR with REVERSED count digits),
 * score :553; filter = aim-genasm/GenASM/DPU-WRAM-filter/dpu/genasm_filter.c:52-239 (same fill, no
 * traceback, returns the minimum error level).  Bit vectors are `count` 64-bit words, word 0 the MOST
 * significant, count = (m + 64) / 64; the end test reads word 0 with bit (m % 64 ? m % 64 - 1 : 63)
 * (genasmDC.c:389-399,530-541) - kept literally, so m % 64 == 0 never finds an alignment.
 * Where the reference reads memory it never wrote for this pair the result is UNDEFINED there
 * (it depends on what earlier pairs left in the tasklet's MRAM segment): a text byte outside ACGTacgt skips
 * the whole step including the traceback rows (genasmDC.c:430-433), and the traceback may step to text row n
 * (pattern longer than the text consumed).  Those pairs get status ORC_GENASM_UNDEFINED and are excluded
 * from parity; "No alignment found" (score -1, CIGAR buffer untouched) gets ORC_GENASM_NOALIGN.
 * ------------------------------------------------------------------------------------------ */
#define ORC_GENASM_UNDEFINED 3
#define ORC_GENASM_NOALIGN 4
typedef unsigned long long u64_t;

static int genasm_code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

static void genasm_shl1(u64_t *dst, const u64_t *src, int count)
{   /* genasmDC.c:474-479: word 0 is the most significant */
    dst[0] = src[0] << 1;
    for (int a = 1; a < count; a++) { dst[a - 1] |= (src[a] >> 63); dst[a] = src[a] << 1; }
}

static void genasm_align(const orc_params *p, const char *pattern, const char *text, int m, int n, orc_result *res,
                         char *cigar, int with_tb)
{
    const int k = p->max_score;
    const int count = (int)((m + 64) / 64.0);
    const u64_t max = ~0ULL;
    res->max_operations = m + n;
    res->begin_offset = 0;
    res->end_offset = 0;
    res->score = -1;
    res->status = ORC_OK;
    /* pattern bitmasks (genasmDC.c:40-88) */
    u64_t *pm = (u64_t *)malloc(sizeof(u64_t) * 4 * count);
    for (int i = 0; i < 4 * count; i++) pm[i] = max;
    for (int i = 0; i < m; i++) {
        const int index = count - ((m - i - 1) / 64) - 1;
        const u64_t bit = ~(1ULL << ((m - i - 1) % 64));
        const int c = genasm_code(pattern[i]);
        if (c >= 0) pm[c * count + index] &= bit;
        else if ((pattern[i] == 'N' || pattern[i] == 'n') && !p->variant) for (int q = 0; q < 4; q++) pm[q * count + index] &= bit;
    }
    const int len1 = (k + 1) * count;
    u64_t *R = (u64_t *)malloc(sizeof(u64_t) * len1), *oldR = (u64_t *)malloc(sizeof(u64_t) * len1);
    u64_t *sub = (u64_t *)malloc(sizeof(u64_t) * count * 4), *ins = sub + count, *mat = ins + count, *del = mat + count;
    /* traceback matrix [n][k+1][4][count] (genasmDC.c:386) + which text rows were written */
    u64_t *tb = with_tb ? (u64_t *)malloc(sizeof(u64_t) * (size_t)(n > 0 ? n : 1) * (k + 1) * 4 * count) : NULL;
    char *row_written = with_tb ? (char *)calloc((size_t)n + 1, 1) : NULL;
    const int rem = m % 64;
    const u64_t max1 = rem == 0 ? (1ULL << 63) : (1ULL << (rem - 1));
    for (int i = 0; i < len1; i++) R[i] = max;
    for (int x = 1; x < k + 1; x++) {   /* genasmDC.c:406-425 */
        if ((x % 64) == 0) {
            const int ind = count - (x / 64);
            for (int y = count - 1; y >= ind && y >= 0; y--) R[x * count + y] = 0ULL;
        } else {
            const int ind = count - 1 - (x / 64);
            for (int y = count - 1; y > ind && y >= 0; y--) R[x * count + y] = 0ULL;
            if (ind >= 0) R[x * count + ind] = max << (x % 64);
        }
    }
    for (int i = n - 1; i >= 0; i--) {   /* genasmDC.c:428-527 */
        const int c = genasm_code(text[i]);
        if (c < 0) continue;
        const u64_t *cur = pm + c * count;
        memcpy(oldR, R, sizeof(u64_t) * len1);
        genasm_shl1(R, oldR, count);
        for (int a = 0; a < count; a++) R[a] |= cur[a];
        if (with_tb) {
            row_written[i] = 1;
            for (int a = 0; a < count; a++) {
                u64_t *t = tb + (((size_t)i * (k + 1) + 0) * 4) * count;
                t[0 * count + a] = R[a]; t[1 * count + a] = max; t[2 * count + a] = max; t[3 * count + a] = max;
            }
        }
        for (int d = 1; d <= k; d++) {
            int index = (d - 1) * count;
            for (int a = 0; a < count; a++) del[a] = oldR[index + a];
            genasm_shl1(sub, del, count);
            genasm_shl1(ins, R + index, count);
            index += count;
            genasm_shl1(mat, oldR + index, count);
            for (int a = 0; a < count; a++) mat[a] |= cur[a];
            for (int a = 0; a < count; a++) R[index + a] = del[a] & sub[a] & ins[a] & mat[a];
            if (with_tb) {
                u64_t *t = tb + (((size_t)i * (k + 1) + d) * 4) * count;
                for (int a = 0; a < count; a++) {
                    t[0 * count + a] = mat[a]; t[1 * count + a] = sub[a]; t[2 * count + a] = ins[a]; t[3 * count + a] = del[a];
                }
            }
        }
    }
    int minError = -1;
    for (int t = 0; t <= k; t++) if ((R[t * count] & max1) == 0) { minError = t; break; }   /* genasmDC.c:530-541 */
    if (!with_tb) { res->score = minError; res->max_operations = 0; goto out; }   /* genasm_filter.c:226-238 */
    if (minError == -1) { res->status = ORC_GENASM_NOALIGN; goto out; }   /* genasmDC.c:543-547 */
    for (int i = 0; i < n; i++) if (!row_written[i]) { res->status = ORC_GENASM_UNDEFINED; goto out; }
    {   /* genasmTB (genasmDC.c:90-336) */
        int curPattern = m - 1, curText = 0, curError = minError, c = 0;
        int countM = 0, countS = 0, countOpen = 0, countExtend = 0, charCount = 0, isFirst = 1;
        char lastChar = '0';
        const char subChar = p->variant ? 'S' : 'X';
        u64_t mask = max1;
#define GENASM_FLUSH() do { if (!isFirst) { int num = charCount; while (num != 0) { cigar[c++] = (char)(num % 10 + '0'); num /= 10; } cigar[c++] = lastChar; } } while (0)
        while (curPattern >= 0 && curError >= 0) {
            if (curText >= n) { res->status = ORC_GENASM_UNDEFINED; goto out; }   /* reads rows the pair never wrote */
            const int ind = count - (curPattern / 64) - 1;
            const u64_t *t = tb + (((size_t)curText * (k + 1) + curError) * 4) * count;
            const u64_t t0 = t[0 * count + ind], t1 = t[1 * count + ind], t2 = t[2 * count + ind], t3 = t[3 * count + ind];
            if (lastChar == 'I' && (t2 & mask) == 0) {          /* affine-insertion: always an extension */
                curPattern -= 1; curError -= 1;
                charCount += 1; countExtend += 1;
            } else if (lastChar == 'D' && (t3 & mask) == 0) {   /* affine-deletion: always an extension */
                curText += 1; curError -= 1;
                charCount += 1; countExtend += 1;
            } else if ((t0 & mask) == 0) {                      /* match */
                curText += 1; curPattern -= 1;
                if (lastChar == 'M') charCount += 1;
                else { GENASM_FLUSH(); charCount = 1; lastChar = 'M'; }
                countM += 1;
            } else if ((t1 & mask) == 0) {                      /* substitution */
                curText += 1; curPattern -= 1; curError -= 1;
                if (lastChar == subChar) charCount += 1;
                else { GENASM_FLUSH(); charCount = 1; lastChar = subChar; }
                countS += 1;
            } else if ((t3 & mask) == 0) {                      /* deletion (lastChar != 'D' here): opens */
                curText += 1; curError -= 1;
                GENASM_FLUSH(); charCount = 1; lastChar = 'D'; countOpen += 1;
            } else if ((t2 & mask) == 0) {                      /* insertion (lastChar != 'I' here): opens */
                curPattern -= 1; curError -= 1;
                GENASM_FLUSH(); charCount = 1; lastChar = 'I'; countOpen += 1;
            } else { res->status = ORC_GENASM_UNDEFINED; goto out; }   /* the reference would spin forever */
            if (curPattern >= 0) mask = 1ULL << (curPattern % 64);
            isFirst = 0;
        }
        { int num = charCount; while (num != 0) { cigar[c++] = (char)(num % 10 + '0'); num /= 10; } }
        cigar[c++] = lastChar;
        cigar[c] = '\0';
#undef GENASM_FLUSH
        res->max_operations = c + 1;
        res->end_offset = c;
        res->score = countM * p->match + countS * p->mismatch + countOpen * (p->gap_open + p->gap_ext) + countExtend * p->gap_ext;
    }
out:
    free(pm); free(R); free(oldR); free(sub); free(tb); free(row_written);
}

/* ------------------------------------------------------------------------------------------
 * Batch driver: contiguous partition of pairs over host threads, like the host's per-DPU split
 * (WFA/DPU-MRAM/host/host.c:191-209).  Buffers are laid out as host.c:126-127 lays them out:
 * pair i's pattern at patterns + i*read_size, ops at ops + i*2*read_size.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_params *p;
    uint32_t first, last;
    const int32_t *plen, *tlen;
    const char *patterns, *texts;
    orc_result *results;
    char *ops;
} orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    const orc_params *p = j->p;
    size_t rs = (size_t)p->read_size;
    void *tab = NULL;
    if (p->algo == ORC_ALGO_NW) tab = malloc(sizeof(int16_t) * (rs + 2) * (rs + 2));
    else if (p->algo == ORC_ALGO_SWG) tab = malloc(sizeof(swg_cell) * (rs + 2) * (rs + 2));
    for (uint32_t i = j->first; i < j->last; ++i) {
        const char *pat = j->patterns + (size_t)i * rs, *txt = j->texts + (size_t)i * rs;
        char *ops = j->ops ? j->ops + (size_t)i * 2 * rs : NULL;
        if (p->algo == ORC_ALGO_GENASM_DC) genasm_align(p, pat, txt, j->plen[i], j->tlen[i], &j->results[i], ops, 1);
        else if (p->algo == ORC_ALGO_GENASM_FILTER) genasm_align(p, pat, txt, j->plen[i], j->tlen[i], &j->results[i], NULL, 0);
        else if (p->algo == ORC_ALGO_WFA) wfa_align(p, pat, txt, j->plen[i], j->tlen[i], &j->results[i], ops);
        else if (p->algo == ORC_ALGO_NW) nw_align(p, pat, txt, j->plen[i], j->tlen[i], &j->results[i], ops, (int16_t *)tab);
        else swg_align(p, pat, txt, j->plen[i], j->tlen[i], &j->results[i], ops, (swg_cell *)tab);
    }
    free(tab);
    return NULL;
}

int orc_align_batch(const orc_params *p, uint32_t n, const int32_t *plen, const int32_t *tlen,
                    const char *patterns, const char *texts, orc_result *results, char *ops, int nthreads)
{
    if (!p || p->algo < 0 || p->algo > 4 || p->read_size <= 0) return -1;
    for (uint32_t i = 0; i < n; ++i)
        if (plen[i] < 0 || tlen[i] < 0 || plen[i] > p->read_size || tlen[i] > p->read_size) return -2;
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
    orc_job *jobs = (orc_job *)calloc((size_t)nthreads, sizeof(orc_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    uint32_t per = (n + (uint32_t)nthreads - 1) / (uint32_t)nthreads;
    for (int t = 0; t < nthreads; ++t) {
        uint32_t a = (uint32_t)t * per, b = a + per;
        if (a > n) a = n;
        if (b > n) b = n;
        orc_job j = { p, a, b, plen, tlen, patterns, texts, results, (p->backtrace || p->algo == ORC_ALGO_GENASM_DC) ? ops : NULL };
        jobs[t] = j;
        if (nthreads == 1) orc_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Checker for full-size parity runs (bench.py "parity", tests/test_gpu_fullsize.py): align every
 * pair here and compare with a candidate's output in place - score, begin/end offsets,
 * max_operations and the op bytes of the valid span [begin_offset, end_offset) - without a second
 * n x 2*read_size op array.  `cand_results` uses the product's 24-byte result layout
 * (include/aim_b200.h aim_result: the reference's result_t with status in the pad word + idx).
 * Returns the number of mismatching pairs through *mismatches and the lowest mismatching index
 * through *first_bad (UINT32_MAX when none).  stride/offset select pairs offset, offset+stride, ...
 * (stride 1 = all) so a bounded sample can be spread over the whole batch.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t max_operations, begin_offset, end_offset, score, status;
    uint32_t idx;
} orc_cand_result;

typedef struct {
    orc_job job;
    const orc_cand_result *cand;
    const char *cand_ops;
    uint32_t stride, offset;
    uint64_t mismatches, checked;
    uint32_t first_bad;
} orc_check_job;

static void *orc_check_worker(void *arg)
{
    orc_check_job *c = (orc_check_job *)arg;
    orc_job *j = &c->job;
    const orc_params *p = j->p;
    size_t rs = (size_t)p->read_size;
    const int has_ops = (p->backtrace || p->algo == ORC_ALGO_GENASM_DC) && c->cand_ops;
    void *tab = NULL;
    char *row = (char *)malloc(2 * rs + 16);
    if (p->algo == ORC_ALGO_NW) tab = malloc(sizeof(int16_t) * (rs + 2) * (rs + 2));
    else if (p->algo == ORC_ALGO_SWG) tab = malloc(sizeof(swg_cell) * (rs + 2) * (rs + 2));
    c->first_bad = UINT32_MAX;
    uint32_t i = j->first;
    if (c->stride > 1) { uint32_t r = i % c->stride; i += (c->offset + c->stride - r) % c->stride; }
    for (; i < j->last; i += c->stride) {
        const char *pat = j->patterns + (size_t)i * rs, *txt = j->texts + (size_t)i * rs;
        orc_result r;
        memset(&r, 0, sizeof r);
        char *ops = has_ops || p->backtrace ? row : NULL;
        if (p->algo == ORC_ALGO_GENASM_DC) genasm_align(p, pat, txt, j->plen[i], j->tlen[i], &r, row, 1);
        else if (p->algo == ORC_ALGO_GENASM_FILTER) genasm_align(p, pat, txt, j->plen[i], j->tlen[i], &r, NULL, 0);
        else if (p->algo == ORC_ALGO_WFA) wfa_align(p, pat, txt, j->plen[i], j->tlen[i], &r, ops);
        else if (p->algo == ORC_ALGO_NW) nw_align(p, pat, txt, j->plen[i], j->tlen[i], &r, ops, (int16_t *)tab);
        else swg_align(p, pat, txt, j->plen[i], j->tlen[i], &r, ops, (swg_cell *)tab);
        const orc_cand_result *g = &c->cand[i];
        int bad = g->score != r.score || g->status != r.status;
        if (p->backtrace || p->algo == ORC_ALGO_GENASM_DC)
            bad |= g->begin_offset != r.begin_offset || g->end_offset != r.end_offset || g->max_operations != r.max_operations;
        if (!bad && has_ops && r.status == ORC_OK && r.end_offset > r.begin_offset && r.begin_offset >= 0)
            bad = memcmp(c->cand_ops + (size_t)i * 2 * rs + r.begin_offset, row + r.begin_offset, (size_t)(r.end_offset - r.begin_offset)) != 0;
        ++c->checked;
        if (bad) { ++c->mismatches; if (i < c->first_bad) c->first_bad = i; }
    }
    free(tab); free(row);
    return NULL;
}

int orc_check_batch(const orc_params *p, uint32_t n, const int32_t *plen, const int32_t *tlen, const char *patterns,
                    const char *texts, const void *cand_results, const char *cand_ops, uint32_t stride, uint32_t offset,
                    int nthreads, uint64_t *checked, uint64_t *mismatches, uint32_t *first_bad)
{
    if (!p || p->algo < 0 || p->algo > 4 || p->read_size <= 0 || !cand_results) return -1;
    for (uint32_t i = 0; i < n; ++i)
        if (plen[i] < 0 || tlen[i] < 0 || plen[i] > p->read_size || tlen[i] > p->read_size) return -2;
    if (stride < 1) stride = 1;
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
    orc_check_job *jobs = (orc_check_job *)calloc((size_t)nthreads, sizeof(orc_check_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    uint32_t per = (n + (uint32_t)nthreads - 1) / (uint32_t)nthreads;
    for (int t = 0; t < nthreads; ++t) {
        uint32_t a = (uint32_t)t * per, b = a + per;
        if (a > n) a = n;
        if (b > n) b = n;
        orc_job j = { p, a, b, plen, tlen, patterns, texts, NULL, NULL };
        jobs[t].job = j;
        jobs[t].cand = (const orc_cand_result *)cand_results;
        jobs[t].cand_ops = cand_ops;
        jobs[t].stride = stride;
        jobs[t].offset = offset % stride;
        if (nthreads == 1) orc_check_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, orc_check_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    uint64_t mm = 0, ck = 0;
    uint32_t fb = UINT32_MAX;
    for (int t = 0; t < nthreads; ++t) { mm += jobs[t].mismatches; ck += jobs[t].checked; if (jobs[t].first_bad < fb) fb = jobs[t].first_bad; }
    if (checked) *checked = ck;
    if (mismatches) *mismatches = mm;
    if (first_bad) *first_bad = fb;
    free(jobs); free(th);
    return 0;
}
