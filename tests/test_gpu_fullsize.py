"""GPU, at (a slice of) BASELINE.json's full sizes: bit-exact equality with the CPU oracle (score, span, op bytes:
oracle.check aligns every pair on all host threads and compares in place - a tie-break difference in the backtrace
fails here), plus the size-independent properties: every CIGAR is a valid edit script of its pair, consumes exactly
plen/tlen, and re-scores to the reported score under the run's penalties."""
import os

import numpy as np
import pytest

import aim_b200 as A
from oracle import oracle as O

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _oracle_equal(algo, res, ops, plen, tlen, pats, txts, stride=1, **kw):
    r = O.check(algo, plen, tlen, pats, txts, np.ascontiguousarray(res), ops, nthreads=THREADS, stride=stride, **kw)
    assert r["mismatches"] == 0, f"{r['mismatches']} of {r['pairs_checked']} pairs differ from the oracle, first at pair {r['first_bad']}"
    return r["pairs_checked"]


def _check_cigars(res, ops, plen, tlen, pats, txts, x, o, e, linear_gap=None, sample=4000, seed=0):
    n = len(res)
    # vectorised: op counts inside the span
    cols = np.arange(ops.shape[1])[None, :]
    span = (cols >= res["begin_offset"][:, None]) & (cols < res["end_offset"][:, None])
    nM = ((ops == ord("M")) & span).sum(1)
    nX = ((ops == ord("X")) & span).sum(1)
    nI = ((ops == ord("I")) & span).sum(1)
    nD = ((ops == ord("D")) & span).sum(1)
    assert np.array_equal(nM + nX + nD, plen), "CIGAR does not consume the pattern"
    assert np.array_equal(nM + nX + nI, tlen), "CIGAR does not consume the text"
    rng = np.random.default_rng(seed)
    for i in rng.choice(n, size=min(sample, n), replace=False):
        cig = ops[i, res["begin_offset"][i]:res["end_offset"][i]]
        v = h = 0
        score = 0
        prev = 0
        for c in cig:
            if c == 77:
                assert pats[i, v] == txts[i, h]; v += 1; h += 1
            elif c == 88:
                assert pats[i, v] != txts[i, h]; v += 1; h += 1; score += x
            elif c == 73:
                h += 1; score += (linear_gap if linear_gap is not None else (e if prev == 73 else o + e))
            else:
                v += 1; score += (linear_gap if linear_gap is not None else (e if prev == 68 else o + e))
            prev = c
        assert (v, h) == (plen[i], tlen[i])
        assert score == res["score"][i], f"pair {i}: CIGAR re-scores to {score}, reported {res['score'][i]}"


def test_cfg4_wfa_adaptive_1m_pairs():
    ms, rs = A.derive_knobs("wfa", 150, 0.04)
    n = 1_000_000
    plen, tlen, pats, txts = A.generate_pairs(4, n, 150, 0.04, rs)
    res, ops, _ = A.align_batch(A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=True, reduce=True), plen, tlen, pats, txts)
    assert int((res["status"] != 0).sum()) == 0
    assert int((res["score"] > ms).sum()) == 0  # 6 edits cost at most 6*5 = 30
    _check_cigars(res, ops, plen, tlen, pats, txts, 3, 4, 1)
    assert _oracle_equal("wfa", res, ops, plen, tlen, pats, txts, max_score=ms, read_size=rs, backtrace=True, reduce=True) == n
    # score-only run agrees on scores
    res2, _, _ = A.align_batch(A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=False, reduce=True), plen, tlen, pats, txts)
    assert np.array_equal(res["score"], res2["score"])


def test_cfg5_long_reads_scores_and_cigars():
    ms, rs = A.derive_knobs("wfa", 10000, 0.10)
    n = 256
    plen, tlen, pats, txts = A.generate_pairs(5, n, 10000, 0.10, rs)
    res, _, _ = A.align_batch(A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=False, reduce=True), plen, tlen, pats, txts)
    assert int((res["status"] != 0).sum()) == 0
    assert res["score"].min() > 2500 and res["score"].max() <= ms
    assert _oracle_equal("wfa", res, None, plen, tlen, pats, txts, max_score=ms, read_size=rs, backtrace=False, reduce=True) == n
    resb, ops, _ = A.align_batch(A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=True, reduce=True), plen, tlen, pats, txts)
    assert np.array_equal(res["score"], resb["score"])
    _check_cigars(resb, ops, plen, tlen, pats, txts, 3, 4, 1, sample=16)
    # with backtrace (beyond what the reference can run at this size): the oracle has no WRAM guard
    assert _oracle_equal("wfa", resb, ops, plen, tlen, pats, txts, stride=4, max_score=ms, read_size=rs, backtrace=True, reduce=True) == n // 4


def test_cfg3_swg_200k_pairs_oracle_equal():
    ms, rs = A.derive_knobs("swg", 250, 0.04, 4, 6, 2)
    n = 200_000
    plen, tlen, pats, txts = A.generate_pairs(3, n, 250, 0.04, rs)
    kw = dict(max_score=ms, read_size=rs, mismatch=4, gap_open=6, gap_ext=2, backtrace=True)
    res, ops, _ = A.align_batch(A.AlignParams(algo="swg", **kw), plen, tlen, pats, txts)
    assert int((res["status"] != 0).sum()) == 0
    assert _oracle_equal("swg", res, ops, plen, tlen, pats, txts, **kw) == n


def test_cfg2_nw_1m_pairs_oracle_equal():
    ms, rs = A.derive_knobs("nw", 100, 0.01)
    n = 1_000_000
    plen, tlen, pats, txts = A.generate_pairs(2, n, 100, 0.01, rs)
    kw = dict(max_score=ms, read_size=rs, backtrace=True)
    res, ops, _ = A.align_batch(A.AlignParams(algo="nw", **kw), plen, tlen, pats, txts)
    assert _oracle_equal("nw", res, ops, plen, tlen, pats, txts, **kw) == n


def test_cfg2_nw_linear_gap_cigars_on_dataset(golden_case):
    e, kw, (plen, tlen, pats, txts) = golden_case("cfg2_nw_err")
    res, ops, _ = A.align_batch(A.AlignParams(**kw), plen, tlen, pats, txts)
    _check_cigars(res, ops, plen, tlen, pats, txts, 3, 0, 0, linear_gap=4, sample=2000)
    okw = {k: v for k, v in kw.items() if k in ("max_score", "read_size", "match", "mismatch", "gap_open", "gap_ext", "backtrace", "reduce")}
    assert _oracle_equal(kw["algo"], res, ops, plen, tlen, pats, txts, **okw) == len(plen)
