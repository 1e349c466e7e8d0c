"""CPU: the full-size parity checker (oracle.check / orc_check_batch) finds what it must find: a wrong score, a wrong
span, one flipped op byte inside a span; ignores bytes outside the span; and its sampling visits offset, offset+stride, ..."""
import numpy as np
import pytest

import aim_b200 as A
from oracle import oracle as O


@pytest.mark.parametrize("algo,kw,l,e", [
    ("wfa", dict(reduce=True), 150, 0.04),
    ("nw", dict(), 100, 0.02),
    ("swg", dict(mismatch=4, gap_open=6, gap_ext=2), 80, 0.05),
])
def test_check_batch_detects_differences(algo, kw, l, e):
    ms, rs = A.derive_knobs(algo, l, e, kw.get("mismatch", 3), kw.get("gap_open", 4), kw.get("gap_ext", 1))
    n = 600
    plen, tlen, pats, txts = A.generate_pairs(11, n, l, e, rs)
    okw = dict(max_score=ms, read_size=rs, backtrace=True, **kw)
    res, ops = O.align(algo, plen, tlen, pats, txts, nthreads=2, **okw)
    cand = np.zeros(n, A.RESULT_DTYPE)
    for f in ("max_operations", "begin_offset", "end_offset", "score", "status"):
        cand[f] = res[f]
    cand["idx"] = np.arange(n)
    r = O.check(algo, plen, tlen, pats, txts, cand, ops, nthreads=3, **okw)
    assert r == {"pairs_checked": n, "mismatches": 0, "first_bad": None}
    # bytes outside the span do not count
    o2 = ops.copy()
    i = int(np.argmax(res["begin_offset"] > 0))
    o2[i, 0] ^= 0x55
    assert O.check(algo, plen, tlen, pats, txts, cand, o2, nthreads=2, **okw)["mismatches"] == 0
    # one op byte inside the span, one score, one begin_offset
    o2[17, res["begin_offset"][17]] ^= 1
    c2 = cand.copy()
    c2["score"][40] += 1
    c2["begin_offset"][300] -= 1
    r = O.check(algo, plen, tlen, pats, txts, c2, o2, nthreads=4, **okw)
    assert r["mismatches"] == 3 and r["first_bad"] == 17
    # strided sample: 40 is visited with stride 20 offset 0, 17 and 300 + others are not all
    r = O.check(algo, plen, tlen, pats, txts, c2, o2, nthreads=3, stride=20, offset=0, **okw)
    assert r["pairs_checked"] == n // 20 and r["mismatches"] == 2 and r["first_bad"] == 40
    r = O.check(algo, plen, tlen, pats, txts, c2, o2, nthreads=3, stride=20, offset=17, **okw)
    assert r["mismatches"] == 1 and r["first_bad"] == 17


def test_check_batch_score_only():
    ms, rs = A.derive_knobs("wfa", 150, 0.04)
    plen, tlen, pats, txts = A.generate_pairs(5, 300, 150, 0.04, rs)
    res, _ = O.align("wfa", plen, tlen, pats, txts, max_score=ms, read_size=rs, backtrace=False, reduce=True)
    cand = np.zeros(300, A.RESULT_DTYPE)
    cand["score"] = res["score"]
    cand["begin_offset"] = -7  # not compared without backtrace
    assert O.check("wfa", plen, tlen, pats, txts, cand, None, max_score=ms, read_size=rs, backtrace=False, reduce=True)["mismatches"] == 0
    cand["score"][299] = 0
    assert O.check("wfa", plen, tlen, pats, txts, cand, None, max_score=ms, read_size=rs, backtrace=False, reduce=True)["first_bad"] == 299
