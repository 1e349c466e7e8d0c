"""SWG/DPU-WRAM semantics: `variant = 1` = int8 cells when MAX_SCORE < 127
(SWG/DPU-WRAM/common/common.h:71-79).  Golden = the UNMODIFIED SWG/DPU-WRAM program's output bytes."""
import json
import lzma

import numpy as np
import pytest

from conftest import GOLDEN, md5_bytes, oracle_results_to_aim, render_output, assert_same_alignment
import aim_b200 as A
from oracle import oracle as O

MAN = {e["name"]: e for e in json.loads((GOLDEN / "swg8" / "manifest.json").read_text())}


def load(name, tmp_path):
    e = MAN[name]
    f = tmp_path / "in.pairs"
    f.write_bytes(lzma.open(GOLDEN / "swg8" / f"{name}.pairs.xz").read())
    return e, A.read_pairs(f, e["params"]["read_size"])


@pytest.mark.parametrize("name", sorted(MAN))
def test_oracle_int8_matches_reference_bytes(name, tmp_path):
    e, (plen, tlen, pats, txts) = load(name, tmp_path)
    p = e["params"]
    kw = dict(max_score=p["max_score"], read_size=p["read_size"], mismatch=p["mismatch"], gap_open=p["gap_open"], gap_ext=p["gap_ext"])
    res, ops = O.align("swg", plen, tlen, pats, txts, backtrace=True, variant=1, **kw)
    out = render_output(oracle_results_to_aim(res), ops, p["read_size"], True, tmp_path)
    assert md5_bytes(out) == e["md5"]
    assert out == lzma.open(GOLDEN / "swg8" / f"{name}.out.xz").read()
    r16, _ = O.align("swg", plen, tlen, pats, txts, backtrace=True, variant=0, **kw)
    differs = int((r16["score"] != res["score"]).sum())
    assert (differs == 0) == name.endswith("nowrap"), f"{differs} pairs differ from the int16 semantics"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MAN))
def test_gpu_int8_variant_matches_reference_bytes(name, tmp_path):
    e, (plen, tlen, pats, txts) = load(name, tmp_path)
    p = e["params"]
    kw = dict(max_score=p["max_score"], read_size=p["read_size"], mismatch=p["mismatch"], gap_open=p["gap_open"], gap_ext=p["gap_ext"])
    res, ops, _ = A.align_batch(A.AlignParams(algo="swg", backtrace=True, variant=1, **kw), plen, tlen, pats, txts)
    out = render_output(res, ops, p["read_size"], True, tmp_path)
    assert md5_bytes(out) == e["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("backtrace", [True, False])
def test_gpu_int8_variant_ragged_vs_oracle(backtrace):
    rs = 120
    n = 2000
    plen, tlen, pats, txts = A.generate_pairs(31, n, 100, 0.05, rs, nthreads=4)
    rng = np.random.default_rng(5)
    plen, tlen = plen.copy(), tlen.copy()
    cut = rng.integers(0, n, 300)
    plen[cut[:150]] = rng.integers(0, np.maximum(plen[cut[:150]], 1))
    tlen[cut[150:]] = rng.integers(0, np.maximum(tlen[cut[150:]], 1))
    kw = dict(max_score=30, read_size=rs, mismatch=4, gap_open=6, gap_ext=2)
    exp, eops = O.align("swg", plen, tlen, pats, txts, backtrace=backtrace, variant=1, nthreads=8, **kw)
    got, gops, _ = A.align_batch(A.AlignParams(algo="swg", backtrace=backtrace, variant=1, **kw), plen, tlen, pats, txts)
    ok = exp["status"] == 0
    assert np.array_equal(got["status"] == 0, ok)
    assert_same_alignment(got[ok], None if gops is None else gops[ok], exp[ok], None if eops is None else eops[ok], backtrace, "swg int8")
    # max_score >= 127 -> int16 cells, the variant flag changes nothing
    kw16 = dict(kw, max_score=127)
    a, _, _ = A.align_batch(A.AlignParams(algo="swg", variant=1, **kw16), plen, tlen, pats, txts)
    b, _, _ = A.align_batch(A.AlignParams(algo="swg", variant=0, **kw16), plen, tlen, pats, txts)
    assert np.array_equal(a["score"], b["score"])
