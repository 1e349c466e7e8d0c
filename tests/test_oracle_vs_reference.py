"""CPU, authoring container only: pin the oracle restatement against the reference ITSELF, compiled
natively from /root/reference by oracle/refbuild.py, on fresh seeded inputs (not the committed
goldens).  Skipped where the reference tree is absent (the GPU box)."""
import numpy as np
import pytest

import aim_b200 as A
from conftest import md5_bytes, oracle_results_to_aim, render_output
from oracle import oracle as O
from oracle import refbuild as rb

pytestmark = pytest.mark.skipif(not rb.reference_available(), reason="/root/reference not present")

CASES = [
    # alg, mem, knobs, (seed, n, length, error)
    ("wfa", "mram", dict(max_score=30, read_size=168, backtrace=True, reduce=True), (101, 3000, 150, 0.04)),
    ("wfa", "wram", dict(max_score=30, read_size=168, backtrace=True, reduce=True), (102, 3000, 150, 0.04)),
    ("wfa", "mram", dict(max_score=12, read_size=168, backtrace=True, reduce=False), (103, 2000, 150, 0.04)),
    ("wfa", "mram", dict(max_score=60, read_size=120, mismatch=5, gap_o=2, gap_e=3, backtrace=True, reduce=True), (104, 2000, 100, 0.06)),
    ("wfa", "mram", dict(max_score=400, read_size=2104, backtrace=False, reduce=True), (105, 40, 2000, 0.05)),
    ("nw", "mram", dict(max_score=8, read_size=112, mismatch=3, gap_o=4, backtrace=True), (106, 600, 100, 0.04)),
    ("nw", "wram", dict(max_score=8, read_size=64, mismatch=2, gap_o=1, backtrace=True), (107, 1500, 50, 0.10)),
    ("swg", "mram", dict(max_score=80, read_size=272, mismatch=4, gap_o=6, gap_e=2, backtrace=True), (108, 150, 250, 0.04)),
    ("swg", "mram", dict(max_score=20, read_size=64, mismatch=4, gap_o=6, gap_e=2, backtrace=True), (109, 1500, 50, 0.10)),
]


@pytest.mark.parametrize("alg,mem,kw,gen", CASES, ids=[f"{c[0]}-{c[1]}-{c[3][0]}" for c in CASES])
def test_port_equals_reference(alg, mem, kw, gen, tmp_path):
    seed, n, length, err = gen
    if alg == "wfa" and mem == "wram" and kw["backtrace"]:
        # WFA/DPU-WRAM's give-up path is undefined behaviour (wfa.c WRAM:345,368-375): keep scores <= MAX_SCORE
        kw = dict(kw, max_score=60)
    binary = rb.build_ref(alg, mem, **kw)
    plen, tlen, pats, txts = A.generate_pairs(seed, n, length, err, kw["read_size"])
    pairs = tmp_path / "in.pairs"
    A.write_pairs(pairs, plen, tlen, pats, txts)
    ref_out = tmp_path / "ref.out"
    rb.run_ref(binary, pairs, ref_out, n + 7)
    res, ops = O.align(alg, plen, tlen, pats, txts, max_score=kw["max_score"], read_size=kw["read_size"],
                       match=kw.get("match", 0), mismatch=kw.get("mismatch", 3), gap_open=kw.get("gap_o", 4),
                       gap_ext=kw.get("gap_e", 1), backtrace=kw["backtrace"], reduce=kw.get("reduce", False))
    out = render_output(oracle_results_to_aim(res), ops, kw["read_size"], kw["backtrace"], tmp_path)
    assert md5_bytes(out) == rb.md5(ref_out)
    assert int((res["status"] != 0).sum()) == 0


def test_reference_is_thread_count_invariant(tmp_path):
    """The multithreaded CPU baseline (one host thread per simulated DPU) prints the same bytes."""
    kw = dict(max_score=30, read_size=168, backtrace=True, reduce=True)
    binary = rb.build_ref("wfa", "mram", **kw)
    plen, tlen, pats, txts = A.generate_pairs(7, 4096, 150, 0.04, 168)
    pairs = tmp_path / "in.pairs"
    A.write_pairs(pairs, plen, tlen, pats, txts)
    a, b = tmp_path / "a.out", tmp_path / "b.out"
    rb.run_ref(binary, pairs, a, 4096)
    rb.run_ref(binary, pairs, b, 4096, nr_dpus=8, threads=4)
    assert rb.md5(a) == rb.md5(b)
