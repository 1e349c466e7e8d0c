"""GPU: the CUDA path, called through the C ABI (aim_align_batch), is bit-exact with (a) the bytes
the reference wrote for the golden vectors and (b) the CPU oracle on seeded inputs, including
ragged/empty inputs, pattern longer than text (flat-array aliasing), non-ACGT bytes, give-up,
score-only and multi-chunk batches."""
import lzma

import numpy as np
import pytest

import aim_b200 as A
from conftest import (GOLDEN, MANIFEST, assert_same_alignment, md5_bytes, oracle_kwargs, oracle_results_to_aim,
                      render_output)
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_golden_bytes(name, golden_case, tmp_path):
    e, kw, (plen, tlen, pats, txts) = golden_case(name)
    res, ops, _ = A.align_batch(A.AlignParams(**kw), plen, tlen, pats, txts)
    assert int((res["status"] != 0).sum()) == 0
    out = render_output(res, ops, kw["read_size"], kw["backtrace"], tmp_path)
    if "output" in e:
        want = lzma.open(GOLDEN / e["output"]).read()
        if out != want:
            gl, wl = out.split(b"\n"), want.split(b"\n")
            first = next(i for i, (a, b) in enumerate(zip(gl, wl)) if a != b)
            raise AssertionError(f"{name}: first differing line {first}: got {gl[first]!r} want {wl[first]!r}")
    assert md5_bytes(out) == e["md5"]


def _ragged(seed, n, read_size, lo, hi, err=0.05, dirty=False):
    """Pairs with independent lengths in [lo, hi], including empty sequences and |plen - tlen| large."""
    rng = np.random.default_rng(seed)
    plen = rng.integers(lo, hi + 1, n).astype(np.int32)
    tlen = np.clip(plen + rng.integers(-12, 13, n), lo, hi).astype(np.int32)
    pats = np.zeros((n, read_size), np.uint8)
    txts = np.zeros((n, read_size), np.uint8)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    for i in range(n):
        p = alpha[rng.integers(0, 4, plen[i])]
        t = np.resize(p, tlen[i]).copy() if plen[i] else alpha[rng.integers(0, 4, tlen[i])]
        flips = rng.random(tlen[i]) < err
        t[flips] = alpha[rng.integers(0, 4, int(flips.sum()))]
        pats[i, :plen[i]] = p
        txts[i, :tlen[i]] = t
        if dirty and i % 4 == 0 and plen[i] and tlen[i]:
            pats[i, rng.integers(0, plen[i])] = ord("N")
            txts[i, rng.integers(0, tlen[i])] = ord("n")
    # a few hand-made corner cases
    plen[0], tlen[0] = 0, 0
    if n > 3:
        plen[1], tlen[1] = 0, min(hi, 5)
        plen[2], tlen[2] = min(hi, 7), 0
        plen[3], tlen[3] = hi, max(lo, hi // 3)  # pattern much longer than text
        pats[3, :hi] = alpha[rng.integers(0, 4, hi)]
    return plen, tlen, pats, txts


CASES = [
    ("wfa", dict(max_score=40, read_size=64, backtrace=True, reduce=True), (1, 3000, 0, 60)),
    ("wfa", dict(max_score=40, read_size=64, backtrace=True, reduce=False), (2, 3000, 0, 60)),
    ("wfa", dict(max_score=40, read_size=64, backtrace=False, reduce=True), (3, 3000, 0, 60)),
    ("wfa", dict(max_score=9, read_size=64, backtrace=True, reduce=True), (4, 3000, 0, 60)),      # many give-ups
    ("wfa", dict(max_score=60, read_size=64, mismatch=1, gap_open=1, gap_ext=1, backtrace=True, reduce=True), (5, 2000, 0, 60)),
    ("wfa", dict(max_score=90, read_size=64, mismatch=7, gap_open=2, gap_ext=5, backtrace=True, reduce=True), (6, 2000, 0, 60)),
    ("wfa", dict(max_score=120, read_size=304, backtrace=True, reduce=True), (7, 1500, 200, 300)),
    ("wfa", dict(max_score=700, read_size=1504, backtrace=True, reduce=True), (8, 64, 1200, 1500)),  # long-read arena mode
    ("wfa", dict(max_score=700, read_size=1504, backtrace=False, reduce=True), (9, 64, 1200, 1500)),
    ("wfa", dict(max_score=700, read_size=1504, backtrace=True, reduce=False), (10, 48, 1200, 1500)),
    # score-only long reads: windowed-ring kernel; without trimming the window is outgrown -> hand-over list
    ("wfa", dict(max_score=700, read_size=1504, backtrace=False, reduce=False), (20, 48, 1200, 1500)),
    ("wfa", dict(max_score=2500, read_size=3008, backtrace=False, reduce=True), (21, 300, 2000, 3000)),
    ("wfa", dict(max_score=2500, read_size=3008, mismatch=2, gap_open=3, gap_ext=2, backtrace=False, reduce=True), (22, 200, 2000, 3000)),
    ("nw", dict(max_score=0, read_size=64, mismatch=3, gap_open=4, backtrace=True), (11, 3000, 0, 60)),
    ("nw", dict(max_score=0, read_size=64, mismatch=1, gap_open=1, backtrace=True), (12, 3000, 0, 60)),
    ("nw", dict(max_score=0, read_size=64, mismatch=3, gap_open=4, backtrace=False), (13, 3000, 0, 60)),
    ("nw", dict(max_score=0, read_size=304, mismatch=3, gap_open=4, backtrace=True), (14, 600, 200, 300)),
    ("swg", dict(max_score=20, read_size=64, mismatch=4, gap_open=6, gap_ext=2, backtrace=True), (15, 3000, 0, 60)),
    ("swg", dict(max_score=200, read_size=64, mismatch=4, gap_open=6, gap_ext=2, backtrace=True), (16, 3000, 0, 60)),
    ("swg", dict(max_score=20, read_size=64, match=-1, mismatch=2, gap_open=3, gap_ext=1, backtrace=True), (17, 3000, 0, 60)),
    ("swg", dict(max_score=30, read_size=64, mismatch=4, gap_open=6, gap_ext=2, backtrace=False), (18, 3000, 0, 60)),
    ("swg", dict(max_score=80, read_size=304, mismatch=4, gap_open=6, gap_ext=2, backtrace=True), (19, 600, 200, 300)),
]


@pytest.mark.parametrize("algo,kw,gen", CASES, ids=[f"{c[0]}-{c[2][0]}" for c in CASES])
@pytest.mark.parametrize("dirty", [False, True], ids=["acgt", "dirty"])
def test_vs_oracle_ragged(algo, kw, gen, dirty):
    _vs_oracle_ragged(algo, kw, gen, dirty)


DP_CASES = [c for c in CASES if c[0] in ("nw", "swg")]


@pytest.mark.parametrize("algo,kw,gen", DP_CASES, ids=[f"{c[0]}-{c[2][0]}" for c in DP_CASES])
def test_vs_oracle_ragged_literal_dp_kernel(algo, kw, gen, monkeypatch):
    """NW/SWG are normally served by the register-strip / shared-memory-row kernels (aim_dp_fast.cu); the
    literal int16 kernel (aim_dp.cu) that backs them up for over-long reads must stay bit-exact too."""
    monkeypatch.setenv("AIM_DP_MODE", "literal")
    _vs_oracle_ragged(algo, kw, gen, True)


SCAN_CASES = DP_CASES + [
    ("swg", dict(max_score=80, read_size=272, mismatch=4, gap_open=6, gap_ext=2, backtrace=True), (24, 800, 225, 262)),   # config 3's geometry
    ("swg", dict(max_score=80, read_size=272, mismatch=4, gap_open=6, gap_ext=2, backtrace=False), (25, 400, 200, 260)),
    ("swg", dict(max_score=30, read_size=160, mismatch=1, gap_open=1, gap_ext=1, backtrace=True), (26, 800, 100, 150)),
    ("nw", dict(max_score=0, read_size=112, mismatch=3, gap_open=4, backtrace=True), (27, 2000, 80, 104)),                # config 2's
    ("nw", dict(max_score=0, read_size=528, mismatch=2, gap_open=3, backtrace=True), (28, 200, 380, 520)),
]


@pytest.mark.parametrize("mode,minb,tb,grid", [("0", "0", "fused", "0"), ("1", "0", "fused", "0"), ("2", "0", "fused", "0"), ("1", "10", "fused", "3"),
                                               ("1", "6", "kernel", "0"), ("2", "8", "fused", "2"), ("1", "0", "kernel", "5"), ("2", "0", "kernel", "0")])
@pytest.mark.parametrize("algo,kw,gen", SCAN_CASES, ids=[f"{c[0]}-{c[2][0]}" for c in SCAN_CASES])
def test_vs_oracle_ragged_scan_kernel(algo, kw, gen, mode, minb, tb, grid, monkeypatch):
    """Aliased pairs (pattern longer than text) through dp_scan_kernel - the row spread over the lanes of a sub-warp, the
    horizontal gap as a min-plus scan (aim_dp_scan.cuh) - in both block geometries and all register budgets, with the traceback
    inside the fill kernel (every warp walks its last 32 pairs, one per lane) and as a kernel of its own, with a grid of a few
    blocks (every warp then serves many rounds of 32 pairs), and with the kernel switched off (dp_row_kernel).  Lengths differ by
    up to 12, so each batch also holds pairs the scan kernel must leave to dp_row_kernel (more tail cells than a block has
    columns) and non-aliased ones (dp2_strip_kernel)."""
    monkeypatch.setenv("AIM_DP_SCAN", mode)
    monkeypatch.setenv("AIM_DP_SCAN_MINB", minb)
    monkeypatch.setenv("AIM_DP_SCAN_TB", tb)
    monkeypatch.setenv("AIM_DP_SCAN_GRID", grid)
    _vs_oracle_ragged(algo, kw, gen, True)


@pytest.mark.parametrize("g", ["8", "32"])
def test_long_read_kernel_lane_groups(g, monkeypatch):
    """The windowed-ring long-read kernel with 8 and 32 lanes per pair (default 16)."""
    monkeypatch.setenv("AIM_WFA_LONG_G", g)
    _vs_oracle_ragged("wfa", dict(max_score=2500, read_size=3008, backtrace=False, reduce=True), (23, 200, 2000, 3000), True)


@pytest.mark.parametrize("backtrace", [False, True])
def test_long_read_second_pass_window(backtrace, monkeypatch):
    """Long reads run in a 128-diagonal window first; a pair whose wavefront outgrows it (112 diagonals) is handed to a second
    pass in the 256-diagonal window, and from there (224) to the warp-per-pair kernel.  A block of 120-260 inserted bases
    forces the wavefront to span that many diagonals (trimming never cuts diagonal tlen - plen): all three routes are taken."""
    rng = np.random.default_rng(77)
    n, rs = 40, 3008
    alpha = np.frombuffer(b"ACGT", np.uint8)
    plen, tlen = np.zeros(n, np.int32), np.zeros(n, np.int32)
    pats, txts = np.zeros((n, rs), np.uint8), np.zeros((n, rs), np.uint8)
    for i in range(n):
        pl = int(rng.integers(2200, 2600))
        ins = 0 if i % 4 == 0 else int(rng.integers(120, 261))
        p = alpha[rng.integers(0, 4, pl)]
        at = int(rng.integers(200, pl - 200))
        t = np.concatenate([p[:at], alpha[rng.integers(0, 4, ins)], p[at:]])
        flips = rng.random(len(t)) < 0.03
        t[flips] = alpha[rng.integers(0, 4, int(flips.sum()))]
        plen[i], tlen[i] = pl, len(t)
        pats[i, :pl], txts[i, :len(t)] = p, t
    kw = dict(max_score=2500, read_size=rs, backtrace=backtrace, reduce=True)
    exp, exp_ops = O.align("wfa", plen, tlen, pats, txts, nthreads=8, **kw)
    got = {}
    for wc in ("128", "256"):
        monkeypatch.setenv("AIM_WFA_LONG_WC", wc)
        res, ops, _ = A.align_batch(A.AlignParams(algo="wfa", **kw), plen, tlen, pats, txts)
        assert_same_alignment(res, ops, exp, exp_ops, backtrace, what=f"long reads, first window {wc}")
        got[wc] = (res.tobytes(), None if ops is None else ops.tobytes())
    assert got["128"] == got["256"]


def test_dp_long_rows_fall_back_to_literal_kernel():
    """(2*READ_SIZE+2)*max penalty >= 32767: int16 truncation is possible, the literal kernel must serve."""
    kw = dict(max_score=50, read_size=2304, mismatch=6, gap_open=7, backtrace=True)
    plen, tlen, pats, txts = _ragged(31, 24, kw["read_size"], 1800, 2300, dirty=True)
    res, ops, _ = A.align_batch(A.AlignParams(algo="nw", **kw), plen, tlen, pats, txts)
    exp, exp_ops = O.align("nw", plen, tlen, pats, txts, nthreads=8, **kw)
    assert_same_alignment(res, ops, exp, exp_ops, True, what="nw long rows")


def _vs_oracle_ragged(algo, kw, gen, dirty):
    seed, n, lo, hi = gen
    plen, tlen, pats, txts = _ragged(seed, n, kw["read_size"], lo, hi, dirty=dirty)
    res, ops, _ = A.align_batch(A.AlignParams(algo=algo, **kw), plen, tlen, pats, txts)
    exp, exp_ops = O.align(algo, plen, tlen, pats, txts, nthreads=8, **kw)
    # SWG dead ends (reference: exit(1)) must be flagged on the same pairs
    assert_same_alignment(res, ops, exp, exp_ops, kw["backtrace"] , what=f"{algo} seed {seed}")
    assert np.array_equal(res["idx"], np.arange(n, dtype=np.uint32))


def test_multichunk_and_idx_base():
    """More pairs than one transfer chunk: order, idx and results survive the double-buffered pipeline."""
    ms, rs = A.derive_knobs("wfa", 100, 0.02)
    n = 700_000
    plen, tlen, pats, txts = A.generate_pairs(21, n, 100, 0.02, rs)
    kw = dict(max_score=ms, read_size=rs, backtrace=True, reduce=True)
    res, ops, phase = A.align_batch(A.AlignParams(algo="wfa", **kw), plen, tlen, pats, txts, idx_base=1000)
    exp, exp_ops = O.align("wfa", plen, tlen, pats, txts, nthreads=16, **kw)
    assert np.array_equal(res["idx"], np.arange(n, dtype=np.uint32) + 1000)
    assert_same_alignment(res, ops, exp, exp_ops, True, what="multichunk")
    assert all(p > 0 for p in phase)


def test_pinned_buffers_match_pageable():
    ms, rs = A.derive_knobs("wfa", 150, 0.04)
    n = 50_000
    kw = dict(max_score=ms, read_size=rs, backtrace=True, reduce=True)
    pin = [A.PinnedArray((n,), np.int32), A.PinnedArray((n,), np.int32), A.PinnedArray((n, rs), np.uint8), A.PinnedArray((n, rs), np.uint8)]
    A.generate_pairs(22, n, 150, 0.04, rs, out=tuple(p.array for p in pin))
    pres, pops = A.PinnedArray((n,), A.RESULT_DTYPE), A.PinnedArray((n, 2 * rs), np.uint8)
    res, ops, _ = A.align_batch(A.AlignParams(algo="wfa", **kw), *(p.array for p in pin), results=pres.array, ops=pops.array)
    res2, ops2, _ = A.align_batch(A.AlignParams(algo="wfa", **kw), *(p.array.copy() for p in pin))
    assert_same_alignment(res, ops, res2, ops2, True, what="pinned vs pageable")


def test_errors_are_reported():
    rs = 64
    plen = np.array([10, 70], np.int32)
    tlen = np.array([10, 10], np.int32)
    z = np.zeros((2, rs), np.uint8)
    with pytest.raises(A.AimError) as ei:
        A.align_batch(A.AlignParams(algo="wfa", max_score=10, read_size=rs), plen, tlen, z, z)
    assert ei.value.code == -2  # AIM_ERR_LENGTH (host.c:119-123)
    with pytest.raises(A.AimError) as ei:
        A.align_batch(A.AlignParams(algo="wfa", max_score=10, read_size=rs, mismatch=0), plen[:1], tlen[:1], z[:1], z[:1])
    assert ei.value.code == -1  # penalty validation of the run scripts


def test_arena_overflow_is_recovered():
    """Long reads with backtrace and a deliberately small history arena (1 MiB per pair; a 10 kbp pair needs about that much):
    pairs that outgrow it are re-aligned with a larger arena instead of ending the run (reference: "Out of memory MRAM",
    dpu_allocator_mram.c:6-10), and the result is the oracle's."""
    ms, rs = A.derive_knobs("wfa", 10000, 0.10)
    n = 24
    plen, tlen, pats, txts = A.generate_pairs(55, n, 10000, 0.10, rs)
    kw = dict(max_score=ms, read_size=rs, backtrace=True, reduce=True)
    res, ops, _ = A.align_batch(A.AlignParams(algo="wfa", arena_mb=1, **kw), plen, tlen, pats, txts)
    assert int((res["status"] != 0).sum()) == 0
    exp, exp_ops = O.align("wfa", plen, tlen, pats, txts, nthreads=8, **kw)
    assert_same_alignment(res, ops, exp, exp_ops, True, what="wfa long reads, small arena")
