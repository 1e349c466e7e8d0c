"""Compact transfers (include/aim_b200.h: aim_pack_pairs / aim_align_packed / aim_write_results_packed): same scores and the
same CIGAR text as the reference prints (edit_cigar_print, WFA/DPU-MRAM/host/host.c:69-89), with 2-bit sequences in and
run-length CIGAR rows out."""
import lzma
import os

import numpy as np
import pytest

from conftest import GOLDEN, MANIFEST, md5_bytes, oracle_results_to_aim
import aim_b200 as A
from oracle import oracle as O


def py_pack(row: bytes, length: int, words: int):
    out = []
    for w in range(words):
        v = 0
        for b in range(16):
            j = 16 * w + b
            code = ((row[j] >> 1) & 3) if j < length else 0
            v = (v << 2) | code
        out.append(v)
    return out


def test_pack_pairs_matches_restatement_and_flags_non_acgt():
    rs = 168
    plen, tlen, pats, txts = A.generate_pairs(4, 300, 150, 0.04, rs, nthreads=1)
    pats, txts = pats.copy(), txts.copy()
    pats[5, 10] = ord("N")
    txts[64, 3] = ord("a")
    pats[7, plen[7]:] = ord("N")  # bytes beyond the sequence do not count
    words = A.packed_row_bytes(rs) // 4
    assert A.packed_row_bytes(rs) == 48 and A.packed_row_bytes(272) == 80
    for threads in (1, 3):
        packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs, nthreads=threads)
        assert packed.shape == (300, 2, words)
        for i in (0, 5, 7, 64, 299):
            assert list(packed[i, 0]) == py_pack(bytes(pats[i]), plen[i], words)
            assert list(packed[i, 1]) == py_pack(bytes(txts[i]), tlen[i], words)
        set_bits = [i for i in range(300) if (flags[i >> 5] >> (i & 31)) & 1]
        assert set_bits == [5, 64]


@pytest.mark.parametrize("variant", ["avx2", "bmi2", "swar"])
@pytest.mark.parametrize("rs", [168, 40, 256])
def test_pack_pairs_variants_on_ragged_rows(variant, rs, monkeypatch):
    """Every packer variant (AVX2 32 bases per step / BMI2 pext / plain 64-bit) against the restatement: every length 0..READ_SIZE (full
    and partial 32-base steps, READ_SIZE not a multiple of 32), a byte outside ACGT at every position class (first, last counted, first
    uncounted), lower case, and rows whose bytes beyond the sequence are not bases."""
    if variant != "avx2":
        monkeypatch.setenv("AIM_NO_AVX2", "1")
    if variant == "swar":
        monkeypatch.setenv("AIM_NO_BMI2", "1")
    rng = np.random.default_rng(rs)
    n = rs + 1 + 64
    alpha = np.frombuffer(b"ACGT", np.uint8)
    plen = np.concatenate([np.arange(rs + 1), rng.integers(0, rs + 1, 64)]).astype(np.int32)
    tlen = plen[::-1].copy()
    pats = alpha[rng.integers(0, 4, (n, rs))]
    txts = alpha[rng.integers(0, 4, (n, rs))]
    want_flag = set()
    for i in range(n):
        pats[i, plen[i]:] = rng.integers(0, 256, rs - plen[i])   # bytes beyond the sequence do not count
        k = i % 7
        if k == 1 and plen[i] > 0: pats[i, 0] = ord("N"); want_flag.add(i)
        if k == 2 and plen[i] > 0: pats[i, plen[i] - 1] = ord("a"); want_flag.add(i)
        if k == 3 and tlen[i] > 33: txts[i, 33] = 0; want_flag.add(i)
        if k == 4 and tlen[i] < rs: txts[i, tlen[i]] = ord("N")  # first uncounted byte: not flagged
    words = A.packed_row_bytes(rs) // 4
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs, nthreads=2)
    for i in range(n):
        assert list(packed[i, 0]) == py_pack(bytes(pats[i]), plen[i], words) or i in want_flag, (variant, i)
        assert list(packed[i, 1]) == py_pack(bytes(txts[i]), tlen[i], words) or i in want_flag, (variant, i)
    got_flag = {i for i in range(n) if (flags[i >> 5] >> (i & 31)) & 1}
    assert got_flag == want_flag


@pytest.mark.gpu
@pytest.mark.parametrize("reduce", [True, False])
def test_packed_path_equals_oracle(reduce, monkeypatch):
    """Scores, spans and CIGAR text bit-exact vs the oracle; flagged / overflowing pairs reported; odd chunking and idx_base."""
    monkeypatch.setenv("AIM_CHUNK_PAIRS", "4096")  # several chunks
    rs, ms = 168, 30
    n = 20_000
    plen, tlen, pats, txts = A.generate_pairs(41, n, 150, 0.04, rs)
    pats, txts = pats.copy(), txts.copy()
    dirty = [3, 4097, 12_345]
    for i in dirty:
        pats[i, 7] = ord("N")
    plen, tlen = plen.copy(), tlen.copy()
    plen[10], tlen[11] = 0, 0
    plen[12] = tlen[12] = 0
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs)
    params = A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=True, reduce=reduce)
    res, cig, phase = A.align_packed(params, plen, tlen, packed, flags, cigar_pitch=64, idx_base=1000)
    exp, eops = O.align("wfa", plen, tlen, pats, txts, max_score=ms, read_size=rs, backtrace=True, reduce=reduce, nthreads=8)
    want = A.cigar_strings(oracle_results_to_aim(exp), eops)
    assert np.array_equal(res["idx"], np.arange(1000, 1000 + n, dtype=np.uint32))
    ok = np.ones(n, bool)
    ok[dirty] = False
    assert np.array_equal(res["status"][dirty], [5, 5, 5])
    assert int((res["status"][ok] != 0).sum()) == 0
    for f in ("score", "max_operations", "begin_offset", "end_offset"):
        assert np.array_equal(res[f][ok], exp[f][ok]), f
    for i in np.nonzero(ok)[0]:
        got = bytes(cig[i]).split(b"\0", 1)[0].decode()
        assert got == want[i], (i, got, want[i])
    assert all(p >= 0 for p in phase)
    # a row too small for some CIGARs: those pairs say so, the others are unchanged
    res2, cig2, _ = A.align_packed(params, plen, tlen, packed, flags, cigar_pitch=16)
    over = res2["status"] == 6
    assert over.any() and np.array_equal(res2["score"][ok], exp["score"][ok])
    for i in np.nonzero(ok & ~over)[0][:2000]:
        assert bytes(cig2[i]).split(b"\0", 1)[0].decode() == want[i]
    assert all(len(want[i]) >= 15 for i in np.nonzero(over)[0])


@pytest.mark.gpu
def test_packed_path_writes_the_reference_bytes(tmp_path):
    """Golden case cfg4 (reference output committed): pack -> aim_align_packed -> aim_write_results_packed == the reference's file."""
    e = MANIFEST["cfg4_wfa_adaptive_synth"]
    p = e["params"]
    f = tmp_path / "in.pairs"
    f.write_bytes(lzma.open(GOLDEN / e["input"]).read())
    rs = p["read_size"]
    plen, tlen, pats, txts = A.read_pairs(f, rs)
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs)
    params = A.AlignParams(algo="wfa", max_score=p["max_score"], read_size=rs, backtrace=True, reduce=True)
    res, cig, _ = A.align_packed(params, plen, tlen, packed, flags, cigar_pitch=96)
    out = tmp_path / "out"
    A.write_results_packed(out, res, cig)
    assert md5_bytes(out.read_bytes()) == e["md5"]


@pytest.mark.gpu
def test_packed_entry_rejects_what_it_does_not_serve():
    rs = 168
    plen, tlen, pats, txts = A.generate_pairs(4, 64, 150, 0.04, rs)
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs)
    with pytest.raises(A.AimError):
        A.align_packed(A.AlignParams(algo="nw", max_score=30, read_size=rs, backtrace=True), plen, tlen, packed, flags)
    with pytest.raises(A.AimError):
        A.align_packed(A.AlignParams(algo="wfa", max_score=30, read_size=rs, backtrace=False), plen, tlen, packed, flags)
    with pytest.raises(A.AimError):
        A.align_packed(A.AlignParams(algo="wfa", max_score=30, read_size=rs, backtrace=True), plen, tlen, packed, flags, cigar_pitch=20)


@pytest.mark.gpu
@pytest.mark.parametrize("algo,kw,length,error", [
    ("wfa", dict(max_score=30, read_size=168, reduce=True), 150, 0.04),
    ("nw", dict(max_score=8, read_size=112), 100, 0.02),
    ("swg", dict(max_score=40, read_size=112, mismatch=4, gap_open=6, gap_ext=2), 100, 0.04),
    ("genasm_dc", dict(max_score=5, read_size=112), 100, 0.01),
])
def test_cigar_rows_from_reference_layout_inputs(algo, kw, length, error, monkeypatch):
    """aim_align_batch_cigars: the reference's input layout, the CIGAR text it prints out - every algorithm, several chunks,
    pairs with bytes outside ACGT served like in aim_align_batch."""
    monkeypatch.setenv("AIM_CHUNK_MB", "1")
    rs = kw["read_size"]
    n = 40_000
    plen, tlen, pats, txts = A.generate_pairs(55, n, length, error, rs)
    pats = pats.copy()
    if algo != "genasm_dc":
        pats[[9, 20_001], 5] = ord("N")
    params = A.AlignParams(algo=algo, backtrace=True, **kw)
    res, cig, _ = A.align_batch_cigars(params, plen, tlen, pats, txts, cigar_pitch=96, idx_base=3)
    okw = {k: v for k, v in kw.items()}
    exp, eops = O.align(algo, plen, tlen, pats, txts, backtrace=True, nthreads=8, **okw)
    assert np.array_equal(res["idx"], np.arange(3, 3 + n, dtype=np.uint32))
    for f in ("score", "status", "max_operations", "begin_offset", "end_offset"):
        assert np.array_equal(res[f], exp[f]), f
    if algo == "genasm_dc":
        want = [bytes(eops[i, :exp["end_offset"][i]]).decode() for i in range(n)]
    else:
        want = A.cigar_strings(oracle_results_to_aim(exp), eops)
    for i in range(n):
        got = bytes(cig[i]).split(b"\0", 1)[0].decode()
        assert got == want[i], (i, got, want[i])
    # a row too small: flagged, scores intact
    res2, cig2, _ = A.align_batch_cigars(params, plen, tlen, pats, txts, cigar_pitch=16)
    over = res2["status"] == 6
    assert np.array_equal(res2["score"], exp["score"]) and all(len(want[i]) >= 16 for i in np.nonzero(over)[0])
    assert all(len(want[i]) < 16 for i in np.nonzero((res2["status"] == 0) & (exp["status"] == 0))[0])


def test_packed_writer_format_and_length_check(tmp_path):
    """aim_write_results_packed = the reference's output lines from CIGAR rows; aim_pack_pairs rejects lengths beyond READ_SIZE
    like get_reads does (host.c:119-123)."""
    res = np.zeros(3, A.RESULT_DTYPE)
    res["idx"] = [10, 11, 12]
    res["score"] = [5, 0, 31]
    cig = np.zeros((3, 16), np.uint8)
    for i, t in enumerate((b"40M1D59M", b"100M", b"1M")):
        cig[i, :len(t)] = np.frombuffer(t, np.uint8)
    out = tmp_path / "o"
    A.write_results_packed(out, res, cig)
    assert out.read_bytes() == b"10, 5, \n40M1D59M\n11, 0, \n100M\n12, 31, \n1M\n"
    rs = 64
    plen, tlen, pats, txts = A.generate_pairs(1, 40, 50, 0.04, rs, nthreads=1)
    bad = plen.copy()
    bad[33] = rs + 1
    with pytest.raises(A.AimError) as ei:
        A.pack_pairs(bad, tlen, pats, txts, rs)
    assert ei.value.code == -2  # AIM_ERR_LENGTH
    # the last flag word covers fewer than 32 pairs
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs)
    assert flags.shape == (2,) and int(flags.sum()) == 0


@pytest.mark.parametrize("threads", ["1", "3"])
def test_packed_writer_blocks_stay_in_order(tmp_path, monkeypatch, threads):
    monkeypatch.setenv("AIM_IO_THREADS", threads)
    n = 100_003  # several 32 K blocks, the last one partial
    res = np.zeros(n, A.RESULT_DTYPE)
    res["idx"] = np.arange(n)
    res["score"] = np.arange(n) % 31
    cig = np.zeros((n, 16), np.uint8)
    rows = [(b"%dM1X%dM" % (i % 90 + 1, i % 7 + 1)) for i in range(n)]
    for i, t in enumerate(rows):
        cig[i, :len(t)] = np.frombuffer(t, np.uint8)
    out = tmp_path / "o"
    A.write_results_packed(out, res, cig)
    want = b"".join(b"%d, %d, \n%s\n" % (i, i % 31, rows[i]) for i in range(n))
    assert out.read_bytes() == want


@pytest.mark.gpu
@pytest.mark.parametrize("length,error,pitch_of", [(150, 0.04, lambda rs: 2 * rs), (50, 0.04, lambda rs: rs), (50, 0.04, lambda rs: 2 * rs)])
def test_packed_large_pitch_and_small_read_size(length, error, pitch_of):
    """cigar_pitch up to 2*read_size is accepted as the header says (the CIGAR rows have their own device buffer), and
    read_size = 64 works with cigar_pitch = 64 (the rows used to share the text buffer)."""
    ms, rs = A.derive_knobs("wfa", length, error)
    n = 5000
    plen, tlen, pats, txts = A.generate_pairs(43, n, length, error, rs)
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs)
    pitch = pitch_of(rs) // 16 * 16
    params = A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=True, reduce=True)
    res, cig, _ = A.align_packed(params, plen, tlen, packed, flags, cigar_pitch=pitch)
    exp, eops = O.align("wfa", plen, tlen, pats, txts, max_score=ms, read_size=rs, backtrace=True, reduce=True, nthreads=8)
    want = A.cigar_strings(oracle_results_to_aim(exp), eops)
    assert int((res["status"] != 0).sum()) == 0
    assert np.array_equal(res["score"], exp["score"])
    assert [bytes(r).split(b"\0", 1)[0].decode() for r in cig] == want
