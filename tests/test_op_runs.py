"""aim_align_batch brings the op rows back as RUN rows (include/aim_b200.h, "how aim_align_batch brings the op rows back"; the
rows the reference pulls with dpu_push_xfer FROM_DPU, WFA/DPU-MRAM/host/host.c:316-326).
CPU: the host half (aim_expand_op_runs) on hand-made run rows.  GPU: the caller's buffers are byte-identical with and without
the run rows, on inputs where no row / a few rows / most rows overflow their run row, over many chunks, pinned and pageable."""
import os

import numpy as np
import pytest

import aim_b200 as A


def _encode(row: bytes, pitch: int) -> bytes:
    """Python restatement of op_runs_kernel's format: 32-bit words - the run count, then position | length (1..255) << 16 | op << 24 per
    run of bytes other than 'M'; first word 0xffffffff = more runs than the row holds."""
    words = []
    i = 0
    while i < len(row):
        j = i
        while j < len(row) and row[j] == row[i]:
            j += 1
        if row[i] != ord("M"):
            s = i
            while s < j:
                words.append(s | (min(j - s, 255) << 16) | (row[i] << 24))
                s += 255
        i = j
    cap = pitch // 4 - 1
    if len(words) > cap:
        return np.array([0xFFFFFFFF] + [0] * cap, np.uint32).tobytes()
    return np.array([len(words)] + words + [0] * (cap - len(words)), np.uint32).tobytes()


def test_pitch_rule():
    assert A.op_runs_pitch(168) == 64 and A.op_runs_pitch(112) == 48 and A.op_runs_pitch(256) == 96
    assert A.op_runs_pitch(32) == 32 and A.op_runs_pitch(24) == 0 and A.op_runs_pitch(1024) == 384 and A.op_runs_pitch(1032) == 0


def test_download_bytes_rule():
    wfa = dict(algo="wfa", mismatch=3, gap_open=4, gap_ext=1, read_size=168, reduce=True)
    assert A.op_rows_download_bytes(A.AlignParams(max_score=30, backtrace=True, **wfa)) == 44   # 1 + 30 / 3 words
    assert A.op_rows_download_bytes(A.AlignParams(max_score=400, backtrace=True, **wfa)) == 64  # the READ_SIZE rule is the smaller one
    assert A.op_rows_download_bytes(A.AlignParams(max_score=30, backtrace=False, **wfa)) == 0   # no op rows at all
    assert A.op_rows_download_bytes(A.AlignParams(algo="nw", max_score=4, read_size=112, backtrace=True)) == 48
    assert A.op_rows_download_bytes(A.AlignParams(algo="genasm_dc", max_score=5, read_size=120)) == 64
    assert A.op_rows_download_bytes(A.AlignParams(algo="genasm_dc", max_score=30, read_size=168)) == 0
    assert A.op_rows_download_bytes(A.AlignParams(algo="genasm_filter", max_score=2, read_size=120)) == 0


@pytest.mark.parametrize("rs,n", [(168, 5000), (32, 300), (1024, 100), (112, 1)])
def test_expand_restores_rows_and_lists_overflows(rs, n):
    rng = np.random.default_rng(rs)
    pitch = A.op_runs_pitch(rs)
    rows = np.full((n, 2 * rs), ord("M"), np.uint8)
    for i in range(n):
        kind = i % 5
        k = 0 if kind == 0 else int(rng.integers(1, 8)) if kind < 4 else 2 * rs  # all 'M' / a few edits / noise (overflows)
        pos = rng.integers(0, 2 * rs, size=k)
        rows[i, pos] = rng.choice(np.frombuffer(b"XID", np.uint8), size=k)
    if n > 3:
        rows[3, :] = ord("D")        # one run longer than 255 when 2*rs > 255
        rows[2, 0] = ord("X")        # a run of one at each end
        rows[2, -1] = ord("I")
    runs = np.frombuffer(b"".join(_encode(bytes(r), pitch) for r in rows), np.uint8).reshape(n, pitch)
    want_ov = np.flatnonzero(runs.view(np.uint32)[:, 0] == 0xFFFFFFFF)
    got = np.full((n, 2 * rs), 0x55, np.uint8)
    got, ov = A.expand_op_runs(runs, rs, got)
    assert list(ov) == list(want_ov)
    keep = np.ones(n, bool)
    keep[want_ov] = False
    assert (got[keep] == rows[keep]).all()
    assert (got[~keep] == 0x55).all()  # overflowing rows are left to the caller (aim_align_batch fetches them as they are)


def test_expand_rejects_bad_arguments():
    with pytest.raises(A.AimError):
        A.expand_op_runs(np.zeros((4, 14), np.uint8), 168)     # pitch not a multiple of 4
    with pytest.raises(A.AimError):
        A.expand_op_runs(np.zeros((4, 64), np.uint8), 2048)    # rows this wide are never run rows


def test_expand_treats_a_run_outside_the_row_as_overflow():
    rs, pitch = 64, 32
    runs = np.zeros((3, pitch // 4), np.uint32)
    runs[0, :2] = [1, 100 | (28 << 16) | (ord("X") << 24)]   # ops 100..127
    runs[1, :2] = [1, 120 | (28 << 16) | (ord("X") << 24)]   # would end at 148 > 128
    runs[2, 0] = 9                                           # more runs than a 32-byte row holds
    ops, ov = A.expand_op_runs(runs.view(np.uint8).reshape(3, pitch), rs)
    assert list(ov) == [1, 2] and bytes(ops[0]) == b"M" * 100 + b"X" * 28


def _mixed_pairs(n, length, rs, noisy_every, seed):
    """generate_dataset pairs; every `noisy_every`-th text replaced by random bases (its alignment is mostly X / gaps)."""
    plen, tlen, pats, txts = A.generate_pairs(seed, n, length, 0.04, rs)
    if noisy_every:
        rng = np.random.default_rng(seed)
        idx = np.arange(0, n, noisy_every)
        txts[idx, :length] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(len(idx), length))
        tlen[idx] = length
    return plen, tlen, pats, txts


@pytest.mark.gpu
@pytest.mark.parametrize("algo,length,rs,max_score,noisy_every,chunk_mb", [
    ("wfa", 150, 168, 400, 0, None),       # no overflow
    ("wfa", 150, 168, 30, 0, None),        # config 4: the run row is cut down to what a score of 30 can need (44 bytes)
    ("wfa", 150, 168, 30, 3000, 1),        # ... and unrelated pairs give up with untouched rows
    ("wfa", 150, 168, 400, 3000, 1),       # a few overflowing rows per chunk, many chunks: fetched one by one
    ("wfa", 150, 168, 400, 2, 1),          # half the rows overflow: the chunk's rows are fetched as they are
    ("nw", 100, 112, 40, 97, None),
    ("swg", 250, 256, 400, 50, 1),
    ("nw", 24, 32, 40, 5, None),           # smallest served READ_SIZE
])
def test_batch_rows_identical_with_and_without_run_rows(algo, length, rs, max_score, noisy_every, chunk_mb, monkeypatch):
    n = 40_000 if algo != "swg" else 12_000
    arrays = _mixed_pairs(n, length, rs, noisy_every, 31)
    params = A.AlignParams(algo=algo, mismatch=3, gap_open=4, gap_ext=1, max_score=max_score, read_size=rs, backtrace=True, reduce=False)
    if chunk_mb:
        monkeypatch.setenv("AIM_CHUNK_MB", str(chunk_mb))
    monkeypatch.setenv("AIM_SPARSE_OPS", "0")
    res0, ops0, _ = A.align_batch(params, *arrays)
    monkeypatch.delenv("AIM_SPARSE_OPS")
    res1, ops1, _ = A.align_batch(params, *arrays)                      # pageable buffers
    assert res0.tobytes() == res1.tobytes()
    assert (ops0 == ops1).all()
    pr, po = A.PinnedArray((n,), A.RESULT_DTYPE), A.PinnedArray((n, 2 * rs), np.uint8)
    po.array[:] = 0
    res2, ops2, _ = A.align_batch(params, *arrays, results=pr.array, ops=po.array)  # pinned output buffers
    assert res0.tobytes() == res2.tobytes() and (ops0 == ops2).all()
    if noisy_every and max_score > 100:
        pitch = A.op_runs_pitch(rs)
        spans_long = sum(_encode(bytes(ops0[i]), pitch)[:4] == b"\xff\xff\xff\xff" for i in range(0, n, noisy_every))
        assert spans_long > 0, "the case was meant to overflow some run rows"


@pytest.mark.gpu
def test_two_gpus_share_the_host_pool():
    if A.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    arrays = _mixed_pairs(200_000, 150, 168, 1000, 5)
    params = A.AlignParams(algo="wfa", mismatch=3, gap_open=4, gap_ext=1, max_score=400, read_size=168, backtrace=True, reduce=True)
    res1, ops1, _ = A.align_batch(params, *arrays)
    params2 = A.AlignParams(**{**params.__dict__, "ngpus": 2})
    res2, ops2, _ = A.align_batch(params2, *arrays)
    assert res1.tobytes() == res2.tobytes() and (ops1 == ops2).all()


@pytest.mark.gpu
@pytest.mark.parametrize("k,rs,chunk_mb", [(5, 120, None), (5, 120, 1), (11, 128, 1)])
def test_genasm_dc_strings_identical_with_and_without_string_heads(k, rs, chunk_mb, monkeypatch):
    """GenASM-DC: the op rows hold the DPU's CIGAR strings; only the head of every row (room for the longest string k error levels can
    make) crosses PCIe and the host copies each string with its NUL into the caller's row.  k = 11 at READ_SIZE 128 leaves no room
    (the heads would be as long as the rows): that case moves the rows as they are."""
    n = 60_000
    arrays = A.generate_pairs(9, n, 100, 0.03, rs)
    params = A.AlignParams(algo="genasm_dc", max_score=k, read_size=rs)
    if chunk_mb:
        monkeypatch.setenv("AIM_CHUNK_MB", str(chunk_mb))
    monkeypatch.setenv("AIM_SPARSE_OPS", "0")
    res0, ops0, _ = A.align_batch(params, *arrays)
    monkeypatch.delenv("AIM_SPARSE_OPS")
    res1, ops1, _ = A.align_batch(params, *arrays)
    assert res0.tobytes() == res1.tobytes()
    ends = res0["end_offset"]
    assert ends.max() > 0
    for i in range(0, n, 7):
        e = int(ends[i])
        assert bytes(ops0[i, :e + 1]) == bytes(ops1[i, :e + 1]) and ops1[i, e] == 0, i


def test_expand_from_several_threads_at_once():
    """One process driving several GPUs has one coordinator thread per GPU calling into the same host pool: jobs are serialised, every
    caller gets its own rows back."""
    import threading
    rs, n = 168, 20_000
    pitch = A.op_runs_pitch(rs)
    rng = np.random.default_rng(3)
    jobs = []
    for t in range(4):
        rows = np.full((n, 2 * rs), ord("M"), np.uint8)
        pos = rng.integers(0, 2 * rs, size=(n, 5))
        for j in range(5):
            rows[np.arange(n), pos[:, j]] = ord("XIDX"[(t + j) % 4])
        runs = np.frombuffer(b"".join(_encode(bytes(r), pitch) for r in rows[:2000]), np.uint8).reshape(2000, pitch)
        jobs.append((rows[:2000], runs))
    errs = []

    def work(rows, runs):
        try:
            for _ in range(10):
                got, ov = A.expand_op_runs(runs, rs)
                assert len(ov) == 0 and (got == rows).all()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=j) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
