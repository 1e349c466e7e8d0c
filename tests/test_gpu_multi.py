"""GPU, needs >= 2 devices (skipped otherwise): in-process sharding (aim_params.ngpus) of aim_align_batch (chunks pulled from one
queue by one host thread + stream set per GPU) and aim_align_packed (contiguous index ranges), results in pair order
(host.c:201-209 per DPU)."""
import numpy as np
import pytest

import aim_b200 as A

pytestmark = pytest.mark.gpu


def _need(n):
    if A.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("algo,kw", [("wfa", dict(max_score=30, read_size=168, backtrace=True, reduce=True)),
                                     ("nw", dict(max_score=4, read_size=168, backtrace=True)),
                                     ("genasm_dc", dict(max_score=6, read_size=168))])
def test_sharded_batch_equals_single_gpu(algo, kw):
    _need(2)
    rs = kw["read_size"]
    n = 300_001  # not a multiple of anything
    arrays = A.generate_pairs(90, n, 150, 0.04 if algo != "genasm_dc" else 0.01, rs)
    one, ops1, _ = A.align_batch(A.AlignParams(algo=algo, **kw), *arrays, idx_base=7)
    for g in (2, min(A.device_count(), 4)):
        many, opsg, _ = A.align_batch(A.AlignParams(algo=algo, ngpus=g, **kw), *arrays, idx_base=7)
        assert np.array_equal(one, many), f"{algo}: results differ with ngpus={g}"
        for i in range(0, n, 997):
            b, e = (0, int(one["end_offset"][i]) + 1) if algo == "genasm_dc" else (int(one["begin_offset"][i]), int(one["end_offset"][i]))
            assert np.array_equal(ops1[i, b:e], opsg[i, b:e]), (algo, g, i)


def test_sharded_packed_equals_single_gpu():
    _need(2)
    rs = 168
    n = 250_037
    plen, tlen, pats, txts = A.generate_pairs(91, n, 150, 0.04, rs)
    pats = pats.copy()
    pats[[5, 125_020, 250_000], 3] = ord("N")
    packed, flags = A.pack_pairs(plen, tlen, pats, txts, rs)
    p1 = A.AlignParams(algo="wfa", max_score=30, read_size=rs, backtrace=True, reduce=True)
    r1, c1, _ = A.align_packed(p1, plen, tlen, packed, flags)
    for g in (2, min(A.device_count(), 3)):
        pg = A.AlignParams(algo="wfa", max_score=30, read_size=rs, backtrace=True, reduce=True, ngpus=g)
        rg, cg, _ = A.align_packed(pg, plen, tlen, packed, flags)
        assert np.array_equal(r1, rg)
        assert np.array_equal(r1["status"][[5, 125_020, 250_000]], [5, 5, 5])
        ok = r1["status"] == 0
        rows1 = [bytes(c1[i]).split(b"\0", 1)[0] for i in np.nonzero(ok)[0][::211]]
        rowsg = [bytes(cg[i]).split(b"\0", 1)[0] for i in np.nonzero(ok)[0][::211]]
        assert rows1 == rowsg


def test_chunk_queue_long_reads_equals_single_gpu():
    """Variable-cost pairs (long reads, adaptive) through the chunk queue: same bytes as one GPU, whatever GPU took which chunk."""
    _need(2)
    ms, rs = A.derive_knobs("wfa", 3000, 0.10)
    n = 20_003
    arrays = A.generate_pairs(92, n, 3000, 0.10, rs)
    kw = dict(algo="wfa", max_score=ms, read_size=rs, backtrace=False, reduce=True)
    one, _, _ = A.align_batch(A.AlignParams(**kw), *arrays, idx_base=11)
    for g in (2, A.device_count()):
        many, _, _ = A.align_batch(A.AlignParams(ngpus=g, **kw), *arrays, idx_base=11)
        assert np.array_equal(one, many), f"results differ with ngpus={g}"
