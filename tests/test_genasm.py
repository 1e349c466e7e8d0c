"""GenASM-DC / GenASM-filter (aim-genasm submodule).

CPU (`-m "not gpu"`): the oracle restatement reproduces the reference's output lines on every golden case
(tests/golden/make_golden_genasm.py), except on the pairs it flags as not being a function of the pair.
GPU (`-m gpu`): the CUDA path, through the C ABI, is bit-exact against the oracle (score, status, CIGAR string,
max_operations) and reproduces the reference's output lines.
"""
import json
import lzma

import numpy as np
import pytest

from conftest import GOLDEN
import aim_b200 as A
from oracle import oracle as O

GMAN = {e["name"]: e for e in json.loads((GOLDEN / "genasm" / "manifest.json").read_text())}


def load_case(name, tmp_path):
    e = GMAN[name]
    f = tmp_path / (name + ".pairs")
    f.write_bytes(lzma.open(GOLDEN / e["input"]).read())
    rs = e["params"]["read_size"]
    want = A.pairs_to_process(A.count_pairs(f), e["n_arg"], 1)
    arrays = A.read_pairs(f, rs, want)
    ref_lines = lzma.open(GOLDEN / e["output"]).read().decode("latin-1").split("\n")[:-1]
    return e, arrays, ref_lines


def render(algo, res, ops, i):
    """One output line as aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:286-296 (DC) / DPU-WRAM-filter/host/host.c:272 prints it."""
    if algo == "genasm_filter":
        return f"{i}, {res['score'][i]}"
    return f"{i}, {res['score'][i]}, {bytes(ops[i, :res['end_offset'][i]]).decode('latin-1')}"


def check_lines(algo, res, ops, ref_lines, flagged):
    assert len(ref_lines) == len(res)
    bad = [i for i in range(len(res)) if not flagged[i] and render(algo, res, ops, i) != ref_lines[i]]
    assert not bad, f"{len(bad)} lines differ, first {bad[0]}: {render(algo, res, ops, bad[0])!r} vs {ref_lines[bad[0]]!r}"


@pytest.mark.parametrize("name", sorted(GMAN))
def test_oracle_matches_reference_lines(name, tmp_path):
    e, (plen, tlen, pats, txts), ref_lines = load_case(name, tmp_path)
    p = e["params"]
    res, ops = O.align(e["algo"], plen, tlen, pats, txts, max_score=p["max_score"], read_size=p["read_size"],
                       variant=int(e["variant"] == "mram" and e["algo"] == "genasm_dc"), nthreads=4)
    flagged = res["status"] != 0
    # the flagged share stays small: the check is not vacuous
    assert flagged.sum() <= 0.02 * len(res) + 2, f"{flagged.sum()} of {len(res)} pairs flagged"
    check_lines(e["algo"], res, ops, ref_lines, flagged)


def test_genasm_variant_difference():
    """DPU-MRAM-DC prints substitutions as 'S' and has no 'N' wildcard; the WRAM sources print 'X'."""
    rs = 72
    plen, tlen, pats, txts = A.generate_pairs(6, 200, 60, 0.05, rs, nthreads=1)
    r0, o0 = O.align("genasm_dc", plen, tlen, pats, txts, max_score=6, read_size=rs, variant=0)
    r1, o1 = O.align("genasm_dc", plen, tlen, pats, txts, max_score=6, read_size=rs, variant=1)
    assert np.array_equal(r0["score"], r1["score"])
    s0 = b"".join(bytes(o0[i, :r0["end_offset"][i]]) for i in range(200))
    s1 = b"".join(bytes(o1[i, :r1["end_offset"][i]]) for i in range(200))
    assert b"X" in s0 and b"S" not in s0 and s0.replace(b"X", b"S") == s1


# ---------------------------------------------------------------------------- GPU
def gpu_align(algo, arrays, k, rs, variant=0, **kw):
    plen, tlen, pats, txts = arrays
    params = A.AlignParams(algo=algo, max_score=k, read_size=rs, variant=variant, **kw)
    res, ops, _ = A.align_batch(params, plen, tlen, pats, txts)
    return res, ops


def assert_same(algo, got, gops, exp, eops, what):
    for f in ("score", "status", "max_operations", "begin_offset", "end_offset"):
        bad = np.nonzero(got[f] != exp[f])[0]
        assert bad.size == 0, f"{what}: {f} differs at pairs {bad[:8]} got {got[f][bad[:8]]} want {exp[f][bad[:8]]}"
    if algo == "genasm_dc":
        for i in range(len(got)):
            e = int(exp["end_offset"][i])
            assert bytes(gops[i, :e + 1]) == bytes(eops[i, :e]) + b"\0", f"{what}: CIGAR of pair {i}: {bytes(gops[i, :e + 1])!r} vs {bytes(eops[i, :e])!r}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GMAN))
def test_gpu_matches_oracle_and_reference(name, tmp_path):
    e, arrays, ref_lines = load_case(name, tmp_path)
    p = e["params"]
    variant = int(e["variant"] == "mram" and e["algo"] == "genasm_dc")
    exp, eops = O.align(e["algo"], *arrays, max_score=p["max_score"], read_size=p["read_size"], variant=variant, nthreads=8)
    got, gops = gpu_align(e["algo"], arrays, p["max_score"], p["read_size"], variant)
    assert_same(e["algo"], got, gops, exp, eops, name)
    check_lines(e["algo"], got, gops, ref_lines, got["status"] != 0)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["genasm_dc", "genasm_filter"])
@pytest.mark.parametrize("length,error,k", [(100, 0.02, 5), (100, 0.05, 12), (150, 0.04, 30), (200, 0.06, 40), (64, 0.05, 4),
                                            (128, 0.03, 8), (300, 0.04, 70), (40, 0.1, 2)])
def test_gpu_ragged_sweep(algo, length, error, k):
    """Ragged lengths (including empty sequences and m % 64 == 0), every kernel instantiation (word counts 2/4/8, 1/2/4 levels per
    lane, 4..32 lanes per pair), penalties other than the defaults."""
    import math
    rs = math.ceil((length * (1 + error) + 7) / 8) * 8
    n = 3000
    plen, tlen, pats, txts = A.generate_pairs(100 + length, n, length, error, rs, nthreads=4)
    rng = np.random.default_rng(length)
    plen, tlen = plen.copy(), tlen.copy()
    cut = rng.integers(0, n, 200)
    plen[cut[:100]] = rng.integers(0, np.maximum(plen[cut[:100]], 1))
    tlen[cut[100:]] = rng.integers(0, np.maximum(tlen[cut[100:]], 1))
    plen[:4] = [0, 0, min(64, rs), min(64, rs)]
    tlen[:4] = [0, 5, min(64, rs), 0]
    kw = dict(mismatch=4, gap_open=6, gap_ext=2)
    exp, eops = O.align(algo, plen, tlen, pats, txts, max_score=k, read_size=rs, nthreads=8, **kw)
    got, gops = gpu_align(algo, (plen, tlen, pats, txts), k, rs, **kw)
    assert_same(algo, got, gops, exp, eops, f"{algo} l={length} k={k}")


@pytest.mark.gpu
def test_gpu_genasm_multichunk_and_device_resident():
    """200 K pairs cross several pipeline chunks; idx follows idx_base."""
    rs, k = 112, 5
    n = 200_000
    arrays = A.generate_pairs(77, n, 100, 0.01, rs)
    got, gops = gpu_align("genasm_dc", arrays, k, rs)
    sample = np.arange(0, n, 37)
    exp, eops = O.align("genasm_dc", arrays[0][sample], arrays[1][sample], arrays[2][sample], arrays[3][sample], max_score=k, read_size=rs, nthreads=8)
    assert np.array_equal(got["idx"], np.arange(n, dtype=np.uint32))
    assert_same("genasm_dc", got[sample], gops[sample], exp, eops, "multichunk")
    # every defined CIGAR consumes the whole pattern: sum of M/X/I run lengths == plen
    ok = np.nonzero(got["status"] == 0)[0][:2000]
    for i in ok:
        s = bytes(gops[i, :got["end_offset"][i]]).decode()
        tot, num = 0, ""
        for ch in s:
            if ch.isdigit():
                num += ch
            else:
                cnt = int(num[::-1]); num = ""
                if ch in "MXI":
                    tot += cnt
        assert tot == arrays[0][i]
