"""dp_scan_kernel's algorithm on the CPU: the per-lane code of the kernel (aim_b200/csrc/aim_dp_scan.cuh, compiled by g++ as a
lane-by-lane model, tests/model/dp_scan_model.cpp) against the oracle, bit for bit: score, span and op bytes of every aliased
pair the kernel's classification would hand it (reference: SWG/DPU-MRAM/dpu/swg.c:66-217, NW/DPU-WRAM/dpu/nw.c:67-153 with
num_cols = text_len + 1 <= pattern_len).  The GPU tests (test_gpu_parity.py, test_gpu_fullsize.py) check the kernel itself."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import aim_b200.api as A
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "model" / "dp_scan_model.cpp"
LIB = ROOT / "build" / "libdp_scan_model.so"


@pytest.fixture(scope="module")
def model():
    LIB.parent.mkdir(exist_ok=True)
    hdr = ROOT / "aim_b200" / "csrc" / "aim_dp_scan.cuh"
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", f"-I{ROOT / 'aim_b200' / 'csrc'}",
                        str(SRC), "-o", str(LIB)], check=True)
    lib = C.CDLL(str(LIB))
    lib.scan_model_align.restype = C.c_int
    lib.scan_model_align.argtypes = [C.c_int] * 9 + [C.c_uint32] + [C.c_void_p] * 7 + [C.c_int]
    return lib


def run_model(lib, algo, c, g, x, o, e, ms, rs, plen, tlen, pats, txts, extra=0):
    n = len(plen)
    res = np.zeros((n, 5), np.int32)
    ops = np.zeros((n, 2 * rs), np.uint8)
    served = np.zeros(n, np.uint8)
    rc = lib.scan_model_align(0 if algo == "nw" else 1, c, g, x, o, e, ms, rs, 1, n, plen.ctypes.data, tlen.ctypes.data, pats.ctypes.data,
                              txts.ctypes.data, res.ctypes.data, ops.ctypes.data, served.ctypes.data, extra)
    assert rc == 0
    return res, ops, served.astype(bool)


def compare(algo, res, ops, served, plen, tlen, pats, txts, x, o, e, ms, rs):
    ref, rops = O.align(algo, plen, tlen, pats, txts, max_score=ms, read_size=rs, mismatch=x, gap_open=o, gap_ext=e, backtrace=True, nthreads=8)
    idx = np.flatnonzero(served)
    assert len(idx) > 0
    for k, name in enumerate(("max_operations", "begin_offset", "end_offset", "score", "status")):
        bad = idx[res[idx, k] != ref[name][idx]]
        assert len(bad) == 0, f"{name}: pair {bad[0]} (plen {plen[bad[0]]}, tlen {tlen[bad[0]]}): {res[bad[0], k]} != {ref[name][bad[0]]}"
    for i in idx:
        b, en = ref["begin_offset"][i], ref["end_offset"][i]
        assert bytes(ops[i, b:en]) == bytes(rops[i, b:en]), f"ops of pair {i} (plen {plen[i]}, tlen {tlen[i]})"
    return len(idx)


def ragged(seed, n, rs, lo, hi, dmax, related=True):
    """Aliased pairs of every shape: text_len in [lo, hi], pattern_len = text_len + 1..dmax; related = the text is an edited copy."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    plen, tlen = np.zeros(n, np.int32), np.zeros(n, np.int32)
    pats, txts = np.zeros((n, rs), np.uint8), np.zeros((n, rs), np.uint8)
    for i in range(n):
        tl = int(rng.integers(lo, hi + 1))
        pl = min(rs, tl + int(rng.integers(1, dmax + 1)))
        p = acgt[rng.integers(0, 4, pl)]
        if related:
            t = p.copy()
            for _ in range(int(rng.integers(0, 1 + pl // 8))):
                t[rng.integers(0, pl)] = acgt[rng.integers(0, 4)]
            cut = rng.choice(pl, pl - tl, replace=False)
            t = np.delete(t, cut)
        else:
            t = acgt[rng.integers(0, 4, tl)]
        plen[i], tlen[i] = pl, tl
        pats[i, :pl], txts[i, :tl] = p, t
    return plen, tlen, pats, txts


@pytest.mark.parametrize("c,g", [(16, 8), (8, 16)])
def test_model_config3_swg(model, c, g):
    x, o, e, ms, rs = 4, 6, 2, 80, 272
    plen, tlen, pats, txts = A.generate_pairs(3, 1500, 250, 0.04, rs)
    res, ops, served = run_model(model, "swg", c, g, x, o, e, ms, rs, plen, tlen, pats, txts)
    assert served.sum() > 500
    compare("swg", res, ops, served, plen, tlen, pats, txts, x, o, e, ms, rs)


@pytest.mark.parametrize("c,g", [(8, 8), (4, 16)])
def test_model_config2_nw(model, c, g):
    x, o, e, ms, rs = 3, 4, 1, 4, 112
    plen, tlen, pats, txts = A.generate_pairs(2, 3000, 100, 0.01, rs)
    res, ops, served = run_model(model, "nw", c, g, x, o, e, ms, rs, plen, tlen, pats, txts)
    assert served.sum() > 500
    compare("nw", res, ops, served, plen, tlen, pats, txts, x, o, e, ms, rs)


@pytest.mark.parametrize("algo,c,g,rs,related,extra", [
    ("swg", 16, 8, 272, True, 0), ("swg", 16, 8, 272, False, 3), ("swg", 8, 16, 272, True, 2), ("swg", 8, 8, 112, False, 0),
    ("swg", 16, 16, 400, True, 0), ("swg", 4, 16, 128, True, 1), ("nw", 16, 8, 272, True, 0), ("nw", 8, 8, 112, False, 2),
    ("nw", 16, 16, 400, False, 0), ("nw", 8, 16, 200, True, 0)])
def test_model_ragged(model, algo, c, g, rs, related, extra):
    x, o, e, ms = (4, 6, 2, 80) if algo == "swg" else (3, 4, 1, 10)
    hi = min(2 * c * g, rs - 1)
    plen, tlen, pats, txts = ragged(11 * c + g + rs, 300, rs, 1, hi, c + 2, related)
    res, ops, served = run_model(model, algo, c, g, x, o, e, ms, rs, plen, tlen, pats, txts, extra)
    assert served.sum() > 150
    compare(algo, res, ops, served, plen, tlen, pats, txts, x, o, e, ms, rs)


def test_model_other_penalties(model):
    for (x, o, e, ms) in ((1, 0, 1, 30), (5, 3, 3, 7), (2, 10, 1, 200)):
        plen, tlen, pats, txts = ragged(x * 100 + o, 200, 160, 1, 128, 8, True)
        res, ops, served = run_model(model, "swg", 8, 8, x, o, e, ms, 160, plen, tlen, pats, txts)
        compare("swg", res, ops, served, plen, tlen, pats, txts, x, o, e, ms, 160)


def test_model_random_sweep(model):
    """Many small random cases over penalties, geometries and length ranges (the corners a fixed case list misses: text_len of a few
    bases, text_len at the last column a geometry holds, d = C, equal / unrelated sequences, non-ACGT bytes)."""
    rng = np.random.default_rng(2024)
    geos = [(4, 16), (8, 8), (8, 16), (16, 8), (16, 16)]
    checked = 0
    for it in range(40):
        c, g = geos[it % len(geos)]
        algo = "swg" if it % 3 else "nw"
        cols = 2 * c * g
        rs = int(min(528, (cols + 8 * int(rng.integers(0, 5)) + 7) // 8 * 8))
        x = int(rng.integers(1, 7))
        o = int(rng.integers(1, 9))
        e = int(rng.integers(1, 5))
        ms = int(rng.integers(1, 120))
        lo = int(rng.choice([1, 2, max(1, cols - 20)]))
        hi = min(cols, rs - 1)
        plen, tlen, pats, txts = ragged(1000 + it, 60, rs, lo, hi, c + 1, related=bool(it % 2))
        if it % 5 == 0:  # bytes outside ACGT are compared as bytes (swg.c:206, nw.c:143)
            for i in range(0, 60, 7):
                pats[i, rng.integers(0, plen[i])] = ord("N")
                txts[i, rng.integers(0, tlen[i])] = ord("n")
        if it % 7 == 0:  # a few identical-prefix pairs: text = pattern cut at the end
            for i in range(1, 60, 9):
                txts[i, :tlen[i]] = pats[i, :tlen[i]]
        res, ops, served = run_model(model, algo, c, g, x, o, e, ms, rs, plen, tlen, pats, txts, extra=it % 3)
        if served.sum() == 0:
            continue
        checked += compare(algo, res, ops, served, plen, tlen, pats, txts, x, o, e, ms, rs)
    assert checked > 1000
