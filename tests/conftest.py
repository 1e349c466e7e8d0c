"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200 box.

oracle/ is used in tests only as the checker.  /root/reference is never read by gpu tests.
"""
from __future__ import annotations

import hashlib
import json
import lzma
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: long-running")


def _ensure_built():
    lib = ROOT / "aim_b200" / "libaim_b200.so"
    if not lib.exists():
        import __graft_entry__ as ge
        ge.build()


_ensure_built()
MANIFEST = {e["name"]: e for e in json.loads((GOLDEN / "manifest.json").read_text())}


def md5_bytes(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def params_from_entry(entry):
    """manifest params (reference -D names) -> aim_b200.AlignParams kwargs."""
    p = entry["params"]
    return dict(algo=entry["algo"], match=p.get("match", 0), mismatch=p.get("mismatch", 3), gap_open=p.get("gap_o", 4),
                gap_ext=p.get("gap_e", 1), max_score=p["max_score"], read_size=p["read_size"],
                backtrace=bool(p.get("backtrace", False)), reduce=bool(p.get("reduce", False)))


def oracle_kwargs(kw):
    return {k: v for k, v in kw.items() if k != "algo"}


@pytest.fixture(scope="session")
def golden_case(tmp_path_factory):
    """name -> (entry, AlignParams kwargs, (plen, tlen, patterns, texts) as the host would align them)."""
    import aim_b200 as A
    cache = {}
    tmp = tmp_path_factory.mktemp("golden")

    def load(name):
        if name in cache:
            return cache[name]
        e = MANIFEST[name]
        kw = params_from_entry(e)
        f = tmp / (name + ".pairs")
        f.write_bytes(lzma.open(GOLDEN / e["input"]).read())
        total = A.count_pairs(f)
        want = A.pairs_to_process(total, e["n_arg"], 1)
        arrays = A.read_pairs(f, kw["read_size"], want)
        cache[name] = (e, kw, arrays)
        return cache[name]

    return load


def render_output(results, ops, read_size, backtrace, tmp_path) -> bytes:
    """Bytes the reference's printer would write (host.c:332-353) for these results."""
    import aim_b200 as A
    out = Path(tmp_path) / f"render-{os.getpid()}-{id(results)}.out"
    A.write_results(out, results, ops, read_size, backtrace)
    data = out.read_bytes()
    out.unlink()
    return data


def oracle_results_to_aim(res):
    import aim_b200 as A
    r = np.zeros(len(res), A.RESULT_DTYPE)
    for f in ("max_operations", "begin_offset", "end_offset", "score", "status"):
        r[f] = res[f]
    r["idx"] = np.arange(len(res), dtype=np.uint32)
    return r


def assert_same_alignment(got, got_ops, exp, exp_ops, backtrace, what=""):
    """Bit-exact: score, status, CIGAR span and the op bytes inside the span."""
    for f in ("score", "max_operations", "end_offset", "status"):
        bad = np.nonzero(got[f] != exp[f])[0]
        assert bad.size == 0, f"{what}: {f} differs at pairs {bad[:8]} got {got[f][bad[:8]]} want {exp[f][bad[:8]]}"
    if not backtrace:
        return
    bad = np.nonzero(got["begin_offset"] != exp["begin_offset"])[0]
    assert bad.size == 0, f"{what}: begin_offset differs at pairs {bad[:8]}"
    for i in range(len(got)):
        b, e = int(exp["begin_offset"][i]), int(exp["end_offset"][i])
        if not np.array_equal(got_ops[i, b:e], exp_ops[i, b:e]):
            raise AssertionError(f"{what}: ops differ at pair {i}: got {bytes(got_ops[i, b:e])!r} want {bytes(exp_ops[i, b:e])!r}")
