"""Generate the committed golden vectors by running the UNMODIFIED reference (safaad/aim), built
natively by oracle/refbuild.py, on small seeded inputs.  Run in the authoring container (needs
/root/reference):   python tests/golden/make_golden.py

Every case = an input pair file (tests/golden/<name>.pairs.xz, or one of the two reference
Datasets under tests/golden/datasets/) + the reference's own output (<name>.out.xz, or its md5 for
the full datasets) + the knobs, all listed in tests/golden/manifest.json.
"""
from __future__ import annotations

import json
import lzma
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import aim_b200 as A  # noqa: E402  (host-side generator / file writer only)
from oracle import refbuild as rb  # noqa: E402

G = Path(__file__).resolve().parent
WFA = dict(alg="wfa", mem="mram", mismatch=3, gap_o=4, gap_e=1)

CASES = [
    # name, ref build kwargs, input spec
    ("cfg1_wfa_sample", dict(WFA, max_score=5, read_size=112, backtrace=True), ("dataset", "sample-l100-e1-40K", 40000)),
    ("cfg1_wfa_sample_giveup", dict(WFA, max_score=2, read_size=112, backtrace=True), ("dataset", "sample-l100-e1-40K", 40000)),
    ("cfg1_wfa_err", dict(WFA, max_score=5, read_size=112, backtrace=True), ("dataset", "ERR240727-l100-e1-30000Pairs", 30000)),
    ("cfg2_nw_err", dict(alg="nw", mem="wram", max_score=4, read_size=112, mismatch=3, gap_o=4, backtrace=True), ("dataset", "ERR240727-l100-e1-30000Pairs", 30000)),
    ("cfg2_nw_sample", dict(alg="nw", mem="wram", max_score=4, read_size=112, mismatch=3, gap_o=4, backtrace=True), ("dataset", "sample-l100-e1-40K", 40000)),
    ("swg_sample", dict(alg="swg", mem="mram", max_score=5, read_size=112, mismatch=3, gap_o=4, gap_e=1, backtrace=True), ("dataset", "sample-l100-e1-40K", 40000)),
    ("cfg3_swg_synth", dict(alg="swg", mem="mram", max_score=80, read_size=272, match=0, mismatch=4, gap_o=6, gap_e=2, backtrace=True), ("synth", 3, 400, 250, 0.04)),
    ("nw_l250_synth", dict(alg="nw", mem="wram", max_score=40, read_size=272, mismatch=3, gap_o=4, backtrace=True), ("synth", 7, 400, 250, 0.04)),
    ("swg_l250_scoreonly", dict(alg="swg", mem="mram", max_score=80, read_size=272, match=0, mismatch=4, gap_o=6, gap_e=2, backtrace=False), ("synth", 3, 400, 250, 0.04)),
    ("cfg4_wfa_adaptive_synth", dict(WFA, max_score=30, read_size=168, backtrace=True, reduce=True), ("synth", 4, 1500, 150, 0.04)),
    ("wfa_exact_l150_synth", dict(WFA, max_score=30, read_size=168, backtrace=True, reduce=False), ("synth", 4, 1500, 150, 0.04)),
    ("wfa_l150_scoreonly", dict(WFA, max_score=30, read_size=168, backtrace=False, reduce=True), ("synth", 4, 1500, 150, 0.04)),
    ("wfa_l150_giveup", dict(WFA, max_score=18, read_size=168, backtrace=True, reduce=True), ("synth", 4, 1500, 150, 0.04)),
    ("wfa_x2o3e2_l150", dict(alg="wfa", mem="mram", mismatch=2, gap_o=3, gap_e=2, max_score=30, read_size=168, backtrace=True, reduce=True), ("synth", 11, 800, 150, 0.04)),
    ("cfg5_wfa_long_scoreonly", dict(WFA, max_score=5000, read_size=11008, backtrace=False, reduce=True), ("synth", 5, 10, 10000, 0.10)),
    ("cfg5_wfa_long_bt", dict(WFA, max_score=5000, read_size=11008, backtrace=True, reduce=True, big_wram=True), ("synth", 5, 10, 10000, 0.10)),
    ("wfa_l1000_e5_bt", dict(WFA, max_score=250, read_size=1056, backtrace=True, reduce=True, big_wram=True), ("synth", 9, 60, 1000, 0.05)),
    ("wfa_nonacgt", dict(WFA, max_score=30, read_size=168, backtrace=True, reduce=True), ("dirty", 13, 300, 150, 0.04)),
    ("nw_nonacgt", dict(alg="nw", mem="wram", max_score=30, read_size=168, mismatch=3, gap_o=4, backtrace=True), ("dirty", 13, 300, 150, 0.04)),
]


def make_input(spec, read_size: int, path: Path) -> int:
    kind = spec[0]
    if kind == "dataset":
        raw = lzma.open(G / "datasets" / (spec[1] + ".xz")).read()
        path.write_bytes(raw)
        return spec[2]
    _, seed, n, length, err = spec
    plen, tlen, pats, txts = A.generate_pairs(seed, n, length, err, read_size, nthreads=1)
    if kind == "dirty":  # sprinkle bytes outside ACGT: the reference compares raw bytes (wfa.c:209)
        rng = np.random.default_rng(seed)
        for i in range(0, n, 3):
            for arr, ln in ((pats, plen), (txts, tlen)):
                pos = int(rng.integers(0, ln[i]))
                arr[i, pos] = ord(rng.choice(list("NnacgtRY")))
            if i % 2 == 0:  # the same unusual byte on both sides at an aligned spot -> a match
                pos = int(rng.integers(0, min(plen[i], tlen[i]) // 4))
                pats[i, pos] = txts[i, pos] = ord("N")
    A.write_pairs(path, plen, tlen, pats, txts)
    return n


def main() -> None:
    manifest = []
    for name, kw, spec in CASES:
        kw = dict(kw)
        alg, mem = kw.pop("alg"), kw.pop("mem")
        binary = rb.build_ref(alg, mem, **kw)
        with tempfile.TemporaryDirectory() as tmp:
            pairs = Path(tmp) / "in.pairs"
            n_arg = make_input(spec, kw["read_size"], pairs)
            out = Path(tmp) / "ref.out"
            rb.run_ref(binary, pairs, out, n_arg)
            entry = dict(name=name, algo=alg, variant=mem, n_arg=n_arg, md5=rb.md5(out),
                         lines=out.read_bytes().count(b"\n"), reference_binary=binary.name,
                         params={k: (int(v) if not isinstance(v, bool) else v) for k, v in kw.items() if k != "big_wram"})
            if spec[0] == "dataset":
                entry["input"] = f"datasets/{spec[1]}.xz"
            else:
                entry["input"] = f"{name}.pairs.xz"
                entry["generator"] = dict(kind=spec[0], seed=spec[1], n=spec[2], length=spec[3], error=spec[4])
                (G / entry["input"]).write_bytes(lzma.compress(pairs.read_bytes(), preset=9))
                entry["output"] = f"{name}.out.xz"
                (G / entry["output"]).write_bytes(lzma.compress(out.read_bytes(), preset=9))
            manifest.append(entry)
            print(name, entry["md5"], entry["lines"])
    (G / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")


if __name__ == "__main__":
    main()
