"""Golden vectors for GenASM-DC / GenASM-filter: outputs of the UNMODIFIED reference
(aim-genasm submodule, built natively by oracle/refbuild.py) on small seeded inputs and on the two reference
Datasets.  Run in the authoring container (needs /root/reference):  python tests/golden/make_golden_genasm.py

Lines of pairs whose reference output is not a function of the pair (oracle/aim_oracle.c: a text byte outside
ACGTacgt, traceback reaching text row n, "No alignment found") are kept as the reference wrote them here; the
tests skip exactly the pairs the oracle flags.
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
OUT = G / "genasm"

CASES = [
    # name, algo, reference directory variant, max_score (k), read_size, input spec
    ("dc_wram_sample", "genasm_dc", "wram", 5, 112, ("dataset", "sample-l100-e1-40K", 40000)),
    ("dc_mram_sample", "genasm_dc", "mram", 5, 112, ("dataset", "sample-l100-e1-40K", 40000)),
    ("dc_wram_err", "genasm_dc", "wram", 5, 112, ("dataset", "ERR240727-l100-e1-30000Pairs", 30000)),
    ("filter_wram_sample", "genasm_filter", "wram", 5, 112, ("dataset", "sample-l100-e1-40K", 40000)),
    ("filter_mram_err", "genasm_filter", "mram", 5, 112, ("dataset", "ERR240727-l100-e1-30000Pairs", 30000)),
    ("filter_wram_sample_k2", "genasm_filter", "wram", 2, 112, ("dataset", "sample-l100-e1-40K", 40000)),
    ("dc_wram_l150_k30", "genasm_dc", "wram", 30, 168, ("synth", 4, 600, 150, 0.04)),
    ("dc_wram_l150_k9", "genasm_dc", "wram", 9, 168, ("synth", 4, 600, 150, 0.04)),
    ("dc_wram_l250_k50", "genasm_dc", "wram", 50, 264, ("synth", 3, 200, 250, 0.04)),
    ("filter_wram_l250_k50", "genasm_filter", "wram", 50, 264, ("synth", 3, 200, 250, 0.04)),
    ("dc_wram_l60_k3", "genasm_dc", "wram", 3, 72, ("synth", 6, 800, 60, 0.03)),
    # DC: unusual bytes in the PATTERN only - a text byte outside ACGTacgt leaves traceback rows unwritten and the
    # reference's traceback then spins forever on whatever the rows hold (genasmDC.c:107-322 has no final else)
    ("dc_wram_dirty", "genasm_dc", "wram", 30, 168, ("dirty_pattern", 13, 300, 150, 0.04)),
    ("dc_mram_dirty", "genasm_dc", "mram", 30, 168, ("dirty_pattern", 13, 300, 150, 0.04)),
    ("filter_wram_dirty", "genasm_filter", "wram", 30, 168, ("dirty", 13, 300, 150, 0.04)),
]


def make_input(spec, read_size: int, path: Path) -> int:
    kind = spec[0]
    if kind == "dataset":
        path.write_bytes(lzma.open(G / "datasets" / (spec[1] + ".xz")).read())
        return spec[2]
    _, seed, n, length, err = spec
    plen, tlen, pats, txts = A.generate_pairs(seed, n, length, err, read_size, nthreads=1)
    if kind.startswith("dirty"):  # N wildcards in the pattern (genasmDC.c:75-81), lower case, and bytes the text loop skips
        rng = np.random.default_rng(seed)
        for i in range(0, n, 2):
            pos = int(rng.integers(0, plen[i]))
            pats[i, pos] = ord(rng.choice(list("NnacgtRY")))
            if i % 6 == 0 and kind == "dirty":
                pos = int(rng.integers(0, tlen[i]))
                txts[i, pos] = ord(rng.choice(list("NacgtR")))
    A.write_pairs(path, plen, tlen, pats, txts)
    return n


def main() -> None:
    OUT.mkdir(exist_ok=True)
    manifest = []
    for name, alg, mem, k, rs, spec in CASES:
        binary = rb.build_ref(alg, mem, max_score=k, read_size=rs, backtrace=(alg == "genasm_dc"))
        with tempfile.TemporaryDirectory() as tmp:
            pairs = Path(tmp) / "in.pairs"
            n_arg = make_input(spec, rs, pairs)
            out = Path(tmp) / "ref.out"
            rb.run_ref(binary, pairs, out, n_arg, timeout=300)
            entry = dict(name=name, algo=alg, variant=mem, n_arg=n_arg, md5=rb.md5(out), lines=out.read_bytes().count(b"\n"),
                         reference_binary=binary.name, params=dict(max_score=k, read_size=rs, mismatch=3, gap_o=4, gap_e=1))
            if spec[0] == "dataset":
                entry["input"] = f"datasets/{spec[1]}.xz"
            else:
                entry["input"] = f"genasm/{name}.pairs.xz"
                entry["generator"] = dict(kind=spec[0], seed=spec[1], n=spec[2], length=spec[3], error=spec[4])
                (G / entry["input"]).write_bytes(lzma.compress(pairs.read_bytes(), preset=9))
            entry["output"] = f"genasm/{name}.out.xz"
            (G / entry["output"]).write_bytes(lzma.compress(out.read_bytes(), preset=9))
            manifest.append(entry)
            print(name, entry["md5"], entry["lines"])
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")


if __name__ == "__main__":
    main()
