"""Golden vectors for the SWG/DPU-WRAM int8 semantics (cells are int8 when MAX_SCORE < 127,
SWG/DPU-WRAM/common/common.h:71-79): outputs of the UNMODIFIED SWG/DPU-WRAM program, built natively, on inputs where the
8-bit cells wrap (so the results differ from the int16 SWG/DPU-MRAM program on every pair).
Run in the authoring container:  python tests/golden/make_golden_swg8.py
"""
from __future__ import annotations

import json
import lzma
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import aim_b200 as A  # noqa: E402  (host-side generator / file writer only)
from oracle import refbuild as rb  # noqa: E402

OUT = Path(__file__).resolve().parent / "swg8"
CASES = [  # name, seed, n, length, error, read_size, max_score, mismatch, gap_o, gap_e
    ("swg8_l100_x4o6e2", 21, 400, 100, 0.01, 112, 8, 4, 6, 2),
    ("swg8_l130_default", 22, 300, 130, 0.02, 144, 13, 3, 4, 1),
    ("swg8_l60_nowrap", 23, 400, 60, 0.03, 72, 9, 3, 4, 1),
]


def main() -> None:
    OUT.mkdir(exist_ok=True)
    manifest = []
    for name, seed, n, length, err, rs, ms, x, o, e in CASES:
        binary = rb.build_ref("swg", "wram", max_score=ms, read_size=rs, mismatch=x, gap_o=o, gap_e=e, backtrace=True)
        plen, tlen, pats, txts = A.generate_pairs(seed, n, length, err, rs, nthreads=1)
        with tempfile.TemporaryDirectory() as tmp:
            pairs, out = Path(tmp) / "in.pairs", Path(tmp) / "ref.out"
            A.write_pairs(pairs, plen, tlen, pats, txts)
            rb.run_ref(binary, pairs, out, n, timeout=300)
            (OUT / f"{name}.pairs.xz").write_bytes(lzma.compress(pairs.read_bytes(), preset=9))
            (OUT / f"{name}.out.xz").write_bytes(lzma.compress(out.read_bytes(), preset=9))
            manifest.append(dict(name=name, n=n, md5=rb.md5(out), reference_binary=binary.name,
                                 params=dict(max_score=ms, read_size=rs, mismatch=x, gap_open=o, gap_ext=e),
                                 generator=dict(seed=seed, n=n, length=length, error=err)))
            print(name, manifest[-1]["md5"])
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1) + "\n")


if __name__ == "__main__":
    main()
