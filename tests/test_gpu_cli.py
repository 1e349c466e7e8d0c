"""GPU: the process-level boundary.  `build/host <pairs> <out> <N>` (knobs via the environment, set by
the run-*-pim-*.py wrappers) writes byte-for-byte what the reference host wrote on its own Datasets,
and prints the reference's phase lines (WFA/DPU-MRAM/host/host.c:189-330)."""
import lzma
import re
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import GOLDEN, MANIFEST, md5_bytes

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _dataset(tmp_path, name):
    f = tmp_path / name
    f.write_bytes(lzma.open(GOLDEN / "datasets" / (name + ".xz")).read())
    return f


def test_config1_wfa_via_wrapper(tmp_path):
    f = _dataset(tmp_path, "sample-l100-e1-40K")
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "run-wfa-pim-mram.py"), "-i", str(f), "-o", str(out),
                        "-l", "100", "-e", "0.01", "-n", "40000", "-b"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert md5_bytes(out.read_bytes()) == MANIFEST["cfg1_wfa_sample"]["md5"] == "63dfdb4ed4be17b9735e0febef6deeb7"
    for pat in (r"^Allocated 1 DPU\(s\)$", r"^NumReads per dpu = 40000$", r"^Copying data to DPU$", r"^CPU-DPU: [0-9.]+ ms$",
                r"^Run program on DPU\(s\)$", r"^DPU Kernel: [0-9.]+ ms$", r"^Retrieve results$", r"^DPU-CPU: [0-9.]+ ms$"):
        assert re.search(pat, r.stdout, re.M), f"missing stdout line {pat}: {r.stdout}"
    assert (tmp_path / "dpu-out").exists()


def test_config2_nw_via_wrapper_and_dpus_rule(tmp_path):
    f = _dataset(tmp_path, "ERR240727-l100-e1-30000Pairs")
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "run-nw-pim-wram.py"), "-i", str(f), "-o", str(out),
                        "-l", "100", "-e", "0.01", "-n", "30000", "-b", "-d", "1"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert md5_bytes(out.read_bytes()) == MANIFEST["cfg2_nw_err"]["md5"]
    assert "DPU Kernel Time:" in r.stdout and "DPU-CPU Time:" in r.stdout  # NW/DPU-WRAM/host/host.c:300,329
    assert (tmp_path / "dpu_out").exists()
    # N = 1001 over 4 "DPUs": 4 * roundup8(250) = 1024 pairs are aligned (host.c:191)
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "run-nw-pim-mram.py"), "-i", str(f), "-o", str(out),
                        "-l", "100", "-e", "0.01", "-n", "1001", "-d", "4"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert out.read_bytes().count(b"\n") == 1024
    assert "NumReads per dpu = 256" in r.stdout


def test_read_longer_than_read_size_exits_zero(tmp_path):  # host.c:119-123
    f = _dataset(tmp_path, "sample-l100-e1-40K")
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "run-wfa-pim-wram.py"), "-i", str(f), "-o", str(tmp_path / "o"),
                        "-l", "50", "-e", "0.01", "-n", "100"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "READ LENGTH less than length of the input reads" in r.stdout


def test_cigar_rows_and_op_row_fallback_write_the_same_bytes(tmp_path):
    """build/host prints GPU-built CIGAR rows when they fit (config-4 golden case) and falls back to the op rows when some CIGAR is
    longer than a row (l=1000 golden case: ~50 edits per pair); AIM_CIGAR_ROWS=0 forces the op rows.  Same bytes every way."""
    import os
    for name, args in (("cfg4_wfa_adaptive_synth", ["-l", "150", "-e", "0.04", "-b", "-r"]),
                       ("wfa_l1000_e5_bt", ["-l", "1000", "-e", "0.05", "-b", "-r"])):
        e = MANIFEST[name]
        f = tmp_path / (name + ".pairs")
        f.write_bytes(lzma.open(GOLDEN / e["input"]).read())
        for env_extra in ({}, {"AIM_CIGAR_ROWS": "0"}):
            out = tmp_path / "out"
            r = subprocess.run([sys.executable, str(ROOT / "scripts" / "run-wfa-pim-mram.py"), "-i", str(f), "-o", str(out),
                                "-n", str(e["n_arg"])] + args, capture_output=True, text=True, cwd=tmp_path,
                               env=dict(os.environ, **env_extra))
            assert r.returncode == 0, r.stdout + r.stderr
            assert md5_bytes(out.read_bytes()) == e["md5"], (name, env_extra)
