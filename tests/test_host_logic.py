"""CPU: the drop-in boundary's host logic (no GPU): knob derivation, pair-file reader quirks, result
writer format, pairs-to-process rule, generator, exported C-ABI symbols."""
import ctypes
import math
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import aim_b200 as A
from aim_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def py_knobs(algo, l, e, x, g, a):
    """The run scripts' own arithmetic (run-wfa-pim-mram.py:58-67, run-nw-pim-mram.py:51-60)."""
    w = l * e
    ms = math.ceil(max(w * x, w * g)) if algo == "nw" else math.ceil(max(w * x, w * (g + a)))
    rs = math.ceil((((l + w) + 7) / 8)) * 8
    return int(ms), int(rs)


@pytest.mark.parametrize("algo,l,e,x,g,a", [
    ("wfa", 100, 0.01, 3, 4, 1), ("nw", 100, 0.01, 3, 4, 1), ("swg", 250, 0.04, 4, 6, 2), ("wfa", 150, 0.04, 3, 4, 1),
    ("wfa", 10000, 0.10, 3, 4, 1), ("wfa", 151, 0.07, 3, 4, 1), ("swg", 333, 0.013, 5, 2, 3), ("nw", 77, 0.09, 2, 9, 1),
])
def test_knobs_match_script_arithmetic(algo, l, e, x, g, a):
    assert A.derive_knobs(algo, l, e, x, g, a) == py_knobs(algo, l, e, x, g, a)


def test_knobs_of_the_five_configs():
    assert A.derive_knobs("wfa", 100, 0.01) == (5, 112)
    assert A.derive_knobs("nw", 100, 0.01, 3, 4) == (4, 112)
    assert A.derive_knobs("swg", 250, 0.04, 4, 6, 2) == (80, 272)
    assert A.derive_knobs("wfa", 150, 0.04) == (30, 168)
    assert A.derive_knobs("wfa", 10000, 0.10) == (5000, 11008)


def test_pairs_to_process_rule():  # host.c:191,201-209
    assert A.pairs_to_process(20000, 40000, 1) == 20000   # dataset names overstate: 40K lines = 20K pairs
    assert A.pairs_to_process(20000, 1001, 1) == 1008      # N is rounded up to a multiple of 8 per DPU
    assert A.pairs_to_process(20000, 1001, 4) == 4 * 256
    assert A.pairs_to_process(100, 1000, 3) == 100


def test_read_pairs_drops_first_and_last_char(tmp_path):  # host.c:112-117
    f = tmp_path / "p"
    f.write_bytes(b">ACGT\n<ACGA\n?TTTT\n!TTT\n>GG\n<GGC")  # markers unchecked; last line lacks '\n' and loses a base
    plen, tlen, pats, txts = A.read_pairs(f, 8)
    assert plen.tolist() == [4, 4, 2] and tlen.tolist() == [4, 3, 2]
    assert bytes(pats[0, :4]) == b"ACGT" and bytes(txts[1, :3]) == b"TTT" and bytes(txts[2, :2]) == b"GG"
    assert A.count_pairs(f) == 3


def test_read_pairs_odd_line_count_and_too_long(tmp_path):
    f = tmp_path / "p"
    f.write_bytes(b">ACGT\n<ACGA\n>AAAA\n")
    assert len(A.read_pairs(f, 8)[0]) == 1
    g = tmp_path / "q"
    g.write_bytes(b">" + b"A" * 20 + b"\n<ACGT\n")
    with pytest.raises(A.AimError) as ei:
        A.read_pairs(g, 16)
    assert ei.value.code == -2


def test_result_writer_format(tmp_path):  # host.c:340-350 + 69-89: "%d, %d, \n" then the RLE CIGAR line
    rs = 8
    res = np.zeros(2, A.RESULT_DTYPE)
    ops = np.full((2, 2 * rs), ord("M"), np.uint8)
    res[0] = (8, 0, 8, 7, 0, 0)
    ops[0, :8] = np.frombuffer(b"MMXMMIDM", np.uint8)
    res[1] = (6, 5, 6, 11, 0, 1)       # give-up: begin = max_ops-1 -> "1M"
    out = tmp_path / "o"
    A.write_results(out, res, ops, rs, True)
    assert out.read_bytes() == b"0, 7, \n2M1X2M1I1D1M\n1, 11, \n1M\n"
    A.write_results(out, res, None, rs, False)
    assert out.read_bytes() == b"0, 7, \n1, 11, \n"
    assert A.cigar_strings(res, ops) == ["2M1X2M1I1D1M", "1M"]


def test_generator_is_deterministic_and_thread_independent():
    a = A.generate_pairs(4, 5000, 150, 0.04, 168, nthreads=1)
    b = A.generate_pairs(4, 5000, 150, 0.04, 168, nthreads=7)
    c = A.generate_pairs(4, 2500, 150, 0.04, 168, first_pair=2500, nthreads=3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for x, y in zip(a, c):
        assert np.array_equal(x[2500:], y)
    plen, tlen, pats, txts = a
    assert (plen == 150).all() and tlen.min() >= 144 and tlen.max() <= 156
    assert set(np.unique(pats[:, :150])) == set(b"ACGT")
    # 6 edits per pair, so the texts differ from the patterns but not wildly
    same = (pats[:, :140] == txts[:, :140]).mean()
    assert 0.3 < same < 0.999


def test_pairs_file_roundtrip(tmp_path):
    plen, tlen, pats, txts = A.generate_pairs(9, 300, 100, 0.05, 112)
    f = tmp_path / "r.pairs"
    A.write_pairs(f, plen, tlen, pats, txts)
    p2, t2, pa2, tx2 = A.read_pairs(f, 112)
    assert np.array_equal(plen, p2) and np.array_equal(tlen, t2) and np.array_equal(pats, pa2) and np.array_equal(txts, tx2)


def test_c_abi_exports_every_declared_symbol():
    header = (ROOT / "include" / "aim_b200.h").read_text()
    declared = set(re.findall(r"\b(aim_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/aim_b200.h but not exported"
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    assert lib.aim_abi_version() == 1
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "orc_align_batch" not in nm, "the product library must not contain the oracle"


def test_product_does_not_import_the_oracle():
    for f in list((ROOT / "aim_b200").rglob("*.py")) + list((ROOT / "aim_b200" / "csrc").glob("*")) + [ROOT / "tools" / "host.cpp"]:
        txt = f.read_text(errors="ignore")
        assert "oracle" not in txt.lower() or f.name == "run_pim.py", f"{f} mentions the oracle"


def test_no_gpu_means_loud_error_not_fallback():
    if A.device_count() > 0:
        pytest.skip("a GPU is visible")
    plen, tlen, pats, txts = A.generate_pairs(1, 8, 50, 0.04, 56)
    with pytest.raises(A.AimError) as ei:
        A.align_batch(A.AlignParams(algo="wfa", max_score=10, read_size=56), plen, tlen, pats, txts)
    assert ei.value.code == -4  # AIM_ERR_NO_DEVICE


def test_run_wrappers_option_surface():
    def run(script, *argv):
        return subprocess.run([sys.executable, str(ROOT / "scripts" / script), *argv], capture_output=True, text=True)
    r = run("run-wfa-pim-mram.py", "-i", "in", "-o", "o", "-l", "150", "-e", "0.04", "-n", "1000", "-b", "-r", "-d", "2", "-t", "3", "--dry-run")
    assert r.returncode == 0
    env = dict(l.split("=", 1) for l in r.stdout.splitlines() if re.match(r"^[A-Z_]+=", l))
    assert env["MAX_SCORE"] == "30" and env["READ_SIZE"] == "168" and env["REDUCE"] == "1" and env["BACKTRACE"] == "1"
    assert env["GAP_O"] == "4" and env["GAP_E"] == "1" and env["NR_DPUS"] == "2" and env["AIM_ALGO"] == "wfa"
    r = run("run-nw-pim-wram.py", "-i", "in", "-l", "100", "-e", "0.01", "-n", "30000", "-g", "4", "--dry-run")
    env = dict(l.split("=", 1) for l in r.stdout.splitlines() if re.match(r"^[A-Z_]+=", l))
    assert env["MAX_SCORE"] == "4" and env["READ_SIZE"] == "112" and env["GAP_I"] == "4" and env["GAP_D"] == "4" and env["BACKTRACE"] == "0"
    assert run("run-nw-pim-wram.py", "-i", "in", "-l", "100", "-e", "0.01", "-n", "3", "-a", "1").returncode == 2  # NW has no -a
    r = run("run-swg-pim-mram.py", "-i", "in", "-l", "250", "-e", "0.04", "-n", "10", "-x", "0", "--dry-run")
    assert r.returncode == 255 and "Wrong affine gap penalties" in r.stdout  # exit(-1), run-swg-pim-mram.py:44-46


def test_host_cli_argument_errors(tmp_path):  # host.c:150-184
    host = ROOT / "build" / "host"
    r = subprocess.run([str(host)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and r.stdout == "wrong number of arguments\n"
    r = subprocess.run([str(host), str(tmp_path / "missing"), str(tmp_path / "o"), "10"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "couldn't be opened" in r.stderr
    f = tmp_path / "p"
    f.write_bytes(b">ACGT\n<ACGA\n")
    r = subprocess.run([str(host), str(f), str(tmp_path / "o"), "1"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and r.stdout == "Allocated DPUs more than needed\n"


def test_wfa_schedule_sizes_history():
    """The data-independent schedule bound used to size the shared-memory history is an upper bound of the
    oracle's actual widths (oracle as checker)."""
    from oracle import oracle as O
    plen, tlen, pats, txts = A.generate_pairs(5, 200, 150, 0.04, 168)
    res, _ = O.align("wfa", plen, tlen, pats, txts, max_score=30, read_size=168, backtrace=False, reduce=False)
    assert res["score"].max() <= 30


@pytest.mark.parametrize("threads", ["1", "2", "3", "7", "16"])
def test_parallel_reader_and_writer_match_a_plain_python_restatement(tmp_path, monkeypatch, threads):
    """The mmap/multi-thread reader (host.c:91-134 semantics) and writer (host.c:332-353, 69-89) against a
    line-by-line Python restatement, with slice boundaries falling anywhere (tiny file, many threads)."""
    monkeypatch.setenv("AIM_IO_THREADS", threads)
    rng = np.random.default_rng(int(threads))
    lines = []
    for i in range(101):  # odd number of lines: the last pattern has no text and is dropped
        ln = int(rng.integers(0, 24))
        lines.append((">" if i % 2 == 0 else "<") + "".join(rng.choice(list("ACGT"), ln)))
    body = "\n".join(lines)  # no trailing newline: the final line loses its last base (T11)
    f = tmp_path / "ragged.pairs"
    f.write_text(body)
    rs = 24
    assert A.count_pairs(f) == 50
    plen, tlen, pats, txts = A.read_pairs(f, rs)
    assert len(plen) == 50
    for i in range(50):
        for arr, ln_arr, line in ((pats, plen, lines[2 * i] + "\n"), (txts, tlen, lines[2 * i + 1] + "\n")):
            want = line[1:-1]
            assert ln_arr[i] == len(want)
            assert bytes(arr[i, :len(want)]).decode() == want
            assert not arr[i, len(want):].any()
    f2 = tmp_path / "nl.pairs"
    f2.write_text(body + "\n")
    assert A.count_pairs(f2) == 50  # 101 lines -> 50 pairs
    p2 = A.read_pairs(f2, rs, 7)
    assert len(p2[0]) == 7

    n = 300
    res = np.zeros(n, A.RESULT_DTYPE)
    res["idx"] = np.arange(n) + 5
    res["score"] = rng.integers(-3, 4000, n)
    ops = rng.choice(np.frombuffer(b"MMMMMMXID", np.uint8), (n, 2 * rs))
    res["begin_offset"] = rng.integers(0, 2 * rs - 1, n)
    res["end_offset"] = [int(rng.integers(b + 1, 2 * rs + 1)) for b in res["begin_offset"]]
    out = tmp_path / "w.out"
    A.write_results(out, res, ops, rs, True)
    want = []
    for i in range(n):
        want.append(f"{int(res['idx'][i])}, {int(res['score'][i])}, ")
        span = bytes(ops[i, res["begin_offset"][i]:res["end_offset"][i]]).decode()
        cig, j = "", 0
        while j < len(span):
            k = j
            while k < len(span) and span[k] == span[j]:
                k += 1
            cig += f"{k - j}{span[j]}"
            j = k
        want.append(cig)
    assert out.read_text() == "\n".join(want) + "\n"
