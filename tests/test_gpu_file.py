"""GPU: `host <pairs> <out> <N>` through aim_align_file (pair file parsed and output formatted ON THE GPU) writes the bytes the
reference host writes - checked against the unmodified reference binaries (oracle/_ref, prebuilt) on adversarial pair files
(no trailing newline, odd line count, CRLF, empty and one-character lines, N below / above the pair count, multi-DPU rounding,
lines much shorter than READ_SIZE), on many-chunk runs, and against the batch path (host-side get_reads + printer)."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import aim_b200 as A
from oracle import refbuild as rb

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
HOST = ROOT / "build" / "host"
WFA = dict(alg="wfa", mem="mram", max_score=30, read_size=168, mismatch=3, gap_o=4, gap_e=1, backtrace=True, reduce=True)
NW = dict(alg="nw", mem="wram", max_score=4, read_size=112, mismatch=3, gap_o=4, backtrace=True)


def _env(kw, nr_dpus=1, **extra):
    e = dict(os.environ, AIM_ALGO=kw["alg"], AIM_VARIANT=kw["mem"], MAX_SCORE=str(kw["max_score"]), READ_SIZE=str(kw["read_size"]),
             MISMATCH=str(kw["mismatch"]), BACKTRACE="1" if kw.get("backtrace") else "0", REDUCE="1" if kw.get("reduce") else "0",
             NR_DPUS=str(nr_dpus))
    if kw["alg"] == "nw":
        e.update(GAP_I=str(kw["gap_o"]), GAP_D=str(kw["gap_o"]))
    else:
        e.update(GAP_O=str(kw["gap_o"]), GAP_E=str(kw["gap_e"]))
    e.update({k: str(v) for k, v in extra.items()})
    return e


class _R:
    def __init__(self, returncode, stdout="", stderr=""):
        self.returncode, self.stdout, self.stderr = returncode, stdout, stderr


def _ours(kw, pairs, out, n, cwd, nr_dpus=1, cli=False, **extra):
    """`host <pairs> <out> <N>`: as a process (cli=True) or - same code path, no process start-up and CUDA init per case - in
    process: aim_align_file (the stream path host.cpp calls) / read_pairs + align_batch + write_results (its batch path)."""
    if cli:
        return subprocess.run([str(HOST), str(pairs), str(out), str(n)], cwd=cwd, capture_output=True, text=True, env=_env(kw, nr_dpus, **extra), timeout=600)
    params = A.AlignParams(algo=kw["alg"], mismatch=kw["mismatch"], gap_open=kw["gap_o"], gap_ext=kw.get("gap_e", 1), max_score=kw["max_score"],
                           read_size=kw["read_size"], backtrace=bool(kw.get("backtrace")), reduce=bool(kw.get("reduce")), ngpus=int(extra.get("AIM_NGPUS", 1)))
    old = {k: os.environ.get(k) for k in ("AIM_FILE_CHUNK_MB", "AIM_IO_THREADS")}
    try:
        for k in old:
            if k in extra:
                os.environ[k] = str(extra[k])
            else:
                os.environ.pop(k, None)
        if extra.get("AIM_HOST_PATH") == "batch":
            want = A.pairs_to_process(A.count_pairs(pairs), n, nr_dpus)
            try:
                arrays = A.read_pairs(pairs, kw["read_size"], want)
            except A.AimError as e:
                if e.code == -2:
                    Path(out).write_bytes(b"")
                    return _R(0, "READ LENGTH less than length of the input reads")
                raise
            p1 = A.AlignParams(**{**params.__dict__, "ngpus": 1})
            res, ops, _ = A.align_batch(p1, *arrays)
            A.write_results(out, res, ops, kw["read_size"], p1.backtrace)
            return _R(0)
        try:
            A.align_file(params, pairs, out, n, nr_dpus)
        except A.AimError as e:
            if e.code == -2:
                return _R(0, "READ LENGTH less than length of the input reads")
            return _R(1, "", str(e))
        return _R(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _pairs_text(n, length, error, rs, seed=77):
    plen, tlen, pats, txts = A.generate_pairs(seed, n, length, error, rs)
    lines = []
    for i in range(n):
        lines.append(b">" + bytes(pats[i, :plen[i]]))
        lines.append(b"<" + bytes(txts[i, :tlen[i]]))
    return lines


def _variants():
    base = _pairs_text(257, 150, 0.04, 168)
    v = {}
    v["plain"] = b"\n".join(base) + b"\n"
    v["no_trailing_newline"] = b"\n".join(base)
    v["odd_line_count"] = b"\n".join(base[:-1]) + b"\n"
    v["odd_line_count_no_newline"] = b"\n".join(base[:-1])
    v["crlf"] = b"\r\n".join(base[:60]) + b"\r\n"
    short = list(base[:40])
    short[6] = b">"          # one-character line: length -1 in the reference, clamped to 0
    short[9] = b"<A"
    short[12] = b""          # empty line
    v["tiny_lines"] = b"\n".join(short) + b"\n"
    v["non_acgt"] = b"\n".join(base[:50]).replace(b"ACG", b"ANG", 7).replace(b"T", b"t", 3) + b"\n"
    return v


@pytest.mark.parametrize("name", sorted(_variants()))
@pytest.mark.parametrize("n_arg,nr_dpus", [(1000, 1), (101, 4)])
def test_adversarial_files_match_reference_binary(name, n_arg, nr_dpus, tmp_path):
    data = _variants()[name]
    pairs = tmp_path / "in.pairs"
    pairs.write_bytes(data)
    binary = rb.build_ref(**WFA)
    try:
        rb.run_ref(binary, pairs, tmp_path / "ref.out", n_arg, nr_dpus=nr_dpus, timeout=120)
    except RuntimeError:
        pytest.skip("the reference itself fails on this input")
    want = (tmp_path / "ref.out").read_bytes()
    for extra in ({}, {"AIM_HOST_PATH": "batch"}, {"cli": name in ("no_trailing_newline", "tiny_lines")}):
        r = _ours(WFA, pairs, tmp_path / "our.out", n_arg, tmp_path, nr_dpus, **extra)
        assert r.returncode == 0, r.stdout + r.stderr
        assert (tmp_path / "our.out").read_bytes() == want, (name, n_arg, nr_dpus, extra)


@pytest.mark.parametrize("kw,length,error,n", [(WFA, 150, 0.04, 60_000), (NW, 100, 0.01, 50_000), (WFA, 20, 0.05, 120_000)])
def test_many_chunks_equal_batch_path(kw, length, error, n, tmp_path):
    """Chunks of 1 MiB (dozens of them; with 20-base reads a chunk holds more pairs than its row buffers: the reader cuts earlier)
    and the default chunking write the same bytes as the batch path, which the golden tests pin on the reference's outputs."""
    plen, tlen, pats, txts = A.generate_pairs(5, n, length, error, kw["read_size"])
    pairs = tmp_path / "in.pairs"
    A.write_pairs(pairs, plen, tlen, pats, txts)
    outs = []
    for extra in ({"AIM_HOST_PATH": "batch"}, {}, {"AIM_FILE_CHUNK_MB": 1, "AIM_IO_THREADS": 3}):
        r = _ours(kw, pairs, tmp_path / "o", n + 8, tmp_path, **extra)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((tmp_path / "o").read_bytes())
    assert outs[0].count(b"\n") == (2 if kw.get("backtrace") else 1) * n
    assert all(o == outs[0] for o in outs[1:])
    # the same file without its final newline and with a dangling pattern line: end-of-file handling in the LAST of many chunks
    data = pairs.read_bytes()
    for tail in (data[:-1], data + b">ACGT\n", data + b">ACGT"):
        pairs.write_bytes(tail)
        got = []
        for extra in ({"AIM_HOST_PATH": "batch"}, {"AIM_FILE_CHUNK_MB": 1}):
            r = _ours(kw, pairs, tmp_path / "o", n + 8, tmp_path, **extra)
            assert r.returncode == 0, r.stdout + r.stderr
            got.append((tmp_path / "o").read_bytes())
        assert got[0] == got[1] and got[0].count(b"\n") == (2 if kw.get("backtrace") else 1) * n


def test_too_long_read_leaves_empty_output_and_exits_zero(tmp_path):  # host.c:119-123
    lines = _pairs_text(3000, 150, 0.04, 168)
    lines[4001] = b"<" + b"ACGT" * 60  # 240 > READ_SIZE 168, deep inside the file
    pairs = tmp_path / "in.pairs"
    pairs.write_bytes(b"\n".join(lines) + b"\n")
    r = _ours(WFA, pairs, tmp_path / "o", 3000, tmp_path, cli=True, AIM_FILE_CHUNK_MB=1)
    assert r.returncode == 0 and "READ LENGTH less than length of the input reads" in r.stdout
    assert (tmp_path / "o").read_bytes() == b""


def test_score_only_and_two_gpus(tmp_path):
    n = 40_000
    plen, tlen, pats, txts = A.generate_pairs(6, n, 150, 0.04, 168)
    pairs = tmp_path / "in.pairs"
    A.write_pairs(pairs, plen, tlen, pats, txts)
    kw = dict(WFA, backtrace=False)
    a = _ours(kw, pairs, tmp_path / "a", n, tmp_path, AIM_HOST_PATH="batch")
    b = _ours(kw, pairs, tmp_path / "b", n, tmp_path, AIM_FILE_CHUNK_MB=1)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert (tmp_path / "a").read_bytes() == (tmp_path / "b").read_bytes()
    if A.device_count() >= 2:
        c = _ours(WFA, pairs, tmp_path / "c", n, tmp_path, AIM_FILE_CHUNK_MB=1, AIM_NGPUS=2)
        d = _ours(WFA, pairs, tmp_path / "d", n, tmp_path)
        assert c.returncode == 0 and d.returncode == 0, c.stderr + d.stderr
        assert (tmp_path / "c").read_bytes() == (tmp_path / "d").read_bytes()


@pytest.mark.parametrize("alg,mem,k,length,rs", [("genasm_dc", "wram", 5, 100, 120), ("genasm_dc", "mram", 8, 150, 168), ("genasm_filter", "wram", 2, 100, 120)])
def test_genasm_stream_equals_batch_path(alg, mem, k, length, rs, tmp_path):
    """The GenASM hosts' output lines (`idx, score, CIGAR` / `idx, score`, aim-genasm/GenASM/DPU-WRAM-DC/host/host.c:286-296) formatted
    on the GPU by aim_align_file = the batch path's host printer (which the golden GenASM cases pin), also over many 1 MiB chunks."""
    n = 30_000
    plen, tlen, pats, txts = A.generate_pairs(11, n, length, 0.03, rs)
    pairs = tmp_path / "in.pairs"
    A.write_pairs(pairs, plen, tlen, pats, txts)
    kw = dict(alg=alg, mem=mem, max_score=k, read_size=rs, mismatch=3, gap_o=4, gap_e=1)
    outs = []
    for extra in ({"AIM_HOST_PATH": "batch"}, {}, {"AIM_FILE_CHUNK_MB": 1, "AIM_IO_THREADS": 3}):
        r = _ours(kw, pairs, tmp_path / "o", n + 8, tmp_path, cli=True, **extra)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((tmp_path / "o").read_bytes())
    assert outs[0].count(b"\n") == n
    assert all(o == outs[0] for o in outs[1:])
