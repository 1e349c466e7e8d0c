"""include/dpu.h + libaim_dpu.so: the reference's host.c files compile UNCHANGED against the UPMEM host-API adapter.

CPU part (here): the library loads and exports every function dpu.h declares; where /root/reference is present the six
hosts compile against it, and their host-side plumbing (MRAM images, gather/scatter, the 8/16-byte request and 24/32-byte
result layouts, multi-DPU partition) reproduces the golden outputs with the GPU entry points interposed by a test stub
(tests/stub/aim_stub.c = the CPU oracle behind aim_align_batch, LD_PRELOAD).  GPU part (-m gpu): the prebuilt hosts under
build/upmem_hosts run on the B200 and reproduce the same bytes - see tests/test_gpu_upmem_adapter.py.
"""
import ctypes as C
import lzma
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import GOLDEN, MANIFEST, ROOT, md5_bytes

sys.path.insert(0, str(ROOT / "tools"))
import build_upmem_hosts as B  # noqa: E402

REF = Path(os.environ.get("AIM_REFERENCE_ROOT", "/root/reference"))
have_ref = (REF / "WFA" / "DPU-MRAM" / "host" / "host.c").exists()
CASES = ["cfg1_wfa_sample", "cfg1_wfa_err", "cfg2_nw_sample", "swg_sample", "cfg4_wfa_adaptive_synth", "wfa_l150_scoreonly",
         "wfa_l150_giveup", "swg_l250_scoreonly", "nw_nonacgt"]


def test_library_exports_every_declared_function():
    lib = C.CDLL(str(ROOT / "aim_b200" / "libaim_dpu.so"))
    hdr = (ROOT / "include" / "dpu.h").read_text()
    names = set(re.findall(r"^(?:dpu_error_t|void|const char \*|struct \w+)\s*\*?\s*((?:aim_)?dpu_\w+)\(", hdr, re.M))
    assert {"dpu_alloc", "dpu_free", "dpu_get_nr_dpus", "dpu_prepare_xfer", "dpu_push_xfer", "dpu_launch", "dpu_log_read",
            "aim_dpu_load", "dpu_error_to_string", "aim_dpu_iterator_from", "aim_dpu_iterator_at"} <= names
    for n in names:
        assert hasattr(lib, n), f"libaim_dpu.so does not export {n}"


def host_kwargs(e):
    p = e["params"]
    return dict(max_score=p["max_score"], read_size=p["read_size"], match=p.get("match", 0), mismatch=p.get("mismatch", 3),
                gap_o=p.get("gap_o", 4), gap_e=p.get("gap_e", 1), backtrace=bool(p.get("backtrace")), reduce=bool(p.get("reduce")))


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    from oracle import oracle as O
    O.build()
    out = tmp_path_factory.mktemp("stub") / "libaim_stub.so"
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "stub" / "aim_stub.c"), "-o", str(out),
                    str(O.LIB), f"-Wl,-rpath,{O.LIB.parent}"], check=True)
    return out


@pytest.mark.skipif(not have_ref, reason="reference tree not present")
@pytest.mark.parametrize("name,nr_dpus", [(n, 1) for n in CASES] + [("cfg1_wfa_sample", 7), ("cfg2_nw_sample", 64), ("cfg4_wfa_adaptive_synth", 3)])
def test_unmodified_hosts_plumbing_matches_golden(name, nr_dpus, stub, tmp_path):
    e = MANIFEST[name]
    host = B.build_host(e["algo"], e["variant"], nr_dpus=nr_dpus, reference=REF, **host_kwargs(e))
    pairs = tmp_path / "in.pairs"
    pairs.write_bytes(lzma.open(GOLDEN / e["input"]).read())
    r = subprocess.run([str(host), str(pairs), str(tmp_path / "out"), str(e["n_arg"])], cwd=tmp_path, capture_output=True, text=True,
                       env=dict(os.environ, LD_PRELOAD=str(stub)))
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"Allocated {nr_dpus} DPU(s)" in r.stdout and "DPU Kernel" in r.stdout
    got = (tmp_path / "out").read_bytes()
    assert got.count(b"\n") == e["lines"]
    assert md5_bytes(got) == e["md5"]


@pytest.mark.skipif(not have_ref, reason="reference tree not present")
def test_no_gpu_fails_loudly(tmp_path):
    e = MANIFEST["cfg1_wfa_sample"]
    host = B.build_host(e["algo"], e["variant"], reference=REF, **host_kwargs(e))
    pairs = tmp_path / "in.pairs"
    pairs.write_bytes(lzma.open(GOLDEN / e["input"]).read()[:4000])
    import aim_b200 as A
    if A.device_count() > 0:
        pytest.skip("a GPU is visible")
    r = subprocess.run([str(host), str(pairs), str(tmp_path / "out"), "16"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
