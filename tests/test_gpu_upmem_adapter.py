"""GPU: the reference's host programs, compiled UNCHANGED against include/dpu.h + libaim_dpu.so (tools/build_upmem_hosts.py,
prebuilt into build/upmem_hosts by __graft_entry__.build() where /root/reference exists), run on the B200 and write the
reference's output bytes (md5 of SURVEY.md App. C for the two Datasets)."""
import lzma
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN, MANIFEST, ROOT, md5_bytes

sys.path.insert(0, str(ROOT / "tools"))
import build_upmem_hosts as B  # noqa: E402

pytestmark = pytest.mark.gpu
CASES = [("cfg1_wfa_sample", 1), ("cfg1_wfa_sample", 7), ("cfg1_wfa_err", 1), ("cfg2_nw_sample", 1), ("cfg2_nw_err", 64), ("swg_sample", 1),
         ("cfg3_swg_synth", 1), ("cfg4_wfa_adaptive_synth", 3), ("wfa_l150_scoreonly", 1), ("wfa_l150_giveup", 1), ("wfa_nonacgt", 1),
         ("cfg5_wfa_long_scoreonly", 1)]


def _host(e, nr_dpus):
    p = e["params"]
    name = B.host_name(e["algo"], e["variant"], max_score=p["max_score"], read_size=p["read_size"], match=p.get("match", 0),
                       mismatch=p.get("mismatch", 3), gap_o=p.get("gap_o", 4), gap_e=p.get("gap_e", 1), backtrace=bool(p.get("backtrace")),
                       reduce=bool(p.get("reduce")), nr_dpus=nr_dpus)
    path = ROOT / "build" / "upmem_hosts" / name
    assert path.exists(), f"{path} missing: __graft_entry__.build() prebuilds it where the reference tree exists"
    return path


@pytest.mark.parametrize("name,nr_dpus", CASES)
def test_unmodified_reference_host_on_b200(name, nr_dpus, tmp_path):
    e = MANIFEST[name]
    pairs = tmp_path / "in.pairs"
    pairs.write_bytes(lzma.open(GOLDEN / e["input"]).read())
    r = subprocess.run([str(_host(e, nr_dpus)), str(pairs), str(tmp_path / "out"), str(e["n_arg"])], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"Allocated {nr_dpus} DPU(s)" in r.stdout
    got = (tmp_path / "out").read_bytes()
    assert got.count(b"\n") == e["lines"]
    assert md5_bytes(got) == e["md5"]


def test_two_gpus_through_the_adapter(tmp_path):
    import aim_b200 as A
    if A.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    e = MANIFEST["cfg1_wfa_sample"]
    pairs = tmp_path / "in.pairs"
    pairs.write_bytes(lzma.open(GOLDEN / e["input"]).read())
    r = subprocess.run([str(_host(e, 7)), str(pairs), str(tmp_path / "out"), str(e["n_arg"])], cwd=tmp_path, capture_output=True, text=True,
                       env=dict(os.environ, AIM_NGPUS="all"), timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert md5_bytes((tmp_path / "out").read_bytes()) == e["md5"]
