"""CPU: the oracle restatement (oracle/aim_oracle.c) reproduces every golden vector, i.e. the bytes
the UNMODIFIED reference wrote for those inputs (tests/golden/make_golden.py)."""
import lzma

import pytest

from conftest import GOLDEN, MANIFEST, md5_bytes, oracle_kwargs, oracle_results_to_aim, render_output
from oracle import oracle as O


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_oracle_matches_reference_output(name, golden_case, tmp_path):
    e, kw, (plen, tlen, pats, txts) = golden_case(name)
    res, ops = O.align(kw["algo"], plen, tlen, pats, txts, nthreads=4, **oracle_kwargs(kw))
    out = render_output(oracle_results_to_aim(res), ops, kw["read_size"], kw["backtrace"], tmp_path)
    assert out.count(b"\n") == e["lines"]
    assert md5_bytes(out) == e["md5"]
    if "output" in e:
        assert out == lzma.open(GOLDEN / e["output"]).read()


def test_known_answers_from_survey():
    # md5 of the reference's output on its own Datasets
    assert MANIFEST["cfg1_wfa_sample"]["md5"] == "63dfdb4ed4be17b9735e0febef6deeb7"
    assert MANIFEST["cfg2_nw_sample"]["md5"] == "1bb055852cd6112bd40d47a54ff5d0b9"
    assert MANIFEST["cfg2_nw_err"]["md5"] == "10d03e8742d9b200930e747c49d77ab7"
    assert MANIFEST["swg_sample"]["md5"] == "63dfdb4ed4be17b9735e0febef6deeb7"
