"""Committed golden vectors (tests/golden/contract_v1.json, contract_v2.json, made by tools/gen_golden.py): for each
version of the numerical contract the oracle must reproduce them bit-for-bit on CPU, and the CUDA engine must
reproduce them bit-for-bit on the B200."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases as G       # noqa: E402

DOCS = {cv: json.load(open(os.path.join(HERE, "golden", f"contract_v{cv}.json"))) for cv in (1, 2)}


def _check_cases(amh, eng, cv):
    doc = DOCS[cv]
    assert doc["contract_version"] == cv
    with amh.contract(cv):
        cases = G.build_cases(amh)
        assert sorted(c[0] for c in cases) == sorted(doc["cases"].keys())
        for case in cases:
            got = G.run_case(amh, eng, case)
            want = doc["cases"][case[0]]
            for k, enc in want.items():
                w = G.decode(enc)
                g = got[k]
                if w.dtype == np.float64:
                    assert np.array_equal(g.view(np.uint64), w.view(np.uint64)), f"v{cv} {case[0]}: {k} differs"
                else:
                    assert np.array_equal(g, w), f"v{cv} {case[0]}: {k} differs"


@pytest.mark.parametrize("cv", [1, 2])
def test_oracle_reproduces_golden_noise(oracle, cv):
    assert oracle.contract_version() == 2                  # the default of new runs; v1 stays selectable
    dp = C.POINTER(C.c_double)
    for rec in DOCS[cv]["noise"]:
        d = rec["d"]
        z = np.empty(d); e = C.c_double()
        oracle.lib.amho_probe_step_noise_cv(cv, C.c_uint64(rec["seed"]), C.c_uint64(rec["step"]), d, z.ctypes.data_as(dp), C.byref(e))
        assert np.array_equal(z.view(np.uint64), G.decode(rec["z"]).view(np.uint64))
        assert e.value == G.decode(rec["e"])[0]


@pytest.mark.parametrize("cv", [1, 2])
def test_oracle_reproduces_golden_cases(amh, oracle, cv):
    _check_cases(amh, oracle, cv)


@pytest.mark.gpu
@pytest.mark.parametrize("cv", [1, 2])
def test_cuda_reproduces_golden_cases(amh, cuda, cv):
    assert cuda.contract_version() == 2
    _check_cases(amh, cuda, cv)
