"""Committed golden vectors (tests/golden/contract_v1.json, made by tools/gen_golden.py): the oracle must
reproduce them bit-for-bit on CPU, and the CUDA engine must reproduce them bit-for-bit on the B200."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases as G       # noqa: E402

DOC = json.load(open(os.path.join(HERE, "golden", "contract_v1.json")))


def _check_cases(amh, eng):
    cases = G.build_cases(amh)
    assert sorted(c[0] for c in cases) == sorted(DOC["cases"].keys())
    for case in cases:
        got = G.run_case(amh, eng, case)
        want = DOC["cases"][case[0]]
        for k, enc in want.items():
            w = G.decode(enc)
            g = got[k]
            if w.dtype == np.float64:
                assert np.array_equal(g.view(np.uint64), w.view(np.uint64)), f"{case[0]}: {k} differs"
            else:
                assert np.array_equal(g, w), f"{case[0]}: {k} differs"


def test_oracle_reproduces_golden_noise(oracle):
    assert DOC["contract_version"] == oracle.contract_version()
    dp = C.POINTER(C.c_double)
    for rec in DOC["noise"]:
        d = rec["d"]
        z = np.empty(d); e = C.c_double()
        oracle.lib.amho_probe_step_noise(C.c_uint64(rec["seed"]), C.c_uint64(rec["step"]), d, z.ctypes.data_as(dp), C.byref(e))
        assert np.array_equal(z.view(np.uint64), G.decode(rec["z"]).view(np.uint64))
        assert e.value == G.decode(rec["e"])[0]


def test_oracle_reproduces_golden_cases(amh, oracle):
    _check_cases(amh, oracle)


@pytest.mark.gpu
def test_cuda_reproduces_golden_cases(amh, cuda):
    assert DOC["contract_version"] == cuda.contract_version()
    _check_cases(amh, cuda)
