#!/usr/bin/env python3
"""Regenerates tests/golden/contract_v1.json and contract_v2.json from the CPU oracle (see tests/golden/cases.py for why
the oracle, not the Julia reference, is the generator), one file per version of the numerical contract
(include/amh_contract.h).  Run from the repository root:  python tools/gen_golden.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import amh_b200 as amh          # noqa: E402
import cases as G               # noqa: E402

orc = amh.Engine(lib_path=os.path.join(ROOT, "oracle", "libamh_oracle.so"), prefix="amho_")
dp = C.POINTER(C.c_double)
for cv in (1, 2):
    doc = {"contract_version": cv, "generator": "tools/gen_golden.py (CPU oracle)", "noise": [], "cases": {}}
    for seed, step, d in [(0, 0, 1), (1, 1, 2), (0xDEADBEEFCAFEF00D, 7, 5), (2 ** 64 - 1, 2 ** 40 + 3, 32), (42, 100000, 64)]:
        z = np.empty(d); e = C.c_double()
        orc.lib.amho_probe_step_noise_cv(cv, C.c_uint64(seed), C.c_uint64(step), d, z.ctypes.data_as(dp), C.byref(e))
        doc["noise"].append({"seed": seed, "step": step, "d": d, "z": G.encode(z), "e": G.encode(np.array([e.value]))})
    with amh.contract(cv):
        for case in G.build_cases(amh):
            res = G.run_case(amh, orc, case)
            doc["cases"][case[0]] = {k: G.encode(v) for k, v in res.items()}
    path = os.path.join(ROOT, "tests", "golden", f"contract_v{cv}.json")
    json.dump(doc, open(path, "w"), separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(doc["cases"]), "cases")
