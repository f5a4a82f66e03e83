#!/usr/bin/env python3
"""Regenerates tests/golden/contract_v1.json from the CPU oracle (see tests/golden/cases.py for why the
oracle, not the Julia reference, is the generator).  Run from the repository root:  python tools/gen_golden.py"""
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
doc = {"contract_version": orc.contract_version(), "generator": "tools/gen_golden.py (CPU oracle)", "noise": [], "cases": {}}
dp = C.POINTER(C.c_double)
for seed, step, d in [(0, 0, 1), (1, 1, 2), (0xDEADBEEFCAFEF00D, 7, 5), (2 ** 64 - 1, 2 ** 40 + 3, 32), (42, 100000, 64)]:
    z = np.empty(d); e = C.c_double()
    orc.lib.amho_probe_step_noise(C.c_uint64(seed), C.c_uint64(step), d, z.ctypes.data_as(dp), C.byref(e))
    doc["noise"].append({"seed": seed, "step": step, "d": d, "z": G.encode(z), "e": G.encode(np.array([e.value]))})
for case in G.build_cases(amh):
    res = G.run_case(amh, orc, case)
    doc["cases"][case[0]] = {k: G.encode(v) for k, v in res.items()}
path = os.path.join(ROOT, "tests", "golden", "contract_v1.json")
json.dump(doc, open(path, "w"), separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes;", len(doc["cases"]), "cases")
