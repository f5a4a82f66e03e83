import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_LIB = os.path.join(ROOT, "oracle", "libamh_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_oracle():
    if not os.path.exists(ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_LIB


@pytest.fixture(scope="session")
def amh():
    import amh_b200
    return amh_b200


@pytest.fixture(scope="session")
def oracle(amh):
    """the CPU oracle behind the same ABI (prefix amho_) -- the checker, never the product"""
    return amh.Engine(lib_path=_ensure_oracle(), prefix="amho_")


@pytest.fixture(scope="session")
def cuda(amh):
    """the product: libamh_b200.so on cuda:0; fails loudly if missing"""
    return amh.default_engine(0)


def make_spd(d, seed, lo=1.0, hi=100.0):
    """Sigma = Q diag(lambda) Q' with log-spaced eigenvalues"""
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(d, d)))
    lam = np.exp(np.linspace(np.log(lo), np.log(hi), d))
    S = (Q * lam) @ Q.T
    return (S + S.T) / 2
