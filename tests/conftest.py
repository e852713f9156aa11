import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU parity oracle (test infrastructure)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def pg():
    """The CUDA library bound to cuda:0.  Fails loudly when it cannot be loaded."""
    import plan_b200
    from plan_b200 import _lib as L
    lib = L.lib()
    L.check(lib.pg_init(0))
    return lib


@pytest.fixture(scope="session")
def sf01_host(oracle):
    """dbgen-equivalent SF0.1 tables on the host (CPU generator)."""
    orders, line = oracle.gen_orders_lineitem(0.1)
    cust = oracle.gen_customer(0.1)
    return {"orders": orders, "lineitem": line, "customer": cust}
