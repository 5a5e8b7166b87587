import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def shuf_l3k10():
    """Deterministic subk=6 permutation used by every L3K10 test and golden vector (seed 1)."""
    from public_kssd_b200 import synth
    return synth.make_shuf_table(6, 1)


@pytest.fixture(scope="session")
def shuf_s5():
    from public_kssd_b200 import synth
    return synth.make_shuf_table(5, 2)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_ctx_l3k10(shuf_l3k10):
    from public_kssd_b200 import kssd
    ctx = kssd.Context(10, 6, 3, shuf_l3k10, device=0, shuf_id=4242)
    yield ctx
    ctx.close()
