import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


CPU_PORT = os.path.join(ROOT, "oracle", "_ref", "libhzw_cpu_port.so")
CUDA_LIB = os.path.join(ROOT, "circuits_b200", "libhzwitness.so")


@pytest.fixture(scope="session")
def built():
    """Build both libraries once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    if not (os.path.exists(CPU_PORT) and os.path.exists(CUDA_LIB)):
        g.build()
    return True


@pytest.fixture(scope="session")
def small_bb():
    from circuits_b200.synth import small_batch
    return small_batch(3, 16, 2, 2)


@pytest.fixture(scope="session")
def small_oracle(small_bb):
    """Oracle witness + symbol table of the small batch (a few seconds of big-int work)."""
    from oracle import hermez as hz
    from oracle.circuit import symbols
    w, main, _ = hz.rollup_main_witness(3, 16, 2, 2, small_bb.getInput())
    return w, symbols(main)


def cpu_port_calculator(*params, **kw):
    """The CPU port behind the product's own binding class (checker side only)."""
    from circuits_b200.witness_calculator import WitnessCalculator
    return WitnessCalculator(*params, lib_path=CPU_PORT, **kw)
