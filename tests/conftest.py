import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are skipped (not failed) on a box without a CUDA device; the product itself still fails
    loudly there (tests/test_cpu_port_parity.py::test_product_fails_loudly_without_gpu)."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_devices() == 0:
        skip = pytest.mark.skip(reason="no CUDA device on this box")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


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


MIXED32 = (12, 32, 7, 3)


@pytest.fixture(scope="session")
def mixed32():
    """Every transaction type in one batch at the BASELINE tree depth (nLevels = 32)."""
    from circuits_b200.synth import mixed_batch
    return mixed_batch(MIXED32[1], MIXED32[0], MIXED32[2], MIXED32[3])


@pytest.fixture(scope="session")
def mixed32_oracle(mixed32):
    """Its witness from the independent big-integer oracle (about half a minute)."""
    from oracle import hermez as hz
    w, _, _ = hz.rollup_main_witness(*MIXED32, mixed32.getInput())
    return w
