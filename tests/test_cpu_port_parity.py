"""CPU (no GPU) checks: the circuit templates shared by the CUDA kernels, compiled for the host
(oracle/cpu_port), against the independent big-integer oracle; the C ABI surface; `.wtns`."""
import ctypes
import os
import struct

import pytest

from conftest import cpu_port_calculator, CUDA_LIB, ROOT
from oracle.field import P


def test_cuda_library_exports_every_declared_symbol(built):
    """include/hzwitness.h <-> libhzwitness.so; no compute call (no GPU here)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "hzwitness.h")).read()
    declared = sorted(set(re.findall(r"\b(hzw_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 12
    lib = ctypes.CDLL(CUDA_LIB)
    for name in declared:
        assert getattr(lib, name) is not None, name
    from circuits_b200.witness_calculator import EXPORTS
    assert sorted(EXPORTS) == declared


def test_product_fails_loudly_without_gpu(built):
    """No CPU fallback: without a CUDA device construction must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from circuits_b200.witness_calculator import WitnessCalculator, WitnessError
    with pytest.raises(WitnessError):
        WitnessCalculator(3, 16, 2, 2)


def test_rollup_main_small_matches_oracle_signal_by_signal(built, small_bb, small_oracle):
    w_ref, sym_ref = small_oracle
    wc = cpu_port_calculator(3, 16, 2, 2)
    assert wc.NVars == len(w_ref)
    w = wc.calculateWitness(small_bb.getInput(), True)
    assert w == w_ref
    assert w[0] == 1 and w[1] == small_bb.getHashInputs()
    # the `.sym` equivalent: every name, aliases included, resolves to the same index
    sym = wc.symbols("main")
    assert sym == sym_ref
    assert wc.getSignal(w, "main.hashGlobalInputs") == small_bb.getHashInputs()
    wc.assertOut(w, {"hashGlobalInputs": small_bb.getHashInputs()})


def test_wtns_container_layout(built, small_bb, tmp_path):
    wc = cpu_port_calculator(3, 16, 2, 2)
    wb = wc.calculateBinWitness(small_bb.getInput(), False)
    path = str(tmp_path / "w.wtns")
    wc.writeWtns(wb, path)
    raw = open(path, "rb").read()
    assert raw[:4] == b"wtns"
    ver, nsec = struct.unpack("<II", raw[4:12])
    assert (ver, nsec) == (2, 2)
    sid, size, n8 = struct.unpack("<IQI", raw[12:28])
    assert (sid, size, n8) == (1, 40, 32)
    assert int.from_bytes(raw[28:60], "little") == P
    nwit, = struct.unpack("<I", raw[60:64])
    assert nwit == wc.NVars
    sid2, size2 = struct.unpack("<IQ", raw[64:76])
    assert (sid2, size2) == (2, 32 * wc.NVars)
    assert raw[76:] == bytes(wb)
    from oracle.wtns import wtns_bytes, parse_wtns  # the oracle's own writer
    vals = [int.from_bytes(wb[32 * i:32 * i + 32], "little") for i in range(wc.NVars)]
    assert wtns_bytes(vals) == raw
    assert parse_wtns(raw) == vals


def test_unknown_and_missing_inputs(built, small_bb):
    from circuits_b200.witness_calculator import WitnessError
    wc = cpu_port_calculator(3, 16, 2, 2)
    inp = dict(small_bb.getInput())
    inp["bogus"] = 1
    with pytest.raises(WitnessError, match="Signal bogus is not an input of the circuit"):
        wc.calculateWitness(inp)
    inp = dict(small_bb.getInput())
    del inp["s"]
    with pytest.raises(WitnessError, match="Not all inputs have been set"):
        wc.calculateWitness(inp)


def test_constraint_violation_is_reported_like_the_runtime(built, small_bb):
    """test/rollup-main.test.js:873-877 pattern: tamper an input, expect "Constraint doesn't match"."""
    from circuits_b200.witness_calculator import WitnessError
    wc = cpu_port_calculator(3, 16, 2, 2)
    inp = dict(small_bb.getInput())
    inp["balance1"] = list(inp["balance1"])
    inp["balance1"][0] += 1          # sender state no longer hashes into the old root
    with pytest.raises(WitnessError, match="Constraint doesn't match 1 != 0") as e:
        wc.calculateWitness(inp, True)
    assert "main.rollupTx[0].processor1.checkOldInput" in str(e.value)
    # without sanityCheck the witness is emitted regardless (production CLI behaviour)
    w = wc.calculateWitness(inp, False)
    assert len(w) == wc.NVars


def test_inverse_hint_streams_give_the_same_witness(built, small_bb, small_oracle, monkeypatch):
    """The device evaluates the EdDSA ladders with inverse hints (csrc/hints.cuh: projective
    Edwards walk + one batched inversion, constant window tables recorded once).  The same code
    runs on the host with HZW_CPU_HINTS=1: signal by signal equal to the oracle, and equal to the
    hint-free evaluation on a signed nLevels=32 batch."""
    w_ref, _ = small_oracle
    monkeypatch.setenv("HZW_CPU_HINTS", "1")
    wc = cpu_port_calculator(3, 16, 2, 2)
    assert wc.calculateWitness(small_bb.getInput(), True) == w_ref
    from circuits_b200.synth import bench_batch
    prm = (12, 32, 4, 2)
    bb = bench_batch(*prm, n_accounts=8)
    wh = cpu_port_calculator(*prm)
    packed = wh.packInput(bb.getInput())
    with_hints = bytes(wh.calculateBinWitness(packed, True))
    monkeypatch.delenv("HZW_CPU_HINTS")
    plain = bytes(cpu_port_calculator(*prm).calculateBinWitness(packed, True))
    assert with_hints == plain
    assert int.from_bytes(plain[32:64], "little") == bb.getHashInputs()
