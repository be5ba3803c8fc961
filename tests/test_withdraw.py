"""Withdraw(nLevels) (/root/reference/src/withdraw.circom, config 5 of BASELINE.json): exit-tree inclusion
witnesses.  The reference's test (test/withdraw.test.js:39-171) builds 4 exits, checks hashGlobalInputs of each
and that a tampered balance raises "Constraint doesn't match 1 != 0"."""
import pytest

from conftest import cpu_port_calculator


def _wc(lib="cpu", nLevels=32):
    from circuits_b200.witness_calculator import WitnessCalculator
    if lib == "cpu":
        return cpu_port_calculator(1, nLevels, 0, 1, circuit="withdraw")
    return WitnessCalculator(1, nLevels, 0, 1, circuit="withdraw")


def test_withdraw_cpu_port_equals_oracle_signal_by_signal(built):
    from circuits_b200.synth import withdraw_inputs
    from oracle import hermez as hz
    from oracle.circuit import symbols
    cases = withdraw_inputs(16, 4)
    wc = _wc("cpu", 16)
    for inp, h in cases:
        w_ref, main, _ = hz.withdraw_witness(16, inp)
        w = wc.calculateWitness(inp, True)
        assert w == w_ref
        assert w[1] == h
    assert wc.symbols("main") == symbols(main)


def test_withdraw_tampered_balance_is_rejected(built):
    from circuits_b200.synth import withdraw_inputs
    from circuits_b200.witness_calculator import WitnessError
    inp, _ = withdraw_inputs(32, 3)[1]
    wc = _wc("cpu")
    bad = dict(inp)
    bad["balance"] = inp["balance"] + 1
    with pytest.raises(WitnessError, match="Constraint doesn't match 1 != 0") as e:
        wc.calculateWitness(bad, True)
    assert "main.smtVerify.checkRoot" in str(e.value)


@pytest.mark.gpu
def test_withdraw_batch_on_gpu_equals_cpu_port(built):
    from circuits_b200.synth import withdraw_inputs
    cases = withdraw_inputs(32, 37)
    gpu, cpu = _wc("gpu"), _wc("cpu")
    packed = b"".join(bytes(gpu.packInput(inp)) for inp, _ in cases)
    wg = bytes(gpu.calculateBatch(len(cases), packed, True))
    wr = bytes(cpu.calculateBatch(len(cases), packed, True))
    assert len(wg) == len(wr) == 32 * gpu.NVars * len(cases)
    assert wg == wr
    for k, (_, h) in enumerate(cases):
        off = 32 * gpu.NVars * k + 32
        assert int.from_bytes(wg[off:off + 32], "little") == h


@pytest.mark.gpu
def test_withdraw_tampered_balance_is_rejected_on_gpu(built):
    from circuits_b200.synth import withdraw_inputs
    from circuits_b200.witness_calculator import WitnessError
    inp, _ = withdraw_inputs(32, 3)[2]
    bad = dict(inp)
    bad["balance"] = inp["balance"] + 1
    with pytest.raises(WitnessError, match="Constraint doesn't match 1 != 0"):
        _wc("gpu").calculateWitness(bad, True)
