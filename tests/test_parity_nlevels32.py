"""Parity at the tree depth of the BASELINE configurations (nLevels = 32, SMT depth 33), against the independent
Python oracle DIRECTLY -- not through the CPU port, which shares the template source with the device code:

  * RollupMain(12, 32, 7, 3) with every transaction type of /root/reference/test/rollup-main.test.js:65-878 and
    rollup-main-L1.test.js (deposits, transfers, exits, force txs, a nullified L1, NOP padding, two fee txs);
  * Withdraw(32) (/root/reference/test/withdraw.test.js:39-171).

The CPU half pins the CPU port to the oracle at this depth too, so that the byte-for-byte GPU = CPU-port checks
at the full batch sizes (tests/test_full_batch.py) rest on a chain of equalities with no gap."""
import pytest

from conftest import MIXED32, cpu_port_calculator


def _first_diff(a, b):
    return next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), None)


def test_cpu_port_nlevels32_equals_oracle(built, mixed32, mixed32_oracle):
    wc = cpu_port_calculator(*MIXED32)
    w = wc.calculateWitness(mixed32.getInput(), True)
    assert len(w) == len(mixed32_oracle)
    assert w == mixed32_oracle, "first differing index %s" % _first_diff(w, mixed32_oracle)
    assert w[1] == mixed32.getHashInputs()


def test_withdraw32_cpu_port_equals_oracle(built):
    from circuits_b200.synth import withdraw_inputs
    from oracle import hermez as hz
    wc = cpu_port_calculator(1, 32, 0, 1, circuit="withdraw")
    for inp, h in withdraw_inputs(32, 3):
        w_ref, _, _ = hz.withdraw_witness(32, inp)
        assert wc.calculateWitness(inp, True) == w_ref
        assert w_ref[1] == h


@pytest.mark.gpu
def test_gpu_nlevels32_equals_oracle_host_and_resident_paths(built, mixed32, mixed32_oracle):
    """The same witness three ways: through hzw_calculate into a host buffer (the end-to-end path with its copies
    around the rebuilt image levels: 12 slots >= 8 take it), left resident in HBM and read back with
    hzw_witness_read, and with the sanity check on."""
    from circuits_b200.witness_calculator import WitnessCalculator
    wc = WitnessCalculator(*MIXED32)
    packed = wc.packInput(mixed32.getInput())
    ref = b"".join(x.to_bytes(32, "little") for x in mixed32_oracle)
    host = bytes(wc.calculateBinWitness(packed, True))
    assert len(host) == len(ref)
    if host != ref:
        raise AssertionError("host path: first differing witness index %d" %
                             next(i for i in range(wc.NVars) if host[32 * i:32 * i + 32] != ref[32 * i:32 * i + 32]))
    st = wc.lastStats()
    assert st["gpu_launches"] >= 15 and st["d2h_bytes"] < len(ref)        # image levels did not cross the link
    wc.calculateResident(packed, False)
    res = bytes(wc.readWitness())
    if res != ref:
        raise AssertionError("resident path: first differing witness index %d" %
                             next(i for i in range(wc.NVars) if res[32 * i:32 * i + 32] != ref[32 * i:32 * i + 32]))
    ptr, dev, _ = wc.witnessDevice()
    assert ptr and dev == 0
    idx = wc.signalIndex("main.rollupTx[7].processor1.newRoot")
    assert int.from_bytes(bytes(wc.readWitness(idx, 1)), "little") == mixed32_oracle[idx]


@pytest.mark.gpu
def test_gpu_withdraw32_equals_oracle(built):
    from circuits_b200.synth import withdraw_inputs
    from circuits_b200.witness_calculator import WitnessCalculator
    from oracle import hermez as hz
    cases = withdraw_inputs(32, 5)
    gpu = WitnessCalculator(1, 32, 0, 1, circuit="withdraw")
    packed = b"".join(bytes(gpu.packInput(inp)) for inp, _ in cases)
    wg = bytes(gpu.calculateBatch(len(cases), packed, True))
    for k, (inp, h) in enumerate(cases[:3]):
        w_ref, _, _ = hz.withdraw_witness(32, inp)
        got = wg[32 * gpu.NVars * k:32 * gpu.NVars * (k + 1)]
        assert got == b"".join(x.to_bytes(32, "little") for x in w_ref), "instance %d" % k
        assert w_ref[1] == h
