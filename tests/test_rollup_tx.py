"""RollupTx(nLevels, maxFeeTx) as a circuit of its own (HZW_CIRCUIT_ROLLUP_TX; config 2 of BASELINE.json): the
reference's test/rollup-tx.test.js:56-920 compiles `component main = RollupTx(16, 16)` and feeds it the per-slot
inputs a batch produces for every transaction type.  Here the inputs of every slot of the mixed batch (all
transaction types, nLevels = 32) are lifted from the RollupMain witness of the oracle, and the standalone witness
of the CPU port / the GPU is compared signal by signal with the oracle's standalone RollupTx, names included."""
import pytest

from conftest import MIXED32, cpu_port_calculator


def _slot_inputs(w, sym, slot, nLevels, maxFeeTx):
    """The RollupTx inputs of `slot`, read by name from a RollupMain witness."""
    from oracle.hermez import ROLLUP_TX_INPUTS
    out = {}
    for nm in ROLLUP_TX_INPUTS:
        base = "main.rollupTx[%d].%s" % (slot, nm)
        if base in sym:
            out[nm] = w[sym[base]]
        else:
            vals, k = [], 0
            while "%s[%d]" % (base, k) in sym:
                vals.append(w[sym["%s[%d]" % (base, k)]])
                k += 1
            assert vals, base
            out[nm] = vals
    return out


@pytest.fixture(scope="module")
def slots(mixed32):
    from oracle import hermez as hz
    from oracle.circuit import symbols
    nTx, nLevels, _, maxFeeTx = MIXED32
    w, main, _ = hz.rollup_main_witness(*MIXED32, mixed32.getInput())
    sym = symbols(main)
    return [(_slot_inputs(w, sym, i, nLevels, maxFeeTx), w[sym["main.rollupTx[%d].newStateRoot" % i]]) for i in range(nTx)]


def _check(wc, slots, which):
    from oracle import hermez as hz
    from oracle.circuit import symbols
    nLevels, maxFeeTx = MIXED32[1], MIXED32[3]
    names = None
    for i in which:
        inp, new_root = slots[i]
        ref, main, _ = hz.rollup_tx_witness(nLevels, maxFeeTx, inp)
        w = wc.calculateWitness(inp, True)
        assert len(w) == len(ref) == wc.NVars
        bad = [k for k in range(len(w)) if w[k] != ref[k]]
        assert not bad, "slot %d: %d entries differ, first %s" % (i, len(bad), bad[:5])
        assert wc.getSignal(w, "main.newStateRoot") == new_root
        if names is None:
            names = symbols(main)
            assert wc.symbols("main") == names


def test_rollup_tx_cpu_port_equals_oracle(built, slots):
    wc = cpu_port_calculator(1, MIXED32[1], 0, MIXED32[3], circuit="rollup-tx")
    _check(wc, slots, [0, 4, 7, 11])


def test_rollup_tx_rejects_a_wrong_state(built, slots):
    from circuits_b200.witness_calculator import WitnessError
    wc = cpu_port_calculator(1, MIXED32[1], 0, MIXED32[3], circuit="rollup-tx")
    inp = dict(slots[8][0])
    inp["oldStateRoot"] = inp["oldStateRoot"] + 1
    with pytest.raises(WitnessError, match="Constraint doesn't match"):
        wc.calculateWitness(inp, True)


@pytest.mark.gpu
def test_rollup_tx_gpu_equals_oracle_every_tx_type(built, slots):
    from circuits_b200.witness_calculator import WitnessCalculator
    wc = WitnessCalculator(1, MIXED32[1], 0, MIXED32[3], circuit="rollup-tx")
    _check(wc, slots, range(len(slots)))


@pytest.mark.gpu
def test_rollup_tx_gpu_batch_equals_cpu_port(built, slots):
    from circuits_b200.witness_calculator import WitnessCalculator
    gpu = WitnessCalculator(1, MIXED32[1], 0, MIXED32[3], circuit="rollup-tx")
    cpu = cpu_port_calculator(1, MIXED32[1], 0, MIXED32[3], circuit="rollup-tx")
    packed = b"".join(bytes(gpu.packInput(inp)) for inp, _ in slots)
    assert bytes(gpu.calculateBatch(len(slots), packed, True)) == bytes(cpu.calculateBatch(len(slots), packed, True))
