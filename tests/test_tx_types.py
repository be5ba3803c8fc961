"""Every transaction type in one batch (the cases of /root/reference/test/rollup-main.test.js and
rollup-main-L1.test.js): the CPU port against the oracle signal by signal, the GPU against the CPU port."""
import pytest

from conftest import cpu_port_calculator

PRM = (12, 16, 7, 3)


@pytest.fixture(scope="module")
def mixed():
    from circuits_b200.synth import mixed_batch
    return mixed_batch(PRM[1], PRM[0], PRM[2], PRM[3])


def test_mixed_batch_cpu_port_equals_oracle(built, mixed):
    from oracle import hermez as hz
    inp = mixed.getInput()
    assert sum(inp["onChain"]) == 7 and inp["newExit"].count(1) == 2       # two exit leaves inserted, one of them updated again
    w_ref, _, _ = hz.rollup_main_witness(*PRM, inp)
    wc = cpu_port_calculator(*PRM)
    w = wc.calculateWitness(inp, True)
    assert w == w_ref
    assert w[1] == mixed.getHashInputs()


@pytest.mark.gpu
def test_mixed_batch_gpu_equals_cpu_port(built, mixed):
    from circuits_b200.witness_calculator import WitnessCalculator
    gpu, cpu = WitnessCalculator(*PRM), cpu_port_calculator(*PRM)
    packed = gpu.packInput(mixed.getInput())
    wg = bytes(gpu.calculateBinWitness(packed, True))
    assert wg == bytes(cpu.calculateBinWitness(packed, True))
    assert int.from_bytes(wg[32:64], "little") == mixed.getHashInputs()


def _invalid_key_batch():
    """An account created with a BabyJubJub y that has no x (as /root/reference/test/rollup-main-L1.test.js:106-121),
    then an L2 transfer FROM it: the verifier's Ax = sqrt(non-residue) = 0 path of pointbits.circom.  Such a tx
    cannot carry a valid signature, so it is only evaluated without sanity checks."""
    from circuits_b200 import float40
    from circuits_b200.field import P
    from circuits_b200.rollupdb import RollupDB, Account
    from circuits_b200.synth import _deposit
    a, d = 168700, 168696
    y = 5
    while True:
        n = (1 - y * y) * pow(a - d * y * y, P - 2, P) % P
        if n != 0 and pow(n, (P - 1) // 2, P) != 1:
            break
        y += 1
    bits = [(y >> i) & 1 for i in range(256)]
    db = RollupDB(chain_id=0)
    good = Account(3)
    bb = db.buildBatch(3, 16, 2, 2)
    bb.addTx(dict(fromIdx=0, loadAmountF=float40.fix2float(900), tokenID=1, fromBjjCompressed=bits, fromEthAddr=77, toIdx=0, onChain=1))
    bb.addTx(_deposit(good, 100))
    bb.build()
    db.consolidate(bb)
    bb = db.buildBatch(3, 16, 2, 2)
    tx = dict(fromIdx=256, toIdx=257, tokenID=1, amountF=float40.fix2float(10), nonce=0, userFee=0, onChain=0,
              s=12345, r8x=6789, r8y=1011)
    bb.addTx(tx)
    bb.build()
    return bb


def test_invalid_key_sender_cpu_port_equals_oracle(built):
    from oracle import hermez as hz
    bb = _invalid_key_batch()
    inp = bb.getInput()
    w_ref, _, _ = hz.rollup_main_witness(3, 16, 2, 2, inp, sanity=False)
    wc = cpu_port_calculator(3, 16, 2, 2)
    assert wc.calculateWitness(inp, False) == w_ref


@pytest.mark.gpu
def test_invalid_key_sender_gpu_equals_cpu_port(built):
    from circuits_b200.witness_calculator import WitnessCalculator
    bb = _invalid_key_batch()
    gpu, cpu = WitnessCalculator(3, 16, 2, 2), cpu_port_calculator(3, 16, 2, 2)
    packed = gpu.packInput(bb.getInput())
    assert bytes(gpu.calculateBinWitness(packed, False)) == bytes(cpu.calculateBinWitness(packed, False))
