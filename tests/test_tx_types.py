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
