"""ONE context over several GPUs of a box (`hzw_params.n_gpus`): `hzw_calculate` must land one whole witness in
the caller's buffer, byte for byte the witness of a single GPU (`./circuit input.json witness.wtns` writes one
file: /root/reference/tools/helpers/actions.js:139-143).  The exchange between the two phases is done inside the
library (pack kernels storing into every peer); nothing here calls a collective.

Needs >= 2 GPUs: skipped on a one-GPU box (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py`)."""
import numpy as np
import pytest

from conftest import MIXED32, _cuda_devices, cpu_port_calculator

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_cuda_devices() < 2, reason="needs two GPUs on one box")]


def _first_diff(a, b):
    a32 = a.view(np.uint64).reshape(-1, 4)
    b32 = b.view(np.uint64).reshape(-1, 4)
    bad = np.nonzero((a32 != b32).any(axis=1))[0]
    return None if bad.size == 0 else (int(bad[0]), int(bad.size))


def _packed(prm, inp):
    cpu = cpu_port_calculator(*prm)
    packed = np.frombuffer(bytes(cpu.packInput(inp)), dtype=np.uint8).copy()
    cpu.close()
    return packed


def _check(prm, packed, expect_hash, ref=None, worlds=None):
    from circuits_b200.witness_calculator import WitnessCalculator
    if ref is None:
        cpu = cpu_port_calculator(*prm)
        ref = np.empty(32 * cpu.NVars, dtype=np.uint8)
        cpu.calculateBinWitness(packed, False, out=ref)
        cpu.close()
    assert int.from_bytes(bytes(ref[32:64]), "little") == expect_hash
    out = np.empty_like(ref)
    for n in worlds or [w for w in (2, 3, 4, 8) if w <= _cuda_devices() and w <= prm[0]]:
        wc = WitnessCalculator(*prm, n_gpus=n)
        for rep in range(2):
            out[:] = 0xA5
            wc.calculateBinWitness(packed, False, out=out)
            assert _first_diff(out, ref) is None, ("n_gpus=%d host buffer, run %d" % (n, rep), _first_diff(out, ref))
        st = wc.lastStats()
        assert st["gpu_launches"] > 20 * n
        # resident: every rank keeps its ranges on its own device; hzw_witness_read gathers them
        out[:] = 0x5A
        ms = wc.calculateResident(packed, False)
        assert ms > 0
        wc.readWitness(out=out)
        assert _first_diff(out, ref) is None, ("n_gpus=%d resident" % n, _first_diff(out, ref))
        wc.close()
    return ref


def test_mixed_batch_every_tx_type(built, mixed32):
    _check(MIXED32, _packed(MIXED32, mixed32.getInput()), mixed32.getHashInputs())


def test_uneven_split_and_dense_tree(built):
    from circuits_b200.synth import bench_batch
    prm = (26, 10, 5, 4)          # 26 slots do not divide by 3 / 4 / 8 (the pre-image stays a whole number of bytes)
    bb = bench_batch(*prm, n_accounts=700)
    _check(prm, _packed(prm, bb.getInput()), bb.getHashInputs())


def test_constraint_violation_is_reported_from_the_owning_rank(built, small_bb):
    from circuits_b200.witness_calculator import WitnessCalculator, WitnessError
    prm = (3, 16, 2, 2)
    inp = dict(small_bb.getInput())
    inp["imStateRoot"] = list(inp["imStateRoot"])
    inp["imStateRoot"][1] = int(inp["imStateRoot"][1]) + 1         # checked in the last slot: rank 1 of 2
    wc = WitnessCalculator(*prm, n_gpus=2)
    with pytest.raises(WitnessError, match="Constraint doesn't match"):
        wc.calculateBinWitness(inp, True)
    # and the context is still usable
    w = wc.calculateBinWitness(small_bb.getInput(), True)
    assert int.from_bytes(bytes(w[32:64]), "little") == small_bb.getHashInputs()
    wc.close()


@pytest.mark.parametrize("cfg", ["rollup-main-344", "rollup-main-2048"])
def test_full_batch_on_all_gpus_equals_one_gpu(built, cfg):
    """The benchmark sizes: the witness of n_gpus = all equals the witness of one GPU (itself compared with the CPU
    port byte for byte in tests/test_full_batch.py)."""
    import bench
    from circuits_b200.witness_calculator import WitnessCalculator
    prm = bench.CONFIGS[cfg]
    inp, expect_hash = bench.make_input(prm)
    packed = _packed(prm, inp)
    one = WitnessCalculator(*prm)
    ref = np.empty(32 * one.NVars, dtype=np.uint8)
    one.calculateBinWitness(packed, False, out=ref)
    one.close()
    n = max(w for w in (2, 4, 8) if w <= _cuda_devices())
    _check(prm, packed, expect_hash, ref=ref, worlds=[n])
