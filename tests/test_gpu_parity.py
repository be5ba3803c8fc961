"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle."""
import hashlib

import pytest

from conftest import cpu_port_calculator

pytestmark = pytest.mark.gpu


def _cuda(*params):
    from circuits_b200.witness_calculator import WitnessCalculator
    return WitnessCalculator(*params)


def test_small_batch_equals_oracle_signal_by_signal(built, small_bb, small_oracle):
    w_ref, sym_ref = small_oracle
    wc = _cuda(3, 16, 2, 2)
    assert wc.NVars == len(w_ref)
    w = wc.calculateWitness(small_bb.getInput(), True)
    bad = [i for i in range(len(w)) if w[i] != w_ref[i]]
    assert not bad, (len(bad), bad[:8])
    assert wc.symbols("main.rollupTx[1].processor1.levels[3]") == {
        k: v for k, v in sym_ref.items() if k.startswith("main.rollupTx[1].processor1.levels[3]")}
    assert wc.lastStats()["gpu_launches"] > 0


def test_nlevels32_batch_equals_cpu_port_and_hashlib(built):
    """24 slots at nLevels=32 (8 L1 deposits + 16 signed transfers): byte-exact against the CPU
    port (itself pinned to the oracle), and hashGlobalInputs against hashlib's SHA-256."""
    from circuits_b200.synth import bench_batch
    prm = (24, 32, 8, 4)
    bb = bench_batch(*prm, n_accounts=16)
    wc = _cuda(*prm)
    ref = cpu_port_calculator(*prm)
    packed = wc.packInput(bb.getInput())
    wg = bytes(wc.calculateBinWitness(packed, True))
    wr = bytes(ref.calculateBinWitness(packed, True))
    assert len(wg) == len(wr) == 32 * wc.NVars
    if wg != wr:
        first = next(i for i in range(wc.NVars) if wg[32 * i:32 * i + 32] != wr[32 * i:32 * i + 32])
        raise AssertionError("first differing witness index %d" % first)
    assert int.from_bytes(wg[32:64], "little") == bb.getHashInputs()


def test_all_nop_batch(built):
    """Empty `bb.build()` (tools/helpers/actions.js:84-91): every slot is NOP padding."""
    from circuits_b200.rollupdb import RollupDB
    prm = (4, 8, 2, 2)
    bb = RollupDB(chain_id=0).buildBatch(*prm)
    bb.build()
    wc = _cuda(*prm)
    ref = cpu_port_calculator(*prm)
    packed = wc.packInput(bb.getInput())
    wg = bytes(wc.calculateBinWitness(packed, True))
    assert wg == bytes(ref.calculateBinWitness(packed, True))
    assert int.from_bytes(wg[32:64], "little") == bb.getHashInputs()


def test_constraint_violation_on_gpu(built, small_bb):
    from circuits_b200.witness_calculator import WitnessError
    wc = _cuda(3, 16, 2, 2)
    inp = dict(small_bb.getInput())
    inp["balance1"] = list(inp["balance1"])
    inp["balance1"][0] += 1
    with pytest.raises(WitnessError, match="Constraint doesn't match 1 != 0") as e:
        wc.calculateWitness(inp, True)
    assert "main.rollupTx[0].processor1.checkOldInput" in str(e.value)
    assert len(wc.calculateWitness(inp, False)) == wc.NVars


def test_tampered_signature_on_gpu(built, small_bb):
    """A wrong `s` walks the whole signature path (ladder walk, k_ed_fill, prefilled constants) with consistent
    values and only fails at the final comparison: same error as the CPU port with the sanity check on, the same
    witness byte for byte with it off."""
    from circuits_b200.witness_calculator import WitnessError
    wc, ref = _cuda(3, 16, 2, 2), cpu_port_calculator(3, 16, 2, 2)
    inp = dict(small_bb.getInput())
    inp["s"] = list(inp["s"])
    inp["s"][0] = int(inp["s"][0]) + 1
    with pytest.raises(WitnessError, match="Constraint doesn't match") as eg:
        wc.calculateWitness(inp, True)
    with pytest.raises(WitnessError, match="Constraint doesn't match") as ec:
        ref.calculateWitness(inp, True)
    assert "main.rollupTx[0].sigVerifier.eqCheck" in str(eg.value)
    assert str(eg.value) == str(ec.value)
    packed = wc.packInput(inp)
    assert bytes(wc.calculateBinWitness(packed, False)) == bytes(ref.calculateBinWitness(packed, False))


def test_repeated_batches_on_one_context(built, small_bb):
    """Scratch areas, job records and level flags are reused between batches: a context gives the same witness
    for the same input after it has evaluated a different one."""
    wc = _cuda(3, 16, 2, 2)
    inp = small_bb.getInput()
    first = bytes(wc.calculateBinWitness(wc.packInput(inp), True))
    other = dict(inp)
    other["s"] = list(inp["s"]); other["s"][0] = int(inp["s"][0]) + 7
    wc.calculateBinWitness(wc.packInput(other), False)
    assert bytes(wc.calculateBinWitness(wc.packInput(inp), True)) == first


def test_two_shards_on_one_gpu_equal_the_whole(built):
    """The CUDA side of the multi-GPU protocol without a second GPU: two contexts (rank 0 and 1 of 2)
    on this device, their exchange records concatenated by hand where NCCL's all-gather would be.
    The assembled pieces must equal the single-context witness byte for byte."""
    import torch
    from circuits_b200.synth import bench_batch
    from circuits_b200.witness_calculator import ShardedWitnessCalculator
    prm = (24, 32, 8, 4)
    bb = bench_batch(*prm, n_accounts=16)
    whole = _cuda(*prm)
    packed = whole.packInput(bb.getInput())
    want = bytes(whole.calculateBinWitness(packed, True))
    ranks = [ShardedWitnessCalculator(*prm, rank=r, world=2) for r in range(2)]
    for wc in ranks:
        wc.phase1(packed, True)
    recs = torch.cat([wc._rec for wc in ranks])
    shards = []
    for wc in ranks:
        wc._recs.copy_(recs)
        torch.cuda.synchronize()
        buf, _ = wc.phase2(True)
        shards.append((wc.pieces(), bytes(buf)))
    assert sum(c for pcs, _ in shards for _, c in pcs) == whole.NVars
    assert bytes(ShardedWitnessCalculator.assemble(whole.NVars, shards)) == want


@pytest.mark.gpu
@pytest.mark.parametrize("world, prm", [(3, (13, 16, 5, 2)), (5, (11, 16, 4, 3))])
def test_uneven_shards_with_different_owners_of_the_closing_values(built, world, prm, monkeypatch):
    """Shards of unequal size where the last slot and the last fee tx belong to DIFFERENT ranks (and one rank has no fee
    tx at all): the early exchange walks them on their owners (hz_early_exchange: n_s = 1, n_f = 0 on one rank, the other
    way round on another), every rank packs its record ahead of its merkle levels."""
    import torch
    from circuits_b200.synth import bench_batch
    from circuits_b200.witness_calculator import ShardedWitnessCalculator
    monkeypatch.setenv("HZW_EARLY_HASH", "1")         # the early paths are for long SHA walks: forced on at this size
    bb = bench_batch(*prm, n_accounts=16)
    whole = _cuda(*prm)
    packed = whole.packInput(bb.getInput())
    want = bytes(whole.calculateBinWitness(packed, True))
    assert int.from_bytes(want[32:64], "little") == bb.getHashInputs()
    whole.calculateResident(packed, True)             # the single-GPU early walk at this size too
    got = bytearray(len(want)); whole.readWitness(out=got)
    assert bytes(got) == want
    ranks = [ShardedWitnessCalculator(*prm, rank=r, world=world) for r in range(world)]
    for wc in ranks:
        wc.phase1(packed, True)
    recs = torch.cat([wc._rec for wc in ranks])
    shards = []
    for wc in ranks:
        wc._recs.copy_(recs)
        torch.cuda.synchronize()
        buf, _ = wc.phase2(True)
        shards.append((wc.pieces(), bytes(buf)))
    assert sum(c for pcs, _ in shards for _, c in pcs) == whole.NVars
    assert bytes(ShardedWitnessCalculator.assemble(whole.NVars, shards)) == want


def test_copy_paths_without_rebuilt_images(built, monkeypatch):
    """With many ranks per host the image levels are copied like everything else (a rank has no threads to
    rebuild them with): the same witnesses from the end-to-end path and from two shards with HZW_NO_HOLES set."""
    monkeypatch.setenv("HZW_NO_HOLES", "1")
    test_nlevels32_batch_equals_cpu_port_and_hashlib(built)
    test_two_shards_on_one_gpu_equal_the_whole(built)


def test_poseidon_lane_kernel_known_answers(built):
    """The lane-parallel permutation on its own: circomlib's published vector poseidon([1, 2]) and
    seeded random pairs against the oracle's big-integer Poseidon."""
    import random
    from circuits_b200.witness_calculator import poseidon_pairs
    from oracle.poseidon_constants import poseidon_hash
    from oracle.field import P
    rng = random.Random(7)
    quads = [(1, 2, 3, 4)] + [tuple(rng.randrange(P) for _ in range(4)) for _ in range(37)]
    quads += [(0, 0, P - 1, P - 1)]
    for lanes in (False, True):          # transposed kernel, lane-group kernel
        got, _ = poseidon_pairs(quads, lanes=lanes)
        assert got[0][0] == 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a
        for (a, b, c, d), (x, y) in zip(quads, got):
            assert x == poseidon_hash([a, b]) and y == poseidon_hash([c, d])


@pytest.mark.gpu
@pytest.mark.parametrize("t", [4, 5, 6, 7])
def test_transposed_poseidon_of_every_width(built, t):
    """The transposed permutation at the other widths of the path (leaf, state, verifier and message hashes) against the
    oracle's big-integer Poseidon: seeded tuples, the all-zero and all-(p - 1) inputs, a count that is no multiple of 32."""
    import random
    from circuits_b200.witness_calculator import poseidon_width
    from oracle.poseidon_constants import poseidon_hash
    from oracle.field import P
    rng = random.Random(40 + t)
    tuples = [tuple(range(1, t))] + [tuple(rng.randrange(P) for _ in range(t - 1)) for _ in range(41)]
    tuples += [(0,) * (t - 1), (P - 1,) * (t - 1)]
    got, _, _ = poseidon_width(t, tuples)
    for tup, d in zip(tuples, got):
        assert d == poseidon_hash(list(tup)), tup


@pytest.mark.gpu
@pytest.mark.parametrize("t", [3, 4, 5, 6, 7])
def test_digest_only_poseidon(built, t):
    """hzw_poseidon_hash (what the batch builder's trees hash with): digests only, every width."""
    import random
    from circuits_b200.witness_calculator import poseidon_hash_batch
    from oracle.poseidon_constants import poseidon_hash
    from oracle.field import P
    rng = random.Random(90 + t)
    tuples = [tuple(rng.randrange(P) for _ in range(t - 1)) for _ in range(70)] + [(0,) * (t - 1)]
    assert poseidon_hash_batch(t, tuples) == [poseidon_hash(list(x)) for x in tuples]
