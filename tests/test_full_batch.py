"""Whole-witness parity at the sizes the benchmark is quoted on: rollup-main(344, 32, 256, 64) and
rollup-main(2048, 32, 256, 64) -- 8.1 GB -- byte for byte against the CPU port (pinned to the oracle at this tree
depth by tests/test_parity_nlevels32.py), over every way the witness leaves the device:

  * the end-to-end path of hzw_calculate (pieces copied while the rest is computed, image levels rebuilt in the
    caller's buffer by host threads),
  * the same with HZW_NO_HOLES (every entry crosses the link),
  * the device-resident witness read back with hzw_witness_read,
  * two shards on one GPU, assembled.

Plus the corner cases of the image-level copies: a dense tree (no level of any processor is an image), blocks of
processors whose image counts differ, non-uniform groups of slots."""
import os

import numpy as np
import pytest

from conftest import cpu_port_calculator

pytestmark = pytest.mark.gpu


def _first_diff(a, b):
    a32 = a.view(np.uint64).reshape(-1, 4)
    b32 = b.view(np.uint64).reshape(-1, 4)
    bad = np.nonzero((a32 != b32).any(axis=1))[0]
    return None if bad.size == 0 else (int(bad[0]), int(bad.size))


def _reference(prm, packed):
    cpu = cpu_port_calculator(*prm)
    ref = np.empty(32 * cpu.NVars, dtype=np.uint8)
    cpu.calculateBinWitness(packed, False, out=ref)
    cpu.close()
    return ref


def _check_all_paths(prm, packed, expect_hash, shards=True):
    import torch
    from circuits_b200.witness_calculator import ShardedWitnessCalculator, WitnessCalculator
    ref = _reference(prm, packed)
    assert int.from_bytes(bytes(ref[32:64]), "little") == expect_hash
    out = np.empty_like(ref)
    wc = WitnessCalculator(*prm)
    # 1. end-to-end path (twice: the second run re-uses hole records and scratch of the first)
    for rep in range(2):
        out[:] = 0xA5
        wc.calculateBinWitness(packed, False, out=out)
        assert _first_diff(out, ref) is None, ("e2e path, run %d" % rep, _first_diff(out, ref))
    d2h = wc.lastStats()["d2h_bytes"]
    # 2. resident witness
    out[:] = 0x5A
    wc.calculateResident(packed, False)
    wc.readWitness(out=out)
    assert _first_diff(out, ref) is None, ("resident path", _first_diff(out, ref))
    wc.close()
    # 3. every entry copied
    os.environ["HZW_NO_HOLES"] = "1"
    try:
        wc = WitnessCalculator(*prm)
        out[:] = 0xA5
        wc.calculateBinWitness(packed, False, out=out)
        assert _first_diff(out, ref) is None, ("HZW_NO_HOLES path", _first_diff(out, ref))
        assert wc.lastStats()["d2h_bytes"] >= ref.size
        wc.close()
    finally:
        del os.environ["HZW_NO_HOLES"]
    # 4. two shards on this GPU, the all-gather done by hand
    if shards:
        ranks = [ShardedWitnessCalculator(*prm, rank=r, world=2) for r in range(2)]
        for r in ranks:
            r.phase1(packed, False)
        recs = torch.cat([r._rec for r in ranks])
        out[:] = 0x5A
        covered = 0
        for r in ranks:
            r._recs.copy_(recs)
            torch.cuda.synchronize()
            buf = np.empty(32 * r.shardNVars, dtype=np.uint8)
            r.phase2(False, out=buf)
            off = 0
            for idx, cnt in r.pieces():
                out[32 * idx:32 * (idx + cnt)] = buf[off:off + 32 * cnt]
                off += 32 * cnt
                covered += cnt
            r.close()
        assert covered * 32 == ref.size
        assert _first_diff(out, ref) is None, ("sharded path", _first_diff(out, ref))
    return d2h, ref.size


@pytest.mark.parametrize("cfg", ["rollup-main-344", "rollup-main-2048"])
def test_full_witness_equals_cpu_port(built, cfg):
    import bench
    prm = bench.CONFIGS[cfg]
    inp, expect_hash = bench.make_input(prm)
    cpu = cpu_port_calculator(*prm)
    packed = np.frombuffer(bytes(cpu.packInput(inp)), dtype=np.uint8).copy()
    cpu.close()
    d2h, total = _check_all_paths(prm, packed, expect_hash)
    assert d2h < total          # sparse tree: the image levels were rebuilt on the host, not copied


def test_dense_tree_and_mixed_image_counts(built):
    """nLevels = 10 with 700 accounts: leaves sit at depth 9-10 of 11, so most processors have NO image level and
    the blocks of k_smt_levels mix processors with zero and one; the holes path must copy everything else intact."""
    from circuits_b200.synth import bench_batch
    prm = (24, 10, 4, 4)
    bb = bench_batch(*prm, n_accounts=700)
    cpu = cpu_port_calculator(*prm)
    packed = np.frombuffer(bytes(cpu.packInput(bb.getInput())), dtype=np.uint8).copy()
    cpu.close()
    _check_all_paths(prm, packed, bb.getHashInputs())


@pytest.mark.parametrize("prm, accounts", [((45, 24, 11, 5), 40), ((9, 16, 3, 1), 16), ((70, 32, 70, 3), 24)])
def test_ragged_sizes(built, prm, accounts):
    """Sizes that fit none of the kernels' units: slot counts that are no multiple of 8 (ladder and tail blocks,
    groups of the pitched copies), 16 (level blocks) or 32 (warps of the per-slot kernels), an odd number of fee txs, a
    batch of L1 slots only (every key off the curve), depths other than 32 (multiples of 8: the host-side batch builder
    assembles the hash pre-image by bytes); the SHA block range ends inside a word of the packed transfer."""
    from circuits_b200.synth import bench_batch
    bb = bench_batch(*prm, n_accounts=accounts)
    cpu = cpu_port_calculator(*prm)
    packed = np.frombuffer(bytes(cpu.packInput(bb.getInput())), dtype=np.uint8).copy()
    cpu.close()
    _check_all_paths(prm, packed, bb.getHashInputs())


def test_non_uniform_slot_groups(built, mixed32):
    """Every transaction type within one group of eight slots: processors of NOP / exit / insert slots have
    different image counts and offsets, so the pitched copies cannot be used and the copies go piece by piece."""
    from conftest import MIXED32
    cpu = cpu_port_calculator(*MIXED32)
    packed = np.frombuffer(bytes(cpu.packInput(mixed32.getInput())), dtype=np.uint8).copy()
    cpu.close()
    _check_all_paths(MIXED32, packed, mixed32.getHashInputs())


def test_world1_sharded_context_after_another_batch(built, small_bb):
    """A sharded context with world = 1 (what ShardedWitnessCalculator.exchange() handles explicitly) must not
    read hole records of an earlier call: it copies its pieces whole."""
    from circuits_b200.witness_calculator import ShardedWitnessCalculator, WitnessCalculator
    prm = (3, 16, 2, 2)
    inp = small_bb.getInput()
    other = dict(inp)
    other["s"] = list(inp["s"]); other["s"][0] = int(inp["s"][0]) + 7
    whole = WitnessCalculator(*prm)
    want = bytes(whole.calculateBinWitness(whole.packInput(inp), False))
    wc = ShardedWitnessCalculator(*prm, rank=0, world=1)
    wc.calculateShard(wc.packInput(other), False)
    buf, _ = wc.calculateShard(wc.packInput(inp), False)
    assert bytes(ShardedWitnessCalculator.assemble(wc.NVars, [(wc.pieces(), bytes(buf))])) == want


def test_streamed_wtns_file_equals_the_host_buffer(built, tmp_path):
    """hzw_calculate_to_wtns at rollup-main(344, 32, 256, 64): the 1.4 GB file written chunk by chunk from device memory
    holds the header of snarkjs' writer and, behind it, byte for byte what hzw_calculate puts in a host buffer."""
    import bench
    from circuits_b200.witness_calculator import WitnessCalculator
    prm = bench.CONFIGS["rollup-main-344"]
    inp, expect_hash = bench.make_input(prm)
    wc = WitnessCalculator(*prm)
    packed = np.frombuffer(bytes(wc.packInput(inp)), dtype=np.uint8).copy()
    ref = np.empty(32 * wc.NVars, dtype=np.uint8)
    wc.calculateBinWitness(packed, False, out=ref)
    path = str(tmp_path / "witness.wtns")
    wc.calculateToWtns(packed, path, False)
    assert os.path.getsize(path) == 76 + ref.size
    body = np.memmap(path, dtype=np.uint8, mode="r", offset=76)
    assert _first_diff(np.ascontiguousarray(body), ref) is None
    head = open(path, "rb").read(76)
    assert head[:4] == b"wtns" and int.from_bytes(head[60:64], "little") == wc.NVars
    assert int.from_bytes(bytes(body[32:64]), "little") == expect_hash
    wc.close()
