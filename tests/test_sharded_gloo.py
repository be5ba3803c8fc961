"""The multi-GPU protocol on CPU: world_size 2 and 3 `gloo` process groups, each rank driving the CPU
port (checker library, same C ABI and the same two-phase shard protocol as libhzwitness.so).  The
pieces of all ranks, put together, must be the single-process witness byte for byte."""
import os
import sys

import pytest

from conftest import CPU_PORT, ROOT

PRM = (6, 16, 2, 3)


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import pickle
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from circuits_b200.witness_calculator import ShardedWitnessCalculator
    from circuits_b200.synth import bench_batch
    bb = bench_batch(*PRM, n_accounts=6)
    wc = ShardedWitnessCalculator(*PRM, rank=rank, world=world, lib_path=CPU_PORT, record_device="cpu")
    packed = wc.packInput(bb.getInput())
    shard, _ = wc.calculateShard(packed, True)
    assert len(shard) == 32 * wc.shardNVars
    with open(os.path.join(tmpdir, "shard%d.pkl" % rank), "wb") as f:
        pickle.dump((wc.pieces(), bytes(shard), wc.NVars), f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_witness_equals_single_process(built, world, tmp_path):
    import pickle
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from circuits_b200.witness_calculator import ShardedWitnessCalculator, WitnessCalculator
    from circuits_b200.synth import bench_batch
    shards, nvars = [], None
    covered = 0
    for r in range(world):
        with open(os.path.join(str(tmp_path), "shard%d.pkl" % r), "rb") as f:
            pcs, data, nvars = pickle.load(f)
        shards.append((pcs, data))
        covered += sum(c for _, c in pcs)
    assert covered == nvars                      # the pieces tile the witness exactly
    full = ShardedWitnessCalculator.assemble(nvars, shards)
    bb = bench_batch(*PRM, n_accounts=6)
    ref = WitnessCalculator(*PRM, lib_path=CPU_PORT)
    want = bytes(ref.calculateBinWitness(ref.packInput(bb.getInput()), True))
    assert bytes(full) == want
    assert int.from_bytes(want[32:64], "little") == bb.getHashInputs()
