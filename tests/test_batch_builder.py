"""The batch builder's trees with every hash deferred and evaluated in waves (circuits_b200/batchtree.py) against the
host mirror that hashes node by node (circuits_b200/smt.py, after circomlib smt.js as /root/reference/tools/
generate-input.js:61-109 drives it): the circuit inputs of a batch must be identical -- siblings, old keys / values of
insertions that push a leaf down, intermediate and final roots of the state and exit trees.

On the CPU the waves are hashed by the host's Poseidon (the logic under test is the deferred tree); on the GPU by the
transposed Poseidon kernels of libhzwitness.so, which is what the benchmark's batch is built with."""
import random
import time

import pytest


def _py_hasher(t, tuples):
    from circuits_b200.poseidon import poseidon
    return [poseidon(list(x)) for x in tuples]


def _same_inputs(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k], k


def test_deferred_tree_equals_the_node_by_node_tree():
    from circuits_b200.batchtree import DeferredSMT, HashBatch, materialize
    from circuits_b200.smt import SMT
    rng = random.Random(5)
    ref, hb = SMT(), HashBatch(_py_hasher)
    tree = DeferredSMT(hb, 20)
    keys, got, want = [], [], []
    for step in range(60):
        if keys and rng.random() < 0.4:
            k, v = rng.choice(keys), rng.randrange(1 << 200)
            want.append(ref.update(k, v)); got.append(tree.update(k, v))
        else:
            # keys that share long low-bit prefixes: insertions push existing leaves several levels down
            k = rng.choice([rng.randrange(1 << 16), (rng.randrange(8) << 13) | 5, (rng.randrange(16) << 9) | 0x155])
            if k in keys:
                continue
            v = rng.randrange(1 << 200)
            keys.append(k)
            want.append(ref.insert(k, v)); got.append(tree.insert(k, v))
        if step % 17 == 16:
            hb.resolve()                    # several batches over one tree
    hb.resolve()
    assert materialize(got) == want
    assert int(tree.root) == ref.root


def test_mixed_batch_inputs_are_identical():
    from circuits_b200.synth import mixed_batch
    eager = mixed_batch(16)
    lazy = mixed_batch(16, deferred=True, hasher=_py_hasher)
    _same_inputs(eager.getInput(), lazy.getInput())
    assert eager.getHashInputs() == lazy.getHashInputs()
    assert lazy.hashes.launches < 80        # waves, not one call per hash


def test_bench_batch_inputs_are_identical():
    from circuits_b200.synth import bench_batch
    eager = bench_batch(24, 16, 6, 2, n_accounts=40)
    lazy = bench_batch(24, 16, 6, 2, n_accounts=40, deferred=True, hasher=_py_hasher)
    _same_inputs(eager.getInput(), lazy.getInput())
    assert eager.getHashInputs() == lazy.getHashInputs()


@pytest.mark.gpu
def test_gpu_built_batch_equals_the_host_built_one(built):
    from circuits_b200.synth import bench_batch
    t = time.time()
    eager = bench_batch(64, 32, 16, 4, n_accounts=128)
    t_host = time.time() - t
    t = time.time()
    gpu = bench_batch(64, 32, 16, 4, n_accounts=128, deferred=True)
    t_gpu = time.time() - t
    _same_inputs(eager.getInput(), gpu.getInput())
    assert eager.getHashInputs() == gpu.getHashInputs()
    print("batch of 64 over 128 accounts: %.1f s with the hashes on the host, %.1f s on the GPU (%d hashes in %d launches)"
          % (t_host, t_gpu, gpu.hashes.hashes, gpu.hashes.launches))
