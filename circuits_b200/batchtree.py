"""Batch builder on the device, first slice: the sparse Merkle trees of a batch with
every Poseidon hash evaluated on the GPU, level by level.

The host mirror (`smt.SMT`, after circomlib `smt.js` as the reference's `tools/generate-input.js:61-109` drives it through
@hermeznetwork/commonjs) walks the tree once per transaction and hashes 13-33 nodes one after the other: 10^5 sequential
Poseidon permutations per 2048-slot batch.  None of those hashes decides what the NEXT transaction does -- balances,
nonces and keys do -- so the structure of every proof (which node is the sibling at which level, which leaf is pushed
down by an insertion) is known from the key sets alone.  `DeferredSMT` keeps that structure eagerly (key counts per
prefix) and hands out *lazy* values for everything that is a hash: state hashes, leaf hashes, node hashes, roots,
siblings.  `HashBatch.resolve()` then evaluates all of them in waves of equal dependency depth (a state hash, its leaf
hash, then one wave per tree level: < 40 waves per batch), each wave one launch of the transposed Poseidon kernels of
`libhzwitness.so` (`hzw_poseidon_hash`: digests only).

Tree semantics (circomlib): leaf = Poseidon(key, value, 1), node = Poseidon(left, right), empty = 0, key bits consumed
LSB first; a sub-tree holding a single leaf is represented by that leaf at the highest free level.  So the value of the
node with prefix p is a function of the keys under p: none -> 0, one -> its leaf hash, two or more -> the hash of its
children -- which is what `_handle` returns.
"""


class Lazy:
    """A Poseidon hash that has not been evaluated yet: `ins` are ints or other Lazy values."""
    __slots__ = ("ins", "v", "depth")

    def __init__(self, ins):
        self.ins = ins
        self.v = None
        self.depth = 1 + max((x.depth for x in ins if isinstance(x, Lazy) and x.v is None), default=0)

    def __int__(self):
        if self.v is None:
            raise RuntimeError("lazy hash read before HashBatch.resolve()")
        return self.v


def value_of(x):
    return x.v if isinstance(x, Lazy) else x


class HashBatch:
    """Collects lazy hashes; resolve() evaluates them wave by wave with `hasher(t, tuples) -> digests`
    (t = number of inputs + 1).  The default hasher runs the transposed Poseidon kernels on the GPU."""

    def __init__(self, hasher=None):
        self.hasher = hasher or gpu_hasher
        self.pending = []
        self.launches = 0
        self.hashes = 0

    def hash(self, ins):
        h = Lazy(tuple(ins))
        self.pending.append(h)
        return h

    def resolve(self):
        waves = {}
        for h in self.pending:
            if h.v is None:
                waves.setdefault((h.depth, len(h.ins) + 1), []).append(h)
        for key in sorted(waves):
            group = waves[key]
            tuples = [tuple(value_of(x) for x in h.ins) for h in group]
            digests = self.hasher(key[1], tuples)
            self.launches += 1
            self.hashes += len(group)
            for h, d in zip(group, digests):
                h.v = d
                h.depth = 0
        self.pending = []


def gpu_hasher(t, tuples):
    import os
    from . import witness_calculator as wcm
    # one process per GPU (torchrun): every rank hashes on its own device
    return wcm.poseidon_hash_batch(t, tuples, device=int(os.environ.get("LOCAL_RANK", "0")))


def materialize(x):
    """ints in the place of every resolved lazy value of a nested input object"""
    if isinstance(x, Lazy):
        return int(x)
    if isinstance(x, list):
        if x and not isinstance(x[0], (list, dict)):         # flat list (siblings, roots): no recursion per element
            return [y.v if type(y) is Lazy else y for y in x]
        return [materialize(y) for y in x]
    if isinstance(x, dict):
        return {k: materialize(v) for k, v in x.items()}
    return x


class DeferredSMT:
    """`smt.SMT` with the same results (`insert` / `update` return oldRoot, newRoot, siblings, oldKey, oldValue, isOld0),
    every hash among them lazy until the batch is resolved."""

    def __init__(self, batch, max_levels=64):
        self.batch = batch
        self.max_levels = max_levels
        self.cnt = {}        # (level, prefix) -> keys under the prefix (levels 0 .. max_levels)
        self.only = {}       # (level, prefix) -> the key, while there is exactly one
        self.ver = {}        # (level, prefix) -> node hash, while there are two or more
        self.val = {}        # key -> value (int or Lazy)
        self.leaf = {}       # key -> leaf hash (Lazy)
        self.root = 0

    def copy(self, batch=None):
        n = DeferredSMT(batch or self.batch, self.max_levels)
        n.cnt, n.only, n.ver = dict(self.cnt), dict(self.only), dict(self.ver)
        n.val, n.leaf, n.root = dict(self.val), dict(self.leaf), self.root
        return n

    def _handle(self, node):
        c = self.cnt.get(node, 0)
        if c == 0:
            return 0
        if c == 1:
            return self.leaf[self.only[node]]
        return self.ver[node]

    def _walk(self, key):
        """siblings down to the node where the walk of smt.find ends, and that node"""
        sibs, level, prefix = [], 0, 0
        while self.cnt.get((level, prefix), 0) >= 2:
            bit = (key >> level) & 1
            child = prefix | (bit << level)
            sibs.append(self._handle((level + 1, child ^ (1 << level))))
            level, prefix = level + 1, child
        return sibs, (level, prefix)

    def _rehash(self, key):
        """new versions of the nodes above `key` that hold two or more keys, deepest first"""
        self.leaf[key] = self.batch.hash((key, self.val[key], 1))
        path, level, prefix = [], 0, 0
        while self.cnt.get((level, prefix), 0) >= 2:
            path.append((level, prefix))
            prefix |= ((key >> level) & 1) << level
            level += 1
        for level, prefix in reversed(path):
            left, right = (level + 1, prefix), (level + 1, prefix | (1 << level))
            self.ver[(level, prefix)] = self.batch.hash((self._handle(left), self._handle(right)))
        self.root = self._handle((0, 0))

    def insert(self, key, value):
        if key in self.val:
            raise KeyError("key already exists")
        sibs, end = self._walk(key)
        c = self.cnt.get(end, 0)
        old_root = self.root
        if c == 0:
            res = dict(isOld0=True, oldKey=0, oldValue=0)
        else:
            other = self.only[end]
            res = dict(isOld0=False, oldKey=other, oldValue=self.val[other])
        self.val[key] = value
        prefix = 0
        for level in range(self.max_levels + 1):
            node = (level, prefix)
            n = self.cnt.get(node, 0) + 1
            self.cnt[node] = n
            if n == 1:
                self.only[node] = key
            elif n == 2:
                self.only.pop(node, None)
            prefix |= ((key >> level) & 1) << level
        self._rehash(key)
        res.update(oldRoot=old_root, newRoot=self.root, siblings=sibs)
        return res

    def update(self, key, value):
        if key not in self.val:
            raise KeyError("key does not exist")
        sibs, _ = self._walk(key)
        old_root, old_value = self.root, self.val[key]
        self.val[key] = value
        self._rehash(key)
        return dict(oldRoot=old_root, newRoot=self.root, siblings=sibs, oldKey=key, oldValue=old_value, isOld0=False)
