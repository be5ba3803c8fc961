"""In-memory sparse Merkle tree with circomlib semantics (host side).

Mirror of circomlib `smt.js` + `SMTMemDB` [un-vendored] as used by the
reference's input generation (/root/reference/test/rollup-main.test.js:5,
tools/generate-input.js:6).  Leaf = Poseidon(key, value, 1), node =
Poseidon(left, right), empty = 0, key bits consumed LSB-first; a sub-tree
holding a single leaf is represented by that leaf at the highest free level.
The proofs it returns are what SMTProcessor / SMTVerifier expect
(`siblings`, `oldKey`, `oldValue`, `isOld0`).
"""
from .poseidon import poseidon


def hash_leaf(key, value):
    return poseidon([key, value, 1])


def hash_node(left, right):
    return poseidon([left, right])


class SMT:
    def __init__(self):
        self.root = 0
        self.db = {}   # node hash -> ("L", key, value) | ("N", left, right)

    def find(self, key):
        sibs = []
        node = self.root
        level = 0
        while True:
            if node == 0:
                return dict(found=False, siblings=sibs, isOld0=True, notFoundKey=0, notFoundValue=0)
            rec = self.db[node]
            if rec[0] == "L":
                if rec[1] == key:
                    return dict(found=True, siblings=sibs, foundValue=rec[2], isOld0=False)
                return dict(found=False, siblings=sibs, isOld0=False,
                            notFoundKey=rec[1], notFoundValue=rec[2])
            if (key >> level) & 1:
                sibs.append(rec[1])
                node = rec[2]
            else:
                sibs.append(rec[2])
                node = rec[1]
            level += 1

    def _climb(self, key, node, sibs):
        for lvl in range(len(sibs) - 1, -1, -1):
            if (key >> lvl) & 1:
                pair = (sibs[lvl], node)
            else:
                pair = (node, sibs[lvl])
            node = hash_node(*pair)
            self.db[node] = ("N",) + pair
        return node

    def insert(self, key, value):
        res = self.find(key)
        if res["found"]:
            raise KeyError("key already exists")
        proof_sibs = list(res["siblings"])
        sibs = list(proof_sibs)
        leaf = hash_leaf(key, value)
        self.db[leaf] = ("L", key, value)
        if not res["isOld0"]:
            okey = res["notFoundKey"]
            lvl = len(sibs)
            while ((okey >> lvl) & 1) == ((key >> lvl) & 1):
                sibs.append(0)
                lvl += 1
            sibs.append(hash_leaf(okey, res["notFoundValue"]))
        old_root = self.root
        self.root = self._climb(key, leaf, sibs)
        return dict(oldRoot=old_root, newRoot=self.root, siblings=proof_sibs,
                    oldKey=res["notFoundKey"], oldValue=res["notFoundValue"], isOld0=res["isOld0"])

    def update(self, key, value):
        res = self.find(key)
        if not res["found"]:
            raise KeyError("key does not exist")
        sibs = res["siblings"]
        leaf = hash_leaf(key, value)
        self.db[leaf] = ("L", key, value)
        old_root = self.root
        self.root = self._climb(key, leaf, sibs)
        return dict(oldRoot=old_root, newRoot=self.root, siblings=list(sibs),
                    oldKey=key, oldValue=res["foundValue"], isOld0=False)


def pad_siblings(sibs, n):
    if len(sibs) > n:
        raise ValueError("proof deeper than nLevels+1")
    return list(sibs) + [0] * (n - len(sibs))
