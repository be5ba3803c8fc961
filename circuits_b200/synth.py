"""Seeded synthetic batches for tests and the benchmark.

Stands in for /root/reference/tools/generate-input.js (+ helpers/gen-inputs-utils.js), which
drives @hermeznetwork/commonjs with an unseeded `Math.random` and cannot run offline: the state
tree is pre-populated with accounts, then one batch is built with `maxL1Tx` L1
`createAccountDeposit` txs (generate-input.js:88-90) followed by L2 transfers of 20 % of the
sender's balance with `userFee` 176 (gen-inputs-utils.js:27-70), one fee token and one fee
collector (generate-input.js:102-103).
"""
import hashlib
import random

from . import babyjub, float40
from .rollupdb import RollupDB, Account

SEED = 20261017
_B8_DOUBLES = None


def _mul_base8(k):
    """Fixed-base scalar multiplication on a table of 4-bit windows of Base8 (d * 16^w * Base8, d = 1..15), accumulated
    in extended coordinates (unified addition, one inversion at the end): host speed only."""
    global _B8_DOUBLES
    if _B8_DOUBLES is None:
        t, base = [], babyjub.BASE8
        for _ in range(64):
            row, pt = [None], base
            for d in range(1, 16):
                row.append((pt[0], pt[1], pt[0] * pt[1] % babyjub.P))
                pt = babyjub.add(pt, base)
            t.append(row)
            base = pt                      # 16 * base
        _B8_DOUBLES = t
    P, A, D = babyjub.P, babyjub.A, babyjub.D
    X, Y, Z, T = 0, 1, 1, 0
    w = 0
    while k:
        d = k & 15
        if d:
            x2, y2, t2 = _B8_DOUBLES[w][d]
            a, b, c = X * x2 % P, Y * y2 % P, D * T % P * t2 % P
            e = ((X + Y) * (x2 + y2) - a - b) % P
            f, g, h = (Z - c) % P, (Z + c) % P, (b - A * a) % P
            X, Y, Z, T = e * f % P, g * h % P, f * g % P, e * h % P
        k >>= 4
        w += 1
    zi = pow(Z, -1, P)
    return (X * zi % P, Y * zi % P)


class FastAccount(Account):
    """Account whose key generation / signing use the fixed-base table."""

    def __init__(self, seed):
        h = hashlib.sha256(b"hermez-account-%d" % seed).digest()
        kp = babyjub.KeyPair.__new__(babyjub.KeyPair)
        kp.a = int.from_bytes(h, "big") % babyjub.SUBORDER or 1
        kp.s = kp.a * 8
        kp.pub = _mul_base8(kp.a)

        def sign(msg, nonce_seed, kp=kp):
            r = (nonce_seed % (babyjub.SUBORDER - 1)) + 1
            R8 = _mul_base8(r)
            hm = babyjub.poseidon([R8[0], R8[1], kp.pub[0], kp.pub[1], msg])
            return R8, (r + hm * kp.s) % babyjub.SUBORDER
        kp.sign = sign
        self.key = kp
        self.ax, self.ay = kp.pub
        self.sign = babyjub.sign_of(self.ax)
        self.ethAddr = int.from_bytes(hashlib.sha256(b"eth-%d" % seed).digest()[:20], "big")
        self.bjjCompressed = babyjub.pack_bits(kp.pub)
        self._n = 0


def _deposit(acc, amount, token=1):
    return dict(fromIdx=0, loadAmountF=float40.floor_fix2float(amount), tokenID=token,
                fromBjjCompressed=acc.bjjCompressed, fromEthAddr=acc.ethAddr, toIdx=0, onChain=1)


def small_batch(nTx=3, nLevels=16, maxL1Tx=2, maxFeeTx=2):
    """Two deposits consolidated, then one signed transfer with a fee and a fee collector."""
    db = RollupDB(chain_id=0)
    a1, a2 = Account(1), Account(2)
    bb = db.buildBatch(nTx, nLevels, maxL1Tx, maxFeeTx)
    bb.addTx(_deposit(a1, 1000))
    bb.addTx(_deposit(a2, 1000))
    bb.build()
    db.consolidate(bb)
    bb = db.buildBatch(nTx, nLevels, maxL1Tx, maxFeeTx)
    tx = dict(fromIdx=256, toIdx=257, tokenID=1, amountF=float40.fix2float(50), nonce=0, userFee=126, onChain=0)
    a1.sign_tx(tx, 0)
    bb.addTx(tx)
    bb.addToken(1)
    bb.addFeeIdx(257)
    bb.build()
    return bb


def bench_batch(nTx, nLevels, maxL1Tx, maxFeeTx, n_accounts=None, n_keys=64, seed=SEED, n_l2=None, deferred=False, hasher=None):
    """The benchmark batch: `maxL1Tx` deposits + (nTx - maxL1Tx) signed transfers.

    `n_keys` distinct key pairs are shared round-robin by the accounts (key generation is the slow
    part of the host-side setup and is not part of what is measured)."""
    rng = random.Random(seed)
    if n_accounts is None:
        n_accounts = max(2 * nTx, 16)
    keys = [FastAccount(k) for k in range(n_keys)]
    db = RollupDB(chain_id=0, deferred=deferred, hasher=hasher)     # deferred: the hashes of every batch on the GPU (batchtree.py)
    owners = {}
    # populate: batches of createAccountDeposit (generate-input.js:61-82)
    todo = n_accounts
    while todo > 0:
        n = min(todo, maxL1Tx) if maxL1Tx > 0 else 0
        if n == 0:
            break
        pb = db.buildBatch(max(n, 1), nLevels, max(n, 1), 1)
        for _ in range(n):
            acc = keys[len(owners) % n_keys]
            owners[256 + len(owners)] = acc
            pb.addTx(_deposit(acc, rng.randrange(1 << 40, 1 << 96)))
        pb.build()
        db.consolidate(pb)
        todo -= n
    bb = db.buildBatch(nTx, nLevels, maxL1Tx, maxFeeTx)
    n_l1 = maxL1Tx
    for _ in range(n_l1):
        acc = keys[rng.randrange(n_keys)]
        bb.addTx(_deposit(acc, rng.randrange(1 << 40, 1 << 96)))
    idxs = sorted(owners)
    nonces = {}
    bal = {i: db.states[i]["balance"] for i in idxs}
    n_l2 = nTx - n_l1 if n_l2 is None else n_l2
    to_sign = []
    for _ in range(n_l2):
        a, b = rng.sample(idxs, 2)
        amountF = float40.floor_fix2float(bal[a] // 5)
        amount = float40.float2fix(amountF)
        from .feetable import compute_fee
        fee = compute_fee(amount, 176)
        if bal[a] < amount + fee:
            amountF, amount, fee = 0, 0, 0
        tx = dict(fromIdx=a, toIdx=b, tokenID=1, amountF=amountF, nonce=nonces.get(a, 0), userFee=176, onChain=0)
        if deferred:
            to_sign.append((owners[a], tx))
        else:
            owners[a].sign_tx(tx, 0)
        bb.addTx(tx)
        nonces[a] = nonces.get(a, 0) + 1
        bal[a] -= amount + fee
        bal[b] += amount
    if to_sign:
        _sign_batch(to_sign, bb, db.hashes.hasher)
    bb.addToken(1)
    bb.addFeeIdx(idxs[rng.randrange(len(idxs))])
    bb.build()
    return bb


def _sign_batch(to_sign, bb, hasher):
    """EdDSA-Poseidon signatures of a batch with the two hashes of every signature (message hash, width 7; the
    verifier's hash of R8, A, M, width 6) evaluated together: same values as Account.sign_tx one by one."""
    from .rollupdb import build_tx_compressed_data
    msgs = hasher(7, [(build_tx_compressed_data(tx, 0),
                       tx.get("toEthAddr", 0) | (tx.get("amountF", 0) << 160) | (tx.get("maxNumBatch", 0) << 200),
                       tx.get("toBjjAy", 0), tx.get("rqTxCompressedDataV2", 0), tx.get("rqToEthAddr", 0), tx.get("rqToBjjAy", 0))
                      for _, tx in to_sign])
    rs, r8s = [], []
    for (acc, tx), msg in zip(to_sign, msgs):
        acc._n += 1
        seed = int.from_bytes(hashlib.sha256(b"r-%d-%d" % (msg, acc._n)).digest(), "big")
        r = (seed % (babyjub.SUBORDER - 1)) + 1
        rs.append(r)
        r8s.append(_mul_base8(r))
    hms = hasher(6, [(r8[0], r8[1], acc.key.pub[0], acc.key.pub[1], msg) for (acc, _), r8, msg in zip(to_sign, r8s, msgs)])
    for (acc, tx), r, r8, hm in zip(to_sign, rs, r8s, hms):
        tx["r8x"], tx["r8y"], tx["s"] = r8[0], r8[1], (r + hm * acc.key.s) % babyjub.SUBORDER
    # bb.txs holds copies made by addTx, in the order they were added after the L1 txs
    l2 = [t for t in bb.txs if not t.get("onChain")]
    assert len(l2) == len(to_sign)
    for dst, (_, src) in zip(l2, to_sign):
        dst["r8x"], dst["r8y"], dst["s"] = src["r8x"], src["r8y"], src["s"]


def withdraw_inputs(nLevels, n, seed=SEED):
    """`n` Withdraw(nLevels) inputs: the leaves of one exit tree (what `withdrawUtils` / `getExitTreeInfo`
    produce in /root/reference/test/withdraw.test.js:60-103) with their inclusion proofs, plus the
    hashGlobalInputs each must yield (SHA-256 of rootExit|ethAddr|tokenID|balance|idx, mod p:
    src/withdraw.circom:84-176)."""
    from .field import P
    from .rollupdb import hash_state
    from .smt import SMT, pad_siblings
    rng = random.Random(seed)
    tree, leaves = SMT(), []
    used = set()
    for k in range(n):
        idx = rng.randrange(256, 1 << min(nLevels, 40))
        while idx in used:
            idx = rng.randrange(256, 1 << min(nLevels, 40))
        used.add(idx)
        st = dict(tokenID=rng.randrange(1 << 32), nonce=0, sign=rng.randrange(2), balance=rng.randrange(1 << 96),
                  ay=rng.randrange(P), ethAddr=rng.randrange(1 << 160))
        tree.insert(idx, hash_state(st))
        leaves.append((idx, st))
    out = []
    for idx, st in leaves:
        res = tree.find(idx)
        assert res["found"]
        inp = dict(rootExit=tree.root, ethAddr=st["ethAddr"], tokenID=st["tokenID"], balance=st["balance"], idx=idx,
                   sign=st["sign"], ay=st["ay"], siblingsState=pad_siblings(res["siblings"], nLevels + 1))
        pre = ((((tree.root << 160 | st["ethAddr"]) << 32 | st["tokenID"]) << 192 | st["balance"]) << 48) | idx
        h = int.from_bytes(hashlib.sha256(pre.to_bytes(86, "big")).digest(), "big") % P
        out.append((inp, h))
    return out


def mixed_batch(nLevels=16, nTx=12, maxL1Tx=7, maxFeeTx=3, deferred=False, hasher=None):
    """One batch with every transaction type of the protocol, after the cases of
    /root/reference/test/rollup-main.test.js:65-878 and rollup-main-L1.test.js:88-487: L1 createAccountDeposit,
    deposit, createAccountDepositTransfer, depositTransfer, forceTransfer, forceExit, an L1 nullified by a wrong
    token; L2 transfer, exit, a second exit of the same account (update of the exit leaf), transfer with
    request-offset linkage left empty; NOP padding; two fee tokens and collectors."""
    db = RollupDB(chain_id=0, deferred=deferred, hasher=hasher)
    acc = [Account(k + 1) for k in range(4)]
    bb = db.buildBatch(4, nLevels, 4, 1)
    for a, amt, tok in ((acc[0], 10000, 1), (acc[1], 5000, 1), (acc[2], 7000, 2), (acc[3], 0, 1)):
        bb.addTx(_deposit(a, amt, tok))
    bb.build()
    db.consolidate(bb)                      # accounts 256..259
    bb = db.buildBatch(nTx, nLevels, maxL1Tx, maxFeeTx)
    f = float40.fix2float
    new = Account(9)
    # L1 transactions
    bb.addTx(_deposit(new, 300))                                                             # createAccountDeposit -> 260
    bb.addTx(dict(fromIdx=256, loadAmountF=f(500), tokenID=1, fromBjjCompressed=[0] * 256,   # deposit
                  fromEthAddr=acc[0].ethAddr, toIdx=0, onChain=1))
    t = _deposit(Account(10), 800)                                                           # createAccountDepositTransfer -> 261
    t.update(toIdx=257, amountF=f(100))
    bb.addTx(t)
    bb.addTx(dict(fromIdx=257, loadAmountF=f(200), amountF=f(50), tokenID=1, fromBjjCompressed=[0] * 256,   # depositTransfer
                  fromEthAddr=acc[1].ethAddr, toIdx=256, onChain=1))
    bb.addTx(dict(fromIdx=256, loadAmountF=0, amountF=f(70), tokenID=1, fromBjjCompressed=[0] * 256,        # forceTransfer
                  fromEthAddr=acc[0].ethAddr, toIdx=259, onChain=1))
    bb.addTx(dict(fromIdx=257, loadAmountF=0, amountF=f(30), tokenID=1, fromBjjCompressed=[0] * 256,        # forceExit
                  fromEthAddr=acc[1].ethAddr, toIdx=1, onChain=1))
    bb.addTx(dict(fromIdx=258, loadAmountF=f(40), amountF=0, tokenID=1, fromBjjCompressed=[0] * 256,        # deposit with the wrong token: nullified
                  fromEthAddr=acc[2].ethAddr, toIdx=0, onChain=1))
    # L2 transactions
    tx = dict(fromIdx=256, toIdx=257, tokenID=1, amountF=f(150), nonce=0, userFee=126, onChain=0)
    acc[0].sign_tx(tx, 0); bb.addTx(tx)
    tx = dict(fromIdx=256, toIdx=1, tokenID=1, amountF=f(100), nonce=1, userFee=68, onChain=0)              # exit
    acc[0].sign_tx(tx, 0); bb.addTx(tx)
    tx = dict(fromIdx=256, toIdx=1, tokenID=1, amountF=f(20), nonce=2, userFee=0, onChain=0)                # second exit: exit leaf update
    acc[0].sign_tx(tx, 0); bb.addTx(tx)
    tx = dict(fromIdx=258, toIdx=258, tokenID=2, amountF=f(500), nonce=0, userFee=184, onChain=0)           # self transfer, other token
    acc[2].sign_tx(tx, 0); bb.addTx(tx)
    bb.addToken(1); bb.addFeeIdx(259)
    bb.addToken(2); bb.addFeeIdx(258)
    bb.build()
    return bb
