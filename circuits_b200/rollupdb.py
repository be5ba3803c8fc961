"""Host-side rollup state + batch builder: produces `RollupMain` input objects.

Mirror (same method names and input keys) of the pieces of
@hermeznetwork/commonjs 0.3.3 [un-vendored] the reference drives its circuits
with: `RollupDB.buildBatch` / `BatchBuilder.addTx/addToken/addFeeIdx/build/
getInput/getHashInputs` (/root/reference/test/rollup-main.test.js:66-71,
tools/generate-input.js:84-109), `stateUtils.hashState`, `txUtils`.
The input key set and array extents are those of
/root/reference/src/rollup-main.circom:105-196; the state transition follows
the circuit (src/rollup-tx-states.circom, src/balance-updater.circom) so that
every `===` of an honestly built batch holds.
"""
import hashlib

from .field import P
from .poseidon import poseidon
from .smt import SMT, pad_siblings
from .batchtree import DeferredSMT, HashBatch, Lazy, materialize
from . import babyjub, float40
from .feetable import compute_fee

CONST_SIG = 3322668559          # src/decode-tx.circom:353
EXIT_IDX = 1                    # src/rollup-tx-states.circom:138
NULL_IDX = 0
FIRST_IDX = 255                 # last reserved idx; accounts start at 256
ETH_ADDR_ANY = (1 << 160) - 1   # src/rollup-tx-states.circom:130


def hash_state(st):
    """src/lib/hash-state.circom:18-40."""
    e0 = st["tokenID"] + (st["nonce"] << 32) + (st["sign"] << 72)
    return poseidon([e0, st["balance"], st["ay"], st["ethAddr"]])


def build_tx_compressed_data(tx, chain_id):
    """src/decode-tx.circom:79-86 bit layout."""
    return (CONST_SIG | (chain_id << 32) | (tx.get("fromIdx", 0) << 48) | (tx.get("toIdx", 0) << 96)
            | (tx.get("tokenID", 0) << 144) | (tx.get("nonce", 0) << 176)
            | (tx.get("userFee", 0) << 216) | (tx.get("toBjjSign", 0) << 224))


def build_tx_compressed_data_v2(tx):
    """src/decode-tx.circom:179-212 (L2 only; L1 keeps just the toBjjSign bit)."""
    sign = tx.get("toBjjSign", 0) << 216
    if tx.get("onChain"):
        return sign
    return (tx.get("fromIdx", 0) | (tx.get("toIdx", 0) << 48) | (tx.get("amountF", 0) << 96)
            | (tx.get("tokenID", 0) << 136) | (tx.get("nonce", 0) << 168)
            | (tx.get("userFee", 0) << 208) | sign)


def build_hash_sig(tx, chain_id):
    """sigL2Hash, src/decode-tx.circom:254-283."""
    e1 = tx.get("toEthAddr", 0) | (tx.get("amountF", 0) << 160) | (tx.get("maxNumBatch", 0) << 200)
    return poseidon([build_tx_compressed_data(tx, chain_id), e1, tx.get("toBjjAy", 0),
                     tx.get("rqTxCompressedDataV2", 0), tx.get("rqToEthAddr", 0),
                     tx.get("rqToBjjAy", 0)])


class Account:
    """Host mirror of commonjs `HermezAccount`: BJJ key + ethereum address."""

    def __init__(self, seed):
        h = hashlib.sha256(b"hermez-account-%d" % seed).digest()
        self.key = babyjub.KeyPair(int.from_bytes(h, "big"))
        self.ax, self.ay = self.key.pub
        self.sign = babyjub.sign_of(self.ax)
        self.ethAddr = int.from_bytes(hashlib.sha256(b"eth-%d" % seed).digest()[:20], "big")
        self.bjjCompressed = babyjub.pack_bits(self.key.pub)
        self._n = 0

    def sign_tx(self, tx, chain_id):
        msg = build_hash_sig(tx, chain_id)
        self._n += 1
        seed = int.from_bytes(hashlib.sha256(b"r-%d-%d" % (msg, self._n)).digest(), "big")
        R8, S = self.key.sign(msg, seed)
        tx["r8x"], tx["r8y"], tx["s"] = R8[0], R8[1], S
        return tx


class RollupDB:
    def __init__(self, chain_id=0, deferred=False, hasher=None):
        """deferred: the trees of a batch are `batchtree.DeferredSMT`s -- every Poseidon hash of the batch (state, leaf
        and node hashes) is evaluated when the batch is built, in waves, by `hasher(t, tuples)` (default: the transposed
        Poseidon kernels on the GPU) instead of one by one on the host.  Same inputs either way."""
        self.chainID = chain_id
        self.hashes = HashBatch(hasher) if deferred else None
        self.stateTree = DeferredSMT(self.hashes, 48) if deferred else SMT()
        self.states = {}            # idx -> state dict
        self.lastIdx = FIRST_IDX
        self.lastBatch = 0
        self.exitTrees = {}

    def buildBatch(self, nTx, nLevels, maxL1Tx, maxFeeTx):
        return BatchBuilder(self, nTx, nLevels, maxL1Tx, maxFeeTx)

    def consolidate(self, bb):
        if not bb.built:
            raise RuntimeError("batch not built")
        self.stateTree = bb.stateTree
        self.states = bb.states
        self.lastIdx = bb.finalIdx
        self.lastBatch = bb.currentNumBatch
        self.exitTrees[bb.currentNumBatch] = (bb.exitTree, bb.exitStates)

    def getStateByIdx(self, idx):
        return self.states.get(idx)


def _copy_tree(t):
    if isinstance(t, DeferredSMT):
        return t.copy()
    n = SMT()
    n.root = t.root
    n.db = t.db          # node store is append-only and content-addressed: safe to share
    return n


class BatchBuilder:
    def __init__(self, db, nTx, nLevels, maxL1Tx, maxFeeTx):
        self.nTx, self.nLevels, self.maxL1Tx, self.maxFeeTx = nTx, nLevels, maxL1Tx, maxFeeTx
        self.chainID = db.chainID
        self.currentNumBatch = db.lastBatch + 1
        self.hashes = db.hashes
        self.stateTree = _copy_tree(db.stateTree)
        self.states = {k: dict(v) for k, v in db.states.items()}
        self.exitTree = DeferredSMT(self.hashes, 48) if self.hashes else SMT()
        self.exitStates = {}
        self.oldLastIdx = db.lastIdx
        self.finalIdx = db.lastIdx
        self.oldStateRoot = int(db.stateTree.root)
        self.txs = []
        self.feePlanTokens = []
        self.feeIdxs = []
        self.built = False

    # -- reference API ------------------------------------------------------
    def addTx(self, tx):
        if len(self.txs) >= self.nTx:
            raise RuntimeError("batch is full")
        self.txs.append(dict(tx))

    def addToken(self, tokenID):
        if len(self.feePlanTokens) >= self.maxFeeTx:
            raise RuntimeError("fee plan is full")
        self.feePlanTokens.append(tokenID)

    def addFeeIdx(self, idx):
        if len(self.feeIdxs) >= self.maxFeeTx:
            raise RuntimeError("fee idxs full")
        self.feeIdxs.append(idx)

    # -- helpers --------------------------------------------------------------
    def _zeros(self, n):
        return [0] * n

    def _hash_state(self, st):
        if self.hashes is None:
            return hash_state(st)
        e0 = st["tokenID"] + (st["nonce"] << 32) + (st["sign"] << 72)
        return self.hashes.hash((e0, st["balance"], st["ay"], st["ethAddr"]))

    def _process(self, tx, I, accFee):
        """Apply one tx to the trees; append its per-slot inputs to I."""
        L = self.nLevels + 1
        onChain = 1 if tx.get("onChain") else 0
        fromIdx = tx.get("fromIdx", 0)
        toIdx = tx.get("toIdx", 0)
        amountF = tx.get("amountF", 0)
        loadAmountF = tx.get("loadAmountF", 0)
        amount = float40.float2fix(amountF)
        loadAmount = float40.float2fix(loadAmountF)
        tokenID = tx.get("tokenID", 0)
        userFee = tx.get("userFee", 0)
        nonce = tx.get("nonce", 0)
        fromEthAddr = tx.get("fromEthAddr", 0) if onChain else 0
        bjjBits = list(tx.get("fromBjjCompressed") or [0] * 256) if onChain else [0] * 256
        newAccount = 1 if (onChain and fromIdx == 0) else 0
        auxFromIdx = 0
        if newAccount:
            self.finalIdx += 1
            auxFromIdx = self.finalIdx
        finalFromIdx = auxFromIdx if newAccount else fromIdx
        auxToIdx = tx.get("auxToIdx", 0)
        finalToIdx = auxToIdx if (not onChain and toIdx == NULL_IDX) else toIdx
        isExit = finalToIdx == EXIT_IDX
        nop = finalFromIdx == 0
        isAmount = amount != 0
        isLoadAmount = loadAmount != 0

        # ---- state 1 (sender) ------------------------------------------------
        zero_state = dict(tokenID=0, nonce=0, sign=0, balance=0, ay=0, ethAddr=0)
        sib1 = self._zeros(L)
        isOld0_1 = oldKey1 = oldValue1 = 0
        if nop:
            st1 = dict(zero_state)
            s1 = None
        elif newAccount:
            ay = sum(b << i for i, b in enumerate(bjjBits[:254])) % P
            st1 = dict(tokenID=tokenID, nonce=0, sign=bjjBits[255], balance=0, ay=ay,
                       ethAddr=fromEthAddr)
            s1 = dict(st1)
        else:
            st1 = dict(self.states[finalFromIdx])
            s1 = dict(st1)

        # ---- nullifiers (src/rollup-tx-states.circom:260-313) ------------------
        notCreate = onChain and not newAccount
        nullEth = bool(notCreate and isAmount and fromEthAddr != st1["ethAddr"])
        nullTok1 = bool(notCreate and tokenID != st1["tokenID"])

        # ---- state 2 (receiver), needs P2 function -----------------------------
        p2_active = (not nop) and isAmount
        newExit = 0
        st2 = dict(zero_state)
        if p2_active:
            if isExit:
                if finalFromIdx in self.exitStates:
                    st2 = dict(self.exitStates[finalFromIdx])
                else:
                    newExit = 1
            elif finalToIdx == finalFromIdx:
                st2 = None        # self transfer: filled after processor 1
            else:
                if finalToIdx not in self.states:
                    raise ValueError("receiver %d does not exist" % finalToIdx)
                st2 = dict(self.states[finalToIdx])
        isP2Insert = isExit and newExit

        # ---- balances (src/balance-updater.circom:24-112) ----------------------
        fee = compute_fee(amount, userFee) if (not onChain and not nop) else 0
        effLoad = 0 if (not onChain or (nullTok1 and isLoadAmount)) else loadAmount
        effAmt1 = 0 if nop else amount
        bal_sender = 0 if newAccount else st1["balance"]

        def finish_nullifiers(tok2):
            nullTok2 = bool(onChain and isAmount and not isP2Insert and tokenID != tok2)
            return nullEth or nullTok2 or (nullTok1 and isAmount)

        # tokenID2 for the self-transfer case equals the sender's tokenID
        tok2_for_null = (s1["tokenID"] if st2 is None else st2["tokenID"]) if s1 else 0
        nullifyAmount = finish_nullifiers(tok2_for_null)
        effAmt2 = 0 if nullifyAmount else effAmt1
        underflowOk = bal_sender + effLoad - effAmt2 - fee >= 0
        if not underflowOk and not onChain:
            raise ValueError("L2 tx underflow")
        effAmt3 = effAmt2 if underflowOk else 0
        newBalSender = bal_sender + effLoad - effAmt3 - fee

        # ---- processor 1 ---------------------------------------------------------
        if not nop:
            new1 = dict(s1)
            new1["nonce"] = s1["nonce"] + (0 if onChain else 1)
            new1["balance"] = newBalSender
            if newAccount:
                r = self.stateTree.insert(finalFromIdx, self._hash_state(new1))
                isOld0_1 = 1 if r["isOld0"] else 0
                oldKey1, oldValue1 = r["oldKey"], r["oldValue"]
            else:
                r = self.stateTree.update(finalFromIdx, self._hash_state(new1))
            sib1 = pad_siblings(r["siblings"], L)
            self.states[finalFromIdx] = new1

        # ---- processor 2 ---------------------------------------------------------
        sib2 = self._zeros(L)
        isOld0_2 = oldKey2 = oldValue2 = 0
        if p2_active:
            if isExit:
                if newExit:
                    new2 = dict(tokenID=s1["tokenID"], nonce=0, sign=s1["sign"], balance=effAmt3,
                                ay=s1["ay"], ethAddr=s1["ethAddr"])
                    r = self.exitTree.insert(finalFromIdx, self._hash_state(new2))
                    isOld0_2 = 1 if r["isOld0"] else 0
                    oldKey2, oldValue2 = r["oldKey"], r["oldValue"]
                    st2 = dict(zero_state)
                else:
                    new2 = dict(st2)
                    new2["balance"] = st2["balance"] + effAmt3
                    r = self.exitTree.update(finalFromIdx, self._hash_state(new2))
                self.exitStates[finalFromIdx] = new2
            else:
                if st2 is None:
                    st2 = dict(self.states[finalToIdx])
                new2 = dict(st2)
                new2["balance"] = st2["balance"] + effAmt3
                r = self.stateTree.update(finalToIdx, self._hash_state(new2))
                self.states[finalToIdx] = new2
            sib2 = pad_siblings(r["siblings"], L)

        # ---- fees ----------------------------------------------------------------
        if fee:
            for j, tok in enumerate(self._plan()):
                if tok == tokenID:
                    accFee[j] += fee
                    break

        # ---- per-slot inputs -------------------------------------------------------
        txc = dict(tx, onChain=onChain, amountF=amountF)
        I["txCompressedData"].append(build_tx_compressed_data(txc, self.chainID))
        I["amountF"].append(amountF)
        I["txCompressedDataV2"].append(build_tx_compressed_data_v2(txc))
        I["fromIdx"].append(fromIdx)
        I["auxFromIdx"].append(auxFromIdx)
        I["toIdx"].append(toIdx)
        I["auxToIdx"].append(auxToIdx)
        I["toBjjAy"].append(tx.get("toBjjAy", 0))
        I["toEthAddr"].append(tx.get("toEthAddr", 0))
        I["maxNumBatch"].append(tx.get("maxNumBatch", 0))
        I["onChain"].append(onChain)
        I["newAccount"].append(newAccount)
        I["rqOffset"].append(tx.get("rqOffset", 0))
        I["rqTxCompressedDataV2"].append(tx.get("rqTxCompressedDataV2", 0))
        I["rqToEthAddr"].append(tx.get("rqToEthAddr", 0))
        I["rqToBjjAy"].append(tx.get("rqToBjjAy", 0))
        I["s"].append(tx.get("s", 0) if not onChain else 0)
        I["r8x"].append(tx.get("r8x", 0) if not onChain else 0)
        I["r8y"].append(tx.get("r8y", 0) if not onChain else 0)
        I["loadAmountF"].append(loadAmountF)
        I["fromEthAddr"].append(fromEthAddr)
        I["fromBjjCompressed"].append(bjjBits)
        for k in ("tokenID", "nonce", "sign", "balance", "ay", "ethAddr"):
            I[k + "1"].append(st1[k] if not newAccount else
                              {"tokenID": tokenID, "nonce": 0, "sign": st1["sign"], "balance": 0,
                               "ay": st1["ay"], "ethAddr": fromEthAddr}[k])
            I[k + "2"].append(st2[k])
        I["siblings1"].append(sib1)
        I["isOld0_1"].append(isOld0_1)
        I["oldKey1"].append(oldKey1)
        I["oldValue1"].append(oldValue1)
        I["siblings2"].append(sib2)
        I["newExit"].append(newExit)
        I["isOld0_2"].append(isOld0_2)
        I["oldKey2"].append(oldKey2)
        I["oldValue2"].append(oldValue2)
        return dict(onChain=onChain, isAmountNullified=(nullifyAmount or not underflowOk),
                    finalToIdx=finalToIdx, fromIdx=fromIdx, amountF=amountF, userFee=userFee,
                    loadAmountF=loadAmountF, fromEthAddr=fromEthAddr, bjjBits=bjjBits,
                    tokenID=tokenID, toIdx=toIdx)

    def _plan(self):
        return self.feePlanTokens + [0] * (self.maxFeeTx - len(self.feePlanTokens))

    def build(self):
        if self.built:
            raise RuntimeError("batch already built")
        nTx, maxFeeTx, L = self.nTx, self.maxFeeTx, self.nLevels + 1
        keys = ["txCompressedData", "amountF", "txCompressedDataV2", "fromIdx", "auxFromIdx", "toIdx",
                "auxToIdx", "toBjjAy", "toEthAddr", "maxNumBatch", "onChain", "newAccount", "rqOffset",
                "rqTxCompressedDataV2", "rqToEthAddr", "rqToBjjAy", "s", "r8x", "r8y", "loadAmountF",
                "fromEthAddr", "fromBjjCompressed", "tokenID1", "nonce1", "sign1", "balance1", "ay1",
                "ethAddr1", "siblings1", "isOld0_1", "oldKey1", "oldValue1", "tokenID2", "nonce2",
                "sign2", "balance2", "ay2", "ethAddr2", "siblings2", "newExit", "isOld0_2", "oldKey2",
                "oldValue2"]
        I = {k: [] for k in keys}
        I["oldLastIdx"] = self.oldLastIdx
        I["oldStateRoot"] = self.oldStateRoot
        I["globalChainID"] = self.chainID
        I["currentNumBatch"] = self.currentNumBatch
        I["feePlanTokens"] = self._plan()
        I["feeIdxs"] = self.feeIdxs + [0] * (maxFeeTx - len(self.feeIdxs))
        for k in ("imOnChain", "imOutIdx", "imStateRoot", "imExitRoot", "imAccFeeOut"):
            I[k] = []
        accFee = [0] * maxFeeTx
        l1 = [t for t in self.txs if t.get("onChain")]
        l2 = [t for t in self.txs if not t.get("onChain")]
        if len(l1) > self.maxL1Tx:
            raise RuntimeError("too many L1 txs")
        ordered = l1 + l2 + [dict() for _ in range(nTx - len(self.txs))]
        self._meta = []
        for i, tx in enumerate(ordered):
            if not tx:
                tx = dict(fromIdx=0, toIdx=0, onChain=0)
            meta = self._process(tx, I, accFee)
            self._meta.append(meta)
            if i < nTx - 1:
                I["imOnChain"].append(meta["onChain"])
                I["imOutIdx"].append(self.finalIdx)
                I["imStateRoot"].append(self.stateTree.root)
                I["imExitRoot"].append(self.exitTree.root)
                I["imAccFeeOut"].append(list(accFee))
        I["imInitStateRootFee"] = self.stateTree.root
        I["imFinalAccFee"] = list(accFee)
        # fee txs (src/fee-tx.circom:26-111)
        I["imStateRootFee"] = []
        for k in ("tokenID3", "nonce3", "sign3", "balance3", "ay3", "ethAddr3", "siblings3"):
            I[k] = []
        for j in range(maxFeeTx):
            idx = I["feeIdxs"][j]
            if idx == 0:
                st = dict(tokenID=0, nonce=0, sign=0, balance=0, ay=0, ethAddr=0)
                sib = [0] * L
            else:
                st = dict(self.states[idx])
                if st["tokenID"] != I["feePlanTokens"][j]:
                    raise ValueError("fee idx token mismatch")
                new = dict(st)
                new["balance"] = st["balance"] + accFee[j]
                r = self.stateTree.update(idx, self._hash_state(new))
                sib = pad_siblings(r["siblings"], L)
                self.states[idx] = new
            for k in ("tokenID", "nonce", "sign", "balance", "ay", "ethAddr"):
                I[k + "3"].append(st[k])
            I["siblings3"].append(sib)
            if j < maxFeeTx - 1:
                I["imStateRootFee"].append(self.stateTree.root)
        if self.hashes is not None:
            self.hashes.resolve()           # every hash of the batch, wave by wave (batchtree.py)
            I = materialize(I)
        self.newStateRoot = int(self.stateTree.root)
        self.newExitRoot = int(self.exitTree.root)
        self.input = I
        self.built = True
        return self

    def getInput(self):
        if not self.built:
            raise RuntimeError("batch not built")
        return self.input

    # -- public-input hash (src/hash-inputs.circom:23-185) ---------------------
    def getHashInputs(self):
        """SHA-256 of the pretended public inputs, as a field element."""
        nL = self.nLevels
        bits = []

        def put(v, n):
            bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))
        put(self.oldLastIdx, 48)
        put(self.finalIdx, 48)
        put(self.oldStateRoot, 256)
        put(self.newStateRoot, 256)
        put(self.newExitRoot, 256)
        for i in range(self.maxL1Tx):
            m = self._meta[i] if i < len(self._meta) else None
            if m and m["onChain"]:
                put(m["fromEthAddr"], 160)
                bits.extend(m["bjjBits"][::-1])
                put(m["fromIdx"], 48)
                put(m["loadAmountF"], 40)
                put(m["amountF"], 40)
                put(m["tokenID"], 32)
                put(m["toIdx"], 48)
            else:
                bits.extend([0] * 624)
        for m in self._meta:
            put(m["fromIdx"], nL)
            put(m["finalToIdx"] % (1 << nL) if True else 0, nL)
            put(0 if m["isAmountNullified"] else m["amountF"], 40)
            put(0 if m["onChain"] else m["userFee"], 8)
        for idx in self.input["feeIdxs"]:
            put(idx, nL)
        put(self.chainID, 16)
        put(self.currentNumBatch, 32)
        by = bytes(sum(bits[8 * i + k] << (7 - k) for k in range(8)) for i in range(len(bits) // 8))
        return int.from_bytes(hashlib.sha256(by).digest(), "big") % P
