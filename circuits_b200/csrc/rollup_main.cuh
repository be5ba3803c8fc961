// RollupMain (/root/reference/src/rollup-main.circom:82-475) cut into sections that a GPU can
// schedule independently.  The circuit itself makes the cut possible: every cross-slot
// dependency is an *input* (`im*`, :93-99,117-126), so a slot needs nothing from its
// neighbours but main-component inputs.  The layout walk runs the sections in circuit order;
// the kernels give each section to its own thread(s), starting at the witness offset the
// walk recorded in the plan.
#pragma once
#include "templates_hermez.cuh"

struct HzParams { int nTx, nLevels, maxL1Tx, maxFeeTx; };

// main-component inputs in declaration order (src/rollup-main.circom:105-196)
enum HzIn {
    IN_oldLastIdx, IN_oldStateRoot, IN_globalChainID, IN_currentNumBatch, IN_feeIdxs, IN_feePlanTokens,
    IN_imOnChain, IN_imOutIdx, IN_imStateRoot, IN_imExitRoot, IN_imAccFeeOut, IN_imStateRootFee,
    IN_imInitStateRootFee, IN_imFinalAccFee,
    IN_txCompressedData, IN_amountF, IN_txCompressedDataV2, IN_fromIdx, IN_auxFromIdx, IN_toIdx, IN_auxToIdx,
    IN_toBjjAy, IN_toEthAddr, IN_maxNumBatch, IN_onChain, IN_newAccount, IN_rqOffset, IN_rqTxCompressedDataV2,
    IN_rqToEthAddr, IN_rqToBjjAy, IN_s, IN_r8x, IN_r8y, IN_loadAmountF, IN_fromEthAddr, IN_fromBjjCompressed,
    IN_tokenID1, IN_nonce1, IN_sign1, IN_balance1, IN_ay1, IN_ethAddr1, IN_siblings1, IN_isOld0_1, IN_oldKey1,
    IN_oldValue1, IN_tokenID2, IN_nonce2, IN_sign2, IN_balance2, IN_ay2, IN_ethAddr2, IN_siblings2, IN_newExit,
    IN_isOld0_2, IN_oldKey2, IN_oldValue2, IN_tokenID3, IN_nonce3, IN_sign3, IN_balance3, IN_ay3, IN_ethAddr3,
    IN_siblings3, IN_COUNT
};

struct HzPlan {
    HzParams P;
    uint32_t in_off[IN_COUNT];   // witness index of element [0][0] of each input
    uint32_t in_d0[IN_COUNT];    // outer extent (1 for scalars)
    uint32_t in_d1[IN_COUNT];    // inner extent (1 unless 2-D)
    uint32_t n_inputs;
    uint32_t nvars;
    HzSizes sz;
    // hasherInputs
    uint32_t nbits, nblocks;
    uint32_t off_hasher, off_sha_blocks, off_hasher_tail;
    uint32_t idx_newLastIdx, idx_newStateRoot, idx_newExitRoot;
    // per-slot section offsets (device pointers in the device copy)
    const uint32_t* off_decode;
    const uint32_t* off_rollup;
    const uint32_t* off_fee;
};

HZ_HD uint32_t hz_pre_nbits(const HzParams& P) {
    return 48 + 48 + 256 * 3 + (uint32_t)P.maxL1Tx * 624 + (uint32_t)P.nTx * (2 * P.nLevels + 48) + (uint32_t)P.maxFeeTx * P.nLevels + 16 + 32;
}
HZ_HD uint32_t hz_pre_pos_l1(const HzParams&) { return 48 + 48 + 256 * 3; }
HZ_HD uint32_t hz_pre_pos_l2(const HzParams& P) { return hz_pre_pos_l1(P) + (uint32_t)P.maxL1Tx * 624; }

// SHA pre-image as seen by the sections: bit values (compute) or witness indexes (layout)
TM struct PreBits;
template <> struct PreBits<MCompute> {
    uint8_t* b;
    HZ_HD void put(uint32_t pos, const VC& x) { b[pos] = (uint8_t)(x.v.l[0] & 1u); }
    HZ_HD VC get(uint32_t pos) const { VC r; r.v = fr_from_u64(b[pos]); r.sig = true; return r; }
    HZ_HD uint32_t bit(uint32_t pos) const { return b[pos]; }
};
template <> struct PreBits<MLayout> {
    uint32_t* idx;
    HZ_HD void put(uint32_t pos, const VL& x) { idx[pos] = x.idx; }
    HZ_HD VL get(uint32_t pos) const { VL r; r.idx = idx[pos]; return r; }
    HZ_HD uint32_t bit(uint32_t) const { return 0; }
};

TM HZ_HD V<M> hz_in(const Comp<M>& c, const HzPlan& pl, int f, int a = 0, int b = 0) {
    return c.load(pl.in_off[f] + (uint32_t)a * pl.in_d1[f] + (uint32_t)b);
}

// main inputs the dry walk of k_smtjob never reads the VALUE of (fee plan and accumulated fees: the fee accumulator is
// skipped there; siblings: the processors are not evaluated, their jobs carry the siblings' witness index): each is a
// dependent global load at the head of the merkle chain, so the dry walk gets a wired zero instead
TM HZ_HD V<M> hz_in_unless_dry(const Comp<M>& c, const HzPlan& pl, int f, int a = 0, int b = 0) {
    if (hz_is_dry(c)) return c.known(0, Ops<M>::K(0));
    return hz_in(c, pl, f, a, b);
}

// ---- section B/C: decodeTx[i]  (:223-265) ---------------------------------------------------
TM HZ_FN void sec_decode(Comp<M>& c, const HzPlan& pl, int i, DecodeOut<M>& o, PreBits<M>& pre) {
    const HzParams& P = pl.P;
    DecodeIn<M> in;
    in.previousOnChain = i == 0 ? Ops<M>::K(1) : hz_in(c, pl, IN_imOnChain, i - 1);
    in.txCompressedData = hz_in(c, pl, IN_txCompressedData, i);
    in.maxNumBatch = hz_in(c, pl, IN_maxNumBatch, i);
    in.amountF = hz_in(c, pl, IN_amountF, i);
    in.toEthAddr = hz_in(c, pl, IN_toEthAddr, i);
    in.toBjjAy = hz_in(c, pl, IN_toBjjAy, i);
    in.rqTxCompressedDataV2 = hz_in(c, pl, IN_rqTxCompressedDataV2, i);
    in.rqToEthAddr = hz_in(c, pl, IN_rqToEthAddr, i);
    in.rqToBjjAy = hz_in(c, pl, IN_rqToBjjAy, i);
    in.fromEthAddr = hz_in(c, pl, IN_fromEthAddr, i);
    in.fromBjjCompressed.clear(256);
    for (int k = 0; k < 256; k++) in.fromBjjCompressed.put(k, hz_in(c, pl, IN_fromBjjCompressed, i, k));
    in.loadAmountF = hz_in(c, pl, IN_loadAmountF, i);
    in.globalChainID = hz_in(c, pl, IN_globalChainID);
    in.currentNumBatch = hz_in(c, pl, IN_currentNumBatch);
    in.onChain = hz_in(c, pl, IN_onChain, i);
    in.newAccount = hz_in(c, pl, IN_newAccount, i);
    in.auxFromIdx = hz_in(c, pl, IN_auxFromIdx, i);
    in.auxToIdx = hz_in(c, pl, IN_auxToIdx, i);
    in.inIdx = i == 0 ? hz_in(c, pl, IN_oldLastIdx) : hz_in(c, pl, IN_imOutIdx, i - 1);
    if (c.sanity()) {
        // A - binary checks (:207-219)
        if (i < P.nTx - 1) { V<M> x = hz_in(c, pl, IN_imOnChain, i); c.check0(x * (x - 1), __LINE__); }
        c.check0(in.onChain * (in.onChain - 1), __LINE__);
        c.check0(in.newAccount * (in.newAccount - 1), __LINE__);
        for (int k = 0; k < 256; k++) {
            V<M> x = hz_in(c, pl, IN_fromBjjCompressed, i, k);
            c.check0(x * (x - 1), __LINE__);
        }
        V<M> a = hz_in(c, pl, IN_isOld0_1, i), b = hz_in(c, pl, IN_isOld0_2, i);
        c.check0(a * (a - 1), __LINE__);
        c.check0(b * (b - 1), __LINE__);
    }
    DecodeTx(c.sub(Nm("decodeTx", i)), P.nLevels, in, o);
    if (c.sanity()) {
        // C - integrity of the decode intermediary signals (:258-265)
        c.check(o.txCompressedDataV2, hz_in(c, pl, IN_txCompressedDataV2, i), __LINE__);
        if (i < P.nTx - 1) {
            c.check(o.onChain, hz_in(c, pl, IN_imOnChain, i), __LINE__);
            c.check(o.outIdx, hz_in(c, pl, IN_imOutIdx, i), __LINE__);
        }
    }
    // H - L1TxFullData of the first maxL1Tx slots goes into the SHA pre-image (:443-448)
    if (i < P.maxL1Tx) {
        uint32_t p0 = hz_pre_pos_l1(P) + (uint32_t)i * 624;
        for (int k = 0; k < 624; k++) pre.put(p0 + k, o.L1TxFullData.get(k));
    }
}

// ---- section D/E: rollupTx[i]  (:269-389, :427-431) ------------------------------------------
HZ_HD bool hz_roots_deferred(const SlotHook<MCompute>& hk) { return hk.smt != nullptr; }
#ifndef HZ_DEVICE_ONLY
inline bool hz_roots_deferred(const SlotHook<MLayout>&) { return false; }
#endif

// E (:383-389) / G (:427-431): the roots a slot produced against the intermediary inputs
TM HZ_FN void sec_rollup_root_checks(Comp<M>& c, const HzPlan& pl, int i, const V<M>& newStateRoot, const V<M>& newExitRoot) {
    if (!c.sanity()) return;
    if (i < pl.P.nTx - 1) {
        c.check(newStateRoot, hz_in(c, pl, IN_imStateRoot, i), __LINE__);
        c.check(newExitRoot, hz_in(c, pl, IN_imExitRoot, i), __LINE__);
    } else {
        c.check(newStateRoot, hz_in(c, pl, IN_imInitStateRootFee), __LINE__);
    }
}

TM HZ_FN void sec_rollup(Comp<M>& c, const HzPlan& pl, int i, const DecodeOut<M>& d, RollupTxOut<M>& o,
                          PreBits<M>& pre, const SlotHook<M>& hook) {
    const HzParams& P = pl.P;
    const int nTx = P.nTx, L = P.nLevels + 1;
    RollupTxIn<M> t;
    for (int k = 0; k < P.maxFeeTx; k++) {
        t.feePlanTokens[k] = hz_in_unless_dry(c, pl, IN_feePlanTokens, k);
        t.accFeeIn[k] = i == 0 ? Ops<M>::K(0) : hz_in_unless_dry(c, pl, IN_imAccFeeOut, i - 1, k);
    }
    for (int j = 0; j < 3; j++) {
        int q = i + j + 1;
        bool in = q < nTx;
        t.win.futV2[j] = in ? hz_in(c, pl, IN_txCompressedDataV2, q) : Ops<M>::K(0);
        t.win.futEth[j] = in ? hz_in(c, pl, IN_toEthAddr, q) : Ops<M>::K(0);
        t.win.futAy[j] = in ? hz_in(c, pl, IN_toBjjAy, q) : Ops<M>::K(0);
    }
    for (int j = 0; j < 4; j++) {
        int q = i - j - 1;
        bool in = q >= 0;
        t.win.pastV2[j] = in ? hz_in(c, pl, IN_txCompressedDataV2, q) : Ops<M>::K(0);
        t.win.pastEth[j] = in ? hz_in(c, pl, IN_toEthAddr, q) : Ops<M>::K(0);
        t.win.pastAy[j] = in ? hz_in(c, pl, IN_toBjjAy, q) : Ops<M>::K(0);
    }
    t.fromIdx = d.fromIdx; t.auxFromIdx = hz_in(c, pl, IN_auxFromIdx, i);
    t.toIdx = d.toIdx; t.auxToIdx = hz_in(c, pl, IN_auxToIdx, i);
    t.toBjjAy = hz_in(c, pl, IN_toBjjAy, i); t.toBjjSign = d.toBjjSign;
    t.toEthAddr = hz_in(c, pl, IN_toEthAddr, i);
    t.amount = d.amount; t.tokenID = d.tokenID; t.nonce = d.nonce; t.userFee = d.userFee;
    t.rqOffset = hz_in(c, pl, IN_rqOffset, i);
    t.onChain = hz_in(c, pl, IN_onChain, i); t.newAccount = hz_in(c, pl, IN_newAccount, i);
    t.rqTxCompressedDataV2 = hz_in(c, pl, IN_rqTxCompressedDataV2, i);
    t.rqToEthAddr = hz_in(c, pl, IN_rqToEthAddr, i); t.rqToBjjAy = hz_in(c, pl, IN_rqToBjjAy, i);
    t.sigL2Hash = d.sigL2Hash;
    t.s = hz_in(c, pl, IN_s, i); t.r8x = hz_in(c, pl, IN_r8x, i); t.r8y = hz_in(c, pl, IN_r8y, i);
    t.fromEthAddr = hz_in(c, pl, IN_fromEthAddr, i);
    t.fromBjjCompressed.clear(256);
    for (int k = 0; k < 256; k++) t.fromBjjCompressed.put(k, hz_in(c, pl, IN_fromBjjCompressed, i, k));
    t.loadAmountF = hz_in(c, pl, IN_loadAmountF, i);
    t.tokenID1 = hz_in(c, pl, IN_tokenID1, i); t.nonce1 = hz_in(c, pl, IN_nonce1, i); t.sign1 = hz_in(c, pl, IN_sign1, i);
    t.balance1 = hz_in(c, pl, IN_balance1, i); t.ay1 = hz_in(c, pl, IN_ay1, i); t.ethAddr1 = hz_in(c, pl, IN_ethAddr1, i);
    for (int k = 0; k < L; k++) t.siblings1[k] = hz_in_unless_dry(c, pl, IN_siblings1, i, k);
    t.isOld0_1 = hz_in(c, pl, IN_isOld0_1, i); t.oldKey1 = hz_in(c, pl, IN_oldKey1, i); t.oldValue1 = hz_in(c, pl, IN_oldValue1, i);
    t.tokenID2 = hz_in(c, pl, IN_tokenID2, i); t.nonce2 = hz_in(c, pl, IN_nonce2, i); t.sign2 = hz_in(c, pl, IN_sign2, i);
    t.balance2 = hz_in(c, pl, IN_balance2, i); t.newExit = hz_in(c, pl, IN_newExit, i); t.ay2 = hz_in(c, pl, IN_ay2, i);
    t.ethAddr2 = hz_in(c, pl, IN_ethAddr2, i);
    for (int k = 0; k < L; k++) t.siblings2[k] = hz_in_unless_dry(c, pl, IN_siblings2, i, k);
    t.isOld0_2 = hz_in(c, pl, IN_isOld0_2, i); t.oldKey2 = hz_in(c, pl, IN_oldKey2, i); t.oldValue2 = hz_in(c, pl, IN_oldValue2, i);
    t.oldStateRoot = i == 0 ? hz_in(c, pl, IN_oldStateRoot) : hz_in(c, pl, IN_imStateRoot, i - 1);
    t.oldExitRoot = i == 0 ? Ops<M>::K(0) : hz_in(c, pl, IN_imExitRoot, i - 1);
    t.sib1_idx = pl.in_off[IN_siblings1] + (uint32_t)i * pl.in_d1[IN_siblings1];
    t.sib2_idx = pl.in_off[IN_siblings2] + (uint32_t)i * pl.in_d1[IN_siblings2];
    RollupTx(c.sub(Nm("rollupTx", i)), P.nLevels, P.maxFeeTx, t, o, hook);
    if (c.sanity()) {
        // E (:383-389) / G (:427-431); with the processors handed off the roots are checked by the tail
        if (!hz_roots_deferred(hook)) sec_rollup_root_checks(c, pl, i, o.newStateRoot, o.newExitRoot);
        for (int k = 0; k < P.maxFeeTx; k++)
            c.check(o.accFeeOut[k], i < nTx - 1 ? hz_in(c, pl, IN_imAccFeeOut, i, k) : hz_in(c, pl, IN_imFinalAccFee, k), __LINE__);
    }
    if (hz_jobs_only(hook)) return;
    // H - L1L2TxData with the amount nullified (:452-461): what hasherInputs.L1L2TxsData[] holds
    const int nL2 = 2 * P.nLevels + 48;
    uint32_t p0 = hz_pre_pos_l2(P) + (uint32_t)i * nL2;
    for (int j = 0; j < nL2; j++) {
        if (j >= 2 * P.nLevels && j < 2 * P.nLevels + 40) pre.put(p0 + j, d.L1L2TxData.get(j) * (1 - o.isAmountNullified));
        else pre.put(p0 + j, d.L1L2TxData.get(j));
    }
}

// ---- section F/G: feeTx[i]  (:393-426) --------------------------------------------------------
TM HZ_FN V<M> sec_fee(Comp<M>& c, const HzPlan& pl, int i, const FeeHook<M>& hook, bool deferred = false) {
    const HzParams& P = pl.P;
    V<M> sib[HZ_MAX_LEVELS];
    for (int k = 0; k < P.nLevels + 1; k++) sib[k] = hz_in(c, pl, IN_siblings3, i, k);
    V<M> r = FeeTx(c.sub(Nm("feeTx", i)), P.nLevels,
                   i == 0 ? hz_in(c, pl, IN_imInitStateRootFee) : hz_in(c, pl, IN_imStateRootFee, i - 1),
                   hz_in(c, pl, IN_feePlanTokens, i), hz_in(c, pl, IN_feeIdxs, i), hz_in(c, pl, IN_imFinalAccFee, i),
                   hz_in(c, pl, IN_tokenID3, i), hz_in(c, pl, IN_nonce3, i), hz_in(c, pl, IN_sign3, i),
                   hz_in(c, pl, IN_balance3, i), hz_in(c, pl, IN_ay3, i), hz_in(c, pl, IN_ethAddr3, i), sib, hook);
    if (!deferred && c.sanity() && i < P.maxFeeTx - 1) c.check(r, hz_in(c, pl, IN_imStateRootFee, i), __LINE__);
    return r;
}

#ifdef HZ_DEVICE_ONLY
// ---- device only: what sec_rollup needs of decodeTx[i], values only (k_smtjob) ----------------------
// The fields DecodeTx cuts out of txCompressedData (src/decode-tx.circom:96-162) and the decoded amount
// (:166-172, through the same DecodeFloatBin template on a dry context); every one is a signal there.
__device__ __forceinline__ Fr hz_fr_field(const Fr& x, int from, int n) {
    Fr r = fr_zero();
    for (int k = 0; k < n; k++) r.l[k >> 5] |= fr_bit(x, from + k) << (k & 31);
    return r;
}
static __device__ __noinline__ void hz_decode_lite(Comp<MCompute>& c, const HzPlan& pl, int i, DecodeOut<MCompute>& o) {
    const Fr txc = hz_in(c, pl, IN_txCompressedData, i).v;
    auto sigv = [](const Fr& f) { VC r; r.v = f; r.sig = true; return r; };
    o.fromIdx = sigv(hz_fr_field(txc, 48, 48)); o.toIdx = sigv(hz_fr_field(txc, 96, 48));
    o.tokenID = sigv(hz_fr_field(txc, 144, 32)); o.nonce = sigv(hz_fr_field(txc, 176, 40));
    o.userFee = sigv(hz_fr_field(txc, 216, 8)); o.toBjjSign = sigv(hz_fr_field(txc, 224, 1));
    BV<MCompute> am = Num2Bits(c.sub("n2bAmount"), 40, hz_in(c, pl, IN_amountF, i));
    o.amount = DecodeFloatBin(c.sub("dfAmount"), am);
    o.sigL2Hash = sigv(fr_zero());                  // only the signature reads it (k_edjob)
    o.L1L2TxData.clear(2 * pl.P.nLevels + 48); o.L1TxFullData.clear(624);
    o.txCompressedDataV2 = o.outIdx = o.onChain = sigv(fr_zero());
}

// ---- device only: bodies and tops of the SMT processors as their own sections -------------------
__device__ __forceinline__ VC hz_sigv(const Fr& f, bool sig) { VC r; r.v = f; r.sig = sig; return r; }

// head of one SMTProcessor body from its job record (k_smt): everything before the levels; the levels
// themselves are evaluated by k_smt_levels from the HzLevJob and the state-machine tuples left here
static __device__ __noinline__ void sec_smt_head(Fr* w, HzErrBuf* err, int nLevels1, const HzSmtJob& j, HzSmtMid& out, HzLevJob& lev, Fr* smst) {
    Comp<MCompute> c(w, j.cur, err);
    c.solo = true; c.skip_hash = true;
    VC sib[HZ_MAX_LEVELS];
    for (int k = 0; k < nLevels1; k++) sib[k] = c.load(j.sib + k);
    VC f[2];
    f[0] = hz_sigv(j.fnc0, j.sigmask & 32u); f[1] = hz_sigv(j.fnc1, j.sigmask & 64u);
    SMTMid<MCompute> mid;
    SMTHead<MCompute> hd;
    // the old root is not needed by the body (wired signal: no entry); the top gets the real one
    SMTProcessorHead<MCompute>(c, nLevels1, hz_sigv(fr_zero(), true), sib, hz_sigv(j.oldKey, j.sigmask & 1u),
                               hz_sigv(j.oldValue, j.sigmask & 2u), hz_sigv(j.isOld0, j.sigmask & 4u),
                               hz_sigv(j.newKey, j.sigmask & 8u), hz_sigv(j.newValue, j.sigmask & 16u), f, mid, hd);
    // one thread per processor (the leaf hashes are k_leafhash's: Comp::skip_hash)
    for (int k = 0; k < nLevels1 * 6; k++) smst[k] = hd.sm[k / 6][k % 6].v;
    {
        uint64_t bits = 0;
        for (int k = 0; k < nLevels1; k++) bits |= (uint64_t)hd.n2bNew.bit(k) << k;
        lev.lrbits = bits; lev.cur = c.cur; lev.sib = j.sib;           // lev.o1, lev.n1: k_leafhash
        out.enabled = mid.enabled.v; out.oldKey = mid.oldKey.v; out.newKey = mid.newKey.v; out.fnc0 = mid.fnc0.v;
        out.fnc1 = mid.fnc1.v; out.cur = c.cur + hz_smt_levels_size(nLevels1); out.newRootIdx = mid.newRootIdx;
    }
}
__device__ __forceinline__ VC hz_smt_top(Comp<MCompute>& c, const HzSmtMid& m, const VC& oldRoot) {
    SMTMid<MCompute> mid;
    mid.oldRoot = oldRoot; mid.enabled = hz_sigv(m.enabled, true); mid.oldKey = hz_sigv(m.oldKey, true);
    mid.newKey = hz_sigv(m.newKey, true); mid.fnc0 = hz_sigv(m.fnc0, true); mid.fnc1 = hz_sigv(m.fnc1, true);
    mid.oc = hz_sigv(m.oc, true); mid.nc = hz_sigv(m.nc, true); mid.newRootIdx = m.newRootIdx;
    c.cur = m.cur;
    return SMTProcessorTop<MCompute>(c, mid);
}
// tops of processor 1 and 2, s3, s4, s5 and the root checks of slot i (k_tail)
static __device__ __noinline__ void sec_rollup_tail(Fr* w, HzErrBuf* err, const HzPlan& pl, int i, const HzSmtMid* mids, const Fr& isExit_) {
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, mids[0].cur, err, nullptr, &pend);
    VC oldStateRoot = i == 0 ? hz_in(c, pl, IN_oldStateRoot) : hz_in(c, pl, IN_imStateRoot, i - 1);
    VC oldExitRoot = i == 0 ? hz_sigv(fr_zero(), true) : hz_in(c, pl, IN_imExitRoot, i - 1);
    VC isExit = hz_sigv(isExit_, true);
    VC p1 = hz_smt_top(c, mids[0], oldStateRoot);
    VC s3 = Mux1(c.sub("s3"), p1, oldExitRoot, isExit);
    VC p2 = hz_smt_top(c, mids[1], s3);
    VC newStateRoot = Mux1(c.sub("s4"), p2, p1, isExit);
    VC newExitRoot = Mux1(c.sub("s5"), oldExitRoot, p2, isExit);
    sec_rollup_root_checks(c, pl, i, newStateRoot, newExitRoot);
    c.flush_pending();
}
// top of the processor of fee tx i
static __device__ __noinline__ void sec_fee_tail(Fr* w, HzErrBuf* err, const HzPlan& pl, int i, const HzSmtMid& mid) {
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, mid.cur, err, nullptr, &pend);
    VC oldStateRoot = i == 0 ? hz_in(c, pl, IN_imInitStateRootFee) : hz_in(c, pl, IN_imStateRootFee, i - 1);
    VC r = hz_smt_top(c, mid, oldStateRoot);
    if (c.sanity() && i < pl.P.maxFeeTx - 1) c.check(r, hz_in(c, pl, IN_imStateRootFee, i), __LINE__);
    c.flush_pending();
}
#endif

// ---- section H: hasherInputs (src/hash-inputs.circom:23-185) ----------------------------------
// head: everything before the first sha256compression; the pre-image gets its remaining bits
TM HZ_FN void sec_hasher_head(Comp<M>& c, const HzPlan& pl, PreBits<M>& pre, uint32_t* off_blocks, W32<M>* ha) {
    const HzParams& P = pl.P;
    const int nL = P.nLevels, nL2 = 2 * nL + 48;
    V<M> oldLastIdx = c.def("oldLastIdx", hz_in(c, pl, IN_oldLastIdx));
    V<M> newLastIdx = c.def("newLastIdx", c.load(pl.idx_newLastIdx));
    V<M> oldStateRoot = c.def("oldStateRoot", hz_in(c, pl, IN_oldStateRoot));
    V<M> newStateRoot = c.def("newStateRoot", c.load(pl.idx_newStateRoot));
    V<M> newExitRoot = c.def("newExitRoot", c.load(pl.idx_newExitRoot));
    uint32_t p1 = hz_pre_pos_l1(P), p2 = hz_pre_pos_l2(P);
    for (uint32_t k = 0; k < (uint32_t)P.maxL1Tx * 624; k++) c.alias(Nm("L1TxsFullData", (int)k), pre.get(p1 + k));
    for (int i = 0; i < P.nTx; i++)
        for (int j = 0; j < nL2; j++) {
            uint32_t pos = p2 + (uint32_t)i * nL2 + j;
            if (j >= 2 * nL && j < 2 * nL + 40) pre.put(pos, c.def(Nm("L1L2TxsData", i * nL2 + j), Ops<M>::K(pre.bit(pos))));
            else c.alias(Nm("L1L2TxsData", i * nL2 + j), pre.get(pos));
        }
    for (int i = 0; i < P.maxFeeTx; i++) c.alias(Nm("feeTxsData", i), hz_in(c, pl, IN_feeIdxs, i));
    V<M> chain = c.def("globalChainID", hz_in(c, pl, IN_globalChainID));
    V<M> batch = c.def("currentNumBatch", hz_in(c, pl, IN_currentNumBatch));
    // hashInputsOut is wired to n2bHashInputsOut.out
    uint32_t pos = 0;
    {
        BV<M> b = Num2Bits(c.sub("n2bOldLastIdx"), 48, oldLastIdx);
        if (c.sanity()) { uint32_t s = 0; for (int i = nL; i < 48; i++) s += b.bit(i); c.check0(Ops<M>::K(s), __LINE__); }
        for (int k = 0; k < 48; k++) pre.put(pos + k, b.get(47 - k));
        pos += 48;
    }
    {
        BV<M> b = Num2Bits(c.sub("n2bNewLastIdx"), 48, newLastIdx);
        if (c.sanity()) { uint32_t s = 0; for (int i = nL; i < 48; i++) s += b.bit(i); c.check0(Ops<M>::K(s), __LINE__); }
        for (int k = 0; k < 48; k++) pre.put(pos + k, b.get(47 - k));
        pos += 48;
    }
    const char* rn[3] = {"n2bOldStateRoot", "n2bNewStateRoot", "n2bNewExitRoot"};
    for (int r = 0; r < 3; r++) {
        BV<M> b = Num2Bits(c.sub(rn[r]), 256, r == 0 ? oldStateRoot : (r == 1 ? newStateRoot : newExitRoot));
        for (int k = 0; k < 256; k++) pre.put(pos + k, b.get(255 - k));
        pos += 256;
    }
    pos = p2 + (uint32_t)P.nTx * nL2;
    for (int i = 0; i < P.maxFeeTx; i++) {
        BV<M> b = Num2Bits(c.sub(Nm("n2bFeeTxsData", i)), 48, hz_in(c, pl, IN_feeIdxs, i));
        if (c.sanity()) { uint32_t s = 0; for (int k = nL; k < 48; k++) s += b.bit(k); c.check0(Ops<M>::K(s), __LINE__); }
        for (int k = 0; k < nL; k++) pre.put(pos + k, b.get(nL - 1 - k));
        pos += nL;
    }
    {
        BV<M> b = Num2Bits(c.sub("n2bChainID"), 16, chain);
        for (int k = 0; k < 16; k++) pre.put(pos + k, b.get(15 - k));
        pos += 16;
    }
    {
        BV<M> b = Num2Bits(c.sub("n2bCurrentNumBatch"), 32, batch);
        for (int k = 0; k < 32; k++) pre.put(pos + k, b.get(31 - k));
        pos += 32;
    }
    // Sha256(nBits): `in`, `paddedIn[k < nBits]` are wires; padding bits and H constants own entries
    Comp<M> sh = c.sub("inputsHasher");
    const uint32_t nBits = pl.nbits, total = pl.nblocks * 512;
    for (uint32_t k = 0; k < nBits; k++) { sh.alias(Nm("in", (int)k), pre.get(k)); sh.alias(Nm("paddedIn", (int)k), pre.get(k)); }
    pre.put(nBits, sh.def(Nm("paddedIn", (int)nBits), Ops<M>::K(1)));
    for (uint32_t k = nBits + 1; k < total - 64; k++) pre.put(k, sh.def(Nm("paddedIn", (int)k), Ops<M>::K(0)));
    for (uint32_t k = total - 64; k < total; k++) {
        uint32_t kk = total - k - 1;
        pre.put(k, sh.def(Nm("paddedIn", (int)k), Ops<M>::K(kk < 32 ? ((nBits >> kk) & 1u) : 0)));
    }
    const char* han[8] = {"ha0", "ha1", "ha2", "ha3", "ha4", "ha5", "ha6", "ha7"};
    for (int j = 0; j < 8; j++) {
        Comp<M> hc = sh.sub(han[j]);
        ha[j] = w32_def(hc, "out", hz_sha_h(j), 32);
    }
    *off_blocks = sh.cur;
    sh.cur += pl.nblocks * (uint32_t)HZ_SZ_SHA_BLOCK;
    // out[] of Sha256 are wires to the last block; tail follows
}

// one sha256compression; `sh` is the Sha256 component positioned at the block's base
TM HZ_FN void sec_sha_block(Comp<M>& sh, int b, const W32<M>* hin, const W32<M>* msg, W32<M>* outw) {
    Sha256compression(sh.sub(Nm("sha256compression", b)), hin, msg, outw);
}

HZ_HD VC w32_get(const W32<MCompute>& x, int k) { VC r; r.v = fr_from_u64((x.v >> k) & 1u); r.sig = true; return r; }
#ifndef HZ_DEVICE_ONLY
inline VL w32_get(const W32<MLayout>& x, int k) { VL r; r.idx = x.idx[k]; return r; }
#endif

// tail: digest bits -> field element (src/hash-inputs.circom:179-184); `c` = hasherInputs after
// the blocks, `sh` = the Sha256 component (its out[] are wires to the last block)
TM HZ_FN V<M> sec_hasher_tail(Comp<M>& c, Comp<M>& sh, const W32<M>* digest) {
    BV<M> bits; bits.clear(256);
    for (int q = 0; q < 256; q++) {
        V<M> h = w32_get(digest[q >> 5], 31 - (q & 31));      // h[q], MSB first
        sh.alias(Nm("out", q), h);
        bits.put(255 - q, h);
    }
    V<M> o = Bits2Num(c.sub("n2bHashInputsOut"), bits, 0, 256);
    return c.alias("hashInputsOut", o);
}
