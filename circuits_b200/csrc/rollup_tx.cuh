// RollupTx(nLevels, maxFeeTx) as a circuit of its own (/root/reference/src/rollup-tx.circom:78-591 compiled with
// `component main = RollupTx(nLevels, maxFeeTx)`, the way /root/reference/test/rollup-tx.test.js:56-920 drives it;
// config 2 of BASELINE.json).  The template is the very one RollupMain instantiates per slot
// (templates_hermez.cuh: RollupTx) evaluated as the MAIN component: its own signals are never aliased away
// (Comp::is_main), so every input, output and intermediate signal of the template owns a witness entry at its
// declaration point -- inputs and the four outputs interleave, which is why the packed input vector is scattered
// into the witness through `in_idx` instead of copied as one run.
// Instances are independent: the device gives each its own lane group and witness range (hzw_calculate_batch).
#pragma once
#include "rollup_main.cuh"

// main-component inputs in declaration order (src/rollup-tx.circom:93-170); len: 1, or F = maxFeeTx, or L = nLevels + 1
struct HzTInDesc { const char* name; int kind; int n; };    // kind 0: n entries, 1: maxFeeTx, 2: nLevels + 1
static const HzTInDesc HZ_ROLLUPTX_INPUTS[] = {
    {"feePlanTokens", 1, 0}, {"accFeeIn", 1, 0},
    {"futureTxCompressedDataV2", 0, 3}, {"pastTxCompressedDataV2", 0, 4}, {"futureToEthAddr", 0, 3}, {"pastToEthAddr", 0, 4},
    {"futureToBjjAy", 0, 3}, {"pastToBjjAy", 0, 4},
    {"fromIdx", 0, 1}, {"auxFromIdx", 0, 1}, {"toIdx", 0, 1}, {"auxToIdx", 0, 1}, {"toBjjAy", 0, 1}, {"toBjjSign", 0, 1},
    {"toEthAddr", 0, 1}, {"amount", 0, 1}, {"tokenID", 0, 1}, {"nonce", 0, 1}, {"userFee", 0, 1}, {"rqOffset", 0, 1},
    {"onChain", 0, 1}, {"newAccount", 0, 1}, {"rqTxCompressedDataV2", 0, 1}, {"rqToEthAddr", 0, 1}, {"rqToBjjAy", 0, 1},
    {"sigL2Hash", 0, 1}, {"s", 0, 1}, {"r8x", 0, 1}, {"r8y", 0, 1}, {"fromEthAddr", 0, 1}, {"fromBjjCompressed", 0, 256},
    {"loadAmountF", 0, 1}, {"tokenID1", 0, 1}, {"nonce1", 0, 1}, {"sign1", 0, 1}, {"balance1", 0, 1}, {"ay1", 0, 1},
    {"ethAddr1", 0, 1}, {"siblings1", 2, 0}, {"isOld0_1", 0, 1}, {"oldKey1", 0, 1}, {"oldValue1", 0, 1},
    {"tokenID2", 0, 1}, {"nonce2", 0, 1}, {"sign2", 0, 1}, {"balance2", 0, 1}, {"newExit", 0, 1}, {"ay2", 0, 1},
    {"ethAddr2", 0, 1}, {"siblings2", 2, 0}, {"isOld0_2", 0, 1}, {"oldKey2", 0, 1}, {"oldValue2", 0, 1},
    {"oldStateRoot", 0, 1}, {"oldExitRoot", 0, 1},
};
#define HZ_ROLLUPTX_NIN ((int)(sizeof(HZ_ROLLUPTX_INPUTS) / sizeof(HZ_ROLLUPTX_INPUTS[0])))
HZ_HD int hz_rolluptx_in_len(int f, int nLevels, int maxFeeTx) {
    const int kind = HZ_ROLLUPTX_INPUTS[f].kind;
    return kind == 1 ? maxFeeTx : (kind == 2 ? nLevels + 1 : HZ_ROLLUPTX_INPUTS[f].n);
}
HZ_HD uint32_t hz_rolluptx_ninputs(int nLevels, int maxFeeTx) { return 2u * (uint32_t)maxFeeTx + 21u + 22u + 256u + 7u + 2u * ((uint32_t)nLevels + 1u) + 13u + 2u; }
// witness index of packed input k: inputs are declared in one run except for the outputs declared among them --
// accFeeOut[F] after accFeeIn, newStateRoot after oldStateRoot
HZ_HD uint32_t hz_rolluptx_in_index(uint32_t k, int nLevels, int maxFeeTx) {
    const uint32_t F = (uint32_t)maxFeeTx, n = hz_rolluptx_ninputs(nLevels, maxFeeTx);
    if (k < 2u * F) return 1u + k;
    if (k + 1u < n) return 1u + F + k;                      // everything up to oldStateRoot sits behind accFeeOut[F]
    return 1u + F + k + 1u;                                 // oldExitRoot: behind newStateRoot as well
}

TM struct HzInlineHook;
template <> struct HzInlineHook<MCompute> {
    static HZ_HD SlotHook<MCompute> make() { SlotHook<MCompute> h; h.eddsa = nullptr; h.smt = nullptr; h.isExit = nullptr; h.sz = nullptr; h.jobs_only = false; return h; }
};
#ifndef HZ_DEVICE_ONLY
template <> struct HzInlineHook<MLayout> {
    static SlotHook<MLayout> make() { SlotHook<MLayout> h; h.sz_out = nullptr; return h; }
};
#endif

// the whole main component; inputs already sit at their entries (hz_rolluptx_in_index), `c` starts at entry 1
TM HZ_FN void sec_rolluptx_main(Comp<M>& c, int nLevels, int maxFeeTx, RollupTxOut<M>& o) {
    RollupTxIn<M> t;
    uint32_t k = 0;
    auto next = [&]() { V<M> v = c.load(hz_rolluptx_in_index(k, nLevels, maxFeeTx)); k++; return v; };
    const int L = nLevels + 1;
    for (int j = 0; j < maxFeeTx; j++) t.feePlanTokens[j] = next();
    for (int j = 0; j < maxFeeTx; j++) t.accFeeIn[j] = next();
    for (int j = 0; j < 3; j++) t.win.futV2[j] = next();
    for (int j = 0; j < 4; j++) t.win.pastV2[j] = next();
    for (int j = 0; j < 3; j++) t.win.futEth[j] = next();
    for (int j = 0; j < 4; j++) t.win.pastEth[j] = next();
    for (int j = 0; j < 3; j++) t.win.futAy[j] = next();
    for (int j = 0; j < 4; j++) t.win.pastAy[j] = next();
    t.fromIdx = next(); t.auxFromIdx = next(); t.toIdx = next(); t.auxToIdx = next(); t.toBjjAy = next(); t.toBjjSign = next();
    t.toEthAddr = next(); t.amount = next(); t.tokenID = next(); t.nonce = next(); t.userFee = next(); t.rqOffset = next();
    t.onChain = next(); t.newAccount = next(); t.rqTxCompressedDataV2 = next(); t.rqToEthAddr = next(); t.rqToBjjAy = next();
    t.sigL2Hash = next(); t.s = next(); t.r8x = next(); t.r8y = next(); t.fromEthAddr = next();
    t.fromBjjCompressed.clear(256);
    for (int j = 0; j < 256; j++) t.fromBjjCompressed.put(j, next());
    t.loadAmountF = next();
    t.tokenID1 = next(); t.nonce1 = next(); t.sign1 = next(); t.balance1 = next(); t.ay1 = next(); t.ethAddr1 = next();
    t.sib1_idx = hz_rolluptx_in_index(k, nLevels, maxFeeTx);
    for (int j = 0; j < L; j++) t.siblings1[j] = next();
    t.isOld0_1 = next(); t.oldKey1 = next(); t.oldValue1 = next();
    t.tokenID2 = next(); t.nonce2 = next(); t.sign2 = next(); t.balance2 = next(); t.newExit = next(); t.ay2 = next(); t.ethAddr2 = next();
    t.sib2_idx = hz_rolluptx_in_index(k, nLevels, maxFeeTx);
    for (int j = 0; j < L; j++) t.siblings2[j] = next();
    t.isOld0_2 = next(); t.oldKey2 = next(); t.oldValue2 = next();
    t.oldStateRoot = next(); t.oldExitRoot = next();
    SlotHook<M> hook = HzInlineHook<M>::make();      // signature and processors evaluated in place
    c.is_main = true;
    RollupTx(static_cast<Comp<M>&&>(c), nLevels, maxFeeTx, t, o, hook);
}

#ifndef HZ_DEVICE_ONLY
struct HzTLayout { int nLevels = 0, maxFeeTx = 0; uint32_t nvars = 0, n_inputs = 0; };
// layout walk of the standalone circuit: sizes, and names through the sink
inline void hz_rolluptx_walk(int nLevels, int maxFeeTx, LayoutSink* sink, HzTLayout& out) {
    out.nLevels = nLevels; out.maxFeeTx = maxFeeTx; out.n_inputs = hz_rolluptx_ninputs(nLevels, maxFeeTx);
    Comp<MLayout> c(1, sink, "main");
    RollupTxOut<MLayout> o;
    sec_rolluptx_main(c, nLevels, maxFeeTx, o);
    out.nvars = c.cur;
}
#endif
