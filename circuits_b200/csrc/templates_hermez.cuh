// hermeznetwork/circuits templates (/root/reference/src/**/*.circom), generic over the
// evaluation mode (ctx.cuh).  Each function follows one template, citing file:line; signal
// order, wiring and `===` checks as in oracle/hermez.py.
#pragma once
#include "templates_circomlib.cuh"

// feeShiftTable, src/compute-fee.circom:105-109 (protocol constants, copied as data).  The device reads
// its own copy in global memory: a table local to the function is rebuilt on the thread's stack at
// every call (it was a fifth of k_front's time).
static const uint64_t hz_fee_table_host[256] = {
#include "fee_table.inc"
};
#if defined(__CUDACC__)
__device__ const uint64_t hz_fee_table_dev[256] = {
#include "fee_table.inc"
};
#endif
HZ_HD uint64_t hz_fee_shift(int i) {
#if defined(__CUDA_ARCH__)
    return hz_fee_table_dev[i];
#else
    return hz_fee_table_host[i];
#endif
}

// ---------------------------------------------------------------------------
// src/lib/decode-float.circom:12-44
// ---------------------------------------------------------------------------
TM HZ_FN V<M> DecodeFloatBin(Comp<M>&& c, const BV<M>& bits) {
    BV<M> ins; ins.clear(40);
    for (int i = 0; i < 40; i++) ins.put(i, c.def(Nm("in", i), bits.get(i)));
    uint32_t out = c.sig("out");
    for (int i = 0; i < 35; i++) c.alias(Nm("m", i), ins.get(i));
    for (int i = 0; i < 5; i++) c.alias(Nm("e", i), ins.get(35 + i));
    uint32_t pe = c.sigs("pe", 5);
    V<M> p = c.set(pe, 9 * ins.get(35) + 1);
    const uint64_t sq[5] = {0, 100ull, 10000ull, 100000000ull, 10000000000000000ull};
    for (int i = 1; i < 5; i++) p = c.set(pe + i, (p * sq[i] - p) * ins.get(35 + i) + p);
    c.alias("scale10", p);
    return c.set(out, bits_lc<M>(ins, 0, 35) * p);
}

// src/lib/decode-float.circom:51-65 (used only stand-alone; kept for the unit vectors)
TM HZ_FN V<M> DecodeFloat(Comp<M>&& c, const V<M>& x) {
    V<M> in_ = c.def("in", x);
    BV<M> bits = Num2Bits(c.sub("n2b"), 40, in_);
    return c.alias("out", DecodeFloatBin(c.sub("decoder"), bits));
}

// ---------------------------------------------------------------------------
// src/lib/mux256.circom:10-51 -- table inputs are the constants of feeShiftTable
// ---------------------------------------------------------------------------
#define HZ_SZ_MUX256_FEE (8 + 256 + 17 * HZ_SZ_MUX4_WIRED)
HZ_HD bool hz_is_dry(const Comp<MCompute>& c) { return c.dry; }
#ifndef HZ_DEVICE_ONLY
inline bool hz_is_dry(const Comp<MLayout>&) { return false; }
#endif
TM HZ_FN V<M> Mux256Fee(Comp<M>&& c, const V<M>* s) {
    if (hz_is_dry(c)) {
        // values only: the table entry the eight selector bits point at
        uint32_t idx = 0;
        for (int i = 0; i < 8; i++) idx |= vlsb(s[i]) << i;
        V<M> r = Ops<M>::K(hz_fee_shift((int)idx));
        return c.known(0, r);
    }
    V<M> s_[8];
    for (int i = 0; i < 8; i++) s_[i] = c.def(Nm("s", i), s[i]);
    uint32_t in0 = c.cur;
    for (int i = 0; i < 256; i++) c.def(Nm("in", i), Ops<M>::K(hz_fee_shift(i)));
    V<M> outs[16];
    for (int i = 0; i < 16; i++) {
        V<M> k[16];
        for (int j = 0; j < 16; j++) k[j] = c.known(in0 + 16 * i + j, Ops<M>::K(hz_fee_shift(16 * i + j)));
        outs[i] = Mux4(c.sub(Nm("mux", i)), k, s_);
    }
    return c.alias("out", Mux4(c.sub(Nm("mux", 16)), outs, s_ + 4));
}

// ---------------------------------------------------------------------------
// src/compute-fee.circom:12-94
// ---------------------------------------------------------------------------
TM HZ_FN V<M> ComputeFee(Comp<M>&& c, const V<M>& feeSel, const V<M>& amount, const V<M>& applyFee) {
    V<M> feeSel_ = c.def("feeSel", feeSel);
    V<M> amount_ = c.def("amount", amount);
    V<M> applyFee_ = c.def("applyFee", applyFee);
    uint32_t feeOut = c.sig("feeOut");
    Comp<M> mux_c = c.sub_reserve("mux256", HZ_SZ_MUX256_FEE);
    BV<M> selbits = Num2Bits(c.sub("n2bFeeSel"), 8, feeSel_);
    V<M> ss[8];
    for (int i = 0; i < 8; i++) ss[i] = selbits.get(i) * applyFee_;
    V<M> muxo = Mux256Fee(HZ_MV(mux_c), ss);
    V<M> notShifted = c.def("feeOutNotShifted", muxo * amount_);
    V<M> applyShift = c.def("applyShift", 1 - selbits.get(6) * selbits.get(7));
    BV<M> b; b.clear(253);
    for (int i = 0; i < 253; i++) b.put(i, c.def(Nm("bitsFeeOut", i), Ops<M>::K(vbit(notShifted, i))));
    V<M> lcNot = bits_lc<M>(b, 0, 128);
    V<M> lcSh = bits_lc<M>(b, 60, 128);
    if (c.sanity()) {
        uint32_t ovNot = 0, ovSh = 0;
        for (int i = 128; i < 253; i++) ovNot += b.bit(i);
        for (int i = 188; i < 253; i++) ovSh += b.bit(i);
        if (vbit(notShifted, 253)) c.check(bits_lc<M>(b, 0, 253), notShifted, __LINE__);
        c.check0(applyShift * (uint64_t)ovSh, __LINE__);
        c.check0((1 - applyShift) * (uint64_t)ovNot, __LINE__);
    }
    return c.set(feeOut, applyShift * (lcSh - lcNot) + lcNot);
}

// ---------------------------------------------------------------------------
// src/balance-updater.circom:24-112
// ---------------------------------------------------------------------------
TM struct BalanceOut { V<M> newS, newR, isP2Nop, fee2Charge, isAmountNullified; };
TM HZ_FN BalanceOut<M> BalanceUpdater(Comp<M>&& c, const V<M>& oldS, const V<M>& oldR, const V<M>& amount,
                                       const V<M>& loadAmount, const V<M>& feeSelector, const V<M>& onChain,
                                       const V<M>& nop, const V<M>& nullifyLoadAmount, const V<M>& nullifyAmount) {
    V<M> bs = c.def("oldStBalanceSender", oldS);
    V<M> br = c.def("oldStBalanceReceiver", oldR);
    V<M> amount_ = c.def("amount", amount);
    V<M> loadAmount_ = c.def("loadAmount", loadAmount);
    V<M> feeSel = c.def("feeSelector", feeSelector);
    V<M> onChain_ = c.def("onChain", onChain);
    V<M> nop_ = c.def("nop", nop);
    V<M> nla = c.def("nullifyLoadAmount", nullifyLoadAmount);
    V<M> na = c.def("nullifyAmount", nullifyAmount);
    uint32_t newS = c.sig("newStBalanceSender"), newR = c.sig("newStBalanceReceiver"), isP2Nop = c.sig("isP2Nop");
    // fee2Charge and underflowOk are wired to sub-component outputs: no entries of their own
    uint32_t ian = c.sig("isAmountNullified");
    uint32_t ea1 = c.sig("effectiveAmount1"), ea2 = c.sig("effectiveAmount2"), ea3 = c.sig("effectiveAmount3");
    uint32_t ela1 = c.sig("effectiveLoadAmount1"), ela2 = c.sig("effectiveLoadAmount2");
    BalanceOut<M> o;
    o.fee2Charge = c.alias("fee2Charge", ComputeFee(c.sub("computeFee"), feeSel, amount_, (1 - onChain_) * (1 - nop_)));
    V<M> vela1 = c.set(ela1, loadAmount_ * onChain_);
    V<M> vela2 = c.set(ela2, vela1 * (1 - nla));
    V<M> vea1 = c.set(ea1, amount_ * (1 - nop_));
    V<M> vea2 = c.set(ea2, vea1 * (1 - na));
    BV<M> bits = Num2Bits(c.sub("n2bSender"), 193, Ops<M>::pow2(192) + bs + vela2 - vea2 - o.fee2Charge);
    V<M> underflowOk = c.alias("underflowOk", bits.get(192));
    HZ_CHECK0(c, (1 - underflowOk) * (1 - onChain_));
    V<M> vea3 = c.set(ea3, underflowOk * vea2);
    o.newS = c.set(newS, bs + vela2 - vea3 - o.fee2Charge);
    o.newR = c.set(newR, br + vea3);
    V<M> isz = IsZero(c.sub("effectiveAmountIsZero"), vea1);
    o.isAmountNullified = c.set(ian, 1 - (1 - na) * underflowOk);
    o.isP2Nop = c.set(isP2Nop, 1 - isz);
    return o;
}

// ---------------------------------------------------------------------------
// src/fee-accumulator.circom:17-45, 56-91
// ---------------------------------------------------------------------------
TM HZ_FN void FeeAccumulatorStep(Comp<M>&& c, const V<M>& tokenID, const V<M>& plan, const V<M>& selIn,
                                 const V<M>& fee, const V<M>& accIn, V<M>& selOut, V<M>& accOut) {
    V<M> tokenID_ = c.def("tokenID", tokenID);
    V<M> plan_ = c.def("feePlanTokenID", plan);
    V<M> selIn_ = c.def("isSelectedIn", selIn);
    V<M> fee_ = c.def("fee2Charge", fee);
    V<M> accIn_ = c.def("accFeeIn", accIn);
    uint32_t so = c.sig("isSelectedOut");
    V<M> eq = IsEqual(c.sub("isEqual"), tokenID_, plan_);
    selOut = c.set(so, 1 - (1 - eq) * (1 - selIn_));
    accOut = c.alias("accFeeOut", Mux1(c.sub("mux"), accIn_, accIn_ + fee_, eq * (1 - selIn_)));
}
#define HZ_MAX_FEE 64
TM HZ_FN void FeeAccumulator(Comp<M>&& c, int n, const V<M>& tokenID, const V<M>& fee2Charge,
                              const V<M>* feePlanTokenID, const V<M>* accFeeIn, V<M>* accFeeOut) {
    V<M> tokenID_ = c.def("tokenID", tokenID);
    V<M> fee = c.def("fee2Charge", fee2Charge);
    V<M> plan[HZ_MAX_FEE], accIn[HZ_MAX_FEE];
    for (int i = 0; i < n; i++) plan[i] = c.def(Nm("feePlanTokenID", i), feePlanTokenID[i]);
    // accFeeIn may be constants (slot 0): entries of their own then
    for (int i = 0; i < n; i++) accIn[i] = c.def(Nm("accFeeIn", i), accFeeIn[i]);
    V<M> sel = Ops<M>::K(0);
    for (int i = 0; i < n; i++) {
        V<M> so, ao;
        FeeAccumulatorStep(c.sub(Nm("chain", i)), tokenID_, plan[i], sel, fee, accIn[i], so, ao);
        sel = so;
        accFeeOut[i] = c.alias(Nm("accFeeOut", i), ao);
    }
}

// ---------------------------------------------------------------------------
// src/lib/hash-state.circom:18-40
// ---------------------------------------------------------------------------
TM HZ_FN V<M> HashState(Comp<M>&& c, const V<M>& tokenID, const V<M>& nonce, const V<M>& sign,
                        const V<M>& balance, const V<M>& ay, const V<M>& ethAddr) {
    V<M> in[4];
    V<M> tokenID_ = c.def("tokenID", tokenID);
    V<M> nonce_ = c.def("nonce", nonce);
    V<M> sign_ = c.def("sign", sign);
    in[1] = c.def("balance", balance);
    in[2] = c.def("ay", ay);
    in[3] = c.def("ethAddr", ethAddr);
    in[0] = c.def("e0", tokenID_ + nonce_ * Ops<M>::pow2(32) + sign_ * Ops<M>::pow2(72));
    return c.alias("out", Poseidon<M>(c.sub("hash"), in, 4));
}

// ---------------------------------------------------------------------------
// src/lib/utils-bjj.circom:12-28, 37-57
// ---------------------------------------------------------------------------
TM HZ_FN void BitsCompressed2AySign(Comp<M>&& c, const BV<M>& bits, V<M>& ay, V<M>& sign) {
    BV<M> ins; ins.clear(256);
    for (int i = 0; i < 256; i++) ins.put(i, c.def(Nm("bjjCompressed", i), bits.get(i)));
    ay = c.alias("ay", Bits2Num(c.sub("b2nAy"), ins, 0, 254));
    sign = c.alias("sign", ins.get(255));
}
TM HZ_FN V<M> AySign2Ax(Comp<M>&& c, const V<M>& ay, const V<M>& sign) {
    V<M> ay_ = c.def("ay", ay);
    V<M> sign_ = c.def("sign", sign);
    BV<M> bits = Num2Bits(c.sub("n2bAy"), 254, ay_);
    bits.n = 256;
    bits.put(254, Ops<M>::K(0));
    bits.put(255, sign_);
    PtV<M> pt = Bits2Point_Strict(c.sub("b2Point"), bits);
    return c.alias("ax", pt.x);
}

// ---------------------------------------------------------------------------
// src/rq-tx-verifier.circom:19-94
// ---------------------------------------------------------------------------
TM struct RqWin { V<M> futV2[3], pastV2[4], futEth[3], pastEth[4], futAy[3], pastAy[4]; };
TM HZ_FN void RqTxVerifier(Comp<M>&& c, const RqWin<M>& w, const V<M>& rqV2, const V<M>& rqEth,
                            const V<M>& rqAy, const V<M>& rqTxOffset) {
    V<M> fV2[3], pV2[4], fEth[3], pEth[4], fAy[3], pAy[4];
    for (int i = 0; i < 3; i++) fV2[i] = c.def(Nm("futureTxCompressedDataV2", i), w.futV2[i]);
    for (int i = 0; i < 4; i++) pV2[i] = c.def(Nm("pastTxCompressedDataV2", i), w.pastV2[i]);
    for (int i = 0; i < 3; i++) fEth[i] = c.def(Nm("futureToEthAddr", i), w.futEth[i]);
    for (int i = 0; i < 4; i++) pEth[i] = c.def(Nm("pastToEthAddr", i), w.pastEth[i]);
    for (int i = 0; i < 3; i++) fAy[i] = c.def(Nm("futureToBjjAy", i), w.futAy[i]);
    for (int i = 0; i < 4; i++) pAy[i] = c.def(Nm("pastToBjjAy", i), w.pastAy[i]);
    V<M> rqV2_ = c.def("rqTxCompressedDataV2", rqV2);
    V<M> rqEth_ = c.def("rqToEthAddr", rqEth);
    V<M> rqAy_ = c.def("rqToBjjAy", rqAy);
    V<M> off = c.def("rqTxOffset", rqTxOffset);
    Comp<M> m1 = c.sub_reserve("muxTxCompressedDataV2", 10);
    Comp<M> m2 = c.sub_reserve("muxToEthAddr", 10);
    Comp<M> m3 = c.sub_reserve("muxToBjjAy", 10);
    BV<M> bits = Num2Bits(c.sub("n2b"), 3, off);
    V<M> s[3];
    for (int i = 0; i < 3; i++) s[i] = bits.get(i);
    V<M> k[8];
    k[0] = Ops<M>::K(0);
    k[1] = fV2[0]; k[2] = fV2[1]; k[3] = fV2[2]; k[4] = pV2[3]; k[5] = pV2[2]; k[6] = pV2[1]; k[7] = pV2[0];
    V<M> o1 = Mux3(HZ_MV(m1), k, s);
    k[1] = fEth[0]; k[2] = fEth[1]; k[3] = fEth[2]; k[4] = pEth[3]; k[5] = pEth[2]; k[6] = pEth[1]; k[7] = pEth[0];
    V<M> o2 = Mux3(HZ_MV(m2), k, s);
    k[1] = fAy[0]; k[2] = fAy[1]; k[3] = fAy[2]; k[4] = pAy[3]; k[5] = pAy[2]; k[6] = pAy[1]; k[7] = pAy[0];
    V<M> o3 = Mux3(HZ_MV(m3), k, s);
    HZ_CHECK(c, o1, rqV2_);
    HZ_CHECK(c, o2, rqEth_);
    HZ_CHECK(c, o3, rqAy_);
}

// ---------------------------------------------------------------------------
// src/rollup-tx-states.circom:39-314
// ---------------------------------------------------------------------------
TM struct StatesOut {
    V<M> isP1Insert, isP2Insert, key1, key2, P1_fnc0, P1_fnc1, P2_fnc0, P2_fnc1, isExit, verifySignEnabled,
        nop, checkToEthAddr, checkToBjj, nullifyLoadAmount, nullifyAmount;
};
TM HZ_FN StatesOut<M> RollupTxStates(Comp<M>&& c, const V<M>& fromIdx_, const V<M>& toIdx_, const V<M>& toEthAddr_,
                                      const V<M>& auxFromIdx_, const V<M>& auxToIdx_, const V<M>& amount_,
                                      const V<M>& newExit_, const V<M>& loadAmount_, const V<M>& newAccount_,
                                      const V<M>& onChain_, const V<M>& fromEthAddr_, const V<M>& ethAddr1_,
                                      const V<M>& tokenID_, const V<M>& tokenID1_, const V<M>& tokenID2_) {
    V<M> fromIdx = c.def("fromIdx", fromIdx_);
    V<M> toIdx = c.def("toIdx", toIdx_);
    V<M> toEthAddr = c.def("toEthAddr", toEthAddr_);
    V<M> auxFromIdx = c.def("auxFromIdx", auxFromIdx_);
    V<M> auxToIdx = c.def("auxToIdx", auxToIdx_);
    V<M> amount = c.def("amount", amount_);
    V<M> newExit = c.def("newExit", newExit_);
    V<M> loadAmount = c.def("loadAmount", loadAmount_);
    V<M> na = c.def("newAccount", newAccount_);
    V<M> oc = c.def("onChain", onChain_);
    V<M> fromEthAddr = c.def("fromEthAddr", fromEthAddr_);
    V<M> ethAddr1 = c.def("ethAddr1", ethAddr1_);
    V<M> tokenID = c.def("tokenID", tokenID_);
    V<M> tokenID1 = c.def("tokenID1", tokenID1_);
    V<M> tokenID2 = c.def("tokenID2", tokenID2_);
    StatesOut<M> o;
    // outputs in declaration order; key1, key2, isExit and nop are wired to sub-component outputs
    uint32_t s_isP1Insert = c.sig("isP1Insert"), s_isP2Insert = c.sig("isP2Insert");
    uint32_t s_P1_fnc0 = c.sig("P1_fnc0"), s_P1_fnc1 = c.sig("P1_fnc1"), s_P2_fnc0 = c.sig("P2_fnc0"), s_P2_fnc1 = c.sig("P2_fnc1");
    uint32_t s_vse = c.sig("verifySignEnabled");
    uint32_t s_checkToEthAddr = c.sig("checkToEthAddr"), s_checkToBjj = c.sig("checkToBjj");
    uint32_t s_nullifyLoadAmount = c.sig("nullifyLoadAmount"), s_nullifyAmount = c.sig("nullifyAmount");

    V<M> finalFromIdx = c.alias("finalFromIdx", Mux1(c.sub("selectFromIdx"), fromIdx, auxFromIdx, oc * na));
    V<M> toIdxIsZero = IsZero(c.sub("toIdxIsZero"), toIdx);
    V<M> selectAuxToIdx = c.def("selectAuxToIdx", (1 - oc) * toIdxIsZero);
    V<M> finalToIdx = c.alias("finalToIdx", Mux1(c.sub("selectToIdx"), toIdx, auxToIdx, selectAuxToIdx));

    V<M> isToEthAddrAny = IsEqual(c.sub("isToEthAddrAny"), Ops<M>::pow2(160) - 1, toEthAddr);
    o.isExit = c.alias("isExit", IsEqual(c.sub("checkIsExit"), Ops<M>::K(1), finalToIdx));

    V<M> finalFromIdxIsZero = IsZero(c.sub("finalFromIdxIsZero"), finalFromIdx);
    V<M> isFinalFromIdx = c.def("isFinalFromIdx", 1 - finalFromIdxIsZero);
    V<M> loadAmountIsZero = IsZero(c.sub("loadAmountIsZero"), loadAmount);
    V<M> isLoadAmount = c.def("isLoadAmount", 1 - loadAmountIsZero);
    V<M> amountIsZero = IsZero(c.sub("amountIsZero"), amount);
    V<M> isAmount = c.def("isAmount", 1 - amountIsZero);

    HZ_CHECK0(c, (1 - oc) * isLoadAmount);
    HZ_CHECK0(c, (1 - oc) * na);

    o.isP1Insert = c.set(s_isP1Insert, oc * na);
    o.P1_fnc0 = c.set(s_P1_fnc0, o.isP1Insert * isFinalFromIdx);
    o.P1_fnc1 = c.set(s_P1_fnc1, (1 - o.isP1Insert) * isFinalFromIdx);
    V<M> k[4], s[2];
    k[0] = Ops<M>::K(0); k[1] = finalFromIdx; k[2] = finalFromIdx; k[3] = finalFromIdx;
    s[0] = o.P1_fnc0; s[1] = o.P1_fnc1;
    o.key1 = c.alias("key1", Mux2(c.sub("mux1"), k, s));

    o.isP2Insert = c.set(s_isP2Insert, o.isExit * newExit);
    o.P2_fnc0 = c.set(s_P2_fnc0, o.isP2Insert * isFinalFromIdx);
    o.P2_fnc1 = c.set(s_P2_fnc1, (1 - o.isP2Insert) * isFinalFromIdx);
    k[0] = Ops<M>::K(0); k[1] = finalToIdx; k[2] = Ops<M>::K(0); k[3] = finalFromIdx;
    s[0] = isAmount; s[1] = o.isExit;
    o.key2 = c.alias("key2", Mux2(c.sub("mux2"), k, s));

    o.verifySignEnabled = c.set(s_vse, (1 - oc) * isFinalFromIdx);
    o.nop = c.alias("nop", finalFromIdxIsZero);

    uint32_t s_tmpEth = c.sig("tmpCheckToEthAddr"), s_tmpBjj = c.sig("tmpCheckToBjj");
    V<M> tmpEth = c.set(s_tmpEth, (1 - isToEthAddrAny) * selectAuxToIdx);
    V<M> tmpBjj = c.set(s_tmpBjj, isToEthAddrAny * selectAuxToIdx);
    o.checkToEthAddr = c.set(s_checkToEthAddr, tmpEth * (1 - o.nop));
    o.checkToBjj = c.set(s_checkToBjj, tmpBjj * (1 - o.nop));

    V<M> onChainNotCreateAccount = c.def("onChainNotCreateAccount", (1 - na) * oc);
    uint32_t s_anEth = c.sig("applyNullifierEthAddr");
    V<M> shouldCheckEthAddr = c.def("shouldCheckEthAddr", onChainNotCreateAccount * isAmount);
    V<M> checkFromEthAddr = IsEqual(c.sub("checkFromEthAddr"), fromEthAddr, ethAddr1);
    V<M> anEth = c.set(s_anEth, shouldCheckEthAddr * (1 - checkFromEthAddr));

    uint32_t s_anT1 = c.sig("applyNullifierTokenID1");
    V<M> shouldCheckTokenID1 = c.alias("shouldCheckTokenID1", onChainNotCreateAccount);
    V<M> checkTokenID1 = IsEqual(c.sub("checkTokenID1"), tokenID, tokenID1);
    V<M> anT1 = c.set(s_anT1, shouldCheckTokenID1 * (1 - checkTokenID1));

    uint32_t s_anT2 = c.sig("applyNullifierTokenID2");
    V<M> sc20 = c.def("shouldCheckTokenID2_0", oc * isAmount);
    V<M> sc21 = c.def("shouldCheckTokenID2_1", sc20 * (1 - o.isP2Insert));
    V<M> checkTokenID2 = IsEqual(c.sub("checkTokenID2"), tokenID, tokenID2);
    V<M> anT2 = c.set(s_anT2, sc21 * (1 - checkTokenID2));

    o.nullifyLoadAmount = c.set(s_nullifyLoadAmount, anT1 * isLoadAmount);
    uint32_t s_na0 = c.sig("nullifyAmount_0"), s_act = c.sig("applyCheckTokenID1ToAmount");
    V<M> act = c.set(s_act, anT1 * isAmount);
    V<M> na0 = c.set(s_na0, 1 - (1 - anEth) * (1 - anT2));
    o.nullifyAmount = c.set(s_nullifyAmount, 1 - (1 - na0) * (1 - act));
    return o;
}

// ---------------------------------------------------------------------------
// src/decode-tx.circom:44-369
// ---------------------------------------------------------------------------
#define HZ_CONST_SIG 3322668559ull
TM struct DecodeOut {
    BVT<M, 5> L1L2TxData;       // nLevels*2 + 48 bits (<= 144), MSB first
    BVT<M, 20> L1TxFullData;    // 624 bits, MSB first
    V<M> txCompressedDataV2, outIdx, fromIdx, toIdx, tokenID, nonce, userFee, toBjjSign, amount, sigL2Hash, onChain;
};
TM struct DecodeIn {
    V<M> previousOnChain, txCompressedData, maxNumBatch, amountF, toEthAddr, toBjjAy, rqTxCompressedDataV2,
        rqToEthAddr, rqToBjjAy, fromEthAddr, loadAmountF, globalChainID, currentNumBatch, onChain, newAccount,
        auxFromIdx, auxToIdx, inIdx;
    BV<M> fromBjjCompressed;
};

TM HZ_FN void DecodeTx(Comp<M>&& c, int nLevels, const DecodeIn<M>& in, DecodeOut<M>& o) {
    V<M> previousOnChain = c.def("previousOnChain", in.previousOnChain);
    V<M> txCompressedData = c.def("txCompressedData", in.txCompressedData);
    V<M> maxNumBatch = c.def("maxNumBatch", in.maxNumBatch);
    V<M> amountF = c.def("amountF", in.amountF);
    V<M> toEthAddr = c.def("toEthAddr", in.toEthAddr);
    V<M> toBjjAy = c.def("toBjjAy", in.toBjjAy);
    V<M> rqV2 = c.def("rqTxCompressedDataV2", in.rqTxCompressedDataV2);
    V<M> rqEth = c.def("rqToEthAddr", in.rqToEthAddr);
    V<M> rqAy = c.def("rqToBjjAy", in.rqToBjjAy);
    const int nL2 = nLevels * 2 + 48;
    // L1L2TxData: only the userFee bits (last 8) are expressions; the rest are wired
    o.L1L2TxData.clear(nL2);
    uint32_t l1l2_fee = c.cur;
    for (int i = 0; i < 8; i++) c.sig(Nm("L1L2TxData", nLevels * 2 + 40 + i));
    // txCompressedDataV2 is wired to b2nTxCompressedDataV2.out
    V<M> fromEthAddr = c.def("fromEthAddr", in.fromEthAddr);
    BV<M> bjj; bjj.clear(256);
    for (int i = 0; i < 256; i++) bjj.put(i, c.def(Nm("fromBjjCompressed", i), in.fromBjjCompressed.get(i)));
    V<M> loadAmountF = c.def("loadAmountF", in.loadAmountF);
    V<M> globalChainID = c.def("globalChainID", in.globalChainID);
    V<M> currentNumBatch = c.def("currentNumBatch", in.currentNumBatch);
    V<M> oc = c.def("onChain", in.onChain);
    V<M> newAccount = c.def("newAccount", in.newAccount);
    V<M> auxFromIdx = c.def("auxFromIdx", in.auxFromIdx);
    V<M> auxToIdx = c.def("auxToIdx", in.auxToIdx);
    o.L1TxFullData.clear(624);
    uint32_t l1full = c.sigs("L1TxFullData", 624);
    V<M> inIdx = c.def("inIdx", in.inIdx);
    uint32_t s_outIdx = c.sig("outIdx");
    // constSig .. userFee, amount, sigL2Hash, toBjjSign are wired to sub-component outputs

    BV<M> d = Num2Bits(c.sub("n2bData"), 225, txCompressedData);
    V<M> constSig = c.alias("constSig", Bits2Num(c.sub("b2nConstSig"), d, 0, 32));
    V<M> chainID = c.alias("chainID", Bits2Num(c.sub("b2nChainID"), d, 32, 16));
    o.fromIdx = c.alias("fromIdx", Bits2Num(c.sub("b2nFrom"), d, 48, 48));
    if (c.sanity()) { uint32_t s = 0; for (int i = nLevels; i < 48; i++) s += d.bit(48 + i); c.check0(Ops<M>::K(s), __LINE__); }
    o.toIdx = c.alias("toIdx", Bits2Num(c.sub("b2nTo"), d, 96, 48));
    if (c.sanity()) { uint32_t s = 0; for (int i = nLevels; i < 48; i++) s += d.bit(96 + i); c.check0(Ops<M>::K(s), __LINE__); }
    o.tokenID = c.alias("tokenID", Bits2Num(c.sub("b2ntokenID"), d, 144, 32));
    o.nonce = c.alias("nonce", Bits2Num(c.sub("b2nNonce"), d, 176, 40));
    o.userFee = c.alias("userFee", Bits2Num(c.sub("b2nUserFee"), d, 216, 8));
    o.toBjjSign = c.alias("toBjjSign", d.get(224));

    BV<M> am = Num2Bits(c.sub("n2bAmount"), 40, amountF);
    o.amount = c.alias("amount", DecodeFloatBin(c.sub("dfAmount"), am));

    {
        V<M> noc = 1 - oc;
        BV<M> v2; v2.clear(217);
        for (int i = 0; i < 48; i++) v2.put(i, d.get(48 + i) * noc);
        for (int i = 0; i < 48; i++) v2.put(48 + i, d.get(96 + i) * noc);
        for (int i = 0; i < 40; i++) v2.put(96 + i, am.get(i) * noc);
        for (int i = 0; i < 32; i++) v2.put(136 + i, d.get(144 + i) * noc);
        for (int i = 0; i < 40; i++) v2.put(168 + i, d.get(176 + i) * noc);
        for (int i = 0; i < 8; i++) v2.put(208 + i, d.get(216 + i) * noc);
        v2.put(216, d.get(224));
        o.txCompressedDataV2 = c.alias("txCompressedDataV2", Bits2Num(c.sub("b2nTxCompressedDataV2"), v2, 0, 217));
    }

    V<M> toIdxIsZero = IsZero(c.sub("toIdxIsZero"), o.toIdx);
    V<M> selectToIdx = Mux1(c.sub("selectToIdx"), o.toIdx, auxToIdx, (1 - oc) * toIdxIsZero);
    BV<M> ft = Num2Bits(c.sub("n2bFinalToIdx"), nLevels, selectToIdx);
    for (int i = 0; i < nLevels; i++) o.L1L2TxData.put(nLevels - 1 - i, c.alias(Nm("L1L2TxData", nLevels - 1 - i), d.get(48 + i)));
    for (int i = 0; i < nLevels; i++) o.L1L2TxData.put(nLevels * 2 - 1 - i, c.alias(Nm("L1L2TxData", nLevels * 2 - 1 - i), ft.get(i)));
    for (int i = 0; i < 40; i++) o.L1L2TxData.put(nLevels * 2 + 40 - 1 - i, c.alias(Nm("L1L2TxData", nLevels * 2 + 40 - 1 - i), am.get(i)));
    for (int i = 0; i < 8; i++) o.L1L2TxData.put(nL2 - 1 - i, c.set(l1l2_fee + 7 - i, d.get(216 + i) * (1 - oc)));

    Comp<M> b2nElement1_c = c.sub_reserve("b2nElement1", 1);
    BV<M> eth = Num2Bits(c.sub("n2bToEthAddr"), 160, toEthAddr);
    BV<M> mnb = Num2Bits(c.sub("n2bMaxNumBatch"), 32, maxNumBatch);
    V<M> e1;
    {
        BV<M> eb; eb.clear(232);
        bv_copy(eb, 0, eth, 0, 160); bv_copy(eb, 160, am, 0, 40); bv_copy(eb, 200, mnb, 0, 32);
        e1 = Bits2Num(HZ_MV(b2nElement1_c), eb, 0, 232);
    }
    V<M> hin[6];
    hin[0] = txCompressedData; hin[1] = e1; hin[2] = toBjjAy; hin[3] = rqV2; hin[4] = rqEth; hin[5] = rqAy;
    o.sigL2Hash = c.alias("sigL2Hash", Poseidon<M>(c.sub("hashSig"), hin, 6));

    BV<M> feth = Num2Bits(c.sub("n2bFromEthAddr"), 160, fromEthAddr);
    for (int i = 0; i < 160; i++) o.L1TxFullData.put(160 - 1 - i, c.set(l1full + 160 - 1 - i, feth.get(i) * oc));
    for (int i = 0; i < 256; i++) o.L1TxFullData.put(416 - 1 - i, c.set(l1full + 416 - 1 - i, bjj.get(i) * oc));
    for (int i = 0; i < 48; i++) o.L1TxFullData.put(464 - 1 - i, c.set(l1full + 464 - 1 - i, d.get(48 + i) * oc));
    BV<M> la = Num2Bits(c.sub("n2bLoadAmountF"), 40, loadAmountF);
    for (int i = 0; i < 40; i++) o.L1TxFullData.put(504 - 1 - i, c.set(l1full + 504 - 1 - i, la.get(i) * oc));
    for (int i = 0; i < 40; i++) o.L1TxFullData.put(544 - 1 - i, c.set(l1full + 544 - 1 - i, am.get(i) * oc));
    for (int i = 0; i < 32; i++) o.L1TxFullData.put(576 - 1 - i, c.set(l1full + 576 - 1 - i, d.get(144 + i) * oc));
    for (int i = 0; i < 48; i++) o.L1TxFullData.put(624 - 1 - i, c.set(l1full + 624 - 1 - i, d.get(96 + i) * oc));

    V<M> fromIdxIsZero = IsZero(c.sub("fromIdxIsZero"), o.fromIdx);
    HZ_CHECK(c, oc * fromIdxIsZero, newAccount);
    o.outIdx = c.set(s_outIdx, inIdx + oc * newAccount);
    ForceEqualIfEnabled(c.sub("idxChecker"), oc * newAccount, auxFromIdx, o.outIdx);
    HZ_CHECK0(c, (1 - previousOnChain) * oc);
    ForceEqualIfEnabled(c.sub("chainIDChecker"), 1 - oc, globalChainID, chainID);
    ForceEqualIfEnabled(c.sub("constSigChecker"), 1 - oc, constSig, Ops<M>::K(HZ_CONST_SIG));
    V<M> mnbz = IsZero(c.sub("maxNumBatchIsZero"), maxNumBatch);
    V<M> ok = GreaterEqThan(c.sub("isMaxNumBatchOk"), 32, maxNumBatch, currentNumBatch);
    HZ_CHECK0(c, (1 - ok) * (1 - mnbz));
    o.onChain = oc;
}

// ---------------------------------------------------------------------------
// src/rollup-tx.circom:78-591
// ---------------------------------------------------------------------------
TM struct RollupTxIn {
    V<M> feePlanTokens[HZ_MAX_FEE], accFeeIn[HZ_MAX_FEE];
    RqWin<M> win;
    V<M> fromIdx, auxFromIdx, toIdx, auxToIdx, toBjjAy, toBjjSign, toEthAddr, amount, tokenID, nonce, userFee,
        rqOffset, onChain, newAccount, rqTxCompressedDataV2, rqToEthAddr, rqToBjjAy, sigL2Hash, s, r8x, r8y,
        fromEthAddr;
    BV<M> fromBjjCompressed;
    V<M> loadAmountF, tokenID1, nonce1, sign1, balance1, ay1, ethAddr1;
    V<M> siblings1[HZ_MAX_LEVELS];
    V<M> isOld0_1, oldKey1, oldValue1, tokenID2, nonce2, sign2, balance2, newExit, ay2, ethAddr2;
    V<M> siblings2[HZ_MAX_LEVELS];
    V<M> isOld0_2, oldKey2, oldValue2, oldStateRoot, oldExitRoot;
    uint32_t sib1_idx, sib2_idx;      // witness index of siblings1[i][0] / siblings2[i][0] (main inputs)
};
TM struct RollupTxOut { V<M> accFeeOut[HZ_MAX_FEE]; V<M> newStateRoot, newExitRoot, isAmountNullified; };

// ---- hand-offs between the device kernels ------------------------------------------------------
// The device does not walk a slot start to end in one thread group.  The front kernel evaluates
// DecodeTx and RollupTx up to the sub-graphs below and records their inputs as jobs; each
// sub-graph then runs in the kernel shaped for it:
//   phase F (AySign2Ax + EdDSAPoseidonVerifier)       -> HzEddsaJob  -> k_eddsa
//   SMTProcessor bodies (level hashes) of phase J     -> HzSmtJob    -> k_smt     -> HzSmtMid
//   SMTProcessor tops, s3/s4/s5 and the root checks   -> k_tail (needs processor 1's new root)
// Host builds (CPU port, layout walk) pass no job pointers and evaluate everything in place.
struct HzEddsaJob { Fr aySignature, signSignature, enabled, s1Ay, S, R8x, R8y, M; uint32_t cur; uint32_t pad[7]; };
// inputs of one SMTProcessor body; sigmask bit k set = input k is a wired signal (no entry of its
// own), k: 0 oldKey, 1 oldValue, 2 isOld0, 3 newKey, 4 newValue, 5 fnc[0], 6 fnc[1]
// hsel: the selector of the multiplexer in front of oldValue (old state hash against the input oldValue1 / 2): the
// state hashes are evaluated by k_hstate after the job is recorded, which completes oldValue += (1 - hsel) * hash
// ovidx: witness index of that multiplexer's output (HZ_NOIDX: oldValue is wired to the hash), which k_hstate writes
struct HzSmtJob { Fr oldKey, oldValue, isOld0, newKey, newValue, fnc0, fnc1, hsel; uint32_t cur, sib, sigmask, ovidx, pad[4]; };
// what a body leaves for its top; cur = witness cursor where the top continues
struct HzSmtMid { Fr enabled, oldKey, newKey, fnc0, fnc1, oc, nc; uint32_t cur, newRootIdx, pad[6]; };
// what the head of a body leaves for the levels kernel: leaf hashes, the level bits of the new key, where the
// level components start, where the siblings are; the state-machine tuples go to a scratch array
struct HzLevJob { Fr o1, n1; uint64_t lrbits; uint32_t cur, sib, pad[4]; };
struct HzSizes { uint32_t getAx, sigVerifier, phaseJK, p2_off, mulfix_off; };   // p2_off: processor2 start within phases J-K; mulfix_off: sigVerifier.mulFix from the start of phase F

TM struct SlotHook;
template <> struct SlotHook<MCompute> {
    HzEddsaJob* eddsa;              // null: evaluate phase F in place
    HzSmtJob* smt;                  // null: evaluate phases J, K in place; else 2 jobs (processor 1, 2)
    Fr* isExit;                     // with smt: the selector the tail needs
    const HzSizes* sz;
    bool jobs_only = false;         // dry evaluation for the SMT jobs alone (k_smtjob): the fee accumulator is skipped
};
template <> struct SlotHook<MLayout> {
    HzSizes* sz_out;                // sizes recorded by the layout walk (may be null)
    uint32_t* ed_cur = nullptr;     // where phase F of the slot being walked starts (may be null)
    uint32_t* jk_cur = nullptr;     // where phases J-K (processor1) of the slot being walked start (may be null)
};
HZ_HD bool hz_jobs_only(const SlotHook<MCompute>& hk) { return hk.jobs_only; }
#ifndef HZ_DEVICE_ONLY
inline bool hz_jobs_only(const SlotHook<MLayout>&) { return false; }
#endif

TM HZ_FN void hz_phaseF(Comp<M>& c, const V<M>& vse, const V<M>& signSignature, const V<M>& aySignature,
                        const V<M>& s1Ay, const V<M>& S, const V<M>& R8x, const V<M>& R8y, const V<M>& Msg) {
    V<M> ax = AySign2Ax(c.sub("getAx"), aySignature, signSignature);
    EdDSAPoseidonVerifier(c.sub("sigVerifier"), vse, ax, s1Ay, S, R8x, R8y, Msg);
}
HZ_HD void hz_phaseF_hook(Comp<MCompute>& c, const SlotHook<MCompute>& hk, const VC& vse, const VC& signSignature,
                          const VC& aySignature, const VC& s1Ay, const VC& S, const VC& R8x, const VC& R8y, const VC& Msg) {
    if (hk.jobs_only) { c.cur += hk.sz->getAx + hk.sz->sigVerifier; return; }
    if (!hk.eddsa) { hz_phaseF(c, vse, signSignature, aySignature, s1Ay, S, R8x, R8y, Msg); return; }
    HzEddsaJob* j = hk.eddsa;
    j->aySignature = aySignature.v; j->signSignature = signSignature.v; j->enabled = vse.v; j->s1Ay = s1Ay.v;
    j->S = S.v; j->R8x = R8x.v; j->R8y = R8y.v; j->M = Msg.v; j->cur = c.cur;
    c.cur += hk.sz->getAx + hk.sz->sigVerifier;
}
#ifndef HZ_DEVICE_ONLY
inline void hz_phaseF_hook(Comp<MLayout>& c, const SlotHook<MLayout>& hk, const VL& vse, const VL& signSignature,
                           const VL& aySignature, const VL& s1Ay, const VL& S, const VL& R8x, const VL& R8y, const VL& Msg) {
    uint32_t c0 = c.cur;
    V<MLayout> ax = AySign2Ax(c.sub("getAx"), aySignature, signSignature);
    uint32_t c1 = c.cur;
    EdDSAPoseidonVerifier(c.sub("sigVerifier"), vse, ax, s1Ay, S, R8x, R8y, Msg);
    if (hk.sz_out) { hk.sz_out->getAx = c1 - c0; hk.sz_out->sigVerifier = c.cur - c1; hk.sz_out->mulfix_off = HintGen<MLayout>::last_fix_begin() - c0; }
    if (hk.ed_cur) *hk.ed_cur = c0;
}
#endif

// phases J and K of RollupTx (src/rollup-tx.circom:537-590)
TM struct PhaseJKIn {
    V<M> oldStateRoot, oldExitRoot, isExit;
    const V<M>* siblings1; V<M> oldKey1, oldValue1, isOld0_1, key1, newValue1, f1[2];
    const V<M>* siblings2; V<M> oldKey2, oldValue2, isOld0_2, key2, newValue2, f2[2];
    uint32_t sib1_idx, sib2_idx;
    V<M> ins1, ins2;                 // isP1Insert, isP2Insert (what selects oldValue1 / oldValue2)
    uint32_t ov1_idx, ov2_idx;       // entries of s1OldValue.mux.out[0] / s2OldValue.mux.out[0]
};
TM HZ_FN void hz_phaseJK(Comp<M>& c, int nLevels, const PhaseJKIn<M>& in, V<M>& newStateRoot, V<M>& newExitRoot,
                         uint32_t* p2_start = nullptr) {
    // J - smt processors (:537-570)
    V<M> p1 = SMTProcessor(c.sub("processor1"), nLevels + 1, in.oldStateRoot, in.siblings1, in.oldKey1, in.oldValue1,
                           in.isOld0_1, in.key1, in.newValue1, in.f1);
    V<M> s3 = Mux1(c.sub("s3"), p1, in.oldExitRoot, in.isExit);
    if (p2_start) *p2_start = c.cur;
    V<M> p2 = SMTProcessor(c.sub("processor2"), nLevels + 1, s3, in.siblings2, in.oldKey2, in.oldValue2,
                           in.isOld0_2, in.key2, in.newValue2, in.f2);
    // K - select output roots (:577-590)
    // (a RollupTx evaluated as the main component has reserved entries for its two roots: RollupTx sets them)
    newStateRoot = Mux1(c.sub("s4"), p2, p1, in.isExit);
    newExitRoot = Mux1(c.sub("s5"), in.oldExitRoot, p2, in.isExit);
    if (!c.is_main) { newStateRoot = c.alias("newStateRoot", newStateRoot); newExitRoot = c.alias("newExitRoot", newExitRoot); }
}
HZ_HD void hz_smt_job(HzSmtJob* j, uint32_t cur, uint32_t sib, const VC& oldKey, const VC& oldValue, const VC& isOld0,
                      const VC& newKey, const VC& newValue, const VC* f) {
    j->oldKey = oldKey.v; j->oldValue = oldValue.v; j->isOld0 = isOld0.v; j->newKey = newKey.v; j->newValue = newValue.v;
    j->fnc0 = f[0].v; j->fnc1 = f[1].v; j->cur = cur; j->sib = sib;
    j->sigmask = (oldKey.sig ? 1u : 0u) | (oldValue.sig ? 2u : 0u) | (isOld0.sig ? 4u : 0u) | (newKey.sig ? 8u : 0u) |
                 (newValue.sig ? 16u : 0u) | (f[0].sig ? 32u : 0u) | (f[1].sig ? 64u : 0u);
}
HZ_HD void hz_phaseJK_hook(Comp<MCompute>& c, const SlotHook<MCompute>& hk, int nLevels, const PhaseJKIn<MCompute>& in,
                           VC& newStateRoot, VC& newExitRoot) {
    if (!hk.smt) { hz_phaseJK(c, nLevels, in, newStateRoot, newExitRoot); return; }
    // the bodies are jobs; processor 2 starts after processor 1 and s3 -- the tail recomputes the
    // cursors from the body records, here only the two component starts are needed
    (void)nLevels;
    hz_smt_job(&hk.smt[0], c.cur, in.sib1_idx, in.oldKey1, in.oldValue1, in.isOld0_1, in.key1, in.newValue1, in.f1);
    hz_smt_job(&hk.smt[1], c.cur + hk.sz->p2_off, in.sib2_idx, in.oldKey2, in.oldValue2, in.isOld0_2, in.key2, in.newValue2, in.f2);
    hk.smt[0].hsel = in.ins1.v; hk.smt[1].hsel = in.ins2.v;
    hk.smt[0].ovidx = in.ov1_idx; hk.smt[1].ovidx = in.ov2_idx;
    *hk.isExit = in.isExit.v;
    c.cur += hk.sz->phaseJK;
    newStateRoot = vc(fr_zero()); newExitRoot = vc(fr_zero());
}
#ifndef HZ_DEVICE_ONLY
inline void hz_phaseJK_hook(Comp<MLayout>& c, const SlotHook<MLayout>& hk, int nLevels, const PhaseJKIn<MLayout>& in,
                            VL& newStateRoot, VL& newExitRoot) {
    uint32_t c0 = c.cur, p2 = 0;
    hz_phaseJK(c, nLevels, in, newStateRoot, newExitRoot, &p2);
    if (hk.sz_out) { hk.sz_out->phaseJK = c.cur - c0; hk.sz_out->p2_off = p2 - c0; }
    if (hk.jk_cur) *hk.jk_cur = c0;
}
#endif

TM HZ_FN void RollupTx(Comp<M>&& c, int nLevels, int maxFeeTx, const RollupTxIn<M>& i, RollupTxOut<M>& o,
                        const SlotHook<M>& hook) {
    V<M> feePlanTokens[HZ_MAX_FEE], accFeeIn[HZ_MAX_FEE];
    for (int k = 0; k < maxFeeTx; k++) feePlanTokens[k] = c.def(Nm("feePlanTokens", k), i.feePlanTokens[k]);
    for (int k = 0; k < maxFeeTx; k++) accFeeIn[k] = c.def(Nm("accFeeIn", k), i.accFeeIn[k]);
    // accFeeOut[] are wired to feeAccumulator outputs -- unless this template is the main component (standalone
    // RollupTx, HZW_CIRCUIT_ROLLUP_TX): main's signals own their entries, at their declaration points
    const bool mn = c.is_main;
    const uint32_t slot_accOut = mn ? c.sigs("accFeeOut", maxFeeTx) : 0;
    RqWin<M> win;
    for (int k = 0; k < 3; k++) win.futV2[k] = c.def(Nm("futureTxCompressedDataV2", k), i.win.futV2[k]);
    for (int k = 0; k < 4; k++) win.pastV2[k] = c.def(Nm("pastTxCompressedDataV2", k), i.win.pastV2[k]);
    for (int k = 0; k < 3; k++) win.futEth[k] = c.def(Nm("futureToEthAddr", k), i.win.futEth[k]);
    for (int k = 0; k < 4; k++) win.pastEth[k] = c.def(Nm("pastToEthAddr", k), i.win.pastEth[k]);
    for (int k = 0; k < 3; k++) win.futAy[k] = c.def(Nm("futureToBjjAy", k), i.win.futAy[k]);
    for (int k = 0; k < 4; k++) win.pastAy[k] = c.def(Nm("pastToBjjAy", k), i.win.pastAy[k]);
    V<M> fromIdx = c.def("fromIdx", i.fromIdx), auxFromIdx = c.def("auxFromIdx", i.auxFromIdx);
    V<M> toIdx = c.def("toIdx", i.toIdx), auxToIdx = c.def("auxToIdx", i.auxToIdx);
    V<M> toBjjAy = c.def("toBjjAy", i.toBjjAy), toBjjSign = c.def("toBjjSign", i.toBjjSign);
    V<M> toEthAddr = c.def("toEthAddr", i.toEthAddr), amount = c.def("amount", i.amount);
    V<M> tokenID = c.def("tokenID", i.tokenID), nonce = c.def("nonce", i.nonce), userFee = c.def("userFee", i.userFee);
    V<M> rqOffset = c.def("rqOffset", i.rqOffset), oc = c.def("onChain", i.onChain), newAccount = c.def("newAccount", i.newAccount);
    V<M> rqV2 = c.def("rqTxCompressedDataV2", i.rqTxCompressedDataV2), rqEth = c.def("rqToEthAddr", i.rqToEthAddr);
    V<M> rqAy = c.def("rqToBjjAy", i.rqToBjjAy), sigL2Hash = c.def("sigL2Hash", i.sigL2Hash);
    V<M> s_ = c.def("s", i.s), r8x = c.def("r8x", i.r8x), r8y = c.def("r8y", i.r8y);
    V<M> fromEthAddr = c.def("fromEthAddr", i.fromEthAddr);
    BV<M> bjj; bjj.clear(256);
    for (int k = 0; k < 256; k++) bjj.put(k, c.def(Nm("fromBjjCompressed", k), i.fromBjjCompressed.get(k)));
    V<M> loadAmountF = c.def("loadAmountF", i.loadAmountF);
    V<M> tokenID1 = c.def("tokenID1", i.tokenID1), nonce1 = c.def("nonce1", i.nonce1), sign1 = c.def("sign1", i.sign1);
    V<M> balance1 = c.def("balance1", i.balance1), ay1 = c.def("ay1", i.ay1), ethAddr1 = c.def("ethAddr1", i.ethAddr1);
    V<M> siblings1[HZ_MAX_LEVELS], siblings2[HZ_MAX_LEVELS];
    for (int k = 0; k < nLevels + 1; k++) siblings1[k] = c.def(Nm("siblings1", k), i.siblings1[k]);
    V<M> isOld0_1 = c.def("isOld0_1", i.isOld0_1), oldKey1 = c.def("oldKey1", i.oldKey1), oldValue1 = c.def("oldValue1", i.oldValue1);
    V<M> tokenID2 = c.def("tokenID2", i.tokenID2), nonce2 = c.def("nonce2", i.nonce2), sign2 = c.def("sign2", i.sign2);
    V<M> balance2 = c.def("balance2", i.balance2), newExit = c.def("newExit", i.newExit), ay2 = c.def("ay2", i.ay2);
    V<M> ethAddr2 = c.def("ethAddr2", i.ethAddr2);
    for (int k = 0; k < nLevels + 1; k++) siblings2[k] = c.def(Nm("siblings2", k), i.siblings2[k]);
    V<M> isOld0_2 = c.def("isOld0_2", i.isOld0_2), oldKey2 = c.def("oldKey2", i.oldKey2), oldValue2 = c.def("oldValue2", i.oldValue2);
    V<M> oldStateRoot = c.def("oldStateRoot", i.oldStateRoot);
    const uint32_t slot_nsr = mn ? c.sig("newStateRoot") : 0;
    V<M> oldExitRoot = c.def("oldExitRoot", i.oldExitRoot);
    const uint32_t slot_ner = mn ? c.sig("newExitRoot") : 0, slot_ian = mn ? c.sig("isAmountNullified") : 0;
    const uint32_t slot_la = mn ? c.sig("loadAmount") : 0;
    // (as a sub-component) newStateRoot, newExitRoot, isAmountNullified, loadAmount are wired to sub-component outputs

    // A - compute tx-states (:183-209)
    BV<M> la_bits = Num2Bits(c.sub("n2bloadAmountF"), 40, loadAmountF);
    V<M> la_v = DecodeFloatBin(c.sub("dfLoadAmount"), la_bits);
    V<M> loadAmount = mn ? c.set(slot_la, la_v) : c.alias("loadAmount", la_v);
    StatesOut<M> st = RollupTxStates(c.sub("states"), fromIdx, toIdx, toEthAddr, auxFromIdx, auxToIdx, amount,
                                     newExit, loadAmount, newAccount, oc, fromEthAddr, ethAddr1, tokenID, tokenID1, tokenID2);
    // B - check request transaction fields (:213-231)
    RqTxVerifier(c.sub("rqTxVerifier"), win, rqV2, rqEth, rqAy, rqOffset);
    // C - check state fields (:237-292)
    ForceEqualIfEnabled(c.sub("nonceChecker"), 1 - oc, nonce, nonce1);
    ForceEqualIfEnabled(c.sub("checkToEthAddr"), 1 - (1 - st.checkToEthAddr) * (1 - st.checkToBjj), toEthAddr, ethAddr2);
    ForceEqualIfEnabled(c.sub("toBjjAyChecker"), st.checkToBjj, ay2, toBjjAy);
    ForceEqualIfEnabled(c.sub("toBjjSignChecker"), st.checkToBjj, sign2, toBjjSign);
    ForceEqualIfEnabled(c.sub("checkTokenID1"), 1 - oc, tokenID, tokenID1);
    ForceEqualIfEnabled(c.sub("checkTokenID2"), (1 - oc) * (1 - st.isP2Insert), tokenID, tokenID2);
    ForceEqualIfEnabled(c.sub("checkTokenID1L1"), st.isP1Insert, tokenID, tokenID1);
    ForceEqualIfEnabled(c.sub("fromEthAddrChecker"), st.isP1Insert, fromEthAddr, ethAddr1);
    // D - compute old hash states (:297-312)
    V<M> oldSt1Hash = HashState(c.sub("oldSt1Hash"), tokenID1, nonce1, sign1, balance1, ay1, ethAddr1);
    V<M> oldSt2Hash = HashState(c.sub("oldSt2Hash"), tokenID2, nonce2, sign2, balance2, ay2, ethAddr2);
    // E - signal processor selectors (:317-443)
    V<M> bay, bsign;
    BitsCompressed2AySign(c.sub("decodeFromBjj"), bjj, bay, bsign);
    const V<M> zero = Ops<M>::K(0);
    const V<M>& i1 = st.isP1Insert; const V<M>& i2 = st.isP2Insert;
    V<M> s1Balance = Mux1(c.sub("s1Balance"), balance1, zero, i1);
    V<M> s1Sign = Mux1(c.sub("s1Sign"), sign1, bsign, i1);
    V<M> s1Ay = Mux1(c.sub("s1Ay"), ay1, bay, i1);
    V<M> s1Nonce = Mux1(c.sub("s1Nonce"), nonce1, zero, i1);
    V<M> s1EthAddr = Mux1(c.sub("s1EthAddr"), ethAddr1, fromEthAddr, i1);
    V<M> s1TokenID = Mux1(c.sub("s1TokenID"), tokenID1, tokenID, i1);
    V<M> s1OldKey = Mux1(c.sub("s1OldKey"), st.key1, oldKey1, i1);
    const uint32_t ov1_idx = c.cur;       // all three inputs are signals: the multiplexer's only entry is its output
    V<M> s1OldValue = Mux1(c.sub("s1OldValue"), oldSt1Hash, oldValue1, i1);
    V<M> s2Balance = Mux1(c.sub("s2Balance"), balance2, zero, i2);
    V<M> s2Sign = Mux1(c.sub("s2Sign"), sign2, s1Sign, i2);
    V<M> s2Ay = Mux1(c.sub("s2Ay"), ay2, s1Ay, i2);
    V<M> s2Nonce = Mux1(c.sub("s2Nonce"), nonce2, zero, i2);
    V<M> s2EthAddr = Mux1(c.sub("s2EthAddr"), ethAddr2, s1EthAddr, i2);
    V<M> s2TokenID = Mux1(c.sub("s2TokenID"), tokenID2, s1TokenID, i2);
    V<M> s2OldKey = Mux1(c.sub("s2OldKey"), st.key2, oldKey2, i2);
    const uint32_t ov2_idx = c.cur;
    V<M> s2OldValue = Mux1(c.sub("s2OldValue"), oldSt2Hash, oldValue2, i2);
    // F - verify eddsa signature (:454-482)
    const V<M>& vse = st.verifySignEnabled;
    V<M> signSignature = Mux1(c.sub("signSignature"), zero, s1Sign, vse);
    V<M> aySignature = Mux1(c.sub("aySignature"), zero, s1Ay, vse);
    hz_phaseF_hook(c, hook, vse, signSignature, aySignature, s1Ay, s_, r8x, r8y, sigL2Hash);
    // G - update balances (:486-499)
    BalanceOut<M> bo = BalanceUpdater(c.sub("balanceUpdater"), s1Balance, s2Balance, amount, loadAmount, userFee,
                                      oc, st.nop, st.nullifyLoadAmount, st.nullifyAmount);
    o.isAmountNullified = mn ? c.set(slot_ian, bo.isAmountNullified) : c.alias("isAmountNullified", bo.isAmountNullified);
    // H - accumulate fees (:501-512)
    if (!hz_jobs_only(hook)) {
        FeeAccumulator(c.sub("feeAccumulator"), maxFeeTx, tokenID, bo.fee2Charge, feePlanTokens, accFeeIn, o.accFeeOut);
        for (int k = 0; k < maxFeeTx; k++) { if (mn) o.accFeeOut[k] = c.set(slot_accOut + (uint32_t)k, o.accFeeOut[k]); else c.alias(Nm("accFeeOut", k), o.accFeeOut[k]); }
    }
    // I - compute hash new states (:517-532)
    V<M> newSt1Hash = HashState(c.sub("newSt1Hash"), s1TokenID, s1Nonce + (1 - oc), s1Sign, bo.newS, s1Ay, s1EthAddr);
    V<M> newSt2Hash = HashState(c.sub("newSt2Hash"), s2TokenID, s2Nonce, s2Sign, bo.newR, s2Ay, s2EthAddr);
    // J, K - smt processors and output roots (:537-590)
    PhaseJKIn<M> jk;
    jk.oldStateRoot = oldStateRoot; jk.oldExitRoot = oldExitRoot; jk.isExit = st.isExit;
    jk.siblings1 = siblings1; jk.oldKey1 = s1OldKey; jk.oldValue1 = s1OldValue; jk.isOld0_1 = isOld0_1;
    jk.key1 = st.key1; jk.newValue1 = newSt1Hash; jk.f1[0] = st.P1_fnc0; jk.f1[1] = st.P1_fnc1;
    jk.siblings2 = siblings2; jk.oldKey2 = s2OldKey; jk.oldValue2 = s2OldValue; jk.isOld0_2 = isOld0_2;
    jk.key2 = st.key2; jk.newValue2 = newSt2Hash;
    jk.f2[0] = st.P2_fnc0 * bo.isP2Nop; jk.f2[1] = st.P2_fnc1 * bo.isP2Nop;
    jk.sib1_idx = i.sib1_idx; jk.sib2_idx = i.sib2_idx;
    jk.ins1 = st.isP1Insert; jk.ins2 = st.isP2Insert; jk.ov1_idx = ov1_idx; jk.ov2_idx = ov2_idx;
    hz_phaseJK_hook(c, hook, nLevels, jk, o.newStateRoot, o.newExitRoot);
    if (mn) { o.newStateRoot = c.set(slot_nsr, o.newStateRoot); o.newExitRoot = c.set(slot_ner, o.newExitRoot); }
}

// ---------------------------------------------------------------------------
// src/fee-tx.circom:26-111
// ---------------------------------------------------------------------------
TM struct FeeHook;
template <> struct FeeHook<MCompute> { HzSmtJob* smt; uint32_t sib_idx; };   // smt null: evaluate the processor in place
template <> struct FeeHook<MLayout> { int unused; };
HZ_HD bool hz_fee_job(const FeeHook<MCompute>& hk, uint32_t cur, const VC& oldKey, const VC& oldValue, const VC& isOld0,
                      const VC& newKey, const VC& newValue, const VC* f) {
    if (!hk.smt) return false;
    hz_smt_job(hk.smt, cur, hk.sib_idx, oldKey, oldValue, isOld0, newKey, newValue, f);
    hk.smt->hsel = fr_zero(); hk.smt->ovidx = HZ_NOIDX;
    return true;
}
#ifndef HZ_DEVICE_ONLY
inline bool hz_fee_job(const FeeHook<MLayout>&, uint32_t, const VL&, const VL&, const VL&, const VL&, const VL&, const VL*) { return false; }
#endif

TM HZ_FN V<M> FeeTx(Comp<M>&& c, int nLevels, const V<M>& oldStateRoot_, const V<M>& feePlanToken_,
                     const V<M>& feeIdx_, const V<M>& accFee_, const V<M>& tokenID_, const V<M>& nonce_,
                     const V<M>& sign_, const V<M>& balance_, const V<M>& ay_, const V<M>& ethAddr_, const V<M>* siblings,
                     const FeeHook<M>& hook) {
    V<M> oldStateRoot = c.def("oldStateRoot", oldStateRoot_);
    V<M> feePlanToken = c.def("feePlanToken", feePlanToken_);
    V<M> feeIdx = c.def("feeIdx", feeIdx_);
    V<M> accFee = c.def("accFee", accFee_);
    V<M> tokenID = c.def("tokenID", tokenID_);
    V<M> nonce = c.def("nonce", nonce_);
    V<M> sign = c.def("sign", sign_);
    V<M> balance = c.def("balance", balance_);
    V<M> ay = c.def("ay", ay_);
    V<M> ethAddr = c.def("ethAddr", ethAddr_);
    V<M> sib[HZ_MAX_LEVELS];
    for (int k = 0; k < nLevels + 1; k++) sib[k] = c.def(Nm("siblings", k), siblings[k]);
    // newStateRoot is wired to processor.newRoot
    V<M> z = IsZero(c.sub("feeIdxIsZero"), feeIdx);
    ForceEqualIfEnabled(c.sub("tokenIDChecker"), 1 - z, feePlanToken, tokenID);
    V<M> f[2];
    f[0] = c.def("p_fnc0", Ops<M>::K(0));
    f[1] = c.def("p_fnc1", 1 - z);
    V<M> oldH = HashState(c.sub("oldStFeePck"), tokenID, nonce, sign, balance, ay, ethAddr);
    V<M> newH = HashState(c.sub("newStFeePck"), tokenID, nonce, sign, accFee + balance, ay, ethAddr);
    // the processor is the last thing in the component: as a job nothing is left to do here
    if (hz_fee_job(hook, c.cur, feeIdx, oldH, Ops<M>::K(0), feeIdx, newH, f)) return Ops<M>::K(0);
    V<M> nr = SMTProcessor(c.sub("processor"), nLevels + 1, oldStateRoot, sib, feeIdx, oldH, Ops<M>::K(0), feeIdx, newH, f);
    return c.alias("newStateRoot", nr);
}
