"""hermeznetwork/circuits templates restated for the oracle (test infrastructure).

Each function follows one template under /root/reference/src (file:line cited),
declaring and assigning signals in source order on top of `oracle.circomlib`.
"""
from .circuit import Ctx, Comp, Sig, val, witness, symbols, number
from .field import P
from .fee_table import FEE_SHIFT_TABLE
from . import circomlib as cl


# ---------------------------------------------------------------------------
# src/lib/decode-float.circom:12-44
# ---------------------------------------------------------------------------
def DecodeFloatBin(c, bits):
    ins = c.inps("in", bits)
    out = c.sig("out")
    m = c.sigs("m", 35)
    e = c.sigs("e", 5)
    pe = c.sigs("pe", 5)
    scale10 = c.sig("scale10")
    for i in range(35):
        c.set(m[i], ins[i])
    for i in range(5):
        c.set(e[i], ins[i + 35])
    c.set(pe[0], 9 * e[0].v + 1)
    for i in range(1, 5):
        c.set(pe[i], (pe[i - 1].v * (10 ** (2 ** i)) - pe[i - 1].v) * e[i].v + pe[i - 1].v)
    c.set(scale10, pe[4])
    lcm = 0
    for i in range(35):
        lcm += m[i].v << i
    c.set(out, lcm * scale10.v)
    return out


# src/lib/decode-float.circom:51-65
def DecodeFloat(c, x):
    in_ = c.inp("in", x)
    out = c.sig("out")
    n2b_c = c.sub("n2b")
    dec_c = c.sub("decoder")
    bits = cl.Num2Bits(n2b_c, 40, in_)
    o = DecodeFloatBin(dec_c, bits)
    c.set(out, o)
    return out


# ---------------------------------------------------------------------------
# src/lib/mux256.circom:10-51
# ---------------------------------------------------------------------------
def Mux256(c, s, ins):
    s_ = c.inps("s", s)
    in_ = c.inps("in", ins)
    out = c.sig("out")
    mux_c = [c.sub("mux[%d]" % i) for i in range(17)]
    outs = []
    for i in range(16):
        outs.append(cl.Mux4(mux_c[i], in_[16 * i:16 * i + 16], s_[0:4]))
    o = cl.Mux4(mux_c[16], outs, s_[4:8])
    c.set(out, o)
    return out


# ---------------------------------------------------------------------------
# src/compute-fee.circom:12-94
# ---------------------------------------------------------------------------
def ComputeFee(c, feeSel, amount, applyFee):
    feeSel_ = c.inp("feeSel", feeSel)
    amount_ = c.inp("amount", amount)
    applyFee_ = c.inp("applyFee", applyFee)
    feeOut = c.sig("feeOut")
    mux_c = c.sub("mux256")
    selbits = cl.Num2Bits(c.sub("n2bFeeSel"), 8, feeSel_)
    muxo = Mux256(mux_c, [selbits[i].v * applyFee_.v for i in range(8)], FEE_SHIFT_TABLE)
    feeOutNotShifted = c.inp("feeOutNotShifted", muxo.v * amount_.v)
    applyShift = c.inp("applyShift", 1 - selbits[6].v * selbits[7].v)
    bitsFeeOut = c.sigs("bitsFeeOut", 253)
    lcIn = lcNot = lcSh = ovNot = ovSh = 0
    v = feeOutNotShifted.v
    for i in range(253):
        b = (v >> i) & 1
        c.set(bitsFeeOut[i], b)
        lcIn += b << i
        if i >= 60:
            if i < 60 + 128:
                lcSh += b << (i - 60)
            else:
                ovSh += b
        if i < 128:
            lcNot += b << i
        else:
            ovNot += b
    c.check(lcIn, v, "ComputeFee bits")
    c.check(applyShift.v * ovSh, 0, "ComputeFee overflow shifted")
    c.check((1 - applyShift.v) * ovNot, 0, "ComputeFee overflow not shifted")
    c.set(feeOut, applyShift.v * (lcSh - lcNot) + lcNot)
    return feeOut


# ---------------------------------------------------------------------------
# src/balance-updater.circom:24-112
# ---------------------------------------------------------------------------
def BalanceUpdater(c, oldStBalanceSender, oldStBalanceReceiver, amount, loadAmount, feeSelector,
                   onChain, nop, nullifyLoadAmount, nullifyAmount):
    bs = c.inp("oldStBalanceSender", oldStBalanceSender)
    br = c.inp("oldStBalanceReceiver", oldStBalanceReceiver)
    amount_ = c.inp("amount", amount)
    loadAmount_ = c.inp("loadAmount", loadAmount)
    feeSel = c.inp("feeSelector", feeSelector)
    onChain_ = c.inp("onChain", onChain)
    nop_ = c.inp("nop", nop)
    nla = c.inp("nullifyLoadAmount", nullifyLoadAmount)
    na = c.inp("nullifyAmount", nullifyAmount)
    newS = c.sig("newStBalanceSender")
    newR = c.sig("newStBalanceReceiver")
    isP2Nop = c.sig("isP2Nop")
    fee2Charge = c.sig("fee2Charge")
    isAmountNullified = c.sig("isAmountNullified")
    underflowOk = c.sig("underflowOk")
    ea1 = c.sig("effectiveAmount1")
    ea2 = c.sig("effectiveAmount2")
    ea3 = c.sig("effectiveAmount3")
    ela1 = c.sig("effectiveLoadAmount1")
    ela2 = c.sig("effectiveLoadAmount2")
    fee = ComputeFee(c.sub("computeFee"), feeSel, amount_, (1 - onChain_.v) * (1 - nop_.v))
    c.set(fee2Charge, fee)
    c.set(ela1, loadAmount_.v * onChain_.v)
    c.set(ela2, ela1.v * (1 - nla.v))
    c.set(ea1, amount_.v * (1 - nop_.v))
    c.set(ea2, ea1.v * (1 - na.v))
    bits = cl.Num2Bits(c.sub("n2bSender"), 193,
                       (1 << 192) + bs.v + ela2.v - ea2.v - fee2Charge.v)
    c.set(underflowOk, bits[192])
    c.check((1 - underflowOk.v) * (1 - onChain_.v), 0, "BalanceUpdater underflow")
    c.set(ea3, underflowOk.v * ea2.v)
    c.set(newS, bs.v + ela2.v - ea3.v - fee2Charge.v)
    c.set(newR, br.v + ea3.v)
    isz = cl.IsZero(c.sub("effectiveAmountIsZero"), ea1)
    c.set(isAmountNullified, 1 - (1 - na.v) * underflowOk.v)
    c.set(isP2Nop, 1 - isz.v)
    return newS, newR, isP2Nop, fee2Charge, isAmountNullified


# ---------------------------------------------------------------------------
# src/fee-accumulator.circom:17-45, 56-91
# ---------------------------------------------------------------------------
def FeeAccumulatorStep(c, tokenID, feePlanTokenID, isSelectedIn, fee2Charge, accFeeIn):
    tokenID_ = c.inp("tokenID", tokenID)
    plan = c.inp("feePlanTokenID", feePlanTokenID)
    selIn = c.inp("isSelectedIn", isSelectedIn)
    fee = c.inp("fee2Charge", fee2Charge)
    accIn = c.inp("accFeeIn", accFeeIn)
    selOut = c.sig("isSelectedOut")
    accOut = c.sig("accFeeOut")
    eq = cl.IsEqual(c.sub("isEqual"), tokenID_, plan)
    c.set(selOut, 1 - (1 - eq.v) * (1 - selIn.v))
    o = cl.Mux1(c.sub("mux"), accIn, accIn.v + fee.v, eq.v * (1 - selIn.v))
    c.set(accOut, o)
    return selOut, accOut


def FeeAccumulator(c, tokenID, fee2Charge, feePlanTokenID, accFeeIn):
    n = len(feePlanTokenID)
    tokenID_ = c.inp("tokenID", tokenID)
    fee = c.inp("fee2Charge", fee2Charge)
    plan = c.inps("feePlanTokenID", feePlanTokenID)
    accIn = c.inps("accFeeIn", accFeeIn)
    accOut = c.sigs("accFeeOut", n)
    sel = 0
    for i in range(n):
        sel, o = FeeAccumulatorStep(c.sub("chain[%d]" % i), tokenID_, plan[i], sel, fee, accIn[i])
        c.set(accOut[i], o)
    return accOut


# ---------------------------------------------------------------------------
# src/lib/hash-state.circom:18-40
# ---------------------------------------------------------------------------
def HashState(c, tokenID, nonce, sign, balance, ay, ethAddr):
    tokenID_ = c.inp("tokenID", tokenID)
    nonce_ = c.inp("nonce", nonce)
    sign_ = c.inp("sign", sign)
    balance_ = c.inp("balance", balance)
    ay_ = c.inp("ay", ay)
    ethAddr_ = c.inp("ethAddr", ethAddr)
    out = c.sig("out")
    e0 = c.inp("e0", tokenID_.v + nonce_.v * (1 << 32) + sign_.v * (1 << 72))
    o = cl.Poseidon(c.sub("hash"), [e0, balance_, ay_, ethAddr_])
    c.set(out, o)
    return out


# ---------------------------------------------------------------------------
# src/lib/utils-bjj.circom:12-28, 37-57
# ---------------------------------------------------------------------------
def BitsCompressed2AySign(c, bits):
    ins = c.inps("bjjCompressed", bits)
    ay = c.sig("ay")
    sign = c.sig("sign")
    o = cl.Bits2Num(c.sub("b2nAy"), ins[:254])
    c.set(ay, o)
    c.set(sign, ins[255])
    return ay, sign


def AySign2Ax(c, ay, sign):
    ay_ = c.inp("ay", ay)
    sign_ = c.inp("sign", sign)
    ax = c.sig("ax")
    bits = cl.Num2Bits(c.sub("n2bAy"), 254, ay_)
    o = cl.Bits2Point_Strict(c.sub("b2Point"), bits + [0, sign_])
    c.set(ax, o[0])
    return ax


# ---------------------------------------------------------------------------
# src/rq-tx-verifier.circom:19-94
# ---------------------------------------------------------------------------
def RqTxVerifier(c, futureV2, pastV2, futureEth, pastEth, futureAy, pastAy,
                 rqV2, rqEth, rqAy, rqTxOffset):
    fV2 = c.inps("futureTxCompressedDataV2", futureV2)
    pV2 = c.inps("pastTxCompressedDataV2", pastV2)
    fEth = c.inps("futureToEthAddr", futureEth)
    pEth = c.inps("pastToEthAddr", pastEth)
    fAy = c.inps("futureToBjjAy", futureAy)
    pAy = c.inps("pastToBjjAy", pastAy)
    rqV2_ = c.inp("rqTxCompressedDataV2", rqV2)
    rqEth_ = c.inp("rqToEthAddr", rqEth)
    rqAy_ = c.inp("rqToBjjAy", rqAy)
    off = c.inp("rqTxOffset", rqTxOffset)
    m1 = c.sub("muxTxCompressedDataV2")
    m2 = c.sub("muxToEthAddr")
    m3 = c.sub("muxToBjjAy")
    bits = cl.Num2Bits(c.sub("n2b"), 3, off)

    def sel(f, p):
        return [0, f[0], f[1], f[2], p[3], p[2], p[1], p[0]]
    o1 = cl.Mux3(m1, sel(fV2, pV2), bits)
    o2 = cl.Mux3(m2, sel(fEth, pEth), bits)
    o3 = cl.Mux3(m3, sel(fAy, pAy), bits)
    c.check(o1.v, rqV2_.v, "RqTxVerifier txCompressedDataV2")
    c.check(o2.v, rqEth_.v, "RqTxVerifier toEthAddr")
    c.check(o3.v, rqAy_.v, "RqTxVerifier toBjjAy")


# ---------------------------------------------------------------------------
# src/rollup-tx-states.circom:39-314
# ---------------------------------------------------------------------------
def RollupTxStates(c, fromIdx, toIdx, toEthAddr, auxFromIdx, auxToIdx, amount, newExit, loadAmount,
                   newAccount, onChain, fromEthAddr, ethAddr1, tokenID, tokenID1, tokenID2):
    fromIdx = c.inp("fromIdx", fromIdx)
    toIdx = c.inp("toIdx", toIdx)
    toEthAddr = c.inp("toEthAddr", toEthAddr)
    auxFromIdx = c.inp("auxFromIdx", auxFromIdx)
    auxToIdx = c.inp("auxToIdx", auxToIdx)
    amount = c.inp("amount", amount)
    newExit = c.inp("newExit", newExit)
    loadAmount = c.inp("loadAmount", loadAmount)
    newAccount = c.inp("newAccount", newAccount)
    onChain = c.inp("onChain", onChain)
    fromEthAddr = c.inp("fromEthAddr", fromEthAddr)
    ethAddr1 = c.inp("ethAddr1", ethAddr1)
    tokenID = c.inp("tokenID", tokenID)
    tokenID1 = c.inp("tokenID1", tokenID1)
    tokenID2 = c.inp("tokenID2", tokenID2)
    o = {}
    for nm in ["isP1Insert", "isP2Insert", "key1", "key2", "P1_fnc0", "P1_fnc1", "P2_fnc0", "P2_fnc1",
               "isExit", "verifySignEnabled", "nop", "checkToEthAddr", "checkToBjj",
               "nullifyLoadAmount", "nullifyAmount"]:
        o[nm] = c.sig(nm)
    oc, na = onChain.v, newAccount.v

    finalFromIdx = c.sig("finalFromIdx")
    sf = cl.Mux1(c.sub("selectFromIdx"), fromIdx, auxFromIdx, oc * na)
    c.set(finalFromIdx, sf)

    finalToIdx = c.sig("finalToIdx")
    toIdxIsZero = cl.IsZero(c.sub("toIdxIsZero"), toIdx)
    selectAuxToIdx = c.inp("selectAuxToIdx", (1 - oc) * toIdxIsZero.v)
    st = cl.Mux1(c.sub("selectToIdx"), toIdx, auxToIdx, selectAuxToIdx)
    c.set(finalToIdx, st)

    isToEthAddrAny = cl.IsEqual(c.sub("isToEthAddrAny"), (1 << 160) - 1, toEthAddr)
    checkIsExit = cl.IsEqual(c.sub("checkIsExit"), 1, finalToIdx)
    c.set(o["isExit"], checkIsExit)

    finalFromIdxIsZero = cl.IsZero(c.sub("finalFromIdxIsZero"), finalFromIdx)
    isFinalFromIdx = c.inp("isFinalFromIdx", 1 - finalFromIdxIsZero.v)
    loadAmountIsZero = cl.IsZero(c.sub("loadAmountIsZero"), loadAmount)
    isLoadAmount = c.inp("isLoadAmount", 1 - loadAmountIsZero.v)
    amountIsZero = cl.IsZero(c.sub("amountIsZero"), amount)
    isAmount = c.inp("isAmount", 1 - amountIsZero.v)

    c.check((1 - oc) * isLoadAmount.v, 0, "RollupTxStates loadAmount on L2")
    c.check((1 - oc) * na, 0, "RollupTxStates newAccount on L2")

    c.set(o["isP1Insert"], oc * na)
    c.set(o["P1_fnc0"], o["isP1Insert"].v * isFinalFromIdx.v)
    c.set(o["P1_fnc1"], (1 - o["isP1Insert"].v) * isFinalFromIdx.v)
    k1 = cl.Mux2(c.sub("mux1"), [0, finalFromIdx, finalFromIdx, finalFromIdx],
                 [o["P1_fnc0"], o["P1_fnc1"]])
    c.set(o["key1"], k1)

    c.set(o["isP2Insert"], o["isExit"].v * newExit.v)
    c.set(o["P2_fnc0"], o["isP2Insert"].v * isFinalFromIdx.v)
    c.set(o["P2_fnc1"], (1 - o["isP2Insert"].v) * isFinalFromIdx.v)
    k2 = cl.Mux2(c.sub("mux2"), [0, finalToIdx, 0, finalFromIdx], [isAmount, o["isExit"]])
    c.set(o["key2"], k2)

    c.set(o["verifySignEnabled"], (1 - oc) * isFinalFromIdx.v)
    c.set(o["nop"], finalFromIdxIsZero)

    tmpCheckToEthAddr = c.sig("tmpCheckToEthAddr")
    tmpCheckToBjj = c.sig("tmpCheckToBjj")
    c.set(tmpCheckToEthAddr, (1 - isToEthAddrAny.v) * selectAuxToIdx.v)
    c.set(tmpCheckToBjj, isToEthAddrAny.v * selectAuxToIdx.v)
    c.set(o["checkToEthAddr"], tmpCheckToEthAddr.v * (1 - o["nop"].v))
    c.set(o["checkToBjj"], tmpCheckToBjj.v * (1 - o["nop"].v))

    onChainNotCreateAccount = c.inp("onChainNotCreateAccount", (1 - na) * oc)
    applyNullifierEthAddr = c.sig("applyNullifierEthAddr")
    shouldCheckEthAddr = c.inp("shouldCheckEthAddr", onChainNotCreateAccount.v * isAmount.v)
    checkFromEthAddr = cl.IsEqual(c.sub("checkFromEthAddr"), fromEthAddr, ethAddr1)
    c.set(applyNullifierEthAddr, shouldCheckEthAddr.v * (1 - checkFromEthAddr.v))

    applyNullifierTokenID1 = c.sig("applyNullifierTokenID1")
    shouldCheckTokenID1 = c.inp("shouldCheckTokenID1", onChainNotCreateAccount)
    checkTokenID1 = cl.IsEqual(c.sub("checkTokenID1"), tokenID, tokenID1)
    c.set(applyNullifierTokenID1, shouldCheckTokenID1.v * (1 - checkTokenID1.v))

    applyNullifierTokenID2 = c.sig("applyNullifierTokenID2")
    shouldCheckTokenID2_0 = c.inp("shouldCheckTokenID2_0", oc * isAmount.v)
    shouldCheckTokenID2_1 = c.inp("shouldCheckTokenID2_1",
                                  shouldCheckTokenID2_0.v * (1 - o["isP2Insert"].v))
    checkTokenID2 = cl.IsEqual(c.sub("checkTokenID2"), tokenID, tokenID2)
    c.set(applyNullifierTokenID2, shouldCheckTokenID2_1.v * (1 - checkTokenID2.v))

    c.set(o["nullifyLoadAmount"], applyNullifierTokenID1.v * isLoadAmount.v)
    nullifyAmount_0 = c.sig("nullifyAmount_0")
    applyCheckTokenID1ToAmount = c.sig("applyCheckTokenID1ToAmount")
    c.set(applyCheckTokenID1ToAmount, applyNullifierTokenID1.v * isAmount.v)
    c.set(nullifyAmount_0, 1 - (1 - applyNullifierEthAddr.v) * (1 - applyNullifierTokenID2.v))
    c.set(o["nullifyAmount"], 1 - (1 - nullifyAmount_0.v) * (1 - applyCheckTokenID1ToAmount.v))
    return o


# ---------------------------------------------------------------------------
# src/decode-tx.circom:44-369
# ---------------------------------------------------------------------------
CONST_SIG = 3322668559


def DecodeTx(c, nLevels, previousOnChain, txCompressedData, maxNumBatch, amountF, toEthAddr, toBjjAy,
             rqTxCompressedDataV2, rqToEthAddr, rqToBjjAy, fromEthAddr, fromBjjCompressed, loadAmountF,
             globalChainID, currentNumBatch, onChain, newAccount, auxFromIdx, auxToIdx, inIdx):
    previousOnChain = c.inp("previousOnChain", previousOnChain)
    txCompressedData = c.inp("txCompressedData", txCompressedData)
    maxNumBatch = c.inp("maxNumBatch", maxNumBatch)
    amountF = c.inp("amountF", amountF)
    toEthAddr = c.inp("toEthAddr", toEthAddr)
    toBjjAy = c.inp("toBjjAy", toBjjAy)
    rqV2 = c.inp("rqTxCompressedDataV2", rqTxCompressedDataV2)
    rqEth = c.inp("rqToEthAddr", rqToEthAddr)
    rqAy = c.inp("rqToBjjAy", rqToBjjAy)
    L1L2TxData = c.sigs("L1L2TxData", nLevels * 2 + 40 + 8)
    txCompressedDataV2 = c.sig("txCompressedDataV2")
    fromEthAddr = c.inp("fromEthAddr", fromEthAddr)
    bjj = c.inps("fromBjjCompressed", fromBjjCompressed)
    loadAmountF = c.inp("loadAmountF", loadAmountF)
    globalChainID = c.inp("globalChainID", globalChainID)
    currentNumBatch = c.inp("currentNumBatch", currentNumBatch)
    onChain = c.inp("onChain", onChain)
    newAccount = c.inp("newAccount", newAccount)
    auxFromIdx = c.inp("auxFromIdx", auxFromIdx)
    auxToIdx = c.inp("auxToIdx", auxToIdx)
    L1TxFullData = c.sigs("L1TxFullData", 160 + 256 + 48 + 40 + 40 + 32 + 48)
    inIdx = c.inp("inIdx", inIdx)
    outIdx = c.sig("outIdx")
    constSig = c.sig("constSig")
    chainID = c.sig("chainID")
    fromIdx = c.sig("fromIdx")
    toIdx = c.sig("toIdx")
    tokenID = c.sig("tokenID")
    nonce = c.sig("nonce")
    userFee = c.sig("userFee")
    toBjjSign = c.sig("toBjjSign")
    amount = c.sig("amount")
    sigL2Hash = c.sig("sigL2Hash")
    oc = onChain.v

    d = cl.Num2Bits(c.sub("n2bData"), 225, txCompressedData)
    c.set(constSig, cl.Bits2Num(c.sub("b2nConstSig"), d[0:32]))
    c.set(chainID, cl.Bits2Num(c.sub("b2nChainID"), d[32:48]))
    c.set(fromIdx, cl.Bits2Num(c.sub("b2nFrom"), d[48:96]))
    c.check(sum(d[48 + i].v for i in range(nLevels, 48)), 0, "DecodeTx paddingFrom")
    c.set(toIdx, cl.Bits2Num(c.sub("b2nTo"), d[96:144]))
    c.check(sum(d[96 + i].v for i in range(nLevels, 48)), 0, "DecodeTx paddingTo")
    c.set(tokenID, cl.Bits2Num(c.sub("b2ntokenID"), d[144:176]))
    c.set(nonce, cl.Bits2Num(c.sub("b2nNonce"), d[176:216]))
    c.set(userFee, cl.Bits2Num(c.sub("b2nUserFee"), d[216:224]))
    c.set(toBjjSign, d[224])

    am = cl.Num2Bits(c.sub("n2bAmount"), 40, amountF)
    c.set(amount, DecodeFloatBin(c.sub("dfAmount"), am))

    v2bits = ([d[48 + i].v * (1 - oc) for i in range(48)]
              + [d[96 + i].v * (1 - oc) for i in range(48)]
              + [am[i].v * (1 - oc) for i in range(40)]
              + [d[144 + i].v * (1 - oc) for i in range(32)]
              + [d[176 + i].v * (1 - oc) for i in range(40)]
              + [d[216 + i].v * (1 - oc) for i in range(8)]
              + [d[224]])
    c.set(txCompressedDataV2, cl.Bits2Num(c.sub("b2nTxCompressedDataV2"), v2bits))

    toIdxIsZero = cl.IsZero(c.sub("toIdxIsZero"), toIdx)
    selectToIdx = cl.Mux1(c.sub("selectToIdx"), toIdx, auxToIdx, (1 - oc) * toIdxIsZero.v)
    ft = cl.Num2Bits(c.sub("n2bFinalToIdx"), nLevels, selectToIdx)
    for i in range(nLevels):
        c.set(L1L2TxData[nLevels - 1 - i], d[48 + i])
    for i in range(nLevels):
        c.set(L1L2TxData[nLevels * 2 - 1 - i], ft[i])
    for i in range(40):
        c.set(L1L2TxData[nLevels * 2 + 40 - 1 - i], am[i])
    for i in range(8):
        c.set(L1L2TxData[nLevels * 2 + 40 + 8 - 1 - i], d[216 + i].v * (1 - oc))

    b2nElement1_c = c.sub("b2nElement1")
    eth = cl.Num2Bits(c.sub("n2bToEthAddr"), 160, toEthAddr)
    mnb = cl.Num2Bits(c.sub("n2bMaxNumBatch"), 32, maxNumBatch)
    e1 = cl.Bits2Num(b2nElement1_c, eth + am + mnb)
    h = cl.Poseidon(c.sub("hashSig"), [txCompressedData, e1, toBjjAy, rqV2, rqEth, rqAy])
    c.set(sigL2Hash, h)

    feth = cl.Num2Bits(c.sub("n2bFromEthAddr"), 160, fromEthAddr)
    for i in range(160):
        c.set(L1TxFullData[160 - 1 - i], feth[i].v * oc)
    for i in range(256):
        c.set(L1TxFullData[160 + 256 - 1 - i], bjj[i].v * oc)
    for i in range(48):
        c.set(L1TxFullData[160 + 256 + 48 - 1 - i], d[48 + i].v * oc)
    la = cl.Num2Bits(c.sub("n2bLoadAmountF"), 40, loadAmountF)
    for i in range(40):
        c.set(L1TxFullData[160 + 256 + 48 + 40 - 1 - i], la[i].v * oc)
    for i in range(40):
        c.set(L1TxFullData[160 + 256 + 48 + 40 + 40 - 1 - i], am[i].v * oc)
    for i in range(32):
        c.set(L1TxFullData[160 + 256 + 48 + 40 + 40 + 32 - 1 - i], d[144 + i].v * oc)
    for i in range(48):
        c.set(L1TxFullData[160 + 256 + 48 + 40 + 40 + 32 + 48 - 1 - i], d[96 + i].v * oc)

    fromIdxIsZero = cl.IsZero(c.sub("fromIdxIsZero"), fromIdx)
    c.check(oc * fromIdxIsZero.v, newAccount.v, "DecodeTx newAccount")
    c.set(outIdx, inIdx.v + oc * newAccount.v)
    cl.ForceEqualIfEnabled(c.sub("idxChecker"), oc * newAccount.v, auxFromIdx, outIdx)
    c.check((1 - previousOnChain.v) * oc, 0, "DecodeTx L1 after L2")
    cl.ForceEqualIfEnabled(c.sub("chainIDChecker"), 1 - oc, globalChainID, chainID)
    cl.ForceEqualIfEnabled(c.sub("constSigChecker"), 1 - oc, constSig, CONST_SIG)
    mnbz = cl.IsZero(c.sub("maxNumBatchIsZero"), maxNumBatch)
    ok = cl.GreaterEqThan(c.sub("isMaxNumBatchOk"), 32, maxNumBatch, currentNumBatch)
    c.check((1 - ok.v) * (1 - mnbz.v), 0, "DecodeTx maxNumBatch")
    return dict(L1L2TxData=L1L2TxData, txCompressedDataV2=txCompressedDataV2,
                L1TxFullData=L1TxFullData, outIdx=outIdx, fromIdx=fromIdx, toIdx=toIdx,
                tokenID=tokenID, nonce=nonce, userFee=userFee, toBjjSign=toBjjSign,
                amount=amount, sigL2Hash=sigL2Hash, onChain=onChain)


# ---------------------------------------------------------------------------
# src/rollup-tx.circom:78-591
# ---------------------------------------------------------------------------
def RollupTx(c, nLevels, maxFeeTx, i):
    """`i` is a dict of input name -> Sig / int / list, keyed like the template's inputs."""
    feePlanTokens = c.inps("feePlanTokens", i["feePlanTokens"])
    accFeeIn = c.inps("accFeeIn", i["accFeeIn"])
    accFeeOut = c.sigs("accFeeOut", maxFeeTx)
    futureV2 = c.inps("futureTxCompressedDataV2", i["futureTxCompressedDataV2"])
    pastV2 = c.inps("pastTxCompressedDataV2", i["pastTxCompressedDataV2"])
    futureEth = c.inps("futureToEthAddr", i["futureToEthAddr"])
    pastEth = c.inps("pastToEthAddr", i["pastToEthAddr"])
    futureAy = c.inps("futureToBjjAy", i["futureToBjjAy"])
    pastAy = c.inps("pastToBjjAy", i["pastToBjjAy"])
    s = {}
    for nm in ["fromIdx", "auxFromIdx", "toIdx", "auxToIdx", "toBjjAy", "toBjjSign", "toEthAddr",
               "amount", "tokenID", "nonce", "userFee", "rqOffset", "onChain", "newAccount",
               "rqTxCompressedDataV2", "rqToEthAddr", "rqToBjjAy", "sigL2Hash", "s", "r8x", "r8y",
               "fromEthAddr"]:
        s[nm] = c.inp(nm, i[nm])
    bjj = c.inps("fromBjjCompressed", i["fromBjjCompressed"])
    s["loadAmountF"] = c.inp("loadAmountF", i["loadAmountF"])
    for nm in ["tokenID1", "nonce1", "sign1", "balance1", "ay1", "ethAddr1"]:
        s[nm] = c.inp(nm, i[nm])
    siblings1 = c.inps("siblings1", i["siblings1"])
    for nm in ["isOld0_1", "oldKey1", "oldValue1",
               "tokenID2", "nonce2", "sign2", "balance2", "newExit", "ay2", "ethAddr2"]:
        s[nm] = c.inp(nm, i[nm])
    siblings2 = c.inps("siblings2", i["siblings2"])
    for nm in ["isOld0_2", "oldKey2", "oldValue2"]:
        s[nm] = c.inp(nm, i[nm])
    oldStateRoot = c.inp("oldStateRoot", i["oldStateRoot"])
    newStateRoot = c.sig("newStateRoot")
    oldExitRoot = c.inp("oldExitRoot", i["oldExitRoot"])
    newExitRoot = c.sig("newExitRoot")
    isAmountNullified = c.sig("isAmountNullified")
    oc = s["onChain"].v

    # A - compute tx-states (:183-209)
    loadAmount = c.sig("loadAmount")
    la_bits = cl.Num2Bits(c.sub("n2bloadAmountF"), 40, s["loadAmountF"])
    c.set(loadAmount, DecodeFloatBin(c.sub("dfLoadAmount"), la_bits))
    st = RollupTxStates(c.sub("states"), s["fromIdx"], s["toIdx"], s["toEthAddr"], s["auxFromIdx"],
                        s["auxToIdx"], s["amount"], s["newExit"], loadAmount, s["newAccount"],
                        s["onChain"], s["fromEthAddr"], s["ethAddr1"], s["tokenID"], s["tokenID1"],
                        s["tokenID2"])
    # B - check request transaction fields (:213-231)
    RqTxVerifier(c.sub("rqTxVerifier"), futureV2, pastV2, futureEth, pastEth, futureAy, pastAy,
                 s["rqTxCompressedDataV2"], s["rqToEthAddr"], s["rqToBjjAy"], s["rqOffset"])
    # C - check state fields (:237-292)
    F = cl.ForceEqualIfEnabled
    F(c.sub("nonceChecker"), 1 - oc, s["nonce"], s["nonce1"])
    F(c.sub("checkToEthAddr"), 1 - (1 - st["checkToEthAddr"].v) * (1 - st["checkToBjj"].v),
      s["toEthAddr"], s["ethAddr2"])
    F(c.sub("toBjjAyChecker"), st["checkToBjj"], s["ay2"], s["toBjjAy"])
    F(c.sub("toBjjSignChecker"), st["checkToBjj"], s["sign2"], s["toBjjSign"])
    F(c.sub("checkTokenID1"), 1 - oc, s["tokenID"], s["tokenID1"])
    F(c.sub("checkTokenID2"), (1 - oc) * (1 - st["isP2Insert"].v), s["tokenID"], s["tokenID2"])
    F(c.sub("checkTokenID1L1"), st["isP1Insert"], s["tokenID"], s["tokenID1"])
    F(c.sub("fromEthAddrChecker"), st["isP1Insert"], s["fromEthAddr"], s["ethAddr1"])
    # D - compute old hash states (:297-312)
    oldSt1Hash = HashState(c.sub("oldSt1Hash"), s["tokenID1"], s["nonce1"], s["sign1"], s["balance1"],
                           s["ay1"], s["ethAddr1"])
    oldSt2Hash = HashState(c.sub("oldSt2Hash"), s["tokenID2"], s["nonce2"], s["sign2"], s["balance2"],
                           s["ay2"], s["ethAddr2"])
    # E - signal processor selectors (:317-443)
    bay, bsign = BitsCompressed2AySign(c.sub("decodeFromBjj"), bjj)
    M = cl.Mux1
    i1, i2 = st["isP1Insert"], st["isP2Insert"]
    s1Balance = M(c.sub("s1Balance"), s["balance1"], 0, i1)
    s1Sign = M(c.sub("s1Sign"), s["sign1"], bsign, i1)
    s1Ay = M(c.sub("s1Ay"), s["ay1"], bay, i1)
    s1Nonce = M(c.sub("s1Nonce"), s["nonce1"], 0, i1)
    s1EthAddr = M(c.sub("s1EthAddr"), s["ethAddr1"], s["fromEthAddr"], i1)
    s1TokenID = M(c.sub("s1TokenID"), s["tokenID1"], s["tokenID"], i1)
    s1OldKey = M(c.sub("s1OldKey"), st["key1"], s["oldKey1"], i1)
    s1OldValue = M(c.sub("s1OldValue"), oldSt1Hash, s["oldValue1"], i1)
    s2Balance = M(c.sub("s2Balance"), s["balance2"], 0, i2)
    s2Sign = M(c.sub("s2Sign"), s["sign2"], s1Sign, i2)
    s2Ay = M(c.sub("s2Ay"), s["ay2"], s1Ay, i2)
    s2Nonce = M(c.sub("s2Nonce"), s["nonce2"], 0, i2)
    s2EthAddr = M(c.sub("s2EthAddr"), s["ethAddr2"], s1EthAddr, i2)
    s2TokenID = M(c.sub("s2TokenID"), s["tokenID2"], s1TokenID, i2)
    s2OldKey = M(c.sub("s2OldKey"), st["key2"], s["oldKey2"], i2)
    s2OldValue = M(c.sub("s2OldValue"), oldSt2Hash, s["oldValue2"], i2)
    # F - verify eddsa signature (:454-482)
    vse = st["verifySignEnabled"]
    signSignature = M(c.sub("signSignature"), 0, s1Sign, vse)
    aySignature = M(c.sub("aySignature"), 0, s1Ay, vse)
    ax = AySign2Ax(c.sub("getAx"), aySignature, signSignature)
    cl.EdDSAPoseidonVerifier(c.sub("sigVerifier"), vse, ax, s1Ay, s["s"], s["r8x"], s["r8y"],
                             s["sigL2Hash"])
    # G - update balances (:486-499)
    newS, newR, isP2Nop, fee2Charge, ian = BalanceUpdater(
        c.sub("balanceUpdater"), s1Balance, s2Balance, s["amount"], loadAmount, s["userFee"],
        s["onChain"], st["nop"], st["nullifyLoadAmount"], st["nullifyAmount"])
    c.set(isAmountNullified, ian)
    # H - accumulate fees (:501-512)
    fo = FeeAccumulator(c.sub("feeAccumulator"), s["tokenID"], fee2Charge, feePlanTokens, accFeeIn)
    for k in range(maxFeeTx):
        c.set(accFeeOut[k], fo[k])
    # I - compute hash new states (:517-532)
    newSt1Hash = HashState(c.sub("newSt1Hash"), s1TokenID, s1Nonce.v + (1 - oc), s1Sign, newS, s1Ay,
                           s1EthAddr)
    newSt2Hash = HashState(c.sub("newSt2Hash"), s2TokenID, s2Nonce, s2Sign, newR, s2Ay, s2EthAddr)
    # J - smt processors (:537-570)
    p1 = cl.SMTProcessor(c.sub("processor1"), nLevels + 1, oldStateRoot, siblings1, s1OldKey,
                         s1OldValue, s["isOld0_1"], st["key1"], newSt1Hash,
                         [st["P1_fnc0"], st["P1_fnc1"]])
    s3 = M(c.sub("s3"), p1, oldExitRoot, st["isExit"])
    p2 = cl.SMTProcessor(c.sub("processor2"), nLevels + 1, s3, siblings2, s2OldKey, s2OldValue,
                         s["isOld0_2"], st["key2"], newSt2Hash,
                         [st["P2_fnc0"].v * isP2Nop.v, st["P2_fnc1"].v * isP2Nop.v])
    # K - select output roots (:577-590)
    s4 = M(c.sub("s4"), p2, p1, st["isExit"])
    c.set(newStateRoot, s4)
    s5 = M(c.sub("s5"), oldExitRoot, p2, st["isExit"])
    c.set(newExitRoot, s5)
    return dict(accFeeOut=accFeeOut, newStateRoot=newStateRoot, newExitRoot=newExitRoot,
                isAmountNullified=isAmountNullified)


# ---------------------------------------------------------------------------
# src/fee-tx.circom:26-111
# ---------------------------------------------------------------------------
def FeeTx(c, nLevels, oldStateRoot, feePlanToken, feeIdx, accFee, tokenID, nonce, sign, balance, ay,
          ethAddr, siblings):
    oldStateRoot = c.inp("oldStateRoot", oldStateRoot)
    feePlanToken = c.inp("feePlanToken", feePlanToken)
    feeIdx = c.inp("feeIdx", feeIdx)
    accFee = c.inp("accFee", accFee)
    tokenID = c.inp("tokenID", tokenID)
    nonce = c.inp("nonce", nonce)
    sign = c.inp("sign", sign)
    balance = c.inp("balance", balance)
    ay = c.inp("ay", ay)
    ethAddr = c.inp("ethAddr", ethAddr)
    sib = c.inps("siblings", siblings)
    newStateRoot = c.sig("newStateRoot")
    z = cl.IsZero(c.sub("feeIdxIsZero"), feeIdx)
    cl.ForceEqualIfEnabled(c.sub("tokenIDChecker"), 1 - z.v, feePlanToken, tokenID)
    p_fnc0 = c.sig("p_fnc0")
    p_fnc1 = c.sig("p_fnc1")
    c.set(p_fnc0, 0)
    c.set(p_fnc1, 1 - z.v)
    oldH = HashState(c.sub("oldStFeePck"), tokenID, nonce, sign, balance, ay, ethAddr)
    newH = HashState(c.sub("newStFeePck"), tokenID, nonce, sign, accFee.v + balance.v, ay, ethAddr)
    nr = cl.SMTProcessor(c.sub("processor"), nLevels + 1, oldStateRoot, sib, feeIdx, oldH, 0, feeIdx,
                         newH, [p_fnc0, p_fnc1])
    c.set(newStateRoot, nr)
    return newStateRoot


# ---------------------------------------------------------------------------
# src/hash-inputs.circom:23-185
# ---------------------------------------------------------------------------
def HashInputs(c, nLevels, nTx, maxL1Tx, maxFeeTx, oldLastIdx, newLastIdx, oldStateRoot, newStateRoot,
               newExitRoot, L1TxsFullData, L1L2TxsData, feeTxsData, globalChainID, currentNumBatch):
    oldLastIdx = c.inp("oldLastIdx", oldLastIdx)
    newLastIdx = c.inp("newLastIdx", newLastIdx)
    oldStateRoot = c.inp("oldStateRoot", oldStateRoot)
    newStateRoot = c.inp("newStateRoot", newStateRoot)
    newExitRoot = c.inp("newExitRoot", newExitRoot)
    l1 = c.inps("L1TxsFullData", L1TxsFullData)
    l2 = c.inps("L1L2TxsData", L1L2TxsData)
    fee = c.inps("feeTxsData", feeTxsData)
    chain = c.inp("globalChainID", globalChainID)
    batch = c.inp("currentNumBatch", currentNumBatch)
    out = c.sig("hashInputsOut")
    b_old = cl.Num2Bits(c.sub("n2bOldLastIdx"), 48, oldLastIdx)
    c.check(sum(b_old[i].v for i in range(nLevels, 48)), 0, "HashInputs paddingOldLastIdx")
    b_new = cl.Num2Bits(c.sub("n2bNewLastIdx"), 48, newLastIdx)
    c.check(sum(b_new[i].v for i in range(nLevels, 48)), 0, "HashInputs paddingNewLastIdx")
    b_osr = cl.Num2Bits(c.sub("n2bOldStateRoot"), 256, oldStateRoot)
    b_nsr = cl.Num2Bits(c.sub("n2bNewStateRoot"), 256, newStateRoot)
    b_ner = cl.Num2Bits(c.sub("n2bNewExitRoot"), 256, newExitRoot)
    b_fee = []
    for i in range(maxFeeTx):
        b = cl.Num2Bits(c.sub("n2bFeeTxsData[%d]" % i), 48, fee[i])
        c.check(sum(b[j].v for j in range(nLevels, 48)), 0, "HashInputs paddingFeeTxsData")
        b_fee.append(b)
    b_chain = cl.Num2Bits(c.sub("n2bChainID"), 16, chain)
    b_batch = cl.Num2Bits(c.sub("n2bCurrentNumBatch"), 32, batch)
    pre = []
    pre += b_old[::-1]
    pre += b_new[::-1]
    pre += b_osr[::-1]
    pre += b_nsr[::-1]
    pre += b_ner[::-1]
    pre += l1
    pre += l2
    for i in range(maxFeeTx):
        pre += b_fee[i][:nLevels][::-1]
    pre += b_chain[::-1]
    pre += b_batch[::-1]
    h = cl.Sha256(c.sub("inputsHasher"), pre)
    o = cl.Bits2Num(c.sub("n2bHashInputsOut"), [h[255 - i] for i in range(256)])
    c.set(out, o)
    return out


# ---------------------------------------------------------------------------
# src/rollup-main.circom:82-475
# ---------------------------------------------------------------------------
SCALAR_INPUTS = ["oldLastIdx", "oldStateRoot", "globalChainID", "currentNumBatch"]


def rollup_main_input_shapes(nTx, nLevels, maxL1Tx, maxFeeTx):
    """Ordered (name, shape) list of main inputs (src/rollup-main.circom:105-196)."""
    L = nLevels + 1
    return [
        ("oldLastIdx", ()), ("oldStateRoot", ()), ("globalChainID", ()), ("currentNumBatch", ()),
        ("feeIdxs", (maxFeeTx,)), ("feePlanTokens", (maxFeeTx,)),
        ("imOnChain", (nTx - 1,)), ("imOutIdx", (nTx - 1,)), ("imStateRoot", (nTx - 1,)),
        ("imExitRoot", (nTx - 1,)), ("imAccFeeOut", (nTx - 1, maxFeeTx)),
        ("imStateRootFee", (maxFeeTx - 1,)), ("imInitStateRootFee", ()), ("imFinalAccFee", (maxFeeTx,)),
        ("txCompressedData", (nTx,)), ("amountF", (nTx,)), ("txCompressedDataV2", (nTx,)),
        ("fromIdx", (nTx,)), ("auxFromIdx", (nTx,)), ("toIdx", (nTx,)), ("auxToIdx", (nTx,)),
        ("toBjjAy", (nTx,)), ("toEthAddr", (nTx,)), ("maxNumBatch", (nTx,)), ("onChain", (nTx,)),
        ("newAccount", (nTx,)), ("rqOffset", (nTx,)), ("rqTxCompressedDataV2", (nTx,)),
        ("rqToEthAddr", (nTx,)), ("rqToBjjAy", (nTx,)), ("s", (nTx,)), ("r8x", (nTx,)), ("r8y", (nTx,)),
        ("loadAmountF", (nTx,)), ("fromEthAddr", (nTx,)), ("fromBjjCompressed", (nTx, 256)),
        ("tokenID1", (nTx,)), ("nonce1", (nTx,)), ("sign1", (nTx,)), ("balance1", (nTx,)),
        ("ay1", (nTx,)), ("ethAddr1", (nTx,)), ("siblings1", (nTx, L)), ("isOld0_1", (nTx,)),
        ("oldKey1", (nTx,)), ("oldValue1", (nTx,)),
        ("tokenID2", (nTx,)), ("nonce2", (nTx,)), ("sign2", (nTx,)), ("balance2", (nTx,)),
        ("ay2", (nTx,)), ("ethAddr2", (nTx,)), ("siblings2", (nTx, L)), ("newExit", (nTx,)),
        ("isOld0_2", (nTx,)), ("oldKey2", (nTx,)), ("oldValue2", (nTx,)),
        ("tokenID3", (maxFeeTx,)), ("nonce3", (maxFeeTx,)), ("sign3", (maxFeeTx,)),
        ("balance3", (maxFeeTx,)), ("ay3", (maxFeeTx,)), ("ethAddr3", (maxFeeTx,)),
        ("siblings3", (maxFeeTx, L)),
    ]


def _declare_inputs(c, shapes, inputs):
    sig = {}
    for name, shape in shapes:
        if name not in inputs:
            raise KeyError("missing input signal %s" % name)
        v = inputs[name]
        if shape == ():
            sig[name] = c.inp(name, int(v))
        elif len(shape) == 1:
            assert len(v) == shape[0], name
            sig[name] = [c.inp("%s[%d]" % (name, a), int(v[a])) for a in range(shape[0])]
        else:
            assert len(v) == shape[0], name
            sig[name] = [[c.inp("%s[%d][%d]" % (name, a, b), int(v[a][b])) for b in range(shape[1])]
                         for a in range(shape[0])]
    return sig


def RollupMain(c, nTx, nLevels, maxL1Tx, maxFeeTx, inputs):
    hashGlobalInputs = c.sig("hashGlobalInputs")
    I = _declare_inputs(c, rollup_main_input_shapes(nTx, nLevels, maxL1Tx, maxFeeTx), inputs)
    # A - check binary signals (:207-219)
    for i in range(nTx - 1):
        c.check_bool(I["imOnChain"][i], "imOnChain")
    for i in range(nTx):
        c.check_bool(I["onChain"][i], "onChain")
        c.check_bool(I["newAccount"][i], "newAccount")
        for j in range(256):
            c.check_bool(I["fromBjjCompressed"][i][j], "fromBjjCompressed")
        c.check_bool(I["isOld0_1"][i], "isOld0_1")
        c.check_bool(I["isOld0_2"][i], "isOld0_2")
    # B - decode transactions (:223-254)
    dec = []
    for i in range(nTx):
        dec.append(DecodeTx(
            c.sub("decodeTx[%d]" % i), nLevels,
            1 if i == 0 else I["imOnChain"][i - 1],
            I["txCompressedData"][i], I["maxNumBatch"][i], I["amountF"][i], I["toEthAddr"][i],
            I["toBjjAy"][i], I["rqTxCompressedDataV2"][i], I["rqToEthAddr"][i], I["rqToBjjAy"][i],
            I["fromEthAddr"][i], I["fromBjjCompressed"][i], I["loadAmountF"][i],
            I["globalChainID"], I["currentNumBatch"], I["onChain"][i], I["newAccount"][i],
            I["auxFromIdx"][i], I["auxToIdx"][i],
            I["oldLastIdx"] if i == 0 else I["imOutIdx"][i - 1]))
    # C - check integrity decode intermediary signals (:258-265)
    for i in range(nTx):
        c.check(dec[i]["txCompressedDataV2"].v, I["txCompressedDataV2"][i].v, "main txCompressedDataV2")
    for i in range(nTx - 1):
        c.check(dec[i]["onChain"].v, I["imOnChain"][i].v, "main imOnChain")
        c.check(dec[i]["outIdx"].v, I["imOutIdx"][i].v, "main imOutIdx")
    # D - process rollup transactions (:269-379)
    rtx = []
    for i in range(nTx):
        def win(arr, k):
            return arr[k] if 0 <= k < nTx else 0
        ti = dict(
            feePlanTokens=I["feePlanTokens"],
            accFeeIn=[0] * maxFeeTx if i == 0 else I["imAccFeeOut"][i - 1],
            futureTxCompressedDataV2=[win(I["txCompressedDataV2"], i + j + 1) for j in range(3)],
            futureToEthAddr=[win(I["toEthAddr"], i + j + 1) for j in range(3)],
            futureToBjjAy=[win(I["toBjjAy"], i + j + 1) for j in range(3)],
            pastTxCompressedDataV2=[win(I["txCompressedDataV2"], i - j - 1) for j in range(4)],
            pastToEthAddr=[win(I["toEthAddr"], i - j - 1) for j in range(4)],
            pastToBjjAy=[win(I["toBjjAy"], i - j - 1) for j in range(4)],
            fromIdx=dec[i]["fromIdx"], auxFromIdx=I["auxFromIdx"][i],
            toIdx=dec[i]["toIdx"], auxToIdx=I["auxToIdx"][i], toBjjAy=I["toBjjAy"][i],
            toBjjSign=dec[i]["toBjjSign"], toEthAddr=I["toEthAddr"][i],
            amount=dec[i]["amount"], tokenID=dec[i]["tokenID"], nonce=dec[i]["nonce"],
            userFee=dec[i]["userFee"], rqOffset=I["rqOffset"][i], onChain=I["onChain"][i],
            newAccount=I["newAccount"][i], rqTxCompressedDataV2=I["rqTxCompressedDataV2"][i],
            rqToEthAddr=I["rqToEthAddr"][i], rqToBjjAy=I["rqToBjjAy"][i],
            sigL2Hash=dec[i]["sigL2Hash"], s=I["s"][i], r8x=I["r8x"][i], r8y=I["r8y"][i],
            fromEthAddr=I["fromEthAddr"][i], loadAmountF=I["loadAmountF"][i],
            fromBjjCompressed=I["fromBjjCompressed"][i],
            tokenID1=I["tokenID1"][i], nonce1=I["nonce1"][i], sign1=I["sign1"][i],
            balance1=I["balance1"][i], ay1=I["ay1"][i], ethAddr1=I["ethAddr1"][i],
            siblings1=I["siblings1"][i], isOld0_1=I["isOld0_1"][i], oldKey1=I["oldKey1"][i],
            oldValue1=I["oldValue1"][i],
            tokenID2=I["tokenID2"][i], nonce2=I["nonce2"][i], sign2=I["sign2"][i],
            balance2=I["balance2"][i], newExit=I["newExit"][i], ay2=I["ay2"][i],
            ethAddr2=I["ethAddr2"][i], siblings2=I["siblings2"][i], isOld0_2=I["isOld0_2"][i],
            oldKey2=I["oldKey2"][i], oldValue2=I["oldValue2"][i],
            oldStateRoot=I["oldStateRoot"] if i == 0 else I["imStateRoot"][i - 1],
            oldExitRoot=0 if i == 0 else I["imExitRoot"][i - 1])
        rtx.append(RollupTx(c.sub("rollupTx[%d]" % i), nLevels, maxFeeTx, ti))
    # E - check integrity transactions intermediary signals (:383-389)
    for i in range(nTx - 1):
        c.check(rtx[i]["newStateRoot"].v, I["imStateRoot"][i].v, "main imStateRoot")
        c.check(rtx[i]["newExitRoot"].v, I["imExitRoot"][i].v, "main imExitRoot")
        for j in range(maxFeeTx):
            c.check(rtx[i]["accFeeOut"][j].v, I["imAccFeeOut"][i][j].v, "main imAccFeeOut")
    # F - process fee transactions (:393-417)
    ftx = []
    for i in range(maxFeeTx):
        ftx.append(FeeTx(
            c.sub("feeTx[%d]" % i), nLevels,
            I["imInitStateRootFee"] if i == 0 else I["imStateRootFee"][i - 1],
            I["feePlanTokens"][i], I["feeIdxs"][i], I["imFinalAccFee"][i],
            I["tokenID3"][i], I["nonce3"][i], I["sign3"][i], I["balance3"][i], I["ay3"][i],
            I["ethAddr3"][i], I["siblings3"][i]))
    # G - check integrity fee transactions intermediary signals (:422-431)
    for i in range(maxFeeTx - 1):
        c.check(ftx[i].v, I["imStateRootFee"][i].v, "main imStateRootFee")
    c.check(rtx[nTx - 1]["newStateRoot"].v, I["imInitStateRootFee"].v, "main imInitStateRootFee")
    for i in range(maxFeeTx):
        c.check(rtx[nTx - 1]["accFeeOut"][i].v, I["imFinalAccFee"][i].v, "main imFinalAccFee")
    # H - compute global hash input (:435-474)
    l1 = []
    for i in range(maxL1Tx):
        l1 += dec[i]["L1TxFullData"]
    l2 = []
    for i in range(nTx):
        d = dec[i]["L1L2TxData"]
        l2 += d[:2 * nLevels]
        l2 += [d[j].v * (1 - rtx[i]["isAmountNullified"].v) for j in range(2 * nLevels, 2 * nLevels + 40)]
        l2 += d[2 * nLevels + 40:]
    h = HashInputs(c.sub("hasherInputs"), nLevels, nTx, maxL1Tx, maxFeeTx,
                   I["oldLastIdx"], dec[nTx - 1]["outIdx"], I["oldStateRoot"], ftx[maxFeeTx - 1],
                   rtx[nTx - 1]["newExitRoot"], l1, l2, I["feeIdxs"], I["globalChainID"],
                   I["currentNumBatch"])
    c.set(hashGlobalInputs, h)
    return hashGlobalInputs


# ---------------------------------------------------------------------------
# src/withdraw.circom:21-72, 84-176
# ---------------------------------------------------------------------------
def withdraw_input_shapes(nLevels):
    return [("rootExit", ()), ("ethAddr", ()), ("tokenID", ()), ("balance", ()), ("idx", ()),
            ("sign", ()), ("ay", ()), ("siblingsState", (nLevels + 1,))]


def HashInputsWithdrawal(c, nLevels, rootExit, ethAddr, tokenID, balance, idx):
    rootExit = c.inp("rootExit", rootExit)
    ethAddr = c.inp("ethAddr", ethAddr)
    tokenID = c.inp("tokenID", tokenID)
    balance = c.inp("balance", balance)
    idx = c.inp("idx", idx)
    out = c.sig("hashInputsOut")
    b_root = cl.Num2Bits(c.sub("n2bRootExit"), 256, rootExit)
    b_eth = cl.Num2Bits(c.sub("n2bEthAddr"), 160, ethAddr)
    b_tok = cl.Num2Bits(c.sub("n2bTokenID"), 32, tokenID)
    b_bal = cl.Num2Bits(c.sub("n2bBalance"), 192, balance)
    b_idx = cl.Num2Bits(c.sub("n2bIdx"), 48, idx)
    c.check(sum(b_idx[j].v for j in range(nLevels, 48)), 0, "HashInputsWithdrawal paddingIdx")
    pre = b_root[::-1] + b_eth[::-1] + b_tok[::-1] + b_bal[::-1] + b_idx[::-1]
    h = cl.Sha256(c.sub("inputsHasher"), pre)
    o = cl.Bits2Num(c.sub("n2bHashInputsOut"), [h[255 - i] for i in range(256)])
    c.set(out, o)
    return out


def Withdraw(c, nLevels, inputs):
    hashGlobalInputs = c.sig("hashGlobalInputs")
    I = _declare_inputs(c, withdraw_input_shapes(nLevels), inputs)
    st = HashState(c.sub("accountState"), I["tokenID"], 0, I["sign"], I["balance"], I["ay"],
                   I["ethAddr"])
    cl.SMTVerifier(c.sub("smtVerify"), nLevels + 1, 1, I["rootExit"], I["siblingsState"], 0, 0, 0,
                   I["idx"], st, 0)
    h = HashInputsWithdrawal(c.sub("hasherInputs"), nLevels, I["rootExit"], I["ethAddr"],
                             I["tokenID"], I["balance"], I["idx"])
    c.set(hashGlobalInputs, h)
    return hashGlobalInputs


# ---------------------------------------------------------------------------
# drivers
# ---------------------------------------------------------------------------
def run_template(fn, *args, sanity=True, **kw):
    """Instantiate `fn` as the main component; returns (main Comp, ctx, result)."""
    ctx = Ctx(sanity=sanity)
    main = Comp(ctx, "main", is_main=True)
    res = fn(main, *args, **kw)
    return main, ctx, res


def rollup_main_witness(nTx, nLevels, maxL1Tx, maxFeeTx, inputs, sanity=True):
    main, ctx, _ = run_template(RollupMain, nTx, nLevels, maxL1Tx, maxFeeTx, inputs, sanity=sanity)
    return witness(main), main, ctx


ROLLUP_TX_INPUTS = (["feePlanTokens", "accFeeIn", "futureTxCompressedDataV2", "pastTxCompressedDataV2", "futureToEthAddr",
                     "pastToEthAddr", "futureToBjjAy", "pastToBjjAy", "fromIdx", "auxFromIdx", "toIdx", "auxToIdx", "toBjjAy",
                     "toBjjSign", "toEthAddr", "amount", "tokenID", "nonce", "userFee", "rqOffset", "onChain", "newAccount",
                     "rqTxCompressedDataV2", "rqToEthAddr", "rqToBjjAy", "sigL2Hash", "s", "r8x", "r8y", "fromEthAddr",
                     "fromBjjCompressed", "loadAmountF", "tokenID1", "nonce1", "sign1", "balance1", "ay1", "ethAddr1",
                     "siblings1", "isOld0_1", "oldKey1", "oldValue1", "tokenID2", "nonce2", "sign2", "balance2", "newExit",
                     "ay2", "ethAddr2", "siblings2", "isOld0_2", "oldKey2", "oldValue2", "oldStateRoot", "oldExitRoot"])


def rollup_tx_witness(nLevels, maxFeeTx, inputs, sanity=True):
    """RollupTx(nLevels, maxFeeTx) as the main component (`component main = RollupTx(...)`,
    /root/reference/test/rollup-tx.test.js:56-63): inputs keyed by the template's input names
    (/root/reference/src/rollup-tx.circom:93-170)."""
    main, ctx, _ = run_template(RollupTx, nLevels, maxFeeTx, inputs, sanity=sanity)
    return witness(main), main, ctx


def withdraw_witness(nLevels, inputs, sanity=True):
    main, ctx, _ = run_template(Withdraw, nLevels, inputs, sanity=sanity)
    return witness(main), main, ctx
