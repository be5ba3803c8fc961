// Withdraw(nLevels) (/root/reference/src/withdraw.circom:21-72) and HashInputsWithdrawal (:84-176):
// one exit-tree inclusion proof per instance -- HashState of the account (nonce 0), SMTVerifier(nLevels + 1)
// in inclusion mode, SHA-256 of rootExit | ethAddr | tokenID | balance | idx (688 bits, 2 blocks).
// Instances are independent: the device gives each its own lane group and its own witness range
// (hzw_calculate_batch; config 5 of BASELINE.json batches 65 536 of them over 8 GPUs).
#pragma once
#include "rollup_main.cuh"

typedef BVT<MCompute, 32> BVC32;
template <class M> using BV32 = BVT<M, 32>;

// message word t of block b from the padded bit vector (bit k of the word = paddedIn[512 b + 32 t + 31 - k])
HZ_HD W32<MCompute> w32_from_bits(const BVC32& bits, int pos) {
    W32<MCompute> r; r.v = 0;
    for (int k = 0; k < 32; k++) r.v |= bits.bit(pos + 31 - k) << k;
    return r;
}
#ifndef HZ_DEVICE_ONLY
inline W32<MLayout> w32_from_bits(const BVT<MLayout, 32>& bits, int pos) {
    W32<MLayout> r;
    for (int k = 0; k < 32; k++) r.idx[k] = bits.idx[pos + 31 - k];
    return r;
}
#endif

// circomlib sha256.circom for nBits <= 1024 - 65: `in` and `paddedIn[k < nBits]` are wires, the padding bits
// and the initial hash value own entries, then the compressions in order; digest = out words of the last one
TM HZ_FN void Sha256Bits(Comp<M>& sh, BV32<M>& bits, int nBits, W32<M>* digest) {
    const int nblocks = (nBits + 64) / 512 + 1, total = nblocks * 512;
    for (int k = 0; k < nBits; k++) { sh.alias(Nm("in", k), bits.get(k)); sh.alias(Nm("paddedIn", k), bits.get(k)); }
    bits.put(nBits, sh.def(Nm("paddedIn", nBits), Ops<M>::K(1)));
    for (int k = nBits + 1; k < total - 64; k++) bits.put(k, sh.def(Nm("paddedIn", k), Ops<M>::K(0)));
    for (int k = total - 64; k < total; k++) {
        int kk = total - k - 1;
        bits.put(k, sh.def(Nm("paddedIn", k), Ops<M>::K(kk < 32 ? (((uint32_t)nBits >> kk) & 1u) : 0)));
    }
    W32<M> hin[8], outw[8], msg[16];
    const char* han[8] = {"ha0", "ha1", "ha2", "ha3", "ha4", "ha5", "ha6", "ha7"};
    for (int j = 0; j < 8; j++) {
        Comp<M> hc = sh.sub(han[j]);
        hin[j] = w32_def(hc, "out", hz_sha_h(j), 32);
    }
    for (int b = 0; b < nblocks; b++) {
        for (int t = 0; t < 16; t++) msg[t] = w32_from_bits(bits, 512 * b + 32 * t);
        Sha256compression(sh.sub(Nm("sha256compression", b)), hin, msg, outw);
        for (int j = 0; j < 8; j++) hin[j] = outw[j];
    }
    for (int j = 0; j < 8; j++) digest[j] = outw[j];
}

// src/withdraw.circom:84-176
TM HZ_FN V<M> HashInputsWithdrawal(Comp<M>&& c, int nLevels, const V<M>& rootExit_, const V<M>& ethAddr_,
                                   const V<M>& tokenID_, const V<M>& balance_, const V<M>& idx_) {
    V<M> rootExit = c.def("rootExit", rootExit_), ethAddr = c.def("ethAddr", ethAddr_), tokenID = c.def("tokenID", tokenID_);
    V<M> balance = c.def("balance", balance_), idx = c.def("idx", idx_);
    // hashInputsOut is wired to n2bHashInputsOut.out
    BV32<M> pre; pre.clear(1024);
    int pos = 0;
    { BV<M> b = Num2Bits(c.sub("n2bRootExit"), 256, rootExit); for (int k = 0; k < 256; k++) pre.put(pos + k, b.get(255 - k)); pos += 256; }
    { BV<M> b = Num2Bits(c.sub("n2bEthAddr"), 160, ethAddr); for (int k = 0; k < 160; k++) pre.put(pos + k, b.get(159 - k)); pos += 160; }
    { BV<M> b = Num2Bits(c.sub("n2bTokenID"), 32, tokenID); for (int k = 0; k < 32; k++) pre.put(pos + k, b.get(31 - k)); pos += 32; }
    { BV<M> b = Num2Bits(c.sub("n2bBalance"), 192, balance); for (int k = 0; k < 192; k++) pre.put(pos + k, b.get(191 - k)); pos += 192; }
    {
        BV<M> b = Num2Bits(c.sub("n2bIdx"), 48, idx);
        if (c.sanity()) { uint32_t s = 0; for (int k = nLevels; k < 48; k++) s += b.bit(k); c.check0(Ops<M>::K(s), __LINE__); }
        for (int k = 0; k < 48; k++) pre.put(pos + k, b.get(47 - k));
        pos += 48;
    }
    W32<M> digest[8];
    BV<M> bits; bits.clear(256);
    {
        Comp<M> sh = c.sub("inputsHasher");
        Sha256Bits<M>(sh, pre, pos, digest);
        for (int q = 0; q < 256; q++) {
            V<M> h = w32_get(digest[q >> 5], 31 - (q & 31));      // out[q], MSB first
            sh.alias(Nm("out", q), h);
            bits.put(255 - q, h);
        }
    }
    V<M> o = Bits2Num(c.sub("n2bHashInputsOut"), bits, 0, 256);
    return c.alias("hashInputsOut", o);
}

// main-component inputs in declaration order (src/withdraw.circom:25-33)
enum HzWIn { WIN_rootExit, WIN_ethAddr, WIN_tokenID, WIN_balance, WIN_idx, WIN_sign, WIN_ay, WIN_siblingsState, WIN_COUNT };
static const char* const HZ_WITHDRAW_INPUT_NAMES[WIN_COUNT] = {"rootExit", "ethAddr", "tokenID", "balance", "idx", "sign", "ay", "siblingsState"};
HZ_HD uint32_t hz_withdraw_ninputs(int nLevels) { return 7u + (uint32_t)nLevels + 1u; }

// the whole main component; `c` = main positioned after its inputs (witness index 2 + ninputs), inputs at w[2 ..]
TM HZ_FN V<M> sec_withdraw(Comp<M>& c, int nLevels) {
    V<M> in[7], sib[HZ_MAX_LEVELS];
    for (int k = 0; k < 7; k++) in[k] = c.load(2 + (uint32_t)k);
    for (int k = 0; k < nLevels + 1; k++) sib[k] = c.load(9 + (uint32_t)k);
    V<M> st = HashState(c.sub("accountState"), in[WIN_tokenID], Ops<M>::K(0), in[WIN_sign], in[WIN_balance], in[WIN_ay], in[WIN_ethAddr]);
    SMTVerifier(c.sub("smtVerify"), nLevels + 1, Ops<M>::K(1), in[WIN_rootExit], sib, Ops<M>::K(0), Ops<M>::K(0), Ops<M>::K(0),
                in[WIN_idx], st, Ops<M>::K(0));
    return HashInputsWithdrawal(c.sub("hasherInputs"), nLevels, in[WIN_rootExit], in[WIN_ethAddr], in[WIN_tokenID], in[WIN_balance], in[WIN_idx]);
}
