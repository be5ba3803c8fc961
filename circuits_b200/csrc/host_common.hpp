// Host-side pieces shared by the CUDA library (hzw.cu) and the CPU port used as checker
// (oracle/cpu_port): input table, the layout walk that builds the kernel plan and resolves
// signal names, and the `.wtns` writer.  No field arithmetic happens here.
#pragma once
#include "rollup_main.cuh"
#include "withdraw.cuh"
#include "rollup_tx.cuh"
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>

static const char* const HZ_INPUT_NAMES[IN_COUNT] = {
    "oldLastIdx", "oldStateRoot", "globalChainID", "currentNumBatch", "feeIdxs", "feePlanTokens",
    "imOnChain", "imOutIdx", "imStateRoot", "imExitRoot", "imAccFeeOut", "imStateRootFee",
    "imInitStateRootFee", "imFinalAccFee",
    "txCompressedData", "amountF", "txCompressedDataV2", "fromIdx", "auxFromIdx", "toIdx", "auxToIdx",
    "toBjjAy", "toEthAddr", "maxNumBatch", "onChain", "newAccount", "rqOffset", "rqTxCompressedDataV2",
    "rqToEthAddr", "rqToBjjAy", "s", "r8x", "r8y", "loadAmountF", "fromEthAddr", "fromBjjCompressed",
    "tokenID1", "nonce1", "sign1", "balance1", "ay1", "ethAddr1", "siblings1", "isOld0_1", "oldKey1",
    "oldValue1", "tokenID2", "nonce2", "sign2", "balance2", "ay2", "ethAddr2", "siblings2", "newExit",
    "isOld0_2", "oldKey2", "oldValue2", "tokenID3", "nonce3", "sign3", "balance3", "ay3", "ethAddr3",
    "siblings3"};

// extents of every main input (src/rollup-main.circom:105-196); rank 0 = scalar
static inline void hz_input_shape(const HzParams& P, int f, int& rank, uint32_t& d0, uint32_t& d1) {
    const uint32_t nTx = P.nTx, L = P.nLevels + 1, F = P.maxFeeTx;
    rank = 1; d0 = nTx; d1 = 1;
    switch (f) {
        case IN_oldLastIdx: case IN_oldStateRoot: case IN_globalChainID: case IN_currentNumBatch:
        case IN_imInitStateRootFee:
            rank = 0; d0 = 1; break;
        case IN_feeIdxs: case IN_feePlanTokens: case IN_imFinalAccFee:
        case IN_tokenID3: case IN_nonce3: case IN_sign3: case IN_balance3: case IN_ay3: case IN_ethAddr3:
            d0 = F; break;
        case IN_imOnChain: case IN_imOutIdx: case IN_imStateRoot: case IN_imExitRoot:
            d0 = nTx - 1; break;
        case IN_imAccFeeOut: rank = 2; d0 = nTx - 1; d1 = F; break;
        case IN_imStateRootFee: d0 = F - 1; break;
        case IN_fromBjjCompressed: rank = 2; d1 = 256; break;
        case IN_siblings1: case IN_siblings2: rank = 2; d1 = L; break;
        case IN_siblings3: rank = 2; d0 = F; d1 = L; break;
        default: break;
    }
}

struct HzLayout {
    HzPlan plan;
    std::vector<uint32_t> off_decode, off_rollup, off_fee, off_eddsa, off_jk;   // off_eddsa: start of phase F (getAx) per slot; off_jk: of phases J-K (processor1)
    std::vector<uint32_t> off_hs, off_hsig;   // where the Poseidon components of the state hashes (4 per slot, then 2 per fee tx) and of the message hashes start
};

static inline bool hz_params_ok(const HzParams& P) {
    return P.nTx >= 1 && P.nLevels >= 1 && P.nLevels <= 48 && P.maxL1Tx >= 0 && P.maxL1Tx <= P.nTx &&
           P.maxFeeTx >= 1 && P.maxFeeTx <= HZ_MAX_FEE;
}

// The layout walk: runs every section in circuit order with MLayout (no arithmetic).
static inline void hz_layout_walk(const HzParams& P, LayoutSink* sink, HzLayout& R) {
    HzPlan& pl = R.plan;
    memset(&pl, 0, sizeof pl);
    pl.P = P;
    pl.nbits = hz_pre_nbits(P);
    pl.nblocks = (pl.nbits + 64) / 512 + 1;
    HzPoseidonStarts starts;
    hz_poseidon_starts = &starts;
    Comp<MLayout> main(1, sink, "main");
    main.sig("hashGlobalInputs");
    for (int f = 0; f < IN_COUNT; f++) {
        int rank; uint32_t d0, d1;
        hz_input_shape(P, f, rank, d0, d1);
        pl.in_off[f] = main.cur; pl.in_d0[f] = d0; pl.in_d1[f] = d1;
        if (main.named) {
            for (uint32_t a = 0; a < d0; a++)
                for (uint32_t b = 0; b < d1; b++)
                    main.name(rank == 0 ? Nm(HZ_INPUT_NAMES[f]) : (rank == 1 ? Nm(HZ_INPUT_NAMES[f], (int)a) : Nm(HZ_INPUT_NAMES[f], (int)a, (int)b)),
                              main.cur + a * d1 + b);
        }
        main.cur += d0 * d1;
    }
    pl.n_inputs = main.cur - 2;
    std::vector<uint32_t> preidx((size_t)pl.nblocks * 512, HZ_NOIDX);
    PreBits<MLayout> pre; pre.idx = preidx.data();
    std::vector<DecodeOut<MLayout> > dec(P.nTx);
    R.off_decode.resize(P.nTx); R.off_rollup.resize(P.nTx); R.off_fee.resize(P.maxFeeTx); R.off_eddsa.resize(P.nTx); R.off_jk.resize(P.nTx);
    for (int i = 0; i < P.nTx; i++) { R.off_decode[i] = main.cur; sec_decode(main, pl, i, dec[i], pre); }
    pl.idx_newLastIdx = dec[P.nTx - 1].outIdx.idx;
    SlotHook<MLayout> hook; hook.sz_out = &pl.sz;
    for (int i = 0; i < P.nTx; i++) {
        R.off_rollup[i] = main.cur;
        hook.ed_cur = &R.off_eddsa[i];
        hook.jk_cur = &R.off_jk[i];
        RollupTxOut<MLayout> o;
        sec_rollup(main, pl, i, dec[i], o, pre, hook);
        if (i == P.nTx - 1) pl.idx_newExitRoot = o.newExitRoot.idx;
    }
    for (int i = 0; i < P.maxFeeTx; i++) {
        R.off_fee[i] = main.cur;
        FeeHook<MLayout> fh; fh.unused = 0;
        VL r = sec_fee(main, pl, i, fh);
        if (i == P.maxFeeTx - 1) pl.idx_newStateRoot = r.idx;
    }
    pl.off_hasher = main.cur;
    {
        Comp<MLayout> hc = main.sub("hasherInputs");
        W32<MLayout> hin[8], outw[8], msg[16];
        sec_hasher_head(hc, pl, pre, &pl.off_sha_blocks, hin);
        std::string shp = hc.prefix.empty() ? std::string() : hc.prefix + ".inputsHasher";
        Comp<MLayout> sh(pl.off_sha_blocks, sink, shp);
        for (uint32_t b = 0; b < pl.nblocks; b++) {
            for (int t = 0; t < 16; t++) for (int k = 0; k < 32; k++) msg[t].idx[k] = preidx[(size_t)b * 512 + 32 * t + 31 - k];
            sec_sha_block(sh, (int)b, hin, msg, outw);
            for (int j = 0; j < 8; j++) hin[j] = outw[j];
        }
        hc.cur = sh.cur;
        pl.off_hasher_tail = hc.cur;
        VL h = sec_hasher_tail(hc, sh, outw);
        (void)h;
    }
    pl.nvars = main.cur;
    hz_poseidon_starts = nullptr;
    R.off_hs = starts.t5; R.off_hsig = starts.t7;
    if (R.off_hs.size() != (size_t)(4 * P.nTx + 2 * P.maxFeeTx) || R.off_hsig.size() != (size_t)P.nTx) {
        fprintf(stderr, "hzw layout: %zu state hashes and %zu message hashes walked\n", R.off_hs.size(), R.off_hsig.size());
        abort();
    }
    pl.off_decode = R.off_decode.data(); pl.off_rollup = R.off_rollup.data(); pl.off_fee = R.off_fee.data();
}

// ---- Withdraw(nLevels): layout walk -------------------------------------------------------------
struct HzWLayout { int nLevels = 0; uint32_t n_inputs = 0, nvars = 0; };
static inline void hz_withdraw_walk(int nLevels, LayoutSink* sink, HzWLayout& R) {
    R.nLevels = nLevels;
    R.n_inputs = hz_withdraw_ninputs(nLevels);
    Comp<MLayout> main(1, sink, "main");
    main.sig("hashGlobalInputs");
    if (main.named) {
        for (int f = 0; f < WIN_siblingsState; f++) main.name(Nm(HZ_WITHDRAW_INPUT_NAMES[f]), main.cur + (uint32_t)f);
        for (int k = 0; k < nLevels + 1; k++) main.name(Nm("siblingsState", k), main.cur + 7u + (uint32_t)k);
    }
    main.cur += R.n_inputs;
    sec_withdraw(main, nLevels);
    R.nvars = main.cur;
}

// `.wtns` container, snarkjs 0.3.59 wtns_utils.writeBin layout
#define HZ_WTNS_HEADER 76
static inline void hz_wtns_header(uint8_t hdr[HZ_WTNS_HEADER], uint64_t nvars) {
    uint8_t* p = hdr;
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) *p++ = (uint8_t)(v >> (8 * i)); };
    auto u64 = [&](uint64_t v) { for (int i = 0; i < 8; i++) *p++ = (uint8_t)(v >> (8 * i)); };
    memcpy(p, "wtns", 4); p += 4;
    u32(2); u32(2);
    u32(1); u64(40); u32(32);
    Fr P = fr_P();
    for (int i = 0; i < 8; i++) u32(P.l[i]);
    u32((uint32_t)nvars);
    u32(2); u64(32ull * nvars);
}
static inline int hz_write_wtns(const char* path, const uint8_t* witness, uint64_t nvars) {
    FILE* f = fopen(path, "wb");
    if (!f) return -2;
    uint8_t hdr[HZ_WTNS_HEADER];
    hz_wtns_header(hdr, nvars);
    size_t ok = fwrite(hdr, 1, sizeof hdr, f);
    ok += fwrite(witness, 32, nvars, f);
    if (fclose(f) != 0) return -2;
    return ok == sizeof hdr + nvars ? 0 : -2;
}
