// Per tx slot / fee tx kernels that start from the main inputs (see hzw_dev.cuh for the map of the kernels).
#include "hzw_dev.cuh"

// [s0, s0 + ns): the tx slots of this shard (all of them on one GPU)
// One THREAD per slot (Comp::solo): with the hashes gone to the transposed kernels what is left of decodeTx[i] and
// rollupTx[i] is bit plumbing and selections -- scalar work that a lane group would repeat in all of its eight lanes.
// A warp walks 32 consecutive slots; slots of one kind follow the same path through the templates.
#define HZ_FRONT_BLOCK 32
__global__ void __launch_bounds__(HZ_FRONT_BLOCK) k_front(HzPlan pl, Fr* w, uint8_t* preb, HzErrBuf* err, HzEddsaJob* jobs, HzSmtJob* smt, Fr* isExit, int s0, int ns) {
    int g = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const bool real = g < ns;
    if (!real) g = ns - 1;                   // padding lanes redo the last slot (identical stores): the warp stays whole
    const int i = s0 + g;
    PreBits<MCompute> pre; pre.b = preb;
    SlotHook<MCompute> hook; hook.eddsa = &jobs[i]; hook.smt = &smt[2 * i]; hook.isExit = &isExit[i]; hook.sz = &pl.sz;
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, pl.off_decode[i], real ? err : nullptr, nullptr, &pend);
    c.solo = true;
    c.skip_hash = true;                      // the message hash and the four state hashes: k_sighash, k_hstate
    DecodeOut<MCompute> d;
    sec_decode(c, pl, i, d, pre);
    c.cur = pl.off_rollup[i];
    RollupTxOut<MCompute> o;
    sec_rollup(c, pl, i, d, o, pre, hook);
    __syncwarp();
    c.flush_pending();
}

// ---- the EdDSA job of a slot straight from the main inputs -------------------------------------------
// The signature is the longest chain of a slot, and what it needs from the rest of the slot is little:
// the message hash (src/decode-tx.circom:254-283: Poseidon(6) over inputs and one repacked element),
// the signer's key (state 1, or the compressed key of an account being created:
// src/rollup-tx.circom:331-343,454-482) and verifySignEnabled (src/rollup-tx-states.circom:98-103,222).
// This kernel derives them from the inputs with the formulas of those templates, so k_eddsa starts at
// once instead of after k_front.  k_front still records the job the templates arrive at; k_tail compares
// the two when the sanity check is on.
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_edjob(HzPlan pl, const Fr* w, HzEddsaJob* jobs, const uint32_t* off_eddsa, Fr* sig_in, int s0, int ns) {
    bool real; int i = s0 + hz_item(ns, real);
    const int lane = (int)hz_glane();
    auto in = [&](int f) -> Fr { return w[pl.in_off[f] + (uint32_t)i * pl.in_d1[f]]; };
    const Fr txc = in(IN_txCompressedData), oc = in(IN_onChain), na = in(IN_newAccount);
    // states: finalFromIdx = Mux1(fromIdx, auxFromIdx, onChain * newAccount), fromIdx = bits 48..95 of txCompressedData
    Fr fromIdx = fr_zero();
    fromIdx.l[0] = (txc.l[1] >> 16) | (txc.l[2] << 16);
    fromIdx.l[1] = txc.l[2] >> 16;
    const Fr ins1 = fr_mul(oc, na);                                   // isP1Insert
    const Fr fin = fr_add(fr_mul(fr_sub(in(IN_auxFromIdx), fromIdx), ins1), fromIdx);
    const Fr vse = fr_mul(fr_sub(fr_one(), oc), fr_is_zero(fin) ? fr_zero() : fr_one());
    // decodeTx: element 1 of the message = toEthAddr (160 bits) | amountF (40) << 160 | maxNumBatch (32) << 200
    const Fr eth = in(IN_toEthAddr), am = in(IN_amountF), mnb = in(IN_maxNumBatch);
    Fr e1;
#pragma unroll
    for (int k = 0; k < 5; k++) e1.l[k] = eth.l[k];
    e1.l[5] = am.l[0]; e1.l[6] = (am.l[1] & 0xffu) | (mnb.l[0] << 8); e1.l[7] = mnb.l[0] >> 24;
    Fr hs = fr_zero();
    if (lane == 1) hs = txc; else if (lane == 2) hs = e1; else if (lane == 3) hs = in(IN_toBjjAy);
    else if (lane == 4) hs = in(IN_rqTxCompressedDataV2); else if (lane == 5) hs = in(IN_rqToEthAddr); else if (lane == 6) hs = in(IN_rqToBjjAy);
    if (lane >= 1 && lane <= 6) sig_in[6 * (size_t)i + (lane - 1)] = hs;          // k_sighash hashes them into jobs[i].M
    // decodeFromBjj: ay = bits 0..253 (mod p), sign = bit 255; the lanes read every 8th bit each
    uint32_t bl[8];
#pragma unroll
    for (int k = 0; k < 8; k++) bl[k] = 0;
    const uint32_t bjj0 = pl.in_off[IN_fromBjjCompressed] + (uint32_t)i * pl.in_d1[IN_fromBjjCompressed];
    for (int k = lane; k < 256; k += HZ_G) bl[k >> 5] |= (w[bjj0 + k].l[0] & 1u) << (k & 31);
#pragma unroll
    for (int o = 1; o < HZ_G; o <<= 1)
#pragma unroll
        for (int k = 0; k < 8; k++) bl[k] |= __shfl_xor_sync(0xffffffffu, bl[k], o, HZ_G);
    Fr bay;
#pragma unroll
    for (int k = 0; k < 8; k++) bay.l[k] = bl[k];
    const Fr bsign = fr_from_u64(bl[7] >> 31);
    bay.l[7] &= 0x3fffffffu;
    if (fr_geq_p(bay)) fr_sub_p(bay);
    const Fr sign1 = in(IN_sign1), ay1 = in(IN_ay1);
    const Fr s1Sign = fr_add(fr_mul(fr_sub(bsign, sign1), ins1), sign1);
    const Fr s1Ay = fr_add(fr_mul(fr_sub(bay, ay1), ins1), ay1);
    if (lane == 0) {
        HzEddsaJob& j = jobs[i];
        j.aySignature = fr_mul(s1Ay, vse); j.signSignature = fr_mul(s1Sign, vse); j.enabled = vse; j.s1Ay = s1Ay;
        j.S = in(IN_s); j.R8x = in(IN_r8x); j.R8y = in(IN_r8y); j.cur = off_eddsa[i];
    }
}

__global__ void __launch_bounds__(HZ_BLOCK) k_fee_front(HzPlan pl, Fr* w, HzErrBuf* err, HzSmtJob* smt, Fr* hs_in, int f0, int nf) {
    bool real; int i = f0 + hz_item(nf, real);
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, pl.off_fee[i], real ? err : nullptr, nullptr, &pend);
    int hs_n = 0;
    c.skip_hash = true; c.hs_rec = hs_in + 4 * (size_t)(4 * pl.P.nTx + 2 * i); c.hs_n = &hs_n;      // its two state hashes: k_hstate
    FeeHook<MCompute> hook; hook.smt = &smt[i];
    hook.sib_idx = pl.in_off[IN_siblings3] + (uint32_t)i * pl.in_d1[IN_siblings3];
    sec_fee(c, pl, i, hook, true);
    c.flush_pending();
}
