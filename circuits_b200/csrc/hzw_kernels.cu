// libhzwitness.so, CUDA half (sm_100a): device memory, the section kernels and their launch
// order.  Every kernel instantiates the circuit templates in MCompute mode; a section of the
// circuit (rollup_main.cuh) is evaluated by a GROUP of 8 lanes in lock-step (ctx.cuh): glue
// redundantly, Poseidon permutations lane-parallel, stores spread over the lanes.  Each signal is
// stored at the witness index the layout walk assigned.
//
//   k_front      group per tx slot: decodeTx[i] + rollupTx[i] up to the hand-offs (EdDSA, SMT)
//   k_fee_front  group per fee tx: feeTx[i] up to its SMT processor
//   k_smt        group per SMTProcessor (2 per slot + 1 per fee tx): leaf hashes, key bits, state
//                machine and the nLevels+1 levels, old-path and new-path hashes side by side
//   k_eddsa      group per tx slot: getAx + sigVerifier (phase F), concurrent with k_smt
//   k_tail       group per slot / fee tx: processor tops, s3/s4/s5, root checks
//   k_hash_head  hasherInputs up to the first compression (one group)
//   k_sha_chain  SHA-256 chaining values of all blocks (the chain is serial)
//   k_sha_blocks one thread per sha256compression: all bit signals of the block
//   k_hash_tail  digest -> hashGlobalInputs
#define HZ_DEVICE_ONLY
#include "hzw_cuda.h"
#include "../../include/hzwitness.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <chrono>
#include <vector>

#define HZ_NKERN 10  /* timed intervals; k_smt's covers the head and the levels kernels */
static const char* const HZ_KERN_NAMES[HZ_NKERN] = {"k_smtjob", "k_fee_front", "k_smt", "k_tail", "k_hash_head",
                                                    "k_sha_chain", "k_sha_blocks", "k_hash_tail", "k_eddsa", "k_front"};
#define HZ_BLOCK 128                       /* 16 lane groups per block */

struct HzCuda {
    int device;
    HzPlan plan;               // device-pointer version
    Fr* w;                     // witness, nvars entries
    uint8_t* pre;              // SHA pre-image bits, nblocks*512
    uint32_t* chain;           // (nblocks+1) x 8
    uint32_t* msgw;            // nblocks x 16
    uint32_t* shawk;           // nblocks x 64: message schedule + round constants
    HzEddsaJob* jobs;          // nTx: from k_edjob, read by k_eddsa
    HzEddsaJob* jobs_front;    // nTx: the same as the templates of k_front arrive at them (k_tail compares)
    uint32_t* off_eddsa;       // nTx: where phase F of each slot starts
    HzSmtJob* smt;             // 2 nTx + maxFeeTx: from k_smtjob (slots) and k_fee_front
    HzSmtJob* smt_front;       // 2 nTx: the same as k_front arrives at them (k_tail compares)
    uint32_t* off_jk;          // nTx: where phases J-K of each slot start
    HzSmtMid* mids;            // 2 nTx + maxFeeTx
    HzLevJob* lev;             // 2 nTx + maxFeeTx: what k_smt leaves for k_smt_levels
    Fr* smst;                  // per processor (nLevels + 1) x 6 state-machine values
    Fr* img00;                 // the 634 entries of SMTHash2(0, 0) and, at [HZ_IMG00_DIGEST], its digest
    uint8_t* zflag;            // per block of k_smt_levels and level: the level hashed (0, 0) in every chain of the block
    // end-to-end path: the (0, 0) levels at the bottom of every processor are not copied to the host but rebuilt there
    struct HzHole* holes; struct HzHole* h_holes;   // per job: where its levels start, how many bottom levels are (0, 0) images
    Fr* gl; Fr* h_gl;          // per job: the entries between those images (level headers, switchers), compact
    Fr* h_img00;               // host copy of the image
    cudaStream_t stream5, stream6;
    cudaEvent_t ev_pack[2], ev_info[2], ev_sha[2];
    uint64_t last_d2h;
    bool no_holes;             // HZW_NO_HOLES (tests / measurement): copy every entry, rebuild nothing on the host
    std::vector<uint32_t>* h_off_rollup;   // host copy of the slot offsets (plus the end of the last slot)
    Fr* isExit;                // nTx
    Fr* isExit_front;          // nTx: k_front's copy
    Fr* hint_cst;              // HZ_CST_MAX: constant inverses of EscalarMulFix (recorded at create)
    Fr* hint_tab;              // HZ_EDTAB_LEN: its window entries
    Fr* hint_img;              // witness image of EscalarMulFix from the recording pass (constant entries are copied from it)
    Fr* hint_scratch;          // per lane group of k_eddsa: work area of the hint generators
    uint32_t fix_size, win_size, win_stride;   // entries of EscalarMulFix; of the windows + carry adders of its segment 0; per window
    HzErrBuf* err;
    uint32_t *off_decode, *off_rollup, *off_fee;
    uint32_t h_off_decode0, h_off_rollup0, h_off_rollup_mid, h_off_fee0;   // host copies of the piece boundaries
    uint32_t* off_blocks;      // 1
    // shard of this context: rank / world and the ranges it evaluates (everything when world == 1)
    int rank, world, per, s0, ns, f0, nf;
    uint32_t b0, nb;
    uint8_t* recs;             // world x rec_bytes: the gathered exchange records (own one while packing)
    cudaStream_t stream, stream2, stream3, stream4;
    cudaEvent_t ev_half[2], ev_ed[2], ev_done;
    cudaEvent_t ev0, ev1, ev_front, ev_eddsa, ev_lev, ev_zero, ev_fee;
    cudaEvent_t kev[HZ_NKERN + 3];
};

static void hz_set_ranges(HzCuda* d, int rank, int world);
__device__ __forceinline__ void hz_head_err(HzErrBuf* err, uint32_t cur, uint32_t line, const Fr& lhs);
__global__ void k_poseidon_t3(const Fr* in, Fr* sig, Fr* digests, int nperms);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "hzw: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return HZW_E_DEVICE; } } while (0)

// item of this lane group; groups past the end redo the last item (identical stores) so that
// every warp stays whole for the shuffles, and report nothing
__device__ __forceinline__ int hz_item(int n, bool& real) {
    int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) / HZ_G);
    real = g < n;
    return real ? g : n - 1;
}
static inline unsigned hz_grid(int items) { return (unsigned)((items * HZ_G + HZ_BLOCK - 1) / HZ_BLOCK); }

// ---- Poseidon alone (the second metric of BASELINE.json: permutations per second) ----------------
// The SMT level hashes as k_smt evaluates them -- a lane group runs the old-path and the new-path
// Poseidon(2) side by side -- but on free-standing inputs: pair q hashes (in[4q], in[4q+1]) and
// (in[4q+2], in[4q+3]) and stores all 2 x 633 signals of the two components at sig[q * 2 * size].
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_poseidon_pairs(const Fr* in, Fr* sig, Fr* digests, int npairs) {
    bool real; int q = hz_item(npairs, real);
    const uint32_t size = poseidon_size(3);
    Comp<MCompute> ca(sig, (uint32_t)q * 2 * size, nullptr), cb(sig, (uint32_t)q * 2 * size + size, nullptr);
    VC a[2], b[2], oa, ob;
    for (int k = 0; k < 2; k++) { a[k].v = in[4 * q + k]; a[k].sig = true; b[k].v = in[4 * q + 2 + k]; b[k].sig = true; }
    PoseidonDual<MCompute>(static_cast<Comp<MCompute>&&>(ca), a, static_cast<Comp<MCompute>&&>(cb), b, 2, oa, ob);
    if (hz_glane() == 0) { digests[2 * q] = oa.v; digests[2 * q + 1] = ob.v; }
}

// [s0, s0 + ns): the tx slots of this shard (all of them on one GPU)
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_front(HzPlan pl, Fr* w, uint8_t* preb, HzErrBuf* err, HzEddsaJob* jobs, HzSmtJob* smt, Fr* isExit, int s0, int ns) {
    bool real; int i = s0 + hz_item(ns, real);
    PreBits<MCompute> pre; pre.b = preb;
    SlotHook<MCompute> hook; hook.eddsa = &jobs[i]; hook.smt = &smt[2 * i]; hook.isExit = &isExit[i]; hook.sz = &pl.sz;
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, pl.off_decode[i], real ? err : nullptr, nullptr, &pend);
    DecodeOut<MCompute> d;
    sec_decode(c, pl, i, d, pre);
    c.cur = pl.off_rollup[i];
    RollupTxOut<MCompute> o;
    sec_rollup(c, pl, i, d, o, pre, hook);
    c.flush_pending();
}

// ---- the SMT jobs of a slot ahead of k_front --------------------------------------------------------------
// The merkle kernels are the other long chain of a slot, and what they need from rollupTx[i] are the two job
// records (keys, old / new state hashes, function selectors).  This kernel runs the SAME templates as k_front on
// a dry context -- values only: no entry stored, no inverse queued, the signature and the fee accumulator
// skipped, decodeTx reduced to the fields rollupTx reads -- so the merkle kernels start after a fraction of
// k_front's time, and k_front itself runs beside them.  k_front still records the jobs it arrives at; k_tail
// compares the two (always: a difference is an internal error, HzErrBuf.pad).
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_smtjob(HzPlan pl, Fr* w, HzSmtJob* smt, Fr* isExit, const uint32_t* off_jk, int s0, int ns) {
    bool real; int i = s0 + hz_item(ns, real);
    Comp<MCompute> c(w, pl.off_rollup[i], nullptr);
    c.dry = true;
    DecodeOut<MCompute> d;
    hz_decode_lite(c, pl, i, d);
    PreBits<MCompute> pre; pre.b = nullptr;
    SlotHook<MCompute> hook; hook.eddsa = nullptr; hook.smt = &smt[2 * i]; hook.isExit = &isExit[i]; hook.sz = &pl.sz; hook.jobs_only = true;
    RollupTxOut<MCompute> o;
    sec_rollup(c, pl, i, d, o, pre, hook);
    __syncwarp();
    // the dry walk does not keep the witness cursor: the processors start where the layout walk put them
    if (hz_glane() == 0) { smt[2 * i].cur = off_jk[i]; smt[2 * i + 1].cur = off_jk[i] + pl.sz.p2_off; }
}

// ---- the EdDSA job of a slot straight from the main inputs -------------------------------------------
// The signature is the longest chain of a slot, and what it needs from the rest of the slot is little:
// the message hash (src/decode-tx.circom:254-283: Poseidon(6) over inputs and one repacked element),
// the signer's key (state 1, or the compressed key of an account being created:
// src/rollup-tx.circom:331-343,454-482) and verifySignEnabled (src/rollup-tx-states.circom:98-103,222).
// This kernel derives them from the inputs with the formulas of those templates, so k_eddsa starts at
// once instead of after k_front.  k_front still records the job the templates arrive at; k_tail compares
// the two when the sanity check is on.
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_edjob(HzPlan pl, const Fr* w, HzEddsaJob* jobs, const uint32_t* off_eddsa, int s0, int ns) {
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
    const Fr M = hz_poseidon_lanes<7, true>(nullptr, 0, hs, 0);
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
        j.S = in(IN_s); j.R8x = in(IN_r8x); j.R8y = in(IN_r8y); j.M = M; j.cur = off_eddsa[i];
    }
}

__global__ void __launch_bounds__(HZ_BLOCK) k_fee_front(HzPlan pl, Fr* w, HzErrBuf* err, HzSmtJob* smt, int f0, int nf) {
    bool real; int i = f0 + hz_item(nf, real);
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, pl.off_fee[i], real ? err : nullptr, nullptr, &pend);
    FeeHook<MCompute> hook; hook.smt = &smt[i];
    hook.sib_idx = pl.in_off[IN_siblings3] + (uint32_t)i * pl.in_d1[IN_siblings3];
    sec_fee(c, pl, i, hook, true);
    c.flush_pending();
}

// jobs of the shard: 2 per own slot, then 1 per own fee tx
__device__ __forceinline__ int hz_smt_job_index(const HzPlan& pl, int q, int s0, int ns, int f0) {
    return q < 2 * ns ? 2 * s0 + q : 2 * pl.P.nTx + f0 + (q - 2 * ns);
}
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_smt(HzPlan pl, Fr* w, HzErrBuf* err, const HzSmtJob* smt, HzSmtMid* mids, HzLevJob* lev, Fr* smst,
                                                  int s0, int ns, int f0, int nf) {
    bool real; int q = hz_item(2 * ns + nf, real);
    q = hz_smt_job_index(pl, q, s0, ns, f0);
    const int L = pl.P.nLevels + 1;
    sec_smt_head(w, real ? err : nullptr, L, smt[q], mids[q], lev[q], smst + (size_t)q * L * 6);
}

// ---- the levels of all SMT processors, transposed ------------------------------------------------------
// 132 of the 142 Poseidon permutations of a transaction are the old-path / new-path hashes of the
// 2 x 33 levels.  A chain = one path of one processor: 33 Poseidon(2) one after the other plus a few
// selections per level.  Thread (warp r, lane c) holds state element r of chain c, a block of 3 warps
// runs 32 chains (16 processors): the S-box of a partial round is one full warp, the MDS rows meet in
// shared memory once per round (k_poseidon_t3 is this kernel without the level glue).  Warp 0 also
// does the per-level glue of its chain (switcher in front of the hash, root behind it).  Entry
// offsets inside a level follow SMTProcessorLevel (templates_circomlib.cuh): oldRoot, newRoot,
// [oldChild, newChild: bottom level only], aux[4], oldProofHash (634), newProofHash (634),
// oldSwitcher (outL, outR, aux), newSwitcher (L, R, outL, outR, aux).
// Levels below the leaf of a sparse tree hash (0, 0): child and sibling are both zero there, on the old path
// and on the new one.  With 2^12 accounts in a tree of depth 32 that is some 19 of the 33 levels of every
// processor.  Their 634 entries do not depend on anything, so when a level is (0, 0) in EVERY chain of the
// block the 65 rounds are skipped, the digest is the constant, and the entries are copied from an image
// recorded at context creation by k_smt_zero (a wide copy, off the critical chain of the levels).
#define HZ_IMG00_DIGEST 640
__global__ void __launch_bounds__(96) k_smt_levels(HzPlan pl, Fr* w, const HzLevJob* jobs, const Fr* smst, HzSmtMid* mids,
                                                   const Fr* img00, uint8_t* zflag, int s0, int ns, int f0, int nf) {
    __shared__ uint32_t xch[2][3][8][32];
    __shared__ uint32_t hin[2][8][32];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    const int njobs = 2 * ns + nf, L = pl.P.nLevels + 1;
    int q = (int)blockIdx.x * 16 + (c >> 1);
    if (q >= njobs) q = njobs - 1;                    // padding chains redo the last processor (identical stores)
    const int job = hz_smt_job_index(pl, q, s0, ns, f0);
    const int side = c & 1;                           // 0: old path, 1: new path
    const HzLevJob& J = jobs[job];
    const Fr* stj = smst + (size_t)job * L * 6;
    const Fr o1 = J.o1, n1 = J.n1;
    const uint32_t HS = HZ_SZ_SMTHASH2_WIRED;
    const Fr* C = &hz_poseidon_c(3, 0);
    Fr m[3];
    for (int k = 0; k < 3; k++) m[k] = hz_poseidon_mr(3, r)[k];
    Fr child = fr_zero();
    for (int lev = L - 1; lev >= 0; lev--) {
        const bool bottom = lev == L - 1;
        const uint32_t lb = J.cur + (bottom ? 0u : 2u + (uint32_t)(L - 1 - lev) * HZ_SZ_SMT_LEVEL);
        const uint32_t o_aux = lb + 2 + (bottom ? 2u : 0u), o_oph = o_aux + 4, o_nph = o_oph + HS, o_osw = o_nph + HS, o_nsw = o_osw + 3;
        Fr st_top, st_old0, st_bot, st_new1, st_upd;
        if (r == 0) {
            st_top = stj[lev * 6 + 0]; st_old0 = stj[lev * 6 + 1]; st_bot = stj[lev * 6 + 2]; st_new1 = stj[lev * 6 + 3]; st_upd = stj[lev * 6 + 5];
            const Fr sib = w[J.sib + lev];
            const Fr lr = fr_from_u64((J.lrbits >> lev) & 1ull);
            Fr hL, hR;
            if (side == 0) {
                if (bottom) hz_store_fr(&w[lb + 2], fr_zero());
                Fr aux = fr_mul(fr_sub(sib, child), lr);
                hL = fr_add(aux, child); hR = fr_sub(sib, aux);
                hz_store_fr(&w[o_osw], hL); hz_store_fr(&w[o_osw + 1], hR); hz_store_fr(&w[o_osw + 2], aux);
            } else {
                if (bottom) hz_store_fr(&w[lb + 3], fr_zero());
                Fr a1 = fr_mul(child, fr_add(st_top, st_bot)), a2 = fr_mul(sib, st_top);
                Fr Ln = fr_add(a1, fr_mul(n1, st_new1)), Rn = fr_add(a2, fr_mul(o1, st_new1));
                Fr aux = fr_mul(fr_sub(Rn, Ln), lr);
                hL = fr_add(aux, Ln); hR = fr_sub(Rn, aux);
                hz_store_fr(&w[o_aux + 1], a1); hz_store_fr(&w[o_aux + 2], a2);
                hz_store_fr(&w[o_nsw], Ln); hz_store_fr(&w[o_nsw + 1], Rn); hz_store_fr(&w[o_nsw + 2], hL);
                hz_store_fr(&w[o_nsw + 3], hR); hz_store_fr(&w[o_nsw + 4], aux);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) { hin[0][k][c] = hL.l[k]; hin[1][k][c] = hR.l[k]; }
        }
        __syncthreads();
        Fr s = fr_zero();
        if (r > 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) s.l[k] = hin[r - 1][k][c];
        }
        const int allz = __syncthreads_and(fr_is_zero(s) ? 1 : 0);
        if (threadIdx.x == 0) zflag[(size_t)blockIdx.x * L + lev] = allz ? 1 : 0;
        // SMTHash2 (wired L, R) -> Poseidon(2): entry 0 is ark[0].in[0] = 0
        Fr* hw = w + (side == 0 ? o_oph : o_nph);
        if (allz) s = img00[HZ_IMG00_DIGEST];
        else if (r == 0) hz_store_fr(&hw[0], s);
        uint32_t cur = 1;
        for (int i = 0; i < (allz ? 0 : 65); i++) {
            s = fr_add(s, C[3 * i + r]);
            hz_store_fr(&hw[cur + r], s);
            const bool full = i < 4 || i >= 61;
            const uint32_t sg = cur + 6;
            if (full || r == 0) {                       // warp-uniform
                Fr in2 = fr_mul_chain(s, s);
                Fr in4 = fr_mul_chain(in2, in2);
                s = fr_mul_chain(in4, s);
                hz_store_fr(&hw[sg + 3 * r], s); hz_store_fr(&hw[sg + 3 * r + 1], in2); hz_store_fr(&hw[sg + 3 * r + 2], in4);
            }
            uint32_t (*buf)[8][32] = xch[i & 1];
#pragma unroll
            for (int k = 0; k < 8; k++) buf[r][k][c] = s.l[k];
            __syncthreads();
            Fr a[3];
#pragma unroll
            for (int qq = 0; qq < 3; qq++)
#pragma unroll
                for (int k = 0; k < 8; k++) a[qq].l[k] = buf[qq][k][c];
            s = fr_dot_mont_chain<3>(a, m);
            hz_store_fr(&hw[cur + 3 + r], s);
            cur += 6 + (full ? 9 : 3);
        }
        if (r == 0) {
            if (side == 0) {
                Fr a0 = fr_mul(o1, fr_add(fr_add(st_bot, st_new1), st_upd));
                child = fr_add(a0, fr_mul(s, st_top));
                hz_store_fr(&w[o_aux], a0); hz_store_fr(&w[lb], child);
            } else {
                Fr a3 = fr_mul(s, fr_add(fr_add(st_top, st_bot), st_new1));
                child = fr_add(a3, fr_mul(n1, fr_add(st_old0, st_upd)));
                hz_store_fr(&w[o_aux + 3], a3); hz_store_fr(&w[lb + 1], child);
            }
        }
    }
    if (r == 0) { if (side == 0) mids[job].oc = child; else mids[job].nc = child; }
}

// ---- end-to-end path: the bottom (0, 0) levels stay on the device ---------------------------------------------
// Those levels are 40 % of the witness of a sparse tree and all alike: instead of crossing PCIe they are rebuilt in
// the caller's buffer by host threads from one copy of the image (computed on the device at context creation),
// while the rest of the witness is copied around them.  What the host needs per processor: where its levels
// start, how many bottom levels are images (`nz`: the leading run of flagged levels of its block), and the few
// entries between the images (level headers and switchers), which k_hole_pack gathers into a compact array.
struct HzHole { uint32_t cur, nz; };
#define HZ_HS 634u                          /* poseidon_size(3): entries of SMTHash2 (checked at context creation) */
#define HZ_SZL (2u + 4u + 2u * HZ_HS + 3u + 5u)   /* HZ_SZ_SMT_LEVEL */
#define HZ_GL_STRIDE(L) (2u + 14u * (uint32_t)(L))
// level k from the bottom of a processor whose levels start at cur: its first entry, its old-path hash
__host__ __device__ static inline uint32_t hz_lev_base(uint32_t cur, uint32_t k) { return k == 0 ? cur : cur + 2u + k * HZ_SZL; }
__host__ __device__ static inline uint32_t hz_lev_oph(uint32_t cur, uint32_t k) { return hz_lev_base(cur, k) + (k == 0 ? 8u : 6u); }
// witness index of glue entry j of the hole: the bottom header (8), then per level its switchers (8) and the next header (6)
__host__ __device__ static inline uint32_t hz_hole_glue_index(uint32_t cur, uint32_t j) {
    if (j < 8u) return cur + j;
    const uint32_t p = j - 8u, k = p / 14u, r = p % 14u;
    return hz_lev_oph(cur, k) + 2u * HZ_HS + r;          // switchers of level k, running on into the header of level k + 1
}
__global__ void __launch_bounds__(128) k_hole_pack(HzPlan pl, const Fr* w, const HzLevJob* jobs, const uint8_t* zflag, HzHole* holes, Fr* gl,
                                                   int s0, int ns, int f0, int nf) {
    const int njobs = 2 * ns + nf, L = pl.P.nLevels + 1;
    const int q = (int)blockIdx.x;
    if (q >= njobs) return;
    const int job = hz_smt_job_index(pl, q, s0, ns, f0);
    const int b = q / 16;
    int nz = 0;
    while (nz < L && zflag[(size_t)b * L + (L - 1 - nz)]) nz++;
    const uint32_t cur = jobs[job].cur;
    if (threadIdx.x == 0) { holes[job].cur = cur; holes[job].nz = (uint32_t)nz; }
    const uint32_t G = nz ? 2u + 14u * (uint32_t)nz : 0u;
    Fr* dst = gl + (size_t)job * HZ_GL_STRIDE(L);
    for (uint32_t j = threadIdx.x; j < G; j += blockDim.x) dst[j] = w[hz_hole_glue_index(cur, j)];
}

// the entries of the (0, 0) levels k_smt_levels skipped: block (b, lev) copies the image into the 32 chains of
// levels-block b when that level was flagged, a warp per 4 chains
__global__ void __launch_bounds__(256) k_smt_zero(HzPlan pl, Fr* w, const HzLevJob* jobs, const Fr* img00, const uint8_t* zflag,
                                                  int s0, int ns, int f0, int nf) {
    const int njobs = 2 * ns + nf, L = pl.P.nLevels + 1;
    const int b = (int)blockIdx.x, lev = (int)blockIdx.y;
    if (!zflag[(size_t)b * L + lev]) return;
    const int wi = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const uint32_t HS = HZ_SZ_SMTHASH2_WIRED;
    const uint4* src = reinterpret_cast<const uint4*>(img00);
    for (int c = wi * 4; c < wi * 4 + 4; c++) {
        int q = b * 16 + (c >> 1);
        if (q >= njobs) q = njobs - 1;
        const HzLevJob& J = jobs[hz_smt_job_index(pl, q, s0, ns, f0)];
        const bool bottom = lev == L - 1;
        const uint32_t lb = J.cur + (bottom ? 0u : 2u + (uint32_t)(L - 1 - lev) * HZ_SZ_SMT_LEVEL);
        const uint32_t o_oph = lb + 2 + (bottom ? 2u : 0u) + 4;
        uint4* dst = reinterpret_cast<uint4*>(w + ((c & 1) == 0 ? o_oph : o_oph + HS));
        for (uint32_t e = (uint32_t)lane; e < 2 * HS; e += 32) __stcs(&dst[e], src[e]);
    }
}

// one-time: the constant division sites and window entries of EscalarMulFix(253, BASE8)
__global__ void k_fix_consts(Fr* wtmp, Fr* cst, Fr* tab, uint32_t* counts) {
    hz_record_fix_consts(wtmp, cst, tab, counts);
}

// the constant entries of EscalarMulFix(253, BASE8) for every signature of the shard: its recorded image copied
// to sigVerifier.mulFix of each slot before k_eddsa runs (which then only writes the entries that depend on S)
__global__ void __launch_bounds__(256) k_ed_const(HzPlan pl, Fr* w, const uint32_t* off_eddsa, const Fr* img, uint32_t fix_size, int s0, int ns) {
    const uint32_t per = 2 * fix_size;                                  // uint4 pieces per signature
    const uint64_t total = (uint64_t)per * (uint64_t)ns;
    const uint4* src = reinterpret_cast<const uint4*>(img);
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t g = (uint32_t)(e / per), k = (uint32_t)(e % per);
        uint4* dst = reinterpret_cast<uint4*>(w + off_eddsa[s0 + (int)g] + pl.sz.mulfix_off);
        __stcs(&dst[k], src[k]);
    }
}

__global__ void __launch_bounds__(HZ_BLOCK, 3) k_eddsa(HzPlan pl, Fr* w, HzErrBuf* err, const HzEddsaJob* jobs,
                                                    const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch, uint32_t win_size, int s0, int ns) {
    bool real; int i = s0 + hz_item(ns, real);
    const HzEddsaJob& j = jobs[i];
    HzHints h;
    h.cst = cst; h.ci = 0; h.dat = nullptr; h.di = 0; h.rec = nullptr; h.ri = 0;
    h.edtab = tab; h.edrec = nullptr; h.ei = 0; h.lad_ok = 0; h.fill = 1; h.prefilled = 1; h.win_c0 = 0; h.win_size = win_size; h.cimg = img; h.fix_shift = 0; h.scratch = scratch + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) / HZ_G) * HZ_HINT_SCRATCH;   // per group, not per slot: padding groups redo a slot
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, j.cur, real ? err : nullptr, &h, &pend);
    VC vse, sgn, ay, s1Ay, S, R8x, R8y, Msg;
    vse.v = j.enabled; sgn.v = j.signSignature; ay.v = j.aySignature; s1Ay.v = j.s1Ay;
    S.v = j.S; R8x.v = j.R8x; R8y.v = j.R8y; Msg.v = j.M;
    vse.sig = sgn.sig = ay.sig = s1Ay.sig = S.sig = R8x.sig = R8y.sig = Msg.sig = true;
    hz_phaseF(c, vse, sgn, ay, s1Ay, S, R8x, R8y, Msg);
    c.flush_pending();
}

// the entries of the ladder steps of the signatures k_eddsa walked (hints.cuh): a thread per chunk of steps;
// first the 33 chunks of the two EscalarMulAny segments of every signature, then the 11 of the fixed-base adders
#define HZ_FILL_PER_SIG (HZ_FILL_CHUNKS + 83)
__global__ void __launch_bounds__(128) k_ed_fill(Fr* w, HzErrBuf* err, const Fr* tab, const Fr* scratch, uint32_t win_stride, int ns) {
    const int gid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int nseg = HZ_FILL_C0 + HZ_FILL_C1;
    if (gid >= (nseg + HZ_FILL_CF) * ns) {
        // the 83 window multiplexers of every signature
        const int g3 = gid - (nseg + HZ_FILL_CF) * ns, gw = g3 / 83;
        if (gw >= ns) return;
        const Fr* scw = scratch + (size_t)gw * HZ_HINT_SCRATCH;
        const HzFillRec recw = *reinterpret_cast<const HzFillRec*>(scw + HZ_SCR_REC);
        if (recw.ok) hz_fill_winmux(w, tab, recw, win_stride, g3 % 83);
        return;
    }
    const bool fix = gid >= nseg * ns;
    const int g2 = fix ? gid - nseg * ns : gid;
    const int g = fix ? g2 / HZ_FILL_CF : g2 / nseg, k = fix ? g2 % HZ_FILL_CF : g2 % nseg;
    if (g >= ns) return;
    const Fr* scr = scratch + (size_t)g * HZ_HINT_SCRATCH;          // group g of k_eddsa evaluated slot s0 + g
    const HzFillRec rec = *reinterpret_cast<const HzFillRec*>(scr + HZ_SCR_REC);
    if (!rec.ok) return;                                            // degenerate input: the templates divided on the spot
    if (fix) hz_fill_fix(w, err, scr, scr + HZ_SCR_AFF, tab, rec, k);
    else if (k < HZ_FILL_C0) hz_fill_seg(w, err, scr, scr + HZ_SCR_AFF, rec, 0, k);
    else hz_fill_seg(w, err, scr, scr + HZ_SCR_AFF, rec, 1, k - HZ_FILL_C0);
}

__global__ void __launch_bounds__(HZ_BLOCK) k_tail(HzPlan pl, Fr* w, HzErrBuf* err, HzErrBuf* ebuf, const HzSmtMid* mids, const Fr* isExit, const Fr* isExit_front,
                                                  const HzEddsaJob* jobs, const HzEddsaJob* jobs_front, const HzSmtJob* smt, const HzSmtJob* smt_front,
                                                  int s0, int ns, int f0, int nf) {
    bool real; int q = hz_item(ns + nf, real);
    if (q < ns) {
        int i = s0 + q;
        sec_rollup_tail(w, real ? err : nullptr, pl, i, &mids[2 * i], isExit[i]);
        // the jobs k_edjob / k_smtjob derived ahead of k_front against the ones its templates arrived at
        if (real && hz_glane() == 0) {
            const HzEddsaJob &a = jobs[i], &b = jobs_front[i];
            bool same = fr_eq(a.aySignature, b.aySignature) && fr_eq(a.signSignature, b.signSignature) && fr_eq(a.enabled, b.enabled) &&
                        fr_eq(a.s1Ay, b.s1Ay) && fr_eq(a.S, b.S) && fr_eq(a.R8x, b.R8x) && fr_eq(a.R8y, b.R8y) && fr_eq(a.M, b.M) && a.cur == b.cur;
            for (int k = 0; k < 2; k++) {
                const HzSmtJob &p = smt[2 * i + k], &q2 = smt_front[2 * i + k];
                same = same && fr_eq(p.oldKey, q2.oldKey) && fr_eq(p.oldValue, q2.oldValue) && fr_eq(p.isOld0, q2.isOld0) && fr_eq(p.newKey, q2.newKey) &&
                       fr_eq(p.newValue, q2.newValue) && fr_eq(p.fnc0, q2.fnc0) && fr_eq(p.fnc1, q2.fnc1) && p.cur == q2.cur && p.sib == q2.sib && p.sigmask == q2.sigmask;
            }
            same = same && fr_eq(isExit[i], isExit_front[i]);
            if (!same) atomicAdd(&ebuf->pad, 1u);
        }
    }
    else { int i = f0 + (q - ns); sec_fee_tail(w, real ? err : nullptr, pl, i, mids[2 * pl.P.nTx + i]); }
}

// ---- the exchange between shards (one all-gather, done by the host binding) ---------------------
// What hasherInputs needs from slots this GPU did not evaluate: their SHA pre-image bits
// (L1TxFullData, 624 bits for the first maxL1Tx slots, and L1L2TxData with the amount nullified,
// 2 nLevels + 48 bits) and the three values that close the chains (newLastIdx, newStateRoot,
// newExitRoot).  A rank's record: per own slot 624 + nL2 bytes (one bit per byte), then 3 x 32 B
// values and 3 x 4 B flags saying which of them this rank produced.
__host__ __device__ static inline uint32_t hz_rec_slot_bytes(const HzParams& P) { return 624u + 2u * (uint32_t)P.nLevels + 48u; }
__host__ __device__ static inline uint32_t hz_rec_vals_off(const HzParams& P, int per) { return ((uint32_t)per * hz_rec_slot_bytes(P) + 3u) & ~3u; }
__host__ __device__ static inline uint32_t hz_rec_bytes(const HzParams& P, int per) { return (hz_rec_vals_off(P, per) + 3 * 32 + 3 * 4 + 15u) & ~15u; }
__global__ void k_pack(HzPlan pl, const Fr* w, const uint8_t* pre, uint8_t* rec, int s0, int ns, int per, int own_last_slot, int own_last_fee) {
    const HzParams& P = pl.P;
    const uint32_t sb = hz_rec_slot_bytes(P), nL2 = 2u * (uint32_t)P.nLevels + 48u;
    const uint32_t p1 = hz_pre_pos_l1(P), p2 = hz_pre_pos_l2(P);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (uint32_t e = gid; e < (uint32_t)per * sb; e += gsz) {
        uint32_t q = e / sb, k = e % sb, i = (uint32_t)s0 + q;
        uint8_t v = 0;
        if ((int)q < ns) {
            if (k < 624) v = (int)i < P.maxL1Tx ? pre[p1 + i * 624 + k] : 0;
            else v = pre[p2 + i * nL2 + (k - 624)];
        }
        rec[e] = v;
    }
    if (gid == 0) {
        uint32_t* out = reinterpret_cast<uint32_t*>(rec + hz_rec_vals_off(P, per));
        Fr z = fr_zero();
        const Fr& a = own_last_slot ? w[pl.idx_newLastIdx] : z;
        const Fr& b = own_last_fee ? w[pl.idx_newStateRoot] : z;
        const Fr& c = own_last_slot ? w[pl.idx_newExitRoot] : z;
        for (int k = 0; k < 8; k++) { out[k] = a.l[k]; out[8 + k] = b.l[k]; out[16 + k] = c.l[k]; }
        out[24] = (uint32_t)own_last_slot; out[25] = (uint32_t)own_last_fee; out[26] = (uint32_t)own_last_slot;
    }
}
__global__ void k_unpack(HzPlan pl, Fr* w, uint8_t* pre, const uint8_t* recs, int world, int per, uint32_t rec_bytes) {
    const HzParams& P = pl.P;
    const uint32_t sb = hz_rec_slot_bytes(P), nL2 = 2u * (uint32_t)P.nLevels + 48u;
    const uint32_t p1 = hz_pre_pos_l1(P), p2 = hz_pre_pos_l2(P);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const uint32_t tot = (uint32_t)world * (uint32_t)per * sb;
    for (uint32_t e = gid; e < tot; e += gsz) {
        uint32_t r = e / ((uint32_t)per * sb), x = e % ((uint32_t)per * sb);
        uint32_t q = x / sb, k = x % sb, i = r * (uint32_t)per + q;
        if ((int)i >= P.nTx) continue;
        uint8_t v = recs[(size_t)r * rec_bytes + x];
        if (k < 624) { if ((int)i < P.maxL1Tx) pre[p1 + i * 624 + k] = v; }
        else pre[p2 + i * nL2 + (k - 624)] = v;
    }
    if (gid < 3) {
        const uint32_t idx = gid == 0 ? pl.idx_newLastIdx : (gid == 1 ? pl.idx_newStateRoot : pl.idx_newExitRoot);
        for (int r = 0; r < world; r++) {
            const uint32_t* in = reinterpret_cast<const uint32_t*>(recs + (size_t)r * rec_bytes + hz_rec_vals_off(P, per));
            if (in[24 + gid]) { Fr f; for (int k = 0; k < 8; k++) f.l[k] = in[8 * gid + k]; w[idx] = f; }
        }
    }
}

// ---- hasherInputs (src/hash-inputs.circom:23-185) ------------------------------------------------
// The head of the component (everything before the first sha256compression) is bit plumbing:
// every entry is one bit of a main input / of a root, or a padding / IV constant.  It is laid out
// by sec_hasher_head (the layout walk); here each entry gets its own thread.  Entry offsets from
// the component start, in declaration order:
//   L1L2TxsData amount bits (40 per slot, the only L1L2 bits that are not wires: rollup-main.circom:452-461)
//   n2bOldLastIdx.out[48], n2bNewLastIdx.out[48], n2b{OldState,NewState,NewExit}Root.out[256],
//   n2bFeeTxsData[i].out[48], n2bChainID.out[16], n2bCurrentNumBatch.out[32],
//   inputsHasher.paddedIn[nBits .. 512 nblocks), inputsHasher.ha0..ha7.out[32]
struct HzHeadOff { uint32_t l1l2, n2b, pad, iv, end; };
__host__ __device__ static inline HzHeadOff hz_head_off(const HzPlan& pl) {
    HzHeadOff o;
    o.l1l2 = 0;
    o.n2b = (uint32_t)pl.P.nTx * 40;
    o.pad = o.n2b + 48 + 48 + 3 * 256 + (uint32_t)pl.P.maxFeeTx * 48 + 16 + 32;
    o.iv = o.pad + (pl.nblocks * 512 - pl.nbits);
    o.end = o.iv + 256;
    return o;
}
__device__ __forceinline__ void hz_put_bit(Fr* p, uint32_t bit) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(bit, 0, 0, 0));
    __stcs(reinterpret_cast<uint4*>(p) + 1, make_uint4(0, 0, 0, 0));
}
__device__ __forceinline__ void hz_head_err(HzErrBuf* err, uint32_t cur, uint32_t line, const Fr& lhs) {
    if (!err) return;
    uint32_t k = atomicAdd(&err->count, 1u);
    if (k < HZ_MAX_ERR) { err->rec[k].cur = cur; err->rec[k].line = line; err->rec[k].lhs = lhs; err->rec[k].rhs = fr_zero(); }
}
// one Num2Bits(n) of `value`: entries at w[e0 .. e0+n); the pre-image takes its low m bits MSB first
// (pre[p0 + k] = bit m-1-k); bits at or above `chk` must be zero (:54-99), and value < 2^n (Num2Bits)
__device__ __forceinline__ void hz_head_n2b(Fr* w, uint8_t* pre, HzErrBuf* err, uint32_t tid, uint32_t e0, int n,
                                           const Fr& value, uint32_t p0, int m, int chk) {
    if (tid < (uint32_t)n) {
        uint32_t b = fr_bit(value, (int)tid);
        hz_put_bit(&w[e0 + tid], b);
        if ((int)tid < m) pre[p0 + (uint32_t)(m - 1) - tid] = (uint8_t)b;
        if ((int)tid >= chk && b) hz_head_err(err, e0 + (uint32_t)n, __LINE__, fr_one());
    }
    if (tid == 0 && n < 254) {
        bool ok = true;
        for (int i = n; i < 254; i++) ok = ok && !fr_bit(value, i);
        if (!ok) hz_head_err(err, e0 + (uint32_t)n, __LINE__, value);
    }
}
__global__ void k_hash_head(HzPlan pl, Fr* w, uint8_t* pre, HzErrBuf* err, uint32_t* off_blocks) {
    const HzParams& P = pl.P;
    const HzHeadOff o = hz_head_off(pl);
    const uint32_t nL = (uint32_t)P.nLevels, nL2 = 2 * nL + 48, base = pl.off_hasher;
    const uint32_t p2 = hz_pre_pos_l2(P), total = pl.nblocks * 512;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (gid == 0) { *off_blocks = base + o.end; hz_put_bit(&w[0], 1); }
    // amount bits of L1L2TxsData as k_front left them in the pre-image (already nullified)
    for (uint32_t e = gid; e < (uint32_t)P.nTx * 40; e += gsz) {
        uint32_t i = e / 40, jj = e % 40;
        hz_put_bit(&w[base + o.l1l2 + e], pre[p2 + i * nL2 + 2 * nL + jj]);
    }
    // the Num2Bits blocks: one 256-thread tile each
    const uint32_t nN2b = 2 + 3 + (uint32_t)P.maxFeeTx + 2;
    for (uint32_t q = blockIdx.x; q < nN2b; q += gridDim.x) {
        uint32_t tid = threadIdx.x, e0 = base + o.n2b;
        if (q == 0) hz_head_n2b(w, pre, err, tid, e0, 48, w[pl.in_off[IN_oldLastIdx]], 0, 48, (int)nL);
        else if (q == 1) hz_head_n2b(w, pre, err, tid, e0 + 48, 48, w[pl.idx_newLastIdx], 48, 48, (int)nL);
        else if (q < 5) {
            uint32_t r = q - 2;
            uint32_t src = r == 0 ? pl.in_off[IN_oldStateRoot] : (r == 1 ? pl.idx_newStateRoot : pl.idx_newExitRoot);
            hz_head_n2b(w, pre, err, tid, e0 + 96 + 256 * r, 256, w[src], 96 + 256 * r, 256, 256);
        } else if (q < 5 + (uint32_t)P.maxFeeTx) {
            uint32_t i = q - 5;
            hz_head_n2b(w, pre, err, tid, e0 + 96 + 768 + 48 * i, 48, w[pl.in_off[IN_feeIdxs] + i],
                        p2 + (uint32_t)P.nTx * nL2 + i * nL, (int)nL, (int)nL);
        } else if (q == 5 + (uint32_t)P.maxFeeTx) {
            hz_head_n2b(w, pre, err, tid, e0 + 96 + 768 + 48 * (uint32_t)P.maxFeeTx, 16, w[pl.in_off[IN_globalChainID]],
                        p2 + (uint32_t)P.nTx * nL2 + (uint32_t)P.maxFeeTx * nL, 16, 16);
        } else {
            hz_head_n2b(w, pre, err, tid, e0 + 96 + 768 + 48 * (uint32_t)P.maxFeeTx + 16, 32, w[pl.in_off[IN_currentNumBatch]],
                        p2 + (uint32_t)P.nTx * nL2 + (uint32_t)P.maxFeeTx * nL + 16, 32, 32);
        }
    }
    // padding: 1, zeros, 64-bit big-endian length
    for (uint32_t k = pl.nbits + gid; k < total; k += gsz) {
        uint32_t b;
        if (k == pl.nbits) b = 1;
        else if (k < total - 64) b = 0;
        else { uint32_t kk = total - k - 1; b = kk < 32 ? ((pl.nbits >> kk) & 1u) : 0; }
        hz_put_bit(&w[base + o.pad + (k - pl.nbits)], b);
        pre[k] = (uint8_t)b;
    }
    // initial hash value, LSB-first bit signals
    for (uint32_t e = gid; e < 256; e += gsz) hz_put_bit(&w[base + o.iv + e], (hz_sha_h((int)(e >> 5)) >> (e & 31)) & 1u);
}

__constant__ uint32_t HZ_SHA_K_DEV[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
__device__ __forceinline__ uint32_t hz_rr(uint32_t x, int r) { return __funnelshift_r(x, x, r); }
// one round on v = (a, b, c, d, e, f, g, h)
__device__ __forceinline__ void hz_sha_round(uint32_t* v, uint32_t wt, int t) {
    uint32_t S1 = hz_rr(v[4], 6) ^ hz_rr(v[4], 11) ^ hz_rr(v[4], 25);
    uint32_t ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
    uint32_t t1 = v[7] + S1 + ch + HZ_SHA_K_DEV[t] + wt;
    uint32_t S0 = hz_rr(v[0], 2) ^ hz_rr(v[0], 13) ^ hz_rr(v[0], 22);
    uint32_t mj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
    uint32_t t2 = S0 + mj;
    v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = v[3] + t1; v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = t1 + t2;
}
__device__ __forceinline__ uint32_t hz_sha_sched(uint32_t w2, uint32_t w7, uint32_t w15, uint32_t w16) {
    uint32_t s0 = hz_rr(w15, 7) ^ hz_rr(w15, 18) ^ (w15 >> 3);
    uint32_t s1 = hz_rr(w2, 17) ^ hz_rr(w2, 19) ^ (w2 >> 10);
    return w16 + s0 + w7 + s1;
}

// chaining values of all blocks.  The chain is serial (block b needs the digest of block b-1), so one thread
// walks it -- but only the 64 rounds: the message schedule does not depend on the chaining value, so all
// threads of the block first expand it for every SHA block (W[t] + K[t], 64 words per block), and the walking
// thread reads the next block's words while it works on the current one.
__device__ __forceinline__ void hz_sha_round_wk(uint32_t* v, uint32_t wk) {
    uint32_t S1 = hz_rr(v[4], 6) ^ hz_rr(v[4], 11) ^ hz_rr(v[4], 25);
    uint32_t ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
    uint32_t t1 = (v[7] + wk) + (S1 + ch);
    uint32_t S0 = hz_rr(v[0], 2) ^ hz_rr(v[0], 13) ^ hz_rr(v[0], 22);
    uint32_t mj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
    uint32_t t2 = S0 + mj;
    v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = v[3] + t1; v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = t1 + t2;
}
// Launched twice per batch, [0, nblocks / 2) (with the expansion of every block's schedule) and the rest, so that
// the bit signals of the first half's blocks (k_sha_blocks) are written while the second half is walked.
__global__ void __launch_bounds__(256) k_sha_chain(HzPlan pl, const uint8_t* preb, uint32_t* chain, uint32_t* msgw, uint32_t* wk,
                                                   uint32_t b_begin, uint32_t b_end) {
    if (b_begin == 0) {
    uint32_t nw = pl.nblocks * 16;
    for (uint32_t t = threadIdx.x; t < nw; t += blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(preb + (size_t)t * 32);
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            uint4 x = p[q];
            uint32_t ws[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int b = 0; b < 4; b++) v = (v << 1) | ((ws[k] >> (8 * b)) & 1u);
        }
        msgw[t] = v;
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < pl.nblocks; b += blockDim.x) {
        uint32_t wv[16];
#pragma unroll
        for (int t = 0; t < 16; t++) wv[t] = msgw[(size_t)b * 16 + t];
#pragma unroll
        for (int t = 0; t < 64; t++) {
            if (t >= 16) wv[t & 15] = hz_sha_sched(wv[(t - 2) & 15], wv[(t - 7) & 15], wv[(t - 15) & 15], wv[t & 15]);
            wk[(size_t)b * 64 + t] = wv[t & 15] + HZ_SHA_K_DEV[t];
        }
    }
    __syncthreads();
    }
    if (threadIdx.x != 0) return;
    uint32_t st[8];
    if (b_begin == 0) { for (int j = 0; j < 8; j++) { st[j] = hz_sha_h(j); chain[j] = st[j]; } }
    else { for (int j = 0; j < 8; j++) st[j] = chain[(size_t)b_begin * 8 + j]; }
    const uint4* wk4 = reinterpret_cast<const uint4*>(wk);
    uint4 nxt[16];
#pragma unroll
    for (int q = 0; q < 16; q++) nxt[q] = wk4[(size_t)b_begin * 16 + q];
    for (uint32_t b = b_begin; b < b_end; b++) {
        uint4 cur[16];
#pragma unroll
        for (int q = 0; q < 16; q++) cur[q] = nxt[q];
        const uint32_t bn = b + 1 < pl.nblocks ? b + 1 : b;
#pragma unroll
        for (int q = 0; q < 16; q++) nxt[q] = wk4[(size_t)bn * 16 + q];
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = st[j];
#pragma unroll
        for (int q = 0; q < 16; q++) {
            hz_sha_round_wk(v, cur[q].x); hz_sha_round_wk(v, cur[q].y); hz_sha_round_wk(v, cur[q].z); hz_sha_round_wk(v, cur[q].w);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) { st[j] += v[j]; chain[(size_t)(b + 1) * 8 + j] = st[j]; }
    }
}

// all bit signals of the sha256compression components: one CUDA block per SHA block, one thread per
// unit (templates_circomlib.cuh: 64 rounds, 48 message-schedule words, digest + final sums).  A
// thread recomputes the word values it needs from the block's chaining value and message words
// (a few hundred integer instructions) and then emits the unit's ~450 bit signals.
#define HZ_SHA_THREADS 128
__global__ void __launch_bounds__(HZ_SHA_THREADS) k_sha_blocks(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain, const uint32_t* msgw, const uint32_t* off_blocks, uint32_t b0) {
    const uint32_t b = b0 + blockIdx.x;
    const int u = (int)threadIdx.x;
    const uint32_t base = *off_blocks + b * (uint32_t)HZ_SZ_SHA_BLOCK;
    uint32_t wv[64], hin[8];
#pragma unroll
    for (int t = 0; t < 16; t++) wv[t] = msgw[(size_t)b * 16 + t];
#pragma unroll
    for (int t = 16; t < 64; t++) wv[t] = hz_sha_sched(wv[t - 2], wv[t - 7], wv[t - 15], wv[t - 16]);
#pragma unroll
    for (int j = 0; j < 8; j++) hin[j] = chain[(size_t)b * 8 + j];
    Comp<MCompute> c(w, base, nullptr);
    c.solo = true;                          // one thread per unit: it stores every entry of its unit
    if (u < 64) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = hin[j];
        for (int t = 0; t < u; t++) hz_sha_round(v, wv[t], t);
        W32<MCompute> vv[8], wt, na, ne;
        for (int j = 0; j < 8; j++) vv[j].v = v[j];
        wt.v = wv[u];
        Sha256UnitRound<MCompute>(c, base, u, vv, wt, na, ne);
    } else if (u < 64 + 48) {
        int t = u - 64 + 16;
        W32<MCompute> a, b7, c15, d16;
        a.v = wv[t - 2]; b7.v = wv[t - 7]; c15.v = wv[t - 15]; d16.v = wv[t - 16];
        Sha256UnitSigmaPlus<MCompute>(c, base, t, a, b7, c15, d16);
    } else if (u == 64 + 48) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = hin[j];
        for (int t = 0; t < 64; t++) hz_sha_round(v, wv[t], t);
        for (int j = 0; j < 8; j++) hz_sha_out_word(c, j, hin[j] + v[j]);
        for (int j = 0; j < 8; j++) {
            W32<MCompute> h, f;
            h.v = hin[j]; f.v = v[j];
            W32<MCompute> fs = Sha256UnitFinal<MCompute>(c, base, j, h, f);
            if (err && fs.v != chain[(size_t)(b + 1) * 8 + j]) hz_head_err(err, base, __LINE__, fr_from_u64(fs.v));
        }
    }
}

__global__ void k_hash_tail(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain) {
    W32<MCompute> dig[8];
    for (int j = 0; j < 8; j++) dig[j].v = chain[(size_t)pl.nblocks * 8 + j];
    Comp<MCompute> tc(w, pl.off_hasher_tail, threadIdx.x < HZ_G ? err : nullptr);
    Comp<MCompute> shc(w, pl.off_hasher_tail, nullptr);
    VC h = sec_hasher_tail(tc, shc, dig);
    tc.put(1, h.v);
}

HzCuda* hz_cuda_create(int device, const HzPlan& ph, const uint32_t* off_decode, const uint32_t* off_rollup,
                       const uint32_t* off_fee, const uint32_t* off_eddsa, const uint32_t* off_jk, int* status) {
    *status = HZW_E_DEVICE;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        fprintf(stderr, "hzw: no CUDA device -- this library has no CPU path\n");
        return nullptr;
    }
    if (device < 0 || device >= n) { *status = HZW_E_PARAM; return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    HzCuda* d = new HzCuda();
    memset(d, 0, sizeof *d);
    d->device = device;
    d->plan = ph;
    hz_set_ranges(d, 0, 1);
    const HzParams& P = ph.P;
    bool ok = true;
    if (ph.off_hasher + hz_head_off(ph).end != ph.off_sha_blocks) {
        fprintf(stderr, "hzw: hasherInputs head layout %u does not match the walk %u\n", ph.off_hasher + hz_head_off(ph).end, ph.off_sha_blocks);
        delete d; *status = HZW_E_PARAM; return nullptr;
    }
    ok = ok && cudaMalloc(&d->w, 32ull * ph.nvars) == cudaSuccess;
    ok = ok && cudaMalloc(&d->pre, (size_t)ph.nblocks * 512) == cudaSuccess;
    ok = ok && cudaMalloc(&d->chain, (size_t)(ph.nblocks + 1) * 8 * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&d->msgw, (size_t)ph.nblocks * 16 * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&d->shawk, (size_t)ph.nblocks * 64 * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&d->jobs, sizeof(HzEddsaJob) * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->jobs_front, sizeof(HzEddsaJob) * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_eddsa, 4ull * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_jk, 4ull * P.nTx) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_jk, off_jk, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMalloc(&d->smt_front, sizeof(HzSmtJob) * 2 * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->isExit_front, sizeof(Fr) * P.nTx) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_eddsa, off_eddsa, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMalloc(&d->smt, sizeof(HzSmtJob) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->mids, sizeof(HzSmtMid) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->lev, sizeof(HzLevJob) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->smst, sizeof(Fr) * 6 * (size_t)(P.nLevels + 1) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->isExit, sizeof(Fr) * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->img00, sizeof(Fr) * (HZ_IMG00_DIGEST + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->zflag, (size_t)((2 * P.nTx + P.maxFeeTx + 15) / 16 + 1) * (P.nLevels + 1)) == cudaSuccess;
    {
        const size_t nj = (size_t)2 * P.nTx + P.maxFeeTx, gs = HZ_GL_STRIDE(P.nLevels + 1);
        ok = ok && cudaMalloc(&d->holes, sizeof(HzHole) * nj) == cudaSuccess && cudaMallocHost(&d->h_holes, sizeof(HzHole) * nj) == cudaSuccess;
        ok = ok && cudaMalloc(&d->gl, sizeof(Fr) * gs * nj) == cudaSuccess && cudaMallocHost(&d->h_gl, sizeof(Fr) * gs * nj) == cudaSuccess;
        ok = ok && cudaMallocHost(&d->h_img00, sizeof(Fr) * HZ_HS) == cudaSuccess;
        ok = ok && cudaStreamCreate(&d->stream5) == cudaSuccess && cudaStreamCreate(&d->stream6) == cudaSuccess;
        for (int k = 0; k < 2; k++) ok = ok && cudaEventCreateWithFlags(&d->ev_pack[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_info[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_sha[k], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&d->hint_cst, sizeof(Fr) * HZ_CST_MAX) == cudaSuccess;
    ok = ok && cudaMalloc(&d->hint_tab, sizeof(Fr) * HZ_EDTAB_LEN) == cudaSuccess;
    ok = ok && cudaMalloc(&d->hint_scratch, sizeof(Fr) * HZ_HINT_SCRATCH * (size_t)hz_grid(P.nTx) * (HZ_BLOCK / HZ_G)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->err, sizeof(HzErrBuf)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_decode, 4ull * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_rollup, 4ull * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_fee, 4ull * P.maxFeeTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_blocks, 4) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_decode, off_decode, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_rollup, off_rollup, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_fee, off_fee, 4ull * P.maxFeeTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaStreamCreate(&d->stream) == cudaSuccess && cudaStreamCreate(&d->stream2) == cudaSuccess;
    ok = ok && cudaStreamCreate(&d->stream3) == cudaSuccess && cudaStreamCreate(&d->stream4) == cudaSuccess;
    for (int k = 0; k < 2; k++) ok = ok && cudaEventCreateWithFlags(&d->ev_half[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_ed[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming) == cudaSuccess;
    d->no_holes = getenv("HZW_NO_HOLES") != nullptr;
    d->h_off_rollup = new std::vector<uint32_t>(off_rollup, off_rollup + P.nTx);
    d->h_off_rollup->push_back(off_fee[0]);
    d->h_off_decode0 = off_decode[0]; d->h_off_rollup0 = off_rollup[0]; d->h_off_rollup_mid = off_rollup[P.nTx / 2]; d->h_off_fee0 = off_fee[0];
    ok = ok && cudaEventCreate(&d->ev0) == cudaSuccess && cudaEventCreate(&d->ev1) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_front, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_eddsa, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_lev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_fee, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_zero, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < HZ_NKERN + 3; k++) ok = ok && cudaEventCreate(&d->kev[k]) == cudaSuccess;
    // kernels that are to share an SM must agree on its L1 / shared-memory split: k_smt_levels stages the MDS
    // exchange in shared memory, k_eddsa (running beside it on the second stream) uses none -- without a common
    // carve-out the second kernel only gets SMs the first has left
    {
        const int carve = 15;    // percent of the 228 KB: room for four k_smt_levels blocks
        cudaFuncSetAttribute(k_front, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_fee_front, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_edjob, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_smtjob, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_eddsa, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_smt, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_smt_levels, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_tail, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    // the per-thread frames of the circuit templates are large (sibling arrays, state machines)
    ok = ok && cudaDeviceSetLimit(cudaLimitStackSize, 80 * 1024) == cudaSuccess;
    if (ok) {
        // the entries and the digest of SMTHash2(0, 0), once (k_poseidon_t3 on one permutation of two zero inputs)
        Fr* zin = nullptr;
        ok = cudaMalloc(&zin, sizeof(Fr) * 2) == cudaSuccess && cudaMemset(zin, 0, sizeof(Fr) * 2) == cudaSuccess;
        if (ok) {
            k_poseidon_t3<<<1, 96, 0, d->stream>>>(zin, d->img00, d->img00 + HZ_IMG00_DIGEST, 1);
            ok = cudaStreamSynchronize(d->stream) == cudaSuccess;
            ok = ok && cudaMemcpy(d->h_img00, d->img00, sizeof(Fr) * HZ_HS, cudaMemcpyDeviceToHost) == cudaSuccess;
            if (ok && (ph.sz.phaseJK == 0 || HZ_HS != 1u + 2u * 3u * 65u + 3u * (8u * 3u + 57u))) ok = false;
        }
        cudaFree(zin);
    }
    if (ok) {
        // record the input-independent part of the EdDSA fixed-base multiplication once
        uint32_t* counts = nullptr; uint32_t hc[5] = {0, 0, 0, 0, 0};
        ok = cudaMalloc(&d->hint_img, sizeof(Fr) * ((size_t)ph.sz.sigVerifier + 64)) == cudaSuccess && cudaMalloc(&counts, 20) == cudaSuccess;
        Fr* wtmp = d->hint_img;
        if (ok) {
            k_fix_consts<<<1, 32, 0, d->stream>>>(wtmp, d->hint_cst, d->hint_tab, counts);
            ok = cudaMemcpyAsync(hc, counts, 20, cudaMemcpyDeviceToHost, d->stream) == cudaSuccess &&
                 cudaStreamSynchronize(d->stream) == cudaSuccess;
            if (ok && (hc[0] > HZ_CST_MAX || hc[1] != HZ_EDTAB_LEN)) {
                fprintf(stderr, "hzw: hint table sizes %u / %u do not match the templates\n", hc[0], hc[1]);
                ok = false;
            }
            d->fix_size = hc[4]; d->win_size = hc[3]; d->win_stride = hc[3] / 83;
            if (ok && (hc[3] == 0 || hc[3] % 83 != 0 || hc[4] > ph.sz.sigVerifier || ph.sz.mulfix_off + hc[4] > ph.sz.getAx + ph.sz.sigVerifier)) {
                fprintf(stderr, "hzw: EscalarMulFix image (%u entries, windows %u) does not fit the layout\n", hc[4], hc[3]);
                ok = false;
            }
        }
        cudaFree(counts);
    }
    if (!ok) {
        fprintf(stderr, "hzw: CUDA allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        hz_cuda_destroy(d);
        return nullptr;
    }
    d->plan.off_decode = d->off_decode; d->plan.off_rollup = d->off_rollup; d->plan.off_fee = d->off_fee;
    *status = 0;
    return d;
}

void hz_cuda_destroy(HzCuda* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    cudaFree(d->w); cudaFree(d->pre); cudaFree(d->chain); cudaFree(d->msgw); cudaFree(d->shawk); cudaFree(d->jobs); cudaFree(d->jobs_front); cudaFree(d->off_eddsa); cudaFree(d->off_jk); cudaFree(d->smt_front); cudaFree(d->isExit_front); cudaFree(d->err);
    cudaFree(d->smt); cudaFree(d->mids); cudaFree(d->isExit); cudaFree(d->lev); cudaFree(d->smst); cudaFree(d->img00); cudaFree(d->zflag); cudaFree(d->holes); cudaFree(d->gl); cudaFreeHost(d->h_holes); cudaFreeHost(d->h_gl); cudaFreeHost(d->h_img00);
    cudaFree(d->hint_cst); cudaFree(d->hint_tab); cudaFree(d->hint_img); cudaFree(d->hint_scratch); cudaFree(d->recs);
    cudaFree(d->off_decode); cudaFree(d->off_rollup); cudaFree(d->off_fee); cudaFree(d->off_blocks);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->stream2) cudaStreamDestroy(d->stream2);
    if (d->stream3) cudaStreamDestroy(d->stream3);
    if (d->stream4) cudaStreamDestroy(d->stream4);
    if (d->stream5) cudaStreamDestroy(d->stream5);
    if (d->stream6) cudaStreamDestroy(d->stream6);
    for (int k = 0; k < 2; k++) { if (d->ev_pack[k]) cudaEventDestroy(d->ev_pack[k]); if (d->ev_info[k]) cudaEventDestroy(d->ev_info[k]); if (d->ev_sha[k]) cudaEventDestroy(d->ev_sha[k]); }
    for (int k = 0; k < 2; k++) { if (d->ev_half[k]) cudaEventDestroy(d->ev_half[k]); if (d->ev_ed[k]) cudaEventDestroy(d->ev_ed[k]); }
    if (d->ev_done) cudaEventDestroy(d->ev_done);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev_front) cudaEventDestroy(d->ev_front);
    if (d->ev_eddsa) cudaEventDestroy(d->ev_eddsa);
    if (d->ev_lev) cudaEventDestroy(d->ev_lev);
    if (d->ev_fee) cudaEventDestroy(d->ev_fee);
    if (d->ev_zero) cudaEventDestroy(d->ev_zero);
    for (int k = 0; k < HZ_NKERN + 3; k++) if (d->kev[k]) cudaEventDestroy(d->kev[k]);
    delete d->h_off_rollup;
    delete d;
}

static void hz_set_ranges(HzCuda* d, int rank, int world) {
    const HzParams& P = d->plan.P;
    d->rank = rank; d->world = world;
    d->per = (P.nTx + world - 1) / world;
    d->s0 = rank * d->per < P.nTx ? rank * d->per : P.nTx;
    d->ns = (d->s0 + d->per <= P.nTx ? d->per : P.nTx - d->s0);
    int perf = (P.maxFeeTx + world - 1) / world;
    d->f0 = rank * perf < P.maxFeeTx ? rank * perf : P.maxFeeTx;
    d->nf = (d->f0 + perf <= P.maxFeeTx ? perf : P.maxFeeTx - d->f0);
    uint32_t perb = (d->plan.nblocks + (uint32_t)world - 1) / (uint32_t)world;
    d->b0 = (uint32_t)rank * perb < d->plan.nblocks ? (uint32_t)rank * perb : d->plan.nblocks;
    d->nb = d->b0 + perb <= d->plan.nblocks ? perb : d->plan.nblocks - d->b0;
}
int hz_cuda_set_shard(HzCuda* d, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return HZW_E_PARAM;
    CK(cudaSetDevice(d->device));
    hz_set_ranges(d, rank, world);
    if (d->recs) { cudaFree(d->recs); d->recs = nullptr; }
    if (world > 1) CK(cudaMalloc(&d->recs, (size_t)world * hz_rec_bytes(d->plan.P, d->per)));
    return 0;
}
uint64_t hz_cuda_exchange_bytes(const HzCuda* d) { return hz_rec_bytes(d->plan.P, d->per); }

// phase 1: inputs to the device, then everything per slot / fee tx of the shard
static int hz_phase1(HzCuda* d, const uint8_t* inputs, int sanity) {
    const HzPlan& pl = d->plan;
    cudaStream_t s = d->stream, s2 = getenv("HZW_SERIAL_STREAMS") ? d->stream : d->stream2;    // (measurement switch)
    HzErrBuf* err = sanity ? d->err : nullptr;
    CK(cudaMemsetAsync(d->err, 0, sizeof(HzErrBuf), s));
    CK(cudaMemsetAsync(d->pre, 0, (size_t)pl.nblocks * 512, s));
    if (inputs) CK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d->w) + 64, inputs, 32ull * pl.n_inputs, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(d->ev0, s));
    const int s0 = d->s0, ns = d->ns, f0 = d->f0, nf = d->nf;
    // the signatures are the longest chain of a slot and need little from the rest of it: their jobs come
    // straight from the inputs (k_edjob) and they run on the second stream from the start, beside k_front
    // and the merkle levels -- no kernel fills the machine alone
    CK(cudaStreamWaitEvent(s2, d->ev0, 0));
    CK(cudaEventRecord(d->kev[9], s2));
    if (ns > 0) {
        k_edjob<<<hz_grid(ns), HZ_BLOCK, 0, s2>>>(pl, d->w, d->jobs, d->off_eddsa, s0, ns);
        k_ed_const<<<592, 256, 0, s2>>>(pl, d->w, d->off_eddsa, d->hint_img, d->fix_size, s0, ns);
            k_eddsa<<<hz_grid(ns), HZ_BLOCK, 0, s2>>>(pl, d->w, err, d->jobs, d->hint_cst, d->hint_tab, d->hint_img, d->hint_scratch, d->win_size, s0, ns);
        k_ed_fill<<<(unsigned)((HZ_FILL_PER_SIG * ns + 127) / 128), 128, 0, s2>>>(d->w, err, d->hint_tab, d->hint_scratch, d->win_stride, ns);
    }
    CK(cudaEventRecord(d->kev[10], s2));
    CK(cudaEventRecord(d->ev_eddsa, s2));
    // the fee transactions need nothing but inputs either: third stream, joined before the merkle kernels
    CK(cudaStreamWaitEvent(d->stream3, d->ev0, 0));
    if (nf > 0) k_fee_front<<<hz_grid(nf), HZ_BLOCK, 0, d->stream3>>>(pl, d->w, err, d->smt + 2 * pl.P.nTx, f0, nf);
    CK(cudaEventRecord(d->ev_fee, d->stream3));
    // k_front on a stream of its own: the merkle kernels take their jobs from k_smtjob
    CK(cudaStreamWaitEvent(d->stream4, d->ev0, 0));
    CK(cudaEventRecord(d->kev[11], d->stream4));
    if (ns > 0) k_front<<<hz_grid(ns), HZ_BLOCK, 0, d->stream4>>>(pl, d->w, d->pre, err, d->jobs_front, d->smt_front, d->isExit_front, s0, ns);
    CK(cudaEventRecord(d->kev[12], d->stream4));
    CK(cudaEventRecord(d->ev_front, d->stream4));
    CK(cudaEventRecord(d->kev[0], s));
    if (ns > 0) k_smtjob<<<hz_grid(ns), HZ_BLOCK, 0, s>>>(pl, d->w, d->smt, d->isExit, d->off_jk, s0, ns);
    CK(cudaEventRecord(d->kev[1], s));
    CK(cudaStreamWaitEvent(s, d->ev_fee, 0));
    CK(cudaEventRecord(d->kev[2], s));
    if (2 * ns + nf > 0) {
        k_smt<<<hz_grid(2 * ns + nf), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->smt, d->mids, d->lev, d->smst, s0, ns, f0, nf);
        const unsigned lb = (unsigned)((2 * ns + nf + 15) / 16);
        k_smt_levels<<<lb, 96, 0, s>>>(pl, d->w, d->lev, d->smst, d->mids, d->img00, d->zflag, s0, ns, f0, nf);
        // the skipped (0, 0) levels are filled in beside the tail and the hash of the inputs
        CK(cudaEventRecord(d->ev_lev, s));
        CK(cudaStreamWaitEvent(d->stream3, d->ev_lev, 0));
        k_smt_zero<<<dim3(lb, (unsigned)(pl.P.nLevels + 1)), 256, 0, d->stream3>>>(pl, d->w, d->lev, d->img00, d->zflag, s0, ns, f0, nf);
        if (d->world > 1 && !d->no_holes) {
            // a shard's pieces are copied around the image levels too (hz_cuda_shard_phase2): their records go to the host now
            const uint32_t gs = HZ_GL_STRIDE(pl.P.nLevels + 1);
            k_hole_pack<<<(unsigned)(2 * ns + nf), 128, 0, s>>>(pl, d->w, d->lev, d->zflag, d->holes, d->gl, s0, ns, f0, nf);
            CK(cudaEventRecord(d->ev_pack[0], s));
            CK(cudaStreamWaitEvent(d->stream5, d->ev_pack[0], 0));
            if (ns > 0) {
                CK(cudaMemcpyAsync(d->h_holes + 2 * s0, d->holes + 2 * s0, sizeof(HzHole) * 2 * ns, cudaMemcpyDeviceToHost, d->stream5));
                CK(cudaMemcpyAsync(d->h_gl + (size_t)2 * s0 * gs, d->gl + (size_t)2 * s0 * gs, sizeof(Fr) * gs * 2 * ns, cudaMemcpyDeviceToHost, d->stream5));
            }
            if (nf > 0) {
                const size_t j = (size_t)2 * pl.P.nTx + f0;
                CK(cudaMemcpyAsync(d->h_holes + j, d->holes + j, sizeof(HzHole) * nf, cudaMemcpyDeviceToHost, d->stream5));
                CK(cudaMemcpyAsync(d->h_gl + j * gs, d->gl + j * gs, sizeof(Fr) * gs * nf, cudaMemcpyDeviceToHost, d->stream5));
            }
            CK(cudaEventRecord(d->ev_info[0], d->stream5));
        }
    }
    CK(cudaEventRecord(d->ev_zero, d->stream3));
    CK(cudaEventRecord(d->kev[3], s));
    CK(cudaStreamWaitEvent(s, d->ev_front, 0));       // the tail compares the jobs, the hash of the inputs reads k_front's pre-image bits
    if (ns + nf > 0) k_tail<<<hz_grid(ns + nf), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->err, d->mids, d->isExit, d->isExit_front, d->jobs, d->jobs_front, d->smt, d->smt_front, s0, ns, f0, nf);
    CK(cudaEventRecord(d->kev[4], s));
    return 0;
}
// phase 2: hasherInputs; every shard walks the whole SHA chain (it is serial and short) and expands
// its own range of blocks
static int hz_phase2(HzCuda* d, int sanity) {
    const HzPlan& pl = d->plan;
    cudaStream_t s = d->stream;
    HzErrBuf* err = sanity ? d->err : nullptr;
    k_hash_head<<<148, 256, 0, s>>>(pl, d->w, d->pre, d->rank == 0 ? err : nullptr, d->off_blocks);
    CK(cudaEventRecord(d->kev[5], s));
    // the chain in two halves: the blocks of the first half are expanded (stream 5) while the second half is walked
    const uint32_t mid = pl.nblocks / 2, bend = d->b0 + d->nb;
    k_sha_chain<<<1, 256, 0, s>>>(pl, d->pre, d->chain, d->msgw, d->shawk, 0, mid);
    const uint32_t n1 = d->b0 < mid ? (bend < mid ? bend : mid) - d->b0 : 0;       // own blocks below mid
    if (n1 > 0) {
        CK(cudaEventRecord(d->ev_sha[0], s));
        CK(cudaStreamWaitEvent(d->stream5, d->ev_sha[0], 0));
        k_sha_blocks<<<n1, HZ_SHA_THREADS, 0, d->stream5>>>(pl, d->w, err, d->chain, d->msgw, d->off_blocks, d->b0);
        CK(cudaEventRecord(d->ev_sha[1], d->stream5));
    }
    k_sha_chain<<<1, 256, 0, s>>>(pl, d->pre, d->chain, d->msgw, d->shawk, mid, pl.nblocks);
    CK(cudaEventRecord(d->kev[6], s));
    if (d->nb > n1) k_sha_blocks<<<d->nb - n1, HZ_SHA_THREADS, 0, s>>>(pl, d->w, err, d->chain, d->msgw, d->off_blocks, d->b0 + n1);
    if (n1 > 0) CK(cudaStreamWaitEvent(s, d->ev_sha[1], 0));
    CK(cudaEventRecord(d->kev[7], s));
    k_hash_tail<<<1, 32, 0, s>>>(pl, d->w, d->rank == 0 ? err : nullptr, d->chain);
    CK(cudaEventRecord(d->kev[8], s));
    CK(cudaStreamWaitEvent(s, d->ev_eddsa, 0));
    CK(cudaStreamWaitEvent(s, d->ev_zero, 0));
    CK(cudaGetLastError());
    CK(cudaEventRecord(d->ev1, s));
    return 0;
}

// One batch end to end with the witness copied to the host WHILE it is computed.  The 8 GB copy over
// PCIe (~150 ms) dwarfs the kernels, so pieces leave as soon as they are final: header, inputs and the
// decodeTx ranges after k_front; the rollupTx ranges of the first half of the slots while the second
// half is evaluated; the rest (second half, fee txs, hasherInputs) at the end.
void hz_host_copy_stream(uint8_t* dst, const uint8_t* src, size_t n);   // hzw_host.cpp: non-temporal stores
void hz_host_copy_fence();
// host side of the holes: the images and the glue entries of jobs [j0, j1) written into the caller's buffer
static void hz_fill_holes(const HzCuda* d, uint8_t* out, int j0, int j1, int part, int parts) {
    const uint32_t gs = HZ_GL_STRIDE(d->plan.P.nLevels + 1);
    for (int job = j0 + part; job < j1; job += parts) {
        const HzHole h = d->h_holes[job];
        if (h.nz == 0) continue;
        const uint8_t* g = reinterpret_cast<const uint8_t*>(d->h_gl + (size_t)job * gs);
        memcpy(out + 32ull * h.cur, g, 8 * 32);
        for (uint32_t k = 0; k < h.nz; k++) {
            const uint32_t oph = hz_lev_oph(h.cur, k);
            hz_host_copy_stream(out + 32ull * oph, reinterpret_cast<const uint8_t*>(d->h_img00), 32ull * HZ_HS);
            hz_host_copy_stream(out + 32ull * (oph + HZ_HS), reinterpret_cast<const uint8_t*>(d->h_img00), 32ull * HZ_HS);
            // switchers of level k and, unless it is the last image level, the header of level k + 1
            memcpy(out + 32ull * (oph + 2 * HZ_HS), g + 32ull * (8 + 14 * k), (k + 1 < h.nz ? 14 : 8) * 32);
        }
    }
    hz_host_copy_fence();
}
// device-to-host copies of witness ranges around the holes (the image levels the host rebuilds itself).  `out` is
// where witness entry 0 would be: the end-to-end path passes the caller's buffer, the sharded path a pointer
// shifted so that the piece being copied lands at its place in the shard buffer.
struct HzCopier {
    HzCuda* d; uint8_t* out; const uint8_t* wb; unsigned npiece;
    // the pieces alternate between two copy streams: one stream leaves the link idle between two copies
    cudaError_t piece(uint64_t a, uint64_t b) {
        if (b <= a) return cudaSuccess;
        d->last_d2h += 32 * (b - a);
        return cudaMemcpyAsync(out + 32 * a, wb + 32 * a, 32 * (b - a), cudaMemcpyDeviceToHost, (npiece++ & 1u) ? d->stream6 : d->stream3);
    }
    // [a, b) of the witness around the holes of jobs [j0, j1) (in memory order)
    cudaError_t pieces_around(uint64_t a, uint64_t b, int j0, int j1) {
        uint64_t at = a;
        for (int job = j0; job < j1; job++) {
            const HzHole h = d->h_holes[job];
            if (h.nz == 0) continue;
            const uint64_t h0 = h.cur, h1 = (uint64_t)h.cur + 2u + (uint64_t)h.nz * HZ_SZL;
            if (h0 < at || h1 > b) continue;                       // not inside this range
            cudaError_t e = piece(at, h0);
            if (e != cudaSuccess) return e;
            at = h1;
        }
        return piece(at, b);
    }
    // the slots [i0, i1) around the holes of their processors.  Eight consecutive slots share a block of
    // k_smt_levels and so the number of image levels; slots in the middle of the batch are laid out alike.  Such a
    // group goes as three pitched copies (before, between and after the two holes, one row per slot) instead of
    // 24 pieces: the copy engine spends microseconds per copy whatever its size.
    cudaError_t slots_around(int i0, int i1) {
        const std::vector<uint32_t>& off = *d->h_off_rollup;
        for (int g0 = i0; g0 < i1; g0 += 8) {
            const int g1 = g0 + 8 < i1 ? g0 + 8 : i1, n = g1 - g0;
            const uint64_t pitch = (uint64_t)off[g0 + 1] - off[g0];
            const HzHole a0 = d->h_holes[2 * g0], b0 = d->h_holes[2 * g0 + 1];
            bool uniform = n > 1 && a0.nz > 0 && b0.nz > 0;
            for (int i = g0; i < g1 && uniform; i++) {
                const HzHole a = d->h_holes[2 * i], b = d->h_holes[2 * i + 1];
                uniform = (uint64_t)off[i + 1] - off[i] == pitch && a.nz == a0.nz && b.nz == b0.nz &&
                          a.cur - off[i] == a0.cur - off[g0] && b.cur - off[i] == b0.cur - off[g0];
            }
            if (!uniform) {
                cudaError_t e = pieces_around(off[g0], off[g1], 2 * g0, 2 * g1);
                if (e != cudaSuccess) return e;
                continue;
            }
            const uint64_t base = off[g0];
            const uint64_t r1s = a0.cur - base, r1e = r1s + 2u + (uint64_t)a0.nz * HZ_SZL, r2s = b0.cur - base, r2e = r2s + 2u + (uint64_t)b0.nz * HZ_SZL;
            const uint64_t seg[3][2] = {{0, r1s}, {r1e, r2s}, {r2e, pitch}};
            for (int k = 0; k < 3; k++) {
                if (seg[k][1] <= seg[k][0]) continue;
                const uint64_t at = 32 * (base + seg[k][0]), wdt = 32 * (seg[k][1] - seg[k][0]);
                d->last_d2h += wdt * (uint64_t)n;
                cudaError_t e = cudaMemcpy2DAsync(out + at, 32 * pitch, wb + at, 32 * pitch, wdt, (size_t)n, cudaMemcpyDeviceToHost, (npiece++ & 1u) ? d->stream6 : d->stream3);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    }
};
static int hz_calculate_overlapped(HzCuda* d, const uint8_t* inputs, int sanity, uint8_t* out) {
    const HzPlan& pl = d->plan;
    const HzParams& P = pl.P;
    cudaStream_t s = d->stream, s2 = d->stream2, s3 = d->stream3, s5 = d->stream5;
    HzErrBuf* err = sanity ? d->err : nullptr;
    const uint8_t* wb = reinterpret_cast<const uint8_t*>(d->w);
    d->last_d2h = 0;
    const cudaStream_t s6 = d->stream6;
    HzCopier cp{d, out, wb, 0};
    auto piece = [&](uint64_t a, uint64_t b) { return cp.piece(a, b); };
    auto pieces_around = [&](uint64_t a, uint64_t b, int j0, int j1) { return cp.pieces_around(a, b, j0, j1); };
    auto slots_around = [&](int i0, int i1) { return cp.slots_around(i0, i1); };
    const uint32_t gs = HZ_GL_STRIDE(P.nLevels + 1);
    CK(cudaMemsetAsync(d->err, 0, sizeof(HzErrBuf), s));
    CK(cudaMemsetAsync(d->pre, 0, (size_t)pl.nblocks * 512, s));
    CK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d->w) + 64, inputs, 32ull * pl.n_inputs, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(d->ev0, s));
    CK(cudaStreamWaitEvent(s2, d->ev0, 0));
    k_edjob<<<hz_grid(P.nTx), HZ_BLOCK, 0, s2>>>(pl, d->w, d->jobs, d->off_eddsa, 0, P.nTx);
    CK(cudaStreamWaitEvent(d->stream4, d->ev0, 0));
    k_front<<<hz_grid(P.nTx), HZ_BLOCK, 0, d->stream4>>>(pl, d->w, d->pre, err, d->jobs_front, d->smt_front, d->isExit_front, 0, P.nTx);
    CK(cudaEventRecord(d->ev_front, d->stream4));
    k_smtjob<<<hz_grid(P.nTx), HZ_BLOCK, 0, s>>>(pl, d->w, d->smt, d->isExit, d->off_jk, 0, P.nTx);
    CK(cudaStreamWaitEvent(s3, d->ev_front, 0));
    CK(cudaStreamWaitEvent(s6, d->ev_front, 0));
    const int F = P.maxFeeTx, h0 = P.nTx / 2;
    k_fee_front<<<hz_grid(F), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->smt + 2 * P.nTx, 0, F);
    // w[0], w[1] are written by the hasher at the end: entries [2, decodeTx[0]) and the decode ranges go now
    CK(piece(2, d->h_off_rollup0));
    for (int h = 0; h < 2; h++) {
        const int s0 = h == 0 ? 0 : h0, ns = h == 0 ? h0 : P.nTx - h0, nf = h == 1 ? F : 0;
        if (ns > 0) {
            k_ed_const<<<592, 256, 0, s2>>>(pl, d->w, d->off_eddsa, d->hint_img, d->fix_size, s0, ns);
            k_eddsa<<<hz_grid(ns), HZ_BLOCK, 0, s2>>>(pl, d->w, err, d->jobs, d->hint_cst, d->hint_tab, d->hint_img, d->hint_scratch, d->win_size, s0, ns);
            k_ed_fill<<<(unsigned)((HZ_FILL_PER_SIG * ns + 127) / 128), 128, 0, s2>>>(d->w, err, d->hint_tab, d->hint_scratch, d->win_stride, ns);
        }
        CK(cudaEventRecord(d->ev_ed[h], s2));
        if (2 * ns + nf > 0) {
            k_smt<<<hz_grid(2 * ns + nf), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->smt, d->mids, d->lev, d->smst, s0, ns, 0, nf);
            const unsigned lb = (unsigned)((2 * ns + nf + 15) / 16);
            k_smt_levels<<<lb, 96, 0, s>>>(pl, d->w, d->lev, d->smst, d->mids, d->img00, d->zflag, s0, ns, 0, nf);
            k_hole_pack<<<(unsigned)(2 * ns + nf), 128, 0, s>>>(pl, d->w, d->lev, d->zflag, d->holes, d->gl, s0, ns, 0, nf);
            CK(cudaEventRecord(d->ev_pack[h], s));
            k_smt_zero<<<dim3(lb, (unsigned)(P.nLevels + 1)), 256, 0, s>>>(pl, d->w, d->lev, d->img00, d->zflag, s0, ns, 0, nf);
            // the hole records and glue entries of this half's jobs to the host, on a stream of their own
            CK(cudaStreamWaitEvent(s5, d->ev_pack[h], 0));
            CK(cudaMemcpyAsync(d->h_holes + 2 * s0, d->holes + 2 * s0, sizeof(HzHole) * 2 * ns, cudaMemcpyDeviceToHost, s5));
            CK(cudaMemcpyAsync(d->h_gl + (size_t)2 * s0 * gs, d->gl + (size_t)2 * s0 * gs, sizeof(Fr) * gs * 2 * ns, cudaMemcpyDeviceToHost, s5));
            if (nf > 0) {
                CK(cudaMemcpyAsync(d->h_holes + 2 * P.nTx, d->holes + 2 * P.nTx, sizeof(HzHole) * nf, cudaMemcpyDeviceToHost, s5));
                CK(cudaMemcpyAsync(d->h_gl + (size_t)2 * P.nTx * gs, d->gl + (size_t)2 * P.nTx * gs, sizeof(Fr) * gs * nf, cudaMemcpyDeviceToHost, s5));
            }
            d->last_d2h += (sizeof(HzHole) + sizeof(Fr) * gs) * (size_t)(2 * ns + nf);
        }
        CK(cudaEventRecord(d->ev_info[h], s5));
        if (h == 0) CK(cudaStreamWaitEvent(s, d->ev_front, 0));
        if (ns + nf > 0) k_tail<<<hz_grid(ns + nf), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->err, d->mids, d->isExit, d->isExit_front, d->jobs, d->jobs_front, d->smt, d->smt_front, s0, ns, 0, nf);
        CK(cudaStreamWaitEvent(s, d->ev_ed[h], 0));
        CK(cudaEventRecord(d->ev_half[h], s));
    }
    CK(cudaEventRecord(d->ev_eddsa, s2));
    int st = hz_phase2(d, sanity);
    if (st) return st;
    CK(cudaEventRecord(d->ev_done, s));
    // everything is queued; from here the host follows the device.  Half 0: its hole records, then host threads
    // rebuild its images while the copy engine moves the rest of its range around them
    const bool trace = getenv("HZW_TRACE") != nullptr;           // (measurement switch) host-side timeline on stderr
    const auto t00 = std::chrono::steady_clock::now();
    auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t00).count(); };
    unsigned nthr = std::thread::hardware_concurrency();
    // two cores stay free for this thread (it issues a few thousand copies while the workers run) and the driver's
    nthr = nthr == 0 ? 4 : (nthr > 18 ? 16 : (nthr > 3 ? nthr - 2 : 1));
    if (getenv("HZW_FILL_THREADS")) nthr = (unsigned)atoi(getenv("HZW_FILL_THREADS")) > 0 ? (unsigned)atoi(getenv("HZW_FILL_THREADS")) : nthr;
    std::vector<std::thread> workers;
    CK(cudaEventSynchronize(d->ev_info[0]));
    if (d->no_holes) for (int j = 0; j < 2 * h0; j++) d->h_holes[j].nz = 0;
    const double t_info0 = ms_since();
    for (unsigned t = 0; t < nthr; t++) workers.emplace_back(hz_fill_holes, d, out, 0, 2 * h0, (int)t, (int)nthr);
    CK(cudaStreamWaitEvent(s3, d->ev_half[0], 0));
    CK(cudaStreamWaitEvent(s6, d->ev_half[0], 0));
    cudaError_t ce = slots_around(0, h0);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(d->ev_info[1]);
    if (ce == cudaSuccess) {
        if (d->no_holes) for (int j = 2 * h0; j < 2 * P.nTx + F; j++) d->h_holes[j].nz = 0;
        for (unsigned t = 0; t < nthr; t++) workers.emplace_back(hz_fill_holes, d, out, 2 * h0, 2 * P.nTx + F, (int)t, (int)nthr);
        ce = cudaStreamWaitEvent(s3, d->ev_done, 0);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s6, d->ev_done, 0);
    }
    if (ce == cudaSuccess) ce = slots_around(h0, P.nTx);
    if (ce == cudaSuccess) ce = pieces_around(d->h_off_fee0, pl.nvars, 2 * P.nTx, 2 * P.nTx + F);
    if (ce == cudaSuccess) ce = piece(0, 2);
    const double t_issued = ms_since();
    for (auto& t : workers) t.join();
    const double t_filled = ms_since();
    if (trace) {
        cudaStreamSynchronize(s3); cudaStreamSynchronize(s6);
        fprintf(stderr, "hzw trace: hole records of half 0 at %.1f ms, copies issued at %.1f, host fill done at %.1f, copies done at %.1f (%u threads, %.2f GB copied)\n",
                t_info0, t_issued, t_filled, ms_since(), nthr, d->last_d2h / 1e9);
    }
    CK(ce);
    return 0;
}

int hz_cuda_calculate(HzCuda* d, const uint8_t* inputs, int sanity, uint8_t* witness_out, float* kernel_ms,
                      HzErrBuf* errs, uint64_t* launches, uint64_t* h2d, uint64_t* d2h) {
    CK(cudaSetDevice(d->device));
    if (d->world != 1) return HZW_E_PARAM;        // sharded contexts go through the two phases
    const HzPlan& pl = d->plan;
    cudaStream_t s = d->stream;
    const bool overlapped = witness_out && inputs && pl.P.nTx >= 8;
    if (overlapped) {
        int st = hz_calculate_overlapped(d, inputs, sanity, witness_out);
        if (st) return st;
    } else {
        int st = hz_phase1(d, inputs, sanity);
        if (st) return st;
        st = hz_phase2(d, sanity);
        if (st) return st;
        if (witness_out) CK(cudaMemcpyAsync(witness_out, d->w, 32ull * pl.nvars, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (overlapped) { CK(cudaStreamSynchronize(d->stream3)); CK(cudaStreamSynchronize(d->stream6)); }
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, d->ev0, d->ev1));
    if (launches) *launches = overlapped ? 24 : 15;        // kernels of ours per batch (see the launch sequences above)
    if (h2d) *h2d = inputs ? 32ull * pl.n_inputs : 0;
    if (d2h) *d2h = (overlapped ? d->last_d2h : (witness_out ? 32ull * pl.nvars : 0)) + sizeof(HzErrBuf);
    return 0;
}

// sharded phase 1: own slots / fee txs, then this rank's exchange record into `record_out` (device memory)
int hz_cuda_shard_phase1(HzCuda* d, const uint8_t* inputs, int sanity, void* record_out, HzErrBuf* errs) {
    CK(cudaSetDevice(d->device));
    if (!record_out) return HZW_E_PARAM;
    int st = hz_phase1(d, inputs, sanity);
    if (st) return st;
    const HzParams& P = d->plan.P;
    int last_slot = d->ns > 0 && d->s0 + d->ns == P.nTx, last_fee = d->nf > 0 && d->f0 + d->nf == P.maxFeeTx;
    k_pack<<<148, 256, 0, d->stream>>>(d->plan, d->w, d->pre, reinterpret_cast<uint8_t*>(record_out), d->s0, d->ns, d->per, last_slot, last_fee);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));         // the host binding runs the all-gather on its own stream
    return 0;
}
// sharded phase 2: `records` = the gathered records of all ranks (device memory); the shard's witness
// pieces are copied to `shard_out` (host) back to back in hz_cuda_pieces order
int hz_cuda_shard_phase2(HzCuda* d, const void* records, int sanity, uint8_t* shard_out, const uint64_t* pieces, int npieces,
                         float* kernel_ms, HzErrBuf* errs, uint64_t* d2h) {
    CK(cudaSetDevice(d->device));
    if (!records) return HZW_E_PARAM;
    cudaStream_t s = d->stream;
    k_unpack<<<148, 256, 0, s>>>(d->plan, d->w, d->pre, reinterpret_cast<const uint8_t*>(records), d->world, d->per, hz_rec_bytes(d->plan.P, d->per));
    int st = hz_phase2(d, sanity);
    if (st) return st;
    uint64_t bytes = 0;
    std::vector<std::thread> workers;
    d->last_d2h = 0;
    if (shard_out) {
        // as on the end-to-end path of one GPU: the image levels of the shard's processors are rebuilt in the shard
        // buffer by host threads, the pieces are copied around them
        const HzParams& P = d->plan.P;
        unsigned nthr = std::thread::hardware_concurrency();
        nthr = nthr == 0 ? 4 : (nthr > 18 ? 16 : (nthr > 3 ? nthr - 2 : 1));
        nthr = nthr / (unsigned)d->world > 0 ? nthr / (unsigned)d->world : 1;      // the ranks of a box share its cores
        // rebuilding the images pays at every rank count measured: with eight ranks a single thread per rank still
        // beats eight copy engines writing the whole 8.1 GB into the host's memory (80 ms against 97 ms per batch)
        const bool holes = d->ns + d->nf > 0 && !d->no_holes;
        CK(cudaEventRecord(d->ev_done, s));
        CK(cudaStreamWaitEvent(d->stream3, d->ev_done, 0));
        CK(cudaStreamWaitEvent(d->stream6, d->ev_done, 0));
        if (holes) CK(cudaEventSynchronize(d->ev_info[0]));
        const uint32_t r0 = d->ns > 0 ? (*d->h_off_rollup)[d->s0] : 0xffffffffu;
        const int fj0 = 2 * P.nTx + d->f0, fj1 = fj0 + d->nf;
        cudaError_t ce = cudaSuccess;
        for (int k = 0; k < npieces && ce == cudaSuccess; k++) {
            const uint64_t a = pieces[2 * k], n = pieces[2 * k + 1];
            uint8_t* out_v = shard_out + bytes - 32ull * a;          // where witness entry 0 would be for this piece
            HzCopier cp{d, out_v, reinterpret_cast<const uint8_t*>(d->w), (unsigned)k};
            if (holes && d->ns > 0 && a == r0 && n == (uint64_t)(*d->h_off_rollup)[d->s0 + d->ns] - r0) {
                for (unsigned t = 0; t < nthr; t++) workers.emplace_back(hz_fill_holes, d, out_v, 2 * d->s0, 2 * (d->s0 + d->ns), (int)t, (int)nthr);
                ce = cp.slots_around(d->s0, d->s0 + d->ns);
            } else if (holes && d->nf > 0) {
                // fee transactions' processors, if they lie in this piece
                bool inside = false;
                for (int j = fj0; j < fj1; j++) inside = inside || (d->h_holes[j].nz > 0 && d->h_holes[j].cur >= a && d->h_holes[j].cur < a + n);
                if (inside) {
                    // only the holes inside [a, a + n) are skipped by the copies; the fill must match: all of them lie in the fee range
                    bool all_in = true;
                    for (int j = fj0; j < fj1; j++)
                        if (d->h_holes[j].nz > 0) all_in = all_in && d->h_holes[j].cur >= a && (uint64_t)d->h_holes[j].cur + 2u + (uint64_t)d->h_holes[j].nz * HZ_SZL <= a + n;
                    if (all_in) {
                        workers.emplace_back(hz_fill_holes, d, out_v, fj0, fj1, 0, 1);
                        ce = cp.pieces_around(a, a + n, fj0, fj1);
                    } else ce = cp.piece(a, a + n);
                } else ce = cp.piece(a, a + n);
            } else ce = cp.piece(a, a + n);
            bytes += 32ull * n;
        }
        for (auto& t : workers) t.join();
        CK(ce);
        if (holes) d->last_d2h += (sizeof(HzHole) + sizeof(Fr) * HZ_GL_STRIDE(P.nLevels + 1)) * (size_t)(2 * d->ns + d->nf);
    }
    CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(d->stream2));
    CK(cudaStreamSynchronize(d->stream3));
    CK(cudaStreamSynchronize(d->stream6));
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, d->ev0, d->ev1));
    if (d2h) *d2h = (shard_out ? d->last_d2h : 0) + sizeof(HzErrBuf);
    return 0;
}

int hz_cuda_kernel_ms(HzCuda* d, float* ms, const char** names, int cap) {
    int n = cap < HZ_NKERN ? cap : HZ_NKERN;
    for (int k = 0; k < n; k++) {
        // kernels 0..7 are bracketed by kev[k], kev[k+1] on the main stream; k_eddsa by kev[9], kev[10]
        cudaEvent_t a = k < 8 ? d->kev[k] : (k == 8 ? d->kev[9] : d->kev[11]), b = k < 8 ? d->kev[k + 1] : (k == 8 ? d->kev[10] : d->kev[12]);
        if (cudaEventElapsedTime(&ms[k], a, b) != cudaSuccess) return HZW_E_DEVICE;
        if (names) names[k] = HZ_KERN_NAMES[k];
    }
    return n;
}

// ---- the same permutations, transposed ------------------------------------------------------------
// Thread (warp r, lane c) holds state element r of permutation c: a block of 3 warps runs 32 permutations.
// A partial round's single S-box is then a full warp of 32 different permutations (warp 0) instead of one
// active lane in eight, and the two other warps issue nothing while they wait for it; the state
// elements meet in shared memory (word-major, conflict-free) once per round for the MDS rows, whose
// constants are uniform per warp.  Same signals at the same indexes as the lane-group form.
#define HZ_PT_CHAINS 32
__global__ void __launch_bounds__(96) k_poseidon_t3(const Fr* in, Fr* sig, Fr* digests, int nperms) {
    __shared__ uint32_t xch[2][3][8][HZ_PT_CHAINS];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    int p = (int)blockIdx.x * HZ_PT_CHAINS + c;
    const bool real = p < nperms;
    if (!real) p = nperms - 1;
    const uint32_t size = 1 + 2 * 3 * 65 + 3 * (8 * 3 + 57);
    Fr* w = sig + (size_t)p * size;
    // inputs are wires (no entries); entry 0 is ark[0].in[0] = 0
    Fr s = r == 0 ? fr_zero() : in[2 * p + (r - 1)];
    if (r == 0 && real) hz_store_fr(&w[0], s);
    uint32_t cur = 1;
    const Fr* C = &hz_poseidon_c(3, 0);
    Fr m[3];
    for (int k = 0; k < 3; k++) m[k] = hz_poseidon_mr(3, r)[k];
    for (int i = 0; i < 65; i++) {
        s = fr_add(s, C[3 * i + r]);
        if (real) hz_store_fr(&w[cur + r], s);
        const bool full = i < 4 || i >= 61;
        const uint32_t sg = cur + 6;
        if (full || r == 0) {                       // warp-uniform
            Fr in2 = fr_mul_chain(s, s);
            Fr in4 = fr_mul_chain(in2, in2);
            s = fr_mul_chain(in4, s);
            if (real) { hz_store_fr(&w[sg + 3 * r], s); hz_store_fr(&w[sg + 3 * r + 1], in2); hz_store_fr(&w[sg + 3 * r + 2], in4); }
        }
        uint32_t (*buf)[8][HZ_PT_CHAINS] = xch[i & 1];
#pragma unroll
        for (int k = 0; k < 8; k++) buf[r][k][c] = s.l[k];
        __syncthreads();
        Fr a[3];
#pragma unroll
        for (int q = 0; q < 3; q++)
#pragma unroll
            for (int k = 0; k < 8; k++) a[q].l[k] = buf[q][k][c];
        s = fr_dot_mont_chain<3>(a, m);
        if (real) hz_store_fr(&w[cur + 3 + r], s);
        cur += 6 + (full ? 9 : 3);
    }
    if (r == 0 && real) digests[p] = s;
}

extern "C" int hzw_poseidon_bench(int device, uint64_t npairs, int reps, const uint8_t* inputs_le32 /* 4 x npairs */,
                                  uint8_t* digests_le32 /* 2 x npairs, nullable */, float* ms_per_rep) {
    if (npairs == 0 || npairs > (1u << 22) || !inputs_le32) return HZW_E_PARAM;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return HZW_E_DEVICE;
    CK(cudaSetDevice(device));
    const uint32_t size = 1 + 2 * 3 * 65 + 3 * (8 * 3 + 57);   // poseidon_size(3): entries of a wired Poseidon(2)
    Fr *in = nullptr, *sig = nullptr, *dg = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaMalloc(&in, 128ull * npairs));
    CK(cudaMalloc(&sig, 64ull * size * npairs));
    CK(cudaMalloc(&dg, 64ull * npairs));
    CK(cudaMemcpy(in, inputs_le32, 128ull * npairs, cudaMemcpyHostToDevice));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // reps < 0: the lane-group form (as inside k_smt); reps > 0: the transposed form
    const bool lanes = reps < 0;
    if (lanes) reps = -reps;
    const int nperms = 2 * (int)npairs;
    const unsigned tgrid = (unsigned)((nperms + HZ_PT_CHAINS - 1) / HZ_PT_CHAINS);
    if (lanes) k_poseidon_pairs<<<hz_grid((int)npairs), HZ_BLOCK>>>(in, sig, dg, (int)npairs);      // warm-up
    else k_poseidon_t3<<<tgrid, 96>>>(in, sig, dg, nperms);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) {
        if (lanes) k_poseidon_pairs<<<hz_grid((int)npairs), HZ_BLOCK>>>(in, sig, dg, (int)npairs);
        else k_poseidon_t3<<<tgrid, 96>>>(in, sig, dg, nperms);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_rep) *ms_per_rep = ms / (reps > 0 ? reps : 1);
    if (digests_le32) CK(cudaMemcpy(digests_le32, dg, 64ull * npairs, cudaMemcpyDeviceToHost));
    cudaFree(in); cudaFree(sig); cudaFree(dg); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// ---- measured ceiling of the pipe the field arithmetic runs on ------------------------------------
// 32 x 32 -> 64-bit multiply-adds (IMAD.WIDE.U32) per second with nothing else in the way: 16 independent
// accumulator chains per thread, full occupancy.  This is the denominator of the integer roofline in
// bench.py (MEASURED_PEAKS.json only has HBM and bf16 figures).
__global__ void __launch_bounds__(256) k_imad_peak(uint64_t* out, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x + k), "r"(y));
    }
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) r ^= acc[k];
    if (r == 0x1234567ull) out[0] = r;          // keeps the chains alive
}
extern "C" int hzw_int_peak(int device, double* lane_products_per_s) {
    int n = 0;
    if (!lane_products_per_s) return HZW_E_PARAM;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return HZW_E_DEVICE;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    uint64_t* out = nullptr;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    k_imad_peak<<<blocks, 256>>>(out, 3, 5, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_imad_peak<<<blocks, 256>>>(out, 3, 5, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *lane_products_per_s = (double)blocks * 256.0 * 16.0 * iters / (best * 1e-3);
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// ---- Withdraw(nLevels): one lane group per instance, each with its own witness range ----------------
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_withdraw(Fr* w, uint32_t nvars, int nLevels, int n, HzErrBuf* err) {
    bool real; int i = hz_item(n, real);
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w + (size_t)i * nvars, 2 + hz_withdraw_ninputs(nLevels), real ? err : nullptr, nullptr, &pend);
    VC h = sec_withdraw(c, nLevels);
    c.put(1, h.v);
    c.put(0, fr_one());
    c.flush_pending();
}
int hz_cuda_probe(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        fprintf(stderr, "hzw: no CUDA device -- this library has no CPU path\n");
        return HZW_E_DEVICE;
    }
    return device < 0 || device >= n ? HZW_E_PARAM : 0;
}
int hz_cuda_withdraw(int device, int nLevels, uint32_t nvars, uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out,
                     float* kernel_ms, HzErrBuf* errs) {
    CK(cudaSetDevice(device));
    CK(cudaDeviceSetLimit(cudaLimitStackSize, 80 * 1024));
    const uint32_t nin = 7u + (uint32_t)nLevels + 1u;          // hz_withdraw_ninputs (a device function in this translation unit)
    const uint64_t chunk = 2048;                 // instances per launch: bounds the device witness buffer (~5.5 GB at nLevels 32)
    const uint64_t cap = n < chunk ? n : chunk;
    Fr* w = nullptr; HzErrBuf* err = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaMalloc(&w, 32ull * nvars * cap));
    CK(cudaMalloc(&err, sizeof(HzErrBuf)));
    CK(cudaMemset(err, 0, sizeof(HzErrBuf)));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float total = 0;
    for (uint64_t k0 = 0; k0 < n; k0 += chunk) {
        const uint64_t m = n - k0 < chunk ? n - k0 : chunk;
        // inputs go to entries [2, 2 + nin) of each instance
        CK(cudaMemcpy2D(reinterpret_cast<uint8_t*>(w) + 64, 32ull * nvars, inputs + 32ull * nin * k0, 32ull * nin, 32ull * nin, m, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e0));
        k_withdraw<<<hz_grid((int)m), HZ_BLOCK>>>(w, nvars, nLevels, (int)m, sanity ? err : nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaGetLastError());
        CK(cudaMemcpy(out + 32ull * nvars * k0, w, 32ull * nvars * m, cudaMemcpyDeviceToHost));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    CK(cudaMemcpy(errs, err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost));
    if (kernel_ms) *kernel_ms = total;
    cudaFree(w); cudaFree(err); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
