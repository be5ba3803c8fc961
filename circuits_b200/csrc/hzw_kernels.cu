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

#define HZ_NKERN 9
static const char* const HZ_KERN_NAMES[HZ_NKERN] = {"k_front", "k_fee_front", "k_smt", "k_tail", "k_hash_head",
                                                    "k_sha_chain", "k_sha_blocks", "k_hash_tail", "k_eddsa"};
#define HZ_BLOCK 128                       /* 16 lane groups per block */

struct HzCuda {
    int device;
    HzPlan plan;               // device-pointer version
    Fr* w;                     // witness, nvars entries
    uint8_t* pre;              // SHA pre-image bits, nblocks*512
    uint32_t* chain;           // (nblocks+1) x 8
    uint32_t* msgw;            // nblocks x 16
    HzEddsaJob* jobs;          // nTx
    HzSmtJob* smt;             // 2 nTx + maxFeeTx
    HzSmtMid* mids;            // 2 nTx + maxFeeTx
    Fr* isExit;                // nTx
    Fr* hint_cst;              // HZ_CST_MAX: constant inverses of EscalarMulFix (recorded at create)
    Fr* hint_tab;              // HZ_EDTAB_LEN: its window entries
    Fr* hint_scratch;          // per lane group of k_eddsa: work area of the hint generators
    HzErrBuf* err;
    uint32_t *off_decode, *off_rollup, *off_fee;
    uint32_t* off_blocks;      // 1
    cudaStream_t stream, stream2;
    cudaEvent_t ev0, ev1, ev_front, ev_eddsa;
    cudaEvent_t kev[HZ_NKERN + 2];
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "hzw: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return HZW_E_DEVICE; } } while (0)

// item of this lane group; groups past the end redo the last item (identical stores) so that
// every warp stays whole for the shuffles, and report nothing
__device__ __forceinline__ int hz_item(int n, bool& real) {
    int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) / HZ_G);
    real = g < n;
    return real ? g : n - 1;
}
static inline unsigned hz_grid(int items) { return (unsigned)((items * HZ_G + HZ_BLOCK - 1) / HZ_BLOCK); }

__global__ void __launch_bounds__(HZ_BLOCK) k_front(HzPlan pl, Fr* w, uint8_t* preb, HzErrBuf* err, HzEddsaJob* jobs, HzSmtJob* smt, Fr* isExit) {
    bool real; int i = hz_item(pl.P.nTx, real);
    PreBits<MCompute> pre; pre.b = preb;
    SlotHook<MCompute> hook; hook.eddsa = &jobs[i]; hook.smt = &smt[2 * i]; hook.isExit = &isExit[i]; hook.sz = &pl.sz;
    Comp<MCompute> c(w, pl.off_decode[i], real ? err : nullptr);
    DecodeOut<MCompute> d;
    sec_decode(c, pl, i, d, pre);
    c.cur = pl.off_rollup[i];
    RollupTxOut<MCompute> o;
    sec_rollup(c, pl, i, d, o, pre, hook);
}

__global__ void __launch_bounds__(HZ_BLOCK) k_fee_front(HzPlan pl, Fr* w, HzErrBuf* err, HzSmtJob* smt) {
    bool real; int i = hz_item(pl.P.maxFeeTx, real);
    Comp<MCompute> c(w, pl.off_fee[i], real ? err : nullptr);
    FeeHook<MCompute> hook; hook.smt = &smt[i];
    hook.sib_idx = pl.in_off[IN_siblings3] + (uint32_t)i * pl.in_d1[IN_siblings3];
    sec_fee(c, pl, i, hook, true);
}

__global__ void __launch_bounds__(HZ_BLOCK) k_smt(HzPlan pl, Fr* w, HzErrBuf* err, const HzSmtJob* smt, HzSmtMid* mids, int njobs) {
    bool real; int q = hz_item(njobs, real);
    sec_smt_body(w, real ? err : nullptr, pl.P.nLevels + 1, smt[q], mids[q]);
}

// one-time: the constant division sites and window entries of EscalarMulFix(253, BASE8)
__global__ void k_fix_consts(Fr* wtmp, Fr* cst, Fr* tab, uint32_t* counts) {
    hz_record_fix_consts(wtmp, cst, tab, counts);
}

__global__ void __launch_bounds__(HZ_BLOCK) k_eddsa(HzPlan pl, Fr* w, HzErrBuf* err, const HzEddsaJob* jobs,
                                                    const Fr* cst, const Fr* tab, Fr* scratch) {
    bool real; int i = hz_item(pl.P.nTx, real);
    const HzEddsaJob& j = jobs[i];
    HzHints h;
    h.cst = cst; h.ci = 0; h.dat = nullptr; h.di = 0; h.rec = nullptr; h.ri = 0;
    h.edtab = tab; h.edrec = nullptr; h.ei = 0; h.scratch = scratch + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) / HZ_G) * HZ_HINT_SCRATCH;   // per group, not per slot: padding groups redo a slot
    Comp<MCompute> c(w, j.cur, real ? err : nullptr, &h);
    VC vse, sgn, ay, s1Ay, S, R8x, R8y, Msg;
    vse.v = j.enabled; sgn.v = j.signSignature; ay.v = j.aySignature; s1Ay.v = j.s1Ay;
    S.v = j.S; R8x.v = j.R8x; R8y.v = j.R8y; Msg.v = j.M;
    vse.sig = sgn.sig = ay.sig = s1Ay.sig = S.sig = R8x.sig = R8y.sig = Msg.sig = true;
    hz_phaseF(c, vse, sgn, ay, s1Ay, S, R8x, R8y, Msg);
}

__global__ void __launch_bounds__(HZ_BLOCK) k_tail(HzPlan pl, Fr* w, HzErrBuf* err, const HzSmtMid* mids, const Fr* isExit) {
    bool real; int i = hz_item(pl.P.nTx + pl.P.maxFeeTx, real);
    if (i < pl.P.nTx) sec_rollup_tail(w, real ? err : nullptr, pl, i, &mids[2 * i], isExit[i]);
    else sec_fee_tail(w, real ? err : nullptr, pl, i - pl.P.nTx, mids[2 * pl.P.nTx + (i - pl.P.nTx)]);
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

// chaining values of all blocks.  The chain is serial (block b needs the digest of block b-1), so
// one thread walks it, with the message schedule in a 16-word register window; the other threads
// only pack the pre-image bits into words first.
__global__ void __launch_bounds__(256) k_sha_chain(HzPlan pl, const uint8_t* preb, uint32_t* chain, uint32_t* msgw) {
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
    if (threadIdx.x != 0) return;
    uint32_t st[8];
    for (int j = 0; j < 8; j++) { st[j] = hz_sha_h(j); chain[j] = st[j]; }
    for (uint32_t b = 0; b < pl.nblocks; b++) {
        uint32_t wv[16], v[8];
#pragma unroll
        for (int t = 0; t < 16; t++) wv[t] = msgw[(size_t)b * 16 + t];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = st[j];
#pragma unroll
        for (int t = 0; t < 64; t++) {
            if (t >= 16) wv[t & 15] = hz_sha_sched(wv[(t - 2) & 15], wv[(t - 7) & 15], wv[(t - 15) & 15], wv[t & 15]);
            hz_sha_round(v, wv[t & 15], t);
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
__global__ void __launch_bounds__(HZ_SHA_THREADS) k_sha_blocks(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain, const uint32_t* msgw, const uint32_t* off_blocks) {
    const uint32_t b = blockIdx.x;
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
                       const uint32_t* off_fee, int* status) {
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
    ok = ok && cudaMalloc(&d->jobs, sizeof(HzEddsaJob) * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->smt, sizeof(HzSmtJob) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->mids, sizeof(HzSmtMid) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->isExit, sizeof(Fr) * P.nTx) == cudaSuccess;
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
    ok = ok && cudaEventCreate(&d->ev0) == cudaSuccess && cudaEventCreate(&d->ev1) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_front, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_eddsa, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < HZ_NKERN + 2; k++) ok = ok && cudaEventCreate(&d->kev[k]) == cudaSuccess;
    // the per-thread frames of the circuit templates are large (sibling arrays, state machines)
    ok = ok && cudaDeviceSetLimit(cudaLimitStackSize, 64 * 1024) == cudaSuccess;
    if (ok) {
        // record the input-independent part of the EdDSA fixed-base multiplication once
        Fr* wtmp = nullptr; uint32_t* counts = nullptr; uint32_t hc[2] = {0, 0};
        ok = cudaMalloc(&wtmp, sizeof(Fr) * ((size_t)ph.sz.sigVerifier + 64)) == cudaSuccess && cudaMalloc(&counts, 8) == cudaSuccess;
        if (ok) {
            k_fix_consts<<<1, 32, 0, d->stream>>>(wtmp, d->hint_cst, d->hint_tab, counts);
            ok = cudaMemcpyAsync(hc, counts, 8, cudaMemcpyDeviceToHost, d->stream) == cudaSuccess &&
                 cudaStreamSynchronize(d->stream) == cudaSuccess;
            if (ok && (hc[0] > HZ_CST_MAX || hc[1] != HZ_EDTAB_LEN)) {
                fprintf(stderr, "hzw: hint table sizes %u / %u do not match the templates\n", hc[0], hc[1]);
                ok = false;
            }
        }
        cudaFree(wtmp); cudaFree(counts);
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
    cudaFree(d->w); cudaFree(d->pre); cudaFree(d->chain); cudaFree(d->msgw); cudaFree(d->jobs); cudaFree(d->err);
    cudaFree(d->smt); cudaFree(d->mids); cudaFree(d->isExit);
    cudaFree(d->hint_cst); cudaFree(d->hint_tab); cudaFree(d->hint_scratch);
    cudaFree(d->off_decode); cudaFree(d->off_rollup); cudaFree(d->off_fee); cudaFree(d->off_blocks);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->stream2) cudaStreamDestroy(d->stream2);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev_front) cudaEventDestroy(d->ev_front);
    if (d->ev_eddsa) cudaEventDestroy(d->ev_eddsa);
    for (int k = 0; k < HZ_NKERN + 2; k++) if (d->kev[k]) cudaEventDestroy(d->kev[k]);
    delete d;
}

int hz_cuda_calculate(HzCuda* d, const uint8_t* inputs, int sanity, uint8_t* witness_out, float* kernel_ms,
                      HzErrBuf* errs, uint64_t* launches, uint64_t* h2d, uint64_t* d2h) {
    CK(cudaSetDevice(d->device));
    const HzPlan& pl = d->plan;
    const HzParams& P = pl.P;
    cudaStream_t s = d->stream;
    HzErrBuf* err = sanity ? d->err : nullptr;
    CK(cudaMemsetAsync(d->err, 0, sizeof(HzErrBuf), s));
    CK(cudaMemsetAsync(d->pre, 0, (size_t)pl.nblocks * 512, s));
    if (inputs) CK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d->w) + 64, inputs, 32ull * pl.n_inputs, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(d->ev0, s));
    const int nTx = P.nTx, F = P.maxFeeTx, njobs = 2 * nTx + F;
    cudaStream_t s2 = d->stream2;
    CK(cudaEventRecord(d->kev[0], s));
    k_front<<<hz_grid(nTx), HZ_BLOCK, 0, s>>>(pl, d->w, d->pre, err, d->jobs, d->smt, d->isExit);
    CK(cudaEventRecord(d->kev[1], s));
    CK(cudaEventRecord(d->ev_front, s));
    // the signatures run beside the merkle levels: neither fills the machine alone
    CK(cudaStreamWaitEvent(s2, d->ev_front, 0));
    CK(cudaEventRecord(d->kev[9], s2));
    k_eddsa<<<hz_grid(nTx), HZ_BLOCK, 0, s2>>>(pl, d->w, err, d->jobs, d->hint_cst, d->hint_tab, d->hint_scratch);
    CK(cudaEventRecord(d->kev[10], s2));
    CK(cudaEventRecord(d->ev_eddsa, s2));
    k_fee_front<<<hz_grid(F), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->smt + 2 * nTx);
    CK(cudaEventRecord(d->kev[2], s));
    k_smt<<<hz_grid(njobs), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->smt, d->mids, njobs);
    CK(cudaEventRecord(d->kev[3], s));
    k_tail<<<hz_grid(nTx + F), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->mids, d->isExit);
    CK(cudaEventRecord(d->kev[4], s));
    k_hash_head<<<148, 256, 0, s>>>(pl, d->w, d->pre, err, d->off_blocks);
    CK(cudaEventRecord(d->kev[5], s));
    k_sha_chain<<<1, 256, 0, s>>>(pl, d->pre, d->chain, d->msgw);
    CK(cudaEventRecord(d->kev[6], s));
    k_sha_blocks<<<pl.nblocks, HZ_SHA_THREADS, 0, s>>>(pl, d->w, err, d->chain, d->msgw, d->off_blocks);
    CK(cudaEventRecord(d->kev[7], s));
    k_hash_tail<<<1, 32, 0, s>>>(pl, d->w, err, d->chain);
    CK(cudaEventRecord(d->kev[8], s));
    CK(cudaStreamWaitEvent(s, d->ev_eddsa, 0));
    CK(cudaGetLastError());
    CK(cudaEventRecord(d->ev1, s));
    if (witness_out) CK(cudaMemcpyAsync(witness_out, d->w, 32ull * pl.nvars, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, d->ev0, d->ev1));
    if (launches) *launches = HZ_NKERN;
    if (h2d) *h2d = inputs ? 32ull * pl.n_inputs : 0;
    if (d2h) *d2h = (witness_out ? 32ull * pl.nvars : 0) + sizeof(HzErrBuf);
    return 0;
}

int hz_cuda_kernel_ms(HzCuda* d, float* ms, const char** names, int cap) {
    int n = cap < HZ_NKERN ? cap : HZ_NKERN;
    for (int k = 0; k < n; k++) {
        // kernels 0..7 are bracketed by kev[k], kev[k+1] on the main stream; k_eddsa by kev[9], kev[10]
        cudaEvent_t a = k < 8 ? d->kev[k] : d->kev[9], b = k < 8 ? d->kev[k + 1] : d->kev[10];
        if (cudaEventElapsedTime(&ms[k], a, b) != cudaSuccess) return HZW_E_DEVICE;
        if (names) names[k] = HZ_KERN_NAMES[k];
    }
    return n;
}
