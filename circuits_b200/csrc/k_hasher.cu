// hasherInputs (SHA-256 over bit signals), the shard exchange records, Withdraw (see hzw_dev.cuh).
#include "hzw_dev.cuh"

// ---- the exchange between shards (one all-gather, done by the host binding) ---------------------
// What hasherInputs needs from slots this GPU did not evaluate: their SHA pre-image bits
// (L1TxFullData, 624 bits for the first maxL1Tx slots, and L1L2TxData with the amount nullified,
// 2 nLevels + 48 bits) and the three values that close the chains (newLastIdx, newStateRoot,
// newExitRoot).  A rank's record: per own slot 624 + nL2 bytes (one bit per byte), then 3 x 32 B
// values and 3 x 4 B flags saying which of them this rank produced.
__global__ void k_pack(HzPlan pl, const Fr* w, const uint8_t* pre, HzPeerRecs dst, int s0, int ns, int per, int own_last_slot, int own_last_fee) {
    // dst.p[0 .. dst.n): where this rank's record goes -- one buffer when the caller runs the collective (NCCL in
    // the Python binding), the slot of this rank in EVERY peer's gathered array when the library drives all the GPUs of
    // a box itself (hz_cuda_multi_phase1): the stores then travel over NVLink as they are produced, and the pack
    // kernel IS the all-gather
    const HzParams& P = pl.P;
    const uint32_t sb = hz_rec_slot_bytes(P), nL2 = 2u * (uint32_t)P.nLevels + 48u;
    const uint32_t p1 = hz_pre_pos_l1(P), p2 = hz_pre_pos_l2(P);
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // four record bytes per store: peer stores are 32-bit words ((per * sb) is padded up to a word by hz_rec_vals_off)
    const uint32_t words = hz_rec_vals_off(P, per) / 4u;
    for (uint32_t x = gid; x < words; x += gsz) {
        uint32_t v4 = 0;
        for (uint32_t b = 0; b < 4; b++) {
            const uint32_t e = 4u * x + b;
            if (e >= (uint32_t)per * sb) break;
            uint32_t q = e / sb, k = e % sb, i = (uint32_t)s0 + q;
            uint32_t v = 0;
            if ((int)q < ns) {
                if (k < 624) v = (int)i < P.maxL1Tx ? pre[p1 + i * 624 + k] : 0;
                else v = pre[p2 + i * nL2 + (k - 624)];
            }
            v4 |= v << (8u * b);
        }
        for (int r = 0; r < dst.n; r++) reinterpret_cast<uint32_t*>(dst.p[r])[x] = v4;
    }
    if (gid == 0) {
        Fr z = fr_zero();
        const Fr& a = own_last_slot ? w[pl.idx_newLastIdx] : z;
        const Fr& b = own_last_fee ? w[pl.idx_newStateRoot] : z;
        const Fr& c = own_last_slot ? w[pl.idx_newExitRoot] : z;
        for (int r = 0; r < dst.n; r++) {
            uint32_t* out = reinterpret_cast<uint32_t*>(dst.p[r] + hz_rec_vals_off(P, per));
            for (int k = 0; k < 8; k++) { out[k] = a.l[k]; out[8 + k] = b.l[k]; out[16 + k] = c.l[k]; }
            out[24] = (uint32_t)own_last_slot; out[25] = (uint32_t)own_last_fee; out[26] = (uint32_t)own_last_slot;
        }
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
__device__ __forceinline__ void hz_put_bit(Fr* p, uint32_t bit) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(bit, 0, 0, 0));
    __stcs(reinterpret_cast<uint4*>(p) + 1, make_uint4(0, 0, 0, 0));
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
// The walk is one dependency chain of 49 024 rounds, so what counts is the number of DEPENDENT instructions per round:
// h and d are old values (e and a of three rounds ago), so h + wk and d + h + wk are formed off the chain and the new
// e and a are one three-input add behind their sigma: rotate -> xor3 -> add3, three instructions per round on either
// chain (five with the textbook grouping t1 = h + S1 + ch + wk; e = d + t1; a = t1 + S0 + maj).
__device__ __forceinline__ uint32_t hz_add3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("{\n .reg .u32 t;\n add.u32 t, %1, %2;\n add.u32 %0, t, %3;\n}" : "=r"(r) : "r"(a), "r"(b), "r"(c));   // ptxas fuses the pair into one IADD3
    return r;
}
__device__ __forceinline__ void hz_sha_round_wk(uint32_t* v, uint32_t wk) {
    const uint32_t hw = v[7] + wk, dhw = v[3] + hw;                         // off the chains
    const uint32_t S1 = hz_rr(v[4], 6) ^ hz_rr(v[4], 11) ^ hz_rr(v[4], 25);
    const uint32_t ch = (v[4] & v[5]) ^ (~v[4] & v[6]);
    const uint32_t e_new = hz_add3(S1, ch, dhw);
    const uint32_t t1 = hz_add3(S1, ch, hw);
    const uint32_t S0 = hz_rr(v[0], 2) ^ hz_rr(v[0], 13) ^ hz_rr(v[0], 22);
    const uint32_t mj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
    const uint32_t a_new = hz_add3(S0, mj, t1);
    v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = e_new; v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = a_new;
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

// ---- Withdraw(nLevels): one lane group per instance, each with its own witness range ----------------
// The entries of the sha256compression blocks are bit signals, 23 M of them at 2048 slots (0.7 GB as field elements).
// On their way to a host buffer they travel packed: 32 consecutive entries per word (lane = bit), and *notbit is
// raised if an entry is anything but 0 or 1 -- the host then copies the range as it is.  count is a multiple of 32.
__global__ void __launch_bounds__(256) k_bitpack(const Fr* w, uint64_t count, uint32_t* packed, uint32_t* notbit) {
    const uint64_t nwords = count / 32;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = (int)(threadIdx.x & 31);
    bool bad = false;
    for (uint64_t wd = warp; wd < nwords; wd += nwarps) {
        const uint4* p = reinterpret_cast<const uint4*>(w + 32 * wd + lane);
        const uint4 a = __ldcs(p), b = __ldcs(p + 1);
        bad = bad || (a.x > 1u) || ((a.y | a.z | a.w | b.x | b.y | b.z | b.w) != 0u);
        const uint32_t bits = __ballot_sync(0xffffffffu, a.x & 1u);
        if (lane == 0) packed[wd] = bits;
    }
    if (bad) atomicOr(notbit, 1u);
}
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_withdraw(Fr* w, uint32_t nvars, int nLevels, int n, HzErrBuf* err) {
    bool real; int i = hz_item(n, real);
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w + (size_t)i * nvars, 2 + hz_withdraw_ninputs(nLevels), real ? err : nullptr, nullptr, &pend);
    VC h = sec_withdraw(c, nLevels);
    c.put(1, h.v);
    c.put(0, fr_one());
    c.flush_pending();
}
// ---- standalone RollupTx(nLevels, maxFeeTx): one lane group per instance, the template evaluated as the main component
__global__ void __launch_bounds__(HZ_BLOCK, 2) k_rolluptx(Fr* w, uint32_t nvars, int nLevels, int maxFeeTx, int n, HzErrBuf* err) {
    bool real; int i = hz_item(n, real);
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w + (size_t)i * nvars, 1, real ? err : nullptr, nullptr, &pend);
    RollupTxOut<MCompute> o;
    sec_rolluptx_main(c, nLevels, maxFeeTx, o);
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
// Withdraw batches: a context keeps two witness buffers of `cap` instances each, so that the device->host copy of one
// chunk (when the caller wants the witnesses) runs beside the kernel of the next; with witness_out == NULL the
// witnesses stay in HBM (the last two chunks of them) and only the public output of every instance -- witness entry 1,
// hashGlobalInputs -- goes back.
struct HzWdCuda {
    int device, nLevels, maxFeeTx; uint32_t nvars, nin; uint64_t cap;
    Fr* w[2]; HzErrBuf* err; cudaStream_t st[2]; cudaEvent_t e0, e1, done[2];
};
HzWdCuda* hz_cuda_withdraw_create(int device, int nLevels, int maxFeeTx, uint32_t n_inputs, uint32_t nvars, int* status) {
    *status = hz_cuda_probe(device);
    if (*status) return nullptr;
    *status = HZW_E_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    HzWdCuda* d = new HzWdCuda();
    memset(d, 0, sizeof *d);
    d->device = device; d->nLevels = nLevels; d->maxFeeTx = maxFeeTx; d->nvars = nvars;
    d->nin = n_inputs;
    d->cap = 0;                                   // buffers are sized by the first batch (hz_wd_reserve)
    bool ok = cudaMalloc(&d->err, sizeof(HzErrBuf)) == cudaSuccess;
    for (int k = 0; k < 2; k++) ok = ok && cudaStreamCreate(&d->st[k]) == cudaSuccess && cudaEventCreateWithFlags(&d->done[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreate(&d->e0) == cudaSuccess && cudaEventCreate(&d->e1) == cudaSuccess;
    ok = ok && cudaDeviceSetLimit(cudaLimitStackSize, 80 * 1024) == cudaSuccess;
    if (!ok) { hz_cuda_withdraw_destroy(d); return nullptr; }
    *status = 0;
    return d;
}
void hz_cuda_withdraw_destroy(HzWdCuda* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    for (int k = 0; k < 2; k++) { cudaFree(d->w[k]); if (d->st[k]) cudaStreamDestroy(d->st[k]); if (d->done[k]) cudaEventDestroy(d->done[k]); }
    cudaFree(d->err);
    if (d->e0) cudaEventDestroy(d->e0);
    if (d->e1) cudaEventDestroy(d->e1);
    delete d;
}
static int hz_wd_reserve(HzWdCuda* d, uint64_t n) {
    // instances per launch: 4096 fill the machine with lane groups several times over (12 GB per buffer at nLevels 32)
    const uint64_t big = d->maxFeeTx > 0 ? 1024 : 4096;
    const uint64_t want = n < big ? n : big;
    if (want <= d->cap) return 0;
    for (int k = 0; k < 2; k++) { cudaFree(d->w[k]); d->w[k] = nullptr; }
    d->cap = 0;
    for (int k = 0; k < 2; k++) CK(cudaMalloc(&d->w[k], 32ull * d->nvars * want));
    d->cap = want;
    return 0;
}
// instances [0, n): inputs n x nin x 32 B (host); out (host, nullable) n x nvars x 32 B; pub_out (host, nullable) n x 32 B
int hz_cuda_withdraw_batch(HzWdCuda* d, uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, uint8_t* pub_out,
                           float* kernel_ms, HzErrBuf* errs, uint64_t* launches) {
    CK(cudaSetDevice(d->device));
    if (n == 0) { memset(errs, 0, sizeof(HzErrBuf)); if (kernel_ms) *kernel_ms = 0; return 0; }
    int st = hz_wd_reserve(d, n);
    if (st) return st;
    const uint64_t chunk = d->cap;
    CK(cudaMemsetAsync(d->err, 0, sizeof(HzErrBuf), d->st[0]));
    CK(cudaEventRecord(d->e0, d->st[0]));
    CK(cudaStreamWaitEvent(d->st[1], d->e0, 0));
    uint64_t nl = 0; int c = 0;
    for (uint64_t k0 = 0; k0 < n; k0 += chunk, c++) {
        const uint64_t m = n - k0 < chunk ? n - k0 : chunk;
        const int b = c & 1;
        Fr* w = d->w[b];
        const uint8_t* in0 = inputs + 32ull * d->nin * k0;
        if (d->maxFeeTx > 0) {
            // the inputs of a RollupTx sit in three runs around the outputs declared among them (hz_rolluptx_in_index)
            const uint32_t F = (uint32_t)d->maxFeeTx, nin = d->nin;
            const uint32_t run[3][3] = {{0, 2 * F, 1}, {2 * F, nin - 1 - 2 * F, 1 + 3 * F}, {nin - 1, 1, 1 + F + nin}};   // first input, count, witness index
            for (int q = 0; q < 3; q++)
                CK(cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(w) + 32ull * run[q][2], 32ull * d->nvars, in0 + 32ull * run[q][0], 32ull * nin,
                                     32ull * run[q][1], m, cudaMemcpyHostToDevice, d->st[b]));
            k_rolluptx<<<hz_grid((int)m), HZ_BLOCK, 0, d->st[b]>>>(w, d->nvars, d->nLevels, d->maxFeeTx, (int)m, sanity ? d->err : nullptr);
        } else {
            // inputs go to entries [2, 2 + nin) of each instance
            CK(cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(w) + 64, 32ull * d->nvars, in0, 32ull * d->nin, 32ull * d->nin, m, cudaMemcpyHostToDevice, d->st[b]));
            k_withdraw<<<hz_grid((int)m), HZ_BLOCK, 0, d->st[b]>>>(w, d->nvars, d->nLevels, (int)m, sanity ? d->err : nullptr);
        }
        nl++;
        CK(cudaGetLastError());
        if (out) CK(cudaMemcpyAsync(out + 32ull * d->nvars * k0, w, 32ull * d->nvars * m, cudaMemcpyDeviceToHost, d->st[b]));
        if (pub_out) CK(cudaMemcpy2DAsync(pub_out + 32ull * k0, 32, reinterpret_cast<uint8_t*>(w) + 32, 32ull * d->nvars, 32, m, cudaMemcpyDeviceToHost, d->st[b]));
    }
    CK(cudaEventRecord(d->done[1], d->st[1]));
    CK(cudaStreamWaitEvent(d->st[0], d->done[1], 0));
    CK(cudaEventRecord(d->e1, d->st[0]));
    CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, d->st[0]));
    CK(cudaStreamSynchronize(d->st[0]));
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, d->e0, d->e1));
    if (launches) *launches = nl;
    return 0;
}
