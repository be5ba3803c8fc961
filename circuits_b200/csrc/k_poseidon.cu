// The Poseidon permutations outside the merkle levels, transposed (see hzw_dev.cuh for the map of the kernels).
//
// A slot has four state hashes (HashState, width 5: src/lib/hash-state.circom via src/rollup-tx.circom:297-312,517-532),
// a fee tx two (src/fee-tx.circom), and a slot's message hash has width 7 (src/decode-tx.circom:254-283).  In the
// lane-group kernels each of them kept eight lanes busy for 68 rounds with one active lane in the partial rounds, one
// after the other.  Here thread (warp r, lane c) holds state element r of permutation c, a block of T warps runs 32
// permutations (as k_smt_levels does for the level hashes): all state hashes of the batch are one launch of 260 blocks,
// and their digests complete the job records of the merkle processors.  The components' own entries are written here;
// the template kernels (k_front, k_fee_front, k_smtjob) step over them (Comp::skip_hash).
#include "hzw_dev.cuh"

template <int T> __device__ __forceinline__ Fr hz_poseidon_tr(uint32_t (*xch)[T][8][32], int r, int c, Fr s, Fr* hw, bool real) {
    const int rp = poseidon_rp(T), R = 8 + rp;
    const Fr* C = &hz_poseidon_c(T, 0);
    Fr m[T];
#pragma unroll
    for (int k = 0; k < T; k++) m[k] = hz_poseidon_mr(T, r)[k];
    if (r == 0 && real) hz_store_fr(&hw[0], s);                 // ark[0].in[0] = 0 (s is zero for element 0)
    uint32_t cur = 1;
    for (int i = 0; i < R; i++) {
        s = fr_add(s, C[T * i + r]);
        if (real) hz_store_fr(&hw[cur + r], s);
        const bool full = i < 4 || i >= rp + 4;
        const uint32_t sg = cur + 2 * T;
        if (full || r == 0) {                                   // warp-uniform
            Fr in2 = fr_mul_chain(s, s);
            Fr in4 = fr_mul_chain(in2, in2);
            s = fr_mul_chain(in4, s);
            if (real) { hz_store_fr(&hw[sg + 3 * r], s); hz_store_fr(&hw[sg + 3 * r + 1], in2); hz_store_fr(&hw[sg + 3 * r + 2], in4); }
        }
        uint32_t (*buf)[8][32] = xch[i & 1];
#pragma unroll
        for (int k = 0; k < 8; k++) buf[r][k][c] = s.l[k];
        __syncthreads();
        Fr a[T];
#pragma unroll
        for (int q = 0; q < T; q++)
#pragma unroll
            for (int k = 0; k < 8; k++) a[q].l[k] = buf[q][k][c];
        s = fr_dot_mont_chain<T>(a, m);
        if (real) hz_store_fr(&hw[cur + T + r], s);
        cur += 2 * T + (full ? 3 * T : 3);
    }
    return s;
}

// all HashState permutations of the shard: 4 per own slot (old state 1, old state 2, new state 1, new state 2), then 2
// per own fee tx (old, new).  hs_in: their 4 inputs each, recorded by k_smtjob / k_fee_front; hs_cur: where each
// component starts (layout walk).  The digests complete the job records: newValue, and oldValue behind its multiplexer.
__global__ void __launch_bounds__(160) k_hstate(Fr* w, const Fr* hs_in, const uint32_t* hs_cur, HzSmtJob* smt, int nTx, int s0, int ns, int f0, int nf) {
    __shared__ uint32_t xch[2][5][8][32];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    const int P = 4 * ns + 2 * nf;
    int p = (int)blockIdx.x * 32 + c;
    const bool real = p < P;
    if (!real) p = P - 1;
    int idx, jobi; bool is_new;
    if (p < 4 * ns) { const int i = s0 + p / 4, k = p % 4; idx = 4 * i + k; jobi = 2 * i + (k & 1); is_new = k >= 2; }
    else { const int q = p - 4 * ns, i = f0 + q / 2, k = q % 2; idx = 4 * nTx + 2 * i + k; jobi = 2 * nTx + i; is_new = k == 1; }
    const Fr s = r == 0 ? fr_zero() : hs_in[4 * (size_t)idx + (r - 1)];
    const Fr d = hz_poseidon_tr<5>(xch, r, c, s, w + hs_cur[idx], real);
    if (r == 0 && real) {
        HzSmtJob* j = &smt[jobi];
        if (is_new) j->newValue = d;
        else {
            const Fr ov = fr_add(j->oldValue, fr_mul(fr_sub(fr_one(), j->hsel), d));
            j->oldValue = ov;
            if (j->ovidx != HZ_NOIDX) hz_store_fr(&w[j->ovidx], ov);          // s1OldValue / s2OldValue .mux.out[0]
        }
    }
}

// the message hash of every own slot: sig_in = its 6 inputs (k_edjob), hsig_cur = where decodeTx[i].hashSig starts
__global__ void __launch_bounds__(224) k_sighash(Fr* w, const Fr* sig_in, const uint32_t* hsig_cur, HzEddsaJob* jobs, int s0, int ns) {
    __shared__ uint32_t xch[2][7][8][32];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    int p = (int)blockIdx.x * 32 + c;
    const bool real = p < ns;
    if (!real) p = ns - 1;
    const int i = s0 + p;
    const Fr s = r == 0 ? fr_zero() : sig_in[6 * (size_t)i + (r - 1)];
    const Fr d = hz_poseidon_tr<7>(xch, r, c, s, w + hsig_cur[i], real);
    if (r == 0 && real) jobs[i].M = d;
}

// the two leaf hashes of every processor of the shard: hash1Old = Poseidon(oldKey, oldValue, 1), hash1New =
// Poseidon(newKey, newValue, 1) (circomlib smthash_poseidon.circom); their digests go to the level job.  The
// component starts after the processor's own inputs: newRoot, the inputs that are not wired signals, `enabled`;
// its first entry is the constant third input.
__global__ void __launch_bounds__(128) k_leafhash(HzPlan pl, Fr* w, const HzSmtJob* smt, HzLevJob* lev, int s0, int ns, int f0, int nf) {
    __shared__ uint32_t xch[2][4][8][32];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    const int P = 2 * (2 * ns + nf);
    int p = (int)blockIdx.x * 32 + c;
    const bool real = p < P;
    if (!real) p = P - 1;
    const int job = hz_smt_job_index(pl, p >> 1, s0, ns, f0);
    const bool is_new = (p & 1) != 0;
    const HzSmtJob& j = smt[job];
    const uint32_t h1 = j.cur + 2u + (uint32_t)__popc((~j.sigmask) & 0x7fu) + (is_new ? HZ_SZ_SMTHASH1_WIRED : 0u);
    Fr s = fr_zero();
    if (r == 1) s = is_new ? j.newKey : j.oldKey;
    else if (r == 2) s = is_new ? j.newValue : j.oldValue;
    else if (r == 3) s = fr_one();
    if (r == 3 && real) hz_store_fr(&w[h1], s);                    // h.inputs[2] = 1 owns an entry
    const Fr d = hz_poseidon_tr<4>(xch, r, c, s, w + h1 + 1, real);
    if (r == 0 && real) { if (is_new) lev[job].n1 = d; else lev[job].o1 = d; }
}

// ---- permutations/s per width (the second metric of BASELINE.md) --------------------------------------------------
// nperms independent Poseidon(T - 1) hashes with every signal of the component stored (entries per permutation:
// 1 + 2 T R + 3 T 8 + 3 rp, the wired size of circomlib's Poseidon), transposed as above: in = T - 1 inputs per
// permutation.  T = 3 is k_poseidon_t3 (k_merkle.cu), the loop of the level hashes with its chain in the Montgomery domain.
template <int T> __global__ void __launch_bounds__(32 * T) k_poseidon_tr_bench(const Fr* in, Fr* sig, Fr* digests, int nperms, uint32_t size) {
    __shared__ uint32_t xch[2][T][8][32];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    int p = (int)blockIdx.x * 32 + c;
    const bool real = p < nperms;
    if (!real) p = nperms - 1;
    const Fr s = r == 0 ? fr_zero() : in[(size_t)(T - 1) * p + (r - 1)];
    const Fr d = hz_poseidon_tr<T>(xch, r, c, s, sig + (size_t)p * size, real);
    if (r == 0 && real) digests[p] = d;
}
extern "C" int hzw_poseidon_bench_t(int device, int t, uint64_t nperms, int reps, const uint8_t* inputs_le32 /* (t - 1) x nperms x 32 B */,
                                    uint8_t* digests_le32 /* nperms x 32 B, nullable */, float* ms_per_rep, uint64_t* signal_bytes) {
    if (t < 4 || t > 7 || nperms == 0 || nperms > (1u << 20) || reps <= 0 || !inputs_le32) return HZW_E_PARAM;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return HZW_E_DEVICE;
    CK(cudaSetDevice(device));
    const int rp = t == 4 ? 56 : (t == 7 ? 63 : 60), R = 8 + rp;      // poseidon_rp(t) (templates_circomlib.cuh), host side
    const uint32_t size = 1u + 2u * (uint32_t)(t * R) + 3u * (uint32_t)(t * 8) + 3u * (uint32_t)rp;
    Fr *in = nullptr, *sig = nullptr, *dg = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaMalloc(&in, 32ull * (t - 1) * nperms));
    CK(cudaMalloc(&sig, 32ull * size * nperms));
    CK(cudaMalloc(&dg, 32ull * nperms));
    CK(cudaMemcpy(in, inputs_le32, 32ull * (t - 1) * nperms, cudaMemcpyHostToDevice));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const unsigned grid = (unsigned)((nperms + 31) / 32);
    for (int r = -1; r < reps; r++) {                       // r = -1: warm-up
        if (r == 0) CK(cudaEventRecord(e0));
        if (t == 4) k_poseidon_tr_bench<4><<<grid, 128>>>(in, sig, dg, (int)nperms, size);
        else if (t == 5) k_poseidon_tr_bench<5><<<grid, 160>>>(in, sig, dg, (int)nperms, size);
        else if (t == 6) k_poseidon_tr_bench<6><<<grid, 192>>>(in, sig, dg, (int)nperms, size);
        else k_poseidon_tr_bench<7><<<grid, 224>>>(in, sig, dg, (int)nperms, size);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_rep) *ms_per_rep = ms / reps;
    if (signal_bytes) *signal_bytes = 32ull * size * nperms;
    if (digests_le32) CK(cudaMemcpy(digests_le32, dg, 32ull * nperms, cudaMemcpyDeviceToHost));
    cudaFree(in); cudaFree(sig); cudaFree(dg); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// ---- digests only: the hashes of the batch builder (circuits_b200/batchtree.py) ------------------------------------
// n independent Poseidon(T - 1) hashes, transposed as above, nothing stored but the digest: what the sparse Merkle trees
// of a batch need, one launch per wave of equal dependency depth.
template <int T> __global__ void __launch_bounds__(32 * T) k_poseidon_digest(const Fr* in, Fr* digests, int nperms) {
    __shared__ uint32_t xch[2][T][8][32];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    int p = (int)blockIdx.x * 32 + c;
    const bool real = p < nperms;
    if (!real) p = nperms - 1;
    const Fr s = r == 0 ? fr_zero() : in[(size_t)(T - 1) * p + (r - 1)];
    const Fr d = hz_poseidon_tr<T>(xch, r, c, s, nullptr, false);
    if (r == 0 && real) digests[p] = d;
}
extern "C" int hzw_poseidon_hash(int device, int t, uint64_t n, const uint8_t* inputs_le32 /* (t - 1) x n x 32 B */, uint8_t* digests_le32 /* n x 32 B */) {
    if (t < 3 || t > 7 || n == 0 || n > (1u << 24) || !inputs_le32 || !digests_le32) return HZW_E_PARAM;
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0 || device < 0 || device >= nd) return HZW_E_DEVICE;
    CK(cudaSetDevice(device));
    Fr *in = nullptr, *dg = nullptr;
    CK(cudaMalloc(&in, 32ull * (t - 1) * n));
    if (cudaMalloc(&dg, 32ull * n) != cudaSuccess) { cudaFree(in); return HZW_E_DEVICE; }
    cudaError_t e = cudaMemcpy(in, inputs_le32, 32ull * (t - 1) * n, cudaMemcpyHostToDevice);
    const unsigned grid = (unsigned)((n + 31) / 32);
    if (e == cudaSuccess) {
        if (t == 3) k_poseidon_digest<3><<<grid, 96>>>(in, dg, (int)n);
        else if (t == 4) k_poseidon_digest<4><<<grid, 128>>>(in, dg, (int)n);
        else if (t == 5) k_poseidon_digest<5><<<grid, 160>>>(in, dg, (int)n);
        else if (t == 6) k_poseidon_digest<6><<<grid, 192>>>(in, dg, (int)n);
        else k_poseidon_digest<7><<<grid, 224>>>(in, dg, (int)n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(digests_le32, dg, 32ull * n, cudaMemcpyDeviceToHost);
    cudaFree(in); cudaFree(dg);
    if (e != cudaSuccess) { fprintf(stderr, "hzw: CUDA error %s in hzw_poseidon_hash\n", cudaGetErrorString(e)); return HZW_E_DEVICE; }
    return 0;
}
