// libhzwitness.so, CUDA half (sm_100a): device memory, the section kernels and their launch
// order.  Every kernel instantiates the circuit templates in MCompute mode; a thread owns one
// section of the circuit (rollup_main.cuh) and stores each signal at its witness index.
//
//   k_slots      one thread per tx slot: decodeTx[i] + rollupTx[i] (signature sub-graph handed off)
//   k_eddsa      one thread per tx slot: getAx + sigVerifier (phase F of rollup-tx.circom)
//   k_fee        one thread per fee tx
//   k_hash_head  hasherInputs up to the first compression (one thread)
//   k_sha_chain  SHA-256 chaining values of all blocks (one thread; the chain is serial)
//   k_sha_blocks one thread per sha256compression: all bit signals of the block
//   k_hash_tail  digest -> hashGlobalInputs
#define HZ_DEVICE_ONLY
#include "hzw_cuda.h"
#include "../../include/hzwitness.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

#define HZ_NKERN 7
static const char* const HZ_KERN_NAMES[HZ_NKERN] = {"k_slots", "k_eddsa", "k_fee", "k_hash_head", "k_sha_chain", "k_sha_blocks", "k_hash_tail"};

struct HzCuda {
    int device;
    HzPlan plan;               // device-pointer version
    Fr* w;                     // witness, nvars entries
    uint8_t* pre;              // SHA pre-image bits, nblocks*512
    uint32_t* chain;           // (nblocks+1) x 8
    uint32_t* msgw;            // nblocks x 16
    HzEddsaJob* jobs;          // nTx
    HzErrBuf* err;
    uint32_t *off_decode, *off_rollup, *off_fee;
    uint32_t* off_blocks;      // 1
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    cudaEvent_t kev[HZ_NKERN + 1];
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "hzw: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return HZW_E_DEVICE; } } while (0)

__global__ void k_slots(HzPlan pl, Fr* w, uint8_t* preb, HzErrBuf* err, HzEddsaJob* jobs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.P.nTx) return;
    PreBits<MCompute> pre; pre.b = preb;
    EddsaHook<MCompute> hook; hook.job = &jobs[i]; hook.sz = &pl.sz;
    Comp<MCompute> c(w, pl.off_decode[i], err);
    DecodeOut<MCompute> d;
    sec_decode(c, pl, i, d, pre);
    c.cur = pl.off_rollup[i];
    RollupTxOut<MCompute> o;
    sec_rollup(c, pl, i, d, o, pre, hook);
}

__global__ void k_eddsa(HzPlan pl, Fr* w, HzErrBuf* err, const HzEddsaJob* jobs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.P.nTx) return;
    const HzEddsaJob& j = jobs[i];
    Comp<MCompute> c(w, j.cur, err);
    VC vse, sgn, ay, s1Ay, S, R8x, R8y, Msg;
    vse.v = j.enabled; sgn.v = j.signSignature; ay.v = j.aySignature; s1Ay.v = j.s1Ay;
    S.v = j.S; R8x.v = j.R8x; R8y.v = j.R8y; Msg.v = j.M;
    vse.sig = sgn.sig = ay.sig = s1Ay.sig = S.sig = R8x.sig = R8y.sig = Msg.sig = true;
    hz_phaseF(c, vse, sgn, ay, s1Ay, S, R8x, R8y, Msg);
}

__global__ void k_fee(HzPlan pl, Fr* w, HzErrBuf* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.P.maxFeeTx) return;
    Comp<MCompute> c(w, pl.off_fee[i], err);
    sec_fee(c, pl, i);
}

__global__ void k_hash_head(HzPlan pl, Fr* w, uint8_t* preb, HzErrBuf* err, uint32_t* off_blocks) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    PreBits<MCompute> pre; pre.b = preb;
    Comp<MCompute> hc(w, pl.off_hasher, err);
    W32<MCompute> ha[8];
    uint32_t ob = 0;
    sec_hasher_head(hc, pl, pre, &ob, ha);
    *off_blocks = ob;
    Fr one = fr_one();
    w[0] = one;
}

__global__ void k_sha_chain(HzPlan pl, const uint8_t* preb, uint32_t* chain, uint32_t* msgw) {
    // pack the message words in parallel, then one thread walks the chain
    uint32_t nw = pl.nblocks * 16;
    for (uint32_t t = threadIdx.x; t < nw; t += blockDim.x) {
        uint32_t v = 0;
        for (int k = 0; k < 32; k++) v = (v << 1) | preb[(size_t)t * 32 + k];
        msgw[t] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t st[8];
    for (int j = 0; j < 8; j++) { st[j] = hz_sha_h(j); chain[j] = st[j]; }
    for (uint32_t b = 0; b < pl.nblocks; b++) {
        hz_sha256_compress(st, &msgw[(size_t)b * 16]);
        for (int j = 0; j < 8; j++) chain[(size_t)(b + 1) * 8 + j] = st[j];
    }
}

__global__ void k_sha_blocks(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain, const uint32_t* msgw, const uint32_t* off_blocks) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (int)pl.nblocks) return;
    Comp<MCompute> sh(w, *off_blocks + (uint32_t)b * HZ_SZ_SHA_BLOCK, err);
    W32<MCompute> hin[8], msg[16], outw[8];
    for (int j = 0; j < 8; j++) hin[j].v = chain[(size_t)b * 8 + j];
    for (int t = 0; t < 16; t++) msg[t].v = msgw[(size_t)b * 16 + t];
    sec_sha_block(sh, b, hin, msg, outw);
}

__global__ void k_hash_tail(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    W32<MCompute> dig[8];
    for (int j = 0; j < 8; j++) dig[j].v = chain[(size_t)pl.nblocks * 8 + j];
    Comp<MCompute> tc(w, pl.off_hasher_tail, err);
    Comp<MCompute> shc(w, pl.off_hasher_tail, err);
    VC h = sec_hasher_tail(tc, shc, dig);
    w[1] = h.v;
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
    ok = ok && cudaMalloc(&d->w, 32ull * ph.nvars) == cudaSuccess;
    ok = ok && cudaMalloc(&d->pre, (size_t)ph.nblocks * 512) == cudaSuccess;
    ok = ok && cudaMalloc(&d->chain, (size_t)(ph.nblocks + 1) * 8 * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&d->msgw, (size_t)ph.nblocks * 16 * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&d->jobs, sizeof(HzEddsaJob) * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->err, sizeof(HzErrBuf)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_decode, 4ull * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_rollup, 4ull * P.nTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_fee, 4ull * P.maxFeeTx) == cudaSuccess;
    ok = ok && cudaMalloc(&d->off_blocks, 4) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_decode, off_decode, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_rollup, off_rollup, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_fee, off_fee, 4ull * P.maxFeeTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaStreamCreate(&d->stream) == cudaSuccess;
    ok = ok && cudaEventCreate(&d->ev0) == cudaSuccess && cudaEventCreate(&d->ev1) == cudaSuccess;
    for (int k = 0; k <= HZ_NKERN; k++) ok = ok && cudaEventCreate(&d->kev[k]) == cudaSuccess;
    // the per-thread frames of the circuit templates are large (sibling arrays, state machines)
    ok = ok && cudaDeviceSetLimit(cudaLimitStackSize, 64 * 1024) == cudaSuccess;
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
    cudaFree(d->off_decode); cudaFree(d->off_rollup); cudaFree(d->off_fee); cudaFree(d->off_blocks);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    for (int k = 0; k <= HZ_NKERN; k++) if (d->kev[k]) cudaEventDestroy(d->kev[k]);
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
    const int T = 32;
    CK(cudaEventRecord(d->kev[0], s));
    k_slots<<<(P.nTx + T - 1) / T, T, 0, s>>>(pl, d->w, d->pre, err, d->jobs);
    CK(cudaEventRecord(d->kev[1], s));
    k_eddsa<<<(P.nTx + T - 1) / T, T, 0, s>>>(pl, d->w, err, d->jobs);
    CK(cudaEventRecord(d->kev[2], s));
    k_fee<<<(P.maxFeeTx + T - 1) / T, T, 0, s>>>(pl, d->w, err);
    CK(cudaEventRecord(d->kev[3], s));
    k_hash_head<<<1, 32, 0, s>>>(pl, d->w, d->pre, err, d->off_blocks);
    CK(cudaEventRecord(d->kev[4], s));
    k_sha_chain<<<1, 256, 0, s>>>(pl, d->pre, d->chain, d->msgw);
    CK(cudaEventRecord(d->kev[5], s));
    k_sha_blocks<<<(pl.nblocks + T - 1) / T, T, 0, s>>>(pl, d->w, err, d->chain, d->msgw, d->off_blocks);
    CK(cudaEventRecord(d->kev[6], s));
    k_hash_tail<<<1, 32, 0, s>>>(pl, d->w, err, d->chain);
    CK(cudaEventRecord(d->kev[7], s));
    CK(cudaGetLastError());
    CK(cudaEventRecord(d->ev1, s));
    if (witness_out) CK(cudaMemcpyAsync(witness_out, d->w, 32ull * pl.nvars, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, d->ev0, d->ev1));
    if (launches) *launches = 7;
    if (h2d) *h2d = inputs ? 32ull * pl.n_inputs : 0;
    if (d2h) *d2h = (witness_out ? 32ull * pl.nvars : 0) + sizeof(HzErrBuf);
    return 0;
}

int hz_cuda_kernel_ms(HzCuda* d, float* ms, const char** names, int cap) {
    int n = cap < HZ_NKERN ? cap : HZ_NKERN;
    for (int k = 0; k < n; k++) {
        if (cudaEventElapsedTime(&ms[k], d->kev[k], d->kev[k + 1]) != cudaSuccess) return HZW_E_DEVICE;
        if (names) names[k] = HZ_KERN_NAMES[k];
    }
    return n;
}
