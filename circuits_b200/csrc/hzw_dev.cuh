// Shared by the CUDA translation units of libhzwitness.so (sm_100a): the device state of a context, the
// constants of the witness layout the kernels and their launch code agree on, and the kernel prototypes.
// The kernels are split over several .cu files only so that they compile in parallel:
//   k_slots.cu    k_front, k_fee_front, k_edjob          (per tx slot / fee tx, from the main inputs)
//   k_smtjob.cu   k_smtjob                               (SMT job records ahead of k_front)
//   k_merkle.cu   k_smt, k_smt_levels, k_smt_zero, k_hole_pack, k_tail, the Poseidon-only kernels
//   k_sig.cu      k_fix_consts, k_ed_const, k_eddsa, k_eddsa_tail, k_ed_fill
//   k_ladder.cu   k_ed_ladder, k_ed_hints
//   k_poseidon.cu k_hstate, k_sighash, k_leafhash        (state, message and leaf hashes, transposed)
//   k_hasher.cu   k_pack, k_unpack, k_hash_head, k_sha_chain, k_sha_blocks, k_hash_tail, k_withdraw
//   hzw_run.cu    device memory, streams and the launch order (hz_cuda_*)
#pragma once
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
#define HZ_MAXPROF 40
#define HZ_BLOCK 128                       /* 16 lane groups per block */

struct HzHole { uint32_t cur, nz; };
#define HZ_MAX_GPUS 16
struct HzPeerRecs { uint8_t* p[HZ_MAX_GPUS]; int n; };   // destinations of one rank's exchange record (k_pack)
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
    Fr* hs_in;                 // 4 x (4 nTx + 2 maxFeeTx): inputs of every state hash (k_smtjob, k_fee_front -> k_hstate)
    uint32_t* hs_cur;          // 4 nTx + 2 maxFeeTx: where their Poseidon components start
    Fr* sig_in;                // 6 x nTx: inputs of the message hashes (k_edjob -> k_sighash)
    uint32_t* hsig_cur;        // nTx: where decodeTx[i].hashSig starts
    HzSmtMid* mids;            // 2 nTx + maxFeeTx
    HzLevJob* lev;             // 2 nTx + maxFeeTx: what k_smt leaves for k_smt_levels
    Fr* smst;                  // per processor (nLevels + 1) x 6 state-machine values
    Fr* img00;                 // the 634 entries of SMTHash2(0, 0) and, at [HZ_IMG00_DIGEST], its digest
    uint8_t* zflag;            // per block of k_smt_levels and level: the level hashed (0, 0) in every chain of the block
    // the early walk (hz_early_hash): the last slot and the last fee tx once more, ahead of the batch, on work areas of their own
    HzSmtJob* e_smt; HzSmtMid* e_mids; HzLevJob* e_lev; Fr* e_smst; Fr* e_isExit; Fr* e_hs_in; uint8_t* e_zflag;
    cudaStream_t stream7; cudaEvent_t ev_edjob, ev_hash;
    bool early;                // this batch's hasherInputs were launched by hz_early_hash
    bool xearly;               // sharded context: the exchange record is packed on stream 7 ahead of the merkle levels, phase 2 hashes there
    // end-to-end path: the (0, 0) levels at the bottom of every processor are not copied to the host but rebuilt there
    struct HzHole* holes; struct HzHole* h_holes;   // per job: where its levels start, how many bottom levels are (0, 0) images
    Fr* gl; Fr* h_gl;          // per job: the entries between those images (level headers, switchers), compact
    Fr* h_img00;               // host copy of the image
    uint32_t* bitpack; uint32_t* h_bitpack;   // the SHA block signals packed 32 per word (+ one flag word at the end) and their host copy
    cudaEvent_t ev_bits;
    cudaStream_t stream5, stream6;
    cudaEvent_t ev_pack[2], ev_info[2], ev_sha[2];
    uint64_t last_d2h;
    uint64_t n_launch;         // kernels of ours launched since the start of the current calculate / shard phase 1
    bool no_holes;             // HZW_NO_HOLES (tests / measurement): copy every entry, rebuild nothing on the host
    std::vector<uint32_t>* h_off_rollup;   // host copy of the slot offsets (plus the end of the last slot)
    Fr* isExit;                // nTx
    Fr* isExit_front;          // nTx: k_front's copy
    Fr* hint_cst;              // HZ_CST_MAX: constant inverses of EscalarMulFix (recorded at create)
    Fr* hint_tab;              // HZ_EDTAB_LEN: its window entries
    Fr* hint_img;              // witness image of EscalarMulFix from the recording pass (constant entries are copied from it)
    Fr* hint_scratch;          // per lane group of k_eddsa: work area of the hint generators
    Fr* ed_tab;                // the window entries in the cached twisted-Edwards form the ladder adds (k_ed_tab, once)
    Fr* ed_pts;                // per signature: the points of every ladder step (k_ed_ladder -> k_ed_conv)
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
    cudaEvent_t ev0, ev1, ev_front, ev_eddsa, ev_lev, ev_zero, ev_fee, ev_jobs, ev_heads;
    cudaEvent_t ev_xchg;       // this rank's record has been stored into every peer (hz_cuda_multi_phase1)
    cudaEvent_t kev[HZ_NKERN + 4];
    // layout map (a circom .sym applied when the witness leaves the device): compiler's index -> own index, and the
    // witness gathered into the compiler's layout
    uint32_t* map_src; Fr* w_out; uint64_t n_out;
    // profile mode (hz_cuda_profile_step): every launch of a batch on ONE stream, an event in front of each
    bool prof; int n_prof; cudaEvent_t pev[HZ_MAXPROF + 1]; const char* pname[HZ_MAXPROF];
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

// jobs of the shard: 2 per own slot, then 1 per own fee tx
__device__ __forceinline__ int hz_smt_job_index(const HzPlan& pl, int q, int s0, int ns, int f0) {
    return q < 2 * ns ? 2 * s0 + q : 2 * pl.P.nTx + f0 + (q - 2 * ns);
}
#define HZ_IMG00_DIGEST 640
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
#define HZ_FILL_PER_SIG (HZ_FILL_CHUNKS + 83)
__host__ __device__ static inline uint32_t hz_rec_slot_bytes(const HzParams& P) { return 624u + 2u * (uint32_t)P.nLevels + 48u; }
__host__ __device__ static inline uint32_t hz_rec_vals_off(const HzParams& P, int per) { return ((uint32_t)per * hz_rec_slot_bytes(P) + 3u) & ~3u; }
__host__ __device__ static inline uint32_t hz_rec_bytes(const HzParams& P, int per) { return (hz_rec_vals_off(P, per) + 3 * 32 + 3 * 4 + 15u) & ~15u; }
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
#define HZ_SHA_THREADS 128
#define HZ_PT_CHAINS 32

// ---- kernels ---------------------------------------------------------------------------------------------
__global__ void k_poseidon_pairs(const Fr* in, Fr* sig, Fr* digests, int npairs);
__global__ void k_poseidon_t3(const Fr* in, Fr* sig, Fr* digests, int nperms);
__global__ void k_front(HzPlan pl, Fr* w, uint8_t* preb, HzErrBuf* err, HzEddsaJob* jobs, HzSmtJob* smt, Fr* isExit, int s0, int ns);
__global__ void k_smtjob(HzPlan pl, Fr* w, HzSmtJob* smt, Fr* isExit, const uint32_t* off_jk, Fr* hs_in, int s0, int ns);
__global__ void k_hstate(Fr* w, const Fr* hs_in, const uint32_t* hs_cur, HzSmtJob* smt, int nTx, int s0, int ns, int f0, int nf);
__global__ void k_leafhash(HzPlan pl, Fr* w, const HzSmtJob* smt, HzLevJob* lev, int s0, int ns, int f0, int nf);
__global__ void k_sighash(Fr* w, const Fr* sig_in, const uint32_t* hsig_cur, HzEddsaJob* jobs, int s0, int ns);
__global__ void k_edjob(HzPlan pl, const Fr* w, HzEddsaJob* jobs, const uint32_t* off_eddsa, Fr* sig_in, int s0, int ns);
__global__ void k_fee_front(HzPlan pl, Fr* w, HzErrBuf* err, HzSmtJob* smt, Fr* hs_in, int f0, int nf);
__global__ void k_smt(HzPlan pl, Fr* w, HzErrBuf* err, const HzSmtJob* smt, HzSmtMid* mids, HzLevJob* lev, Fr* smst, int s0, int ns, int f0, int nf);
__global__ void k_smt_levels(HzPlan pl, Fr* w, const HzLevJob* jobs, const Fr* smst, HzSmtMid* mids, const Fr* img00, uint8_t* zflag, int s0, int ns, int f0, int nf);
__global__ void k_hole_pack(HzPlan pl, const Fr* w, const HzLevJob* jobs, const uint8_t* zflag, HzHole* holes, Fr* gl, int s0, int ns, int f0, int nf);
__global__ void k_smt_zero(HzPlan pl, Fr* w, const HzLevJob* jobs, const Fr* img00, const uint8_t* zflag, int s0, int ns, int f0, int nf);
__global__ void k_fix_consts(Fr* wtmp, Fr* cst, Fr* tab, uint32_t* counts);
__global__ void k_ed_const(HzPlan pl, Fr* w, const uint32_t* off_eddsa, const Fr* img, uint32_t fix_size, int s0, int ns);
__global__ void k_eddsa(HzPlan pl, Fr* w, HzErrBuf* err, const HzEddsaJob* jobs, const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch, uint32_t win_size, int s0, int ns);
__global__ void k_eddsa_tail(Fr* w, HzErrBuf* err, const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch, uint32_t win_size, uint32_t fix_size, int ns);
__global__ void k_ed_ladder(Fr* scratch, const Fr* edtab, Fr* pts, int ns);
__global__ void k_ed_conv(Fr* scratch, const Fr* tab, const Fr* pts, int ns);
__global__ void k_ed_tab(const Fr* tab, Fr* edtab, int npoints, int init_point);
__global__ void k_ed_hints(Fr* scratch, int ns);
__global__ void k_ed_fill(Fr* w, HzErrBuf* err, const Fr* tab, const Fr* scratch, uint32_t win_stride, int ns);
__global__ void k_tail(HzPlan pl, Fr* w, HzErrBuf* err, HzErrBuf* ebuf, const HzSmtMid* mids, const Fr* isExit, const Fr* isExit_front,
                       const HzEddsaJob* jobs, const HzEddsaJob* jobs_front, const HzSmtJob* smt, const HzSmtJob* smt_front, int s0, int ns, int f0, int nf);
__global__ void k_pack(HzPlan pl, const Fr* w, const uint8_t* pre, HzPeerRecs dst, int s0, int ns, int per, int own_last_slot, int own_last_fee);
__global__ void k_unpack(HzPlan pl, Fr* w, uint8_t* pre, const uint8_t* recs, int world, int per, uint32_t rec_bytes);
__global__ void k_hash_head(HzPlan pl, Fr* w, uint8_t* pre, HzErrBuf* err, uint32_t* off_blocks);
__global__ void k_bitpack(const Fr* w, uint64_t count, uint32_t* packed, uint32_t* notbit);
__global__ void k_sha_chain(HzPlan pl, const uint8_t* preb, uint32_t* chain, uint32_t* msgw, uint32_t* wk, uint32_t b_begin, uint32_t b_end);
__global__ void k_sha_blocks(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain, const uint32_t* msgw, const uint32_t* off_blocks, uint32_t b0);
__global__ void k_hash_tail(HzPlan pl, Fr* w, HzErrBuf* err, const uint32_t* chain);
__global__ void k_withdraw(Fr* w, uint32_t nvars, int nLevels, int n, HzErrBuf* err);
__global__ void k_rolluptx(Fr* w, uint32_t nvars, int nLevels, int maxFeeTx, int n, HzErrBuf* err);
__device__ __forceinline__ void hz_head_err(HzErrBuf* err, uint32_t cur, uint32_t line, const Fr& lhs) {
    if (!err) return;
    uint32_t k = atomicAdd(&err->count, 1u);
    if (k < HZ_MAX_ERR) { err->rec[k].cur = cur; err->rec[k].line = line; err->rec[k].lhs = lhs; err->rec[k].rhs = fr_zero(); }
}
