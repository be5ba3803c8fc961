// libhzwitness.so, CUDA half (sm_100a): device memory, streams and the launch order of the section kernels
// (kernels: k_*.cu; shared declarations: hzw_dev.cuh).
#include "hzw_dev.cuh"

static void hz_set_ranges(HzCuda* d, int rank, int world);
// called in front of every launch of a batch: counts it and, in profile mode, brackets it with an event
static inline void hz_mark(HzCuda* d, const char* name, cudaStream_t st) {
    d->n_launch++;
    if (d->prof && d->n_prof < HZ_MAXPROF) { cudaEventRecord(d->pev[d->n_prof], st); d->pname[d->n_prof++] = name; }
}

HzCuda* hz_cuda_create(int device, const HzPlan& ph, const uint32_t* off_decode, const uint32_t* off_rollup,
                       const uint32_t* off_fee, const uint32_t* off_eddsa, const uint32_t* off_jk,
                       const uint32_t* off_hs, const uint32_t* off_hsig, int* status) {
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
    {
        const size_t nhs = (size_t)4 * P.nTx + 2 * P.maxFeeTx;
        ok = ok && cudaMalloc(&d->hs_in, sizeof(Fr) * 4 * nhs) == cudaSuccess && cudaMalloc(&d->hs_cur, 4 * nhs) == cudaSuccess;
        ok = ok && cudaMemcpy(d->hs_cur, off_hs, 4 * nhs, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMalloc(&d->sig_in, sizeof(Fr) * 6 * (size_t)P.nTx) == cudaSuccess && cudaMalloc(&d->hsig_cur, 4ull * P.nTx) == cudaSuccess;
        ok = ok && cudaMemcpy(d->hsig_cur, off_hsig, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&d->isExit_front, sizeof(Fr) * P.nTx) == cudaSuccess;
    ok = ok && cudaMemcpy(d->off_eddsa, off_eddsa, 4ull * P.nTx, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMalloc(&d->smt, sizeof(HzSmtJob) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->mids, sizeof(HzSmtMid) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->lev, sizeof(HzLevJob) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->smst, sizeof(Fr) * 6 * (size_t)(P.nLevels + 1) * (2 * P.nTx + P.maxFeeTx)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->isExit, sizeof(Fr) * P.nTx) == cudaSuccess;
    {
        const size_t nj = (size_t)2 * P.nTx + P.maxFeeTx, nhs = (size_t)4 * P.nTx + 2 * P.maxFeeTx;
        ok = ok && cudaMalloc(&d->e_smt, sizeof(HzSmtJob) * nj) == cudaSuccess && cudaMalloc(&d->e_mids, sizeof(HzSmtMid) * nj) == cudaSuccess;
        ok = ok && cudaMalloc(&d->e_lev, sizeof(HzLevJob) * nj) == cudaSuccess && cudaMalloc(&d->e_smst, sizeof(Fr) * 6 * (size_t)(P.nLevels + 1) * nj) == cudaSuccess;
        ok = ok && cudaMalloc(&d->e_isExit, sizeof(Fr) * P.nTx) == cudaSuccess && cudaMalloc(&d->e_hs_in, sizeof(Fr) * 4 * nhs) == cudaSuccess;
        ok = ok && cudaMalloc(&d->e_zflag, (size_t)2 * (P.nLevels + 1)) == cudaSuccess;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);         // its few small launches should not queue behind the batch's
        ok = ok && cudaStreamCreateWithPriority(&d->stream7, cudaStreamDefault, hi) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&d->ev_edjob, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_hash, cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&d->img00, sizeof(Fr) * (HZ_IMG00_DIGEST + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->zflag, (size_t)((2 * P.nTx + P.maxFeeTx + 15) / 16 + 1) * (P.nLevels + 1)) == cudaSuccess;
    {
        const size_t nj = (size_t)2 * P.nTx + P.maxFeeTx, gs = HZ_GL_STRIDE(P.nLevels + 1);
        ok = ok && cudaMalloc(&d->holes, sizeof(HzHole) * nj) == cudaSuccess && cudaMallocHost(&d->h_holes, sizeof(HzHole) * nj) == cudaSuccess;
        ok = ok && cudaMalloc(&d->gl, sizeof(Fr) * gs * nj) == cudaSuccess && cudaMallocHost(&d->h_gl, sizeof(Fr) * gs * nj) == cudaSuccess;
        ok = ok && cudaMallocHost(&d->h_img00, sizeof(Fr) * HZ_HS) == cudaSuccess;
        const size_t nbw = (size_t)ph.nblocks * HZ_SZ_SHA_BLOCK / 32 + 1;
        ok = ok && cudaMalloc(&d->bitpack, 4 * nbw) == cudaSuccess && cudaMallocHost(&d->h_bitpack, 4 * nbw) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&d->ev_bits, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamCreate(&d->stream5) == cudaSuccess && cudaStreamCreate(&d->stream6) == cudaSuccess;
        for (int k = 0; k < 2; k++) ok = ok && cudaEventCreateWithFlags(&d->ev_pack[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_info[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_sha[k], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&d->hint_cst, sizeof(Fr) * HZ_CST_MAX) == cudaSuccess;
    ok = ok && cudaMalloc(&d->hint_tab, sizeof(Fr) * HZ_EDTAB_LEN) == cudaSuccess;
    ok = ok && cudaMalloc(&d->hint_scratch, sizeof(Fr) * HZ_HINT_SCRATCH * (size_t)hz_grid(P.nTx) * (HZ_BLOCK / HZ_G)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->ed_tab, sizeof(Fr) * (4 * (HZ_EDTAB_LEN / 2) + 4)) == cudaSuccess;
    ok = ok && cudaMalloc(&d->ed_pts, sizeof(Fr) * HZ_PTS * (size_t)P.nTx) == cudaSuccess;
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
    ok = ok && cudaEventCreateWithFlags(&d->ev_xchg, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&d->ev_jobs, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&d->ev_heads, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < HZ_NKERN + 4; k++) ok = ok && cudaEventCreate(&d->kev[k]) == cudaSuccess;
    for (int k = 0; k <= HZ_MAXPROF; k++) ok = ok && cudaEventCreate(&d->pev[k]) == cudaSuccess;
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
        cudaFuncSetAttribute(k_hstate, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_leafhash, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_sighash, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_ed_ladder, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_ed_conv, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_eddsa_tail, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    }
    // the per-thread frames of the circuit templates are large (sibling arrays, state machines)
    ok = ok && cudaDeviceSetLimit(cudaLimitStackSize, 80 * 1024) == cudaSuccess;
    if (ok) {
        // the entries and the digest of SMTHash2(0, 0), once (k_poseidon_t3 on one permutation of two zero inputs)
        Fr* zin = nullptr;
        ok = cudaMalloc(&zin, sizeof(Fr) * 2) == cudaSuccess && cudaMemset(zin, 0, sizeof(Fr) * 2) == cudaSuccess;
        if (ok) {
            d->n_launch++, k_poseidon_t3<<<1, 96, 0, d->stream>>>(zin, d->img00, d->img00 + HZ_IMG00_DIGEST, 1);
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
            d->n_launch++, k_fix_consts<<<1, 32, 0, d->stream>>>(wtmp, d->hint_cst, d->hint_tab, counts);
            ok = cudaMemcpyAsync(hc, counts, 20, cudaMemcpyDeviceToHost, d->stream) == cudaSuccess &&
                 cudaStreamSynchronize(d->stream) == cudaSuccess;
            if (ok && (hc[0] > HZ_CST_MAX || hc[1] != HZ_EDTAB_LEN)) {
                fprintf(stderr, "hzw: hint table sizes %u / %u do not match the templates\n", hc[0], hc[1]);
                ok = false;
            }
            d->fix_size = hc[4]; d->win_size = hc[3]; d->win_stride = hc[3] / 83;
            // the window entries once more, in the form the ladder walk adds (k_ladder.cu)
            d->n_launch++, k_ed_tab<<<(HZ_EDTAB_LEN / 2 + 63) / 64, 64, 0, d->stream>>>(d->hint_tab, d->ed_tab, HZ_EDTAB_LEN / 2, 83 * 8);
            ok = ok && cudaStreamSynchronize(d->stream) == cudaSuccess;
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
    cudaFree(d->hs_in); cudaFree(d->hs_cur); cudaFree(d->sig_in); cudaFree(d->hsig_cur);
    cudaFree(d->smt); cudaFree(d->mids); cudaFree(d->isExit); cudaFree(d->lev); cudaFree(d->smst); cudaFree(d->img00); cudaFree(d->zflag); cudaFree(d->holes); cudaFree(d->gl); cudaFreeHost(d->h_holes); cudaFreeHost(d->h_gl); cudaFreeHost(d->h_img00); cudaFree(d->bitpack); cudaFreeHost(d->h_bitpack); if (d->ev_bits) cudaEventDestroy(d->ev_bits);
    cudaFree(d->hint_cst); cudaFree(d->hint_tab); cudaFree(d->hint_img); cudaFree(d->hint_scratch); cudaFree(d->ed_tab); cudaFree(d->ed_pts); cudaFree(d->recs);
    cudaFree(d->off_decode); cudaFree(d->off_rollup); cudaFree(d->off_fee); cudaFree(d->off_blocks);
    cudaFree(d->map_src); cudaFree(d->w_out);
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
    if (d->ev_xchg) cudaEventDestroy(d->ev_xchg);
    if (d->ev_jobs) cudaEventDestroy(d->ev_jobs);
    if (d->ev_heads) cudaEventDestroy(d->ev_heads);
    for (int k = 0; k < HZ_NKERN + 4; k++) if (d->kev[k]) cudaEventDestroy(d->kev[k]);
    cudaFree(d->e_smt); cudaFree(d->e_mids); cudaFree(d->e_lev); cudaFree(d->e_smst); cudaFree(d->e_isExit); cudaFree(d->e_hs_in); cudaFree(d->e_zflag);
    if (d->stream7) cudaStreamDestroy(d->stream7);
    if (d->ev_edjob) cudaEventDestroy(d->ev_edjob);
    if (d->ev_hash) cudaEventDestroy(d->ev_hash);
    for (int k = 0; k <= HZ_MAXPROF; k++) if (d->pev[k]) cudaEventDestroy(d->pev[k]);
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
uint64_t hz_cuda_launches(const HzCuda* d) { return d->n_launch; }

static int hz_early_hash(HzCuda* d, int sanity);
static int hz_early_exchange(HzCuda* d);
// phase 1: inputs to the device, then everything per slot / fee tx of the shard
static int hz_phase1(HzCuda* d, const uint8_t* inputs, int sanity) {
    const HzPlan& pl = d->plan;
    cudaStream_t s = d->stream;
    // profile mode (hzw_profile_step): every launch on one stream with an event in front of it
    const cudaStream_t s2 = d->prof ? s : d->stream2, s3 = d->prof ? s : d->stream3, s4 = d->prof ? s : d->stream4, s5 = d->prof ? s : d->stream5;
    HzErrBuf* err = sanity ? d->err : nullptr;
    CK(cudaMemsetAsync(d->err, 0, sizeof(HzErrBuf), s));
    CK(cudaMemsetAsync(d->pre, 0, (size_t)pl.nblocks * 512, s));
    if (inputs) CK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(d->w) + 64, inputs, 32ull * pl.n_inputs, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(d->ev0, s));
    const int s0 = d->s0, ns = d->ns, f0 = d->f0, nf = d->nf;
    // the merkle chain is the longest of the batch: its first kernel is launched first
    CK(cudaEventRecord(d->kev[0], s));
    if (ns > 0) hz_mark(d, "k_smtjob", s), k_smtjob<<<(unsigned)((ns + 31) / 32), 32, 0, s>>>(pl, d->w, d->smt, d->isExit, d->off_jk, d->hs_in, s0, ns);
    // the signatures are the longest chain of a slot and need little from the rest of it: their jobs come
    // straight from the inputs (k_edjob) and they run on the second stream from the start, beside k_front
    // and the merkle levels -- no kernel fills the machine alone
    CK(cudaStreamWaitEvent(s2, d->ev0, 0));
    CK(cudaEventRecord(d->kev[9], s2));
    if (ns > 0) {
        hz_mark(d, "k_edjob", s2), k_edjob<<<hz_grid(ns), HZ_BLOCK, 0, s2>>>(pl, d->w, d->jobs, d->off_eddsa, d->sig_in, s0, ns);
        hz_mark(d, "k_sighash", s2), k_sighash<<<(unsigned)((ns + 31) / 32), 224, 0, s2>>>(d->w, d->sig_in, d->hsig_cur, d->jobs, s0, ns);
        CK(cudaEventRecord(d->ev_edjob, s2));
        hz_mark(d, "k_ed_const", s2), k_ed_const<<<592, 256, 0, s2>>>(pl, d->w, d->off_eddsa, d->hint_img, d->fix_size, s0, ns);
            hz_mark(d, "k_eddsa", s2), k_eddsa<<<hz_grid(ns), HZ_BLOCK, 0, s2>>>(pl, d->w, err, d->jobs, d->hint_cst, d->hint_tab, d->hint_img, d->hint_scratch, d->win_size, s0, ns);
            hz_mark(d, "k_ed_ladder", s2), k_ed_ladder<<<(unsigned)((ns + 7) / 8), 96, 0, s2>>>(d->hint_scratch, d->ed_tab, d->ed_pts, ns);
            hz_mark(d, "k_ed_conv", s2), k_ed_conv<<<(unsigned)(((size_t)ns * HZ_LAD_TOT + 127) / 128), 128, 0, s2>>>(d->hint_scratch, d->hint_tab, d->ed_pts, ns);
            hz_mark(d, "k_ed_hints", s2), k_ed_hints<<<(unsigned)((ns + 3) / 4), 128, 0, s2>>>(d->hint_scratch, ns);
            hz_mark(d, "k_eddsa_tail", s2), k_eddsa_tail<<<(unsigned)((ns + 7) / 8), HZ_BLOCK, 0, s2>>>(d->w, err, d->hint_cst, d->hint_tab, d->hint_img, d->hint_scratch, d->win_size, d->fix_size, ns);
        hz_mark(d, "k_ed_fill", s2), k_ed_fill<<<(unsigned)((HZ_FILL_PER_SIG * ns + 127) / 128), 128, 0, s2>>>(d->w, err, d->hint_tab, d->hint_scratch, d->win_stride, ns);
    }
    CK(cudaEventRecord(d->kev[10], s2));
    CK(cudaEventRecord(d->ev_eddsa, s2));
    // the fee transactions need nothing but inputs either: third stream, joined before the merkle kernels
    CK(cudaStreamWaitEvent(s3, d->ev0, 0));
    if (nf > 0) hz_mark(d, "k_fee_front", s3), k_fee_front<<<hz_grid(nf), HZ_BLOCK, 0, s3>>>(pl, d->w, err, d->smt + 2 * pl.P.nTx, d->hs_in, f0, nf);
    CK(cudaEventRecord(d->ev_fee, s3));
    // k_front on a stream of its own: the merkle kernels take their jobs from k_smtjob
    CK(cudaStreamWaitEvent(s4, d->ev0, 0));
    CK(cudaEventRecord(d->kev[11], s4));
    if (ns > 0) hz_mark(d, "k_front", s4), k_front<<<(unsigned)((ns + 31) / 32), 32, 0, s4>>>(pl, d->w, d->pre, err, d->jobs_front, d->smt_front, d->isExit_front, s0, ns);
    CK(cudaEventRecord(d->kev[12], s4));
    CK(cudaEventRecord(d->ev_front, s4));
    // (pays when the SHA walk is long: 766 blocks at 2048 slots, 9.4 -> 9.1 ms; with 150 blocks at 344 slots the extra
    // launches cost more than the walk they hide, 4.5 -> 4.9 ms)
    const bool forced = getenv("HZW_EARLY_HASH") != nullptr;          // (tests: the early paths at small sizes)
    const bool long_walk = (pl.nblocks >= 512 || forced) && pl.P.nTx >= 2 && pl.P.maxFeeTx >= 1 && !getenv("HZW_NO_EARLY_HASH");
    d->early = d->world == 1 && !d->prof && long_walk;
    if (d->early) { int st = hz_early_hash(d, sanity); if (st) return st; }
    d->xearly = d->world > 1 && !d->prof && long_walk;
    if (d->xearly) { int st = hz_early_exchange(d); if (st) return st; }
    CK(cudaStreamWaitEvent(s, d->ev_fee, 0));
    // the state hashes of the slots and fee txs, transposed: their digests complete the job records
    // the processor heads need the jobs but not the hashes: they run on stream 3 beside the two hash kernels
    CK(cudaEventRecord(d->ev_jobs, s));
    CK(cudaStreamWaitEvent(s3, d->ev_jobs, 0));
    if (2 * ns + nf > 0) hz_mark(d, "k_smt", s3), k_smt<<<(unsigned)((2 * ns + nf + 31) / 32), 32, 0, s3>>>(pl, d->w, err, d->smt, d->mids, d->lev, d->smst, s0, ns, f0, nf);
    CK(cudaEventRecord(d->ev_heads, s3));
    if (4 * ns + 2 * nf > 0) hz_mark(d, "k_hstate", s), k_hstate<<<(unsigned)((4 * ns + 2 * nf + 31) / 32), 160, 0, s>>>(d->w, d->hs_in, d->hs_cur, d->smt, pl.P.nTx, s0, ns, f0, nf);
    CK(cudaEventRecord(d->kev[1], s));
    CK(cudaEventRecord(d->kev[2], s));
    if (2 * ns + nf > 0) {
        hz_mark(d, "k_leafhash", s), k_leafhash<<<(unsigned)((2 * (2 * ns + nf) + 31) / 32), 128, 0, s>>>(pl, d->w, d->smt, d->lev, s0, ns, f0, nf);
        CK(cudaStreamWaitEvent(s, d->ev_heads, 0));
        const unsigned lb = (unsigned)((2 * ns + nf + 15) / 16);
        hz_mark(d, "k_smt_levels", s), k_smt_levels<<<lb, 96, 0, s>>>(pl, d->w, d->lev, d->smst, d->mids, d->img00, d->zflag, s0, ns, f0, nf);
        // the skipped (0, 0) levels are filled in beside the tail and the hash of the inputs
        CK(cudaEventRecord(d->ev_lev, s));
        CK(cudaStreamWaitEvent(s3, d->ev_lev, 0));
        hz_mark(d, "k_smt_zero", s3), k_smt_zero<<<dim3(lb, (unsigned)(pl.P.nLevels + 1)), 256, 0, s3>>>(pl, d->w, d->lev, d->img00, d->zflag, s0, ns, f0, nf);
        if (d->world > 1 && !d->no_holes) {
            // a shard's pieces are copied around the image levels too (hz_cuda_shard_phase2): their records go to the host now
            const uint32_t gs = HZ_GL_STRIDE(pl.P.nLevels + 1);
            hz_mark(d, "k_hole_pack", s), k_hole_pack<<<(unsigned)(2 * ns + nf), 128, 0, s>>>(pl, d->w, d->lev, d->zflag, d->holes, d->gl, s0, ns, f0, nf);
            CK(cudaEventRecord(d->ev_pack[0], s));
            CK(cudaStreamWaitEvent(s5, d->ev_pack[0], 0));
            if (ns > 0) {
                CK(cudaMemcpyAsync(d->h_holes + 2 * s0, d->holes + 2 * s0, sizeof(HzHole) * 2 * ns, cudaMemcpyDeviceToHost, s5));
                CK(cudaMemcpyAsync(d->h_gl + (size_t)2 * s0 * gs, d->gl + (size_t)2 * s0 * gs, sizeof(Fr) * gs * 2 * ns, cudaMemcpyDeviceToHost, s5));
            }
            if (nf > 0) {
                const size_t j = (size_t)2 * pl.P.nTx + f0;
                CK(cudaMemcpyAsync(d->h_holes + j, d->holes + j, sizeof(HzHole) * nf, cudaMemcpyDeviceToHost, s5));
                CK(cudaMemcpyAsync(d->h_gl + j * gs, d->gl + j * gs, sizeof(Fr) * gs * nf, cudaMemcpyDeviceToHost, s5));
            }
            CK(cudaEventRecord(d->ev_info[0], s5));
        }
    }
    CK(cudaEventRecord(d->ev_zero, s3));
    CK(cudaEventRecord(d->kev[3], s));
    CK(cudaStreamWaitEvent(s, d->ev_front, 0));       // the tail compares the jobs, the hash of the inputs reads k_front's pre-image bits
    if (ns + nf > 0) hz_mark(d, "k_tail", s), k_tail<<<hz_grid(ns + nf), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->err, d->mids, d->isExit, d->isExit_front, d->jobs, d->jobs_front, d->smt, d->smt_front, s0, ns, f0, nf);
    CK(cudaEventRecord(d->kev[13], s));
    return 0;
}
// hasherInputs on stream `s`: every shard walks the whole SHA chain (it is serial and short) and expands its own range
// of blocks, the first half of them on `s5` while the second half of the chain is walked
static int hz_launch_hash(HzCuda* d, int sanity, cudaStream_t s, cudaStream_t s5) {
    const HzPlan& pl = d->plan;
    HzErrBuf* err = sanity ? d->err : nullptr;
    CK(cudaEventRecord(d->kev[4], s));
    hz_mark(d, "k_hash_head", s), k_hash_head<<<148, 256, 0, s>>>(pl, d->w, d->pre, d->rank == 0 ? err : nullptr, d->off_blocks);
    CK(cudaEventRecord(d->kev[5], s));
    const uint32_t mid = pl.nblocks / 2, bend = d->b0 + d->nb;
    hz_mark(d, "k_sha_chain", s), k_sha_chain<<<1, 256, 0, s>>>(pl, d->pre, d->chain, d->msgw, d->shawk, 0, mid);
    const uint32_t n1 = d->b0 < mid ? (bend < mid ? bend : mid) - d->b0 : 0;       // own blocks below mid
    if (n1 > 0) {
        CK(cudaEventRecord(d->ev_sha[0], s));
        CK(cudaStreamWaitEvent(s5, d->ev_sha[0], 0));
        hz_mark(d, "k_sha_blocks", s5), k_sha_blocks<<<n1, HZ_SHA_THREADS, 0, s5>>>(pl, d->w, err, d->chain, d->msgw, d->off_blocks, d->b0);
        CK(cudaEventRecord(d->ev_sha[1], s5));
    }
    hz_mark(d, "k_sha_chain", s), k_sha_chain<<<1, 256, 0, s>>>(pl, d->pre, d->chain, d->msgw, d->shawk, mid, pl.nblocks);
    CK(cudaEventRecord(d->kev[6], s));
    if (d->nb > n1) hz_mark(d, "k_sha_blocks", s), k_sha_blocks<<<d->nb - n1, HZ_SHA_THREADS, 0, s>>>(pl, d->w, err, d->chain, d->msgw, d->off_blocks, d->b0 + n1);
    if (n1 > 0) CK(cudaStreamWaitEvent(s, d->ev_sha[1], 0));
    CK(cudaEventRecord(d->kev[7], s));
    hz_mark(d, "k_hash_tail", s), k_hash_tail<<<1, 32, 0, s>>>(pl, d->w, d->rank == 0 ? err : nullptr, d->chain);
    CK(cudaEventRecord(d->kev[8], s));
    return 0;
}
// The hash of the inputs is a serial walk (766 blocks at 2048 slots) that needs, besides the pre-image bits k_front
// leaves, three values from the END of the batch: newLastIdx and newExitRoot of the last slot, newStateRoot of the last
// fee tx (src/rollup-main.circom:437-441).  Every root before those is a main input (imStateRoot, imExitRoot,
// imInitStateRootFee, imStateRootFee), so the last slot and the last fee tx can be walked on their own: this puts them
// through the same kernels ahead of the batch on a high-priority stream and work areas of their own (the batch
// evaluates them again with everything else and stores the same entries), and the SHA walk runs beside the merkle
// levels of the batch instead of behind them.
// the walk itself: the last slot (n_s = 1) and / or the last fee tx (n_f = 1) on stream 7, up to their tails
static int hz_early_walk(HzCuda* d, int n_s, int n_f) {
    const HzPlan& pl = d->plan;
    const cudaStream_t s7 = d->stream7;
    const int ls = pl.P.nTx - 1, lf = pl.P.maxFeeTx - 1;
    if (n_s) d->n_launch++, k_smtjob<<<1, 32, 0, s7>>>(pl, d->w, d->e_smt, d->e_isExit, d->off_jk, d->e_hs_in, ls, 1);
    if (n_f) d->n_launch++, k_fee_front<<<hz_grid(1), HZ_BLOCK, 0, s7>>>(pl, d->w, nullptr, d->e_smt + 2 * pl.P.nTx, d->e_hs_in, lf, 1);
    d->n_launch++, k_smt<<<1, 32, 0, s7>>>(pl, d->w, nullptr, d->e_smt, d->e_mids, d->e_lev, d->e_smst, ls, n_s, lf, n_f);
    d->n_launch++, k_hstate<<<1, 160, 0, s7>>>(d->w, d->e_hs_in, d->hs_cur, d->e_smt, pl.P.nTx, ls, n_s, lf, n_f);
    d->n_launch++, k_leafhash<<<1, 128, 0, s7>>>(pl, d->w, d->e_smt, d->e_lev, ls, n_s, lf, n_f);
    d->n_launch++, k_smt_levels<<<1, 96, 0, s7>>>(pl, d->w, d->e_lev, d->e_smst, d->e_mids, d->img00, d->e_zflag, ls, n_s, lf, n_f);
    CK(cudaStreamWaitEvent(s7, d->ev_front, 0));
    CK(cudaStreamWaitEvent(s7, d->ev_edjob, 0));
    d->n_launch++, k_tail<<<hz_grid(n_s + n_f), HZ_BLOCK, 0, s7>>>(pl, d->w, nullptr, d->err, d->e_mids, d->e_isExit, d->isExit_front, d->jobs, d->jobs_front, d->e_smt, d->smt_front, ls, n_s, lf, n_f);
    return 0;
}
static int hz_early_hash(HzCuda* d, int sanity) {
    CK(cudaStreamWaitEvent(d->stream7, d->ev0, 0));
    int st = hz_early_walk(d, 1, 1);
    if (st) return st;
    st = hz_launch_hash(d, sanity, d->stream7, d->stream5);
    if (st) return st;
    CK(cudaEventRecord(d->ev_hash, d->stream7));
    return 0;
}
// A shard does the same for the exchange: its record (pre-image bits of its slots, and the closing values if it owns
// the last slot / fee tx) is complete after k_front and the early walk, long before its merkle levels are.  The record
// is packed on stream 7 then, the collective and the SHA walk of phase 2 run beside the levels, and the main stream only
// joins at the end.
static int hz_early_exchange(HzCuda* d) {
    const HzParams& P = d->plan.P;
    const int n_s = d->ns > 0 && d->s0 + d->ns == P.nTx, n_f = d->nf > 0 && d->f0 + d->nf == P.maxFeeTx;
    CK(cudaStreamWaitEvent(d->stream7, d->ev0, 0));
    if (n_s + n_f > 0) return hz_early_walk(d, n_s, n_f);
    CK(cudaStreamWaitEvent(d->stream7, d->ev_front, 0));
    return 0;
}
// phase 2: hasherInputs (unless the early walk has launched it already), then everything joins the main stream
static int hz_phase2(HzCuda* d, int sanity) {
    cudaStream_t s = d->stream;
    if (d->xearly) {
        int st = hz_launch_hash(d, sanity, d->stream7, d->stream5);
        if (st) return st;
        CK(cudaEventRecord(d->ev_hash, d->stream7));
    }
    if (d->early || d->xearly) CK(cudaStreamWaitEvent(s, d->ev_hash, 0));
    else {
        int st = hz_launch_hash(d, sanity, s, d->prof ? s : d->stream5);
        if (st) return st;
    }
    if (d->prof) CK(cudaEventRecord(d->pev[d->n_prof], s));
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
void hz_host_expand_bits(uint8_t* out, const uint32_t* words, size_t nwords);
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
    d->early = false; d->xearly = false;
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
    d->n_launch++, k_edjob<<<hz_grid(P.nTx), HZ_BLOCK, 0, s2>>>(pl, d->w, d->jobs, d->off_eddsa, d->sig_in, 0, P.nTx);
    d->n_launch++, k_sighash<<<(unsigned)((P.nTx + 31) / 32), 224, 0, s2>>>(d->w, d->sig_in, d->hsig_cur, d->jobs, 0, P.nTx);
    CK(cudaStreamWaitEvent(d->stream4, d->ev0, 0));
    d->n_launch++, k_front<<<(unsigned)((P.nTx + 31) / 32), 32, 0, d->stream4>>>(pl, d->w, d->pre, err, d->jobs_front, d->smt_front, d->isExit_front, 0, P.nTx);
    CK(cudaEventRecord(d->ev_front, d->stream4));
    d->n_launch++, k_smtjob<<<(unsigned)((P.nTx + 31) / 32), 32, 0, s>>>(pl, d->w, d->smt, d->isExit, d->off_jk, d->hs_in, 0, P.nTx);
    CK(cudaStreamWaitEvent(s3, d->ev_front, 0));
    CK(cudaStreamWaitEvent(s6, d->ev_front, 0));
    const int F = P.maxFeeTx, h0 = P.nTx / 2;
    d->n_launch++, k_fee_front<<<hz_grid(F), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->smt + 2 * P.nTx, d->hs_in, 0, F);
    d->n_launch++, k_hstate<<<(unsigned)((4 * P.nTx + 2 * F + 31) / 32), 160, 0, s>>>(d->w, d->hs_in, d->hs_cur, d->smt, P.nTx, 0, P.nTx, 0, F);
    // w[0], w[1] are written by the hasher at the end: entries [2, decodeTx[0]) and the decode ranges go now
    CK(piece(2, d->h_off_rollup0));
    for (int h = 0; h < 2; h++) {
        const int s0 = h == 0 ? 0 : h0, ns = h == 0 ? h0 : P.nTx - h0, nf = h == 1 ? F : 0;
        if (ns > 0) {
            d->n_launch++, k_ed_const<<<592, 256, 0, s2>>>(pl, d->w, d->off_eddsa, d->hint_img, d->fix_size, s0, ns);
            d->n_launch++, k_eddsa<<<hz_grid(ns), HZ_BLOCK, 0, s2>>>(pl, d->w, err, d->jobs, d->hint_cst, d->hint_tab, d->hint_img, d->hint_scratch, d->win_size, s0, ns);
            d->n_launch++, k_ed_ladder<<<(unsigned)((ns + 7) / 8), 96, 0, s2>>>(d->hint_scratch, d->ed_tab, d->ed_pts, ns);
            d->n_launch++, k_ed_conv<<<(unsigned)(((size_t)ns * HZ_LAD_TOT + 127) / 128), 128, 0, s2>>>(d->hint_scratch, d->hint_tab, d->ed_pts, ns);
            d->n_launch++, k_ed_hints<<<(unsigned)((ns + 3) / 4), 128, 0, s2>>>(d->hint_scratch, ns);
            d->n_launch++, k_eddsa_tail<<<(unsigned)((ns + 7) / 8), HZ_BLOCK, 0, s2>>>(d->w, err, d->hint_cst, d->hint_tab, d->hint_img, d->hint_scratch, d->win_size, d->fix_size, ns);
            d->n_launch++, k_ed_fill<<<(unsigned)((HZ_FILL_PER_SIG * ns + 127) / 128), 128, 0, s2>>>(d->w, err, d->hint_tab, d->hint_scratch, d->win_stride, ns);
        }
        CK(cudaEventRecord(d->ev_ed[h], s2));
        if (2 * ns + nf > 0) {
            d->n_launch++, k_leafhash<<<(unsigned)((2 * (2 * ns + nf) + 31) / 32), 128, 0, s>>>(pl, d->w, d->smt, d->lev, s0, ns, 0, nf);
            d->n_launch++, k_smt<<<(unsigned)((2 * ns + nf + 31) / 32), 32, 0, s>>>(pl, d->w, err, d->smt, d->mids, d->lev, d->smst, s0, ns, 0, nf);
            const unsigned lb = (unsigned)((2 * ns + nf + 15) / 16);
            d->n_launch++, k_smt_levels<<<lb, 96, 0, s>>>(pl, d->w, d->lev, d->smst, d->mids, d->img00, d->zflag, s0, ns, 0, nf);
            d->n_launch++, k_hole_pack<<<(unsigned)(2 * ns + nf), 128, 0, s>>>(pl, d->w, d->lev, d->zflag, d->holes, d->gl, s0, ns, 0, nf);
            CK(cudaEventRecord(d->ev_pack[h], s));
            d->n_launch++, k_smt_zero<<<dim3(lb, (unsigned)(P.nLevels + 1)), 256, 0, s>>>(pl, d->w, d->lev, d->img00, d->zflag, s0, ns, 0, nf);
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
        if (ns + nf > 0) d->n_launch++, k_tail<<<hz_grid(ns + nf), HZ_BLOCK, 0, s>>>(pl, d->w, err, d->err, d->mids, d->isExit, d->isExit_front, d->jobs, d->jobs_front, d->smt, d->smt_front, s0, ns, 0, nf);
        CK(cudaStreamWaitEvent(s, d->ev_ed[h], 0));
        CK(cudaEventRecord(d->ev_half[h], s));
    }
    CK(cudaEventRecord(d->ev_eddsa, s2));
    int st = hz_phase2(d, sanity);
    if (st) return st;
    // the bit signals of the SHA blocks cross the link packed (k_bitpack) and are expanded by the host threads
    // (whole words of 32 entries; the few entries left over go with the piece behind them)
    const uint64_t sha0 = pl.off_sha_blocks, sha_n = (uint64_t)pl.nblocks * HZ_SZ_SHA_BLOCK / 32 * 32;
    const bool packed = sha_n >= (1u << 15) && !d->no_holes && !getenv("HZW_NO_BITPACK");
    if (packed) {
        CK(cudaMemsetAsync(d->bitpack + sha_n / 32, 0, 4, s));
        d->n_launch++, k_bitpack<<<1184, 256, 0, s>>>(d->w + sha0, sha_n, d->bitpack, d->bitpack + sha_n / 32);
        CK(cudaMemcpyAsync(d->h_bitpack, d->bitpack, 4 * (sha_n / 32 + 1), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(d->ev_bits, s));
        d->last_d2h += 4 * (sha_n / 32 + 1);
    }
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
    // from here on no early return: the workers are joined before an error is reported
    cudaError_t ce = cudaStreamWaitEvent(s3, d->ev_half[0], 0);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s6, d->ev_half[0], 0);
    if (ce == cudaSuccess) ce = slots_around(0, h0);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(d->ev_info[1]);
    if (ce == cudaSuccess) {
        if (d->no_holes) for (int j = 2 * h0; j < 2 * P.nTx + F; j++) d->h_holes[j].nz = 0;
        for (unsigned t = 0; t < nthr; t++) workers.emplace_back(hz_fill_holes, d, out, 2 * h0, 2 * P.nTx + F, (int)t, (int)nthr);
        ce = cudaStreamWaitEvent(s3, d->ev_done, 0);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s6, d->ev_done, 0);
    }
    if (ce == cudaSuccess) ce = slots_around(h0, P.nTx);
    bool expand = false;
    if (packed) {
        if (ce == cudaSuccess) ce = pieces_around(d->h_off_fee0, sha0, 2 * P.nTx, 2 * P.nTx + F);
        if (ce == cudaSuccess) ce = piece(sha0 + sha_n, pl.nvars);
        if (ce == cudaSuccess) ce = cudaEventSynchronize(d->ev_bits);
        if (ce == cudaSuccess) {
            expand = d->h_bitpack[sha_n / 32] == 0;
            if (!expand) ce = piece(sha0, sha0 + sha_n);          // an entry was not a bit: the range as it is
        }
    } else if (ce == cudaSuccess) ce = pieces_around(d->h_off_fee0, pl.nvars, 2 * P.nTx, 2 * P.nTx + F);
    if (ce == cudaSuccess) ce = piece(0, 2);
    if (expand) {
        const size_t nw = sha_n / 32, per = (nw + nthr - 1) / nthr;
        for (unsigned t = 0; t < nthr; t++) {
            const size_t w0 = (size_t)t * per, w1 = w0 + per < nw ? w0 + per : nw;
            if (w0 < w1) workers.emplace_back([=]() { hz_host_expand_bits(out + 32 * (sha0 + 32 * w0), d->h_bitpack + w0, w1 - w0); hz_host_copy_fence(); });
        }
    }
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
    d->n_launch = 0;
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
    if (launches) *launches = d->n_launch;
    if (h2d) *h2d = inputs ? 32ull * pl.n_inputs : 0;
    if (d2h) *d2h = (overlapped ? d->last_d2h : (witness_out ? 32ull * pl.nvars : 0)) + sizeof(HzErrBuf);
    return 0;
}

// sharded phase 1: own slots / fee txs, then this rank's exchange record into `record_out` (device memory)
int hz_cuda_shard_phase1(HzCuda* d, const uint8_t* inputs, int sanity, void* record_out, HzErrBuf* errs) {
    CK(cudaSetDevice(d->device));
    if (!record_out) return HZW_E_PARAM;
    d->n_launch = 0;
    int st = hz_phase1(d, inputs, sanity);
    if (st) return st;
    const HzParams& P = d->plan.P;
    int last_slot = d->ns > 0 && d->s0 + d->ns == P.nTx, last_fee = d->nf > 0 && d->f0 + d->nf == P.maxFeeTx;
    HzPeerRecs dst; memset(&dst, 0, sizeof dst);
    dst.n = 1; dst.p[0] = reinterpret_cast<uint8_t*>(record_out);
    const cudaStream_t ps = d->xearly ? d->stream7 : d->stream;      // early exchange: the record is complete long before the levels are
    d->n_launch++, k_pack<<<148, 256, 0, ps>>>(d->plan, d->w, d->pre, dst, d->s0, d->ns, d->per, last_slot, last_fee);
    CK(cudaGetLastError());
    if (d->xearly) memset(errs, 0, sizeof(HzErrBuf));      // the kernels are still running: phase 2 reports what they find
    else CK(cudaMemcpyAsync(errs, d->err, sizeof(HzErrBuf), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(ps));                // the host binding runs the all-gather on its own stream
    return 0;
}
// sharded phase 2: `records` = the gathered records of all ranks (device memory); the shard's witness
// pieces are copied to `shard_out` (host) back to back in hz_cuda_pieces order
int hz_cuda_shard_phase2(HzCuda* d, const void* records, int sanity, uint8_t* shard_out, const uint64_t* pieces, int npieces,
                         float* kernel_ms, HzErrBuf* errs, uint64_t* d2h, bool in_place) {
    CK(cudaSetDevice(d->device));
    if (!records) records = d->recs;              // the library's own all-gather left them there (hz_cuda_multi_phase1)
    if (!records) return HZW_E_PARAM;
    cudaStream_t s = d->stream;
    d->n_launch++, k_unpack<<<148, 256, 0, d->xearly ? d->stream7 : s>>>(d->plan, d->w, d->pre, reinterpret_cast<const uint8_t*>(records), d->world, d->per, hz_rec_bytes(d->plan.P, d->per));
    int st = hz_phase2(d, sanity);
    if (st) return st;
    uint64_t bytes = 0;
    std::vector<std::thread> workers;
    d->last_d2h = 0;
    // this rank's SHA blocks leave packed, as on the end-to-end path of one GPU (k_bitpack)
    const uint64_t sha_a = (uint64_t)d->plan.off_sha_blocks + (uint64_t)d->b0 * HZ_SZ_SHA_BLOCK, sha_n = (uint64_t)d->nb * HZ_SZ_SHA_BLOCK / 32 * 32;
    const bool packed = shard_out && sha_n >= (1u << 15) && !d->no_holes && !getenv("HZW_NO_BITPACK");
    if (packed) {
        CK(cudaMemsetAsync(d->bitpack + sha_n / 32, 0, 4, s));
        d->n_launch++, k_bitpack<<<1184, 256, 0, s>>>(d->w + sha_a, sha_n, d->bitpack, d->bitpack + sha_n / 32);
        CK(cudaMemcpyAsync(d->h_bitpack, d->bitpack, 4 * (sha_n / 32 + 1), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(d->ev_bits, s));
    }
    uint8_t* sha_out = nullptr;                    // where witness entry 0 would be for the piece with the SHA blocks
    if (shard_out) {
        // as on the end-to-end path of one GPU: the image levels of the shard's processors are rebuilt in the shard
        // buffer by host threads, the pieces are copied around them
        const HzParams& P = d->plan.P;
        unsigned nthr = std::thread::hardware_concurrency();
        nthr = nthr == 0 ? 4 : (nthr > 18 ? 16 : (nthr > 3 ? nthr - 2 : 1));
        nthr = nthr / (unsigned)d->world > 0 ? nthr / (unsigned)d->world : 1;      // the ranks of a box share its cores
        // rebuilding the images pays at every rank count measured: with eight ranks a single thread per rank still
        // beats eight copy engines writing the whole 8.1 GB into the host's memory (80 ms against 97 ms per batch)
        const bool holes = d->world > 1 && d->ns + d->nf > 0 && !d->no_holes;      // phase 1 packs the hole records only for world > 1
        CK(cudaEventRecord(d->ev_done, s));
        CK(cudaStreamWaitEvent(d->stream3, d->ev_done, 0));
        CK(cudaStreamWaitEvent(d->stream6, d->ev_done, 0));
        if (holes) CK(cudaEventSynchronize(d->ev_info[0]));
        const uint32_t r0 = d->ns > 0 ? (*d->h_off_rollup)[d->s0] : 0xffffffffu;
        const int fj0 = 2 * P.nTx + d->f0, fj1 = fj0 + d->nf;
        cudaError_t ce = cudaSuccess;
        for (int k = 0; k < npieces && ce == cudaSuccess; k++) {
            const uint64_t a = pieces[2 * k], n = pieces[2 * k + 1];
            // where witness entry 0 would be for this piece: the pieces back to back in a shard buffer, or each at its
            // place in ONE whole-witness buffer that every GPU of the box copies into (in_place)
            uint8_t* out_v = in_place ? shard_out : shard_out + bytes - 32ull * a;
            HzCopier cp{d, out_v, reinterpret_cast<const uint8_t*>(d->w), (unsigned)k};
            if (packed && a == sha_a && n >= sha_n) { sha_out = out_v; ce = cp.piece(a + sha_n, a + n); }
            else if (holes && d->ns > 0 && a == r0 && n == (uint64_t)(*d->h_off_rollup)[d->s0 + d->ns] - r0) {
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
        if (sha_out && ce == cudaSuccess) {
            ce = cudaEventSynchronize(d->ev_bits);
            d->last_d2h += 4 * (sha_n / 32 + 1);
            if (ce == cudaSuccess && d->h_bitpack[sha_n / 32] != 0) {           // an entry was not a bit: the range as it is
                HzCopier cp{d, sha_out, reinterpret_cast<const uint8_t*>(d->w), 0u};
                ce = cp.piece(sha_a, sha_a + sha_n);
            } else if (ce == cudaSuccess) {
                const size_t nw = sha_n / 32, per = (nw + nthr - 1) / nthr;
                for (unsigned t = 0; t < nthr; t++) {
                    const size_t w0 = (size_t)t * per, w1 = w0 + per < nw ? w0 + per : nw;
                    if (w0 < w1) workers.emplace_back([=]() { hz_host_expand_bits(sha_out + 32 * (sha_a + 32 * w0), d->h_bitpack + w0, w1 - w0); hz_host_copy_fence(); });
                }
            }
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

// ---- all the GPUs of one box driven by ONE process (hzw_params.n_gpus > 1; hzw_host.cpp: MultiBackend) ----------
// Every device of the group must be able to store into the others' memory: the exchange is done by the pack kernel.
int hz_cuda_enable_peer(HzCuda* d, HzCuda* peer) {
    CK(cudaSetDevice(d->device));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, d->device, peer->device));
    if (!can) { fprintf(stderr, "hzw: device %d cannot access device %d (no peer-to-peer path)\n", d->device, peer->device); return HZW_E_DEVICE; }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
    CK(e);
    return 0;
}
uint8_t* hz_cuda_recs(HzCuda* d) { return d->recs; }
// phase 1 of rank d->rank, nothing waited for: own slots and fee txs, then k_pack stores this rank's record into
// slot d->rank of EVERY rank's gathered array (peer_recs[r], device memory of rank r) -- compute step and all-gather
// in one kernel over NVLink -- and ev_xchg marks it
int hz_cuda_multi_phase1(HzCuda* d, const uint8_t* inputs, int sanity, uint8_t* const* peer_recs, int world) {
    CK(cudaSetDevice(d->device));
    if (world != d->world || world > HZ_MAX_GPUS) return HZW_E_PARAM;
    d->n_launch = 0;
    int st = hz_phase1(d, inputs, sanity);
    if (st) return st;
    const HzParams& P = d->plan.P;
    int last_slot = d->ns > 0 && d->s0 + d->ns == P.nTx, last_fee = d->nf > 0 && d->f0 + d->nf == P.maxFeeTx;
    HzPeerRecs dst; memset(&dst, 0, sizeof dst);
    dst.n = world;
    for (int r = 0; r < world; r++) dst.p[r] = peer_recs[r] + (size_t)d->rank * hz_rec_bytes(P, d->per);
    const cudaStream_t ps = d->xearly ? d->stream7 : d->stream;
    d->n_launch++, k_pack<<<148, 256, 0, ps>>>(d->plan, d->w, d->pre, dst, d->s0, d->ns, d->per, last_slot, last_fee);
    CK(cudaGetLastError());
    CK(cudaEventRecord(d->ev_xchg, ps));
    return 0;
}
// rank d's main stream goes on only when `peer`'s record has landed (called after every rank has queued its phase 1)
int hz_cuda_multi_wait(HzCuda* d, HzCuda* peer) {
    CK(cudaSetDevice(d->device));
    CK(cudaStreamWaitEvent(d->xearly ? d->stream7 : d->stream, peer->ev_xchg, 0));
    return 0;
}

// introspection for the tests of the signature pipeline: Fr entries of the per-signature work areas
// (which = 0: hint scratch of signature g, 1: its ladder points, 2: the cached window table)
int hz_cuda_debug_read(HzCuda* d, int which, uint64_t g, uint64_t first, uint64_t count, uint8_t* out) {
    CK(cudaSetDevice(d->device));
    const Fr* base = which == 0 ? d->hint_scratch + g * HZ_HINT_SCRATCH : (which == 1 ? d->ed_pts + g * HZ_PTS : d->ed_tab);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, base + first, sizeof(Fr) * count, cudaMemcpyDeviceToHost));
    return 0;
}
int hz_cuda_witness_device(HzCuda* d, void** ptr, int32_t* device, void** stream) {
    *ptr = d->w;
    if (device) *device = d->device;
    if (stream) *stream = d->stream;
    return 0;
}
int hz_cuda_witness_read(HzCuda* d, uint64_t first, uint64_t count, uint8_t* out) {
    CK(cudaSetDevice(d->device));
    CK(cudaMemcpyAsync(out, d->w + first, 32ull * count, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    return 0;
}

// One batch (inputs already resident) with every launch on the main stream and an event in front of each: the
// duration of each kernel ALONE, measured live -- what the roofline of a kernel is quoted on (the timed steps run
// the same launches on five streams, where an interval between two events also holds the kernels beside it).
int hz_cuda_profile_step(HzCuda* d, float* ms, const char** names, int cap) {
    CK(cudaSetDevice(d->device));
    if (d->world != 1) return HZW_E_PARAM;
    d->prof = true; d->n_prof = 0; d->n_launch = 0;
    int st = hz_phase1(d, nullptr, 0);
    if (!st) st = hz_phase2(d, 0);
    d->prof = false;
    if (st) return st;
    CK(cudaStreamSynchronize(d->stream));
    int n = d->n_prof < cap ? d->n_prof : cap;
    for (int k = 0; k < n; k++) {
        if (cudaEventElapsedTime(&ms[k], d->pev[k], d->pev[k + 1]) != cudaSuccess) return HZW_E_DEVICE;
        names[k] = d->pname[k];
    }
    return n;
}

// processor levels of the shard in total and how many of them the last batch flagged as (0, 0) images (skipped by
// k_smt_levels, copied by k_smt_zero): what the algorithmic work of those two kernels is counted from
int hz_cuda_image_levels(HzCuda* d, uint64_t out[2]) {
    CK(cudaSetDevice(d->device));
    const int L = d->plan.P.nLevels + 1, nj = 2 * d->ns + d->nf, lb = (nj + 15) / 16;
    std::vector<uint8_t> z((size_t)lb * L);
    if (lb > 0) CK(cudaMemcpy(z.data(), d->zflag, z.size(), cudaMemcpyDeviceToHost));
    out[0] = (uint64_t)nj * L; out[1] = 0;
    for (int b = 0; b < lb; b++) {
        const int in_block = b + 1 < lb ? 16 : nj - 16 * b;
        for (int l = 0; l < L; l++) if (z[(size_t)b * L + l]) out[1] += (uint64_t)in_block;
    }
    return 0;
}

int hz_cuda_kernel_ms(HzCuda* d, float* ms, const char** names, int cap) {
    int n = cap < HZ_NKERN ? cap : HZ_NKERN;
    for (int k = 0; k < n; k++) {
        // kernels 0..7 are bracketed by kev[k], kev[k+1] on the main stream; k_eddsa by kev[9], kev[10]
        cudaEvent_t a = k < 8 ? d->kev[k] : (k == 8 ? d->kev[9] : d->kev[11]), b = k == 3 ? d->kev[13] : (k < 8 ? d->kev[k + 1] : (k == 8 ? d->kev[10] : d->kev[12]));
        if (cudaEventElapsedTime(&ms[k], a, b) != cudaSuccess) return HZW_E_DEVICE;
        if (names) names[k] = HZ_KERN_NAMES[k];
    }
    return n;
}
