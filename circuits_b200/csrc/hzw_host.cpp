// libhzwitness.so, host half: the C ABI of include/hzwitness.h on top of the CUDA backend.
// There is no CPU evaluation path in this library: if CUDA is unavailable hzw_create fails
// with HZW_E_DEVICE.
#include "hzw_api.hpp"
#include "hzw_cuda.h"
#define HZ_MAX_GPUS_HOST 16

struct CudaBackend : HzBackend {
    HzCuda* d;
    explicit CudaBackend(HzCuda* d_) : d(d_) {}
    ~CudaBackend() override { hz_cuda_destroy(d); }
    int calculate(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (!resident && !out) return HZW_E_PARAM;
        return hz_cuda_calculate(d, inputs, sanity, resident ? nullptr : out, ms, errs, &launches, &h2d, &d2h);
    }
    int kernel_ms(float* ms, const char** names, int cap) override { return hz_cuda_kernel_ms(d, ms, names, cap); }
    int profile_step(float* ms, const char** names, int cap) override { return hz_cuda_profile_step(d, ms, names, cap); }
    int image_levels(uint64_t out[2]) override { return hz_cuda_image_levels(d, out); }
    int witness_device(void** ptr, int32_t* device, void** stream) override { return hz_cuda_witness_device(d, ptr, device, stream); }
    int witness_read(uint64_t first, uint64_t count, uint8_t* out) override { return hz_cuda_witness_read(d, first, count, out); }
    int debug_read(int which, uint64_t g, uint64_t first, uint64_t count, uint8_t* out) override { return hz_cuda_debug_read(d, which, g, first, count, out); }
    // a circom .sym as a layout map: the batch is evaluated in this library's layout, a gather kernel puts it into the
    // compiler's, and that is what leaves the device (one plain copy: the map decides which entries exist at all)
    uint64_t n_out = 0;
    int set_layout_map(const uint32_t* src, uint64_t n) override { n_out = n; return hz_cuda_set_map(d, src, n); }
    int calculate_mapped(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (!n_out || (!resident && !out)) return HZW_E_PARAM;
        int st = hz_cuda_calculate(d, inputs, sanity, nullptr, ms, errs, &launches, &h2d, &d2h);
        if (!st) st = hz_cuda_remap(d);
        launches++;
        if (!st && !resident) { st = hz_cuda_read_mapped(d, 0, n_out, out); d2h += 32ull * n_out; }
        return st;
    }
    int read_mapped(uint64_t first, uint64_t count, uint8_t* out) override { return hz_cuda_read_mapped(d, first, count, out); }
    int calculate_to_wtns(const uint8_t* inputs, int sanity, const char* path, float* ms, HzErrBuf* errs) override {
        int st = n_out ? calculate_mapped(inputs, sanity, nullptr, true, ms, errs) : calculate(inputs, sanity, nullptr, true, ms, errs);
        if (st || errs->count || errs->pad) return st;          // a violated constraint: no file, as the reference's binary
        uint8_t hdr[HZ_WTNS_HEADER];
        hz_wtns_header(hdr, n_out ? n_out : L->plan.nvars);
        uint64_t moved = 0;
        st = hz_cuda_stream_wtns(d, path, hdr, sizeof hdr, &moved);
        d2h += moved;
        return st;
    }
    int r1cs_check(const HzR1cs& R, const uint8_t* witness, uint64_t* n_bad, uint64_t* first_bad, float* ms) override {
        return hz_cuda_r1cs_check(d, R, witness, n_bad, first_bad, ms);
    }
    const HzLayout* L = nullptr;
    HzShard sh;
    int set_shard(int rank, int world) override {
        int st = hz_cuda_set_shard(d, rank, world);
        if (!st) sh = hz_shard_ranges(L->plan, rank, world);
        return st;
    }
    uint64_t exchange_bytes() override { return hz_cuda_exchange_bytes(d); }
    int pieces(uint64_t* table, int cap) override { return hz_shard_pieces(*L, sh, table, cap); }
    int shard_phase1(const uint8_t* inputs, int sanity, void* record_out, HzErrBuf* errs) override {
        h2d = inputs ? 32ull * L->plan.n_inputs : 0;
        return hz_cuda_shard_phase1(d, inputs, sanity, record_out, errs);
    }
    int withdraw_batch(uint64_t, const uint8_t*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int shard_phase2(const void* records, int sanity, uint8_t* shard_out, float* ms, HzErrBuf* errs) override {
        uint64_t t[32]; int n = hz_shard_pieces(*L, sh, t, 16);
        int st = hz_cuda_shard_phase2(d, records, sanity, shard_out, t, n, ms, errs, &d2h);
        launches = hz_cuda_launches(d);
        return st;
    }
};

// ---- one process, all the GPUs of a box (hzw_params.n_gpus > 1) --------------------------------------------------
// One CUDA context per device, rank r of n evaluating its range of slots / fee txs / SHA blocks exactly as a rank of
// the multi-process form does; the difference is who runs the exchange and where the witness lands:
//   * the all-gather is done by the pack kernels themselves, storing each rank's record into every peer over NVLink
//     (hz_cuda_multi_phase1), ordered by one event per rank -- no host synchronisation between the two phases;
//   * every GPU copies the pieces it owns straight to their offsets in the caller's ONE witness buffer over its own
//     PCIe link (`./circuit input.json witness.wtns` produces one file: /root/reference/tools/helpers/actions.js:139-143).
// A host thread per device issues that device's work; they meet once, after queueing phase 1.
#include <thread>
struct MultiBackend : HzBackend {
    std::vector<std::unique_ptr<CudaBackend> > rk;
    const HzLayout* L = nullptr;
    int n = 0;
    std::vector<float> ms_rank;
    int run(const uint8_t* inputs, int sanity, uint8_t* out, float* ms, HzErrBuf* errs) {
        std::vector<int> st(n, 0);
        std::vector<HzErrBuf> eb(n);
        std::vector<uint8_t*> recs(n);
        for (int r = 0; r < n; r++) { recs[r] = hz_cuda_recs(rk[r]->d); memset(&eb[r], 0, sizeof(HzErrBuf)); }
        {
            std::vector<std::thread> th;
            for (int r = 0; r < n; r++) th.emplace_back([&, r]() { st[r] = hz_cuda_multi_phase1(rk[r]->d, inputs, sanity, recs.data(), n); });
            for (auto& t : th) t.join();
        }
        for (int r = 0; r < n; r++) if (st[r]) return st[r];       // nothing waits on a peer yet: safe to stop here
        ms_rank.assign(n, 0.f);
        {
            std::vector<std::thread> th;
            for (int r = 0; r < n; r++) th.emplace_back([&, r]() {
                for (int q = 0; q < n && !st[r]; q++) if (q != r) st[r] = hz_cuda_multi_wait(rk[r]->d, rk[q]->d);
                if (st[r]) return;
                uint64_t t[32]; int np = hz_shard_pieces(*L, rk[r]->sh, t, 16);
                uint64_t moved = 0;
                st[r] = hz_cuda_shard_phase2(rk[r]->d, nullptr, sanity, out, t, np, &ms_rank[r], &eb[r], &moved, true);
                rk[r]->d2h = moved; rk[r]->launches = hz_cuda_launches(rk[r]->d);
            });
            for (auto& t : th) t.join();
        }
        launches = 0; d2h = 0; h2d = inputs ? (uint64_t)n * 32ull * L->plan.n_inputs : 0;
        float worst = 0.f;
        for (int r = 0; r < n; r++) {
            if (st[r]) return st[r];
            launches += rk[r]->launches; d2h += rk[r]->d2h;
            if (ms_rank[r] > worst) worst = ms_rank[r];
        }
        if (ms) *ms = worst;                       // device time of the slowest rank (ev0 -> ev1 on its own device)
        // first failing constraint over all ranks: the records of every rank, lowest witness cursor wins in hz_report
        memset(errs, 0, sizeof(HzErrBuf));
        for (int r = 0; r < n; r++) {
            errs->pad += eb[r].pad;
            const uint32_t k = eb[r].count < HZ_MAX_ERR ? eb[r].count : HZ_MAX_ERR;
            for (uint32_t j = 0; j < k; j++) {
                if (errs->count < HZ_MAX_ERR) errs->rec[errs->count++] = eb[r].rec[j];
                else {
                    uint32_t hi = 0;
                    for (uint32_t q = 1; q < HZ_MAX_ERR; q++) if (errs->rec[q].cur > errs->rec[hi].cur) hi = q;
                    if (eb[r].rec[j].cur < errs->rec[hi].cur) errs->rec[hi] = eb[r].rec[j];
                }
            }
        }
        return 0;
    }
    int calculate(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (!resident && !out) return HZW_E_PARAM;
        return run(inputs, sanity, resident ? nullptr : out, ms, errs);
    }
    int kernel_ms(float* ms, const char** names, int cap) override { return hz_cuda_kernel_ms(rk[0]->d, ms, names, cap); }
    int set_shard(int, int) override { return HZW_E_PARAM; }       // the context already spans its GPUs
    uint64_t exchange_bytes() override { return hz_cuda_exchange_bytes(rk[0]->d); }
    // pieces of rank `cap < 0 ? -cap - 1 : 0`... the table of rank 0; hzw_rank_pieces gives the others
    int pieces(uint64_t* table, int cap) override { return hz_shard_pieces(*L, rk[0]->sh, table, cap); }
    int rank_pieces(int r, uint64_t* table, int cap) override { return r < 0 || r >= n ? HZW_E_PARAM : hz_shard_pieces(*L, rk[r]->sh, table, cap); }
    int shard_phase1(const uint8_t*, int, void*, HzErrBuf*) override { return HZW_E_PARAM; }
    int shard_phase2(const void*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int withdraw_batch(uint64_t, const uint8_t*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    // resident witness: rank r's pieces live on rank r's device
    int witness_device(void** ptr, int32_t* device, void** stream) override { return hz_cuda_witness_device(rk[0]->d, ptr, device, stream); }
    int rank_witness_device(int r, void** ptr, int32_t* device, void** stream) override {
        return r < 0 || r >= n ? HZW_E_PARAM : hz_cuda_witness_device(rk[r]->d, ptr, device, stream);
    }
    int witness_read(uint64_t first, uint64_t count, uint8_t* out) override {
        // entry ranges are read from the rank that owns them
        for (int r = 0; r < n; r++) {
            uint64_t t[32]; int np = hz_shard_pieces(*L, rk[r]->sh, t, 16);
            for (int k = 0; k < np; k++) {
                const uint64_t a = t[2 * k] > first ? t[2 * k] : first, b = t[2 * k] + t[2 * k + 1] < first + count ? t[2 * k] + t[2 * k + 1] : first + count;
                if (b > a) { int st = hz_cuda_witness_read(rk[r]->d, a, b - a, out + 32ull * (a - first)); if (st) return st; }
            }
        }
        return 0;
    }
};

HzBackend* hz_backend_create(const hzw_params& prm, const HzLayout& L, int* status) {
    const int n = prm.n_gpus > 1 ? prm.n_gpus : 1;
    if (n > HZ_MAX_GPUS_HOST || n > L.plan.P.nTx) { *status = HZW_E_PARAM; return nullptr; }
    std::vector<std::unique_ptr<CudaBackend> > rk;
    for (int r = 0; r < n; r++) {
        HzCuda* d = hz_cuda_create(prm.device + r, L.plan, L.off_decode.data(), L.off_rollup.data(), L.off_fee.data(), L.off_eddsa.data(), L.off_jk.data(), L.off_hs.data(), L.off_hsig.data(), status);
        if (!d) return nullptr;
        rk.emplace_back(new CudaBackend(d));
        rk.back()->L = &L;
        rk.back()->sh = hz_shard_ranges(L.plan, 0, 1);
    }
    if (n == 1) return rk[0].release();
    MultiBackend* m = new MultiBackend();
    m->L = &L; m->n = n;
    for (int r = 0; r < n; r++) {
        int st = rk[r]->set_shard(r, n);
        for (int q = 0; q < n && !st; q++) if (q != r) st = hz_cuda_enable_peer(rk[r]->d, rk[q]->d);
        if (st) { *status = st; delete m; return nullptr; }
    }
    m->rk = std::move(rk);
    *status = 0;
    return m;
}

// Withdraw(nLevels): instances are independent -- no exchange at all.  One device state per GPU (n_gpus > 1: the
// instances of a batch are split into contiguous ranges, one host thread per device; config 5 of BASELINE.json).
struct HzBatchShape { uint32_t n_inputs, nvars; };
struct CudaWithdrawBackend : HzBackend {
    std::vector<HzWdCuda*> dev; HzBatchShape WL;
    ~CudaWithdrawBackend() override { for (auto* d : dev) hz_cuda_withdraw_destroy(d); }
    int calculate(const uint8_t*, int, uint8_t*, bool, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int set_shard(int, int) override { return HZW_E_PARAM; }
    uint64_t exchange_bytes() override { return 0; }
    int pieces(uint64_t*, int) override { return 0; }
    int shard_phase1(const uint8_t*, int, void*, HzErrBuf*) override { return HZW_E_PARAM; }
    int shard_phase2(const void*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int run(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, uint8_t* pub, float* ms, HzErrBuf* errs) {
        const int g = (int)dev.size();
        std::vector<int> st(g, 0); std::vector<float> t(g, 0.f); std::vector<HzErrBuf> eb(g); std::vector<uint64_t> nl(g, 0);
        const uint64_t per = (n + g - 1) / g;
        auto work = [&](int r) {
            const uint64_t k0 = per * r < n ? per * r : n, m = k0 + per <= n ? per : n - k0;
            memset(&eb[r], 0, sizeof(HzErrBuf));
            st[r] = hz_cuda_withdraw_batch(dev[r], m, inputs + 32ull * WL.n_inputs * k0, sanity, out ? out + 32ull * WL.nvars * k0 : nullptr,
                                           pub ? pub + 32ull * k0 : nullptr, &t[r], &eb[r], &nl[r]);
        };
        if (g == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int r = 0; r < g; r++) th.emplace_back(work, r);
            for (auto& x : th) x.join();
        }
        memset(errs, 0, sizeof(HzErrBuf));
        float worst = 0.f; launches = 0;
        for (int r = 0; r < g; r++) {
            if (st[r]) return st[r];
            if (t[r] > worst) worst = t[r];
            launches += nl[r];
            if (eb[r].count && !errs->count) *errs = eb[r];          // instances are independent: the first range that failed
        }
        if (ms) *ms = worst;
        h2d = 32ull * WL.n_inputs * n; d2h = (out ? 32ull * WL.nvars * n : 0) + (pub ? 32ull * n : 0);
        return 0;
    }
    int withdraw_batch(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, float* ms, HzErrBuf* errs) override {
        return run(n, inputs, sanity, out, nullptr, ms, errs);
    }
    int withdraw_resident(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* pub, float* ms, HzErrBuf* errs) override {
        return run(n, inputs, sanity, nullptr, pub, ms, errs);
    }
};
static HzBackend* hz_batch_backend(const hzw_params& prm, int nLevels, int maxFeeTx, uint32_t n_inputs, uint32_t nvars, int* status) {
    const int n = prm.n_gpus > 1 ? prm.n_gpus : 1;
    if (n > HZ_MAX_GPUS_HOST) { *status = HZW_E_PARAM; return nullptr; }
    CudaWithdrawBackend* b = new CudaWithdrawBackend();
    b->WL.n_inputs = n_inputs; b->WL.nvars = nvars;
    for (int r = 0; r < n; r++) {
        HzWdCuda* d = hz_cuda_withdraw_create(prm.device + r, nLevels, maxFeeTx, n_inputs, nvars, status);
        if (!d) { delete b; return nullptr; }
        b->dev.push_back(d);
    }
    *status = 0;
    return b;
}
HzBackend* hz_backend_create_withdraw(const hzw_params& prm, const HzWLayout& WL, int* status) {
    return hz_batch_backend(prm, WL.nLevels, 0, WL.n_inputs, WL.nvars, status);
}
HzBackend* hz_backend_create_rolluptx(const hzw_params& prm, const HzTLayout& TL, int* status) {
    return hz_batch_backend(prm, TL.nLevels, TL.maxFeeTx, TL.n_inputs, TL.nvars, status);
}

// ---- streaming copy for the host side of the end-to-end path (hzw_run.cu: hz_fill_holes) -----------------
// The rebuilt images are written once and not read again by this process: non-temporal stores keep them out of
// the caches and spare the read-for-ownership of every line.  n is a multiple of 32, dst is 32-byte aligned.
#include <immintrin.h>
__attribute__((target("avx2"))) static void hz_copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
    for (size_t i = 0; i < n; i += 32)
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)));
}
void hz_host_copy_stream(uint8_t* dst, const uint8_t* src, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0 && (n & 31u) == 0) hz_copy_nt_avx2(dst, src, n);
    else memcpy(dst, src, n);
}
void hz_host_copy_fence() { _mm_sfence(); }
// bit k of words[] -> the 32-byte field element 0 or 1 at out + 32 k (the packed transfer of the SHA block signals)
__attribute__((target("avx2"))) static void hz_expand_bits_avx2(uint8_t* out, const uint32_t* words, size_t nwords) {
    const __m256i one = _mm256_set_epi64x(0, 0, 0, 1), zero = _mm256_setzero_si256();
    for (size_t i = 0; i < nwords; i++) {
        const uint32_t wv = words[i];
        __m256i* o = reinterpret_cast<__m256i*>(out + 1024 * i);
        for (int b = 0; b < 32; b++) _mm256_stream_si256(o + b, (wv >> b) & 1u ? one : zero);
    }
}
void hz_host_expand_bits(uint8_t* out, const uint32_t* words, size_t nwords) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0) { hz_expand_bits_avx2(out, words, nwords); return; }
    for (size_t i = 0; i < nwords; i++)
        for (int b = 0; b < 32; b++) { uint8_t* o = out + 1024 * i + 32 * b; memset(o, 0, 32); o[0] = (uint8_t)((words[i] >> b) & 1u); }
}
