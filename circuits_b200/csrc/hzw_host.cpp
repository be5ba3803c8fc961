// libhzwitness.so, host half: the C ABI of include/hzwitness.h on top of the CUDA backend.
// There is no CPU evaluation path in this library: if CUDA is unavailable hzw_create fails
// with HZW_E_DEVICE.
#include "hzw_api.hpp"
#include "hzw_cuda.h"

struct CudaBackend : HzBackend {
    HzCuda* d;
    explicit CudaBackend(HzCuda* d_) : d(d_) {}
    ~CudaBackend() override { hz_cuda_destroy(d); }
    int calculate(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (!resident && !out) return HZW_E_PARAM;
        return hz_cuda_calculate(d, inputs, sanity, resident ? nullptr : out, ms, errs, &launches, &h2d, &d2h);
    }
    int kernel_ms(float* ms, const char** names, int cap) override { return hz_cuda_kernel_ms(d, ms, names, cap); }
    int witness_device(void** ptr, int32_t* device, void** stream) override { return hz_cuda_witness_device(d, ptr, device, stream); }
    int witness_read(uint64_t first, uint64_t count, uint8_t* out) override { return hz_cuda_witness_read(d, first, count, out); }
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

HzBackend* hz_backend_create(const hzw_params& prm, const HzLayout& L, int* status) {
    HzCuda* d = hz_cuda_create(prm.device, L.plan, L.off_decode.data(), L.off_rollup.data(), L.off_fee.data(), L.off_eddsa.data(), L.off_jk.data(), L.off_hs.data(), L.off_hsig.data(), status);
    if (!d) return nullptr;
    CudaBackend* b = new CudaBackend(d);
    b->L = &L;
    b->sh = hz_shard_ranges(L.plan, 0, 1);
    return b;
}

// Withdraw(nLevels): instances are independent; the backend holds no per-circuit device state
struct CudaWithdrawBackend : HzBackend {
    int device; HzWLayout WL;
    int calculate(const uint8_t*, int, uint8_t*, bool, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int set_shard(int, int) override { return HZW_E_PARAM; }
    uint64_t exchange_bytes() override { return 0; }
    int pieces(uint64_t*, int) override { return 0; }
    int shard_phase1(const uint8_t*, int, void*, HzErrBuf*) override { return HZW_E_PARAM; }
    int shard_phase2(const void*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int withdraw_batch(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, float* ms, HzErrBuf* errs) override {
        launches = 1; h2d = 32ull * WL.n_inputs * n; d2h = 32ull * WL.nvars * n;
        return hz_cuda_withdraw(device, WL.nLevels, WL.nvars, n, inputs, sanity, out, ms, errs);
    }
};
HzBackend* hz_backend_create_withdraw(const hzw_params& prm, const HzWLayout& WL, int* status) {
    int st = hz_cuda_probe(prm.device);
    if (st) { *status = st; return nullptr; }
    CudaWithdrawBackend* b = new CudaWithdrawBackend();
    b->device = prm.device; b->WL = WL;
    return b;
}

// ---- streaming copy for the host side of the end-to-end path (hzw_kernels.cu: hz_fill_holes) -----------------
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
