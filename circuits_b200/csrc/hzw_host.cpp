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
};

HzBackend* hz_backend_create(const hzw_params& prm, const HzLayout& L, int* status) {
    HzCuda* d = hz_cuda_create(prm.device, L.plan, L.off_decode.data(), L.off_rollup.data(), L.off_fee.data(), status);
    if (!d) return nullptr;
    return new CudaBackend(d);
}
