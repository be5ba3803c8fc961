// The formats around the witness path, device side (formats.hpp has the readers):
//   k_remap        the witness gathered into the COMPILER's layout (a circom .sym applied as a layout map): entry v of
//                  the output is entry src[v] of this library's own layout -- one coalesced 32-byte store per thread,
//                  the loads scattered but each a whole sector
//   k_r1cs_check   A.w * B.w = C.w for every constraint of an .r1cs, one thread per constraint over CSR term lists:
//                  the sanity check of the reference's tester (`checkConstraints`), independent of the template kernels
#include "hzw_dev.cuh"
#include "formats.hpp"

__global__ void __launch_bounds__(256) k_remap(const Fr* __restrict__ w, const uint32_t* __restrict__ src, Fr* __restrict__ out, uint64_t n) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint4* p = reinterpret_cast<const uint4*>(w + src[v]);
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    __stcs(reinterpret_cast<uint4*>(out + v), a);
    __stcs(reinterpret_cast<uint4*>(out + v) + 1, b);
}

__device__ __forceinline__ Fr hz_lc(const Fr* __restrict__ w, const uint32_t* __restrict__ wire, const Fr* __restrict__ coef, uint64_t a, uint64_t b) {
    Fr s = fr_zero();
    for (uint64_t t = a; t < b; t++) {
        const Fr c = coef[t], x = w[wire[t]];
        // the coefficients of a circom circuit are mostly 1, -1 and powers of two: spare the product for 1
        Fr one = fr_one();
        s = fr_add(s, fr_eq(c, one) ? x : fr_mul(c, x));
    }
    return s;
}
__global__ void __launch_bounds__(128) k_r1cs_check(const Fr* __restrict__ w, const uint64_t* __restrict__ off, const uint32_t* __restrict__ wire,
                                                    const Fr* __restrict__ coef, uint32_t m, unsigned long long* res /* [0] bad, [1] first bad */) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const Fr A = hz_lc(w, wire, coef, off[3 * (uint64_t)c], off[3 * (uint64_t)c + 1]);
    const Fr B = hz_lc(w, wire, coef, off[3 * (uint64_t)c + 1], off[3 * (uint64_t)c + 2]);
    const Fr C = hz_lc(w, wire, coef, off[3 * (uint64_t)c + 2], off[3 * (uint64_t)c + 3]);
    if (!fr_eq(fr_mul(A, B), C)) { atomicAdd(&res[0], 1ull); atomicMin(&res[1], (unsigned long long)c); }
}

// ---- layout map --------------------------------------------------------------------------------------------------
int hz_cuda_set_map(HzCuda* d, const uint32_t* src, uint64_t n_out) {
    CK(cudaSetDevice(d->device));
    cudaFree(d->map_src); cudaFree(d->w_out); d->map_src = nullptr; d->w_out = nullptr; d->n_out = 0;
    if (!src || n_out == 0) return 0;
    CK(cudaMalloc(&d->map_src, 4ull * n_out));
    CK(cudaMalloc(&d->w_out, 32ull * n_out));
    CK(cudaMemcpy(d->map_src, src, 4ull * n_out, cudaMemcpyHostToDevice));
    d->n_out = n_out;
    return 0;
}
// the resident witness (own layout, d->w) gathered into the compiler's layout (d->w_out) on the main stream
int hz_cuda_remap(HzCuda* d) {
    CK(cudaSetDevice(d->device));
    if (!d->n_out) return HZW_E_PARAM;
    d->n_launch++, k_remap<<<(unsigned)((d->n_out + 255) / 256), 256, 0, d->stream>>>(d->w, d->map_src, d->w_out, d->n_out);
    CK(cudaGetLastError());
    return 0;
}
int hz_cuda_read_mapped(HzCuda* d, uint64_t first, uint64_t count, uint8_t* out) {
    CK(cudaSetDevice(d->device));
    if (!d->n_out || first + count > d->n_out) return HZW_E_PARAM;
    CK(cudaMemcpyAsync(out, d->w_out + first, 32ull * count, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    return 0;
}

// ---- constraint check --------------------------------------------------------------------------------------------
// witness: host buffer in the layout the .r1cs numbers its wires in (n_wires entries), or null = the witness resident
// on the device (the gathered one when a layout map is set, else this library's own)
int hz_cuda_r1cs_check(HzCuda* d, const HzR1cs& R, const uint8_t* witness, uint64_t* n_bad, uint64_t* first_bad, float* ms) {
    CK(cudaSetDevice(d->device));
    const Fr* w = d->n_out ? d->w_out : d->w;
    const uint64_t have = d->n_out ? d->n_out : d->plan.nvars;
    Fr* wtmp = nullptr;
    if (witness) {
        CK(cudaMalloc(&wtmp, 32ull * R.n_wires));
        CK(cudaMemcpy(wtmp, witness, 32ull * R.n_wires, cudaMemcpyHostToDevice));
        w = wtmp;
    } else if (R.n_wires > have) return HZW_E_PARAM;
    uint64_t* off = nullptr; uint32_t* wire = nullptr; Fr* coef = nullptr; unsigned long long* res = nullptr;
    const size_t nt = R.wire.size();
    int rc = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if (cudaMalloc(&off, 8 * R.off.size()) != cudaSuccess || cudaMalloc(&wire, 4 * (nt ? nt : 1)) != cudaSuccess ||
            cudaMalloc(&coef, 32 * (nt ? nt : 1)) != cudaSuccess || cudaMalloc(&res, 16) != cudaSuccess) { rc = HZW_E_DEVICE; break; }
        const unsigned long long init[2] = {0ull, ~0ull};
        if (cudaMemcpy(off, R.off.data(), 8 * R.off.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
            (nt && cudaMemcpy(wire, R.wire.data(), 4 * nt, cudaMemcpyHostToDevice) != cudaSuccess) ||
            (nt && cudaMemcpy(coef, R.coef.data(), 32 * nt, cudaMemcpyHostToDevice) != cudaSuccess) ||
            cudaMemcpy(res, init, 16, cudaMemcpyHostToDevice) != cudaSuccess) { rc = HZW_E_DEVICE; break; }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, d->stream);
        if (R.n_constraints) { d->n_launch++; k_r1cs_check<<<(R.n_constraints + 127) / 128, 128, 0, d->stream>>>(w, off, wire, coef, R.n_constraints, res); }
        cudaEventRecord(e1, d->stream);
        unsigned long long out[2];
        if (cudaMemcpyAsync(out, res, 16, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess || cudaStreamSynchronize(d->stream) != cudaSuccess) { rc = HZW_E_DEVICE; break; }
        if (n_bad) *n_bad = out[0];
        if (first_bad) *first_bad = out[1];
        if (ms) cudaEventElapsedTime(ms, e0, e1);
    } while (0);
    if (rc) fprintf(stderr, "hzw: CUDA error in the constraint check: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaFree(off); cudaFree(wire); cudaFree(coef); cudaFree(res); cudaFree(wtmp);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

// ---- streaming `.wtns` writer --------------------------------------------------------------------------------------
// The witness resident on the device (the gathered one under a layout map) goes to the file in 64 MiB chunks through two
// pinned buffers: chunk k + 1 crosses PCIe while chunk k is written -- the host never holds the 8 GB witness, and the
// file is complete when the call returns (`./circuit input.json witness.wtns`, /root/reference/tools/helpers/actions.js:139-143).
#include <fcntl.h>
#include <unistd.h>
int hz_cuda_stream_wtns(HzCuda* d, const char* path, const uint8_t* header, size_t header_len, uint64_t* d2h) {
    CK(cudaSetDevice(d->device));
    const Fr* w = d->n_out ? d->w_out : d->w;
    const uint64_t n = d->n_out ? d->n_out : d->plan.nvars;
    const size_t chunk = (size_t)64 << 20;
    const uint64_t total = 32ull * n;
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return HZW_E_IO;
    uint8_t* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int rc = 0;
    auto put = [&](const uint8_t* p, size_t len) { while (len) { ssize_t k = write(fd, p, len); if (k <= 0) return false; p += k; len -= (size_t)k; } return true; };
    do {
        if (!put(header, header_len)) { rc = HZW_E_IO; break; }
        for (int k = 0; k < 2; k++) if (cudaMallocHost(&buf[k], chunk) != cudaSuccess || cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) rc = HZW_E_DEVICE;
        if (rc) break;
        const uint64_t nchunks = (total + chunk - 1) / chunk;
        auto issue = [&](uint64_t k) {
            const uint64_t at = k * chunk, len = total - at < chunk ? total - at : chunk;
            return cudaMemcpyAsync(buf[k & 1], reinterpret_cast<const uint8_t*>(w) + at, len, cudaMemcpyDeviceToHost, d->stream) == cudaSuccess &&
                   cudaEventRecord(ev[k & 1], d->stream) == cudaSuccess;
        };
        if (nchunks && !issue(0)) { rc = HZW_E_DEVICE; break; }
        for (uint64_t k = 0; k < nchunks && !rc; k++) {
            if (cudaEventSynchronize(ev[k & 1]) != cudaSuccess) { rc = HZW_E_DEVICE; break; }
            if (k + 1 < nchunks && !issue(k + 1)) { rc = HZW_E_DEVICE; break; }      // crosses the link while chunk k is written
            const uint64_t at = k * chunk, len = total - at < chunk ? total - at : chunk;
            if (!put(buf[k & 1], len)) rc = HZW_E_IO;
        }
        cudaStreamSynchronize(d->stream);
    } while (0);
    if (close(fd) != 0 && !rc) rc = HZW_E_IO;
    for (int k = 0; k < 2; k++) { if (buf[k]) cudaFreeHost(buf[k]); if (ev[k]) cudaEventDestroy(ev[k]); }
    if (d2h) *d2h = total;
    return rc;
}
