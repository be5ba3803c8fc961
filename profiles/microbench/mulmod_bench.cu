// Micro-benchmark behind DESIGN.md section 4: cycles per canonical modular multiplication (fr_mul_chain) for one
// warp per SM sub-partition up to full occupancy, with one or two independent dependency chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mulmod_bench mulmod_bench.cu && ./mulmod_bench
#define HZ_DEVICE_ONLY
#include "../../circuits_b200/csrc/fr.cuh"
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP> __global__ void k_chain(Fr* out, int iters) {
    Fr x[ILP], y;
    for (int k = 0; k < ILP; k++) { x[k] = fr_from_u64(threadIdx.x + 3 + 7 * k); x[k].l[7] = 0x12345678u + k; }
    y = fr_from_u64(blockIdx.x + 11); y.l[6] = 0x9abcdef1u; y.l[7] = 0x0fedcba9u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = fr_mul_chain(x[k], y);
    }
    Fr r = x[0];
    for (int k = 1; k < ILP; k++) r = fr_add(r, x[k]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dot3(Fr* out, int iters) {
    Fr s[3], c[3];
    for (int k = 0; k < 3; k++) { s[k] = fr_from_u64(threadIdx.x + 3 + 7 * k); s[k].l[7] = 0x12345678u + k; c[k] = fr_from_u64(blockIdx.x + k); c[k].l[7] = 0x0fedcba9u; }
    for (int i = 0; i < iters; i++) { Fr r = fr_dot_mont_chain<3>(s, c); s[i % 3] = r; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s[0];
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    Fr* out; cudaMalloc(&out, 32ull * 148 * 64 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    printf("SMs %d, clock %d kHz\n", p.multiProcessorCount, p.clockRate);
    for (int warps = 4; warps <= 64; warps *= 2) {              // warps per SM (4 = one per sub-partition)
        for (int ilp = 1; ilp <= 2; ilp++) {
            dim3 grid(p.multiProcessorCount), block(32 * warps > 1024 ? 1024 : 32 * warps);
            if (32 * warps > 1024) grid.x *= (32 * warps) / 1024;
            if (ilp == 1) k_chain<1><<<grid, block>>>(out, 10); else k_chain<2><<<grid, block>>>(out, 10);
            cudaEventRecord(e0);
            if (ilp == 1) k_chain<1><<<grid, block>>>(out, iters); else k_chain<2><<<grid, block>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double cyc = ms * 1e-3 * 1.965e9;
            printf("warps/SM %2d ilp %d: %.3f ms, %.0f cycles per mulmod per warp (latency), %.1f cycles per mulmod per sub-partition (throughput)\n",
                   warps, ilp, ms, cyc / (iters * ilp) * 1.0 / 1.0 / 1.0 * 1.0 / 1.0 / 1.0 * 1.0, cyc / (iters * ilp * (warps / 4.0)));
        }
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        dim3 grid(p.multiProcessorCount), block(32 * warps);
        k_dot3<<<grid, block>>>(out, 10);
        cudaEventRecord(e0); k_dot3<<<grid, block>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double cyc = ms * 1e-3 * 1.965e9;
        printf("dot3 warps/SM %2d: %.0f cycles per row per warp, %.1f per sub-partition\n", warps, cyc / iters, cyc / (iters * (warps / 4.0)));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
