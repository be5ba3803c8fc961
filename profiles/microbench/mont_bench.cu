// Micro-benchmark of the two multipliers (round 2): the canonical Barrett product fr_mul_chain against the
// Montgomery-domain product fr_mont_mul (+ fr_mont_redc off the chain), as a dependent chain per thread, from one
// warp per SM sub-partition up to full occupancy; also checks the Montgomery routines against the Barrett one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mont_bench mont_bench.cu && ./mont_bench
#define HZ_DEVICE_ONLY
#include "../../circuits_b200/csrc/fr.cuh"
#include "../../circuits_b200/csrc/fr_montgen.inc"
#include <cstdio>
#include <cuda_runtime.h>

// R^2 mod p: x -> x~ = mont_mul(x, R2)
__device__ __forceinline__ Fr fr_R2() {
    Fr r; r.l[0] = 0xae216da7u; r.l[1] = 0x1bb8e645u; r.l[2] = 0xe35c59e3u; r.l[3] = 0x53fe3ab1u;
    r.l[4] = 0x53bb8085u; r.l[5] = 0x8c49833du; r.l[6] = 0x7f4e44a5u; r.l[7] = 0x0216d0b1u; return r;
}
template <int MODE> __global__ void k_chain(Fr* out, int iters) {
    Fr x = fr_from_u64(threadIdx.x + 3); x.l[7] = 0x12345678u;
    Fr y = fr_from_u64(blockIdx.x + 11); y.l[6] = 0x9abcdef1u; y.l[7] = 0x0fedcba9u;
    Fr acc = fr_zero();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) x = fr_mul_chain(x, y);                       // Barrett, canonical
        else if (MODE == 1) { Fr t; fr_mont_mul(t, x, y); x = t; }   // Montgomery chain only
        else { Fr t, c; fr_mont_mul(t, x, y); x = t; fr_mont_redc(c, t); acc.l[0] ^= c.l[i & 7]; }   // chain + canonical value of every product
    }
    x.l[0] ^= acc.l[0];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_check(int* bad, int n) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    Fr x = fr_from_u64(0x9e3779b97f4a7c15ull * (g + 1)); x.l[3] = 0xdeadbeefu * (g + 7); x.l[7] = (0x1234567u * (g + 3)) & 0x2fffffffu;
    Fr y = fr_from_u64(0xc2b2ae3d27d4eb4full * (g + 5)); y.l[5] = 0xcafef00du ^ g; y.l[7] = (0x7654321u * (g + 9)) & 0x2fffffffu;
    if (fr_geq_p(x)) fr_sub_p(x);
    if (fr_geq_p(y)) fr_sub_p(y);
    Fr want = fr_mul_chain(x, y);
    Fr xm, ym, pm, got;
    fr_mont_mul(xm, x, fr_R2()); fr_mont_mul(ym, y, fr_R2());
    fr_mont_mul(pm, xm, ym);
    fr_mont_redc(got, pm);
    Fr s[3] = {xm, ym, pm}, c[3] = {ym, xm, fr_R2()}, d, dg;    // dot3 = 2 x y R + pm R ... checked against the pieces
    fr_mont_dot3(d, s, c); fr_mont_redc(dg, d);
    Fr t1, t2; fr_mont_mul(t1, pm, fr_R2()); fr_mont_redc(t2, t1);           // pm R / R = x y R -> redc -> x y  ... t1 = pm~; t2 = pm = xyR mod p
    Fr w2 = fr_add(fr_add(fr_mul_chain(x, y), fr_mul_chain(x, y)), t2);      // canonical: 2xy*R? see below
    // d = (xm ym + ym xm + pm R2)/R = 2 xyR + pm R = 2 pm + pm~  (mod p);  redc(d) = 2xy + pm (mod p) where pm canonical-residue = xyR mod p
    Fr pmc = pm; while (fr_geq_p(pmc)) fr_sub_p(pmc);
    Fr want2 = fr_add(fr_add(want, want), pmc);
    (void)w2;
    if (!fr_eq(got, want) || !fr_eq(dg, want2)) atomicAdd(bad, 1);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    Fr* out; cudaMalloc(&out, 32ull * 148 * 64 * 1024);
    int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    k_check<<<64, 256>>>(bad, 64 * 256);
    int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("check: %d mismatches of %d (%s)\n", hb, 64 * 256, cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    const char* names[3] = {"barrett (canonical)", "montgomery chain", "montgomery chain + redc of every product"};
    for (int mode = 0; mode < 3; mode++)
        for (int warps = 4; warps <= 64; warps *= 2) {
            dim3 grid(p.multiProcessorCount), block(32 * warps > 1024 ? 1024 : 32 * warps);
            if (32 * warps > 1024) grid.x *= (32 * warps) / 1024;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k_chain<0><<<grid, block>>>(out, iters); else if (mode == 1) k_chain<1><<<grid, block>>>(out, iters); else k_chain<2><<<grid, block>>>(out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double cyc = ms * 1e-3 * 1.965e9;
            printf("%-42s warps/SM %2d: %8.0f cycles per product per warp (latency), %7.1f per sub-partition (throughput)\n", names[mode], warps, cyc / iters, cyc / (iters * (warps / 4.0)));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
