// BN254 scalar field Fr, 8 x 32-bit limbs, canonical (non-Montgomery) residues.
//
// Replaces the reference's generated field library (ffiasm `fr.asm`/`fr.cpp`, built by
// /root/reference/tools/helpers/actions.js:207-231 for the prime at :209) on the device.
// Every witness signal has to be emitted as a canonical 32-byte little-endian residue, so
// values are kept canonical and two multipliers are provided:
//   * fr_mul / fr_sqr  : signal x signal, 512-bit product + Barrett reduction (canonical in/out);
//   * fr_dot_mont      : sum_j s_j * c_j with the constants c_j pre-scaled by R = 2^256
//                        (Montgomery form) -- one Montgomery reduction for the whole row, result
//                        canonical.  This is the Poseidon MDS path.
// The same source compiles for the host (g++) so that the circuit templates can be checked
// on a CPU against the oracle; the product only ever runs the device instantiation.
#pragma once
#include <string.h>
#include <stdint.h>

#if defined(__CUDACC__) && defined(HZ_DEVICE_ONLY)
// CUDA translation unit: device code only (the host half of the library is compiled by g++)
#define HZ_HD __device__ __forceinline__
#define HZ_HDN static __device__ __noinline__
#define HZ_FN static __device__ __noinline__
#elif defined(__CUDACC__)
#define HZ_HD __host__ __device__ __forceinline__
#define HZ_HDN __host__ __device__ __noinline__
#define HZ_FN __host__ __device__ __noinline__
#else
#define HZ_HD inline
#define HZ_HDN inline
#define HZ_FN inline
#endif

struct Fr { uint32_t l[8]; };

// p = 21888242871839275222246405745257275088548364400416034343698204186575808495617
#define FR_P0 0xf0000001u
#define FR_P1 0x43e1f593u
#define FR_P2 0x79b97091u
#define FR_P3 0x2833e848u
#define FR_P4 0x8181585du
#define FR_P5 0xb85045b6u
#define FR_P6 0xe131a029u
#define FR_P7 0x30644e72u
#define FR_NP0 0xefffffffu   // -p^-1 mod 2^32

HZ_HD uint32_t fr_p(int i) {
    switch (i) {
        case 0: return FR_P0; case 1: return FR_P1; case 2: return FR_P2; case 3: return FR_P3;
        case 4: return FR_P4; case 5: return FR_P5; case 6: return FR_P6; default: return FR_P7;
    }
}
// mu = floor(2^512 / p), 9 limbs
HZ_HD uint32_t fr_mu(int i) {
    switch (i) {
        case 0: return 0xe1de9259u; case 1: return 0x20703a6bu; case 2: return 0x9e880ae6u;
        case 3: return 0x14485200u; case 4: return 0x80730147u; case 5: return 0xb074a586u;
        case 6: return 0x23a04a7au; case 7: return 0x4a474626u; default: return 0x00000005u;
    }
}

HZ_HD Fr fr_zero() { Fr r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
HZ_HD Fr fr_from_u64(uint64_t v) { Fr r = fr_zero(); r.l[0] = (uint32_t)v; r.l[1] = (uint32_t)(v >> 32); return r; }
HZ_HD Fr fr_one() { return fr_from_u64(1); }
HZ_HD Fr fr_P() { Fr r; for (int i = 0; i < 8; i++) r.l[i] = fr_p(i); return r; }
HZ_HD bool fr_is_zero(const Fr& a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.l[i]; return o == 0; }
HZ_HD bool fr_eq(const Fr& a, const Fr& b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i]; return o == 0; }
HZ_HD uint32_t fr_bit(const Fr& a, int i) { return (a.l[i >> 5] >> (i & 31)) & 1u; }
// 2^k, k <= 253 (< p)
HZ_HD Fr fr_pow2(int k) { Fr r = fr_zero(); r.l[k >> 5] = 1u << (k & 31); return r; }

// a >= b on raw 256-bit values
HZ_HD bool fr_geq_raw(const Fr& a, const Fr& b) {
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] > b.l[i]) return true;
        if (a.l[i] < b.l[i]) return false;
    }
    return true;
}
HZ_HD bool fr_geq_p(const Fr& a) {
    for (int i = 7; i >= 0; i--) {
        uint32_t p = fr_p(i);
        if (a.l[i] > p) return true;
        if (a.l[i] < p) return false;
    }
    return true;
}
// raw add / sub with carry/borrow out
HZ_HD uint32_t fr_add_raw(Fr& r, const Fr& a, const Fr& b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.l[i] + b.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
HZ_HD uint32_t fr_sub_raw(Fr& r, const Fr& a, const Fr& b) {
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)a.l[i] - (int64_t)b.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)(c & 1);
}
HZ_HD void fr_sub_p(Fr& r) {
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)r.l[i] - (int64_t)fr_p(i); r.l[i] = (uint32_t)c; c >>= 32; }
}
HZ_HD void fr_add_p(Fr& r) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)r.l[i] + fr_p(i); r.l[i] = (uint32_t)c; c >>= 32; }
}
HZ_HD Fr fr_add(const Fr& a, const Fr& b) {
    Fr r; fr_add_raw(r, a, b);             // a + b < 2p < 2^255: no carry out
    if (fr_geq_p(r)) fr_sub_p(r);
    return r;
}
HZ_HD Fr fr_sub(const Fr& a, const Fr& b) {
    Fr r; if (fr_sub_raw(r, a, b)) fr_add_p(r);
    return r;
}
HZ_HD Fr fr_neg(const Fr& a) { return fr_sub(fr_zero(), a); }
// circom's signed view: a "negative" value is one above (p-1)/2
HZ_HD bool fr_is_neg(const Fr& a) {
    // (p-1)/2
    const uint32_t h[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u,
                           0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] > h[i]) return true;
        if (a.l[i] < h[i]) return false;
    }
    return false;
}

// ---- 256 x 256 -> 512 product, operand scanning -------------------------------------------
HZ_HD void fr_mul_wide(uint32_t* t /*16*/, const Fr& a, const Fr& b) {
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.l[i] * b.l[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
}

// ---- Barrett: t (512 bits, < p^2 * small) -> t mod p ---------------------------------------
// q1 = t >> 224 (9 limbs), q2 = q1 * mu, qh = q2 >> 288, r = t - qh*p (mod 2^288), r -= p while >= p.
// Only columns >= 7 of q1*mu are formed (one guard limb): the estimate is low by at most 3.
HZ_HD Fr fr_barrett(const uint32_t* t) {
    uint32_t q2[18];
    for (int i = 0; i < 18; i++) q2[i] = 0;
    for (int i = 0; i < 9; i++) {
        uint64_t c = 0;
        uint32_t qi = t[7 + i];
        int j0 = (7 - i) > 0 ? (7 - i) : 0;
        for (int j = j0; j < 9; j++) {
            c += (uint64_t)qi * fr_mu(j) + q2[i + j];
            q2[i + j] = (uint32_t)c;
            c >>= 32;
        }
        q2[i + 9] = (uint32_t)c;
    }
    // r = t - qh * p on 9 limbs
    uint32_t qp[9];
    for (int i = 0; i < 9; i++) qp[i] = 0;
    for (int i = 0; i < 9; i++) {
        uint64_t c = 0;
        uint32_t qi = q2[9 + i];
        for (int j = 0; j < 8 && i + j < 9; j++) {
            c += (uint64_t)qi * fr_p(j) + qp[i + j];
            qp[i + j] = (uint32_t)c;
            c >>= 32;
        }
        if (i == 0) qp[8] = (uint32_t)c;
    }
    uint32_t r[9];
    int64_t b = 0;
    for (int i = 0; i < 9; i++) { b += (int64_t)t[i] - (int64_t)qp[i]; r[i] = (uint32_t)b; b >>= 32; }
    Fr out;
    for (int k = 0; k < 4; k++) {
        bool ge = r[8] != 0;
        if (!ge) {
            ge = true;
            for (int i = 7; i >= 0; i--) {
                uint32_t p = fr_p(i);
                if (r[i] > p) break;
                if (r[i] < p) { ge = false; break; }
            }
        }
        if (!ge) break;
        int64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (int64_t)r[i] - (int64_t)fr_p(i); r[i] = (uint32_t)c; c >>= 32; }
        r[8] = (uint32_t)((int64_t)r[8] + c);
    }
    for (int i = 0; i < 8; i++) out.l[i] = r[i];
    return out;
}

// ---- generated carry-chain products (tools/gen_frmul.py): the device multiplier ----------------
#include "fr_mulgen.inc"

// d = r - k (8 limbs); returns the borrow (0 / 1)
HZ_HD uint32_t fr_sub8(uint32_t* d, const uint32_t* r, const uint32_t* k) {
#if defined(__CUDA_ARCH__)
    uint32_t brw;
    asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20; "
        "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24; "
        "subc.u32 %8, 0, 0;"
        : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7]), "=&r"(brw)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(k[0]), "r"(k[1]), "r"(k[2]), "r"(k[3]), "r"(k[4]), "r"(k[5]), "r"(k[6]), "r"(k[7]));
    return brw & 1u;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)r[i] - (int64_t)k[i]; d[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)(c & 1);
#endif
}
// r < 4p (8 limbs)  ->  r mod p, branch-free: subtract 2p, then p, when they fit
HZ_HD void fr_fold4p(uint32_t* r) {
    const uint32_t p2[8] = {0xe0000002u, 0x87c3eb27u, 0xf372e122u, 0x5067d090u, 0x0302b0bau, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
    const uint32_t p1[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t d[8];
    uint32_t b = fr_sub8(d, r, p2);
    for (int i = 0; i < 8; i++) r[i] = b ? r[i] : d[i];
    b = fr_sub8(d, r, p1);
    for (int i = 0; i < 8; i++) r[i] = b ? r[i] : d[i];
}
// canonical x canonical -> canonical: 512-bit product, Barrett quotient from the top 9 limbs
// (error <= 3, so the remainder is below 4p < 2^256 and 8 limbs carry it)
HZ_HD Fr fr_mul_chain(const Fr& a, const Fr& b) {
    uint32_t t[16], q3[9], qp[8];
    fr_gen_mulwide(t, a, b);
    fr_gen_qmu(q3, t);
    fr_gen_qp(qp, q3);
    Fr r;
    fr_sub8(r.l, t, qp);
    fr_fold4p(r.l);
    return r;
}
// sum_{k<N} s[k] * cR[k] with cR in Montgomery form -> canonical (N <= 7: below 3p before folding)
template <int N> HZ_HD Fr fr_dot_mont_chain(const Fr* s, const Fr* cR) {
    uint32_t T[17], m[8], r[9];
    if (N == 1) { fr_gen_mulwide(T, s[0], cR[0]); T[16] = 0; }
    else if (N == 2) fr_gen_dot2(T, s, cR);
    else if (N == 3) fr_gen_dot3(T, s, cR);
    else if (N == 4) fr_gen_dot4(T, s, cR);
    else if (N == 5) fr_gen_dot5(T, s, cR);
    else if (N == 6) fr_gen_dot6(T, s, cR);
    else fr_gen_dot7(T, s, cR);
    fr_gen_mlow(m, T);
    fr_gen_mpadd(r, m, T);
    Fr o;
    for (int i = 0; i < 8; i++) o.l[i] = r[i];
    fr_fold4p(o.l);
    return o;
}

#ifdef HZ_DEVICE_ONLY
// Most products in the glue of the circuits have a boolean signal on one side (selectors, state
// machine flags, switchers): those are selections, not multiplications.
HZ_HDN Fr fr_mul(const Fr& a, const Fr& b) {
    const uint32_t bh = b.l[1] | b.l[2] | b.l[3] | b.l[4] | b.l[5] | b.l[6] | b.l[7];
    if (bh == 0 && b.l[0] <= 1u) return b.l[0] ? a : b;
    const uint32_t ah = a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7];
    if (ah == 0 && a.l[0] <= 1u) return a.l[0] ? b : a;
    return fr_mul_chain(a, b);
}
// k < 2^20: eight multiply-adds, the quotient from the top limbs, one conditional subtraction
HZ_HD Fr fr_mul_small(const Fr& x, uint32_t k) {
    uint32_t t[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)x.l[i] * k; t[i] = (uint32_t)c; c >>= 32; }
    t[8] = (uint32_t)c;
    // q = floor(t / p) or one less: (t >> 224) * floor(2^288 / p) >> 64
    const uint64_t thi = ((uint64_t)t[8] << 32) | t[7];
    const uint32_t q = (uint32_t)__umul64hi(thi, 0x54a474626ull);
    const uint32_t p1[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t qp[8];
    c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)p1[i] * q; qp[i] = (uint32_t)c; c >>= 32; }
    Fr r;
    fr_sub8(r.l, t, qp);                     // the remainder is below 2p < 2^256: 8 limbs carry it
    uint32_t d[8];
    uint32_t b = fr_sub8(d, r.l, p1);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = b ? r.l[i] : d[i];
    return r;
}
HZ_HDN Fr fr_mul_u64_dev(const Fr& a, uint64_t k) {
    if (k < (1ull << 20)) return fr_mul_small(a, (uint32_t)k);
    return fr_mul_chain(a, fr_from_u64(k));
}
#elif defined(__SIZEOF_INT128__) && defined(__x86_64__)
// Host (the CPU port, i.e. the CPU baseline of the benchmark): 4 x 64-bit limbs with 128-bit products, the arithmetic
// the reference's ffiasm field uses (4 x 4 64-bit Montgomery CIOS).  Witness entries are canonical
// residues, so a product of two of them is two Montgomery steps: redc(a b) = a b / R, redc(that * R^2) = a b; values
// 0 and 1 are handled as the reference's "short" elements are, without a multiplication.
typedef unsigned __int128 hz_u128;
static const uint64_t HZ_P64[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t HZ_R2_64[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x216d0b17f4e44a5ull};
#define HZ_NINV64 0xc2e1f593efffffffull
static inline void hz_mul64_acc(uint64_t* T /* 9 limbs */, const uint64_t* a, const uint64_t* b) {     // T += a * b
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) { hz_u128 x = (hz_u128)a[i] * b[j] + T[i + j] + c; T[i + j] = (uint64_t)x; c = (uint64_t)(x >> 64); }
        for (int k = i + 4; c && k < 9; k++) { hz_u128 x = (hz_u128)T[k] + c; T[k] = (uint64_t)x; c = (uint64_t)(x >> 64); }
    }
}
// T (9 limbs, below 8 p^2) -> T / 2^256 mod p, canonical
static inline void hz_redc64(uint64_t* T, uint64_t* out) {
    for (int i = 0; i < 4; i++) {
        const uint64_t m = T[i] * HZ_NINV64;
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) { hz_u128 x = (hz_u128)m * HZ_P64[j] + T[i + j] + c; T[i + j] = (uint64_t)x; c = (uint64_t)(x >> 64); }
        for (int k = i + 4; c && k < 9; k++) { hz_u128 x = (hz_u128)T[k] + c; T[k] = (uint64_t)x; c = (uint64_t)(x >> 64); }
    }
    uint64_t r[5] = {T[4], T[5], T[6], T[7], T[8]};
    for (;;) {                                   // below 3 p: at most two subtractions
        uint64_t d[4]; uint64_t bw = 0;
        for (int j = 0; j < 4; j++) { hz_u128 x = (hz_u128)r[j] - HZ_P64[j] - bw; d[j] = (uint64_t)x; bw = (uint64_t)(x >> 64) & 1u; }
        if (r[4] == 0 && bw) break;
        r[4] -= bw;
        for (int j = 0; j < 4; j++) r[j] = d[j];
    }
    for (int j = 0; j < 4; j++) out[j] = r[j];
}
static inline Fr hz_mont_pair64(const uint64_t* T0 /* 9 limbs, consumed */) {
    uint64_t T[9], u[4];
    for (int i = 0; i < 9; i++) T[i] = T0[i];
    hz_redc64(T, u);
    for (int i = 0; i < 9; i++) T[i] = 0;
    hz_mul64_acc(T, u, HZ_R2_64);
    hz_redc64(T, u);
    Fr r; memcpy(r.l, u, 32);
    return r;
}
HZ_HDN Fr fr_mul(const Fr& a, const Fr& b) {
    const uint32_t bh = b.l[1] | b.l[2] | b.l[3] | b.l[4] | b.l[5] | b.l[6] | b.l[7];
    if (bh == 0 && b.l[0] <= 1u) return b.l[0] ? a : b;
    const uint32_t ah = a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7];
    if (ah == 0 && a.l[0] <= 1u) return a.l[0] ? b : a;
    uint64_t x[4], y[4], T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    memcpy(x, a.l, 32); memcpy(y, b.l, 32);
    hz_mul64_acc(T, x, y);
    return hz_mont_pair64(T);
}
#define HZ_HOST_MONT64 1
#else
HZ_HDN Fr fr_mul(const Fr& a, const Fr& b) {
    uint32_t t[16];
    fr_mul_wide(t, a, b);
    return fr_barrett(t);
}
#endif
HZ_HD Fr fr_sqr(const Fr& a) { return fr_mul(a, a); }
#ifdef HZ_DEVICE_ONLY
HZ_HD Fr fr_mul_u64(const Fr& a, uint64_t k) { return fr_mul_u64_dev(a, k); }
#else
HZ_HD Fr fr_mul_u64(const Fr& a, uint64_t k) { return fr_mul(a, fr_from_u64(k)); }
#endif

// ---- Montgomery row: sum_{j<n} s[j] * cR[j]  (cR = c * 2^256 mod p)  -> canonical -----------
// n <= 7: the unreduced sum stays below 7 p^2 < 2^511.
HZ_HDN Fr fr_dot_mont(const Fr* s, const Fr* cR, int n) {
#ifdef HZ_HOST_MONT64
    {
        // host: 64-bit limbs, n products accumulated on 9 limbs, one Montgomery reduction (the constants carry the R)
        uint64_t T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, x[4], y[4], u[4];
        for (int k = 0; k < n; k++) { memcpy(x, s[k].l, 32); memcpy(y, cR[k].l, 32); hz_mul64_acc(T, x, y); }
        hz_redc64(T, u);
        Fr r; memcpy(r.l, u, 32);
        return r;
    }
#endif
    uint32_t t[17];
    for (int i = 0; i < 17; i++) t[i] = 0;
    for (int k = 0; k < n; k++) {
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
            uint32_t ai = s[k].l[i];
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)ai * cR[k].l[j] + t[i + j];
                t[i + j] = (uint32_t)c;
                c >>= 32;
            }
            // propagate
            for (int j = i + 8; c != 0 && j < 17; j++) { c += t[j]; t[j] = (uint32_t)c; c >>= 32; }
        }
    }
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * FR_NP0;
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)m * fr_p(j) + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        for (int j = i + 8; c != 0 && j < 17; j++) { c += t[j]; t[j] = (uint32_t)c; c >>= 32; }
    }
    uint32_t hi = t[16];
    Fr r;
    for (int i = 0; i < 8; i++) r.l[i] = t[8 + i];
    for (int k = 0; k < 4; k++) {
        if (hi == 0 && !fr_geq_p(r)) break;
        int64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (int64_t)r.l[i] - (int64_t)fr_p(i); r.l[i] = (uint32_t)c; c >>= 32; }
        hi = (uint32_t)((int64_t)hi + c);
    }
    return r;
}

// ---- inversion: binary extended Euclid, inv(0) = 0 (circom's `1/0` in `<--` evaluates to 0
//      in the wasm/C field libraries; IsZero relies on it) -----------------------------------
HZ_HD void fr_shr1(Fr& a, uint32_t top) {
    for (int i = 0; i < 7; i++) a.l[i] = (a.l[i] >> 1) | (a.l[i + 1] << 31);
    a.l[7] = (a.l[7] >> 1) | (top << 31);
}
HZ_HD void fr_half_mod(Fr& x) {         // x/2 mod p
    if (x.l[0] & 1u) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (uint64_t)x.l[i] + fr_p(i); x.l[i] = (uint32_t)c; c >>= 32; }
        fr_shr1(x, (uint32_t)c);
    } else {
        fr_shr1(x, 0);
    }
}
#ifdef HZ_HOST_MONT64
// host: the same binary extended Euclid on 64-bit limbs, runs of trailing zeros taken in one go
static inline uint64_t hz_sub64(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t bw = 0;
    for (int j = 0; j < 4; j++) { hz_u128 x = (hz_u128)a[j] - b[j] - bw; r[j] = (uint64_t)x; bw = (uint64_t)(x >> 64) & 1u; }
    return bw;
}
static inline void hz_half_mod64(uint64_t* x) {
    uint64_t top = 0;
    if (x[0] & 1u) { uint64_t c = 0; for (int j = 0; j < 4; j++) { hz_u128 t = (hz_u128)x[j] + HZ_P64[j] + c; x[j] = (uint64_t)t; c = (uint64_t)(t >> 64); } top = c; }
    for (int j = 0; j < 3; j++) x[j] = (x[j] >> 1) | (x[j + 1] << 63);
    x[3] = (x[3] >> 1) | (top << 63);
}
static inline void hz_shr1_64(uint64_t* x) { for (int j = 0; j < 3; j++) x[j] = (x[j] >> 1) | (x[j + 1] << 63); x[3] >>= 1; }
static inline bool hz_is_one64(const uint64_t* x) { return x[0] == 1 && (x[1] | x[2] | x[3]) == 0; }
static inline bool hz_geq64(const uint64_t* a, const uint64_t* b) {
    for (int j = 3; j >= 0; j--) { if (a[j] > b[j]) return true; if (a[j] < b[j]) return false; }
    return true;
}
static inline void hz_submod64(uint64_t* x, const uint64_t* y) {          // x = x - y mod p
    if (hz_sub64(x, x, y)) { uint64_t c = 0; for (int j = 0; j < 4; j++) { hz_u128 t = (hz_u128)x[j] + HZ_P64[j] + c; x[j] = (uint64_t)t; c = (uint64_t)(t >> 64); } }
}
HZ_HDN Fr fr_inv(const Fr& a) {
    if (fr_is_zero(a)) return a;
    uint64_t u[4], v[4] = {HZ_P64[0], HZ_P64[1], HZ_P64[2], HZ_P64[3]}, x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
    memcpy(u, a.l, 32);
    while (!hz_is_one64(u) && !hz_is_one64(v)) {
        while (!(u[0] & 1u)) { hz_shr1_64(u); hz_half_mod64(x1); }
        while (!(v[0] & 1u)) { hz_shr1_64(v); hz_half_mod64(x2); }
        if (hz_geq64(u, v)) { hz_sub64(u, u, v); hz_submod64(x1, x2); }
        else { hz_sub64(v, v, u); hz_submod64(x2, x1); }
    }
    Fr r; memcpy(r.l, hz_is_one64(u) ? x1 : x2, 32);
    return r;
}
#elif !defined(HZ_DEVICE_ONLY)
HZ_HDN Fr fr_inv(const Fr& a) {
    if (fr_is_zero(a)) return a;
    Fr u = a, v = fr_P(), x1 = fr_one(), x2 = fr_zero();
    Fr one = fr_one();
    while (!fr_eq(u, one) && !fr_eq(v, one)) {
        if (!(u.l[0] & 1u)) { fr_shr1(u, 0); fr_half_mod(x1); }
        else if (!(v.l[0] & 1u)) { fr_shr1(v, 0); fr_half_mod(x2); }
        else if (fr_geq_raw(u, v)) { fr_sub_raw(u, u, v); x1 = fr_sub(x1, x2); }
        else { fr_sub_raw(v, v, u); x2 = fr_sub(x2, x1); }
    }
    return fr_eq(u, one) ? x1 : x2;
}
#else
// ---- device inversion --------------------------------------------------------------------------
// (1) the same binary extended Euclid, on registers with carry-chain PTX;
// (2) warp-cooperative batching: on the device a warp is four 8-lane groups, each group holding one
//     value in all its lanes.  When the whole warp arrives together the four values are inverted
//     with ONE Euclid run (Montgomery's trick over the groups, two shuffle rounds), which every
//     lane executes on the same number -- no divergence.  If the warp is not whole at the call
//     (a group took a fallback path), every group inverts on its own.
__device__ __forceinline__ uint32_t hz_add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t cy;
    asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20; "
        "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24; "
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cy)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return cy;
}
__device__ __forceinline__ void hz_shr8(uint32_t* a, uint32_t top) {
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = __funnelshift_r(a[i], a[i + 1], 1);
    a[7] = __funnelshift_r(a[7], top, 1);
}
// x <- x / 2 mod p
__device__ __forceinline__ void hz_half8(uint32_t* x) {
    const uint32_t p1[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    if (x[0] & 1u) {
        uint32_t t[8];
        uint32_t cy = hz_add8(t, x, p1);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = t[i];
        hz_shr8(x, cy);
    } else {
        hz_shr8(x, 0);
    }
}
// x <- x - y mod p
__device__ __forceinline__ void hz_submod8(uint32_t* x, const uint32_t* y) {
    const uint32_t p1[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t d[8], e[8];
    uint32_t b = fr_sub8(d, x, y);
    hz_add8(e, d, p1);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = b ? e[i] : d[i];
}
__device__ __forceinline__ bool hz_is_one8(const uint32_t* a) {
    return (a[0] == 1u) && ((a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0u);
}
// 2^-k mod p, k = 0..512: turns Kaliski's almost inverse a^-1 2^k into the inverse with one multiplication
__device__ const Fr hz_inv2k_dev[513] = {
#include "fr_inv2k.inc"
};
__device__ __forceinline__ void hz_shl8(uint32_t* a) {
#pragma unroll
    for (int i = 7; i > 0; i--) a[i] = __funnelshift_l(a[i - 1], a[i], 1);
    a[0] <<= 1;
}
// a != 0, a < p.  Kaliski's almost Montgomery inverse (1995): the same subtract-and-halve walk on (u, v) as the
// binary extended Euclid, but the cofactors are only doubled and added -- plain 256-bit shifts and additions
// (r, s < 2p < 2^255) instead of a modular halving / modular subtraction per step -- and the 2^k they pick up
// (254 <= k <= 508) is divided out at the end: a^-1 = (p - r) 2^-k.
static __device__ __noinline__ Fr fr_inv_euclid(const Fr& a) {
    const uint32_t p1[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t u[8], v[8], r[8], sv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = p1[i]; v[i] = a.l[i]; r[i] = 0u; sv[i] = i == 0 ? 1u : 0u; }
    int k = 0;
    while ((v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) != 0u) {
        if (!(u[0] & 1u)) { hz_shr8(u, 0); hz_shl8(sv); }
        else if (!(v[0] & 1u)) { hz_shr8(v, 0); hz_shl8(r); }
        else {
            uint32_t d[8], t[8];
            const uint32_t b = fr_sub8(d, u, v);
            const uint32_t dz = d[0] | d[1] | d[2] | d[3] | d[4] | d[5] | d[6] | d[7];
            if (!b && dz != 0u) {                    // u > v
                hz_shr8(d, 0);
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = d[i];
                hz_add8(t, r, sv);
#pragma unroll
                for (int i = 0; i < 8; i++) r[i] = t[i];
                hz_shl8(sv);
            } else {                                 // v >= u
                fr_sub8(d, v, u);
                hz_shr8(d, 0);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = d[i];
                hz_add8(t, sv, r);
#pragma unroll
                for (int i = 0; i < 8; i++) sv[i] = t[i];
                hz_shl8(r);
            }
        }
        k++;
    }
    uint32_t d[8];
    if (!fr_sub8(d, r, p1)) {
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = d[i];
    }
    Fr x;
    fr_sub8(x.l, p1, r);
    return fr_mul_chain(x, hz_inv2k_dev[k]);
}
// a != 0: the binary extended Euclid it replaces (kept for reference)
static __device__ __noinline__ Fr fr_inv_euclid_binary(const Fr& a) {
    uint32_t u[8], v[8], x1[8], x2[8];
    const uint32_t p1[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = a.l[i]; v[i] = p1[i]; x1[i] = i == 0 ? 1u : 0u; x2[i] = 0u; }
    while (!hz_is_one8(u) && !hz_is_one8(v)) {
        if (!(u[0] & 1u)) { hz_shr8(u, 0); hz_half8(x1); }
        else if (!(v[0] & 1u)) { hz_shr8(v, 0); hz_half8(x2); }
        else {
            uint32_t d[8];
            uint32_t b = fr_sub8(d, u, v);
            if (!b) {
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = d[i];
                hz_submod8(x1, x2);
            } else {
                fr_sub8(d, v, u);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = d[i];
                hz_submod8(x2, x1);
            }
        }
    }
    Fr r;
    bool fu = hz_is_one8(u);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = fu ? x1[i] : x2[i];
    return r;
}
__device__ __forceinline__ Fr hz_shfl_xor_fr(const Fr& v, int m) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], m);
    return r;
}
static __device__ __noinline__ Fr fr_inv(const Fr& a) {
    const bool z = fr_is_zero(a);
    Fr x = a;
    if (z) x.l[0] = 1u;
    if (__activemask() != 0xffffffffu) {
        Fr r = fr_inv_euclid(x);
        if (z) r.l[0] = 0u;
        return r;
    }
    Fr v1 = hz_shfl_xor_fr(x, 8);            // the neighbouring group's value
    Fr p2 = fr_mul(x, v1);
    Fr q2 = hz_shfl_xor_fr(p2, 16);          // the other pair's product
    Fr t = fr_inv_euclid(fr_mul(p2, q2));    // the same number in all 32 lanes
    Fr r = fr_mul(fr_mul(t, q2), v1);
    if (z) r = fr_from_u64(0);
    return r;
}
// 32 different values, one per lane, inverted with ONE Euclid run (Montgomery's trick across the warp with two
// shuffle scans); lanes that arrive without the whole warp invert on their own.  inv(0) = 0.
static __device__ __noinline__ Fr fr_inv_lanes32(const Fr& a) {
    const bool z = fr_is_zero(a);
    Fr x = a;
    if (z) x.l[0] = 1u;
    if (__activemask() != 0xffffffffu) {
        Fr r = fr_inv_euclid(x);
        if (z) r.l[0] = 0u;
        return r;
    }
    const int lane = (int)(threadIdx.x & 31u);
    Fr below = x, above = x;                  // inclusive products over the lanes up to / from this one
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.l[i] = __shfl_up_sync(0xffffffffu, below.l[i], d);
        if (lane >= d) below = fr_mul_chain(below, o);
#pragma unroll
        for (int i = 0; i < 8; i++) o.l[i] = __shfl_down_sync(0xffffffffu, above.l[i], d);
        if (lane + d < 32) above = fr_mul_chain(above, o);
    }
    Fr all, lo, hi;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        all.l[i] = __shfl_sync(0xffffffffu, below.l[i], 31);
        lo.l[i] = __shfl_up_sync(0xffffffffu, below.l[i], 1);
        hi.l[i] = __shfl_down_sync(0xffffffffu, above.l[i], 1);
    }
    if (lane == 0) lo = fr_one();
    if (lane == 31) hi = fr_one();
    Fr r = fr_mul_chain(fr_inv_euclid(all), fr_mul_chain(lo, hi));
    if (z) r = fr_zero();
    return r;
}
#endif
HZ_HD Fr fr_div(const Fr& a, const Fr& b) { return fr_mul(a, fr_inv(b)); }

// a^e, e given as 8 limbs
HZ_HDN Fr fr_pow(const Fr& a, const uint32_t* e, int nbits) {
    Fr r = fr_one();
    for (int i = nbits - 1; i >= 0; i--) {
        r = fr_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1u) r = fr_mul(r, a);
    }
    return r;
}
