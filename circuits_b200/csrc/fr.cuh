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
#include <stdint.h>

#if defined(__CUDACC__) && defined(HZ_DEVICE_ONLY)
// CUDA translation unit: device code only (the host half of the library is compiled by g++)
#define HZ_HD __device__ __forceinline__
#define HZ_HDN __device__ __noinline__
#define HZ_FN __device__ __noinline__
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

HZ_HDN Fr fr_mul(const Fr& a, const Fr& b) {
    uint32_t t[16];
    fr_mul_wide(t, a, b);
    return fr_barrett(t);
}
HZ_HD Fr fr_sqr(const Fr& a) { return fr_mul(a, a); }
HZ_HD Fr fr_mul_u64(const Fr& a, uint64_t k) { return fr_mul(a, fr_from_u64(k)); }

// ---- Montgomery row: sum_{j<n} s[j] * cR[j]  (cR = c * 2^256 mod p)  -> canonical -----------
// n <= 7: the unreduced sum stays below 7 p^2 < 2^511.
HZ_HDN Fr fr_dot_mont(const Fr* s, const Fr* cR, int n) {
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
