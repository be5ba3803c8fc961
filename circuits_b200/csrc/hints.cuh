// Data-stream inverse hints for the elliptic-curve ladders of EdDSAPoseidonVerifier
// (circomlib escalarmulany.circom / escalarmulfix.circom, reached from
// /root/reference/src/rollup-tx.circom:472).
//
// The templates work in affine Montgomery coordinates and divide once per point operation; every
// division waits for the point before it.  The points themselves do not need divisions: the
// affine formulas of MontgomeryDouble / MontgomeryAdd are rational maps, and here the same maps
// are evaluated on projective triples (U : V : W), u = U/W, v = V/W:
//     double:  lamda = N / Dd,  N = 3U^2 + 2A U W + W^2,  Dd = 2 V W        -> 1/(2v)      = W / (2V)
//     add:     lamda = N / Dn,  N = V2 W1 - V1 W2,        Dn = U2 W1 - U1 W2 -> 1/(u2 - u1) = W1 W2 / Dn
// Nothing here uses the curve equation, so the triples track the template's values for ANY input,
// including the off-curve points an L1 slot feeds the verifier (Ax from the dummy key, Ay from the
// deposited account: src/rollup-tx.circom:454-482).  All denominators of a component are then
// inverted with ONE field inversion (Montgomery's trick).  If one of them is zero (a degenerate
// input; the template's `1/0 = 0` is not a rational map), no stream is installed and the template
// falls back to its own divisions, so the witness is the same in every case.
#pragma once
#include "ctx.cuh"

struct MtPt { Fr U, V, W; };
#define HZ_MT_A 168698ull

// (u3, v3) = MontgomeryDouble(u, v), projectively: 17 products
HZ_HD MtPt mt_dbl(const MtPt& p) {
    const Fr A = fr_from_u64(HZ_MT_A);
    Fr UU = fr_mul(p.U, p.U), UW = fr_mul(p.U, p.W), WW = fr_mul(p.W, p.W);
    Fr AUW = fr_mul(A, UW);
    Fr N = fr_add(fr_add(fr_add(fr_add(UU, UU), UU), fr_add(AUW, AUW)), WW);
    Fr VW = fr_mul(p.V, p.W);
    Fr Dd = fr_add(VW, VW);
    Fr S = fr_mul(Dd, Dd);
    Fr G = fr_sub(fr_mul(fr_mul(N, N), p.W), fr_mul(fr_add(fr_mul(A, p.W), fr_add(p.U, p.U)), S));
    Fr DS = fr_mul(Dd, S);
    MtPt r;
    r.U = fr_mul(G, Dd);
    r.V = fr_sub(fr_mul(N, fr_sub(fr_mul(p.U, S), G)), fr_mul(p.V, DS));
    r.W = fr_mul(DS, p.W);
    return r;
}
// MontgomeryAdd(in1 = p1, in2 = p2), projectively: 16 products; num / den = 1/(u2 - u1)
HZ_HD MtPt mt_add(const MtPt& p1, const MtPt& p2, Fr& num, Fr& den) {
    const Fr A = fr_from_u64(HZ_MT_A);
    Fr U1W2 = fr_mul(p1.U, p2.W), U2W1 = fr_mul(p2.U, p1.W), V1W2 = fr_mul(p1.V, p2.W), V2W1 = fr_mul(p2.V, p1.W);
    Fr W12 = fr_mul(p1.W, p2.W);
    Fr N = fr_sub(V2W1, V1W2), Dn = fr_sub(U2W1, U1W2);
    num = W12; den = Dn;
    Fr S = fr_mul(Dn, Dn);
    Fr G = fr_sub(fr_mul(fr_mul(N, N), W12), fr_mul(fr_add(fr_mul(A, W12), fr_add(U1W2, U2W1)), S));
    Fr DS = fr_mul(Dn, S);
    MtPt r;
    r.U = fr_mul(G, Dn);
    r.V = fr_sub(fr_mul(N, fr_sub(fr_mul(U1W2, S), G)), fr_mul(V1W2, DS));
    r.W = fr_mul(DS, W12);
    return r;
}

// hint[k] = num[k] / den[k] for k < n with one inversion; false (nothing usable) if a den is zero.
// On the device the 8 lanes of a group run this redundantly on the same arrays: every location
// is either only read or only written with one value here, so the lanes cannot disturb each other.
HZ_HD bool hz_batch_divide(Fr* hint, const Fr* num, const Fr* den, Fr* pre, int n) {
    Fr acc = fr_one();
    bool ok = true;
    for (int k = 0; k < n; k++) {
        pre[k] = acc;
        if (fr_is_zero(den[k])) ok = false;
        acc = fr_mul(acc, den[k]);
    }
    Fr t = fr_inv(acc);                       // also when !ok: the call count must not depend on the data
    if (!ok) return false;
    for (int k = n - 1; k >= 0; k--) {
        hint[k] = fr_mul(num[k], fr_mul(t, pre[k]));
        t = fr_mul(t, den[k]);
    }
    return true;
}

HZ_HD void hz_hints_sync() {
#ifdef HZ_DEVICE_ONLY
    __syncwarp();
#endif
}

// SegmentMulAny(n) on the Edwards input (px, py), scalar bits e[0 .. n): hints in template order,
// per bit i = 0..n-2: doubler of D_i, then adder (in1 = D_{i+1}, in2 = acc_i)
template <int W> HZ_HD void hz_hints_segment_any(HzHints* h, const Fr& px, const Fr& py, const BVT<MCompute, W>& e, int n) {
    if (!h || !h->scratch || 2 * (n - 1) > HZ_HINT_MAX) return;
    hz_hints_sync();
    Fr* hint = h->scratch; Fr* out = hint + HZ_HINT_MAX; Fr* den = out + HZ_HINT_MAX; Fr* pre = den + HZ_HINT_MAX;
    // Edwards2Montgomery: u = (1+y)/(1-y), v = u/x  ->  ((1+y) x : (1+y) : (1-y) x)
    MtPt D;
    Fr y1 = fr_add(fr_one(), py);
    D.U = fr_mul(y1, px); D.V = y1; D.W = fr_mul(fr_sub(fr_one(), py), px);
    h->dat = nullptr; h->di = 0;
    if (fr_is_zero(D.W)) return;
    MtPt acc = D;
    for (int i = 0; i < n - 1; i++) {
        out[2 * i] = D.W;
        den[2 * i] = fr_add(D.V, D.V);
        D = mt_dbl(D);
        MtPt s = mt_add(D, acc, out[2 * i + 1], den[2 * i + 1]);
        if (e.bit(i + 1)) acc = s;
    }
    if (hz_batch_divide(hint, out, den, pre, 2 * (n - 1))) h->dat = hint;
}

// the accumulating adders of SegmentMulFix(nW) of segment `seg`: acc_0 = dblLast, acc_{i+1} =
// MontgomeryAdd(in1 = acc_i, in2 = window_i[sel_i]); the window entries (affine Montgomery points)
// come from the table recorded with the constants
template <int W> HZ_HD void hz_hints_segment_fix(HzHints* h, const BVT<MCompute, W>& e, int nW, int seg) {
    if (!h || !h->scratch || !h->edtab || nW > HZ_HINT_MAX) return;
    hz_hints_sync();
    Fr* hint = h->scratch; Fr* out = hint + HZ_HINT_MAX; Fr* den = out + HZ_HINT_MAX; Fr* pre = den + HZ_HINT_MAX;
    const Fr* tab = h->edtab + (seg == 0 ? 0 : HZ_EDTAB_SEG1);
    MtPt acc; acc.U = tab[nW * 16]; acc.V = tab[nW * 16 + 1]; acc.W = fr_one();
    h->dat = nullptr; h->di = 0;
    for (int i = 0; i < nW; i++) {
        int sel = (int)(e.bit(3 * i) | (e.bit(3 * i + 1) << 1) | (e.bit(3 * i + 2) << 2));
        MtPt w; w.U = tab[i * 16 + 2 * sel]; w.V = tab[i * 16 + 2 * sel + 1]; w.W = fr_one();
        acc = mt_add(acc, w, out[i], den[i]);
    }
    if (hz_batch_divide(hint, out, den, pre, nW)) h->dat = hint;
}
HZ_HD void hz_hints_done(HzHints* h) { if (h) { h->dat = nullptr; h->di = 0; } }

// record mode: an entry of the window tables (affine Montgomery point)
HZ_HD void hz_hints_record_point(HzHints* h, const Fr& u, const Fr& v) {
    if (!h || !h->edrec) return;
    h->edrec[h->ei++] = u;
    h->edrec[h->ei++] = v;
}

#ifdef HZ_DEVICE_ONLY
// ---- the device's way: all ladders of one signature at once, spread over the lanes of a group ----
// The doubling chain D_t = 2^t P runs through both segments of EscalarMulAny (Montgomery2Edwards
// followed by Edwards2Montgomery at the segment boundary is the identity on (u, v)), 253 doublings
// that nothing can shorten.  Everything else rides along in other lanes, one step behind:
//   lane 0  D_{t+1} = double(D_t)                                  t = 0..252
//   lane 1  segment 0 adder  (in1 = D_t, in2 = acc0), t = 1..147;  acc0 <- sum when e[t]
//   lane 2  segment 1 adder  (in1 = D_t, in2 = acc1), t = 149..253; acc1 starts as D_148
//   lane 3  EscalarMulFix segment 0 accumulating adder (in1 = accF, in2 = window_t[sel_t]), t = 0..82
// A doubling and an addition differ only in how the slope lamda = N / Dn is formed; the rest of
// MontgomeryAdd / MontgomeryDouble is the same map (u3 = lamda^2 - A - u1 - u2, v3 = lamda (u1 - u3) - v1),
// so after a short role-specific head all lanes run ONE instruction stream.  The 588 denominators
// are then inverted together: every lane takes a 74-element chunk of Montgomery's trick.
#define HZ_LAD_N 588
#define HZ_LAD_SCRATCH (4 * 600)
struct MtOps { Fr N, Dn, W12, X1, X2, Y1; };
__device__ __forceinline__ MtOps mt_head_dbl(const MtPt& p) {
    MtOps o;
    Fr UU = fr_mul_chain(p.U, p.U), UW = fr_mul_chain(p.U, p.W), WW = fr_mul_chain(p.W, p.W), VW = fr_mul_chain(p.V, p.W);
    Fr AUW = fr_mul_small(UW, (uint32_t)HZ_MT_A);
    o.N = fr_add(fr_add(fr_add(fr_add(UU, UU), UU), fr_add(AUW, AUW)), WW);
    o.Dn = fr_add(VW, VW);
    o.W12 = WW; o.X1 = UW; o.X2 = UW; o.Y1 = VW;
    return o;
}
__device__ __forceinline__ MtOps mt_head_add(const MtPt& p1, const MtPt& p2) {
    MtOps o;
    o.X1 = fr_mul_chain(p1.U, p2.W); o.X2 = fr_mul_chain(p2.U, p1.W);
    o.Y1 = fr_mul_chain(p1.V, p2.W);
    Fr Y2 = fr_mul_chain(p2.V, p1.W);
    o.W12 = fr_mul_chain(p1.W, p2.W);
    o.N = fr_sub(Y2, o.Y1); o.Dn = fr_sub(o.X2, o.X1);
    return o;
}
__device__ __forceinline__ MtPt mt_tail(const MtOps& o) {
    Fr S = fr_mul_chain(o.Dn, o.Dn);
    Fr NN = fr_mul_chain(o.N, o.N);
    Fr G = fr_sub(fr_mul_chain(NN, o.W12), fr_mul_chain(fr_add(fr_add(fr_mul_small(o.W12, (uint32_t)HZ_MT_A), o.X1), o.X2), S));
    Fr DS = fr_mul_chain(o.Dn, S);
    MtPt r;
    r.U = fr_mul_chain(G, o.Dn);
    r.V = fr_sub(fr_mul_chain(o.N, fr_sub(fr_mul_chain(o.X1, S), G)), fr_mul_chain(o.Y1, DS));
    r.W = fr_mul_chain(DS, o.W12);
    return r;
}
__device__ __forceinline__ Fr hz_shfl8_fr(const Fr& v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src, HZ_G);
    return r;
}
// position of ladder element `el` in the hint array, which is in the order the templates consume:
// [0, 294) segment 0 (doubler, adder per bit), [294] doublers[0], [295, 505) segment 1, [505, 588) fix adders
__device__ __forceinline__ int hz_lad_pos(int el) {
    if (el < 147) return 2 * el;
    if (el == 147) return 294;
    if (el < 253) return 295 + 2 * (el - 148);
    if (el < 400) return 2 * (el - 253) + 1;
    if (el < 505) return 295 + 2 * (el - 400) + 1;
    return el;
}
// px, py: the Edwards input of EscalarMulAny (after its zero-point substitution); e: the 254 scalar
// bits; s: the 253 bits of S.  Leaves h->lad_ok and, if set, the 588 hints in h->scratch[0 .. 588).
__device__ __noinline__ void hz_hints_eddsa_lanes(HzHints* h, const Fr& px, const Fr& py, const BVT<MCompute, 8>& e,
                                                 const BVT<MCompute, 8>& s) {
    h->lad_ok = 0;
    if (!h->scratch || !h->edtab) return;
    __syncwarp();
    Fr* hint = h->scratch; Fr* num = hint + 600; Fr* den = num + 600; Fr* pre = den + 600;
    const int lane = (int)hz_glane();
    MtPt D;
    {
        Fr y1 = fr_add(fr_one(), py);
        D.U = fr_mul_chain(y1, px); D.V = y1; D.W = fr_mul_chain(fr_sub(fr_one(), py), px);
    }
    bool ok = !fr_is_zero(D.W);
    MtPt acc = D;
    const Fr* tab = h->edtab;
    if (lane == 3) { acc.U = tab[83 * 16]; acc.V = tab[83 * 16 + 1]; acc.W = fr_one(); }
    for (int t = 0; t <= 253; t++) {
        if (lane == 2 && t == 148) acc = D;
        bool active;
        int el;
        MtOps o;
        Fr hn, hd;
        if (lane == 0) {
            active = t <= 252; el = t;
            o = mt_head_dbl(D);
            hn = D.W; hd = fr_add(D.V, D.V);
        } else {
            MtPt p1 = D, p2 = acc;
            if (lane == 3) {
                active = t < 83; el = 505 + t;
                int tt = active ? t : 0;
                int sel = (int)(s.bit(3 * tt) | (s.bit(3 * tt + 1) << 1) | (s.bit(3 * tt + 2) << 2));
                p1 = acc; p2.U = tab[tt * 16 + 2 * sel]; p2.V = tab[tt * 16 + 2 * sel + 1]; p2.W = fr_one();
            } else if (lane == 1) { active = t >= 1 && t <= 147; el = 253 + (t - 1); }
            else if (lane == 2) { active = t >= 149; el = 400 + (t - 149); }
            else { active = false; el = 0; }
            o = mt_head_add(p1, p2);
            hn = o.W12; hd = o.Dn;
        }
        MtPt r = mt_tail(o);
        if (active) {
            num[el] = hn; den[el] = hd;
            if (fr_is_zero(hd)) ok = false;
            if (lane == 0 || lane == 3) { if (lane == 0) D = r; else acc = r; }
            else if (e.bit(t)) acc = r;
        }
        // every lane follows lane 0's point
        D.U = hz_shfl8_fr(D.U, 0); D.V = hz_shfl8_fr(D.V, 0); D.W = hz_shfl8_fr(D.W, 0);
    }
    // any zero denominator in the group: no hints at all
    {
        unsigned bad = __ballot_sync(0xffffffffu, !ok);
        unsigned mine = (bad >> ((threadIdx.x & 31u) & ~7u)) & 0xffu;
        ok = mine == 0;
    }
    __syncwarp();
    // Montgomery's trick, a 74-element chunk per lane
    const int c0 = lane * 74, c1 = c0 + 74 < HZ_LAD_N ? c0 + 74 : HZ_LAD_N;
    Fr tot = fr_one();
    for (int k = c0; k < c1; k++) { pre[k] = tot; Fr dk = den[k]; tot = fr_mul_chain(tot, ok ? dk : fr_one()); }
    // inverse of the own chunk total = inverse of the product of all chunk totals times the other totals
    Fr all = fr_one(), others = fr_one();
    for (int m = 0; m < HZ_G; m++) {
        Fr tm = hz_shfl8_fr(tot, m);
        all = fr_mul_chain(all, tm);
        if (m != lane) others = fr_mul_chain(others, tm);
    }
    Fr inv = fr_mul_chain(fr_inv(all), others);
    if (ok) {
        for (int k = c1 - 1; k >= c0; k--) {
            hint[hz_lad_pos(k)] = fr_mul_chain(num[k], fr_mul_chain(inv, pre[k]));
            inv = fr_mul_chain(inv, den[k]);
        }
    }
    __syncwarp();
    h->lad_ok = ok ? 1 : 0;
}
#endif
