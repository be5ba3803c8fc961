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
