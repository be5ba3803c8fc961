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
template <int W> HZ_HD void hz_hints_segment_fix(HzHints* h, const BVT<MCompute, W>& e, int nW, int seg, Fr* area = nullptr) {
    if (!h || !h->scratch || !h->edtab || nW > HZ_HINT_MAX) return;
    hz_hints_sync();
    Fr* hint = area ? area : h->scratch; Fr* out = hint + HZ_HINT_MAX; Fr* den = out + HZ_HINT_MAX; Fr* pre = den + HZ_HINT_MAX;
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
// ---- the device's way: the ladders of a signature walked projectively by k_ed_ladder (k_ladder.cu) ----
// The doubling chain D_t = 2^t P runs through both segments of EscalarMulAny (Montgomery2Edwards
// followed by Edwards2Montgomery at the segment boundary is the identity on (u, v)), 253 doublings
// that nothing can shorten.  The other ladders follow it:
//   D  D_{t+1} = double(D_t)                                       t = 0..252
//   A  segment 0 adder  (in1 = D_t, in2 = acc0), t = 1..147;  acc0 <- sum when e[t]
//      segment 1 adder  (in1 = D_t, in2 = acc1), t = 149..253; acc1 starts as D_148
//   F  EscalarMulFix segment 0 accumulating adder (in1 = accF, in2 = window_t[sel_t]), t = 0..82
// and the 588 slope denominators (plus the W of the captured points) are inverted together by k_ed_hints.
#define HZ_LAD_N 588
// ---- what the ladder leaves for k_ed_fill ---------------------------------------------------------------
// The witness entries of the ladder steps (BitElementMulAny x 252, the 83 accumulating adders of
// EscalarMulFix) are a fixed function of the step's input points and its two hints, so once the ladder has
// run they can be written by independent threads instead of by the lane group that walked it: k_ed_fill
// gives every CHUNK of HZ_FILL_CH consecutive steps to one thread.  A chunk starts from affine points, so
// the ladder also records the projective points at the chunk starts (and the final ones, which the
// templates continue from) and their W join the one batched inversion: "extra" element ex holds 1/W.
//   ex  0..18  D at the chunk starts of segment 0        19..37  its accumulator there
//      38..51  D at the chunk starts of segment 1        52..65  its accumulator there
//      66..76  accumulator of the fixed-base adders at their chunk starts
//      77 D_147 (dbl of segment 0)   78, 79 final accumulators of segments 0, 1   80 final fixed-base accumulator
#define HZ_FILL_CH 8
#define HZ_FILL_C0 19                      /* chunks of segment 0: 147 steps */
#define HZ_FILL_C1 14                      /* chunks of segment 1: 105 steps */
#define HZ_FILL_CF 11                      /* chunks of the 83 fixed-base adders */
#define HZ_FILL_CHUNKS (HZ_FILL_C0 + HZ_FILL_C1 + HZ_FILL_CF)
#define HZ_LAD_EX 81
// the points k_ed_ladder leaves per signature for k_ed_conv, 3 Fr (X, Y, Z) each: D_t (t = 0..253), the segment
// accumulator before step t (t = 1..253), the fixed-base accumulator before step t (t = 0..82), the two final accumulators
#define HZ_PT_D 0
#define HZ_PT_A 762
#define HZ_PT_F 1521
#define HZ_PT_FIN 1770
#define HZ_PTS 1776
#define HZ_LAD_TOT (HZ_LAD_N + HZ_LAD_EX)
// scratch of a lane group: hint[704] num[704] den[704] pre[704] vv[96] aff[176], then the fill record
#define HZ_SCR_ARR 704
#define HZ_SCR_VV (4 * HZ_SCR_ARR)
#define HZ_SCR_AFF (HZ_SCR_VV + 96)
#define HZ_SCR_REC (HZ_SCR_AFF + 176)
// per signature, written by lane 0 of k_eddsa: whether the ladder data is usable, where the three runs of
// step entries start in the witness, the scalar bits
struct HzFillRec { uint32_t ok, base_seg0, base_seg1, base_fix; uint32_t e[8]; uint32_t s[8]; uint32_t base_win; uint32_t cur; uint32_t pad[10]; };
// what the first half of the verifier (k_eddsa) leaves for the second (k_eddsa_tail), after the fill record:
// the point 8 A, `enabled`, R8 (5 field elements); the witness cursor where the second half starts is rec.cur
#define HZ_SCR_MID (HZ_SCR_REC + 4)
// position of ladder element `el` in the hint array, which is in the order the templates consume:
// [0, 294) segment 0 (doubler, adder per bit), [294] doublers[0], [295, 505) segment 1, [505, 588) fix adders
__host__ __device__ __forceinline__ int hz_lad_pos(int el) {
    if (el < 147) return 2 * el;
    if (el == 147) return 294;
    if (el < 253) return 295 + 2 * (el - 148);
    if (el < 400) return 2 * (el - 253) + 1;
    if (el < 505) return 295 + 2 * (el - 400) + 1;
    return el;
}

// ---- k_ed_fill: the entries of the ladder steps from the ladder's data ------------------------------------
// One thread per chunk of HZ_FILL_CH consecutive steps.  Entries per step (templates_circomlib.cuh):
//   BitElementMulAny: doubler.out[2], doubler.lamda, doubler.x1_2, adder.out[2], adder.lamda, selector.out[2]
//   accumulating adder of SegmentMulFix: out[2], lamda
// computed with the formulas of MontgomeryDouble / MontgomeryAdd / Multiplexor2, the divisions taking the
// ladder's hints (each keeps its `===` check when the sanity check is on).
struct AffPt { Fr u, v; };
__device__ __forceinline__ void hz_fill_err(HzErrBuf* err, uint32_t cur, uint32_t line, const Fr& lhs, const Fr& rhs) {
    uint32_t k = atomicAdd(&err->count, 1u);
    if (k < HZ_MAX_ERR) { err->rec[k].cur = cur; err->rec[k].line = line; err->rec[k].lhs = lhs; err->rec[k].rhs = rhs; }
}
// in1 = a, in2 = b, hint = 1 / (b.u - a.u); stores out[2], lamda at w[0..3)
__device__ __forceinline__ AffPt hz_fill_madd(Fr* w, uint32_t at, HzErrBuf* err, const AffPt& a, const AffPt& b, const Fr& hint) {
    const Fr dy = fr_sub(b.v, a.v);
    const Fr lam = fr_mul_chain(dy, hint);
    if (err) { const Fr lhs = fr_mul_chain(lam, fr_sub(b.u, a.u)); if (!fr_eq(lhs, dy)) hz_fill_err(err, at + 3, __LINE__, lhs, dy); }
    AffPt r;
    r.u = fr_sub(fr_sub(fr_sub(fr_mul_chain(lam, lam), fr_from_u64(HZ_MT_A)), a.u), b.u);
    r.v = fr_sub(fr_mul_chain(lam, fr_sub(a.u, r.u)), a.v);
    hz_store_fr(&w[at], r.u); hz_store_fr(&w[at + 1], r.v); hz_store_fr(&w[at + 2], lam);
    return r;
}
__device__ __forceinline__ void hz_fill_seg(Fr* w, HzErrBuf* err, const Fr* hint, const Fr* aff, const HzFillRec& rec, int sg, int k) {
    const int nsteps = sg == 0 ? 147 : 105, from = sg == 0 ? 0 : 148;
    const int i0 = k * HZ_FILL_CH, i1 = i0 + HZ_FILL_CH < nsteps ? i0 + HZ_FILL_CH : nsteps;
    const Fr* hp = hint + (sg == 0 ? 0 : 295);
    const int exD = sg == 0 ? k : 38 + k, exA = sg == 0 ? 19 + k : 52 + k;
    AffPt D, acc;
    D.u = aff[2 * exD]; D.v = aff[2 * exD + 1]; acc.u = aff[2 * exA]; acc.v = aff[2 * exA + 1];
    uint32_t at = (sg == 0 ? rec.base_seg0 : rec.base_seg1) + 9u * (uint32_t)i0;
    for (int i = i0; i < i1; i++, at += 9) {
        const int eb = from + i + 1;
        const bool sel = ((rec.e[eb >> 5] >> (eb & 31)) & 1u) != 0;
        // doubler
        const Fr x1_2 = fr_mul_chain(D.u, D.u);
        const Fr num = fr_add(fr_add(fr_add(fr_add(x1_2, x1_2), x1_2), fr_mul_small(D.u, (uint32_t)(2 * HZ_MT_A))), fr_one());
        const Fr lam = fr_mul_chain(num, hp[2 * i]);
        if (err) { const Fr lhs = fr_mul_chain(lam, fr_add(D.v, D.v)); if (!fr_eq(lhs, num)) hz_fill_err(err, at + 4, __LINE__, lhs, num); }
        AffPt dd;
        dd.u = fr_sub(fr_sub(fr_mul_chain(lam, lam), fr_from_u64(HZ_MT_A)), fr_add(D.u, D.u));
        dd.v = fr_sub(fr_mul_chain(lam, fr_sub(D.u, dd.u)), D.v);
        hz_store_fr(&w[at], dd.u); hz_store_fr(&w[at + 1], dd.v); hz_store_fr(&w[at + 2], lam); hz_store_fr(&w[at + 3], x1_2);
        // adder (in1 = the doubled point, in2 = the accumulator), selector
        const AffPt ad = hz_fill_madd(w, at + 4, err, dd, acc, hp[2 * i + 1]);
        if (sel) acc = ad;
        hz_store_fr(&w[at + 7], acc.u); hz_store_fr(&w[at + 8], acc.v);
        D = dd;
    }
}
__device__ __forceinline__ void hz_fill_fix(Fr* w, HzErrBuf* err, const Fr* hint, const Fr* aff, const Fr* tab, const HzFillRec& rec, int k) {
    const int i0 = k * HZ_FILL_CH, i1 = i0 + HZ_FILL_CH < 83 ? i0 + HZ_FILL_CH : 83;
    AffPt acc;
    acc.u = aff[2 * (66 + k)]; acc.v = aff[2 * (66 + k) + 1];
    uint32_t at = rec.base_fix + 3u * (uint32_t)i0;
    for (int i = i0; i < i1; i++, at += 3) {
        int sel = 0;
        for (int b = 0; b < 3; b++) sel |= (int)((rec.s[(3 * i + b) >> 5] >> ((3 * i + b) & 31)) & 1u) << b;
        AffPt win;
        win.u = tab[i * 16 + 2 * sel]; win.v = tab[i * 16 + 2 * sel + 1];
        acc = hz_fill_madd(w, at, err, acc, win, hint[505 + i]);
    }
}
// the multiplexer of window i of EscalarMulFix segment 0 (MultiMux3 with two outputs over the window's eight
// constant points, selected by three bits of S): out[2], a210[2], a21[2], a20[2], a2[2], a10[2], a1[2], a0[2], s10
__device__ __forceinline__ void hz_fill_winmux(Fr* w, const Fr* tab, const HzFillRec& rec, uint32_t stride, int i) {
    uint32_t b[3];
    for (int k = 0; k < 3; k++) b[k] = (rec.s[(3 * i + k) >> 5] >> ((3 * i + k) & 31)) & 1u;
    const bool s0 = b[0] != 0, s1 = b[1] != 0, s2 = b[2] != 0, s10 = s0 && s1;
    const Fr z = fr_zero();
    const uint32_t at = rec.base_win + stride * (uint32_t)i;
    for (int o = 0; o < 2; o++) {
        Fr q[8];
        for (int j = 0; j < 8; j++) q[j] = tab[i * 16 + 2 * j + o];
        const Fr d210 = fr_sub(fr_add(fr_add(fr_sub(fr_add(fr_sub(fr_sub(q[7], q[6]), q[5]), q[4]), q[3]), q[2]), q[1]), q[0]);
        const Fr v210 = s10 ? d210 : z;
        const Fr v21 = s1 ? fr_add(fr_sub(fr_sub(q[6], q[4]), q[2]), q[0]) : z;
        const Fr v20 = s0 ? fr_add(fr_sub(fr_sub(q[5], q[4]), q[1]), q[0]) : z;
        const Fr v2 = fr_sub(q[4], q[0]);
        const Fr v10 = s10 ? fr_add(fr_sub(fr_sub(q[3], q[2]), q[1]), q[0]) : z;
        const Fr v1 = s1 ? fr_sub(q[2], q[0]) : z;
        const Fr v0 = s0 ? fr_sub(q[1], q[0]) : z;
        const Fr hi = fr_add(fr_add(fr_add(v210, v21), v20), v2);
        const Fr out = fr_add(s2 ? hi : z, fr_add(fr_add(fr_add(v10, v1), v0), q[0]));
        hz_store_fr(&w[at + o], out); hz_store_fr(&w[at + 2 + o], v210); hz_store_fr(&w[at + 4 + o], v21);
        hz_store_fr(&w[at + 6 + o], v20); hz_store_fr(&w[at + 8 + o], v2); hz_store_fr(&w[at + 10 + o], v10);
        hz_store_fr(&w[at + 12 + o], v1); hz_store_fr(&w[at + 14 + o], v0);
    }
    hz_store_fr(&w[at + 16], fr_from_u64(s10 ? 1u : 0u));
}
#endif
