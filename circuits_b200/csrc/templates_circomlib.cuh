// circomlib 0.5.2 templates, generic over the evaluation mode (ctx.cuh).
//
// circomlib is an npm dependency of the reference (package.json:27, lock 0.5.2) that is not
// vendored under /root/reference; each function restates one upstream template
// (`circuits/<file>.circom`, semantics as in SURVEY.md Appendix A) with the reference call
// sites cited.  Signal declaration order, aliasing of `<==` wires and the `===` checks are
// the ones modelled by oracle/circomlib.py -- the parity tests compare the two signal by
// signal.  Where upstream instantiates a sub-component before the ones it is wired after,
// the range is reserved (`sub_reserve`) with its closed-form size.
#pragma once
#include "ctx.cuh"
#include "hints.cuh"
#include "poseidon_constants.inc"

#define HZ_MV(x) static_cast<decltype(x)&&>(x)
#define HZ_BASE8_X_LIMBS {0xbb957051u, 0x2893f3f6u, 0x0534e0b6u, 0x2ab8d801u, 0x9d6277c1u, 0x4eacb2e0u, 0xd63e739bu, 0x0bb77a6au}
#define HZ_BASE8_Y_LIMBS {0x872d7d8bu, 0x4b3c257au, 0xb9e13377u, 0xfce0051fu, 0xd16bf9edu, 0x25572e1cu, 0xf7a0b249u, 0x25797203u}
#define HZ_TONELLI_C_LIMBS {0x725b19f0u, 0x9bd61b6eu, 0x41112ed4u, 0x402d111eu, 0x8ef62abcu, 0x00e0a7ebu, 0xa58a7e85u, 0x2a3c09f0u}
#define HZ_TONELLI_T_LIMBS {0x3e1f593fu, 0x9b970914u, 0x833e8487u, 0x181585d2u, 0x85045b68u, 0x131a029bu, 0x0644e72eu, 0x00000003u}
#define HZ_TONELLI_TM1H_LIMBS {0x1f0fac9fu, 0xcdcb848au, 0x419f4243u, 0x0c0ac2e9u, 0xc2822db4u, 0x098d014du, 0x83227397u, 0x00000001u}
#define HZ_TONELLI_T1H_LIMBS {0x1f0faca0u, 0xcdcb848au, 0x419f4243u, 0x0c0ac2e9u, 0xc2822db4u, 0x098d014du, 0x83227397u, 0x00000001u}
#define TM template <class M>
typedef BVT<MCompute, 8> BVC;
template <class M> using BV = BVT<M, 8>;

// ---------------------------------------------------------------------------
// bitify.circom (src/decode-tx.circom:96 and throughout)
// ---------------------------------------------------------------------------
TM HZ_FN BV<M> Num2Bits(Comp<M>&& c, int n, const V<M>& x) {
    V<M> in_ = c.def("in", x);
    BV<M> out; out.clear(n);
    for (int i = 0; i < n; i++) out.put(i, c.def(Nm("out", i), Ops<M>::K(vbit(in_, i))));
    if (c.sanity() && n < 254) {
        // sum(out[i] 2^i) === in  <=>  no bit of `in` at or above n
        bool ok = true;
        for (int i = n; i < 254; i++) ok = ok && !vbit(in_, i);
        if (!ok) {
            V<M> lc = Ops<M>::K(0);
            for (int i = 0; i < n; i++) if (out.bit(i)) lc = lc + Ops<M>::pow2(i);
            c.check(lc, in_, __LINE__);
        }
    }
    return out;
}

// value of sum(bits[i] 2^i) mod p
TM HZ_FN V<M> bits_lc(const BV<M>& b, int from, int n);
template <> HZ_HD VC bits_lc<MCompute>(const BVC& b, int from, int n) {
    Fr r = fr_zero();
    for (int i = 0; i < n && i < 254; i++) if (b.bit(from + i)) r.l[i >> 5] |= 1u << (i & 31);
    if (fr_geq_p(r)) fr_sub_p(r);
    VC v = vc(r);
    for (int i = 254; i < n; i++) if (b.bit(from + i)) {
        // 2^254, 2^255 mod p
        Fr t = fr_pow2(253); t = fr_add(t, t); if (i == 255) t = fr_add(t, t);
        v = v + vc(t);
    }
    return v;
}
#ifndef HZ_DEVICE_ONLY
template <> inline VL bits_lc<MLayout>(const BVT<MLayout, 8>&, int, int) { return vl(); }
#endif

TM HZ_FN V<M> Bits2Num(Comp<M>&& c, const BV<M>& bits, int from, int n) {
    BV<M> ins; ins.clear(n);
    for (int i = 0; i < n; i++) ins.put(i, c.def(Nm("in", i), bits.get(from + i)));
    return c.def("out", bits_lc<M>(ins, 0, n));
}

// ---------------------------------------------------------------------------
// compconstant.circom / aliascheck.circom
// ---------------------------------------------------------------------------
#define HZ_SZ_COMPCONSTANT 263   /* 127 parts + sout + Num2Bits(135) outs, inputs wired */
TM HZ_FN V<M> CompConstant(Comp<M>&& c, const Fr& ct, const BV<M>& bits) {
    BV<M> ins; ins.clear(254);
    for (int i = 0; i < 254; i++) ins.put(i, c.def(Nm("in", i), bits.get(i)));
    uint32_t parts = c.sigs("parts", 127);
    uint32_t sout = c.sig("sout");
    V<M> total = Ops<M>::K(0);
    for (int i = 0; i < 127; i++) {
        uint32_t clsb = fr_bit(ct, 2 * i), cmsb = fr_bit(ct, 2 * i + 1);
        uint32_t sl = ins.bit(2 * i), sm = ins.bit(2 * i + 1);
        // a = 2^i, b = 2^128 - 2^i ; part value is 0, a or b
        int sel;  // 0: zero, 1: a, 2: b
        if (!cmsb && !clsb) sel = (sm | sl) ? 2 : 0;
        else if (!cmsb && clsb) sel = sm ? 2 : (sl ? 0 : 1);
        else if (cmsb && !clsb) sel = sm ? (sl ? 2 : 0) : 1;
        else sel = (sm & sl) ? 0 : 1;
        V<M> pv = sel == 0 ? Ops<M>::K(0) : (sel == 1 ? Ops<M>::pow2(i) : Ops<M>::pow2(128) - Ops<M>::pow2(i));
        V<M> p = c.set(parts + i, pv);
        total = total + p;
    }
    V<M> s = c.set(sout, total);
    BV<M> nb = Num2Bits(c.sub("num2bits"), 135, s);
    return c.alias("out", nb.get(127));
}

HZ_HD Fr fr_const_pm1() { Fr r = fr_P(); r.l[0] -= 1; return r; }
HZ_HD Fr fr_const_half() {
    Fr r; const uint32_t h[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    for (int i = 0; i < 8; i++) r.l[i] = h[i];
    return r;
}
// BabyJubJub sub-group order l - 1
HZ_HD Fr fr_const_suborder_m1() {
    // l = 2736030358979909402780800718157159386076813972158567259200215660948447373041
    Fr r; const uint32_t h[8] = {0x392126f0u, 0x677297dcu, 0x3920ee0au, 0xab3eedb8u, 0xd0302b0bu, 0x370a08b6u, 0x5c263405u, 0x060c89ceu};
    for (int i = 0; i < 8; i++) r.l[i] = h[i];
    return r;
}

TM HZ_FN void AliasCheck(Comp<M>&& c, const BV<M>& bits) {
    BV<M> ins; ins.clear(254);
    for (int i = 0; i < 254; i++) ins.put(i, c.def(Nm("in", i), bits.get(i)));
    V<M> o = CompConstant(c.sub("compConstant"), fr_const_pm1(), ins);
    HZ_CHECK0(c, o);
}

TM HZ_FN BV<M> Num2Bits_strict(Comp<M>&& c, const V<M>& x) {
    V<M> in_ = c.def("in", x);
    Comp<M> ac = c.sub_reserve("aliasCheck", HZ_SZ_COMPCONSTANT);
    BV<M> bits = Num2Bits(c.sub("n2b"), 254, in_);
    for (int i = 0; i < 254; i++) c.alias(Nm("out", i), bits.get(i));
    AliasCheck(HZ_MV(ac), bits);
    return bits;
}

// ---------------------------------------------------------------------------
// comparators.circom
// ---------------------------------------------------------------------------
TM HZ_FN V<M> IsZero(Comp<M>&& c, const V<M>& x) {
    V<M> in_ = c.def("in", x);
    uint32_t out = c.sig("out");
    // out = -in * inv + 1 is (in == 0) for the inv the template prescribes; the inverse itself may be deferred
    V<M> o = c.set(out, c.iszero_inv("inv", in_));
    HZ_CHECK0(c, in_ * o);
    return o;
}
// IsZero with the inverse supplied (batched inversions)
TM HZ_FN V<M> IsZeroWithInv(Comp<M>&& c, const V<M>& x, const V<M>& xinv) {
    V<M> in_ = c.def("in", x);
    uint32_t out = c.sig("out");
    V<M> inv_ = c.def("inv", xinv);
    V<M> o = c.set(out, 1 - in_ * inv_);
    HZ_CHECK0(c, in_ * o);
    return o;
}

TM HZ_FN V<M> IsEqual(Comp<M>&& c, const V<M>& a, const V<M>& b) {
    V<M> i0 = c.def(Nm("in", 0), a);
    V<M> i1 = c.def(Nm("in", 1), b);
    V<M> o = IsZero(c.sub("isz"), i1 - i0);
    return c.alias("out", o);
}

TM HZ_FN void ForceEqualIfEnabled(Comp<M>&& c, const V<M>& enabled, const V<M>& a, const V<M>& b) {
    V<M> en = c.def("enabled", enabled);
    V<M> i0 = c.def(Nm("in", 0), a);
    V<M> i1 = c.def(Nm("in", 1), b);
    V<M> o = IsZero(c.sub("isz"), i1 - i0);
    HZ_CHECK0(c, (1 - o) * en);
}

TM HZ_FN V<M> LessThan(Comp<M>&& c, int n, const V<M>& a, const V<M>& b) {
    V<M> i0 = c.def(Nm("in", 0), a);
    V<M> i1 = c.def(Nm("in", 1), b);
    uint32_t out = c.sig("out");
    BV<M> bits = Num2Bits(c.sub("n2b"), n + 1, i0 + Ops<M>::pow2(n) - i1);
    return c.set(out, 1 - bits.get(n));
}

TM HZ_FN V<M> GreaterEqThan(Comp<M>&& c, int n, const V<M>& a, const V<M>& b) {
    V<M> i0 = c.def(Nm("in", 0), a);
    V<M> i1 = c.def(Nm("in", 1), b);
    V<M> o = LessThan(c.sub("lt"), n, i1, i0 + 1);
    return c.alias("out", o);
}

// ---------------------------------------------------------------------------
// gates.circom
// ---------------------------------------------------------------------------
TM HZ_FN V<M> XOR(Comp<M>&& c, const V<M>& a, const V<M>& b) {
    V<M> a_ = c.def("a", a);
    V<M> b_ = c.def("b", b);
    return c.def("out", a_ + b_ - 2 * (a_ * b_));
}
TM HZ_FN V<M> AND(Comp<M>&& c, const V<M>& a, const V<M>& b) {
    V<M> a_ = c.def("a", a);
    V<M> b_ = c.def("b", b);
    return c.def("out", a_ * b_);
}
// MultiAND(n) for n in {3, 4} (smtprocessor / smtverifier keysOk)
TM HZ_FN V<M> MultiAND(Comp<M>&& c, const V<M>* xs, int n) {
    V<M> ins[4];
    for (int i = 0; i < n; i++) ins[i] = c.def(Nm("in", i), xs[i]);
    if (n == 1) return c.alias("out", ins[0]);
    if (n == 2) return c.alias("out", AND(c.sub("and1"), ins[0], ins[1]));
    Comp<M> and2 = c.sub_reserve("and2", 1);
    int n1 = n / 2;
    V<M> a0 = MultiAND(c.sub(Nm("ands", 0)), ins, n1);
    V<M> a1 = MultiAND(c.sub(Nm("ands", 1)), ins + n1, n - n1);
    return c.alias("out", AND(HZ_MV(and2), a0, a1));
}

// ---------------------------------------------------------------------------
// mux1..mux4.circom
// ---------------------------------------------------------------------------
TM HZ_FN V<M> Mux1(Comp<M>&& c, const V<M>& c0, const V<M>& c1, const V<M>& s) {
    V<M> k0 = c.def(Nm("c", 0), c0);
    V<M> k1 = c.def(Nm("c", 1), c1);
    V<M> s_ = c.def("s", s);
    Comp<M> m = c.sub("mux");
    m.alias(Nm("c", 0, 0), k0); m.alias(Nm("c", 0, 1), k1); m.alias("s", s_);
    V<M> o = m.def(Nm("out", 0), (k1 - k0) * s_ + k0);
    return c.alias("out", o);
}

TM HZ_FN V<M> Mux2(Comp<M>&& c, const V<M>* cs, const V<M>* s) {
    V<M> k[4], ss[2];
    for (int i = 0; i < 4; i++) k[i] = c.def(Nm("c", i), cs[i]);
    for (int i = 0; i < 2; i++) ss[i] = c.def(Nm("s", i), s[i]);
    Comp<M> m = c.sub("mux");
    for (int i = 0; i < 4; i++) m.alias(Nm("c", 0, i), k[i]);
    for (int i = 0; i < 2; i++) m.alias(Nm("s", i), ss[i]);
    uint32_t out = m.sig(Nm("out", 0));
    uint32_t a10 = m.sig(Nm("a10", 0)), a1 = m.sig(Nm("a1", 0)), a0 = m.sig(Nm("a0", 0));
    m.alias(Nm("a", 0), k[0]);
    V<M> s10 = m.def("s10", ss[1] * ss[0]);
    V<M> v10 = m.set(a10, (k[3] - k[2] - k[1] + k[0]) * s10);
    V<M> v1 = m.set(a1, (k[2] - k[0]) * ss[1]);
    V<M> v0 = m.set(a0, (k[1] - k[0]) * ss[0]);
    V<M> o = m.set(out, v10 + v1 + v0 + k[0]);
    return c.alias("out", o);
}

// MultiMux3 with n outputs (n = 1 via Mux3, n = 2 in WindowMulFix); cs[i][8]
#define HZ_SZ_MULTIMUX3(n) (8 * (n) + 1)
TM HZ_FN void MultiMux3(Comp<M>&& m, int n, const V<M> (*k)[8], const V<M>* ss, V<M>* outs) {
    for (int i = 0; i < n; i++) for (int j = 0; j < 8; j++) m.alias(Nm("c", i, j), k[i][j]);
    for (int i = 0; i < 3; i++) m.alias(Nm("s", i), ss[i]);
    uint32_t out = m.sigs("out", n);
    uint32_t a210 = m.sigs("a210", n), a21 = m.sigs("a21", n), a20 = m.sigs("a20", n), a2 = m.sigs("a2", n);
    uint32_t a10 = m.sigs("a10", n), a1 = m.sigs("a1", n), a0 = m.sigs("a0", n);
    for (int i = 0; i < n; i++) m.alias(Nm("a", i), k[i][0]);
    V<M> s10 = m.def("s10", ss[1] * ss[0]);
    for (int i = 0; i < n; i++) {
        const V<M>* q = k[i];
        V<M> v210 = m.set(a210 + i, (q[7] - q[6] - q[5] + q[4] - q[3] + q[2] + q[1] - q[0]) * s10);
        V<M> v21 = m.set(a21 + i, (q[6] - q[4] - q[2] + q[0]) * ss[1]);
        V<M> v20 = m.set(a20 + i, (q[5] - q[4] - q[1] + q[0]) * ss[0]);
        V<M> v2 = m.set(a2 + i, q[4] - q[0]);
        V<M> v10 = m.set(a10 + i, (q[3] - q[2] - q[1] + q[0]) * s10);
        V<M> v1 = m.set(a1 + i, (q[2] - q[0]) * ss[1]);
        V<M> v0 = m.set(a0 + i, (q[1] - q[0]) * ss[0]);
        outs[i] = m.set(out + i, (v210 + v21 + v20 + v2) * ss[2] + (v10 + v1 + v0 + q[0]));
    }
}
TM HZ_FN V<M> Mux3(Comp<M>&& c, const V<M>* cs, const V<M>* s) {
    V<M> k[1][8], ss[3], o[1];
    for (int i = 0; i < 8; i++) k[0][i] = c.def(Nm("c", i), cs[i]);
    for (int i = 0; i < 3; i++) ss[i] = c.def(Nm("s", i), s[i]);
    MultiMux3<M>(c.sub("mux"), 1, k, ss, o);
    return c.alias("out", o[0]);
}

#define HZ_SZ_MUX4_WIRED 20
TM HZ_FN V<M> Mux4(Comp<M>&& c, const V<M>* cs, const V<M>* s) {
    V<M> k[16], ss[4];
    for (int i = 0; i < 16; i++) k[i] = c.def(Nm("c", i), cs[i]);
    for (int i = 0; i < 4; i++) ss[i] = c.def(Nm("s", i), s[i]);
    Comp<M> m = c.sub("mux");
    for (int i = 0; i < 16; i++) m.alias(Nm("c", 0, i), k[i]);
    for (int i = 0; i < 4; i++) m.alias(Nm("s", i), ss[i]);
    uint32_t out = m.sig(Nm("out", 0));
    uint32_t a3210 = m.sig(Nm("a3210", 0)), a321 = m.sig(Nm("a321", 0)), a320 = m.sig(Nm("a320", 0)), a310 = m.sig(Nm("a310", 0));
    uint32_t a32 = m.sig(Nm("a32", 0)), a31 = m.sig(Nm("a31", 0)), a30 = m.sig(Nm("a30", 0)), a3 = m.sig(Nm("a3", 0));
    uint32_t a210 = m.sig(Nm("a210", 0)), a21 = m.sig(Nm("a21", 0)), a20 = m.sig(Nm("a20", 0)), a10 = m.sig(Nm("a10", 0));
    uint32_t a2 = m.sig(Nm("a2", 0)), a1 = m.sig(Nm("a1", 0)), a0 = m.sig(Nm("a0", 0));
    m.alias(Nm("a", 0), k[0]);
    V<M> s10 = m.def("s10", ss[1] * ss[0]);
    V<M> s20 = m.def("s20", ss[2] * ss[0]);
    V<M> s21 = m.def("s21", ss[2] * ss[1]);
    V<M> s210 = m.def("s210", s21 * ss[0]);
    V<M> v3210 = m.set(a3210, (k[15] - k[14] - k[13] + k[12] - k[11] + k[10] + k[9] - k[8]
                               - k[7] + k[6] + k[5] - k[4] + k[3] - k[2] - k[1] + k[0]) * s210);
    V<M> v321 = m.set(a321, (k[14] - k[12] - k[10] + k[8] - k[6] + k[4] + k[2] - k[0]) * s21);
    V<M> v320 = m.set(a320, (k[13] - k[12] - k[9] + k[8] - k[5] + k[4] + k[1] - k[0]) * s20);
    V<M> v310 = m.set(a310, (k[11] - k[10] - k[9] + k[8] - k[3] + k[2] + k[1] - k[0]) * s10);
    V<M> v32 = m.set(a32, (k[12] - k[8] - k[4] + k[0]) * ss[2]);
    V<M> v31 = m.set(a31, (k[10] - k[8] - k[2] + k[0]) * ss[1]);
    V<M> v30 = m.set(a30, (k[9] - k[8] - k[1] + k[0]) * ss[0]);
    V<M> v3 = m.set(a3, k[8] - k[0]);
    V<M> v210 = m.set(a210, (k[7] - k[6] - k[5] + k[4] - k[3] + k[2] + k[1] - k[0]) * s210);
    V<M> v21 = m.set(a21, (k[6] - k[4] - k[2] + k[0]) * s21);
    V<M> v20 = m.set(a20, (k[5] - k[4] - k[1] + k[0]) * s20);
    V<M> v10 = m.set(a10, (k[3] - k[2] - k[1] + k[0]) * s10);
    V<M> v2 = m.set(a2, (k[4] - k[0]) * ss[2]);
    V<M> v1 = m.set(a1, (k[2] - k[0]) * ss[1]);
    V<M> v0 = m.set(a0, (k[1] - k[0]) * ss[0]);
    V<M> hi = v3210 + v321 + v320 + v310 + v32 + v31 + v30 + v3;
    V<M> lo = v210 + v21 + v20 + v10 + v2 + v1 + v0 + k[0];
    V<M> o = m.set(out, hi * ss[3] + lo);
    return c.alias("out", o);
}

// ---------------------------------------------------------------------------
// switcher.circom
// ---------------------------------------------------------------------------
TM HZ_FN void Switcher(Comp<M>&& c, const V<M>& sel, const V<M>& L, const V<M>& R, V<M>& outL, V<M>& outR) {
    V<M> sel_ = c.def("sel", sel);
    V<M> L_ = c.def("L", L);
    V<M> R_ = c.def("R", R);
    uint32_t oL = c.sig("outL"), oR = c.sig("outR");
    V<M> aux = c.def("aux", (R_ - L_) * sel_);
    outL = c.set(oL, aux + L_);
    outR = c.set(oR, R_ - aux);
}

// ---------------------------------------------------------------------------
// poseidon.circom (src/decode-tx.circom:275, src/lib/hash-state.circom:32, SMT hashes)
// x^5, t = nInputs + 1, R_F = 8, R_P = [56,57,56,60,60,63,64,63][t-2]  
// ---------------------------------------------------------------------------
HZ_HD int poseidon_rp(int t) {
    switch (t) { case 2: return 56; case 3: return 57; case 4: return 56; case 5: return 60; case 6: return 60; default: return 63; }
}
// witness entries of one Poseidon(t-1) whose inputs are all wired signals
HZ_HD uint32_t poseidon_size(int t) {
    int rp = poseidon_rp(t), R = 8 + rp;
    return 1 + 2 * t * R + 3 * (8 * t + rp);
}

TM struct PoseidonMath;
template <> struct PoseidonMath<MCompute> {
    static HZ_HD VC ark(const VC& x, int t, int idx) { return vc(fr_add(x.v, hz_poseidon_c(t, idx))); }
    static HZ_HD VC mix(const VC* s, int t, int r) {
        Fr sv[7];
        for (int j = 0; j < t; j++) sv[j] = s[j].v;
        return vc(fr_dot_mont(sv, hz_poseidon_mr(t, r), t));
    }
};
template <> struct PoseidonMath<MLayout> {
    static HZ_HD VL ark(const VL&, int, int) { return vl(); }
    static HZ_HD VL mix(const VL*, int, int) { return vl(); }
};

TM HZ_FN V<M> Poseidon(Comp<M>&& c, const V<M>* inputs, int n);
TM HZ_FN V<M> PoseidonGeneric(Comp<M>&& c, const V<M>* inputs, int n) {
    const int t = n + 1, rp = poseidon_rp(t), R = 8 + rp;
    V<M> st[7];
    st[0] = Ops<M>::K(0);
    for (int i = 0; i < n; i++) st[i + 1] = c.def(Nm("inputs", i), inputs[i]);
    for (int i = 0; i < R; i++) {
        V<M> a[7];
        {
            Comp<M> ark = c.sub(Nm("ark", i));
            for (int j = 0; j < t; j++) st[j] = ark.def(Nm("in", j), st[j]);
            for (int j = 0; j < t; j++) a[j] = ark.def(Nm("out", j), PoseidonMath<M>::ark(st[j], t, t * i + j));
        }
        Comp<M> mix = c.sub_reserve(Nm("mix", i), t);
        bool full = i < 4 || i >= rp + 4;
        int k = i < 4 ? i : (full ? i - rp : i - 4);
        int ns = full ? t : 1;
        for (int j = 0; j < ns; j++) {
            Nm snm = full ? Nm("sigmaF", k, j) : Nm("sigmaP", k);
            Comp<M> sg = c.sub(snm);
            sg.alias("in", a[j]);
            uint32_t out = sg.sig("out");
            V<M> in2 = sg.def("in2", a[j] * a[j]);
            V<M> in4 = sg.def("in4", in2 * in2);
            a[j] = sg.set(out, in4 * a[j]);
        }
        for (int j = 0; j < t; j++) mix.alias(Nm("in", j), a[j]);
        for (int r = 0; r < t; r++) st[r] = mix.def(Nm("out", r), PoseidonMath<M>::mix(a, t, r));
    }
    return c.alias("out", st[0]);
}

#ifdef HZ_DEVICE_ONLY
// ---- the permutation on a lane group ---------------------------------------------------------
// Lane `base + j` of the 8-lane group holds state element j (j < T); a group runs one permutation
// (base = 0) or two side by side (bases 0 and 4, T <= 4: the old-path / new-path hashes of an SMT
// level).  Per round each lane adds its round constant, the S-box lanes do x -> x^5 (all T lanes
// in a full round, lane `base` alone in a partial one), then every lane gathers the T S-box
// outputs of its permutation with shuffles and reduces its own MDS row (constants in Montgomery
// form: one reduction per row).  Every signal is stored by the lane that computed it, at the
// witness index the layout walk gives it: per round `ark.out[T]`, `mix.out[T]`, then
// (out, in2, in4) per S-box.
__device__ __forceinline__ Fr hz_shfl_fr(const Fr& v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src, HZ_G);
    return r;
}
// DRY: the digest only, no entry is stored (k_edjob: the message hash of a signature before its slot is evaluated)
template <int T, bool DRY = false> __device__ __noinline__ Fr hz_poseidon_lanes(Fr* w, uint32_t cur, Fr s, int base) {
    const int rp = poseidon_rp(T), R = 8 + rp;
    const int lane = (int)hz_glane();
    const int j = lane - base;
    const bool mine = j >= 0 && j < T;
    const int jj = mine ? j : 0;
    const Fr* C = &hz_poseidon_c(T, 0);
    const Fr* Mrow = hz_poseidon_mr(T, jj);
    Fr mreg[T <= 4 ? T : 1];
    if (T <= 4) {
#pragma unroll
        for (int k = 0; k < (T <= 4 ? T : 1); k++) mreg[k] = Mrow[k];
    }
    for (int i = 0; i < R; i++) {
        s = fr_add(s, C[T * i + jj]);
        if (!DRY && mine) hz_store_fr(&w[cur + jj], s);
        const bool full = i < 4 || i >= rp + 4;
        const uint32_t sg = cur + 2 * T;
        if (mine && (full || j == 0)) {
            Fr in2 = fr_mul_chain(s, s);
            Fr in4 = fr_mul_chain(in2, in2);
            s = fr_mul_chain(in4, s);
            if (!DRY) {
                hz_store_fr(&w[sg + 3 * jj], s);
                hz_store_fr(&w[sg + 3 * jj + 1], in2);
                hz_store_fr(&w[sg + 3 * jj + 2], in4);
            }
        }
        Fr a[T];
#pragma unroll
        for (int k = 0; k < T; k++) a[k] = hz_shfl_fr(s, base + k);
        s = fr_dot_mont_chain<T>(a, T <= 4 ? mreg : Mrow);
        if (!DRY && mine) hz_store_fr(&w[cur + T + jj], s);
        cur += 2 * T + (full ? 3 * T : 3);
    }
    return hz_shfl_fr(s, base);
}
__device__ __forceinline__ Fr hz_poseidon_dispatch_dry(int t, const Fr& s) {
    switch (t) {
        case 3: return hz_poseidon_lanes<3, true>(nullptr, 0, s, 0);
        case 4: return hz_poseidon_lanes<4, true>(nullptr, 0, s, 0);
        case 5: return hz_poseidon_lanes<5, true>(nullptr, 0, s, 0);
        case 6: return hz_poseidon_lanes<6, true>(nullptr, 0, s, 0);
        default: return hz_poseidon_lanes<7, true>(nullptr, 0, s, 0);
    }
}
__device__ __forceinline__ Fr hz_poseidon_dispatch(Fr* w, int t, uint32_t cur, const Fr& s, int base) {
    switch (t) {
        case 3: return hz_poseidon_lanes<3>(w, cur, s, base);
        case 4: return hz_poseidon_lanes<4>(w, cur, s, base);
        case 5: return hz_poseidon_lanes<5>(w, cur, s, base);
        case 6: return hz_poseidon_lanes<6>(w, cur, s, base);
        default: return hz_poseidon_lanes<7>(w, cur, s, base);
    }
}
// declares the `inputs` / `ark[0].in[0]` entries exactly as PoseidonGeneric does and returns the
// witness index where round 0's ark.out starts
__device__ __forceinline__ uint32_t hz_poseidon_head(Comp<MCompute>& c, const VC* inputs, int n) {
    for (int i = 0; i < n; i++) c.def(Nm("inputs", i), inputs[i]);
    c.def(Nm("in", 0), vc(fr_zero()));
    return c.cur;
}
template <> __device__ __noinline__ VC Poseidon<MCompute>(Comp<MCompute>&& c, const VC* inputs, int n) {
    const int t = n + 1;
    const uint32_t c0 = c.cur;
    uint32_t n_own = 0;
    for (int i = 0; i < n; i++) n_own += inputs[i].sig ? 0u : 1u;
    if (c.skip_hash) {
        // the permutation is evaluated by a transposed kernel (k_hstate: the state hashes, from the inputs recorded
        // here; k_sighash: the message hash); nothing on this side reads the digest
        if (c.hs_rec && n == 4) {
            const int k = (*c.hs_n)++;
            if (c.solo || hz_glane() == 0) for (int i = 0; i < 4; i++) c.hs_rec[4 * k + i] = inputs[i].v;
        }
        c.cur = c0 + n_own + poseidon_size(t);
        VC z; z.v = fr_zero(); z.sig = true; z.ph = true;
        return z;
    }
    uint32_t cur = hz_poseidon_head(c, inputs, n);
    const int lane = (int)hz_glane();
    Fr s = (lane >= 1 && lane <= n) ? inputs[lane - 1].v : fr_zero();
    Fr o = c.dry ? hz_poseidon_dispatch_dry(t, s) : hz_poseidon_dispatch(c.w, t, cur, s, 0);
    c.cur = c0 + n_own + poseidon_size(t);
    VC r; r.v = o; r.sig = true;
    return r;
}
#else
template <> inline VC Poseidon<MCompute>(Comp<MCompute>&& c, const VC* inputs, int n) {
    return PoseidonGeneric<MCompute>(HZ_MV(c), inputs, n);
}
template <> inline VL Poseidon<MLayout>(Comp<MLayout>&& c, const VL* inputs, int n) {
    if (hz_poseidon_starts && n == 4) hz_poseidon_starts->t5.push_back(c.cur);
    if (hz_poseidon_starts && n == 6) hz_poseidon_starts->t7.push_back(c.cur);
    return PoseidonGeneric<MLayout>(HZ_MV(c), inputs, n);
}
#endif

// two permutations of the same width in two components that do not depend on each other
TM HZ_FN void PoseidonDual(Comp<M>&& cA, const V<M>* inA, Comp<M>&& cB, const V<M>* inB, int n, V<M>& oA, V<M>& oB) {
    oA = Poseidon<M>(HZ_MV(cA), inA, n);
    oB = Poseidon<M>(HZ_MV(cB), inB, n);
}
#ifdef HZ_DEVICE_ONLY
template <> __device__ __noinline__ void PoseidonDual<MCompute>(Comp<MCompute>&& cA, const VC* inA, Comp<MCompute>&& cB,
                                                               const VC* inB, int n, VC& oA, VC& oB) {
    const int t = n + 1;            // t <= 4: both fit the 8-lane group
    uint32_t a0 = cA.cur, b0 = cB.cur, ownA = 0, ownB = 0;
    for (int i = 0; i < n; i++) { ownA += inA[i].sig ? 0u : 1u; ownB += inB[i].sig ? 0u : 1u; }
    if (cA.skip_hash) {
        // both permutations are evaluated by a transposed kernel (k_leafhash), entries and digests
        cA.cur = a0 + ownA + poseidon_size(t);
        cB.cur = b0 + ownB + poseidon_size(t);
        oA.v = fr_zero(); oA.sig = true; oA.ph = true;
        oB = oA;
        return;
    }
    uint32_t curA = hz_poseidon_head(cA, inA, n);
    uint32_t curB = hz_poseidon_head(cB, inB, n);
    const int lane = (int)hz_glane();
    const bool second = lane >= 4;
    const int j = lane & 3;
    const VC* in = second ? inB : inA;
    Fr s = (j >= 1 && j <= n) ? in[j - 1].v : fr_zero();
    Fr o = hz_poseidon_dispatch(cA.w, t, second ? curB : curA, s, second ? 4 : 0);
    // every lane needs both digests: lanes 0..3 hold A's, lanes 4..7 hold B's
    oA.v = hz_shfl_fr(o, 0); oA.sig = true;
    oB.v = hz_shfl_fr(o, 4); oB.sig = true;
    cA.cur = a0 + ownA + poseidon_size(t);
    cB.cur = b0 + ownB + poseidon_size(t);
}
#endif

// ---------------------------------------------------------------------------
// smt/*.circom  (src/rollup-tx.circom:537,559; src/fee-tx.circom:97; src/withdraw.circom:47)
// ---------------------------------------------------------------------------
TM HZ_FN V<M> SMTHash1(Comp<M>&& c, const V<M>& key, const V<M>& value) {
    V<M> in[3];
    in[0] = c.def("key", key);
    in[1] = c.def("value", value);
    in[2] = Ops<M>::K(1);
    return c.alias("out", Poseidon<M>(c.sub("h"), in, 3));
}
TM HZ_FN V<M> SMTHash2(Comp<M>&& c, const V<M>& L, const V<M>& R) {
    V<M> in[2];
    in[0] = c.def("L", L);
    in[1] = c.def("R", R);
    return c.alias("out", Poseidon<M>(c.sub("h"), in, 2));
}
// the two leaf hashes / the two proof hashes of a level: independent, evaluated side by side
// (both components must own their witness range: sub_reserve)
#define HZ_SZ_SMTHASH1_WIRED (poseidon_size(4) + 1)      /* the constant input 1 owns an entry */
TM HZ_FN void SMTHash1Pair(Comp<M>&& cA, const V<M>& keyA, const V<M>& valA, Comp<M>&& cB, const V<M>& keyB,
                           const V<M>& valB, V<M>& oA, V<M>& oB) {
    V<M> a[3], b[3];
    a[0] = cA.def("key", keyA); a[1] = cA.def("value", valA); a[2] = Ops<M>::K(1);
    b[0] = cB.def("key", keyB); b[1] = cB.def("value", valB); b[2] = Ops<M>::K(1);
    PoseidonDual<M>(cA.sub("h"), a, cB.sub("h"), b, 3, oA, oB);
    cA.alias("out", oA); cB.alias("out", oB);
}
TM HZ_FN void SMTHash2Pair(Comp<M>&& cA, const V<M>& LA, const V<M>& RA, Comp<M>&& cB, const V<M>& LB,
                           const V<M>& RB, V<M>& oA, V<M>& oB) {
    V<M> a[2], b[2];
    a[0] = cA.def("L", LA); a[1] = cA.def("R", RA);
    b[0] = cB.def("L", LB); b[1] = cB.def("R", RB);
    PoseidonDual<M>(cA.sub("h"), a, cB.sub("h"), b, 2, oA, oB);
    cA.alias("out", oA); cB.alias("out", oB);
}

// 1 / x[i] for all i with one field inversion (Montgomery's trick); 1/0 = 0 as everywhere else
TM struct BatchInv;
template <> struct BatchInv<MCompute> {
    static HZ_HD void run(Comp<MCompute>& c, const VC* x, VC* inv, int n) {
        Fr acc = fr_one();
        // prefix products over the non-zero entries, kept in inv[]
        for (int i = 0; i < n; i++) {
            inv[i].v = acc; inv[i].sig = false;
            if (!fr_is_zero(x[i].v)) acc = fr_mul(acc, x[i].v);
        }
        Fr t = c.inv1(acc);
        for (int i = n - 1; i >= 0; i--) {
            if (fr_is_zero(x[i].v)) { inv[i].v = fr_zero(); continue; }
            Fr pre = inv[i].v;
            inv[i].v = fr_mul(t, pre);
            t = fr_mul(t, x[i].v);
        }
    }
};
template <> struct BatchInv<MLayout> {
    static HZ_HD void run(Comp<MLayout>&, const VL*, VL* inv, int n) { for (int i = 0; i < n; i++) inv[i] = vl(); }
};

#define HZ_MAX_LEVELS 49
TM HZ_FN void SMTLevIns(Comp<M>&& c, int n, const V<M>& enabled, const V<M>* siblings, V<M>* levIns) {
    V<M> en = c.def("enabled", enabled);
    V<M> sib[HZ_MAX_LEVELS], isz[HZ_MAX_LEVELS];
    for (int i = 0; i < n; i++) sib[i] = c.def(Nm("siblings", i), siblings[i]);
    uint32_t li = c.sigs("levIns", n);
    uint32_t dn = c.sigs("done", n - 2);       // done[n-2] is wired to levIns[n-1]: no entry of its own
    {
        V<M> inv[HZ_MAX_LEVELS];
        BatchInv<M>::run(c, sib, inv, n);
        for (int i = 0; i < n; i++) isz[i] = IsZeroWithInv(c.sub(Nm("isZero", i)), sib[i], inv[i]);
    }
    HZ_CHECK0(c, (isz[n - 1] - 1) * en);
    levIns[n - 1] = c.set(li + n - 1, 1 - isz[n - 2]);
    V<M> done = c.alias(Nm("done", n - 2), levIns[n - 1]);
    for (int i = n - 2; i > 0; i--) {
        levIns[i] = c.set(li + i, (1 - done) * (1 - isz[i - 1]));
        done = c.set(dn + i - 1, levIns[i] + done);
    }
    levIns[0] = c.set(li, 1 - done);
}

struct SMState { /* marker */ };
// prev / st tuples: (top, old0, bot, new1, na, upd)
TM HZ_FN void SMTProcessorSM(Comp<M>&& c, const V<M>& xor_, const V<M>& is0, const V<M>& levIns,
                             const V<M>* fnc, const V<M>* prev, V<M>* st) {
    V<M> x = c.def("xor", xor_);
    V<M> z = c.def("is0", is0);
    V<M> li = c.def("levIns", levIns);
    V<M> f0 = c.def(Nm("fnc", 0), fnc[0]);
    c.def(Nm("fnc", 1), fnc[1]);
    V<M> p_top = c.def("prev_top", prev[0]);
    V<M> p_old0 = c.def("prev_old0", prev[1]);
    V<M> p_bot = c.def("prev_bot", prev[2]);
    V<M> p_new1 = c.def("prev_new1", prev[3]);
    V<M> p_na = c.def("prev_na", prev[4]);
    V<M> p_upd = c.def("prev_upd", prev[5]);
    uint32_t s_top = c.sig("st_top"), s_old0 = c.sig("st_old0"), s_bot = c.sig("st_bot");
    uint32_t s_new1 = c.sig("st_new1"), s_na = c.sig("st_na"), s_upd = c.sig("st_upd");
    V<M> aux1 = c.def("aux1", p_top * li);
    V<M> aux2 = c.def("aux2", aux1 * f0);
    st[0] = c.set(s_top, p_top - aux1);
    st[1] = c.set(s_old0, aux2 * z);
    st[3] = c.set(s_new1, (aux2 - st[1] + p_bot) * x);
    st[2] = c.set(s_bot, (1 - x) * (aux2 - st[1] + p_bot));
    st[5] = c.set(s_upd, aux1 - aux2);
    st[4] = c.set(s_na, p_new1 + p_old0 + p_na + p_upd);
}

#define HZ_SZ_SMTHASH2_WIRED (poseidon_size(3))
TM HZ_FN void SMTProcessorLevel(Comp<M>&& c, const V<M>* st, const V<M>& sibling, const V<M>& old1leaf,
                                const V<M>& new1leaf, const V<M>& newlrbit, const V<M>& oldChild,
                                const V<M>& newChild, V<M>& oldRootOut, V<M>& newRootOut) {
    V<M> st_top = c.def("st_top", st[0]);
    V<M> st_old0 = c.def("st_old0", st[1]);
    V<M> st_bot = c.def("st_bot", st[2]);
    V<M> st_new1 = c.def("st_new1", st[3]);
    c.def("st_na", st[4]);
    V<M> st_upd = c.def("st_upd", st[5]);
    uint32_t oldRoot = c.sig("oldRoot"), newRoot = c.sig("newRoot");
    V<M> sib = c.def("sibling", sibling);
    V<M> o1 = c.def("old1leaf", old1leaf);
    V<M> n1 = c.def("new1leaf", new1leaf);
    V<M> lr = c.def("newlrbit", newlrbit);
    V<M> oc = c.def("oldChild", oldChild);
    V<M> nc = c.def("newChild", newChild);
    uint32_t aux = c.sigs("aux", 4);
    Comp<M> oldProofHash = c.sub_reserve("oldProofHash", HZ_SZ_SMTHASH2_WIRED);
    Comp<M> newProofHash = c.sub_reserve("newProofHash", HZ_SZ_SMTHASH2_WIRED);
    Comp<M> oldSwitcher = c.sub_reserve("oldSwitcher", 3);
    Comp<M> newSwitcher = c.sub_reserve("newSwitcher", 5);
    V<M> oL, oR, nL, nR, oph, nph;
    Switcher(HZ_MV(oldSwitcher), lr, oc, sib, oL, oR);
    V<M> a1 = c.set(aux + 1, nc * (st_top + st_bot));
    V<M> a2 = c.set(aux + 2, sib * st_top);
    Switcher(HZ_MV(newSwitcher), lr, a1 + n1 * st_new1, a2 + o1 * st_new1, nL, nR);
    SMTHash2Pair<M>(HZ_MV(oldProofHash), oL, oR, HZ_MV(newProofHash), nL, nR, oph, nph);
    V<M> a0 = c.set(aux + 0, o1 * (st_bot + st_new1 + st_upd));
    oldRootOut = c.set(oldRoot, a0 + oph * st_top);
    V<M> a3 = c.set(aux + 3, nph * (st_top + st_bot + st_new1));
    newRootOut = c.set(newRoot, a3 + n1 * (st_old0 + st_upd));
}

// SMTProcessor in two parts.  The level hashes (the body: everything up to topSwitcher) need the
// siblings and the leaf values but not `oldRoot`; only the top (checkOldInput, newRoot, keysOk)
// does.  The device evaluates the bodies of all processors of a batch at once and the tops
// afterwards, when processor 2's old root (= processor 1's new root, src/rollup-tx.circom:553-560)
// is known; everywhere else the two run back to back.
TM struct SMTMid { V<M> oldRoot, enabled, oldKey, newKey, fnc0, fnc1, oc, nc; uint32_t newRootIdx; };
// what the head of the body leaves for the levels
TM struct SMTHead { V<M> sib[HZ_MAX_LEVELS]; V<M> hash1Old, hash1New; V<M> sm[HZ_MAX_LEVELS][6]; BV<M> n2bNew; };
// witness entries of the nLevels level components: 1282 each (oldRoot, newRoot, aux[4], two wired SMTHash2, the
// two switchers), the bottom one (instantiated first) with two more for its constant children
#define HZ_SZ_SMT_LEVEL (2 + 4 + 2 * HZ_SZ_SMTHASH2_WIRED + 3 + 5)
HZ_HD uint32_t hz_smt_levels_size(int nLevels) { return (uint32_t)nLevels * HZ_SZ_SMT_LEVEL + 2; }

// everything of the body before the levels: leaf hashes, key bits, level-insert flags, state machine
TM HZ_FN void SMTProcessorHead(Comp<M>& c, int nLevels, const V<M>& oldRoot, const V<M>* siblings,
                               const V<M>& oldKey, const V<M>& oldValue, const V<M>& isOld0,
                               const V<M>& newKey, const V<M>& newValue, const V<M>* fnc, SMTMid<M>& mid, SMTHead<M>& hd) {
    mid.oldRoot = c.def("oldRoot", oldRoot);
    mid.newRootIdx = c.sig("newRoot");
    for (int i = 0; i < nLevels; i++) hd.sib[i] = c.def(Nm("siblings", i), siblings[i]);
    V<M> oldKey_ = c.def("oldKey", oldKey);
    V<M> oldValue_ = c.def("oldValue", oldValue);
    V<M> isOld0_ = c.def("isOld0", isOld0);
    V<M> newKey_ = c.def("newKey", newKey);
    V<M> newValue_ = c.def("newValue", newValue);
    V<M> f[2];
    f[0] = c.def(Nm("fnc", 0), fnc[0]);
    f[1] = c.def(Nm("fnc", 1), fnc[1]);
    V<M> enabled = c.def("enabled", f[0] + f[1] - f[0] * f[1]);
    {
        Comp<M> h1o = c.sub_reserve("hash1Old", HZ_SZ_SMTHASH1_WIRED);
        Comp<M> h1n = c.sub_reserve("hash1New", HZ_SZ_SMTHASH1_WIRED);
        SMTHash1Pair<M>(HZ_MV(h1o), oldKey_, oldValue_, HZ_MV(h1n), newKey_, newValue_, hd.hash1Old, hd.hash1New);
    }
    BV<M> n2bOld = Num2Bits_strict(c.sub("n2bOld"), oldKey_);
    hd.n2bNew = Num2Bits_strict(c.sub("n2bNew"), newKey_);
    V<M> levIns[HZ_MAX_LEVELS];
    SMTLevIns(c.sub("smtLevIns"), nLevels, enabled, hd.sib, levIns);
    V<M> xors[HZ_MAX_LEVELS];
    for (int i = 0; i < nLevels; i++) xors[i] = XOR(c.sub(Nm("xors", i)), n2bOld.get(i), hd.n2bNew.get(i));
    V<M> first[6];
    first[0] = enabled; first[1] = Ops<M>::K(0); first[2] = Ops<M>::K(0); first[3] = Ops<M>::K(0);
    first[4] = 1 - enabled; first[5] = Ops<M>::K(0);
    // (the previous level's tuple is read where it lies: copying it was 10 % of the stall samples of k_smt)
    for (int i = 0; i < nLevels; i++) SMTProcessorSM(c.sub(Nm("sm", i)), xors[i], isOld0_, levIns[i], f, i == 0 ? first : hd.sm[i - 1], hd.sm[i]);
    HZ_CHECK(c, hd.sm[nLevels - 1][4] + hd.sm[nLevels - 1][3] + hd.sm[nLevels - 1][1] + hd.sm[nLevels - 1][5], Ops<M>::K(1));
    mid.enabled = enabled; mid.oldKey = oldKey_; mid.newKey = newKey_; mid.fnc0 = f[0]; mid.fnc1 = f[1];
}
// the levels, bottom up: old-path and new-path roots of level 0
TM HZ_FN void SMTProcessorLevels(Comp<M>& c, int nLevels, const SMTHead<M>& hd, V<M>& oc, V<M>& nc) {
    oc = Ops<M>::K(0); nc = Ops<M>::K(0);
    for (int i = nLevels - 1; i >= 0; i--) {
        V<M> o, n;
        SMTProcessorLevel(c.sub(Nm("levels", i)), hd.sm[i], hd.sib[i], hd.hash1Old, hd.hash1New, hd.n2bNew.get(i), oc, nc, o, n);
        oc = o; nc = n;
    }
}
TM HZ_FN void SMTProcessorBody(Comp<M>& c, int nLevels, const V<M>& oldRoot, const V<M>* siblings,
                               const V<M>& oldKey, const V<M>& oldValue, const V<M>& isOld0,
                               const V<M>& newKey, const V<M>& newValue, const V<M>* fnc, SMTMid<M>& mid) {
    SMTHead<M> hd;
    SMTProcessorHead<M>(c, nLevels, oldRoot, siblings, oldKey, oldValue, isOld0, newKey, newValue, fnc, mid, hd);
    SMTProcessorLevels<M>(c, nLevels, hd, mid.oc, mid.nc);
}

TM HZ_FN V<M> SMTProcessorTop(Comp<M>& c, const SMTMid<M>& mid) {
    V<M> tL, tR;
    Switcher(c.sub("topSwitcher"), mid.fnc0 * mid.fnc1, mid.oc, mid.nc, tL, tR);
    ForceEqualIfEnabled(c.sub("checkOldInput"), mid.enabled, mid.oldRoot, tL);
    V<M> nr = c.set(mid.newRootIdx, mid.enabled * (tR - mid.oldRoot) + mid.oldRoot);
    V<M> eq = IsEqual(c.sub("areKeyEquals"), mid.oldKey, mid.newKey);
    V<M> xs[3];
    xs[0] = 1 - mid.fnc0; xs[1] = mid.fnc1; xs[2] = 1 - eq;
    V<M> ok = MultiAND(c.sub("keysOk"), xs, 3);
    HZ_CHECK0(c, ok);
    return nr;
}

TM HZ_FN V<M> SMTProcessor(Comp<M>&& c, int nLevels, const V<M>& oldRoot, const V<M>* siblings,
                            const V<M>& oldKey, const V<M>& oldValue, const V<M>& isOld0,
                            const V<M>& newKey, const V<M>& newValue, const V<M>* fnc) {
    SMTMid<M> mid;
    SMTProcessorBody<M>(c, nLevels, oldRoot, siblings, oldKey, oldValue, isOld0, newKey, newValue, fnc, mid);
    return SMTProcessorTop<M>(c, mid);
}

// prev / st tuples: (top, i0, iold, inew, na)
TM HZ_FN void SMTVerifierSM(Comp<M>&& c, const V<M>& is0, const V<M>& levIns, const V<M>& fnc,
                            const V<M>* prev, V<M>* st) {
    V<M> z = c.def("is0", is0);
    V<M> li = c.def("levIns", levIns);
    V<M> f = c.def("fnc", fnc);
    V<M> p_top = c.def("prev_top", prev[0]);
    V<M> p_i0 = c.def("prev_i0", prev[1]);
    V<M> p_iold = c.def("prev_iold", prev[2]);
    V<M> p_inew = c.def("prev_inew", prev[3]);
    V<M> p_na = c.def("prev_na", prev[4]);
    uint32_t s_top = c.sig("st_top"), s_i0 = c.sig("st_i0"), s_iold = c.sig("st_iold");
    uint32_t s_inew = c.sig("st_inew"), s_na = c.sig("st_na");
    V<M> a = c.def("prev_top_lev_ins", p_top * li);
    V<M> b = c.def("prev_top_lev_ins_fnc", a * f);
    st[0] = c.set(s_top, p_top - a);
    st[3] = c.set(s_inew, a - b);
    st[2] = c.set(s_iold, b * (1 - z));
    st[1] = c.set(s_i0, a * z);
    st[4] = c.set(s_na, p_na + p_inew + p_iold + p_i0);
}

TM HZ_FN V<M> SMTVerifierLevel(Comp<M>&& c, const V<M>* st, const V<M>& sibling, const V<M>& old1leaf,
                               const V<M>& new1leaf, const V<M>& lrbit, const V<M>& child) {
    V<M> st_top = c.def("st_top", st[0]);
    c.def("st_i0", st[1]);
    V<M> st_iold = c.def("st_iold", st[2]);
    V<M> st_inew = c.def("st_inew", st[3]);
    c.def("st_na", st[4]);
    uint32_t root = c.sig("root");
    V<M> sib = c.def("sibling", sibling);
    V<M> o1 = c.def("old1leaf", old1leaf);
    V<M> n1 = c.def("new1leaf", new1leaf);
    V<M> lr = c.def("lrbit", lrbit);
    V<M> ch = c.def("child", child);
    uint32_t aux = c.sigs("aux", 2);
    Comp<M> proofHash = c.sub_reserve("proofHash", HZ_SZ_SMTHASH2_WIRED);
    V<M> sL, sR;
    Switcher(c.sub("switcher"), lr, ch, sib, sL, sR);
    V<M> ph = SMTHash2(HZ_MV(proofHash), sL, sR);
    V<M> a0 = c.set(aux + 0, ph * st_top);
    V<M> a1 = c.set(aux + 1, o1 * st_iold);
    return c.set(root, a0 + a1 + n1 * st_inew);
}

TM HZ_FN void SMTVerifier(Comp<M>&& c, int nLevels, const V<M>& enabled, const V<M>& root,
                           const V<M>* siblings, const V<M>& oldKey, const V<M>& oldValue,
                           const V<M>& isOld0, const V<M>& key, const V<M>& value, const V<M>& fnc) {
    V<M> en = c.def("enabled", enabled);
    V<M> root_ = c.def("root", root);
    V<M> sib[HZ_MAX_LEVELS];
    for (int i = 0; i < nLevels; i++) sib[i] = c.def(Nm("siblings", i), siblings[i]);
    V<M> oldKey_ = c.def("oldKey", oldKey);
    V<M> oldValue_ = c.def("oldValue", oldValue);
    V<M> isOld0_ = c.def("isOld0", isOld0);
    V<M> key_ = c.def("key", key);
    V<M> value_ = c.def("value", value);
    V<M> fnc_ = c.def("fnc", fnc);
    V<M> hash1Old = SMTHash1(c.sub("hash1Old"), oldKey_, oldValue_);
    V<M> hash1New = SMTHash1(c.sub("hash1New"), key_, value_);
    Num2Bits_strict(c.sub("n2bOld"), oldKey_);
    BV<M> n2bNew = Num2Bits_strict(c.sub("n2bNew"), key_);
    V<M> levIns[HZ_MAX_LEVELS];
    SMTLevIns(c.sub("smtLevIns"), nLevels, en, sib, levIns);
    V<M> sm[HZ_MAX_LEVELS][5];
    for (int i = 0; i < nLevels; i++) {
        V<M> prev[5];
        if (i == 0) {
            prev[0] = en; prev[1] = Ops<M>::K(0); prev[2] = Ops<M>::K(0); prev[3] = Ops<M>::K(0); prev[4] = 1 - en;
        } else {
            for (int k = 0; k < 5; k++) prev[k] = sm[i - 1][k];
        }
        SMTVerifierSM(c.sub(Nm("sm", i)), isOld0_, levIns[i], fnc_, prev, sm[i]);
    }
    HZ_CHECK(c, sm[nLevels - 1][4] + sm[nLevels - 1][2] + sm[nLevels - 1][3] + sm[nLevels - 1][1], Ops<M>::K(1));
    V<M> child = Ops<M>::K(0);
    for (int i = nLevels - 1; i >= 0; i--)
        child = SMTVerifierLevel(c.sub(Nm("levels", i)), sm[i], sib[i], hash1Old, hash1New, n2bNew.get(i), child);
    V<M> eq = IsEqual(c.sub("areKeyEquals"), oldKey_, key_);
    V<M> xs[4];
    xs[0] = fnc_; xs[1] = 1 - isOld0_; xs[2] = eq; xs[3] = en;
    V<M> ok = MultiAND(c.sub("keysOk"), xs, 4);
    HZ_CHECK0(c, ok);
    ForceEqualIfEnabled(c.sub("checkRoot"), en, child, root_);
}

// ---------------------------------------------------------------------------
// babyjub.circom / montgomery.circom   
// ---------------------------------------------------------------------------
#define HZ_BJ_A 168700ull
#define HZ_BJ_D 168696ull
#define HZ_MONT_A 168698ull
TM struct PtV { V<M> x, y; };

HZ_HD Fr fr_base8_x() {
    // 5299619240641551281634865583518297030282874472190772894086521144482721001553
    Fr r; const uint32_t h[8] = HZ_BASE8_X_LIMBS; for (int i = 0; i < 8; i++) r.l[i] = h[i]; return r;
}
HZ_HD Fr fr_base8_y() {
    // 16950150798460657717958625567821834550301663161624707787222815936182638968203
    Fr r; const uint32_t h[8] = HZ_BASE8_Y_LIMBS; for (int i = 0; i < 8; i++) r.l[i] = h[i]; return r;
}

TM struct HintGen;
template <> struct HintGen<MCompute> {
#ifdef HZ_DEVICE_ONLY
    // device: the hints of every ladder of the signature were produced by k_ed_ladder / k_ed_hints before the
    // second half of the verifier runs (c.hints->lad_ok: they are usable)
    static HZ_HD void seg_any(Comp<MCompute>& c, const VC&, const VC&, const BVC&, int, int seg) {
        if (!c.hints) return;
        c.hints->dat = c.hints->lad_ok ? c.hints->scratch + (seg == 0 ? 0 : 295) : nullptr;
        c.hints->di = 0;
    }
    static HZ_HD void mid(Comp<MCompute>& c) {
        if (!c.hints) return;
        c.hints->dat = c.hints->lad_ok ? c.hints->scratch + 294 : nullptr;
        c.hints->di = 0;
    }
    static HZ_HD void seg_fix(Comp<MCompute>& c, const BVC& e, int nW, int seg) {
        if (!c.hints) return;
        if (seg == 0 && c.hints->lad_ok) { c.hints->dat = c.hints->scratch + 505; c.hints->di = 0; }
        else hz_hints_segment_fix(c.hints, e, nW, seg, c.hints->scratch + HZ_SCR_ARR);   // the ladder's hints stay for k_ed_fill
    }
    // the steps of a ladder are left to k_ed_fill: reserve their entries, note where they start and go on from
    // the points the ladder ended at
    static HZ_HD bool fill_on(const Comp<MCompute>& c) { return c.hints && c.hints->fill && c.hints->lad_ok; }
    static HZ_HD VC aff(const Comp<MCompute>& c, int k) { VC r; r.v = c.hints->scratch[HZ_SCR_AFF + k]; r.sig = true; return r; }
    static HZ_HD bool fill_any(Comp<MCompute>& c, int n, int seg, PtV<MCompute>& d, PtV<MCompute>& a) {
        if (!fill_on(c)) return false;
        HzFillRec* rec = reinterpret_cast<HzFillRec*>(c.hints->scratch + HZ_SCR_REC);
        if (hz_glane() == 0) { if (seg == 0) rec->base_seg0 = c.cur; else rec->base_seg1 = c.cur; }
        c.cur += 9u * (uint32_t)(n - 1);
        if (seg == 0) { d.x = aff(c, 2 * 77); d.y = aff(c, 2 * 77 + 1); }      // segment 1's dbl output is wired nowhere
        a.x = aff(c, 2 * (78 + seg)); a.y = aff(c, 2 * (78 + seg) + 1);
        return true;
    }
    // windows and carry adders of segment 0: constants (prefilled by k_ed_const) and the window multiplexers,
    // which only depend on the scalar bits (k_ed_fill); nothing of them is needed here
    static HZ_HD bool fill_windows(Comp<MCompute>& c, int seg, PtV<MCompute>& wbase, PtV<MCompute>& cad, PtV<MCompute>& dblLast) {
        if (seg != 0 || !fill_on(c) || !c.hints->prefilled || c.hints->win_size == 0) return false;
        HzFillRec* rec = reinterpret_cast<HzFillRec*>(c.hints->scratch + HZ_SCR_REC);
        if (hz_glane() == 0) rec->base_win = c.cur;
        c.cur += c.hints->win_size;
        VC z; z.v = fr_zero(); z.sig = true;          // wired signals whose values nothing downstream reads
        wbase.x = wbase.y = cad.x = cad.y = dblLast.x = dblLast.y = z;
        return true;
    }
    static HZ_HD bool fill_fix(Comp<MCompute>& c, int nW, int seg, PtV<MCompute>& acc) {
        if (seg != 0 || !fill_on(c)) return false;
        HzFillRec* rec = reinterpret_cast<HzFillRec*>(c.hints->scratch + HZ_SCR_REC);
        if (hz_glane() == 0) rec->base_fix = c.cur;
        c.cur += 3u * (uint32_t)nW;
        acc.x = aff(c, 2 * 80); acc.y = aff(c, 2 * 80 + 1);
        return true;
    }
#else
    static HZ_HD bool fill_any(Comp<MCompute>&, int, int, PtV<MCompute>&, PtV<MCompute>&) { return false; }
    static HZ_HD bool fill_fix(Comp<MCompute>&, int, int, PtV<MCompute>&) { return false; }
    static HZ_HD bool fill_windows(Comp<MCompute>&, int, PtV<MCompute>&, PtV<MCompute>&, PtV<MCompute>&) { return false; }
    static HZ_HD void seg_any(Comp<MCompute>& c, const VC& px, const VC& py, const BVC& e, int n, int) { hz_hints_segment_any(c.hints, px.v, py.v, e, n); }
    static HZ_HD void mid(Comp<MCompute>&) {}
    static HZ_HD void seg_fix(Comp<MCompute>& c, const BVC& e, int nW, int seg) { hz_hints_segment_fix(c.hints, e, nW, seg); }
#endif
    // record mode: where the windows / carry adders of segment 0 start and how many entries they take
    static HZ_HD void win_begin(Comp<MCompute>& c, int seg) { if (c.hints && c.hints->rec && seg == 0) c.hints->win_c0 = c.cur; }
    static HZ_HD void win_end(Comp<MCompute>& c, int seg) { if (c.hints && c.hints->rec && seg == 0) c.hints->win_size = c.cur - c.hints->win_c0; }
    static HZ_HD void done(Comp<MCompute>& c) { hz_hints_done(c.hints); }
    static HZ_HD void fix_begin(Comp<MCompute>& c) { if (c.hints) c.hints->fix_shift = c.cur; }
    static HZ_HD void record(Comp<MCompute>& c, const VC& u, const VC& v) { hz_hints_record_point(c.hints, u.v, v.v); }
};
#ifndef HZ_DEVICE_ONLY
template <> struct HintGen<MLayout> {
    static void seg_any(Comp<MLayout>&, const VL&, const VL&, const BVT<MLayout, 8>&, int, int) {}
    static void mid(Comp<MLayout>&) {}
    static void seg_fix(Comp<MLayout>&, const BVT<MLayout, 8>&, int, int) {}
    static bool fill_any(Comp<MLayout>&, int, int, PtV<MLayout>&, PtV<MLayout>&) { return false; }
    static bool fill_fix(Comp<MLayout>&, int, int, PtV<MLayout>&) { return false; }
    static bool fill_windows(Comp<MLayout>&, int, PtV<MLayout>&, PtV<MLayout>&, PtV<MLayout>&) { return false; }
    static void win_begin(Comp<MLayout>&, int) {}
    static void win_end(Comp<MLayout>&, int) {}
    static void done(Comp<MLayout>&) {}
    static uint32_t& last_fix_begin() { static uint32_t v = 0; return v; }     // where the last mulFix walked starts
    static void fix_begin(Comp<MLayout>& c) { last_fix_begin() = c.cur; }
    static void record(Comp<MLayout>&, const VL&, const VL&) {}
};
#endif

#define HZ_SZ_BABYADD_WIRED 6
TM HZ_FN PtV<M> BabyAdd(Comp<M>&& c, const V<M>& x1_, const V<M>& y1_, const V<M>& x2_, const V<M>& y2_) {
    V<M> x1 = c.def("x1", x1_), y1 = c.def("y1", y1_), x2 = c.def("x2", x2_), y2 = c.def("y2", y2_);
    uint32_t xout = c.sig("xout"), yout = c.sig("yout");
    V<M> beta = c.def("beta", x1 * y2);
    V<M> gamma = c.def("gamma", y1 * x2);
    V<M> delta = c.def("delta", (y1 - HZ_BJ_A * x1) * (x2 + y2));
    V<M> tau = c.def("tau", beta * gamma);
    V<M> dt = HZ_BJ_D * tau;
    PtV<M> r;
    V<M> qx, qy;
    c.div2(beta + gamma, 1 + dt, delta + HZ_BJ_A * beta - gamma, 1 - dt, qx, qy);
    r.x = c.set(xout, qx);
    HZ_CHECK(c, (1 + dt) * r.x, beta + gamma);
    r.y = c.set(yout, qy);
    HZ_CHECK(c, (1 - dt) * r.y, delta + HZ_BJ_A * beta - gamma);
    return r;
}
TM HZ_FN PtV<M> BabyDbl(Comp<M>&& c, const V<M>& x, const V<M>& y) {
    V<M> x_ = c.def("x", x), y_ = c.def("y", y);
    PtV<M> r = BabyAdd(c.sub("adder"), x_, y_, x_, y_);
    c.alias("xout", r.x); c.alias("yout", r.y);
    return r;
}
TM HZ_FN void BabyCheck(Comp<M>&& c, const V<M>& x, const V<M>& y) {
    V<M> x_ = c.def("x", x), y_ = c.def("y", y);
    V<M> x2 = c.def("x2", x_ * x_);
    V<M> y2 = c.def("y2", y_ * y_);
    HZ_CHECK(c, HZ_BJ_A * x2 + y2, 1 + HZ_BJ_D * (x2 * y2));
}
TM HZ_FN PtV<M> Edwards2Montgomery(Comp<M>&& c, const V<M>& x, const V<M>& y, int hc = 0) {
    V<M> i0 = c.def(Nm("in", 0), x), i1 = c.def(Nm("in", 1), y);
    PtV<M> r;
    if (hc == 1 && c.const_image()) { r.x = c.copy_const(Nm("out", 0)); r.y = c.copy_const(Nm("out", 1)); return r; }
    if (hc == 0) {
        // u = (1+y)/(1-y), v = u/x = (1+y)/((1-y) x): one inversion for both
        V<M> u, v;
        c.div2(1 + i1, 1 - i1, 1 + i1, (1 - i1) * i0, u, v);
        r.x = c.def(Nm("out", 0), u);
        r.y = c.def(Nm("out", 1), v);
    } else {
        r.x = c.def(Nm("out", 0), c.div(1 + i1, 1 - i1, hc));
        r.y = c.def(Nm("out", 1), c.div(r.x, i0, hc));
    }
    HZ_CHECK(c, r.x * (1 - i1), 1 + i1);
    HZ_CHECK(c, r.y * i0, r.x);
    return r;
}
TM HZ_FN PtV<M> Montgomery2Edwards(Comp<M>&& c, const V<M>& u, const V<M>& v, int hc = 0) {
    V<M> i0 = c.def(Nm("in", 0), u), i1 = c.def(Nm("in", 1), v);
    PtV<M> r;
    if (hc == 1 && c.const_image()) { r.x = c.copy_const(Nm("out", 0)); r.y = c.copy_const(Nm("out", 1)); return r; }
    if (hc == 0) {
        V<M> x, y;
        c.div2(i0, i1, i0 - 1, i0 + 1, x, y);
        r.x = c.def(Nm("out", 0), x);
        r.y = c.def(Nm("out", 1), y);
    } else {
        r.x = c.def(Nm("out", 0), c.div(i0, i1, hc));
        r.y = c.def(Nm("out", 1), c.div(i0 - 1, i0 + 1, hc));
    }
    HZ_CHECK(c, r.x * i1, i0);
    HZ_CHECK(c, r.y * (i0 + 1), i0 - 1);
    return r;
}
TM HZ_FN PtV<M> MontgomeryAdd(Comp<M>&& c, const PtV<M>& p1, const PtV<M>& p2, int hc = 0) {
    V<M> a0 = c.def(Nm("in1", 0), p1.x), a1 = c.def(Nm("in1", 1), p1.y);
    V<M> b0 = c.def(Nm("in2", 0), p2.x), b1 = c.def(Nm("in2", 1), p2.y);
    if (hc == 1 && c.const_image()) {
        PtV<M> q;
        q.x = c.copy_const(Nm("out", 0)); q.y = c.copy_const(Nm("out", 1)); c.copy_const("lamda");
        return q;
    }
    uint32_t out = c.sigs("out", 2);
    V<M> lam = c.def("lamda", c.div(b1 - a1, b0 - a0, hc));
    HZ_CHECK(c, lam * (b0 - a0), b1 - a1);
    PtV<M> r;
    r.x = c.set(out, lam * lam - HZ_MONT_A - a0 - b0);
    r.y = c.set(out + 1, lam * (a0 - r.x) - a1);
    return r;
}
TM HZ_FN PtV<M> MontgomeryDouble(Comp<M>&& c, const PtV<M>& p, int hc = 0) {
    V<M> i0 = c.def(Nm("in", 0), p.x), i1 = c.def(Nm("in", 1), p.y);
    if (hc == 1 && c.const_image()) {
        PtV<M> q;
        q.x = c.copy_const(Nm("out", 0)); q.y = c.copy_const(Nm("out", 1)); c.copy_const("lamda"); c.copy_const("x1_2");
        return q;
    }
    uint32_t out = c.sigs("out", 2);
    uint32_t lamda = c.sig("lamda");
    V<M> x1_2 = c.def("x1_2", i0 * i0);
    V<M> num = 3 * x1_2 + (2 * HZ_MONT_A) * i0 + 1;
    V<M> lam = c.set(lamda, c.div(num, 2 * i1, hc));
    HZ_CHECK(c, lam * (2 * i1), num);
    PtV<M> r;
    r.x = c.set(out, lam * lam - HZ_MONT_A - 2 * i0);
    r.y = c.set(out + 1, lam * (i0 - r.x) - i1);
    return r;
}

// ---------------------------------------------------------------------------
// escalarmulany.circom
// ---------------------------------------------------------------------------
TM HZ_FN PtV<M> Multiplexor2(Comp<M>&& c, const V<M>& sel, const PtV<M>& in0, const PtV<M>& in1) {
    V<M> s = c.def("sel", sel);
    V<M> a0 = c.def(Nm("in", 0, 0), in0.x), a1 = c.def(Nm("in", 0, 1), in0.y);
    V<M> b0 = c.def(Nm("in", 1, 0), in1.x), b1 = c.def(Nm("in", 1, 1), in1.y);
    PtV<M> r;
    r.x = c.def(Nm("out", 0), (b0 - a0) * s + a0);
    r.y = c.def(Nm("out", 1), (b1 - a1) * s + a1);
    return r;
}
TM HZ_FN void BitElementMulAny(Comp<M>&& c, const V<M>& sel, PtV<M>& dbl, PtV<M>& add) {
    V<M> s = c.def("sel", sel);
    PtV<M> d, a;
    d.x = c.def(Nm("dblIn", 0), dbl.x); d.y = c.def(Nm("dblIn", 1), dbl.y);
    a.x = c.def(Nm("addIn", 0), add.x); a.y = c.def(Nm("addIn", 1), add.y);
    PtV<M> dd = MontgomeryDouble(c.sub("doubler"), d, 2);
    PtV<M> ad = MontgomeryAdd(c.sub("adder"), dd, a, 2);
    PtV<M> so = Multiplexor2(c.sub("selector"), s, a, ad);
    c.alias(Nm("dblOut", 0), dd.x); c.alias(Nm("dblOut", 1), dd.y);
    c.alias(Nm("addOut", 0), so.x); c.alias(Nm("addOut", 1), so.y);
    dbl = dd; add = so;
}
// wired size of SegmentMulAny(n) (e and p wired signals)
#define HZ_SZ_SEGMENTMULANY(n) (13 + 9 * ((n) - 1))
TM HZ_FN void SegmentMulAny(Comp<M>&& c, const BV<M>& e, int from, int n, const PtV<M>& p, PtV<M>& out, PtV<M>& dblOut, int seg) {
    BV<M> e_; e_.clear(n);
    for (int i = 0; i < n; i++) e_.put(i, c.def(Nm("e", i), e.get(from + i)));
    PtV<M> p_;
    p_.x = c.def(Nm("p", 0), p.x); p_.y = c.def(Nm("p", 1), p.y);
    PtV<M> m = Edwards2Montgomery(c.sub("e2m"), p_.x, p_.y);
    PtV<M> d = m, a = m;
    HintGen<M>::seg_any(c, p_.x, p_.y, e_, n, seg);
    if (!HintGen<M>::fill_any(c, n, seg, d, a))
        for (int i = 0; i < n - 1; i++) BitElementMulAny(c.sub(Nm("bits", i)), e_.get(i + 1), d, a);
    HintGen<M>::done(c);
    c.alias(Nm("dbl", 0), d.x); c.alias(Nm("dbl", 1), d.y);
    PtV<M> ed = Montgomery2Edwards(c.sub("m2e"), a.x, a.y);
    PtV<M> ex = BabyAdd(c.sub("eadder"), ed.x, ed.y, -p_.x, p_.y);
    PtV<M> ls = Multiplexor2(c.sub("lastSel"), e_.get(0), ex, ed);
    c.alias(Nm("out", 0), ls.x); c.alias(Nm("out", 1), ls.y);
    out = ls; dblOut = d;
}
// EscalarMulAny(254) as used by EdDSAPoseidonVerifier: two segments (148 + 106 bits)
TM HZ_FN PtV<M> EscalarMulAny254(Comp<M>&& c, const BV<M>& e, const PtV<M>& p) {
    const int n = 254, nsegments = 2, nlast = n - 148;
    BV<M> e_; e_.clear(n);
    for (int i = 0; i < n; i++) e_.put(i, c.def(Nm("e", i), e.get(i)));
    PtV<M> p_;
    p_.x = c.def(Nm("p", 0), p.x); p_.y = c.def(Nm("p", 1), p.y);
    uint32_t out = c.sigs("out", 2);
    V<M> zp = IsZero(c.sub("zeropoint"), p_.x);
    PtV<M> base;
    base.x = p_.x + (Ops<M>::KF(fr_base8_x()) - p_.x) * zp;
    base.y = p_.y + (Ops<M>::KF(fr_base8_y()) - p_.y) * zp;
    PtV<M> o0, d0, o1, d1;
    SegmentMulAny(c.sub(Nm("segments", 0)), e_, 0, 148, base, o0, d0, 0);
    (void)nsegments;
    Comp<M> seg1 = c.sub_reserve(Nm("segments", 1), HZ_SZ_SEGMENTMULANY(nlast));
    HintGen<M>::mid(c);
    PtV<M> dd = MontgomeryDouble(c.sub(Nm("doublers", 0)), d0, 2);
    HintGen<M>::done(c);
    PtV<M> ed = Montgomery2Edwards(c.sub(Nm("m2e", 0)), dd.x, dd.y);
    Comp<M> add0 = c.sub_reserve(Nm("adders", 0), HZ_SZ_BABYADD_WIRED);
    SegmentMulAny(HZ_MV(seg1), e_, 148, nlast, ed, o1, d1, 1);
    PtV<M> fin = BabyAdd(HZ_MV(add0), o0.x, o0.y, o1.x, o1.y);
    PtV<M> r;
    r.x = c.set(out, fin.x * (1 - zp));
    r.y = c.set(out + 1, fin.y + (1 - fin.y) * zp);
    return r;
}

// ---------------------------------------------------------------------------
// escalarmulfix.circom
// ---------------------------------------------------------------------------
TM HZ_FN void WindowMulFix(Comp<M>&& c, const V<M>* in3, const PtV<M>& base, PtV<M>& out, PtV<M>& out8) {
    V<M> ins[3];
    for (int i = 0; i < 3; i++) ins[i] = c.def(Nm("in", i), in3[i]);
    PtV<M> b;
    b.x = c.def(Nm("base", 0), base.x); b.y = c.def(Nm("base", 1), base.y);
    Comp<M> mux = c.sub_reserve("mux", HZ_SZ_MULTIMUX3(2));
    PtV<M> pts[8];
    pts[0] = b;
    // the base is a constant point: every division in here reads the constant stream
    pts[1] = MontgomeryDouble(c.sub("dbl2"), b, 1);
    const char* adr[6] = {"adr3", "adr4", "adr5", "adr6", "adr7", "adr8"};
    for (int k = 2; k < 8; k++) pts[k] = MontgomeryAdd(c.sub(adr[k - 2]), b, pts[k - 1], 1);
    for (int k = 0; k < 8; k++) HintGen<M>::record(c, pts[k].x, pts[k].y);
    V<M> kk[2][8], mo[2];
    for (int k = 0; k < 8; k++) { kk[0][k] = pts[k].x; kk[1][k] = pts[k].y; }
    MultiMux3<M>(HZ_MV(mux), 2, kk, ins, mo);
    c.alias(Nm("out8", 0), pts[7].x); c.alias(Nm("out8", 1), pts[7].y);
    c.alias(Nm("out", 0), mo[0]); c.alias(Nm("out", 1), mo[1]);
    out.x = mo[0]; out.y = mo[1]; out8 = pts[7];
}

#define HZ_SZ_MONTDOUBLE_WIRED 4
// e: nW*3 bit values (signals, or constant 0 padding)
TM HZ_FN void SegmentMulFix(Comp<M>&& c, const BV<M>& e, int nW, const PtV<M>& base, PtV<M>& out, PtV<M>& dblOut, int seg) {
    BV<M> e_; e_.clear(nW * 3);
    for (int i = 0; i < nW * 3; i++) e_.put(i, c.def(Nm("e", i), e.get(i)));
    PtV<M> b;
    b.x = c.def(Nm("base", 0), base.x); b.y = c.def(Nm("base", 1), base.y);
    PtV<M> m = Edwards2Montgomery(c.sub("e2m"), b.x, b.y, 1);
    Comp<M> dblLast_c = c.sub_reserve("dblLast", HZ_SZ_MONTDOUBLE_WIRED);
    // windows and carry adders are instantiated first; the accumulating adders follow, so the
    // window outputs are kept: read them back from the witness when accumulating
    PtV<M> wbase = m, cad = m, dblLast, o8;
    PtV<M> win_out[85];
    HintGen<M>::win_begin(c, seg);
    if (!HintGen<M>::fill_windows(c, seg, wbase, cad, dblLast))
    for (int i = 0; i < nW; i++) {
        V<M> in3[3];
        for (int k = 0; k < 3; k++) in3[k] = e_.get(3 * i + k);
        WindowMulFix(c.sub(Nm("windows", i)), in3, wbase, win_out[i], o8);
        wbase = o8;
        if (i < nW - 1) {
            cad = MontgomeryAdd(c.sub(Nm("cadders", i)), cad, o8, 1);
        } else {
            dblLast = MontgomeryDouble(HZ_MV(dblLast_c), o8, 1);
            HintGen<M>::record(c, dblLast.x, dblLast.y);
            cad = MontgomeryAdd(c.sub(Nm("cadders", i)), cad, dblLast, 1);
        }
    }
    HintGen<M>::win_end(c, seg);
    // the accumulating adders depend on the scalar: data stream
    HintGen<M>::seg_fix(c, e_, nW, seg);
    PtV<M> acc = dblLast;
    if (!HintGen<M>::fill_fix(c, nW, seg, acc))
        for (int i = 0; i < nW; i++) acc = MontgomeryAdd(c.sub(Nm("adders", i)), acc, win_out[i], 2);
    HintGen<M>::done(c);
    PtV<M> ed = Montgomery2Edwards(c.sub("m2e"), acc.x, acc.y);
    PtV<M> ced = Montgomery2Edwards(c.sub("cm2e"), cad.x, cad.y, 1);
    PtV<M> cs = BabyAdd(c.sub("cAdd"), ed.x, ed.y, -ced.x, ced.y);
    c.alias(Nm("out", 0), cs.x); c.alias(Nm("out", 1), cs.y);
    c.alias(Nm("dbl", 0), wbase.x); c.alias(Nm("dbl", 1), wbase.y);
    out = cs; dblOut = wbase;
}

// EscalarMulFix(253, BASE8): segment 0 = 249 bits (83 windows), segment 1 = 4 bits (2 windows)
TM HZ_FN PtV<M> EscalarMulFix253(Comp<M>&& c, const BV<M>& e) {
    HintGen<M>::fix_begin(c);
    BV<M> e_; e_.clear(253);
    for (int i = 0; i < 253; i++) e_.put(i, c.def(Nm("e", i), e.get(i)));
    PtV<M> B;
    B.x = Ops<M>::KF(fr_base8_x()); B.y = Ops<M>::KF(fr_base8_y());
    PtV<M> o0, d0, o1, d1;
    {
        BV<M> b0; b0.clear(249);
        bv_copy(b0, 0, e_, 0, 249);
        SegmentMulFix(c.sub(Nm("segments", 0)), b0, 83, B, o0, d0, 0);
    }
    // wired SegmentMulFix(nW = 2) with two constant padding bits: 17 + 45 * 2 + 2
    Comp<M> seg1 = c.sub_reserve(Nm("segments", 1), 109);
    PtV<M> ed = Montgomery2Edwards(c.sub(Nm("m2e", 0)), d0.x, d0.y, 1);
    Comp<M> add0 = c.sub_reserve(Nm("adders", 0), HZ_SZ_BABYADD_WIRED);
    {
        BV<M> b1; b1.clear(6);
        bv_copy(b1, 0, e_, 249, 4);
        b1.put(4, Ops<M>::K(0)); b1.put(5, Ops<M>::K(0));
        SegmentMulFix(HZ_MV(seg1), b1, 2, ed, o1, d1, 1);
    }
    PtV<M> fin = BabyAdd(HZ_MV(add0), o0.x, o0.y, o1.x, o1.y);
    c.alias(Nm("out", 0), fin.x); c.alias(Nm("out", 1), fin.y);
    return fin;
}

// one-time pass over EscalarMulFix(253, BASE8) in record mode: the inverses of its constant
// division sites (in evaluation order) and its window entries as Edwards points (hints.cuh)
HZ_FN void hz_record_fix_consts(Fr* wtmp, Fr* cst, Fr* edtab, uint32_t* counts) {
    HzHints h;
    h.cst = nullptr; h.ci = 0; h.dat = nullptr; h.di = 0; h.rec = cst; h.ri = 0;
    h.edtab = nullptr; h.edrec = edtab; h.ei = 0; h.scratch = nullptr; h.lad_ok = 0; h.fill = 0; h.prefilled = 0; h.win_c0 = 0; h.win_size = 0; h.cimg = nullptr; h.fix_shift = 0;
    Comp<MCompute> c(wtmp, 0, nullptr, &h);
    BVC e; e.clear(253);
    VC z = vc(fr_zero()); z.sig = true;
    for (int i = 0; i < 253; i++) e.put(i, z);
    EscalarMulFix253<MCompute>(c.sub("mulFix"), e);
    counts[0] = h.ri; counts[1] = h.ei; counts[2] = h.win_c0; counts[3] = h.win_size; counts[4] = c.cur;
}

// ---------------------------------------------------------------------------
// eddsaposeidon.circom  (src/rollup-tx.circom:472)
// ---------------------------------------------------------------------------
// The verifier in two halves: everything before the scalar multiplications (bits of S and of the message hash,
// the three doublings of A) and the multiplications with the final comparison.  Host builds and the layout walk
// run them back to back; the device runs each half in a kernel of its own with the ladder walk (k_ed_ladder,
// k_ed_hints) between them, handing over what the second half reads in an EdMid.
TM struct EdMid { V<M> en, R8x, R8y; PtV<M> d3; BV<M> hbits, sbits; };
TM HZ_FN void EdDSAPoseidonVerifierHead(Comp<M>& c, const V<M>& enabled, const V<M>& Ax, const V<M>& Ay,
                                         const V<M>& S, const V<M>& R8x, const V<M>& R8y, const V<M>& Msg, EdMid<M>& m) {
    V<M> en = c.def("enabled", enabled);
    V<M> Ax_ = c.def("Ax", Ax), Ay_ = c.def("Ay", Ay), S_ = c.def("S", S);
    V<M> R8x_ = c.def("R8x", R8x), R8y_ = c.def("R8y", R8y), M_ = c.def("M", Msg);
    m.sbits = Num2Bits(c.sub("snum2bits"), 253, S_);
    {
        BV<M> cb; cb.clear(254);
        bv_copy(cb, 0, m.sbits, 0, 253);
        cb.put(253, Ops<M>::K(0));
        V<M> cc = CompConstant(c.sub("compConstant"), fr_const_suborder_m1(), cb);
        HZ_CHECK0(c, cc * en);
    }
    V<M> hin[5];
    hin[0] = R8x_; hin[1] = R8y_; hin[2] = Ax_; hin[3] = Ay_; hin[4] = M_;
    V<M> h = Poseidon<M>(c.sub("hash"), hin, 5);
    m.hbits = Num2Bits_strict(c.sub("h2bits"), h);
    PtV<M> d1 = BabyDbl(c.sub("dbl1"), Ax_, Ay_);
    PtV<M> d2 = BabyDbl(c.sub("dbl2"), d1.x, d1.y);
    m.d3 = BabyDbl(c.sub("dbl3"), d2.x, d2.y);
    V<M> isz = IsZero(c.sub("isZero"), d2.x);        // dbl3.x is wired to dbl2.xout
    HZ_CHECK0(c, isz * en);
    m.en = en; m.R8x = R8x_; m.R8y = R8y_;
}
TM HZ_FN void EdDSAPoseidonVerifierTail(Comp<M>& c, const EdMid<M>& m) {
    PtV<M> mulAny = EscalarMulAny254(c.sub("mulAny"), m.hbits, m.d3);
    PtV<M> right = BabyAdd(c.sub("addRight"), m.R8x, m.R8y, mulAny.x, mulAny.y);
    PtV<M> left = EscalarMulFix253(c.sub("mulFix"), m.sbits);
    ForceEqualIfEnabled(c.sub("eqCheckX"), m.en, left.x, right.x);
    ForceEqualIfEnabled(c.sub("eqCheckY"), m.en, left.y, right.y);
}
// the second half in the pieces k_eddsa_tail evaluates side by side: the variable-base side up to `right`, the
// fixed-base multiplication (a component of its own: EscalarMulFix253 at its cursor), the two comparisons
TM HZ_FN PtV<M> EdDSAPoseidonVerifierTailAny(Comp<M>& c, const EdMid<M>& m) {
    PtV<M> mulAny = EscalarMulAny254(c.sub("mulAny"), m.hbits, m.d3);
    return BabyAdd(c.sub("addRight"), m.R8x, m.R8y, mulAny.x, mulAny.y);
}
TM HZ_FN void EdDSAPoseidonVerifierTailEq(Comp<M>& c, const EdMid<M>& m, const PtV<M>& left, const PtV<M>& right) {
    ForceEqualIfEnabled(c.sub("eqCheckX"), m.en, left.x, right.x);
    ForceEqualIfEnabled(c.sub("eqCheckY"), m.en, left.y, right.y);
}
TM HZ_FN void EdDSAPoseidonVerifier(Comp<M>&& c, const V<M>& enabled, const V<M>& Ax, const V<M>& Ay,
                                     const V<M>& S, const V<M>& R8x, const V<M>& R8y, const V<M>& Msg) {
    EdMid<M> m;
    EdDSAPoseidonVerifierHead(c, enabled, Ax, Ay, S, R8x, R8y, Msg, m);
    EdDSAPoseidonVerifierTail(c, m);
}

// ---------------------------------------------------------------------------
// pointbits.circom  (src/lib/utils-bjj.circom:46)   SURVEY.md A.4
// ---------------------------------------------------------------------------
// circomlib `sqrt()`: Tonelli-Shanks, 0 for non-residues, result normalised to <= (p-1)/2
HZ_HDN Fr fr_sqrt(const Fr& n) {
    if (fr_is_zero(n)) return n;
    Fr one = fr_one();
    // p - 1 = t 2^28, t odd.  One exponentiation gives everything: x = n^((t-1)/2), r = x n = n^((t+1)/2),
    // tt = r x = n^t, and Euler's criterion n^((p-1)/2) = tt^(2^27) (circomlib evaluates three separate powers;
    // the values are the same)
    const uint32_t e_l[8] = HZ_TONELLI_TM1H_LIMBS;      // (t-1)/2
    const uint32_t cst_l[8] = HZ_TONELLI_C_LIMBS;       // c = 5^t
    Fr x = fr_pow(n, e_l, 225);
    Fr r = fr_mul(x, n);
    Fr t = fr_mul(r, x);
    {
        Fr s = t;
        for (int k = 0; k < 27; k++) s = fr_sqr(s);
        if (!fr_eq(s, one)) return fr_zero();
    }
    int m = 28;
    Fr cst; for (int i = 0; i < 8; i++) cst.l[i] = cst_l[i];
    while (!fr_is_zero(r) && !fr_eq(t, one)) {
        Fr sq = fr_sqr(t);
        int i = 1;
        while (!fr_eq(sq, one)) { i++; sq = fr_sqr(sq); }
        Fr b = cst;
        for (int k = 0; k < m - i - 1; k++) b = fr_sqr(b);
        m = i;
        cst = fr_sqr(b);
        t = fr_mul(t, cst);
        r = fr_mul(r, b);
    }
    if (fr_is_neg(r)) r = fr_neg(r);
    return r;
}
TM struct SqrtMath;
template <> struct SqrtMath<MCompute> {
    // x = sqrt((1 - y^2) / (a - d y^2)), negated when sign == 1
    static HZ_HD VC point_x(const VC& y, uint32_t sign) {
        Fr y2 = fr_sqr(y.v);
        Fr num = fr_sub(fr_one(), y2);
        Fr den = fr_sub(fr_from_u64(HZ_BJ_A), fr_mul(fr_from_u64(HZ_BJ_D), y2));
        Fr x = fr_sqrt(fr_mul(num, fr_inv(den)));
        if (sign) x = fr_neg(x);
        return vc(x);
    }
};
template <> struct SqrtMath<MLayout> {
    static HZ_HD VL point_x(const VL&, uint32_t) { return vl(); }
};

// bits: 256 elements
TM HZ_FN PtV<M> Bits2Point_Strict(Comp<M>&& c, const BV<M>& bits) {
    BV<M> ins; ins.clear(256);
    for (int i = 0; i < 256; i++) ins.put(i, c.def(Nm("in", i), bits.get(i)));
    uint32_t out0 = c.sig(Nm("out", 0));
    AliasCheck(c.sub("aliasCheckY"), ins);
    HZ_CHECK0(c, ins.get(254));
    V<M> y = Bits2Num(c.sub("b2nY"), ins, 0, 254);
    c.alias(Nm("out", 1), y);
    PtV<M> r;
    r.y = y;
    r.x = c.set(out0, SqrtMath<M>::point_x(y, ins.bit(255)));
    BabyCheck(c.sub("babyCheck"), r.x, r.y);
    BV<M> xbits = Num2Bits(c.sub("n2bX"), 254, r.x);
    AliasCheck(c.sub("aliasCheckX"), xbits);
    V<M> sc = CompConstant(c.sub("signCalc"), fr_const_half(), xbits);
    HZ_CHECK(c, sc, ins.get(255));
    return r;
}

// ---------------------------------------------------------------------------
// sha256/*.circom   (src/hash-inputs.circom:113, src/withdraw.circom:135)
// 32-bit words of bit signals; bit k of a W32 is signal k (LSB first).
// ---------------------------------------------------------------------------
template <class M> struct W32;
template <> struct W32<MCompute> { uint32_t v; };
template <> struct W32<MLayout> { uint32_t idx[32]; };

HZ_HD W32<MCompute> w32_def(Comp<MCompute>& c, const char*, uint64_t value, int nout) {
    c.put_bits(value, nout);
    W32<MCompute> r; r.v = (uint32_t)value; return r;
}
HZ_HD void w32_alias(Comp<MCompute>&, const char*, const W32<MCompute>&) {}
HZ_HD W32<MCompute> w32_rotr(const W32<MCompute>& x, int r) { W32<MCompute> o; o.v = (x.v >> r) | (x.v << (32 - r)); return o; }
HZ_HD uint32_t w32_val(const W32<MCompute>& x) { return x.v; }
#ifndef HZ_DEVICE_ONLY
inline W32<MLayout> w32_def(Comp<MLayout>& c, const char* nm, uint64_t, int nout) {
    W32<MLayout> r;
    for (int k = 0; k < 32; k++) r.idx[k] = HZ_NOIDX;
    for (int k = 0; k < nout; k++) { c.name(Nm(nm, k), c.cur); if (k < 32) r.idx[k] = c.cur; c.cur++; }
    return r;
}
inline void w32_alias(Comp<MLayout>& c, const char* nm, const W32<MLayout>& x) {
    if (c.named) for (int k = 0; k < 32; k++) c.name(Nm(nm, k), x.idx[k]);
}
inline W32<MLayout> w32_rotr(const W32<MLayout>& x, int r) { W32<MLayout> o; for (int i = 0; i < 32; i++) o.idx[i] = x.idx[(i + r) & 31]; return o; }
inline uint32_t w32_val(const W32<MLayout>&) { return 0; }
#endif

// naming helpers (layout walk only)
HZ_HD void hz_binsum_names(Comp<MCompute>&, int, const W32<MCompute>&) {}
HZ_HD void hz_sha_in_names(Comp<MCompute>&, const W32<MCompute>*, const W32<MCompute>*) {}
// digest word j: entries out[32 j + k] = bit (31 - k) of v
HZ_HD W32<MCompute> hz_sha_out_word(Comp<MCompute>& c, int, uint32_t v) {
    c.put_bits(v, 32, true);
    W32<MCompute> r; r.v = v; return r;
}
#ifndef HZ_DEVICE_ONLY
inline void hz_binsum_names(Comp<MLayout>& c, int j, const W32<MLayout>& x) {
    if (c.named) for (int k = 0; k < 32; k++) c.name(Nm("in", j, k), x.idx[k]);
}
inline void hz_sha_in_names(Comp<MLayout>& c, const W32<MLayout>* hin, const W32<MLayout>* msg) {
    if (!c.named) return;
    for (int j = 0; j < 8; j++) for (int k = 0; k < 32; k++) c.name(Nm("hin", 32 * j + k), hin[j].idx[k]);
    for (int t = 0; t < 16; t++) for (int k = 0; k < 32; k++) c.name(Nm("inp", 32 * t + 31 - k), msg[t].idx[k]);
}
inline W32<MLayout> hz_sha_out_word(Comp<MLayout>& c, int j, uint32_t) {
    W32<MLayout> r;
    for (int k = 0; k < 32; k++) { c.name(Nm("out", 32 * j + k), c.cur); r.idx[31 - k] = c.cur; c.cur++; }
    return r;
}
#endif

TM HZ_FN W32<M> RotR(Comp<M>&& c, const W32<M>& x, int r) {
    w32_alias(c, "in", x);
    W32<M> o = w32_rotr(x, r);
    w32_alias(c, "out", o);
    return o;
}
// out[i] = in[i + r], zero (a constant entry) at the top r positions
HZ_HD W32<MCompute> ShR(Comp<MCompute>&& c, const W32<MCompute>& x, int r) {
    c.put_bits(0, r);
    W32<MCompute> o; o.v = x.v >> r; return o;
}
#ifndef HZ_DEVICE_ONLY
inline W32<MLayout> ShR(Comp<MLayout>&& c, const W32<MLayout>& x, int r) {
    w32_alias(c, "in", x);
    W32<MLayout> o;
    for (int i = 0; i < 32; i++) {
        if (i + r >= 32) { c.name(Nm("out", i), c.cur); o.idx[i] = c.cur++; }
        else { o.idx[i] = x.idx[i + r]; c.name(Nm("out", i), o.idx[i]); }
    }
    return o;
}
#endif
TM HZ_FN W32<M> Xor3(Comp<M>&& c, const W32<M>& a, const W32<M>& b, const W32<M>& cc) {
    w32_alias(c, "a", a); w32_alias(c, "b", b); w32_alias(c, "c", cc);
    W32<M> o = w32_def(c, "out", w32_val(a) ^ w32_val(b) ^ w32_val(cc), 32);
    w32_def(c, "mid", w32_val(b) & w32_val(cc), 32);
    return o;
}
TM HZ_FN W32<M> SmallSigma(Comp<M>&& c, const W32<M>& x, int ra, int rb, int rc) {
    w32_alias(c, "in", x);
    W32<M> rota = RotR(c.sub("rota"), x, ra);
    W32<M> rotb = RotR(c.sub("rotb"), x, rb);
    W32<M> shrc = ShR(c.sub("shrc"), x, rc);
    W32<M> o = Xor3(c.sub("xor3"), rota, rotb, shrc);
    w32_alias(c, "out", o);
    return o;
}
TM HZ_FN W32<M> BigSigma(Comp<M>&& c, const W32<M>& x, int ra, int rb, int rc) {
    w32_alias(c, "in", x);
    W32<M> rota = RotR(c.sub("rota"), x, ra);
    W32<M> rotb = RotR(c.sub("rotb"), x, rb);
    W32<M> rotc = RotR(c.sub("rotc"), x, rc);
    W32<M> o = Xor3(c.sub("xor3"), rota, rotb, rotc);
    w32_alias(c, "out", o);
    return o;
}
TM HZ_FN W32<M> Ch_t(Comp<M>&& c, const W32<M>& a, const W32<M>& b, const W32<M>& cc) {
    w32_alias(c, "a", a); w32_alias(c, "b", b); w32_alias(c, "c", cc);
    uint32_t av = w32_val(a);
    return w32_def(c, "out", (av & w32_val(b)) | (~av & w32_val(cc)), 32);
}
TM HZ_FN W32<M> Maj_t(Comp<M>&& c, const W32<M>& a, const W32<M>& b, const W32<M>& cc) {
    w32_alias(c, "a", a); w32_alias(c, "b", b); w32_alias(c, "c", cc);
    uint32_t av = w32_val(a), bv = w32_val(b), cv = w32_val(cc);
    W32<M> o = w32_def(c, "out", (av & bv) ^ (av & cv) ^ (bv & cv), 32);
    w32_def(c, "mid", bv & cv, 32);
    return o;
}
HZ_HD int hz_binsum_nout(int nops) { return nops == 2 ? 33 : (nops == 4 ? 34 : 35); }
TM HZ_FN W32<M> BinSum(Comp<M>&& c, const W32<M>* ops, int nops) {
    uint64_t sum = 0;
    for (int j = 0; j < nops; j++) { sum += w32_val(ops[j]); }
    for (int j = 0; j < nops; j++) hz_binsum_names(c, j, ops[j]);
    return w32_def(c, "out", sum, hz_binsum_nout(nops));
}
#define HZ_SZ_SIGMAPLUS 175
#define HZ_SZ_T1 131
#define HZ_SZ_T2 161
#define HZ_SZ_SHA_BLOCK (256 + 48 * HZ_SZ_SIGMAPLUS + 64 * 32 + 64 * HZ_SZ_T1 + 64 * HZ_SZ_T2 + 64 * 33 + 64 * 33 + 8 * 33)
TM HZ_FN W32<M> SigmaPlus(Comp<M>&& c, const W32<M>& in2, const W32<M>& in7, const W32<M>& in15, const W32<M>& in16) {
    w32_alias(c, "in2", in2); w32_alias(c, "in7", in7); w32_alias(c, "in15", in15); w32_alias(c, "in16", in16);
    W32<M> ops[4];
    ops[0] = SmallSigma(c.sub("sigma1"), in2, 17, 19, 10);
    ops[1] = in7;
    ops[2] = SmallSigma(c.sub("sigma0"), in15, 7, 18, 3);
    ops[3] = in16;
    W32<M> o = BinSum<M>(c.sub("sum"), ops, 4);
    w32_alias(c, "out", o);
    return o;
}
TM HZ_FN W32<M> T1(Comp<M>&& c, const W32<M>& h, const W32<M>& e, const W32<M>& f, const W32<M>& g,
                   const W32<M>& k, const W32<M>& w) {
    w32_alias(c, "h", h); w32_alias(c, "e", e); w32_alias(c, "f", f); w32_alias(c, "g", g);
    w32_alias(c, "k", k); w32_alias(c, "w", w);
    Comp<M> ch_c = c.sub_reserve("ch", 32);
    W32<M> ops[5];
    ops[0] = h;
    ops[1] = BigSigma(c.sub("bigsigma1"), e, 6, 11, 25);
    ops[2] = Ch_t(HZ_MV(ch_c), e, f, g);
    ops[3] = k;
    ops[4] = w;
    W32<M> o = BinSum<M>(c.sub("sum"), ops, 5);
    w32_alias(c, "out", o);
    return o;
}
TM HZ_FN W32<M> T2(Comp<M>&& c, const W32<M>& a, const W32<M>& b, const W32<M>& cc) {
    w32_alias(c, "a", a); w32_alias(c, "b", b); w32_alias(c, "c", cc);
    W32<M> ops[2];
    ops[0] = BigSigma(c.sub("bigsigma0"), a, 2, 13, 22);
    ops[1] = Maj_t(c.sub("maj"), a, b, cc);
    W32<M> o = BinSum<M>(c.sub("sum"), ops, 2);
    w32_alias(c, "out", o);
    return o;
}

HZ_HD uint32_t hz_sha_k(int i) {
    const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return K[i];
}
HZ_HD uint32_t hz_sha_h(int i) {
    const uint32_t H[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    return H[i];
}
HZ_HD uint32_t hz_rotr32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
// plain compression on words (the `sha256compression()` helper function upstream uses for `out <--`)
HZ_HD void hz_sha256_compress(uint32_t* st /*8, in/out*/, const uint32_t* blk /*16*/) {
    uint32_t w[64];
    for (int t = 0; t < 16; t++) w[t] = blk[t];
    for (int t = 16; t < 64; t++) {
        uint32_t s0 = hz_rotr32(w[t - 15], 7) ^ hz_rotr32(w[t - 15], 18) ^ (w[t - 15] >> 3);
        uint32_t s1 = hz_rotr32(w[t - 2], 17) ^ hz_rotr32(w[t - 2], 19) ^ (w[t - 2] >> 10);
        w[t] = w[t - 16] + s0 + w[t - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int t = 0; t < 64; t++) {
        uint32_t S1 = hz_rotr32(e, 6) ^ hz_rotr32(e, 11) ^ hz_rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + hz_sha_k(t) + w[t];
        uint32_t S0 = hz_rotr32(a, 2) ^ hz_rotr32(a, 13) ^ hz_rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// hin[j]: chaining word j, LSB-first signals; msg[t]: message word t (bit k = inp[32 t + 31 - k]);
// outw[j]: digest word j as LSB-first view of out[32 j + 31 - k].
// The component is cut into UNITS with fixed offsets from its base so that the device can give
// each unit to its own thread (k_sha_blocks): the digest words, 48 sigmaPlus, 64 rounds
// (ct_k, t1, t2, sume, suma) and 8 final sums.
#define HZ_SHA_OFF_SP 256
#define HZ_SHA_OFF_CK (HZ_SHA_OFF_SP + 48 * HZ_SZ_SIGMAPLUS)
#define HZ_SHA_OFF_T1 (HZ_SHA_OFF_CK + 64 * 32)
#define HZ_SHA_OFF_T2 (HZ_SHA_OFF_T1 + 64 * HZ_SZ_T1)
#define HZ_SHA_OFF_SA (HZ_SHA_OFF_T2 + 64 * HZ_SZ_T2)
#define HZ_SHA_OFF_SE (HZ_SHA_OFF_SA + 64 * 33)
#define HZ_SHA_OFF_FS (HZ_SHA_OFF_SE + 64 * 33)
// message-schedule word t (16 <= t < 64); c = the sha256compression component, base = its start
TM HZ_FN W32<M> Sha256UnitSigmaPlus(Comp<M>& c, uint32_t base, int t, const W32<M>& w2, const W32<M>& w7,
                                     const W32<M>& w15, const W32<M>& w16) {
    c.cur = base + HZ_SHA_OFF_SP + (uint32_t)(t - 16) * HZ_SZ_SIGMAPLUS;
    return SigmaPlus(c.sub(Nm("sigmaPlus", t - 16)), w2, w7, w15, w16);
}
// round t on the working variables v[0..7] = (a, b, c, d, e, f, g, h): returns the new a and e
TM HZ_FN void Sha256UnitRound(Comp<M>& c, uint32_t base, int t, const W32<M>* v, const W32<M>& wt, W32<M>& na, W32<M>& ne) {
    c.cur = base + HZ_SHA_OFF_CK + (uint32_t)t * 32;
    W32<M> kt;
    { Comp<M> kc = c.sub(Nm("ct_k", t)); kt = w32_def(kc, "out", hz_sha_k(t), 32); }
    c.cur = base + HZ_SHA_OFF_T1 + (uint32_t)t * HZ_SZ_T1;
    W32<M> t1 = T1(c.sub(Nm("t1", t)), v[7], v[4], v[5], v[6], kt, wt);
    c.cur = base + HZ_SHA_OFF_T2 + (uint32_t)t * HZ_SZ_T2;
    W32<M> t2 = T2(c.sub(Nm("t2", t)), v[0], v[1], v[2]);
    W32<M> ops[2];
    c.cur = base + HZ_SHA_OFF_SE + (uint32_t)t * 33;
    ops[0] = v[3]; ops[1] = t1;
    ne = BinSum<M>(c.sub(Nm("sume", t)), ops, 2);
    c.cur = base + HZ_SHA_OFF_SA + (uint32_t)t * 33;
    ops[0] = t1; ops[1] = t2;
    na = BinSum<M>(c.sub(Nm("suma", t)), ops, 2);
}
TM HZ_FN W32<M> Sha256UnitFinal(Comp<M>& c, uint32_t base, int j, const W32<M>& hin, const W32<M>& fin) {
    c.cur = base + HZ_SHA_OFF_FS + (uint32_t)j * 33;
    W32<M> ops[2];
    ops[0] = hin; ops[1] = fin;
    return BinSum<M>(c.sub(Nm("fsum", j)), ops, 2);
}
TM HZ_FN void Sha256compression(Comp<M>&& c, const W32<M>* hin, const W32<M>* msg, W32<M>* outw) {
    hz_sha_in_names(c, hin, msg);
    uint32_t st[8], blk[16];
    for (int j = 0; j < 8; j++) st[j] = w32_val(hin[j]);
    for (int t = 0; t < 16; t++) blk[t] = w32_val(msg[t]);
    hz_sha256_compress(st, blk);
    const uint32_t base = c.cur;
    for (int j = 0; j < 8; j++) outw[j] = hz_sha_out_word(c, j, st[j]);
    W32<M> w[64];
    for (int t = 0; t < 64; t++) w[t] = t < 16 ? msg[t] : Sha256UnitSigmaPlus(c, base, t, w[t - 2], w[t - 7], w[t - 15], w[t - 16]);
    W32<M> v[8];
    for (int j = 0; j < 8; j++) v[j] = hin[j];
    for (int t = 0; t < 64; t++) {
        W32<M> na, ne;
        Sha256UnitRound(c, base, t, v, w[t], na, ne);
        v[7] = v[6]; v[6] = v[5]; v[5] = v[4]; v[4] = ne; v[3] = v[2]; v[2] = v[1]; v[1] = v[0]; v[0] = na;
    }
    for (int j = 0; j < 8; j++) {
        W32<M> fs = Sha256UnitFinal(c, base, j, hin[j], v[j]);
        if (c.sanity() && w32_val(fs) != st[j])
            c.check(Ops<M>::K(w32_val(fs)), Ops<M>::K(st[j]), __LINE__);
    }
    c.cur = base + HZ_SZ_SHA_BLOCK;
}
