// The ladders of the EdDSA verifier, off the lane groups (hints.cuh has the plan; launched between k_eddsa and k_eddsa_tail).
//
// k_ed_ladder   The scalar multiplications of a signature are three chains of point operations, each waiting for
//               the point before it: 253 doublings of 8A, the additions into the accumulators of the two
//               EscalarMulAny segments (252), the 83 additions of the fixed-base windows.  None of the points is a
//               witness entry, so the walk stays in the Montgomery domain with lazily reduced values
//               (fr_montgen.inc: 0 <= x~ < 2p) and uses the cheapest form of each map (below).  A chain is walked
//               by FOUR lanes: the products of a point operation come in stages of four independent ones, one per
//               lane, and the lanes trade results with shuffles -- a step costs two product latencies on the curve
//               (four off it) where one thread took eight to sixteen.  A block is three warps over the same 8
//               signatures (doublings / segment adders / fixed-base adders); the doubling warp hands D_t to the
//               adder warp through shared memory, one step behind.  The points of every step are left in memory for
// k_ed_conv     which turns them into what the templates' divisions need (slope numerators / denominators, captured
//               points), one thread per element: products that the chain no longer waits for.
// k_ed_hints    Montgomery's trick over the 588 denominators and the 81 captured W of a signature, a warp per
//               signature (21 elements per lane, the chunk totals combined with shuffle scans, one inversion).
//               Ratios do not depend on the domain: num~ / den~ = num / den, so the hints come out canonical.
#include "hzw_dev.cuh"
#include "fm.cuh"

#define HZ_LAD_SIGS 8
// ---- on the curve: twisted Edwards, extended coordinates -----------------------------------------------------------
// What the templates chain together are Montgomery-form point operations (MontgomeryDouble / MontgomeryAdd, 16 field
// products each when walked projectively).  On the curve they are the group law of BabyJubJub, and the group law is
// cheapest on the twisted Edwards form with a = -1: a x^2 + y^2 = 1 + d x^2 y^2 with a = 168700 becomes
// -x'^2 + y^2 = 1 + d' x'^2 y^2 under x' = s x, s^2 = -a (-a is a square in Fr).  In extended coordinates
// (X : Y : Z : T), T = X Y / Z, a doubling is 4 squarings then 4 products (dbl-2008-hwcd) and an addition of two
// different points 4 products then 4 products with no curve constant in them (add-2008-hwcd-4) -- two stages of four
// independent products, which is what the four lanes of a chain are for.  Nothing the templates need is lost:
// u = (Z + Y) / (Z - Y) and v = s (Z + Y) Z / ((Z - Y) X) give every slope denominator and every captured point back
// as a ratio of products of the stored coordinates (k_ed_conv).  The dedicated addition degenerates exactly where the
// templates' own slope does (equal or opposite points: a zero denominator, which k_ed_hints flags).
// ---- off the curve: the rational maps themselves ---------------------------------------------------------------------
// The key of an L1 slot is (getAx(0, 0), ay of the state) (src/rollup-tx.circom:447-476): not a curve point, and every
// batch has maxL1Tx of those.  There the templates' formulas are just rational maps of (u, v) and are walked as such
// on projective triples (U : V : W): 16 products per step in four stages of four.  A block whose 8 signatures include
// such a key walks all 8 this way (the maps are the group law on the curve too); the fixed-base chain does not depend
// on the key and is always walked on the Edwards form.  rec.pad[0] tells k_ed_conv which form the points are in.
__device__ __forceinline__ Fr fm_const_s() {          // s = sqrt(-168700), canonical
    Fr r; r.l[0] = 0xd8b2d768u; r.l[1] = 0x6c7be2c1u; r.l[2] = 0x321ea117u; r.l[3] = 0x74a5f6cau;
    r.l[4] = 0x84e61177u; r.l[5] = 0x2c7ac0ceu; r.l[6] = 0x05d6e54bu; r.l[7] = 0x0e0ff350u; return r;
}
__device__ __forceinline__ void hz_lad_put(Fr* p, const Fr& v) {
    *reinterpret_cast<uint4*>(p) = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    *(reinterpret_cast<uint4*>(p) + 1) = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// one stage: this lane's product, then everybody's (lane j's product comes back as v[j]).  The four lanes trade through
// shared memory (one __syncwarp per stage; the stages of a step alternate between two buffers, so a lane that is ahead
// never writes what a lane behind still reads).  Shuffles would do, but in code the compiler cannot prove convergent
// (the role of a warp is a function of threadIdx) every one of them is wrapped in its own synchronisation.
typedef uint32_t HzQuadBuf[8][32];
// the product as ONE copy of code that every stage of every role calls: inlined at its 16 call sites the walk is 175 KB
// of straight-line code and its top stall is the instruction fetch
__device__ __noinline__ Fr fm_mul_shared(const Fr a, const Fr b) { return fm_mul(a, b); }
__device__ __forceinline__ void hz_quad_stage(HzQuadBuf& xb, const Fr& a, const Fr& b, Fr& v0, Fr& v1, Fr& v2, Fr& v3) {
    const Fr mine = fm_mul_shared(a, b);
    const int lane = (int)(threadIdx.x & 31u), base = lane & 28;
#pragma unroll
    for (int k = 0; k < 8; k++) xb[k][lane] = mine.l[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) { v0.l[k] = xb[k][base]; v1.l[k] = xb[k][base + 1]; v2.l[k] = xb[k][base + 2]; v3.l[k] = xb[k][base + 3]; }
}
// lane q's pick of four values, as selects (written as a ternary chain the compiler branches per limb, divergently)
__device__ __forceinline__ uint32_t hz_selp(uint32_t x, uint32_t y, int c) {
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(x), "r"(y), "r"(c));
    return r;
}
__device__ __forceinline__ Fr hz_sel4(int q, const Fr& a, const Fr& b, const Fr& c, const Fr& d) {
    Fr r;
    const int lo = q & 1, hi = q & 2;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = hz_selp(hz_selp(d.l[i], c.l[i], lo), hz_selp(b.l[i], a.l[i], lo), hi);
    return r;
}
// x~ * 168698 (the Montgomery-curve coefficient is small: no full product; fr_mul_small takes the lazy range)
__device__ __forceinline__ Fr fm_mulA(const Fr& x) { return fr_mul_small(x, (uint32_t)HZ_MT_A); }

struct FmEd { Fr X, Y, Z, T; };
struct FmPt { Fr U, V, W; };
// second half of a twisted-Edwards operation: (E F : G H : F G : E H), one product per lane
__device__ __forceinline__ FmEd hz_ed_finish(HzQuadBuf* xb, int q, const Fr& E, const Fr& F, const Fr& G, const Fr& H) {
    FmEd r;
    hz_quad_stage(xb[1], hz_sel4(q, E, G, F, E), hz_sel4(q, F, H, G, H), r.X, r.Y, r.Z, r.T);
    return r;
}
__device__ __forceinline__ FmEd hz_ed_dbl(HzQuadBuf* xb, int q, const FmEd& p) {
    Fr A, B, ZZ, tt;
    const Fr in = hz_sel4(q, p.X, p.Y, p.Z, fm_add(p.X, p.Y));
    hz_quad_stage(xb[0], in, in, A, B, ZZ, tt);
    const Fr E = fm_sub(fm_sub(tt, A), B), G = fm_sub(B, A);           // D + B with D = a A = -A
    const Fr F = fm_sub(G, fm_add(ZZ, ZZ)), H = fm_sub(fm_sub(fr_zero(), A), B);
    return hz_ed_finish(xb, q, E, F, G, H);
}
// p + (x2, y2, z2, t2) for two different points, the second given as (Y2 + X2, Y2 - X2, 2 T2, 2 Z2)
__device__ __forceinline__ FmEd hz_ed_add(HzQuadBuf* xb, int q, const FmEd& p, const Fr& c0, const Fr& c1, const Fr& c2, const Fr& c3) {
    Fr A, B, C, D;
    hz_quad_stage(xb[0], hz_sel4(q, fm_sub(p.Y, p.X), fm_add(p.Y, p.X), p.Z, p.T), hz_sel4(q, c0, c1, c2, c3), A, B, C, D);
    return hz_ed_finish(xb, q, fm_add(D, C), fm_sub(B, A), fm_add(B, A), fm_sub(D, C));
}
// second half of both Montgomery-form maps (hints.cuh: mt_tail): from the slope N / Dn and the cross products
__device__ __forceinline__ FmPt hz_mt_finish(HzQuadBuf* xb, int q, const Fr& N, const Fr& Dn, const Fr& S, const Fr& NN, const Fr& W12, const Fr& X1, const Fr& X2, const Fr& Y1) {
    const Fr T = fm_add(fm_add(fm_mulA(W12), X1), X2);
    Fr NNW, TS, DS, X1S;
    hz_quad_stage(xb[0], hz_sel4(q, NN, T, Dn, X1), hz_sel4(q, W12, S, S, S), NNW, TS, DS, X1S);
    const Fr G = fm_sub(NNW, TS);
    Fr U3, Pn, Qn, W3;
    hz_quad_stage(xb[1], hz_sel4(q, G, N, Y1, DS), hz_sel4(q, Dn, fm_sub(X1S, G), DS, W12), U3, Pn, Qn, W3);
    FmPt r; r.U = U3; r.V = fm_sub(Pn, Qn); r.W = W3;
    return r;
}
__device__ __forceinline__ FmPt hz_mt_dbl(HzQuadBuf* xb, int q, const FmPt& p) {
    Fr UU, UW, WW, VW;
    hz_quad_stage(xb[0], hz_sel4(q, p.U, p.U, p.W, p.V), hz_sel4(q, p.U, p.W, p.W, p.W), UU, UW, WW, VW);
    const Fr AUW = fm_mulA(UW);
    const Fr N = fm_add(fm_add(fm_add(fm_add(UU, UU), UU), fm_add(AUW, AUW)), WW), Dn = fm_add(VW, VW);
    Fr S, NN, x2, x3;
    hz_quad_stage(xb[1], hz_sel4(q, Dn, N, Dn, N), hz_sel4(q, Dn, N, Dn, N), S, NN, x2, x3);
    return hz_mt_finish(xb, q, N, Dn, S, NN, WW, UW, UW, VW);
}
// MontgomeryAdd(in1 = p1, in2 = p2)
__device__ __forceinline__ FmPt hz_mt_add(HzQuadBuf* xb, int q, const FmPt& p1, const FmPt& p2) {
    Fr X1, X2, Y1, Y2;
    hz_quad_stage(xb[0], hz_sel4(q, p1.U, p2.U, p1.V, p2.V), hz_sel4(q, p2.W, p1.W, p2.W, p1.W), X1, X2, Y1, Y2);
    const Fr N = fm_sub(Y2, Y1), Dn = fm_sub(X2, X1);
    Fr S, NN, W12, x3;
    hz_quad_stage(xb[1], hz_sel4(q, Dn, N, p1.W, N), hz_sel4(q, Dn, N, p2.W, N), S, NN, W12, x3);
    return hz_mt_finish(xb, q, N, Dn, S, NN, W12, X1, X2, Y1);
}

// One-time: the window entries of EscalarMulFix (affine Montgomery (u, v), hint_tab) in the form the fixed-base chain
// adds: x = u / v, y = (u - 1) / (u + 1), x' = s x  ->  (y + x', y - x', 2 x' y) in the Montgomery domain, 4 Fr per
// point; behind the table the starting accumulator of the fixed-base adders as a full point (X', Y, 1, T).
__global__ void k_ed_tab(const Fr* tab, Fr* edtab, int npoints, int init_point) {
    const int p = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (p >= npoints) return;
    const Fr u = tab[2 * p], v = tab[2 * p + 1];
    const Fr x = fr_mul_chain(fr_mul_chain(u, fr_inv_euclid(v)), fm_const_s());
    const Fr y = fr_mul_chain(fr_sub(u, fr_one()), fr_inv_euclid(fr_add(u, fr_one())));
    const Fr t = fr_mul_chain(x, y);
    edtab[4 * p] = fm_from(fr_add(y, x)); edtab[4 * p + 1] = fm_from(fr_sub(y, x)); edtab[4 * p + 2] = fm_from(fr_add(t, t));
    edtab[4 * p + 3] = fr_zero();
    if (p == init_point) {
        Fr* o = edtab + 4 * (size_t)npoints;
        o[0] = fm_from(x); o[1] = fm_from(y); o[2] = fm_one(); o[3] = fm_from(t);
    }
}

// block = 3 warps over the signatures [8 b, 8 b + 8), four lanes per signature: warp 0 the doubling chain, warp 1 the
// adders of the two EscalarMulAny segments, warp 2 the fixed-base adders.  Signature g leaves its points in
// pts + g * HZ_PTS (hints.cuh), 3 Fr each: (X, Y, Z) on the curve, (U, V, W) off it.
__global__ void __launch_bounds__(96) k_ed_ladder(Fr* scratch, const Fr* edtab, Fr* pts_all, int ns) {
    __shared__ uint32_t dsh[2][4][8][HZ_LAD_SIGS];
    __shared__ HzQuadBuf xball[3][2];
    const int role = (int)(threadIdx.x >> 5), c = (int)((threadIdx.x & 31) >> 2), q = (int)(threadIdx.x & 3);
    HzQuadBuf* xb = xball[role];
    int g = (int)blockIdx.x * HZ_LAD_SIGS + c;
    const bool real = g < ns;
    if (!real) g = ns - 1;                                 // padding lanes redo the last signature and store nothing
    Fr* scr = scratch + (size_t)g * HZ_HINT_SCRATCH;
    Fr* pts = pts_all + (size_t)g * HZ_PTS;
    const HzFillRec rec = *reinterpret_cast<const HzFillRec*>(scr + HZ_SCR_REC);
    // 8A with EscalarMulAny's zero-point substitution (escalarmulany.circom): p.x == 0 -> BASE8
    Fr px = scr[HZ_SCR_MID], py = scr[HZ_SCR_MID + 1];
    if (fr_is_zero(px)) { px = fr_base8_x(); py = fr_base8_y(); }
    // is it a curve point?  168700 x^2 + y^2 = 1 + 168696 x^2 y^2
    bool on;
    {
        const Fr x2 = fr_mul_chain(px, px), y2 = fr_mul_chain(py, py);
        const Fr lhs = fr_add(fr_mul_small(x2, 168700u), y2), rhs = fr_add(fr_one(), fr_mul_small(fr_mul_chain(x2, y2), 168696u));
        on = fr_eq(lhs, rhs);
    }
    const bool ed = __syncthreads_or(on ? 0 : 1) == 0;      // the whole block walks one form
    if (role == 0 && q == 0 && real) reinterpret_cast<HzFillRec*>(scr + HZ_SCR_REC)->pad[0] = ed ? 1u : 0u;
    const Fr xm = fm_from(px), ym = fm_from(py), one = fm_one();
    FmEd D, acc;                                            // off the curve: (U, V, W) in X, Y, Z
    if (ed) {
        D.X = fm_mul(xm, fm_from(fm_const_s())); D.Y = ym; D.Z = one; D.T = fm_mul(D.X, ym);
    } else {
        // Edwards2Montgomery: u = (1+y)/(1-y), v = u/x  ->  ((1+y) x : (1+y) : (1-y) x)
        const Fr y1 = fm_add(one, ym);
        D.X = fm_mul(y1, xm); D.Y = y1; D.Z = fm_mul(fm_sub(one, ym), xm); D.T = fr_zero();
    }
    acc = D;
    Fr w0 = fr_zero(), w1 = fr_zero(), w2 = fr_zero();      // next window entry of the fixed-base adders
    if (role == 2) {
        const Fr* o = edtab + 4 * (size_t)(HZ_EDTAB_LEN / 2);
        acc.X = o[0]; acc.Y = o[1]; acc.Z = o[2]; acc.T = o[3];
        const int sel = (int)(rec.s[0] & 7u);
        w0 = edtab[4 * sel]; w1 = edtab[4 * sel + 1]; w2 = edtab[4 * sel + 2];
    }
    for (int t = 0; t <= 253; t++) {
        if (role == 0) {
            if (real && q < 3) hz_lad_put(pts + HZ_PT_D + 3 * t + q, hz_sel4(q, D.X, D.Y, D.Z, D.Z));
            if (t <= 252) {
                if (ed) D = hz_ed_dbl(xb, q, D);
                else { FmPt p; p.U = D.X; p.V = D.Y; p.W = D.Z; p = hz_mt_dbl(xb, q, p); D.X = p.U; D.Y = p.V; D.Z = p.W; }
                const Fr mine = hz_sel4(q, D.X, D.Y, D.Z, D.T);
                uint32_t (*buf)[8][HZ_LAD_SIGS] = dsh[(t + 1) & 1];
#pragma unroll
                for (int k = 0; k < 8; k++) buf[q][k][c] = mine.l[k];
            }
        } else if (role == 1) {
            if (t >= 1) {
                const uint32_t (*buf)[8][HZ_LAD_SIGS] = dsh[t & 1];
#pragma unroll
                for (int k = 0; k < 8; k++) { D.X.l[k] = buf[0][k][c]; D.Y.l[k] = buf[1][k][c]; D.Z.l[k] = buf[2][k][c]; D.T.l[k] = buf[3][k][c]; }
                if (real && q < 3) hz_lad_put(pts + HZ_PT_A + 3 * (t - 1) + q, hz_sel4(q, acc.X, acc.Y, acc.Z, acc.Z));   // the accumulator BEFORE step t
                if (t == 148) acc = D;                                          // segment 1 starts from D_148
                else {
                    FmEd r;
                    if (ed) r = hz_ed_add(xb, q, acc, fm_add(D.Y, D.X), fm_sub(D.Y, D.X), fm_add(D.T, D.T), fm_add(D.Z, D.Z));
                    else {
                        FmPt p1, p2; p1.U = D.X; p1.V = D.Y; p1.W = D.Z; p2.U = acc.X; p2.V = acc.Y; p2.W = acc.Z;
                        const FmPt p = hz_mt_add(xb, q, p1, p2);
                        r.X = p.U; r.Y = p.V; r.Z = p.W; r.T = fr_zero();
                    }
                    if ((rec.e[t >> 5] >> (t & 31)) & 1u) acc = r;
                }
            }
        } else {
            if (t < 83) {
                if (real && q < 3) hz_lad_put(pts + HZ_PT_F + 3 * t + q, hz_sel4(q, acc.X, acc.Y, acc.Z, acc.Z));
                const Fr c0 = w0, c1 = w1, c2 = w2;
                const int tt = t + 1 < 83 ? t + 1 : 0;
                const int sel = (int)(((rec.s[(3 * tt) >> 5] >> ((3 * tt) & 31)) & 1u) | (((rec.s[(3 * tt + 1) >> 5] >> ((3 * tt + 1) & 31)) & 1u) << 1) |
                                      (((rec.s[(3 * tt + 2) >> 5] >> ((3 * tt + 2) & 31)) & 1u) << 2));
                const Fr* e = edtab + 4 * (size_t)(tt * 8 + sel);
                w0 = e[0]; w1 = e[1]; w2 = e[2];
                // the window point is affine (Z2 = 1): its 2 Z2 is the domain's 2
                acc = hz_ed_add(xb, q, acc, c0, c1, c2, fm_add(one, one));
            }
        }
        __syncthreads();
    }
    if (real && role == 1 && q < 3) hz_lad_put(pts + HZ_PT_FIN + q, hz_sel4(q, acc.X, acc.Y, acc.Z, acc.Z));
    if (real && role == 2 && q < 3) hz_lad_put(pts + HZ_PT_FIN + 3 + q, hz_sel4(q, acc.X, acc.Y, acc.Z, acc.Z));
}

// ---- off the chain: every slope denominator and captured point as products of the stored coordinates ------------
// element e of signature g (one thread each): what k_ed_hints inverts together.  On the curve (rec.pad[0]), with
// u = (Z + Y) / (Z - Y) and v = s (Z + Y) Z / ((Z - Y) X):
//   doubling of D_t                  1 / (2 v)       = X (Z - Y) / (2 s Z (Z + Y))
//   adder, in1 = D_t, in2 = acc      1 / (u2 - u1)   = (Za - Ya)(Zd - Yd) / (2 (Ya Zd - Yd Za))
//   captured point                   (U : V : W)     = ((Z + Y) X : s (Z + Y) Z : (Z - Y) X)
// off it the stored triples are (U : V : W) themselves: W / (2 V), Wd Wa / (Ua Wd - Ud Wa), the point as it is.
// The fixed-base chain is on the curve either way:
//   fixed-base adder, in2 = window   1 / (u2 - u1)   = (Z - Y) / (u2 (Z - Y) - (Z + Y))
__global__ void __launch_bounds__(128) k_ed_conv(Fr* scratch, const Fr* tab, const Fr* pts_all, int ns) {
    const int gid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int g = gid / HZ_LAD_TOT, e = gid % HZ_LAD_TOT;
    if (g >= ns) return;
    Fr* scr = scratch + (size_t)g * HZ_HINT_SCRATCH;
    Fr* num = scr + HZ_SCR_ARR; Fr* den = scr + 2 * HZ_SCR_ARR; Fr* vv = scr + HZ_SCR_VV;
    const Fr* pts = pts_all + (size_t)g * HZ_PTS;
    const HzFillRec* rec = reinterpret_cast<const HzFillRec*>(scr + HZ_SCR_REC);
    const bool ed = rec->pad[0] != 0;
    const Fr sm = fm_from(fm_const_s());
    if (e < 253) {
        const Fr* P = pts + HZ_PT_D + 3 * e;
        const Fr X = P[0], Y = P[1], Z = P[2];
        if (ed) {
            hz_lad_put(&num[e], fm_mul(X, fm_sub(Z, Y)));
            hz_lad_put(&den[e], fm_mul(fm_mul(fm_add(sm, sm), Z), fm_add(Z, Y)));
        } else { hz_lad_put(&num[e], Z); hz_lad_put(&den[e], fm_add(Y, Y)); }
    } else if (e < 505) {
        const int t = e < 400 ? e - 253 + 1 : e - 400 + 149;
        const Fr* Pd = pts + HZ_PT_D + 3 * t; const Fr* Pa = pts + HZ_PT_A + 3 * (t - 1);
        if (ed) {
            const Fr Yd = Pd[1], Zd = Pd[2], Ya = Pa[1], Za = Pa[2];
            const Fr h = fm_sub(fm_mul(Ya, Zd), fm_mul(Yd, Za));
            hz_lad_put(&num[e], fm_mul(fm_sub(Za, Ya), fm_sub(Zd, Yd))); hz_lad_put(&den[e], fm_add(h, h));
        } else {
            const Fr Ud = Pd[0], Wd = Pd[2], Ua = Pa[0], Wa = Pa[2];
            hz_lad_put(&num[e], fm_mul(Wd, Wa)); hz_lad_put(&den[e], fm_sub(fm_mul(Ua, Wd), fm_mul(Ud, Wa)));
        }
    } else if (e < HZ_LAD_N) {
        const int t = e - 505;
        const Fr* Pf = pts + HZ_PT_F + 3 * t;
        const Fr Y = Pf[1], Z = Pf[2];
        const int sel = (int)(((rec->s[(3 * t) >> 5] >> ((3 * t) & 31)) & 1u) | (((rec->s[(3 * t + 1) >> 5] >> ((3 * t + 1) & 31)) & 1u) << 1) |
                              (((rec->s[(3 * t + 2) >> 5] >> ((3 * t + 2) & 31)) & 1u) << 2));
        const Fr u2 = fm_from(tab[t * 16 + 2 * sel]);
        const Fr zy = fm_sub(Z, Y);
        hz_lad_put(&num[e], zy); hz_lad_put(&den[e], fm_sub(fm_mul(u2, zy), fm_add(Z, Y)));
    } else {
        const int ex = e - HZ_LAD_N;
        const Fr* P;
        bool on = ed;
        if (ex < 19) P = pts + HZ_PT_D + 3 * (8 * ex);
        else if (ex < 38) P = pts + HZ_PT_A + 3 * (8 * (ex - 19));                   // accumulator before step 8 k + 1
        else if (ex < 52) P = pts + HZ_PT_D + 3 * (148 + 8 * (ex - 38));
        else if (ex < 66) P = pts + HZ_PT_A + 3 * (148 + 8 * (ex - 52));             // before step 149 + 8 k
        else if (ex < 77) { P = pts + HZ_PT_F + 3 * (8 * (ex - 66)); on = true; }
        else if (ex == 77) P = pts + HZ_PT_D + 3 * 147;
        else if (ex == 78) P = pts + HZ_PT_A + 3 * 147;                               // segment 0's accumulator when step 148 takes over
        else if (ex == 79) P = pts + HZ_PT_FIN;
        else { P = pts + HZ_PT_FIN + 3; on = true; }
        const Fr X = P[0], Y = P[1], Z = P[2];
        if (on) {
            const Fr zpy = fm_add(Z, Y);
            hz_lad_put(&num[HZ_LAD_N + ex], fm_mul(zpy, X));
            hz_lad_put(&vv[ex], fm_mul(fm_mul(sm, zpy), Z));
            hz_lad_put(&den[HZ_LAD_N + ex], fm_mul(fm_sub(Z, Y), X));
        } else { hz_lad_put(&num[HZ_LAD_N + ex], X); hz_lad_put(&vv[ex], Y); hz_lad_put(&den[HZ_LAD_N + ex], Z); }
    }
}

// ---- one inversion per signature ------------------------------------------------------------------------------
// hint[pos(k)] = num[k] / den[k] for the 588 slope denominators, aff[2 ex], aff[2 ex + 1] = U / W, V / W for the 81
// captured points; rec.ok = 0 when a denominator is zero (a degenerate input: the templates then divide on the spot).
#define HZ_HINT_PER 21                      /* 32 lanes x 21 >= 669 */
__device__ __forceinline__ Fr hz_canon2p(const Fr& x) {          // x < 2p -> x mod p
    Fr r = x;
    if (fr_geq_p(r)) fr_sub_p(r);
    return r;
}
__device__ __forceinline__ Fr hz_shfl_up_fr(const Fr& v, int d) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
__device__ __forceinline__ Fr hz_shfl_down_fr(const Fr& v, int d) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}
__global__ void __launch_bounds__(128) k_ed_hints(Fr* scratch, int ns) {
    const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31);
    if (g >= ns) return;                                   // whole warps
    Fr* scr = scratch + (size_t)g * HZ_HINT_SCRATCH;
    Fr* hint = scr; Fr* num = scr + HZ_SCR_ARR; Fr* den = scr + 2 * HZ_SCR_ARR; Fr* pre = scr + 3 * HZ_SCR_ARR;
    Fr* vv = scr + HZ_SCR_VV; Fr* aff = scr + HZ_SCR_AFF;
    // lane l takes the elements l, l + 32, l + 64, ...: Montgomery's trick does not care which elements share a chunk,
    // and this way the 32 lanes of a load touch 32 consecutive field elements
    bool ok = true;
    Fr tot = fr_one();
    for (int k = lane; k < HZ_LAD_TOT; k += 32) {
        const Fr dk = hz_canon2p(den[k]);
        den[k] = dk;
        if (fr_is_zero(dk)) ok = false;
        pre[k] = tot;
        tot = fr_mul_chain(tot, fr_is_zero(dk) ? fr_one() : dk);
    }
    ok = __all_sync(0xffffffffu, ok ? 1 : 0) != 0;
    // products of the chunk totals of the lanes below and above this one
    Fr below = tot;                                        // inclusive scan first
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const Fr o = hz_shfl_up_fr(below, d); if (lane >= d) below = fr_mul_chain(below, o); }
    Fr above = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const Fr o = hz_shfl_down_fr(above, d); if (lane + d < 32) above = fr_mul_chain(above, o); }
    Fr all;
#pragma unroll
    for (int i = 0; i < 8; i++) all.l[i] = __shfl_sync(0xffffffffu, below.l[i], 31);
    Fr lo = hz_shfl_up_fr(below, 1), hi = hz_shfl_down_fr(above, 1);
    if (lane == 0) lo = fr_one();
    if (lane == 31) hi = fr_one();
    Fr inv = fr_mul_chain(fr_inv_euclid(all), fr_mul_chain(lo, hi));      // 1 / (own chunk total)
    if (ok) {
        for (int k = lane + 32 * ((HZ_LAD_TOT - 1 - lane) / 32); k >= 0; k -= 32) {
            const Fr x = fr_mul_chain(inv, pre[k]);
            const Fr nk = hz_canon2p(num[k]);
            if (k < HZ_LAD_N) hint[hz_lad_pos(k)] = fr_mul_chain(nk, x);
            else { aff[2 * (k - HZ_LAD_N)] = fr_mul_chain(nk, x); aff[2 * (k - HZ_LAD_N) + 1] = fr_mul_chain(hz_canon2p(vv[k - HZ_LAD_N]), x); }
            inv = fr_mul_chain(inv, den[k]);
        }
    }
    if (lane == 0) reinterpret_cast<HzFillRec*>(scr + HZ_SCR_REC)->ok = ok ? 1u : 0u;
}
