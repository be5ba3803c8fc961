// EdDSA: the ladder walk and the batched inversion behind it (see hzw_dev.cuh for the map of the kernels).
//
// k_ed_ladder   The scalar multiplications of a signature are three chains of point operations, each waiting for
//               the point before it: 253 doublings of 8A, the additions into the accumulators of the two
//               EscalarMulAny segments (252), the 83 additions of the fixed-base windows.  The points are walked on
//               projective triples with the rational maps of MontgomeryDouble / MontgomeryAdd (hints.cuh), and none
//               of these values is a witness entry, so the walk stays in the Montgomery domain with lazily reduced
//               values (fr_montgen.inc: 0 <= x~ < 2p, 176-instruction products instead of the canonical 312).  One
//               THREAD per chain: the 16 products of a point operation are independent four at a time, which the
//               instruction scheduler interleaves -- a thread keeps its multiplier pipe busy on its own, where a
//               lane group had eight lanes wait for one carry chain.  A block is three warps over the same 32
//               signatures (doublings / segment adders / fixed-base adders); the doubling warp hands D_t to the
//               adder warp through shared memory, one step behind.
// k_ed_hints    Montgomery's trick over the 588 denominators and the 81 captured W of a signature, a warp per
//               signature (21 elements per lane, the chunk totals combined with shuffle scans, one inversion).
//               Ratios do not depend on the domain: num~ / den~ = num / den, so the hints come out canonical.
#include "hzw_dev.cuh"
#include "fm.cuh"

#define HZ_LAD_SIGS 32
struct FmPt { Fr U, V, W; };
// the shared second half of both maps (hints.cuh: mt_tail) in the domain
__device__ __forceinline__ FmPt fm_tail(const Fr& N, const Fr& Dn, const Fr& W12, const Fr& X1, const Fr& X2, const Fr& Y1, const Fr& A) {
    const Fr S = fm_mul(Dn, Dn), NN = fm_mul(N, N);
    const Fr T = fm_add(fm_add(fm_mul(W12, A), X1), X2);
    const Fr G = fm_sub(fm_mul(NN, W12), fm_mul(T, S));
    const Fr DS = fm_mul(Dn, S);
    FmPt r;
    r.U = fm_mul(G, Dn);
    r.V = fm_sub(fm_mul(N, fm_sub(fm_mul(X1, S), G)), fm_mul(Y1, DS));
    r.W = fm_mul(DS, W12);
    return r;
}
__device__ __forceinline__ FmPt fm_dbl(const FmPt& p, const Fr& A) {
    const Fr UU = fm_mul(p.U, p.U), UW = fm_mul(p.U, p.W), WW = fm_mul(p.W, p.W), VW = fm_mul(p.V, p.W);
    const Fr AUW = fm_mul(UW, A);
    const Fr N = fm_add(fm_add(fm_add(fm_add(UU, UU), UU), fm_add(AUW, AUW)), WW);
    return fm_tail(N, fm_add(VW, VW), WW, UW, UW, VW, A);
}
// in1 = p1, in2 = p2; num / den = 1 / (u2 - u1)
__device__ __forceinline__ FmPt fm_padd(const FmPt& p1, const FmPt& p2, const Fr& A, Fr& num, Fr& den) {
    const Fr X1 = fm_mul(p1.U, p2.W), X2 = fm_mul(p2.U, p1.W), Y1 = fm_mul(p1.V, p2.W), Y2 = fm_mul(p2.V, p1.W);
    const Fr W12 = fm_mul(p1.W, p2.W);
    const Fr N = fm_sub(Y2, Y1), Dn = fm_sub(X2, X1);
    num = W12; den = Dn;
    return fm_tail(N, Dn, W12, X1, X2, Y1, A);
}
// in2 affine (its W is 1~): the accumulating adders of the fixed-base windows
__device__ __forceinline__ FmPt fm_padd_aff(const FmPt& p1, const Fr& u2, const Fr& v2, const Fr& A, Fr& num, Fr& den) {
    const Fr X2 = fm_mul(u2, p1.W), Y2 = fm_mul(v2, p1.W);
    const Fr N = fm_sub(Y2, p1.V), Dn = fm_sub(X2, p1.U);
    num = p1.W; den = Dn;
    return fm_tail(N, Dn, p1.W, p1.U, X2, p1.V, A);
}
__device__ __forceinline__ void hz_lad_put(Fr* p, const Fr& v) {
    *reinterpret_cast<uint4*>(p) = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    *(reinterpret_cast<uint4*>(p) + 1) = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ void hz_lad_cap(Fr* scr, int ex, const FmPt& P) {
    hz_lad_put(&scr[HZ_SCR_ARR + HZ_LAD_N + ex], P.U);          // num
    hz_lad_put(&scr[HZ_SCR_VV + ex], P.V);
    hz_lad_put(&scr[2 * HZ_SCR_ARR + HZ_LAD_N + ex], P.W);      // den
}

// block = 3 warps over the signatures [32 b, 32 b + 32): warp 0 the doubling chain, warp 1 the adders of the two
// EscalarMulAny segments, warp 2 the fixed-base adders.  Signature g works on scratch + g * HZ_HINT_SCRATCH.
__global__ void __launch_bounds__(96) k_ed_ladder(Fr* scratch, const Fr* tab, int ns) {
    __shared__ uint32_t dsh[2][3][8][HZ_LAD_SIGS];
    const int role = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    int g = (int)blockIdx.x * HZ_LAD_SIGS + c;
    const bool real = g < ns;
    if (!real) g = ns - 1;                                 // padding lanes redo the last signature and store nothing
    Fr* scr = scratch + (size_t)g * HZ_HINT_SCRATCH;
    Fr* num = scr + HZ_SCR_ARR; Fr* den = scr + 2 * HZ_SCR_ARR;
    const HzFillRec rec = *reinterpret_cast<const HzFillRec*>(scr + HZ_SCR_REC);
    const Fr A = fm_A();
    // 8A with EscalarMulAny's zero-point substitution (escalarmulany.circom): p.x == 0 -> BASE8
    Fr px = scr[HZ_SCR_MID], py = scr[HZ_SCR_MID + 1];
    if (fr_is_zero(px)) { px = fr_base8_x(); py = fr_base8_y(); }
    // Edwards2Montgomery: u = (1+y)/(1-y), v = u/x  ->  ((1+y) x : (1+y) : (1-y) x)
    FmPt D;
    {
        const Fr xm = fm_from(px), ym = fm_from(py), one = fm_one();
        const Fr y1 = fm_add(one, ym);
        D.U = fm_mul(y1, xm); D.V = y1; D.W = fm_mul(fm_sub(one, ym), xm);
    }
    FmPt acc = D;
    Fr wnU = fr_zero(), wnV = fr_zero();
    if (role == 2) {
        acc.U = fm_from(tab[83 * 16]); acc.V = fm_from(tab[83 * 16 + 1]); acc.W = fm_one();
        const int sel = (int)(rec.s[0] & 7u);
        wnU = tab[2 * sel]; wnV = tab[2 * sel + 1];
    }
    for (int t = 0; t <= 253; t++) {
        if (role == 0) {
            if (t <= 252) {
                if (real) {
                    if (t <= 144 && (t & 7) == 0) hz_lad_cap(scr, t >> 3, D);
                    if (t >= 148 && ((t - 148) & 7) == 0) hz_lad_cap(scr, 38 + ((t - 148) >> 3), D);
                    if (t == 147) hz_lad_cap(scr, 77, D);
                    hz_lad_put(&num[t], D.W); hz_lad_put(&den[t], fm_add(D.V, D.V));
                }
                D = fm_dbl(D, A);
                uint32_t (*buf)[8][HZ_LAD_SIGS] = dsh[(t + 1) & 1];
#pragma unroll
                for (int k = 0; k < 8; k++) { buf[0][k][c] = D.U.l[k]; buf[1][k][c] = D.V.l[k]; buf[2][k][c] = D.W.l[k]; }
            }
        } else if (role == 1) {
            if (t >= 1) {
                const uint32_t (*buf)[8][HZ_LAD_SIGS] = dsh[t & 1];
#pragma unroll
                for (int k = 0; k < 8; k++) { D.U.l[k] = buf[0][k][c]; D.V.l[k] = buf[1][k][c]; D.W.l[k] = buf[2][k][c]; }
            }
            if (t == 148) { if (real) hz_lad_cap(scr, 78, acc); acc = D; }
            else if (t >= 1) {
                const bool seg0 = t <= 147;
                if (real) {
                    if (seg0 && t <= 145 && ((t - 1) & 7) == 0) hz_lad_cap(scr, 19 + ((t - 1) >> 3), acc);
                    if (!seg0 && ((t - 149) & 7) == 0) hz_lad_cap(scr, 52 + ((t - 149) >> 3), acc);
                }
                Fr hn, hd;
                const FmPt r = fm_padd(D, acc, A, hn, hd);
                const int el = seg0 ? 253 + (t - 1) : 400 + (t - 149);
                if (real) { hz_lad_put(&num[el], hn); hz_lad_put(&den[el], hd); }
                if ((rec.e[t >> 5] >> (t & 31)) & 1u) acc = r;
            }
        } else {
            if (t < 83) {
                if (real && (t & 7) == 0) hz_lad_cap(scr, 66 + (t >> 3), acc);
                const Fr u2 = fm_from(wnU), v2 = fm_from(wnV);
                const int tt = t + 1 < 83 ? t + 1 : 0;
                const int sel = (int)(((rec.s[(3 * tt) >> 5] >> ((3 * tt) & 31)) & 1u) | (((rec.s[(3 * tt + 1) >> 5] >> ((3 * tt + 1) & 31)) & 1u) << 1) |
                                      (((rec.s[(3 * tt + 2) >> 5] >> ((3 * tt + 2) & 31)) & 1u) << 2));
                wnU = tab[tt * 16 + 2 * sel]; wnV = tab[tt * 16 + 2 * sel + 1];
                Fr hn, hd;
                acc = fm_padd_aff(acc, u2, v2, A, hn, hd);
                if (real) { hz_lad_put(&num[505 + t], hn); hz_lad_put(&den[505 + t], hd); }
            }
        }
        __syncthreads();
    }
    if (real && role == 1) hz_lad_cap(scr, 79, acc);
    if (real && role == 2) hz_lad_cap(scr, 80, acc);
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
    const int c0 = lane * HZ_HINT_PER, c1 = c0 + HZ_HINT_PER < HZ_LAD_TOT ? c0 + HZ_HINT_PER : HZ_LAD_TOT;
    bool ok = true;
    Fr tot = fr_one();
    for (int k = c0; k < c1; k++) {
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
        for (int k = c1 - 1; k >= c0; k--) {
            const Fr x = fr_mul_chain(inv, pre[k]);
            const Fr nk = hz_canon2p(num[k]);
            if (k < HZ_LAD_N) hint[hz_lad_pos(k)] = fr_mul_chain(nk, x);
            else { aff[2 * (k - HZ_LAD_N)] = fr_mul_chain(nk, x); aff[2 * (k - HZ_LAD_N) + 1] = fr_mul_chain(hz_canon2p(vv[k - HZ_LAD_N]), x); }
            inv = fr_mul_chain(inv, den[k]);
        }
    }
    if (lane == 0) reinterpret_cast<HzFillRec*>(scr + HZ_SCR_REC)->ok = ok ? 1u : 0u;
}
