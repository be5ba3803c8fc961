// EdDSA: constants of the fixed-base multiplication, the verifier, the ladder step entries (see hzw_dev.cuh).
#include "hzw_dev.cuh"

// one-time: the constant division sites and window entries of EscalarMulFix(253, BASE8)
__global__ void k_fix_consts(Fr* wtmp, Fr* cst, Fr* tab, uint32_t* counts) {
    hz_record_fix_consts(wtmp, cst, tab, counts);
}

// the constant entries of EscalarMulFix(253, BASE8) for every signature of the shard: its recorded image copied
// to sigVerifier.mulFix of each slot before k_eddsa runs (which then only writes the entries that depend on S)
__global__ void __launch_bounds__(256) k_ed_const(HzPlan pl, Fr* w, const uint32_t* off_eddsa, const Fr* img, uint32_t fix_size, int s0, int ns) {
    const uint32_t per = 2 * fix_size;                                  // uint4 pieces per signature
    const uint64_t total = (uint64_t)per * (uint64_t)ns;
    const uint4* src = reinterpret_cast<const uint4*>(img);
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t g = (uint32_t)(e / per), k = (uint32_t)(e % per);
        uint4* dst = reinterpret_cast<uint4*>(w + off_eddsa[s0 + (int)g] + pl.sz.mulfix_off);
        __stcs(&dst[k], src[k]);
    }
}

// The signature of a slot in two kernels with the ladder walk between them (k_ladder.cu):
//   k_eddsa       getAx (the square root) and the first half of the verifier: bits of S and of the message hash,
//                 the three doublings of A; leaves 8A, the scalar bits and where the second half starts in the
//                 scratch area of its lane group (HzFillRec, HZ_SCR_MID)
//   k_eddsa_tail  the second half: both scalar multiplications with the ladders' hints, the final comparison
__device__ __forceinline__ HzHints hz_sig_hints(const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch, uint32_t win_size) {
    HzHints h;
    h.cst = cst; h.ci = 0; h.dat = nullptr; h.di = 0; h.rec = nullptr; h.ri = 0;
    h.edtab = tab; h.edrec = nullptr; h.ei = 0; h.lad_ok = 0; h.fill = 1; h.prefilled = 1; h.win_c0 = 0; h.win_size = win_size; h.cimg = img; h.fix_shift = 0;
    h.scratch = scratch + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) / HZ_G) * HZ_HINT_SCRATCH;   // per group, not per slot: padding groups redo a slot
    return h;
}
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_eddsa(HzPlan pl, Fr* w, HzErrBuf* err, const HzEddsaJob* jobs,
                                                    const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch, uint32_t win_size, int s0, int ns) {
    bool real; int i = s0 + hz_item(ns, real);
    const HzEddsaJob& j = jobs[i];
    HzHints h = hz_sig_hints(cst, tab, img, scratch, win_size);
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, j.cur, real ? err : nullptr, &h, &pend);
    VC vse, sgn, ay, s1Ay, S, R8x, R8y, Msg;
    vse.v = j.enabled; sgn.v = j.signSignature; ay.v = j.aySignature; s1Ay.v = j.s1Ay;
    S.v = j.S; R8x.v = j.R8x; R8y.v = j.R8y; Msg.v = j.M;
    vse.sig = sgn.sig = ay.sig = s1Ay.sig = S.sig = R8x.sig = R8y.sig = Msg.sig = true;
    VC ax = AySign2Ax(c.sub("getAx"), ay, sgn);
    EdMid<MCompute> m;
    uint32_t cur;
    {
        Comp<MCompute> sv = c.sub("sigVerifier");
        EdDSAPoseidonVerifierHead(sv, vse, ax, s1Ay, S, R8x, R8y, Msg, m);
        cur = sv.cur;
    }
    c.flush_pending();
    if (hz_glane() == 0) {
        HzFillRec* rec = reinterpret_cast<HzFillRec*>(h.scratch + HZ_SCR_REC);
        rec->ok = 0; rec->cur = cur; rec->pad[1] = j.cur + pl.sz.mulfix_off;      // where sigVerifier.mulFix starts
        for (int k = 0; k < 8; k++) { rec->e[k] = m.hbits.val[k]; rec->s[k] = m.sbits.val[k]; }
        Fr* mid = h.scratch + HZ_SCR_MID;
        mid[0] = m.d3.x.v; mid[1] = m.d3.y.v; mid[2] = m.en.v; mid[3] = m.R8x.v; mid[4] = m.R8y.v;
    }
}
// Two sides per signature, in different warps of the same block: warps 0-1 walk the variable-base side (mulAny,
// addRight) of the block's 8 signatures, warps 2-3 the fixed-base multiplication (a component of its own, at the cursor
// k_eddsa noted).  What is left of both after the ladders are conversions and BabyAdds, each waiting for one Euclid
// inversion: 11 in a row on one side, 6 on the other, where one lane group walked all 17.  The fixed-base result meets
// the other side in shared memory for the two comparisons.
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_eddsa_tail(Fr* w, HzErrBuf* err, const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch,
                                                         uint32_t win_size, uint32_t fix_size, int ns) {
    __shared__ Fr left_sh[HZ_BLOCK / HZ_G / 2][2];
    const int side = (int)(threadIdx.x >> 6), gl = (int)((threadIdx.x & 63u) / HZ_G);
    int g = (int)blockIdx.x * (HZ_BLOCK / HZ_G / 2) + gl;
    const bool real = g < ns;
    if (!real) g = ns - 1;                                        // padding groups redo the last signature (identical stores)
    HzHints h = hz_sig_hints(cst, tab, img, scratch, win_size);
    h.scratch = scratch + (size_t)g * HZ_HINT_SCRATCH;               // the ladder kernels worked per signature
    const HzFillRec rec = *reinterpret_cast<const HzFillRec*>(h.scratch + HZ_SCR_REC);
    h.lad_ok = (int)rec.ok;
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, side == 0 ? rec.cur : rec.pad[1], real ? err : nullptr, &h, &pend);
    EdMid<MCompute> m;
    const Fr* mid = h.scratch + HZ_SCR_MID;
    m.d3.x.v = mid[0]; m.d3.y.v = mid[1]; m.en.v = mid[2]; m.R8x.v = mid[3]; m.R8y.v = mid[4];
    m.d3.x.sig = m.d3.y.sig = m.en.sig = m.R8x.sig = m.R8y.sig = true;
    m.hbits.clear(254); m.sbits.clear(253);
    for (int k = 0; k < 8; k++) { m.hbits.val[k] = rec.e[k]; m.hbits.sg[k] = 0xffffffffu; m.sbits.val[k] = rec.s[k]; m.sbits.sg[k] = 0xffffffffu; }
    PtV<MCompute> right, left;
    if (side == 0) right = EdDSAPoseidonVerifierTailAny(c, m);
    else {
        left = EscalarMulFix253(c.sub("mulFix"), m.sbits);
        if (hz_glane() == 0) { left_sh[gl][0] = left.x.v; left_sh[gl][1] = left.y.v; }
    }
    __syncthreads();
    if (side == 0) {
        left.x.v = left_sh[gl][0]; left.y.v = left_sh[gl][1]; left.x.sig = left.y.sig = true;
        c.cur = rec.pad[1] + fix_size;                                 // behind the fixed-base component
        EdDSAPoseidonVerifierTailEq(c, m, left, right);
    }
    c.flush_pending();
}

// the entries of the ladder steps of the signatures k_eddsa walked (hints.cuh): a thread per chunk of steps;
// first the 33 chunks of the two EscalarMulAny segments of every signature, then the 11 of the fixed-base adders
__global__ void __launch_bounds__(128) k_ed_fill(Fr* w, HzErrBuf* err, const Fr* tab, const Fr* scratch, uint32_t win_stride, int ns) {
    const int gid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int nseg = HZ_FILL_C0 + HZ_FILL_C1;
    if (gid >= (nseg + HZ_FILL_CF) * ns) {
        // the 83 window multiplexers of every signature
        const int g3 = gid - (nseg + HZ_FILL_CF) * ns, gw = g3 / 83;
        if (gw >= ns) return;
        const Fr* scw = scratch + (size_t)gw * HZ_HINT_SCRATCH;
        const HzFillRec recw = *reinterpret_cast<const HzFillRec*>(scw + HZ_SCR_REC);
        if (recw.ok) hz_fill_winmux(w, tab, recw, win_stride, g3 % 83);
        return;
    }
    const bool fix = gid >= nseg * ns;
    const int g2 = fix ? gid - nseg * ns : gid;
    const int g = fix ? g2 / HZ_FILL_CF : g2 / nseg, k = fix ? g2 % HZ_FILL_CF : g2 % nseg;
    if (g >= ns) return;
    const Fr* scr = scratch + (size_t)g * HZ_HINT_SCRATCH;          // group g of k_eddsa evaluated slot s0 + g
    const HzFillRec rec = *reinterpret_cast<const HzFillRec*>(scr + HZ_SCR_REC);
    if (!rec.ok) return;                                            // degenerate input: the templates divided on the spot
    if (fix) hz_fill_fix(w, err, scr, scr + HZ_SCR_AFF, tab, rec, k);
    else if (k < HZ_FILL_C0) hz_fill_seg(w, err, scr, scr + HZ_SCR_AFF, rec, 0, k);
    else hz_fill_seg(w, err, scr, scr + HZ_SCR_AFF, rec, 1, k - HZ_FILL_C0);
}
