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
        rec->ok = 0; rec->cur = cur;
        for (int k = 0; k < 8; k++) { rec->e[k] = m.hbits.val[k]; rec->s[k] = m.sbits.val[k]; }
        Fr* mid = h.scratch + HZ_SCR_MID;
        mid[0] = m.d3.x.v; mid[1] = m.d3.y.v; mid[2] = m.en.v; mid[3] = m.R8x.v; mid[4] = m.R8y.v;
    }
}
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_eddsa_tail(Fr* w, HzErrBuf* err, const Fr* cst, const Fr* tab, const Fr* img, Fr* scratch,
                                                         uint32_t win_size, int ns) {
    bool real; const int g = hz_item(ns, real);
    HzHints h = hz_sig_hints(cst, tab, img, scratch, win_size);
    h.scratch = scratch + (size_t)g * HZ_HINT_SCRATCH;               // the ladder kernels worked per signature
    const HzFillRec rec = *reinterpret_cast<const HzFillRec*>(h.scratch + HZ_SCR_REC);
    h.lad_ok = (int)rec.ok;
    HzPending pend; pend.n = 0;
    Comp<MCompute> c(w, rec.cur, real ? err : nullptr, &h, &pend);
    EdMid<MCompute> m;
    const Fr* mid = h.scratch + HZ_SCR_MID;
    m.d3.x.v = mid[0]; m.d3.y.v = mid[1]; m.en.v = mid[2]; m.R8x.v = mid[3]; m.R8y.v = mid[4];
    m.d3.x.sig = m.d3.y.sig = m.en.sig = m.R8x.sig = m.R8y.sig = true;
    m.hbits.clear(254); m.sbits.clear(253);
    for (int k = 0; k < 8; k++) { m.hbits.val[k] = rec.e[k]; m.hbits.sg[k] = 0xffffffffu; m.sbits.val[k] = rec.s[k]; m.sbits.sg[k] = 0xffffffffu; }
    EdDSAPoseidonVerifierTail(c, m);
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
