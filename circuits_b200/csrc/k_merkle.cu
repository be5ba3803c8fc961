// SMT processors: heads, transposed levels, image levels, tops; Poseidon alone (see hzw_dev.cuh).
#include "hzw_dev.cuh"
#include "fm.cuh"

// ---- Poseidon alone (the second metric of BASELINE.json: permutations per second) ----------------
// The SMT level hashes as k_smt evaluates them -- a lane group runs the old-path and the new-path
// Poseidon(2) side by side -- but on free-standing inputs: pair q hashes (in[4q], in[4q+1]) and
// (in[4q+2], in[4q+3]) and stores all 2 x 633 signals of the two components at sig[q * 2 * size].
__global__ void __launch_bounds__(HZ_BLOCK, 3) k_poseidon_pairs(const Fr* in, Fr* sig, Fr* digests, int npairs) {
    bool real; int q = hz_item(npairs, real);
    const uint32_t size = poseidon_size(3);
    Comp<MCompute> ca(sig, (uint32_t)q * 2 * size, nullptr), cb(sig, (uint32_t)q * 2 * size + size, nullptr);
    VC a[2], b[2], oa, ob;
    for (int k = 0; k < 2; k++) { a[k].v = in[4 * q + k]; a[k].sig = true; b[k].v = in[4 * q + 2 + k]; b[k].sig = true; }
    PoseidonDual<MCompute>(static_cast<Comp<MCompute>&&>(ca), a, static_cast<Comp<MCompute>&&>(cb), b, 2, oa, ob);
    if (hz_glane() == 0) { digests[2 * q] = oa.v; digests[2 * q + 1] = ob.v; }
}

// head of every processor body, one THREAD per processor: key bits, level-insert flags (one inversion per warp for the
// 33 sibling tests of its 32 processors), the state machine; leaves the level job and the state-machine tuples
__global__ void __launch_bounds__(32) k_smt(HzPlan pl, Fr* w, HzErrBuf* err, const HzSmtJob* smt, HzSmtMid* mids, HzLevJob* lev, Fr* smst,
                                            int s0, int ns, int f0, int nf) {
    const int n = 2 * ns + nf;
    int q = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const bool real = q < n;
    if (!real) q = n - 1;
    q = hz_smt_job_index(pl, q, s0, ns, f0);
    const int L = pl.P.nLevels + 1;
    sec_smt_head(w, real ? err : nullptr, L, smt[q], mids[q], lev[q], smst + (size_t)q * L * 6);
}

// ---- the levels of all SMT processors, transposed ------------------------------------------------------
// 132 of the 142 Poseidon permutations of a transaction are the old-path / new-path hashes of the
// 2 x 33 levels.  A chain = one path of one processor: 33 Poseidon(2) one after the other plus a few
// selections per level.  Thread (warp r, lane c) holds state element r of chain c, a block of 3 warps
// runs 32 chains (16 processors): the S-box of a partial round is one full warp, the MDS rows meet in
// shared memory once per round (k_poseidon_t3 is this kernel without the level glue).  Warp 0 also
// does the per-level glue of its chain (switcher in front of the hash, root behind it).  Entry
// offsets inside a level follow SMTProcessorLevel (templates_circomlib.cuh): oldRoot, newRoot,
// [oldChild, newChild: bottom level only], aux[4], oldProofHash (634), newProofHash (634),
// oldSwitcher (outL, outR, aux), newSwitcher (L, R, outL, outR, aux).
// Levels below the leaf of a sparse tree hash (0, 0): child and sibling are both zero there, on the old path
// and on the new one.  With 2^12 accounts in a tree of depth 32 that is some 19 of the 33 levels of every
// processor.  Their 634 entries do not depend on anything, so when a level is (0, 0) in EVERY chain of the
// block the 65 rounds are skipped, the digest is the constant, and the entries are copied from an image
// recorded at context creation by k_smt_zero (a wide copy, off the critical chain of the levels).
// ---- one Poseidon(2) permutation per chain, its dependency chain in the Montgomery domain ----------------------
// The 65 rounds of a level hash are one chain of products (x -> x^2 -> x^4 -> x^5 -> MDS row, 260 deep), and the
// levels of a processor are chained on top of that: this is the longest dependency chain of a batch.  Every value on
// it is a witness entry and must be STORED canonical, but the chain itself does not need canonical values: it runs on
// lazily reduced Montgomery residues (fm.cuh: 176-instruction products instead of 312) and the canonical value of an
// entry is recovered beside the chain by one reduction (fr_mont_redc).  In a partial round only element 0 goes through
// the S-box: warp 0 does nothing but the chain there and publishes its intermediate values in shared memory, the two
// other warps -- idle during the S-box -- reduce and store them one round later.
//   xch[b][q]  the three MDS inputs of a round (S-box outputs / passed-through elements), Montgomery
//   e0[b][0..2] element 0 of a partial round for the helpers: state before the round constant, in2, in4
__device__ __forceinline__ void hz_sm_put(uint32_t (*a)[32], int c, const Fr& v) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k][c] = v.l[k];
}
__device__ __forceinline__ Fr hz_sm_get(const uint32_t (*a)[32], int c) {
    Fr v;
#pragma unroll
    for (int k = 0; k < 8; k++) v.l[k] = a[k][c];
    return v;
}
#include "poseidon_c3_mont.inc"
// s_in: canonical input of element r (zero for r = 0); hw: where the component's entries start (entry 0 is
// ark[0].in[0] = 0); returns the canonical mix output of the last round of element r (r = 0: the digest)
__device__ __forceinline__ Fr hz_poseidon3_chain(uint32_t (*xch)[3][8][32], uint32_t (*e0)[3][8][32], int r, int c, const Fr& s_in,
                                                 Fr* hw, bool real, const Fr* m) {
    const Fr* C = &hz_poseidon_c(3, 0);
    if (r == 0 && real) hz_store_fr(&hw[0], s_in);
    Fr sm = fm_from(s_in), sc = s_in;                   // the state element, Montgomery and canonical
    uint32_t cur = 1, prev = 1;
    bool prev_partial = false;
    for (int i = 0; i < 65; i++) {
        const bool full = i < 4 || i >= 61;
        const Fr xm = fm_add(sm, hz_pc3m_dev[3 * i + r]);
        const uint32_t sg = cur + 6;
        if (full || r != 0) {
            // own canonical entries: mix.out[r] of the round before, ark.out[r] of this one
            if (i > 0) { sc = fm_canon(sm); if (real) hz_store_fr(&hw[prev + 3 + r], sc); }
            if (real) hz_store_fr(&hw[cur + r], fr_add(sc, C[3 * i + r]));
        }
        if (full) {
            const Fr in2 = fm_mul(xm, xm), in4 = fm_mul(in2, in2), out = fm_mul(in4, xm);
            hz_sm_put(xch[i & 1][r], c, out);
            if (real) { hz_store_fr(&hw[sg + 3 * r], fm_canon(out)); hz_store_fr(&hw[sg + 3 * r + 1], fm_canon(in2)); hz_store_fr(&hw[sg + 3 * r + 2], fm_canon(in4)); }
        } else if (r == 0) {
            hz_sm_put(e0[i & 1][0], c, sm);
            const Fr in2 = fm_mul(xm, xm), in4 = fm_mul(in2, in2), out = fm_mul(in4, xm);
            hz_sm_put(e0[i & 1][1], c, in2); hz_sm_put(e0[i & 1][2], c, in4);
            hz_sm_put(xch[i & 1][0], c, out);
        } else {
            hz_sm_put(xch[i & 1][r], c, xm);
        }
        if (r != 0 && prev_partial) {
            // element 0 of the partial round before this one, from what warp 0 published then
            const int b = (i - 1) & 1;
            const uint32_t psg = prev + 6;
            if (r == 1) {
                const Fr s0 = fm_canon(hz_sm_get(e0[b][0], c));           // state 0 entering round i - 1 = mix.out[0] of round i - 2
                if (real) {
                    hz_store_fr(&hw[prev - (i - 1 <= 4 ? 15u : 9u) + 3], s0);
                    hz_store_fr(&hw[prev], fr_add(s0, C[3 * (i - 1)]));
                    hz_store_fr(&hw[psg + 1], fm_canon(hz_sm_get(e0[b][1], c)));
                }
            } else {
                if (real) {
                    hz_store_fr(&hw[psg + 2], fm_canon(hz_sm_get(e0[b][2], c)));
                    hz_store_fr(&hw[psg], fm_canon(hz_sm_get(xch[b][0], c)));
                }
            }
        }
        __syncthreads();
        Fr a[3];
#pragma unroll
        for (int q = 0; q < 3; q++) a[q] = hz_sm_get(xch[i & 1][q], c);
        fr_mont_dot3(sm, a, m);
        prev_partial = !full;
        prev = cur;
        cur += 6 + (full ? 9 : 3);
    }
    // mix.out[r] of the last round (a full one)
    sc = fm_canon(sm);
    if (real) hz_store_fr(&hw[prev + 3 + r], sc);
    return sc;
}

// ---- siblings through the bulk-copy engine -----------------------------------------------------------------------
// The sibling of a level is a main input (siblings1 / siblings2 / siblingsState rows of the batch: 33 x 32 B per
// processor in HBM).  Warp 0 needs one per processor and level, at the head of that level's dependency chain; instead
// of a global load there, warp 1 -- idle while warp 0 does the glue of a level -- has the copy engine fetch the NEXT
// level's 16 siblings into shared memory (cp.async.bulk, 32 B each, completion counted on an mbarrier) while the
// current level's 65 rounds run.  Two buffers, one mbarrier each, parity per use.
__device__ __forceinline__ uint32_t hz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hz_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(hz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hz_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(hz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hz_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(hz_smem_u32(dst)), "l"(src), "r"(bytes), "r"(hz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void hz_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(hz_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

__global__ void __launch_bounds__(96) k_smt_levels(HzPlan pl, Fr* w, const HzLevJob* jobs, const Fr* smst, HzSmtMid* mids,
                                                   const Fr* img00, uint8_t* zflag, int s0, int ns, int f0, int nf) {
    __shared__ uint32_t xch[2][3][8][32];
    __shared__ uint32_t e0[2][3][8][32];
    __shared__ uint32_t hin[2][8][32];
    __shared__ __align__(32) Fr sibs[2][16];
    __shared__ __align__(8) uint64_t sbar[2];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    const int njobs = 2 * ns + nf, L = pl.P.nLevels + 1;
    if (threadIdx.x == 0) {
        hz_mbar_init(&sbar[0], 1); hz_mbar_init(&sbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // lane c < 16 of warp 1 fetches the siblings of processor 16 blockIdx.x + c
    uint32_t stage_sib = 0;
    if (r == 1 && c < 16) {
        int qq = (int)blockIdx.x * 16 + c;
        if (qq >= njobs) qq = njobs - 1;
        stage_sib = jobs[hz_smt_job_index(pl, qq, s0, ns, f0)].sib;
    }
    auto stage = [&](int lev) {
        if (r != 1) return;
        uint64_t* b = &sbar[lev & 1];
        if (c == 0) hz_mbar_expect_tx(b, 16u * 32u);
        __syncwarp();
        if (c < 16) hz_bulk_g2s(&sibs[lev & 1][c], &w[stage_sib + (uint32_t)lev], 32u, b);
    };
    uint32_t sph[2] = {0u, 0u};
    stage(L - 1);
    int q = (int)blockIdx.x * 16 + (c >> 1);
    if (q >= njobs) q = njobs - 1;                    // padding chains redo the last processor (identical stores)
    const int job = hz_smt_job_index(pl, q, s0, ns, f0);
    const int side = c & 1;                           // 0: old path, 1: new path
    const HzLevJob& J = jobs[job];
    const Fr* stj = smst + (size_t)job * L * 6;
    const Fr o1 = J.o1, n1 = J.n1;
    const uint32_t HS = HZ_SZ_SMTHASH2_WIRED;
    Fr m[3];
    for (int k = 0; k < 3; k++) m[k] = hz_poseidon_mr(3, r)[k];
    Fr child = fr_zero();
    for (int lev = L - 1; lev >= 0; lev--) {
        const bool bottom = lev == L - 1;
        const uint32_t lb = J.cur + (bottom ? 0u : 2u + (uint32_t)(L - 1 - lev) * HZ_SZ_SMT_LEVEL);
        const uint32_t o_aux = lb + 2 + (bottom ? 2u : 0u), o_oph = o_aux + 4, o_nph = o_oph + HS, o_osw = o_nph + HS, o_nsw = o_osw + 3;
        Fr st_top, st_old0, st_bot, st_new1, st_upd;
        // the next level's siblings start crossing now: their buffer was last read two levels up, several barriers ago
        if (lev > 0) stage(lev - 1);
        if (r == 0) {
            st_top = stj[lev * 6 + 0]; st_old0 = stj[lev * 6 + 1]; st_bot = stj[lev * 6 + 2]; st_new1 = stj[lev * 6 + 3]; st_upd = stj[lev * 6 + 5];
            hz_mbar_wait(&sbar[lev & 1], sph[lev & 1]);
            sph[lev & 1] ^= 1u;
            const Fr sib = sibs[lev & 1][c >> 1];
            const Fr lr = fr_from_u64((J.lrbits >> lev) & 1ull);
            Fr hL, hR;
            if (side == 0) {
                if (bottom) hz_store_fr(&w[lb + 2], fr_zero());
                Fr aux = fr_mul(fr_sub(sib, child), lr);
                hL = fr_add(aux, child); hR = fr_sub(sib, aux);
                hz_store_fr(&w[o_osw], hL); hz_store_fr(&w[o_osw + 1], hR); hz_store_fr(&w[o_osw + 2], aux);
            } else {
                if (bottom) hz_store_fr(&w[lb + 3], fr_zero());
                Fr a1 = fr_mul(child, fr_add(st_top, st_bot)), a2 = fr_mul(sib, st_top);
                Fr Ln = fr_add(a1, fr_mul(n1, st_new1)), Rn = fr_add(a2, fr_mul(o1, st_new1));
                Fr aux = fr_mul(fr_sub(Rn, Ln), lr);
                hL = fr_add(aux, Ln); hR = fr_sub(Rn, aux);
                hz_store_fr(&w[o_aux + 1], a1); hz_store_fr(&w[o_aux + 2], a2);
                hz_store_fr(&w[o_nsw], Ln); hz_store_fr(&w[o_nsw + 1], Rn); hz_store_fr(&w[o_nsw + 2], hL);
                hz_store_fr(&w[o_nsw + 3], hR); hz_store_fr(&w[o_nsw + 4], aux);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) { hin[0][k][c] = hL.l[k]; hin[1][k][c] = hR.l[k]; }
        }
        __syncthreads();
        Fr s = fr_zero();
        if (r > 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) s.l[k] = hin[r - 1][k][c];
        }
        const int allz = __syncthreads_and(fr_is_zero(s) ? 1 : 0);
        if (threadIdx.x == 0) zflag[(size_t)blockIdx.x * L + lev] = allz ? 1 : 0;
        // SMTHash2 (wired L, R) -> Poseidon(2)
        Fr* hw = w + (side == 0 ? o_oph : o_nph);
        if (allz) s = img00[HZ_IMG00_DIGEST];
        else s = hz_poseidon3_chain(xch, e0, r, c, s, hw, true, m);
        if (r == 0) {
            if (side == 0) {
                Fr a0 = fr_mul(o1, fr_add(fr_add(st_bot, st_new1), st_upd));
                child = fr_add(a0, fr_mul(s, st_top));
                hz_store_fr(&w[o_aux], a0); hz_store_fr(&w[lb], child);
            } else {
                Fr a3 = fr_mul(s, fr_add(fr_add(st_top, st_bot), st_new1));
                child = fr_add(a3, fr_mul(n1, fr_add(st_old0, st_upd)));
                hz_store_fr(&w[o_aux + 3], a3); hz_store_fr(&w[lb + 1], child);
            }
        }
    }
    if (r == 0) { if (side == 0) mids[job].oc = child; else mids[job].nc = child; }
}

// ---- end-to-end path: the bottom (0, 0) levels stay on the device ---------------------------------------------
// Those levels are 40 % of the witness of a sparse tree and all alike: instead of crossing PCIe they are rebuilt in
// the caller's buffer by host threads from one copy of the image (computed on the device at context creation),
// while the rest of the witness is copied around them.  What the host needs per processor: where its levels
// start, how many bottom levels are images (`nz`: the leading run of flagged levels of its block), and the few
// entries between the images (level headers and switchers), which k_hole_pack gathers into a compact array.
__global__ void __launch_bounds__(128) k_hole_pack(HzPlan pl, const Fr* w, const HzLevJob* jobs, const uint8_t* zflag, HzHole* holes, Fr* gl,
                                                   int s0, int ns, int f0, int nf) {
    const int njobs = 2 * ns + nf, L = pl.P.nLevels + 1;
    const int q = (int)blockIdx.x;
    if (q >= njobs) return;
    const int job = hz_smt_job_index(pl, q, s0, ns, f0);
    const int b = q / 16;
    int nz = 0;
    while (nz < L && zflag[(size_t)b * L + (L - 1 - nz)]) nz++;
    const uint32_t cur = jobs[job].cur;
    if (threadIdx.x == 0) { holes[job].cur = cur; holes[job].nz = (uint32_t)nz; }
    const uint32_t G = nz ? 2u + 14u * (uint32_t)nz : 0u;
    Fr* dst = gl + (size_t)job * HZ_GL_STRIDE(L);
    for (uint32_t j = threadIdx.x; j < G; j += blockDim.x) dst[j] = w[hz_hole_glue_index(cur, j)];
}

// the entries of the (0, 0) levels k_smt_levels skipped: block (b, lev) copies the image into the 32 chains of
// levels-block b when that level was flagged, a warp per 4 chains
__global__ void __launch_bounds__(256) k_smt_zero(HzPlan pl, Fr* w, const HzLevJob* jobs, const Fr* img00, const uint8_t* zflag,
                                                  int s0, int ns, int f0, int nf) {
    const int njobs = 2 * ns + nf, L = pl.P.nLevels + 1;
    const int b = (int)blockIdx.x, lev = (int)blockIdx.y;
    if (!zflag[(size_t)b * L + lev]) return;
    const int wi = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const uint32_t HS = HZ_SZ_SMTHASH2_WIRED;
    const uint4* src = reinterpret_cast<const uint4*>(img00);
    for (int c = wi * 4; c < wi * 4 + 4; c++) {
        int q = b * 16 + (c >> 1);
        if (q >= njobs) q = njobs - 1;
        const HzLevJob& J = jobs[hz_smt_job_index(pl, q, s0, ns, f0)];
        const bool bottom = lev == L - 1;
        const uint32_t lb = J.cur + (bottom ? 0u : 2u + (uint32_t)(L - 1 - lev) * HZ_SZ_SMT_LEVEL);
        const uint32_t o_oph = lb + 2 + (bottom ? 2u : 0u) + 4;
        uint4* dst = reinterpret_cast<uint4*>(w + ((c & 1) == 0 ? o_oph : o_oph + HS));
        for (uint32_t e = (uint32_t)lane; e < 2 * HS; e += 32) __stcs(&dst[e], src[e]);
    }
}

__global__ void __launch_bounds__(HZ_BLOCK) k_tail(HzPlan pl, Fr* w, HzErrBuf* err, HzErrBuf* ebuf, const HzSmtMid* mids, const Fr* isExit, const Fr* isExit_front,
                                                  const HzEddsaJob* jobs, const HzEddsaJob* jobs_front, const HzSmtJob* smt, const HzSmtJob* smt_front,
                                                  int s0, int ns, int f0, int nf) {
    bool real; int q = hz_item(ns + nf, real);
    if (q < ns) {
        int i = s0 + q;
        sec_rollup_tail(w, real ? err : nullptr, pl, i, &mids[2 * i], isExit[i]);
        // the jobs k_edjob / k_smtjob derived ahead of k_front against the ones its templates arrived at
        if (real && hz_glane() == 0) {
            const HzEddsaJob &a = jobs[i], &b = jobs_front[i];
            bool same = fr_eq(a.aySignature, b.aySignature) && fr_eq(a.signSignature, b.signSignature) && fr_eq(a.enabled, b.enabled) &&
                        fr_eq(a.s1Ay, b.s1Ay) && fr_eq(a.S, b.S) && fr_eq(a.R8x, b.R8x) && fr_eq(a.R8y, b.R8y) && a.cur == b.cur;      // (M and the state hashes come from k_sighash / k_hstate on one side only)
            for (int k = 0; k < 2; k++) {
                const HzSmtJob &p = smt[2 * i + k], &q2 = smt_front[2 * i + k];
                same = same && fr_eq(p.oldKey, q2.oldKey) && fr_eq(p.isOld0, q2.isOld0) && fr_eq(p.newKey, q2.newKey) && fr_eq(p.hsel, q2.hsel) &&
                       fr_eq(p.fnc0, q2.fnc0) && fr_eq(p.fnc1, q2.fnc1) && p.cur == q2.cur && p.sib == q2.sib && p.sigmask == q2.sigmask && p.ovidx == q2.ovidx;
            }
            same = same && fr_eq(isExit[i], isExit_front[i]);
            if (!same) atomicAdd(&ebuf->pad, 1u);
        }
    }
    else { int i = f0 + (q - ns); sec_fee_tail(w, real ? err : nullptr, pl, i, mids[2 * pl.P.nTx + i]); }
}

// ---- the same permutations, transposed ------------------------------------------------------------
// Thread (warp r, lane c) holds state element r of permutation c: a block of 3 warps runs 32 permutations.
// A partial round's single S-box is then a full warp of 32 different permutations (warp 0) instead of one
// active lane in eight, and the two other warps issue nothing while they wait for it; the state
// elements meet in shared memory (word-major, conflict-free) once per round for the MDS rows, whose
// constants are uniform per warp.  Same signals at the same indexes as the lane-group form.
__global__ void __launch_bounds__(96) k_poseidon_t3(const Fr* in, Fr* sig, Fr* digests, int nperms) {
    __shared__ uint32_t xch[2][3][8][HZ_PT_CHAINS];
    __shared__ uint32_t e0[2][3][8][HZ_PT_CHAINS];
    const int r = (int)(threadIdx.x >> 5), c = (int)(threadIdx.x & 31);
    int p = (int)blockIdx.x * HZ_PT_CHAINS + c;
    const bool real = p < nperms;
    if (!real) p = nperms - 1;
    const uint32_t size = 1 + 2 * 3 * 65 + 3 * (8 * 3 + 57);
    Fr m[3];
    for (int k = 0; k < 3; k++) m[k] = hz_poseidon_mr(3, r)[k];
    // inputs are wires (no entries); entry 0 is ark[0].in[0] = 0
    const Fr s = r == 0 ? fr_zero() : in[2 * p + (r - 1)];
    const Fr d = hz_poseidon3_chain(xch, e0, r, c, s, sig + (size_t)p * size, real, m);
    if (r == 0 && real) digests[p] = d;
}

extern "C" int hzw_poseidon_bench(int device, uint64_t npairs, int reps, const uint8_t* inputs_le32 /* 4 x npairs */,
                                  uint8_t* digests_le32 /* 2 x npairs, nullable */, float* ms_per_rep) {
    if (npairs == 0 || npairs > (1u << 22) || !inputs_le32) return HZW_E_PARAM;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return HZW_E_DEVICE;
    CK(cudaSetDevice(device));
    const uint32_t size = 1 + 2 * 3 * 65 + 3 * (8 * 3 + 57);   // poseidon_size(3): entries of a wired Poseidon(2)
    Fr *in = nullptr, *sig = nullptr, *dg = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaMalloc(&in, 128ull * npairs));
    CK(cudaMalloc(&sig, 64ull * size * npairs));
    CK(cudaMalloc(&dg, 64ull * npairs));
    CK(cudaMemcpy(in, inputs_le32, 128ull * npairs, cudaMemcpyHostToDevice));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // reps < 0: the lane-group form (as inside k_smt); reps > 0: the transposed form
    const bool lanes = reps < 0;
    if (lanes) reps = -reps;
    const int nperms = 2 * (int)npairs;
    const unsigned tgrid = (unsigned)((nperms + HZ_PT_CHAINS - 1) / HZ_PT_CHAINS);
    if (lanes) k_poseidon_pairs<<<hz_grid((int)npairs), HZ_BLOCK>>>(in, sig, dg, (int)npairs);      // warm-up
    else k_poseidon_t3<<<tgrid, 96>>>(in, sig, dg, nperms);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) {
        if (lanes) k_poseidon_pairs<<<hz_grid((int)npairs), HZ_BLOCK>>>(in, sig, dg, (int)npairs);
        else k_poseidon_t3<<<tgrid, 96>>>(in, sig, dg, nperms);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_per_rep) *ms_per_rep = ms / (reps > 0 ? reps : 1);
    if (digests_le32) CK(cudaMemcpy(digests_le32, dg, 64ull * npairs, cudaMemcpyDeviceToHost));
    cudaFree(in); cudaFree(sig); cudaFree(dg); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// ---- measured ceiling of the pipe the field arithmetic runs on ------------------------------------
// 32 x 32 -> 64-bit multiply-adds (IMAD.WIDE.U32) per second with nothing else in the way: 16 independent
// accumulator chains per thread, full occupancy.  This is the denominator of the integer roofline in
// bench.py (MEASURED_PEAKS.json only has HBM and bf16 figures).
__global__ void __launch_bounds__(256) k_imad_peak(uint64_t* out, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b + blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x + k), "r"(y));
    }
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) r ^= acc[k];
    if (r == 0x1234567ull) out[0] = r;          // keeps the chains alive
}
extern "C" int hzw_int_peak(int device, double* lane_products_per_s) {
    int n = 0;
    if (!lane_products_per_s) return HZW_E_PARAM;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return HZW_E_DEVICE;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    uint64_t* out = nullptr;
    CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    k_imad_peak<<<blocks, 256>>>(out, 3, 5, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_imad_peak<<<blocks, 256>>>(out, 3, 5, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *lane_products_per_s = (double)blocks * 256.0 * 16.0 * iters / (best * 1e-3);
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
