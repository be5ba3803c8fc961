// CPU port of the witness path -- TEST INFRASTRUCTURE, not a product path.
//
// The circuit templates under circuits_b200/csrc are written once for host and device; this file
// instantiates them for host threads behind the same C ABI as the CUDA library, so that
//   * the `-m "not gpu"` tests can compare every signal with the Python oracle without a GPU, and
//   * bench.py has a compiled multi-threaded CPU baseline (`cpu_baseline`, `--impl reference`)
//     in place of the reference's circom-generated C++ calculator, which cannot be built offline
//     (circom / node / nasm / gmp absent; /root/reference/tools/helpers/actions.js:98-130).
// It is parity-pinned by the independent big-integer oracle (oracle/*.py).  Only tests/ and
// bench.py load the resulting oracle/_ref/libhzw_cpu_port.so; circuits_b200 never does.
#include "../../circuits_b200/csrc/hzw_api.hpp"
#include <thread>
#include <atomic>

template <class F> static void hz_parallel_for(int n, F f) {
    unsigned nt = std::thread::hardware_concurrency();
    const char* e = getenv("HZW_CPU_THREADS");
    if (e && atoi(e) > 0) nt = (unsigned)atoi(e);
    if (nt < 1) nt = 1;
    if ((int)nt > n) nt = (unsigned)n;
    std::atomic<int> next(0);
    auto work = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; f(i); } };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

struct CpuBackend : HzBackend {
    const HzLayout& L;
    // HZW_CPU_HINTS=1: evaluate the EdDSA ladders with the inverse-hint streams of hints.cuh (the
    // device's way) instead of one inversion per division (the reference's way, the default and
    // what the CPU baseline times); the tests run both and expect the same witness
    bool use_hints = false;
    std::vector<Fr> cst, edtab, cimg;
    explicit CpuBackend(const HzLayout& l) : L(l) {
        const char* e = getenv("HZW_CPU_HINTS");
        use_hints = e && atoi(e) > 0;
        if (use_hints) {
            cst.assign(HZ_CST_MAX, fr_zero()); edtab.assign(HZ_EDTAB_LEN, fr_zero());
            cimg.assign((size_t)L.plan.sz.sigVerifier + 64, fr_zero());
            uint32_t counts[2];
            hz_record_fix_consts(cimg.data(), cst.data(), edtab.data(), counts);
            if (counts[0] > HZ_CST_MAX || counts[1] != HZ_EDTAB_LEN) { fprintf(stderr, "hzw cpu port: hint table sizes %u %u\n", counts[0], counts[1]); abort(); }
        }
    }
    // ---- evaluation in two phases (a shard of the batch, or all of it) ---------------------------
    HzShard sh;
    std::vector<Fr> wbuf;                   // sharded runs keep a full-size witness of their own
    std::vector<uint8_t> preb;
    int stride = 1;
    void phase1(Fr* w, const uint8_t* inputs, HzErrBuf* err) {
        const HzPlan& pl = L.plan;
        memset(&w[0], 0, 64);
        w[0].l[0] = 1;
        memcpy(&w[2], inputs, 32ull * pl.n_inputs);
        // bounded sample for the CPU-baseline timing: only every `stride`-th slot / fee tx / block
        stride = 1;
        { const char* e = getenv("HZW_CPU_SAMPLE_STRIDE"); if (e && atoi(e) > 1) stride = atoi(e); }
        preb.assign((size_t)pl.nblocks * 512, 0);
        PreBits<MCompute> pre; pre.b = preb.data();
        SlotHook<MCompute> hook; hook.eddsa = nullptr; hook.smt = nullptr; hook.isExit = nullptr; hook.sz = &pl.sz;
        hz_parallel_for(sh.ns, [&](int q) {
            int i = sh.s0 + q;
            if (i % stride) return;
            std::vector<Fr> scratch(use_hints ? HZ_HINT_SCRATCH : 0);
            HzHints h;
            h.cst = cst.data(); h.ci = 0; h.dat = nullptr; h.di = 0; h.rec = nullptr; h.ri = 0;
            h.edtab = edtab.data(); h.edrec = nullptr; h.ei = 0; h.scratch = scratch.data(); h.lad_ok = 0; h.fill = 0; h.prefilled = 0; h.win_c0 = 0; h.win_size = 0; h.cimg = cimg.empty() ? nullptr : cimg.data(); h.fix_shift = 0;
            Comp<MCompute> c(w, pl.off_decode[i], err, use_hints ? &h : nullptr);
            DecodeOut<MCompute> d;
            PreBits<MCompute> p2 = pre;
            sec_decode(c, pl, i, d, p2);
            c.cur = pl.off_rollup[i];
            RollupTxOut<MCompute> o;
            sec_rollup(c, pl, i, d, o, p2, hook);
        });
        hz_parallel_for(sh.nf, [&](int q) {
            int i = sh.f0 + q;
            if (i % stride) return;
            Comp<MCompute> c(w, pl.off_fee[i], err);
            FeeHook<MCompute> fh; fh.smt = nullptr; fh.sib_idx = 0;
            sec_fee(c, pl, i, fh);
        });
    }
    void phase2(Fr* w, HzErrBuf* err) {
        const HzPlan& pl = L.plan;
        PreBits<MCompute> pre; pre.b = preb.data();
        Comp<MCompute> hc(w, pl.off_hasher, sh.rank == 0 ? err : nullptr);
        uint32_t off_blocks = 0;
        W32<MCompute> ha[8];
        sec_hasher_head(hc, pl, pre, &off_blocks, ha);
        // chaining values, then the blocks in parallel
        std::vector<uint32_t> chain((size_t)(pl.nblocks + 1) * 8), msgw((size_t)pl.nblocks * 16);
        for (int j = 0; j < 8; j++) chain[j] = hz_sha_h(j);
        for (uint32_t b = 0; b < pl.nblocks; b++) {
            for (int t = 0; t < 16; t++) {
                uint32_t v = 0;
                for (int k = 0; k < 32; k++) v = (v << 1) | preb[(size_t)b * 512 + 32 * t + k];
                msgw[(size_t)b * 16 + t] = v;
            }
            uint32_t st[8];
            for (int j = 0; j < 8; j++) st[j] = chain[(size_t)b * 8 + j];
            hz_sha256_compress(st, &msgw[(size_t)b * 16]);
            for (int j = 0; j < 8; j++) chain[(size_t)(b + 1) * 8 + j] = st[j];
        }
        hz_parallel_for((int)sh.nb, [&](int q) {
            int b = (int)sh.b0 + q;
            if (b % stride) return;
            Comp<MCompute> shc(w, off_blocks + (uint32_t)b * HZ_SZ_SHA_BLOCK, err);
            W32<MCompute> hin[8], msg[16], outw[8];
            for (int j = 0; j < 8; j++) hin[j].v = chain[(size_t)b * 8 + j];
            for (int t = 0; t < 16; t++) msg[t].v = msgw[(size_t)b * 16 + t];
            sec_sha_block(shc, b, hin, msg, outw);
        });
        W32<MCompute> dig[8];
        for (int j = 0; j < 8; j++) dig[j].v = chain[(size_t)pl.nblocks * 8 + j];
        Comp<MCompute> tc(w, pl.off_hasher_tail, sh.rank == 0 ? err : nullptr);
        Comp<MCompute> shc(w, pl.off_hasher_tail, nullptr);
        VC h = sec_hasher_tail(tc, shc, dig);
        w[1] = h.v;
    }
    int calculate(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (resident || !out || sh.world != 1) return HZW_E_PARAM;
        (void)ms;
        Fr* w = reinterpret_cast<Fr*>(out);
        HzErrBuf* err = sanity ? errs : nullptr;
        phase1(w, inputs, err);
        phase2(w, err);
        return 0;
    }
    int withdraw_batch(uint64_t, const uint8_t*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    // ---- a circom .sym as a layout map, the .r1cs constraint check: host loops (the checker's side of formats.hpp)
    std::vector<uint32_t> map_src;
    int set_layout_map(const uint32_t* src, uint64_t n) override { map_src.assign(src, src + n); return 0; }
    int calculate_mapped(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (resident || !out || map_src.empty()) return HZW_E_PARAM;
        std::vector<Fr> own(L.plan.nvars);
        int st = calculate(inputs, sanity, reinterpret_cast<uint8_t*>(own.data()), false, ms, errs);
        if (st) return st;
        Fr* o = reinterpret_cast<Fr*>(out);
        for (size_t v = 0; v < map_src.size(); v++) o[v] = own[map_src[v]];
        return 0;
    }
    int calculate_to_wtns(const uint8_t* inputs, int sanity, const char* path, float* ms, HzErrBuf* errs) override {
        const uint64_t n = map_src.empty() ? L.plan.nvars : map_src.size();
        std::vector<uint8_t> w(32 * n);
        int st = map_src.empty() ? calculate(inputs, sanity, w.data(), false, ms, errs) : calculate_mapped(inputs, sanity, w.data(), false, ms, errs);
        if (st || errs->count) return st;
        return hz_write_wtns(path, w.data(), n) ? HZW_E_IO : 0;
    }
    int r1cs_check(const HzR1cs& R, const uint8_t* witness, uint64_t* n_bad, uint64_t* first_bad, float*) override {
        if (!witness) return HZW_E_PARAM;            // nothing stays resident in the CPU port
        const Fr* w = reinterpret_cast<const Fr*>(witness);
        const Fr* coef = reinterpret_cast<const Fr*>(R.coef.data());
        uint64_t bad = 0, first = ~0ull;
        auto lc = [&](uint64_t a, uint64_t b) { Fr s = fr_zero(); for (uint64_t t = a; t < b; t++) s = fr_add(s, fr_mul(coef[t], w[R.wire[t]])); return s; };
        for (uint32_t c = 0; c < R.n_constraints; c++) {
            const Fr A = lc(R.off[3ull * c], R.off[3ull * c + 1]), B = lc(R.off[3ull * c + 1], R.off[3ull * c + 2]), C = lc(R.off[3ull * c + 2], R.off[3ull * c + 3]);
            if (!fr_eq(fr_mul(A, B), C)) { bad++; if (first == ~0ull) first = c; }
        }
        if (n_bad) *n_bad = bad;
        if (first_bad) *first_bad = first;
        return 0;
    }
    // ---- shards: the same protocol as the CUDA library, records in host memory -------------------
    int set_shard(int rank, int world) override { sh = hz_shard_ranges(L.plan, rank, world); return 0; }
    uint64_t exchange_bytes() override { return hz_xrec_bytes(L.plan.P, sh.per); }
    int pieces(uint64_t* table, int cap) override { return hz_shard_pieces(L, sh, table, cap); }
    int shard_phase1(const uint8_t* inputs, int sanity, void* record_out, HzErrBuf* errs) override {
        const HzPlan& pl = L.plan;
        const HzParams& P = pl.P;
        wbuf.assign(pl.nvars, fr_zero());
        phase1(wbuf.data(), inputs, sanity ? errs : nullptr);
        uint8_t* rec = static_cast<uint8_t*>(record_out);
        memset(rec, 0, hz_xrec_bytes(P, sh.per));
        const uint32_t sb = hz_xrec_slot_bytes(P), nL2 = sb - 624, p1 = hz_pre_pos_l1(P), p2 = hz_pre_pos_l2(P);
        for (int q = 0; q < sh.ns; q++) {
            uint32_t i = (uint32_t)(sh.s0 + q);
            if ((int)i < P.maxL1Tx) memcpy(rec + (size_t)q * sb, &preb[p1 + i * 624], 624);
            memcpy(rec + (size_t)q * sb + 624, &preb[p2 + i * nL2], nL2);
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(rec + hz_xrec_vals_off(P, sh.per));
        bool last_slot = sh.ns > 0 && sh.s0 + sh.ns == P.nTx, last_fee = sh.nf > 0 && sh.f0 + sh.nf == P.maxFeeTx;
        if (last_slot) { memcpy(out, wbuf[pl.idx_newLastIdx].l, 32); memcpy(out + 16, wbuf[pl.idx_newExitRoot].l, 32); }
        if (last_fee) memcpy(out + 8, wbuf[pl.idx_newStateRoot].l, 32);
        out[24] = last_slot; out[25] = last_fee; out[26] = last_slot;
        return 0;
    }
    int shard_phase2(const void* records, int sanity, uint8_t* shard_out, float* ms, HzErrBuf* errs) override {
        (void)ms;
        const HzPlan& pl = L.plan;
        const HzParams& P = pl.P;
        const uint8_t* recs = static_cast<const uint8_t*>(records);
        const uint32_t sb = hz_xrec_slot_bytes(P), nL2 = sb - 624, p1 = hz_pre_pos_l1(P), p2 = hz_pre_pos_l2(P);
        const size_t rb = hz_xrec_bytes(P, sh.per);
        for (int r = 0; r < sh.world; r++) {
            for (int q = 0; q < sh.per; q++) {
                uint32_t i = (uint32_t)(r * sh.per + q);
                if ((int)i >= P.nTx) break;
                if ((int)i < P.maxL1Tx) memcpy(&preb[p1 + i * 624], recs + r * rb + (size_t)q * sb, 624);
                memcpy(&preb[p2 + i * nL2], recs + r * rb + (size_t)q * sb + 624, nL2);
            }
            const uint32_t* in = reinterpret_cast<const uint32_t*>(recs + r * rb + hz_xrec_vals_off(P, sh.per));
            if (in[24]) memcpy(wbuf[pl.idx_newLastIdx].l, in, 32);
            if (in[25]) memcpy(wbuf[pl.idx_newStateRoot].l, in + 8, 32);
            if (in[26]) memcpy(wbuf[pl.idx_newExitRoot].l, in + 16, 32);
        }
        phase2(wbuf.data(), sanity ? errs : nullptr);
        if (shard_out) {
            uint64_t t[32]; int n = hz_shard_pieces(L, sh, t, 16); size_t off = 0;
            for (int k = 0; k < n; k++) { memcpy(shard_out + off, &wbuf[t[2 * k]], 32 * t[2 * k + 1]); off += 32 * t[2 * k + 1]; }
        }
        return 0;
    }
};

HzBackend* hz_backend_create(const hzw_params&, const HzLayout& L, int*) {
    CpuBackend* b = new CpuBackend(L);
    b->sh = hz_shard_ranges(L.plan, 0, 1);
    return b;
}

struct CpuWithdrawBackend : HzBackend {
    HzWLayout WL;
    int calculate(const uint8_t*, int, uint8_t*, bool, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int set_shard(int, int) override { return HZW_E_PARAM; }
    uint64_t exchange_bytes() override { return 0; }
    int pieces(uint64_t*, int) override { return 0; }
    int shard_phase1(const uint8_t*, int, void*, HzErrBuf*) override { return HZW_E_PARAM; }
    int shard_phase2(const void*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int withdraw_batch(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, float*, HzErrBuf* errs) override {
        hz_parallel_for((int)n, [&](int i) {
            Fr* w = reinterpret_cast<Fr*>(out) + (size_t)i * WL.nvars;
            memset(&w[0], 0, 64);
            w[0].l[0] = 1;
            memcpy(&w[2], inputs + 32ull * WL.n_inputs * i, 32ull * WL.n_inputs);
            Comp<MCompute> c(w, 2 + WL.n_inputs, sanity ? errs : nullptr);
            VC h = sec_withdraw(c, WL.nLevels);
            w[1] = h.v;
        });
        return 0;
    }
};
// standalone RollupTx(nLevels, maxFeeTx): the template as the main component, one instance per call / per thread
struct CpuRollupTxBackend : HzBackend {
    HzTLayout TL;
    int calculate(const uint8_t*, int, uint8_t*, bool, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int set_shard(int, int) override { return HZW_E_PARAM; }
    uint64_t exchange_bytes() override { return 0; }
    int pieces(uint64_t*, int) override { return 0; }
    int shard_phase1(const uint8_t*, int, void*, HzErrBuf*) override { return HZW_E_PARAM; }
    int shard_phase2(const void*, int, uint8_t*, float*, HzErrBuf*) override { return HZW_E_PARAM; }
    int withdraw_batch(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, float*, HzErrBuf* errs) override {
        hz_parallel_for((int)n, [&](int i) {
            Fr* w = reinterpret_cast<Fr*>(out) + (size_t)i * TL.nvars;
            memset(&w[0], 0, 32);
            w[0].l[0] = 1;
            for (uint32_t k = 0; k < TL.n_inputs; k++)
                memcpy(&w[hz_rolluptx_in_index(k, TL.nLevels, TL.maxFeeTx)], inputs + 32ull * ((size_t)TL.n_inputs * i + k), 32);
            Comp<MCompute> c(w, 1, sanity ? errs : nullptr);
            RollupTxOut<MCompute> o;
            sec_rolluptx_main(c, TL.nLevels, TL.maxFeeTx, o);
        });
        return 0;
    }
};
HzBackend* hz_backend_create_rolluptx(const hzw_params&, const HzTLayout& TL, int*) {
    CpuRollupTxBackend* b = new CpuRollupTxBackend();
    b->TL = TL;
    return b;
}
HzBackend* hz_backend_create_withdraw(const hzw_params&, const HzWLayout& WL, int*) {
    CpuWithdrawBackend* b = new CpuWithdrawBackend();
    b->WL = WL;
    return b;
}
