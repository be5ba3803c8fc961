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
    std::vector<Fr> cst, edtab;
    explicit CpuBackend(const HzLayout& l) : L(l) {
        const char* e = getenv("HZW_CPU_HINTS");
        use_hints = e && atoi(e) > 0;
        if (use_hints) {
            cst.assign(HZ_CST_MAX, fr_zero()); edtab.assign(HZ_EDTAB_LEN, fr_zero());
            std::vector<Fr> tmp((size_t)L.plan.sz.sigVerifier + 64);
            uint32_t counts[2];
            hz_record_fix_consts(tmp.data(), cst.data(), edtab.data(), counts);
            if (counts[0] > HZ_CST_MAX || counts[1] != HZ_EDTAB_LEN) { fprintf(stderr, "hzw cpu port: hint table sizes %u %u\n", counts[0], counts[1]); abort(); }
        }
    }
    int calculate(const uint8_t* inputs, int sanity, uint8_t* out, bool resident, float* ms, HzErrBuf* errs) override {
        if (resident || !out) return HZW_E_PARAM;
        (void)ms;
        const HzPlan& pl = L.plan;
        const HzParams& P = pl.P;
        Fr* w = reinterpret_cast<Fr*>(out);
        memset(&w[0], 0, 64);
        w[0].l[0] = 1;
        memcpy(&w[2], inputs, 32ull * pl.n_inputs);
        HzErrBuf* err = sanity ? errs : nullptr;
        // bounded sample for the CPU-baseline timing: only every `stride`-th slot / fee tx / block
        int stride = 1;
        { const char* e = getenv("HZW_CPU_SAMPLE_STRIDE"); if (e && atoi(e) > 1) stride = atoi(e); }
        std::vector<uint8_t> preb((size_t)pl.nblocks * 512, 0);
        PreBits<MCompute> pre; pre.b = preb.data();
        SlotHook<MCompute> hook; hook.eddsa = nullptr; hook.smt = nullptr; hook.isExit = nullptr; hook.sz = &pl.sz;
        hz_parallel_for(P.nTx, [&](int i) {
            if (i % stride) return;
            std::vector<Fr> scratch(use_hints ? HZ_HINT_SCRATCH : 0);
            HzHints h;
            h.cst = cst.data(); h.ci = 0; h.dat = nullptr; h.di = 0; h.rec = nullptr; h.ri = 0;
            h.edtab = edtab.data(); h.edrec = nullptr; h.ei = 0; h.scratch = scratch.data();
            Comp<MCompute> c(w, pl.off_decode[i], err, use_hints ? &h : nullptr);
            DecodeOut<MCompute> d;
            PreBits<MCompute> p2 = pre;
            sec_decode(c, pl, i, d, p2);
            c.cur = pl.off_rollup[i];
            RollupTxOut<MCompute> o;
            sec_rollup(c, pl, i, d, o, p2, hook);
        });
        hz_parallel_for(P.maxFeeTx, [&](int i) {
            if (i % stride) return;
            Comp<MCompute> c(w, pl.off_fee[i], err);
            FeeHook<MCompute> fh; fh.smt = nullptr; fh.sib_idx = 0;
            sec_fee(c, pl, i, fh);
        });
        Comp<MCompute> hc(w, pl.off_hasher, err);
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
        hz_parallel_for((int)pl.nblocks, [&](int b) {
            if (b % stride) return;
            Comp<MCompute> sh(w, off_blocks + (uint32_t)b * HZ_SZ_SHA_BLOCK, err);
            W32<MCompute> hin[8], msg[16], outw[8];
            for (int j = 0; j < 8; j++) hin[j].v = chain[(size_t)b * 8 + j];
            for (int t = 0; t < 16; t++) msg[t].v = msgw[(size_t)b * 16 + t];
            sec_sha_block(sh, b, hin, msg, outw);
        });
        W32<MCompute> dig[8];
        for (int j = 0; j < 8; j++) dig[j].v = chain[(size_t)pl.nblocks * 8 + j];
        Comp<MCompute> tc(w, pl.off_hasher_tail, err);
        Comp<MCompute> shc(w, pl.off_hasher_tail, err);
        VC h = sec_hasher_tail(tc, shc, dig);
        w[1] = h.v;
        return 0;
    }
};

HzBackend* hz_backend_create(const hzw_params&, const HzLayout& L, int*) { return new CpuBackend(L); }
