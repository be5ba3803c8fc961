// Backend-independent half of the C ABI (include/hzwitness.h): context, input table, layout
// queries, `.wtns`.  The backend (CUDA in hzw.cu; host threads in oracle/cpu_port, the
// checker) supplies hz_backend_create().
#pragma once
#include "host_common.hpp"
#include "../../include/hzwitness.h"
#include <memory>

struct HzBackend {
    virtual ~HzBackend() {}
    // inputs: n_inputs x 32 B; witness_out may be null when resident
    virtual int calculate(const uint8_t* inputs, int sanity, uint8_t* witness_out, bool resident, float* kernel_ms,
                          HzErrBuf* errs) = 0;
    virtual int kernel_ms(float*, const char**, int) { return 0; }
    virtual int profile_step(float*, const char**, int) { return HZW_E_PARAM; }
    virtual int image_levels(uint64_t*) { return HZW_E_PARAM; }
    // sharded evaluation: records live in the backend's memory space (device memory for CUDA)
    virtual int set_shard(int rank, int world) = 0;
    virtual uint64_t exchange_bytes() = 0;
    virtual int pieces(uint64_t* table, int cap) = 0;
    // a context that spans several GPUs (n_gpus > 1): the pieces and the resident witness of its rank r
    virtual int rank_pieces(int r, uint64_t* table, int cap) { return r == 0 ? pieces(table, cap) : HZW_E_PARAM; }
    virtual int rank_witness_device(int r, void** ptr, int32_t* device, void** stream) { return r == 0 ? witness_device(ptr, device, stream) : HZW_E_PARAM; }
    virtual int shard_phase1(const uint8_t* inputs, int sanity, void* record_out, HzErrBuf* errs) = 0;
    virtual int shard_phase2(const void* records, int sanity, uint8_t* shard_out, float* kernel_ms, HzErrBuf* errs) = 0;
    // Withdraw(nLevels): n independent instances, inputs n x ninputs x 32 B, witnesses n x nvars x 32 B
    virtual int withdraw_batch(uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, float* kernel_ms, HzErrBuf* errs) = 0;
    // the same with the witnesses left in device memory: only witness entry 1 (the public output) of every instance comes back
    virtual int withdraw_resident(uint64_t, const uint8_t*, int, uint8_t*, float*, HzErrBuf*) { return HZW_E_PARAM; }
    virtual int debug_read(int, uint64_t, uint64_t, uint64_t, uint8_t*) { return HZW_E_PARAM; }
    // a circom .sym applied as a layout map (formats.hpp): src[v] = own index of the compiler's witness entry v
    virtual int set_layout_map(const uint32_t*, uint64_t) { return HZW_E_PARAM; }
    // the batch with the witness delivered in the compiler's layout (n_out entries)
    virtual int calculate_mapped(const uint8_t*, int, uint8_t*, bool, float*, HzErrBuf*) { return HZW_E_PARAM; }
    virtual int read_mapped(uint64_t, uint64_t, uint8_t*) { return HZW_E_PARAM; }
    // the batch straight into a `.wtns` file (no file when a constraint is violated under sanity)
    virtual int calculate_to_wtns(const uint8_t*, int, const char*, float*, HzErrBuf*) { return HZW_E_PARAM; }
    // A.w * B.w = C.w over an .r1cs; witness null = the resident one
    virtual int r1cs_check(const HzR1cs&, const uint8_t*, uint64_t*, uint64_t*, float*) { return HZW_E_PARAM; }
    // device-resident witness (null / HZW_E_PARAM for a backend without device memory)
    virtual int witness_device(void** ptr, int32_t* device, void** stream) { (void)ptr; (void)device; (void)stream; return HZW_E_PARAM; }
    virtual int witness_read(uint64_t first, uint64_t count, uint8_t* out) { (void)first; (void)count; (void)out; return HZW_E_PARAM; }
    uint64_t launches = 0, h2d = 0, d2h = 0;
};

// ---- shard geometry shared by the backends ------------------------------------------------------
struct HzShard { int rank = 0, world = 1, per = 0, s0 = 0, ns = 0, f0 = 0, nf = 0; uint32_t b0 = 0, nb = 0; };
static inline HzShard hz_shard_ranges(const HzPlan& pl, int rank, int world) {
    const HzParams& P = pl.P;
    HzShard d; d.rank = rank; d.world = world;
    d.per = (P.nTx + world - 1) / world;
    d.s0 = rank * d.per < P.nTx ? rank * d.per : P.nTx;
    d.ns = (d.s0 + d.per <= P.nTx ? d.per : P.nTx - d.s0);
    int perf = (P.maxFeeTx + world - 1) / world;
    d.f0 = rank * perf < P.maxFeeTx ? rank * perf : P.maxFeeTx;
    d.nf = (d.f0 + perf <= P.maxFeeTx ? perf : P.maxFeeTx - d.f0);
    uint32_t perb = (pl.nblocks + (uint32_t)world - 1) / (uint32_t)world;
    d.b0 = (uint32_t)rank * perb < pl.nblocks ? (uint32_t)rank * perb : pl.nblocks;
    d.nb = d.b0 + perb <= pl.nblocks ? perb : pl.nblocks - d.b0;
    return d;
}
// witness pieces a shard owns, (first index, count) pairs in witness order: rank 0 carries the header,
// the main inputs and the head / tail of hasherInputs besides its slots
static inline int hz_shard_pieces(const HzLayout& L, const HzShard& d, uint64_t* table, int cap) {
    const HzPlan& pl = L.plan;
    const HzParams& P = pl.P;
    int n = 0;
    auto add = [&](uint64_t a, uint64_t b) { if (b > a && n < cap) { table[2 * n] = a; table[2 * n + 1] = b - a; n++; } };
    if (d.rank == 0) add(0, 2 + pl.n_inputs);
    if (d.ns > 0) {
        add(L.off_decode[d.s0], d.s0 + d.ns < P.nTx ? L.off_decode[d.s0 + d.ns] : L.off_rollup[0]);
        add(L.off_rollup[d.s0], d.s0 + d.ns < P.nTx ? L.off_rollup[d.s0 + d.ns] : L.off_fee[0]);
    }
    if (d.nf > 0) add(L.off_fee[d.f0], d.f0 + d.nf < P.maxFeeTx ? L.off_fee[d.f0 + d.nf] : pl.off_hasher);
    if (d.rank == 0) add(pl.off_hasher, pl.off_sha_blocks);
    add((uint64_t)pl.off_sha_blocks + (uint64_t)d.b0 * HZ_SZ_SHA_BLOCK, (uint64_t)pl.off_sha_blocks + (uint64_t)(d.b0 + d.nb) * HZ_SZ_SHA_BLOCK);
    if (d.rank == 0) add(pl.off_hasher_tail, pl.nvars);
    return n;
}
// exchange record of one rank (see k_pack in k_hasher.cu): per own slot 624 + nL2 bytes of pre-image
// bits, then newLastIdx / newStateRoot / newExitRoot (3 x 32 B) and 3 x 4 B "this rank produced it" flags
static inline uint32_t hz_xrec_slot_bytes(const HzParams& P) { return 624u + 2u * (uint32_t)P.nLevels + 48u; }
static inline uint32_t hz_xrec_vals_off(const HzParams& P, int per) { return ((uint32_t)per * hz_xrec_slot_bytes(P) + 3u) & ~3u; }
static inline uint32_t hz_xrec_bytes(const HzParams& P, int per) { return (hz_xrec_vals_off(P, per) + 3 * 32 + 3 * 4 + 15u) & ~15u; }
HzBackend* hz_backend_create(const hzw_params& prm, const HzLayout& L, int* status);
HzBackend* hz_backend_create_withdraw(const hzw_params& prm, const HzWLayout& WL, int* status);
HzBackend* hz_backend_create_rolluptx(const hzw_params& prm, const HzTLayout& TL, int* status);

struct hzw_ctx {
    hzw_params prm;
    HzParams P;
    HzLayout L;              // rollup-main
    HzWLayout WL;            // withdraw
    HzTLayout TL;            // standalone rollup-tx
    bool withdraw = false;   // a circuit whose instances are independent (Withdraw or standalone RollupTx): batches, no shards
    bool rtx = false;        // ... and it is the standalone RollupTx
    std::unique_ptr<HzBackend> be;
    HzLayoutMap map;         // set when hzw_create was given a circom .sym: every index that crosses the boundary is the compiler's
    uint64_t own_nvars() const { return rtx ? TL.nvars : (withdraw ? WL.nvars : L.plan.nvars); }
    uint64_t nvars() const { return map.active() ? map.src.size() : own_nvars(); }
    // own index -> the index the caller sees (HZ_NOIDX: the compiler's layout does not hold that signal)
    uint32_t out_index(uint32_t own) const { return map.active() ? (own < map.inv.size() ? map.inv[own] : HZ_NOIDX) : own; }
    uint64_t ninputs() const { return rtx ? TL.n_inputs : (withdraw ? WL.n_inputs : L.plan.n_inputs); }
    void walk(LayoutSink* sk) const {
        if (rtx) { HzTLayout t; hz_rolluptx_walk(TL.nLevels, TL.maxFeeTx, sk, t); }
        else if (withdraw) { HzWLayout t; hz_withdraw_walk(WL.nLevels, sk, t); }
        else { HzLayout t; hz_layout_walk(P, sk, t); }
    }
};

static inline void hz_fr_to_le32(const Fr& f, uint8_t* o) {
    for (int i = 0; i < 8; i++) for (int b = 0; b < 4; b++) o[4 * i + b] = (uint8_t)(f.l[i] >> (8 * b));
}

static inline std::string hz_dec(const Fr& f) {
    // decimal string of a 256-bit value
    uint32_t t[8]; for (int i = 0; i < 8; i++) t[i] = f.l[i];
    std::string s;
    bool nz = true;
    while (nz) {
        uint64_t rem = 0; nz = false;
        for (int i = 7; i >= 0; i--) { uint64_t cur = (rem << 32) | t[i]; t[i] = (uint32_t)(cur / 10); rem = cur % 10; if (t[i]) nz = true; }
        s.insert(s.begin(), (char)('0' + rem));
    }
    return s;
}

extern "C" {

static int hz_apply_sym(hzw_ctx* c, const char* path) {
    // every signal name of the circuit, hashed, against the names of the .sym
    std::vector<std::pair<uint64_t, uint32_t> > names;
    LayoutSink sk; sk.want_names = true; sk.filter = "main"; sk.hashed = &names;
    c->walk(&sk);
    std::string why;
    int rc = hz_load_sym(path, names, (uint32_t)c->own_nvars(), c->map, &why);
    if (rc) { fprintf(stderr, "hzw: layout map %s refused: %s\n", path, why.c_str()); return rc == -1 ? HZW_E_IO : HZW_E_PARAM; }
    return c->be->set_layout_map(c->map.src.data(), c->map.src.size());
}
static int hzw_create_inner(const hzw_params* prm, hzw_ctx** out);
int hzw_create(const hzw_params* prm, const char* layout_map_path, hzw_ctx** out) {
    if (!prm || !out) return HZW_E_PARAM;
    // a layout map is applied by a context of one GPU evaluating one circuit instance per call
    if (layout_map_path && (prm->n_gpus > 1 || prm->circuit_kind != HZW_CIRCUIT_ROLLUP_MAIN)) return HZW_E_PARAM;
    int st = hzw_create_inner(prm, out);
    if (st || !layout_map_path) return st;
    st = hz_apply_sym(*out, layout_map_path);
    if (st) { delete *out; *out = nullptr; }
    return st;
}
static int hzw_create_inner(const hzw_params* prm, hzw_ctx** out) {
    if (prm->circuit_kind == HZW_CIRCUIT_WITHDRAW) {
        if (prm->nLevels < 1 || prm->nLevels > 48) return HZW_E_PARAM;
        hzw_ctx* c = new hzw_ctx();
        c->prm = *prm; c->withdraw = true;
        hz_withdraw_walk(prm->nLevels, nullptr, c->WL);
        int st = 0;
        c->be.reset(hz_backend_create_withdraw(*prm, c->WL, &st));
        if (!c->be) { delete c; return st ? st : HZW_E_DEVICE; }
        *out = c;
        return HZW_OK;
    }
    if (prm->circuit_kind == HZW_CIRCUIT_ROLLUP_TX) {
        if (prm->nLevels < 1 || prm->nLevels + 1 > HZ_MAX_LEVELS || prm->maxFeeTx < 1 || prm->maxFeeTx > HZ_MAX_FEE) return HZW_E_PARAM;
        hzw_ctx* c = new hzw_ctx();
        c->prm = *prm; c->withdraw = true; c->rtx = true;
        hz_rolluptx_walk(prm->nLevels, prm->maxFeeTx, nullptr, c->TL);
        int st = 0;
        c->be.reset(hz_backend_create_rolluptx(*prm, c->TL, &st));
        if (!c->be) { delete c; return st ? st : HZW_E_DEVICE; }
        *out = c;
        return HZW_OK;
    }
    if (prm->circuit_kind != HZW_CIRCUIT_ROLLUP_MAIN) return HZW_E_PARAM;
    HzParams P; P.nTx = prm->nTx; P.nLevels = prm->nLevels; P.maxL1Tx = prm->maxL1Tx; P.maxFeeTx = prm->maxFeeTx;
    if (!hz_params_ok(P)) return HZW_E_PARAM;
    hzw_ctx* c = new hzw_ctx();
    c->prm = *prm; c->P = P;
    hz_layout_walk(P, nullptr, c->L);
    int st = 0;
    c->be.reset(hz_backend_create(*prm, c->L, &st));
    if (!c->be) { delete c; return st ? st : HZW_E_DEVICE; }
    *out = c;
    return HZW_OK;
}
void hzw_destroy(hzw_ctx* ctx) { delete ctx; }
uint64_t hzw_nvars(const hzw_ctx* ctx) { return ctx ? ctx->nvars() : 0; }
uint64_t hzw_ninputs(const hzw_ctx* ctx) { return ctx ? ctx->ninputs() : 0; }

int hzw_input_index(const hzw_ctx* ctx, const char* name, uint64_t* offset, uint64_t* len) {
    if (!ctx || !name) return HZW_E_PARAM;
    if (ctx->rtx) {
        uint64_t at = 0;
        for (int f = 0; f < HZ_ROLLUPTX_NIN; f++) {
            const uint64_t n = (uint64_t)hz_rolluptx_in_len(f, ctx->TL.nLevels, ctx->TL.maxFeeTx);
            if (!strcmp(name, HZ_ROLLUPTX_INPUTS[f].name)) { if (offset) *offset = at; if (len) *len = n; return HZW_OK; }
            at += n;
        }
        return HZW_E_NOSIGNAL;
    }
    if (ctx->withdraw) {
        for (int f = 0; f < WIN_COUNT; f++)
            if (!strcmp(name, HZ_WITHDRAW_INPUT_NAMES[f])) {
                if (offset) *offset = (uint64_t)f;
                if (len) *len = f == WIN_siblingsState ? (uint64_t)ctx->WL.nLevels + 1 : 1;
                return HZW_OK;
            }
        return HZW_E_NOSIGNAL;
    }
    for (int f = 0; f < IN_COUNT; f++)
        if (!strcmp(name, HZ_INPUT_NAMES[f])) {
            if (offset) *offset = ctx->L.plan.in_off[f] - 2;
            if (len) *len = (uint64_t)ctx->L.plan.in_d0[f] * ctx->L.plan.in_d1[f];
            return HZW_OK;
        }
    return HZW_E_NOSIGNAL;
}

static int hz_report(hzw_ctx* ctx, HzErrBuf& eb, hzw_error* err) {
    if (eb.pad != 0) {
        // the jobs derived ahead of k_front differ from the ones its templates arrive at: a bug, never a property of the input
        fprintf(stderr, "hzw: internal error: %u slot(s) with job records that do not match\n", eb.pad);
        if (err) { memset(err, 0, sizeof *err); err->code = HZW_E_DEVICE; }
        return HZW_E_DEVICE;
    }
    if (eb.count == 0) { if (err) memset(err, 0, sizeof *err); return HZW_OK; }
    uint32_t n = eb.count < HZ_MAX_ERR ? eb.count : HZ_MAX_ERR, best = 0;
    for (uint32_t k = 1; k < n; k++) if (eb.rec[k].cur < eb.rec[best].cur) best = k;
    if (err) {
        memset(err, 0, sizeof *err);
        err->code = HZW_E_CONSTRAINT; err->line = eb.rec[best].line; err->witness_index = eb.rec[best].cur;
        hz_fr_to_le32(eb.rec[best].lhs, err->lhs); hz_fr_to_le32(eb.rec[best].rhs, err->rhs);
        // component path: name of the last signal declared before the failing check
        if (ctx->nvars() < 40000000u && eb.rec[best].cur > 1) {
            LayoutSink sk; sk.want_names = true; sk.find_mode = true; sk.find_idx = eb.rec[best].cur - 1;
            ctx->walk(&sk);
            size_t dot = sk.found.rfind('.');
            std::string path = dot == std::string::npos ? sk.found : sk.found.substr(0, dot);
            snprintf(err->where, sizeof err->where, "%s", path.c_str());
        }
    }
    return HZW_E_CONSTRAINT;
}

int hzw_calculate(hzw_ctx* ctx, const uint8_t* inputs, int sanity, uint8_t* witness_out, hzw_error* err) {
    if (!ctx || !inputs || !witness_out) return HZW_E_PARAM;
    HzErrBuf eb; memset(&eb, 0, sizeof eb);
    int st = ctx->map.active() ? ctx->be->calculate_mapped(inputs, sanity, witness_out, false, nullptr, &eb) : ctx->withdraw ? ctx->be->withdraw_batch(1, inputs, sanity, witness_out, nullptr, &eb)
                           : ctx->be->calculate(inputs, sanity, witness_out, false, nullptr, &eb);
    if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
    return hz_report(ctx, eb, err);
}
int hzw_calculate_resident(hzw_ctx* ctx, const uint8_t* inputs, int sanity, float* kernel_ms, hzw_error* err) {
    if (!ctx || ctx->withdraw) return HZW_E_PARAM;
    HzErrBuf eb; memset(&eb, 0, sizeof eb);
    int st = ctx->map.active() ? ctx->be->calculate_mapped(inputs, sanity, nullptr, true, kernel_ms, &eb)
                               : ctx->be->calculate(inputs, sanity, nullptr, true, kernel_ms, &eb);
    if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
    return hz_report(ctx, eb, err);
}
int hzw_calculate_batch(hzw_ctx* ctx, uint64_t n, const uint8_t* inputs, int sanity, uint8_t* witness_out, hzw_error* err) {
    if (!ctx || !inputs || !witness_out) return HZW_E_PARAM;
    if (ctx->withdraw) {
        HzErrBuf eb; memset(&eb, 0, sizeof eb);
        int st = ctx->be->withdraw_batch(n, inputs, sanity, witness_out, nullptr, &eb);
        if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
        return hz_report(ctx, eb, err);
    }
    for (uint64_t k = 0; k < n; k++) {
        int st = hzw_calculate(ctx, inputs + k * 32 * hzw_ninputs(ctx), sanity, witness_out + k * 32 * hzw_nvars(ctx), err);
        if (st) return st;
    }
    return HZW_OK;
}
int hzw_calculate_batch_resident(hzw_ctx* ctx, uint64_t n, const uint8_t* inputs, int sanity, uint8_t* public_out, float* kernel_ms, hzw_error* err) {
    if (!ctx || !inputs || !ctx->withdraw) return HZW_E_PARAM;
    HzErrBuf eb; memset(&eb, 0, sizeof eb);
    int st = ctx->be->withdraw_resident(n, inputs, sanity, public_out, kernel_ms, &eb);
    if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
    return hz_report(ctx, eb, err);
}
int hzw_write_wtns(const hzw_ctx* ctx, const uint8_t* witness, const char* path) {
    if (!ctx || !witness || !path) return HZW_E_PARAM;
    return hz_write_wtns(path, witness, ctx->nvars()) ? HZW_E_IO : HZW_OK;
}
int hzw_calculate_to_wtns(hzw_ctx* ctx, const uint8_t* inputs, int sanity, const char* path, float* kernel_ms, hzw_error* err) {
    if (!ctx || !inputs || !path || ctx->withdraw) return HZW_E_PARAM;
    HzErrBuf eb; memset(&eb, 0, sizeof eb);
    int st = ctx->be->calculate_to_wtns(inputs, sanity, path, kernel_ms, &eb);
    if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
    return hz_report(ctx, eb, err);
}
int hzw_r1cs_check(hzw_ctx* ctx, const char* r1cs_path, const uint8_t* witness, uint64_t* n_constraints, uint64_t* n_bad,
                   uint64_t* first_bad, float* kernel_ms) {
    if (!ctx || !r1cs_path || ctx->withdraw) return HZW_E_PARAM;
    HzR1cs R; std::string why;
    uint8_t prime[32]; hz_fr_to_le32(fr_P(), prime);
    int rc = hz_load_r1cs(r1cs_path, prime, R, &why);
    if (rc) { fprintf(stderr, "hzw: %s: %s\n", r1cs_path, why.c_str()); return rc == -1 ? HZW_E_IO : HZW_E_PARAM; }
    if (R.n_wires > ctx->nvars()) { fprintf(stderr, "hzw: %s has %u wires, the witness %llu entries\n", r1cs_path, R.n_wires, (unsigned long long)ctx->nvars()); return HZW_E_PARAM; }
    if (n_constraints) *n_constraints = R.n_constraints;
    return ctx->be->r1cs_check(R, witness, n_bad, first_bad, kernel_ms);
}
int hzw_witness_device(const hzw_ctx* ctx, void** dev_ptr, int32_t* device, void** stream) {
    if (!ctx || !dev_ptr) return HZW_E_PARAM;
    return ctx->be->witness_device(dev_ptr, device, stream);
}
int hzw_rank_pieces(const hzw_ctx* ctx, int32_t rank, uint64_t* table, int32_t cap) {
    if (!ctx || !table || cap <= 0) return HZW_E_PARAM;
    return ctx->be->rank_pieces(rank, table, cap);
}
int hzw_rank_witness_device(const hzw_ctx* ctx, int32_t rank, void** dev_ptr, int32_t* device, void** stream) {
    if (!ctx || !dev_ptr) return HZW_E_PARAM;
    return ctx->be->rank_witness_device(rank, dev_ptr, device, stream);
}
int hzw_debug_read(hzw_ctx* ctx, int which, uint64_t g, uint64_t first, uint64_t count, uint8_t* out) {
    if (!ctx || !out || ctx->withdraw) return HZW_E_PARAM;
    return ctx->be->debug_read(which, g, first, count, out);
}
int hzw_witness_read(hzw_ctx* ctx, uint64_t first, uint64_t count, uint8_t* out) {
    if (!ctx || !out || first + count > ctx->nvars()) return HZW_E_PARAM;
    if (ctx->map.active()) return ctx->be->read_mapped(first, count, out);
    return ctx->be->witness_read(first, count, out);
}
int hzw_signal_index(const hzw_ctx* ctx, const char* name, uint64_t* index) {
    if (!ctx || !name || !index) return HZW_E_PARAM;
    std::vector<std::pair<std::string, uint32_t> > v;
    LayoutSink sk; sk.want_names = true; sk.filter = name; sk.out = &v;
    ctx->walk(&sk);
    for (auto& e : v) if (e.first == name && e.second != HZ_NOIDX && ctx->out_index(e.second) != HZ_NOIDX) { *index = ctx->out_index(e.second); return HZW_OK; }
    return HZW_E_NOSIGNAL;
}
int hzw_get_signal(const hzw_ctx* ctx, const uint8_t* witness, const char* name, uint8_t out[32]) {
    if (!ctx || !witness || !name || !out) return HZW_E_PARAM;
    std::vector<std::pair<std::string, uint32_t> > v;
    LayoutSink sk; sk.want_names = true; sk.filter = name; sk.out = &v;
    ctx->walk(&sk);
    for (auto& e : v) if (e.first == name && e.second != HZ_NOIDX && ctx->out_index(e.second) != HZ_NOIDX) { memcpy(out, witness + 32ull * ctx->out_index(e.second), 32); return HZW_OK; }
    return HZW_E_NOSIGNAL;
}
int hzw_dump_sym(const hzw_ctx* ctx, const char* prefix, const char* path) {
    if (!ctx || !path) return HZW_E_PARAM;
    std::vector<std::pair<std::string, uint32_t> > v;
    LayoutSink sk; sk.want_names = true; sk.filter = prefix ? prefix : ""; sk.out = &v;
    ctx->walk(&sk);
    FILE* f = fopen(path, "w");
    if (!f) return HZW_E_IO;
    for (auto& e : v) if (ctx->out_index(e.second) != HZ_NOIDX) fprintf(f, "%u,%s\n", ctx->out_index(e.second), e.first.c_str());
    fclose(f);
    return HZW_OK;
}
int hzw_plan_info(const hzw_ctx* ctx, uint64_t out[12]) {
    if (!ctx || !out) return HZW_E_PARAM;
    if (ctx->rtx) { for (int k = 0; k < 12; k++) out[k] = 0; out[0] = ctx->TL.nvars; out[1] = ctx->TL.n_inputs; return HZW_OK; }
    if (ctx->withdraw) { for (int k = 0; k < 12; k++) out[k] = 0; out[0] = ctx->WL.nvars; out[1] = ctx->WL.n_inputs; out[4] = HZ_SZ_SHA_BLOCK; out[5] = 2; return HZW_OK; }
    const HzPlan& pl = ctx->L.plan;
    const HzParams& P = ctx->P;
    out[0] = pl.nvars; out[1] = pl.n_inputs; out[2] = pl.sz.getAx; out[3] = pl.sz.sigVerifier;
    out[4] = HZ_SZ_SHA_BLOCK; out[5] = pl.nblocks;
    out[6] = (P.maxFeeTx > 1 ? ctx->L.off_fee[1] : pl.off_hasher) - ctx->L.off_fee[0];
    out[7] = P.nTx > 2 ? ctx->L.off_decode[2] - ctx->L.off_decode[1] : 0;
    out[8] = P.nTx > 2 ? ctx->L.off_rollup[2] - ctx->L.off_rollup[1] : 0;
    out[9] = pl.sz.phaseJK; out[10] = pl.sz.p2_off; out[11] = pl.off_sha_blocks - pl.off_hasher;
    return HZW_OK;
}
int hzw_last_kernel_ms(const hzw_ctx* ctx, float* ms, const char** names, int cap) {
    if (!ctx || !ms || cap <= 0) return HZW_E_PARAM;
    return ctx->be->kernel_ms(ms, names, cap);
}
int hzw_profile_step(hzw_ctx* ctx, float* ms, const char** names, int cap) {
    if (!ctx || ctx->withdraw || !ms || !names || cap <= 0) return HZW_E_PARAM;
    return ctx->be->profile_step(ms, names, cap);
}
int hzw_image_levels(hzw_ctx* ctx, uint64_t out[2]) {
    if (!ctx || ctx->withdraw || !out) return HZW_E_PARAM;
    return ctx->be->image_levels(out);
}
int hzw_set_shard(hzw_ctx* ctx, int32_t rank, int32_t world) {
    if (!ctx || ctx->withdraw || world < 1 || rank < 0 || rank >= world) return HZW_E_PARAM;
    return ctx->be->set_shard(rank, world);
}
uint64_t hzw_exchange_bytes(const hzw_ctx* ctx) { return ctx ? ctx->be->exchange_bytes() : 0; }
int hzw_shard_pieces(const hzw_ctx* ctx, uint64_t* table, int32_t cap) {
    if (!ctx || !table || cap <= 0) return HZW_E_PARAM;
    return ctx->be->pieces(table, cap);
}
uint64_t hzw_shard_nvars(const hzw_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t t[32]; int n = ctx->be->pieces(t, 16); uint64_t tot = 0;
    for (int k = 0; k < n; k++) tot += t[2 * k + 1];
    return tot;
}
int hzw_shard_phase1(hzw_ctx* ctx, const uint8_t* inputs, int sanity, void* record_out, hzw_error* err) {
    if (!ctx || !record_out) return HZW_E_PARAM;
    HzErrBuf eb; memset(&eb, 0, sizeof eb);
    int st = ctx->be->shard_phase1(inputs, sanity, record_out, &eb);
    if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
    return hz_report(ctx, eb, err);
}
int hzw_shard_phase2(hzw_ctx* ctx, const void* records, int sanity, uint8_t* shard_out, float* kernel_ms, hzw_error* err) {
    if (!ctx || !records) return HZW_E_PARAM;
    HzErrBuf eb; memset(&eb, 0, sizeof eb);
    int st = ctx->be->shard_phase2(records, sanity, shard_out, kernel_ms, &eb);
    if (st) { if (err) { memset(err, 0, sizeof *err); err->code = st; } return st; }
    return hz_report(ctx, eb, err);
}
int hzw_last_stats(const hzw_ctx* ctx, uint64_t* launches, uint64_t* h2d, uint64_t* d2h) {
    if (!ctx) return HZW_E_PARAM;
    if (launches) *launches = ctx->be->launches;
    if (h2d) *h2d = ctx->be->h2d;
    if (d2h) *d2h = ctx->be->d2h;
    return HZW_OK;
}
}  // extern "C"
