// Evaluation contexts for the circuit templates (templates_*.cuh).
//
// The templates are written once, as C++ generic over a mode M:
//   MCompute : computes every signal over Fr and stores it at its witness index
//              (device threads in the CUDA kernels; the same code compiled for the host is the
//              CPU port under oracle/cpu_port used for checking and the CPU baseline);
//   MLayout  : host-only walk with no arithmetic at all -- assigns witness indexes, records
//              section offsets for the kernels (Plan) and resolves signal names (the `.sym`
//              equivalent the reference gets from circom).
//
// Witness numbering follows the model in oracle/circuit.py (circom 0.5.45 behaviour, restated):
// signals are numbered depth-first in declaration order, a sub-component occupies the range
// where it is instantiated, and `a <== b` with a single signal on the right makes `a` an alias
// of `b` (no witness entry of its own).  Index 0 is the constant 1.
#pragma once
#include "fr.cuh"

#ifndef HZ_DEVICE_ONLY
#include "formats.hpp"
#include <string>
#include <vector>
#include <cstdio>
#include <cstdlib>
#endif

struct MCompute {};
struct MLayout {};

// ---- lane groups (device) ------------------------------------------------------------------
// On the device every circuit section is evaluated by a GROUP of HZ_G = 8 consecutive lanes in
// lock-step.  Scalar glue is computed redundantly by all lanes (free under SIMT: the warp
// issues the instruction once either way); what differs per lane is
//   * stores: witness entry `idx` is written by lane (idx mod 8) only, so the 32-byte stores of
//     a section spread over the group instead of serialising on one lane;
//   * Poseidon: lane j holds state element j, the S-boxes of a full round run side by side and
//     the MDS row of each lane is gathered with warp shuffles (templates_circomlib.cuh);
//   * constraint failures: reported by lane 0.
// Host builds (CPU port, layout walk) have no groups: HZ_G = 1.
#ifdef HZ_DEVICE_ONLY
#define HZ_G 8
__device__ __forceinline__ uint32_t hz_glane() { return threadIdx.x & (HZ_G - 1); }
__device__ __forceinline__ void hz_store_fr(Fr* p, const Fr& v) {
    uint4 a = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]), b = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    __stcs(reinterpret_cast<uint4*>(p), a);
    __stcs(reinterpret_cast<uint4*>(p) + 1, b);
}
#else
#define HZ_G 1
#endif

#define HZ_NOIDX 0xFFFFFFFFu

// ---- names: only the layout walk ever looks at them ---------------------------------------
struct Nm {
    const char* s; int i, j;
    HZ_HD Nm(const char* s_) : s(s_), i(-1), j(-1) {}
    HZ_HD Nm(const char* s_, int i_) : s(s_), i(i_), j(-1) {}
    HZ_HD Nm(const char* s_, int i_, int j_) : s(s_), i(i_), j(j_) {}
};

// ---- values ------------------------------------------------------------------------------
template <class M> struct V;
template <> struct V<MCompute> {
    Fr v; bool sig;
    bool ph = false;    // placeholder: the value depends on a hash that another kernel evaluates (Comp::skip_hash);
                        // entries computed from it are not stored here -- the kernel that knows the hash writes them
};
template <> struct V<MLayout> {
    uint32_t idx;   // witness index when the value is a signal, HZ_NOIDX for an expression
};
typedef V<MCompute> VC;
typedef V<MLayout> VL;

HZ_HD VC vc(const Fr& f) { VC r; r.v = f; r.sig = false; return r; }
HZ_HD VL vl() { VL r; r.idx = HZ_NOIDX; return r; }

template <class M> struct Ops;
template <> struct Ops<MCompute> {
    static HZ_HD VC K(uint64_t k) { return vc(fr_from_u64(k)); }
    static HZ_HD VC KF(const Fr& f) { return vc(f); }
    static HZ_HD VC pow2(int k) { return vc(fr_pow2(k)); }
};
template <> struct Ops<MLayout> {
    static HZ_HD VL K(uint64_t) { return vl(); }
    static HZ_HD VL KF(const Fr&) { return vl(); }
    static HZ_HD VL pow2(int) { return vl(); }
};

HZ_HD VC operator+(const VC& a, const VC& b) { VC r = vc(fr_add(a.v, b.v)); r.ph = a.ph || b.ph; return r; }
HZ_HD VC operator-(const VC& a, const VC& b) { VC r = vc(fr_sub(a.v, b.v)); r.ph = a.ph || b.ph; return r; }
HZ_HD VC operator*(const VC& a, const VC& b) { VC r = vc(fr_mul(a.v, b.v)); r.ph = a.ph || b.ph; return r; }
HZ_HD VC operator-(const VC& a) { VC r = vc(fr_neg(a.v)); r.ph = a.ph; return r; }
HZ_HD VC operator+(const VC& a, uint64_t k) { VC r = vc(fr_add(a.v, fr_from_u64(k))); r.ph = a.ph; return r; }
HZ_HD VC operator-(const VC& a, uint64_t k) { VC r = vc(fr_sub(a.v, fr_from_u64(k))); r.ph = a.ph; return r; }
HZ_HD VC operator*(const VC& a, uint64_t k) { VC r = vc(fr_mul_u64(a.v, k)); r.ph = a.ph; return r; }
HZ_HD VC operator-(uint64_t k, const VC& a) { VC r = vc(fr_sub(fr_from_u64(k), a.v)); r.ph = a.ph; return r; }
HZ_HD VC operator+(uint64_t k, const VC& a) { VC r = vc(fr_add(fr_from_u64(k), a.v)); r.ph = a.ph; return r; }
HZ_HD VC operator*(uint64_t k, const VC& a) { VC r = vc(fr_mul_u64(a.v, k)); r.ph = a.ph; return r; }
HZ_HD VC vinv(const VC& a) { return vc(fr_inv(a.v)); }
HZ_HD VC vdiv(const VC& a, const VC& b) { return vc(fr_mul(a.v, fr_inv(b.v))); }
HZ_HD uint32_t vbit(const VC& a, int i) { return fr_bit(a.v, i); }
HZ_HD bool vis0(const VC& a) { return fr_is_zero(a.v); }
// value of a boolean signal as a machine bit
HZ_HD uint32_t vlsb(const VC& a) { return a.v.l[0] & 1u; }

HZ_HD VL operator+(const VL&, const VL&) { return vl(); }
HZ_HD VL operator-(const VL&, const VL&) { return vl(); }
HZ_HD VL operator*(const VL&, const VL&) { return vl(); }
HZ_HD VL operator-(const VL&) { return vl(); }
HZ_HD VL operator+(const VL&, uint64_t) { return vl(); }
HZ_HD VL operator-(const VL&, uint64_t) { return vl(); }
HZ_HD VL operator*(const VL&, uint64_t) { return vl(); }
HZ_HD VL operator-(uint64_t, const VL&) { return vl(); }
HZ_HD VL operator+(uint64_t, const VL&) { return vl(); }
HZ_HD VL operator*(uint64_t, const VL&) { return vl(); }
HZ_HD VL vinv(const VL&) { return vl(); }
HZ_HD VL vdiv(const VL&, const VL&) { return vl(); }
HZ_HD uint32_t vbit(const VL&, int) { return 0; }
HZ_HD bool vis0(const VL&) { return true; }
HZ_HD uint32_t vlsb(const VL&) { return 0; }

// ---- bit vectors (arrays of 0/1-valued values), capacity 32*W bits -----------------------
template <class M, int W> struct BVT;
template <int W> struct BVT<MCompute, W> {
    uint32_t val[W];   // bit values
    uint32_t sg[W];    // 1 = the element is a signal (alias-able), 0 = expression / constant
    int n;
    HZ_HD VC get(int i) const {
        VC r; r.v = fr_from_u64((val[i >> 5] >> (i & 31)) & 1u); r.sig = (sg[i >> 5] >> (i & 31)) & 1u; return r;
    }
    HZ_HD uint32_t bit(int i) const { return (val[i >> 5] >> (i & 31)) & 1u; }
    HZ_HD void put(int i, const VC& x) {
        uint32_t m = 1u << (i & 31);
        val[i >> 5] = (val[i >> 5] & ~m) | ((x.v.l[0] & 1u) ? m : 0);
        sg[i >> 5] = (sg[i >> 5] & ~m) | (x.sig ? m : 0);
    }
    HZ_HD void clear(int n_) { n = n_; for (int i = 0; i < W; i++) { val[i] = 0; sg[i] = 0; } }
};
template <int W> struct BVT<MLayout, W> {
    uint32_t idx[32 * W];
    int n;
    HZ_HD VL get(int i) const { VL r; r.idx = idx[i]; return r; }
    HZ_HD uint32_t bit(int) const { return 0; }
    HZ_HD void put(int i, const VL& x) { idx[i] = x.idx; }
    HZ_HD void clear(int n_) { n = n_; for (int i = 0; i < 32 * W; i++) idx[i] = HZ_NOIDX; }
};

template <class M, int W, int W2>
HZ_HD void bv_copy(BVT<M, W>& d, int dpos, const BVT<M, W2>& s, int spos, int n) {
    for (int i = 0; i < n; i++) d.put(dpos + i, s.get(spos + i));
}

// ---- constraint failures -----------------------------------------------------------------
struct HzErrRec { uint32_t cur; uint32_t line; Fr lhs, rhs; };
#define HZ_MAX_ERR 32
struct HzErrBuf { uint32_t count; uint32_t pad; HzErrRec rec[HZ_MAX_ERR]; };

// ---- layout sink (host) -------------------------------------------------------------------
#ifndef HZ_DEVICE_ONLY
// where the Poseidon components of width 5 (HashState) and 7 (the message hash of DecodeTx) start, in walk order:
// the transposed kernels of the device write their entries (recorded while a walk runs with this set)
struct HzPoseidonStarts { std::vector<uint32_t> t5, t7; };
static thread_local HzPoseidonStarts* hz_poseidon_starts = nullptr;
struct LayoutSink {
    // name recording: every signal whose full name starts with `filter` is reported
    bool want_names = false;
    std::string filter;
    std::vector<std::pair<std::string, uint32_t> >* out = nullptr;
    // ... or, with `hashed` set, as (64-bit hash of the name, index): what resolving a whole .sym against the layout needs
    std::vector<std::pair<uint64_t, uint32_t> >* hashed = nullptr;
    // reverse lookup: name of the first signal declared at witness index `find_idx`
    bool find_mode = false;
    uint32_t find_idx = 0;
    std::string found;
    bool active(const std::string& prefix) const {
        if (!want_names) return false;
        if (find_mode) return true;
        size_t n = prefix.size() < filter.size() ? prefix.size() : filter.size();
        return prefix.compare(0, n, filter, 0, n) == 0;
    }
};
static inline std::string hz_nm(const Nm& n) {
    std::string s(n.s);
    char b[32];
    if (n.i >= 0) { snprintf(b, sizeof b, "[%d]", n.i); s += b; }
    if (n.j >= 0) { snprintf(b, sizeof b, "[%d]", n.j); s += b; }
    return s;
}
#endif

// ---- inverse hints -------------------------------------------------------------------------
// `x <-- a / b` costs a field inversion per call when evaluated as written, and the EdDSA verifier
// has ~1300 of them per transaction, each waiting for the previous one.  A division site can take
// its 1/b from a stream instead (hint class):
//   class 1, constant stream: sites whose denominator does not depend on the inputs (the window
//            tables of EscalarMulFix over BASE8) -- recorded once per context by running the
//            template in record mode;
//   class 2, data stream: the Montgomery ladder of EscalarMulAny and the accumulating adders of
//            EscalarMulFix -- produced per transaction by hints.cuh, which walks the same points
//            in projective Edwards coordinates and inverts all denominators with one inversion.
// With no stream installed (CPU port by default, layout walk) a site inverts on the spot.  A wrong
// hint cannot pass unnoticed: every such site is followed by its `===` check.
#ifdef HZ_DEVICE_ONLY
#define HZ_HINT_SCRATCH 3104               /* hints.cuh: hint / num / den / pre arrays, captured points, fill record */
#else
#define HZ_HINT_SCRATCH (4 * 300)
#endif
#define HZ_HINT_MAX 300                    /* data hints of one component (SegmentMulAny(148): 294) */
#define HZ_EDTAB_SEG1 (83 * 16 + 2)        /* window table: per window 8 points (u, v), then dblLast */
#define HZ_EDTAB_LEN (HZ_EDTAB_SEG1 + 2 * 16 + 2)
#define HZ_CST_MAX 720                     /* constant inverses of EscalarMulFix(253) */
struct HzHints {
    const Fr* cst; uint32_t ci;            // constant stream, read position
    Fr* dat; uint32_t di;                  // data stream of the current component (null: none)
    Fr* rec; uint32_t ri;                  // record mode: constant inverses are appended here
    const Fr* edtab;                       // window entries of EscalarMulFix (affine Montgomery points)
    Fr* edrec; uint32_t ei;                // record mode for the table
    Fr* scratch;                           // HZ_HINT_SCRATCH entries owned by this lane group / thread
    int lad_ok;                            // device: scratch[0 .. 588) holds the hints of all ladders (hints.cuh)
    int fill;                              // device: the entries of the ladder steps are left to k_ed_fill
    int prefilled;                         // device: the constant entries of EscalarMulFix are already in the witness (k_ed_const)
    uint32_t win_c0, win_size;             // the windows / carry adders of its segment 0: where they start (record mode: measured) and their size
    const Fr* cimg; uint32_t fix_shift;    // witness image of EscalarMulFix from the recording pass, and where the
                                           // component being evaluated starts: its constant entries are copied
};

// ---- deferred inverses -----------------------------------------------------------------------
// IsZero needs 1/in only for its `inv` entry: `out` is simply (in == 0).  With a pending list
// installed the entry is reserved, the value queued, and all queued values of a section are
// inverted together (one inversion) when the list is flushed.  Without one (CPU port, the
// reference's way) IsZero inverts on the spot.
#define HZ_PEND_CAP 128
struct HzPending { uint32_t n; uint32_t idx[HZ_PEND_CAP]; Fr val[HZ_PEND_CAP]; };

// ---- component contexts -------------------------------------------------------------------
template <class M> struct Comp;

template <> struct Comp<MCompute> {
    Fr* w;              // witness base
    uint32_t cur;       // next free witness index
    uint32_t end;       // reserved range end (HZ_NOIDX when sharing the parent's cursor)
    Comp* parent;
    HzErrBuf* err;      // null = sanity check off
    HzHints* hints;     // null = every division inverts on the spot
    HzPending* pend;    // null = IsZero inverts on the spot
    bool solo = false;  // device: this thread works alone (no lane group shares its stores)
    bool dry = false;   // values only: nothing is stored, no inverse is queued (k_smtjob derives the SMT jobs of a
                        // slot ahead of k_front by running the same templates this way)
    bool skip_hash = false;   // device: Poseidon components are stepped over -- their entries and digests come from the
                              // transposed kernels (k_hstate, k_sighash); with hs_rec set the inputs are recorded for them
    Fr* hs_rec = nullptr;     // 4 field elements per HashState of this slot / fee tx, in evaluation order
    int* hs_n = nullptr;
    bool is_main = false;     // a template evaluated as the MAIN component (standalone RollupTx): its own signals are never
                              // aliased away -- def / alias give every one an entry; not inherited by sub-components
    typedef VC Val;

    HZ_HD Comp(Fr* w_, uint32_t cur_, HzErrBuf* err_, HzHints* h_ = nullptr, HzPending* p_ = nullptr)
        : w(w_), cur(cur_), end(HZ_NOIDX), parent(nullptr), err(err_), hints(h_), pend(p_) {}
    HZ_HD Comp(Comp* p, bool reserved, uint32_t size) : w(p->w), cur(p->cur), parent(reserved ? nullptr : p), err(p->err), hints(p->hints), pend(p->pend), solo(p->solo), dry(p->dry), skip_hash(p->skip_hash), hs_rec(p->hs_rec), hs_n(p->hs_n) {
        if (reserved) { end = p->cur + size; p->cur += size; } else end = HZ_NOIDX;
    }
    HZ_HD ~Comp() {
        if (parent) parent->cur = cur;
#if !defined(__CUDA_ARCH__) && defined(HZ_DEBUG_LAYOUT)
        if (end != HZ_NOIDX && cur != end) { fprintf(stderr, "reserved size mismatch: %u != %u\n", cur, end); abort(); }
#endif
    }
    Comp(const Comp&) = delete;
    HZ_HD Comp(Comp&& o) : w(o.w), cur(o.cur), end(o.end), parent(o.parent), err(o.err), hints(o.hints), pend(o.pend), solo(o.solo), dry(o.dry), skip_hash(o.skip_hash), hs_rec(o.hs_rec), hs_n(o.hs_n), is_main(o.is_main) { o.parent = nullptr; o.end = HZ_NOIDX; }
    // the `inv` entry and the `out` value of IsZero(x) (comparators.circom): inv <-- x != 0 ? 1/x : 0
    HZ_HD VC iszero_inv(const Nm&, const VC& x) {
        const uint32_t slot = cur++;
        const bool z = fr_is_zero(x.v);
        if (dry) return vc(z ? fr_one() : fr_zero());
        if (pend) {
            if (pend->n == HZ_PEND_CAP) flush_pending();
            pend->idx[pend->n] = slot; pend->val[pend->n] = x.v; pend->n++;
        } else {
            put(slot, fr_inv(x.v));
        }
        return vc(z ? fr_one() : fr_zero());
    }
    HZ_HD void flush_pending() {
        if (!pend || pend->n == 0) return;
        // Montgomery's trick; zeros keep inverse 0
        Fr acc = fr_one();
        Fr pre[HZ_PEND_CAP];
        const int n = (int)pend->n;
        for (int i = 0; i < n; i++) { pre[i] = acc; if (!fr_is_zero(pend->val[i])) acc = fr_mul(acc, pend->val[i]); }
        Fr t = inv1(acc);
        for (int i = n - 1; i >= 0; i--) {
            if (fr_is_zero(pend->val[i])) { put(pend->idx[i], fr_zero()); continue; }
            put(pend->idx[i], fr_mul(t, pre[i]));
            t = fr_mul(t, pend->val[i]);
        }
        pend->n = 0;
    }
    // one field inversion; a thread working alone (solo) shares it with the other lanes of its warp when they all
    // arrive together (each with its own value), a lane group with the other groups of its warp (fr_inv)
    HZ_HD Fr inv1(const Fr& x) {
#ifdef HZ_DEVICE_ONLY
        if (solo) return fr_inv_lanes32(x);
#endif
        return fr_inv(x);
    }
    // two divisions whose denominators are both known: one inversion
    HZ_HD void div2(const VC& a1, const VC& b1, const VC& a2, const VC& b2, VC& q1, VC& q2) {
        // exactly one inversion whatever the data (the device batches inversions across a warp)
        const bool z1 = fr_is_zero(b1.v), z2 = fr_is_zero(b2.v);
        const Fr d1 = z1 ? fr_one() : b1.v, d2 = z2 ? fr_one() : b2.v;
        Fr t = fr_inv(fr_mul(d1, d2));
        q1 = vc(z1 ? fr_zero() : fr_mul(a1.v, fr_mul(t, d2)));
        q2 = vc(z2 ? fr_zero() : fr_mul(a2.v, fr_mul(t, d1)));
    }
    // constant sub-components of EscalarMulFix (hint class 1): their entries do not depend on the
    // inputs, so once recorded they are copied instead of recomputed
    HZ_HD bool const_image() const { return hints && hints->cimg && !hints->rec; }
    HZ_HD VC copy_const(const Nm&) {
        VC r; r.v = hints->cimg[cur - hints->fix_shift]; r.sig = true;
        if (!hints->prefilled) put(cur, r.v);
        cur++;
        return r;
    }
    // a / b at a division site of hint class `cls` (0: none)
    HZ_HD VC div(const VC& a, const VC& b, int cls) {
        if (hints && cls == 1) {
            if (hints->rec) { Fr iv = fr_inv(b.v); hints->rec[hints->ri++] = iv; return vc(fr_mul(a.v, iv)); }
            if (hints->cst) return vc(fr_mul(a.v, hints->cst[hints->ci++]));
        } else if (hints && cls == 2 && hints->dat) {
            return vc(fr_mul(a.v, hints->dat[hints->di++]));
        }
        return vdiv(a, b);
    }

    HZ_HD Comp sub(const Nm&) { return Comp(this, false, 0); }
    HZ_HD Comp sub_reserve(const Nm&, uint32_t size) { return Comp(this, true, size); }
    HZ_HD void skip(const Nm&, uint32_t size) { cur += size; }

    // the one place a witness entry is written
    HZ_HD void put(uint32_t slot, const Fr& v) {
        if (dry) return;
#ifdef HZ_DEVICE_ONLY
        if (solo || ((slot ^ threadIdx.x) & (HZ_G - 1)) == 0) hz_store_fr(&w[slot], v);
#else
        w[slot] = v;
#endif
    }
    // n consecutive entries holding the bits of `bits` (bit k -> entry cur + k, or reversed): on the
    // device the lanes of a group take every 8th entry each
    HZ_HD void put_bits(uint64_t bits, int n, bool msb_first = false) {
        if (dry) { cur += n; return; }
        Fr f = fr_zero();
#ifdef HZ_DEVICE_ONLY
        const int step = solo ? 1 : HZ_G;
        for (int k = solo ? 0 : (int)((threadIdx.x - cur) & (HZ_G - 1)); k < n; k += step) {
#else
        for (int k = 0; k < n; k++) {
#endif
            f.l[0] = (uint32_t)((bits >> (msb_first ? n - 1 - k : k)) & 1u);
#ifdef HZ_DEVICE_ONLY
            hz_store_fr(&w[cur + k], f);
#else
            w[cur + k] = f;
#endif
        }
        cur += n;
    }
    HZ_HD uint32_t sig(const Nm&) { return cur++; }
    HZ_HD uint32_t sigs(const Nm&, int n) { uint32_t s = cur; cur += n; return s; }
    HZ_HD VC set(uint32_t slot, const VC& x) { if (!x.ph) put(slot, x.v); VC r = x; r.sig = true; return r; }
    HZ_HD VC seti(const Nm&, uint32_t slot, const VC& x) { return set(slot, x); }
    HZ_HD VC alias(const Nm& nm, const VC& x) { return is_main ? own(nm, x) : x; }
    // `signal s <== x` / wiring of a sub-component input: alias when x is a signal
    HZ_HD VC def(const Nm& nm, const VC& x) {
        if (is_main) return own(nm, x);
        if (x.sig) return x;
        if (!x.ph) put(cur, x.v);
        cur++; VC r = x; r.sig = true; return r;
    }
    // main-component signals always own their entry
    HZ_HD VC own(const Nm&, const VC& x) { if (!x.ph) put(cur, x.v); cur++; VC r = x; r.sig = true; return r; }
    // only for entries written by an EARLIER kernel or by the caller (main inputs): within a
    // kernel the lanes of a group do not see each other's stores
    HZ_HD VC load(uint32_t idx) const { VC r; r.v = w[idx]; r.sig = true; return r; }
    // a signal whose value the caller already holds (no read-back)
    HZ_HD VC known(uint32_t, const VC& x) const { VC r = x; r.sig = true; return r; }
    HZ_HD bool sanity() const { return err != nullptr; }
    HZ_HD void check(const VC& lhs, const VC& rhs, uint32_t line) {
        if (!err) return;
        if (fr_eq(lhs.v, rhs.v)) return;
#ifdef HZ_DEVICE_ONLY
        if (!solo && hz_glane() != 0) return;
#endif
#ifdef __CUDA_ARCH__
        uint32_t k = atomicAdd(&err->count, 1u);
#else
        uint32_t k = __atomic_fetch_add(&err->count, 1u, __ATOMIC_RELAXED);
#endif
        if (k < HZ_MAX_ERR) { err->rec[k].cur = cur; err->rec[k].line = line; err->rec[k].lhs = lhs.v; err->rec[k].rhs = rhs.v; }
    }
    HZ_HD void check0(const VC& lhs, uint32_t line) { check(lhs, vc(fr_zero()), line); }
};

#ifndef HZ_DEVICE_ONLY
template <> struct Comp<MLayout> {
    uint32_t cur, end;
    Comp* parent;
    LayoutSink* sink;
    std::string prefix;
    bool named;
    bool is_main = false;
    typedef VL Val;

    Comp(uint32_t cur_, LayoutSink* s, const std::string& pre) : cur(cur_), end(HZ_NOIDX), parent(nullptr), sink(s), prefix(pre) {
        named = sink && sink->active(prefix);
    }
    Comp(Comp* p, const Nm& nm, bool reserved, uint32_t size) : cur(p->cur), parent(reserved ? nullptr : p), sink(p->sink) {
        if (reserved) { end = p->cur + size; p->cur += size; } else end = HZ_NOIDX;
        named = false;
        if (sink && sink->want_names) {
            prefix = p->prefix + "." + hz_nm(nm);
            named = sink->active(prefix);
        }
    }
    ~Comp() {
        if (parent) parent->cur = cur;
        if (end != HZ_NOIDX && cur != end) {
            fprintf(stderr, "hzw layout: reserved size mismatch in %s: used %u, reserved up to %u\n", prefix.c_str(), cur, end);
            abort();
        }
    }
    Comp(const Comp&) = delete;
    Comp(Comp&& o) : cur(o.cur), end(o.end), parent(o.parent), sink(o.sink), prefix(o.prefix), named(o.named), is_main(o.is_main) { o.parent = nullptr; o.end = HZ_NOIDX; }

    Comp sub(const Nm& nm) { return Comp(this, nm, false, 0); }
    Comp sub_reserve(const Nm& nm, uint32_t size) { return Comp(this, nm, true, size); }
    void skip(const Nm&, uint32_t size) { cur += size; }

    void name(const Nm& nm, uint32_t idx) {
        if (named && sink->find_mode) {
            if (idx == sink->find_idx && sink->found.empty()) sink->found = prefix + "." + hz_nm(nm);
            return;
        }
        if (named) {
            std::string full = prefix + "." + hz_nm(nm);
            if (full.compare(0, sink->filter.size(), sink->filter) == 0) {
                if (sink->hashed) sink->hashed->push_back(std::make_pair(hz_hash64(full.data(), full.size()), idx));
                else sink->out->push_back(std::make_pair(full, idx));
            }
        }
    }
    uint32_t sig(const Nm& nm) { name(nm, cur); return cur++; }
    uint32_t sigs(const Nm& nm, int n) {
        uint32_t s = cur;
        if (named) for (int i = 0; i < n; i++) name(Nm(nm.s, i), s + i);
        cur += n; return s;
    }
    VL set(uint32_t slot, const VL&) { VL r; r.idx = slot; return r; }
    VL seti(const Nm& nm, uint32_t slot, const VL&) { name(nm, slot); VL r; r.idx = slot; return r; }
    VL alias(const Nm& nm, const VL& x) { if (is_main) return own(nm, x); name(nm, x.idx); return x; }
    VL def(const Nm& nm, const VL& x) {
        if (is_main) return own(nm, x);
        if (x.idx != HZ_NOIDX) { name(nm, x.idx); return x; }
        name(nm, cur); VL r; r.idx = cur++; return r;
    }
    VL own(const Nm& nm, const VL&) { name(nm, cur); VL r; r.idx = cur++; return r; }
    VL load(uint32_t idx) const { VL r; r.idx = idx; return r; }
    VL known(uint32_t idx, const VL&) const { VL r; r.idx = idx; return r; }
    bool const_image() const { return false; }
    VL copy_const(const Nm&) { return vl(); }
    VL iszero_inv(const Nm& nm, const VL&) { name(nm, cur); cur++; return vl(); }
    void flush_pending() {}
    VL div(const VL&, const VL&, int) { return vl(); }
    void div2(const VL&, const VL&, const VL&, const VL&, VL& q1, VL& q2) { q1 = vl(); q2 = vl(); }
    bool sanity() const { return false; }
    void check(const VL&, const VL&, uint32_t) {}
    void check0(const VL&, uint32_t) {}
};
#endif

#define HZ_CHECK(c, lhs, rhs) do { if ((c).sanity()) (c).check((lhs), (rhs), __LINE__); } while (0)
#define HZ_CHECK0(c, lhs) do { if ((c).sanity()) (c).check0((lhs), __LINE__); } while (0)
