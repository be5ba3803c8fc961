"""Generate csrc/fr_mulgen.inc: the multi-limb products behind fr_mul / fr_dot_mont as straight-line
carry chains, once as PTX (`mad.lo.cc.u32` / `madc.hi.cc.u32`, which ptxas fuses into
IMAD.WIDE.U32 with carry for sm_100a) and once as portable C (the same op list, so the host
tests validate the chain construction bit for bit).

    python -m circuits_b200.tools.gen_frmul

A product x[j] * y[i] lands at limb position i + j.  Products whose positions have the same
parity occupy disjoint 64-bit windows, so one row of a schoolbook multiplication is two carry
chains: the even-position windows accumulate in the `e` limbs, the odd-position ones in the `o`
limbs, and e + o is merged once at the end.  No carry ever crosses an asm statement.

Generated routines (all on 32-bit limbs, little endian):
  fr_gen_mulwide   t[0..15]  = a[0..7] * b[0..7]
  fr_gen_qmu       q3[0..8]  = ((t[7..15] * MU) with columns below 7 dropped) >> 288   (Barrett)
  fr_gen_qp        r[0..7]   = (q3[0..8] * P) mod 2^256
  fr_gen_dotN      T[0..16]  = sum_{k<N} s[k] * c[k]                                  (N = 2..7)
  fr_gen_mlow      m[0..7]   = (T[0..7] * (-P^-1 mod 2^256)) mod 2^256
  fr_gen_mpadd     r[0..8]   = (T[0..16] + m * P) >> 256          (Montgomery reduction, second half)
"""
import os

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MU = (1 << 512) // P
NP0 = (-pow(P, -1, 1 << 32)) % (1 << 32)


def limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


class Emit:
    """Collects chains; each chain becomes one asm statement (device) and one C block (host)."""

    def __init__(self):
        self.dev = []
        self.host = []
        self.live = set()       # accumulator variables that hold a value
        self.decl = []          # accumulator variables in order of first use

    def touch(self, v):
        if v not in self.decl:
            self.decl.append(v)

    def chain(self, steps):
        """steps: list of (kind, acc, x, y) with kind in lo/hi/c; x, y = variable name or int."""
        outs, ins, text, hostl = [], [], [], ["{ uint64_t cy_ = 0, s_, p_; (void)p_;"]
        outmap, inmap = {}, {}

        def oref(v, fresh):
            if v not in outmap:
                outmap[v] = len(outs)
                outs.append((v, fresh))
            return "%%%d" % outmap[v]

        # two passes: outputs are numbered first, inputs after
        fresh_of = {}
        for kind, acc, x, y in steps:
            self.touch(acc)
            if acc not in fresh_of:
                fresh_of[acc] = acc not in self.live
            oref(acc, fresh_of[acc])
        nout = len(outs)

        def iref(v):
            if isinstance(v, int):
                return "0x%08x" % v
            if v in outmap:
                # an accumulator used as a multiplicand: only legal when it is not fresh
                assert not fresh_of[v]
                return "%%%d" % outmap[v]
            if v not in inmap:
                inmap[v] = nout + len(ins)
                ins.append(v)
            return "%%%d" % inmap[v]

        written = set()
        first = True
        for k, (kind, acc, x, y) in enumerate(steps):
            d = oref(acc, fresh_of[acc])
            src = d if (not fresh_of[acc] or acc in written) else "0"
            hsrc = acc if (not fresh_of[acc] or acc in written) else "0u"
            cc = ".cc"  # keep the flag live through the chain; the last carry-out is dead
            if kind == "c":
                if first:
                    raise ValueError("carry step cannot start a chain")
                text.append("addc%s.u32 %s, %s, 0;" % (cc, d, src))
                hostl.append("s_ = (uint64_t)%s + cy_; %s = (uint32_t)s_; cy_ = s_ >> 32;" % (hsrc, acc))
            else:
                op = "mad" if first else "madc"
                text.append("%s.%s%s.u32 %s, %s, %s, %s;" % (op, kind, cc, d, iref(x), iref(y), src))
                hx = ("0x%08xu" % x) if isinstance(x, int) else x
                hy = ("0x%08xu" % y) if isinstance(y, int) else y
                part = "(uint32_t)p_" if kind == "lo" else "(p_ >> 32)"
                hostl.append("p_ = (uint64_t)%s * %s; s_ = (uint64_t)%s + %s + cy_; %s = (uint32_t)s_; cy_ = s_ >> 32;"
                             % (hx, hy, hsrc, part, acc))
            written.add(acc)
            first = False
        for acc in written:
            self.live.add(acc)
        cons_out = ", ".join('"%s"(%s)' % ("=&r" if fr else "+r", v) for v, fr in outs)
        cons_in = ", ".join('"r"(%s)' % v for v in ins)
        self.dev.append('asm("%s" : %s : %s);' % (" ".join(text), cons_out, cons_in))
        hostl.append("}")
        self.host.append(" ".join(hostl))

    def row(self, acc_name, pos_js, xs, y, limit=None):
        """One parity class of one row.  pos_js: list of (position, j); xs: multiplicand limbs
        (names or ints) indexed by j; y: the row's multiplier.  Windows above `limit` are cut."""
        steps = []
        last_hi = None
        for pos, j in pos_js:
            if limit is not None and pos > limit:
                continue
            steps.append(("lo", acc_name(pos), xs[j], y))
            last_hi = None
            if limit is None or pos + 1 <= limit:
                steps.append(("hi", acc_name(pos + 1), xs[j], y))
                last_hi = pos + 1
        if not steps:
            return
        # carry out of the chain: a fresh limb of the same accumulator absorbs it; when that limb is
        # live the carry is counted in the `c` limb of that position instead (no ripple: the
        # counters are summed in at the merge)
        if last_hi is not None and acc_name(last_hi) in self.live:
            k = last_hi + 1
            if limit is None or k <= limit:
                if acc_name(k) not in self.live:
                    steps.append(("c", acc_name(k), None, None))
                else:
                    steps.append(("c", "c%d" % k, None, None))
        self.chain(steps)


def schoolbook(em, nx, ny, xs, ys, keep, limit=None):
    """rows over y; keep(i, j) selects products."""
    for i in range(ny):
        for par in (0, 1):
            pj = [(i + j, j) for j in range(nx) if keep(i, j) and (i + j) % 2 == par]
            em.row((lambda p: "e%d" % p) if par == 0 else (lambda p: "o%d" % p), pj, xs, ys[i], limit=limit)


def merge(em, lo, hi, out):
    """out(k) = limb k of sum over the e, o and c accumulators, k = lo..hi (carries run from lo);
    out(k) None = the limb only contributes its carry."""
    have = lambda v: v in em.live
    tmps = []

    def dst(k):
        d = out(k)
        if d is None:
            d = "tmp%d_" % k
            if d not in tmps:
                tmps.append(d)
        return d

    def add_chain(pairs):
        """pairs: list of (dst, a, b) with a/b variable names or None; one add.cc chain."""
        outs = [d for d, _, _ in pairs]
        ins = []
        for _, a, b in pairs:
            for v in (a, b):
                if v is not None and v not in outs and v not in ins:
                    ins.append(v)
        ref = lambda v: "0" if v is None else ("%%%d" % (outs.index(v) if v in outs else len(outs) + ins.index(v)))
        text, hostl = [], ["{ uint64_t cy_ = 0, s_;"]
        for n, (d, a, b) in enumerate(pairs):
            if a is None:
                a, b = b, None
            text.append("%s.cc.u32 %s, %s, %s;" % ("add" if n == 0 else "addc", ref(d), ref(a), ref(b)))
            hostl.append("s_ = (uint64_t)%s + %s + cy_; %s = (uint32_t)s_; cy_ = s_ >> 32;" % (a or "0u", b or "0u", d))
        hostl.append("}")
        # a destination that is also a source must be read-write
        cons = []
        srcs = set(v for _, a, b in pairs for v in (a, b) if v is not None)
        for d in outs:
            cons.append('"%s"(%s)' % ("+r" if d in srcs else "=&r", d))
        em.dev.append('asm("%s" : %s : %s);' % (" ".join(text), ", ".join(cons), ", ".join('"r"(%s)' % i for i in ins)))
        em.host.append(" ".join(hostl))

    ks = list(range(lo, hi + 1))
    first = [(dst(k), ("e%d" % k) if have("e%d" % k) else None, ("o%d" % k) if have("o%d" % k) else None) for k in ks]
    mark = len(em.dev)
    add_chain(first)
    cks = [k for k in ks if have("c%d" % k)]
    if cks:
        second = [(dst(k), dst(k), ("c%d" % k) if have("c%d" % k) else None) for k in range(cks[0], hi + 1)]
        add_chain(second)
    if tmps:
        em.dev.insert(mark, "uint32_t %s;" % ", ".join(tmps))
        em.host.insert(mark, "uint32_t %s;" % ", ".join(tmps))


def gen():
    out = ["// GENERATED by circuits_b200/tools/gen_frmul.py -- do not edit.",
           "// Straight-line carry chains for the multi-limb products of fr.cuh (see the generator).",
           "#pragma once", ""]

    def function(fname, sig, em, prologue="", epilogue="", locals_="", tail="", preload=""):
        accs = list(em.decl)
        out.append("HZ_HD void %s(%s) {" % (fname, sig))
        if prologue:
            out.append("    " + prologue)
        if locals_:
            out.append("    " + locals_)
        if accs:
            out.append("    uint32_t %s;" % ", ".join(accs))
        out.append("#if defined(__CUDA_ARCH__)")
        if preload:
            out.append("    " + preload)
        out.extend("    " + l for l in em.dev)
        out.append("#else")
        # host: fresh accumulators start at zero (the chains read `0u` for them anyway)
        out.append("    " + " ".join("%s = 0;" % v for v in accs))
        if preload:
            out.append("    " + preload)
        out.extend("    " + l for l in em.host)
        out.append("#endif")
        if tail:
            out.append("    " + tail)
        if epilogue:
            out.append("    " + epilogue)
        out.append("}")
        out.append("")

    # ---- t = a * b ---------------------------------------------------------------------------
    em = Emit()
    a = ["a.l[%d]" % j for j in range(8)]
    b = ["b.l[%d]" % i for i in range(8)]
    # operands must be plain identifiers inside asm constraints: copy to locals
    pro = " ".join("const uint32_t a%d = a.l[%d], b%d = b.l[%d];" % (j, j, j, j) for j in range(8))
    a = ["a%d" % j for j in range(8)]
    b = ["b%d" % i for i in range(8)]
    schoolbook(em, 8, 8, a, b, lambda i, j: True)
    merge(em, 0, 15, lambda k: "t[%d]" % k)
    function("fr_gen_mulwide", "uint32_t* t, const Fr& a, const Fr& b", em, prologue=pro)

    # ---- Barrett quotient: q3 = (q1 * MU, columns >= 7) >> 288, q1 = t[7..15] ---------------------
    em = Emit()
    mu = limbs(MU, 9)
    assert MU >> 288 == 0
    pro = " ".join("const uint32_t y%d = t[%d];" % (i, 7 + i) for i in range(9))
    q1 = ["y%d" % i for i in range(9)]
    schoolbook(em, 9, 9, mu, q1, lambda i, j: i + j >= 7)
    merge(em, 7, 17, lambda k: ("q3[%d]" % (k - 9)) if k >= 9 else None)
    function("fr_gen_qmu", "uint32_t* q3, const uint32_t* t", em, prologue=pro)

    # ---- r = (q3 * P) mod 2^256 -----------------------------------------------------------------
    em = Emit()
    pl = limbs(P, 8)
    pro = " ".join("const uint32_t y%d = q3[%d];" % (i, i) for i in range(8))
    q3 = ["y%d" % i for i in range(8)]
    schoolbook(em, 8, 8, pl, q3, lambda i, j: i + j <= 7, limit=7)
    merge(em, 0, 7, lambda k: "r[%d]" % k)
    function("fr_gen_qp", "uint32_t* r, const uint32_t* q3", em, prologue=pro)

    # ---- T[0..16] = sum_k s[k] * c[k], N = 2..7 ---------------------------------------------------
    for N in range(2, 8):
        em = Emit()
        pro = " ".join("const uint32_t s%d_%d = s[%d].l[%d], c%d_%d = c[%d].l[%d];" % (k, j, k, j, k, j, k, j)
                       for k in range(N) for j in range(8))
        for k in range(N):
            schoolbook(em, 8, 8, ["s%d_%d" % (k, j) for j in range(8)], ["c%d_%d" % (k, i) for i in range(8)],
                       lambda i, j: True)
        merge(em, 0, 16, lambda k: "T[%d]" % k)
        function("fr_gen_dot%d" % N, "uint32_t* T, const Fr* s, const Fr* c", em, prologue=pro)

    # ---- Montgomery reduction: r[0..8] = (T + ((T mod R) * NP mod R) * P) / R, R = 2^256 ------------
    em = Emit()
    NP = (-pow(P, -1, 1 << 256)) % (1 << 256)
    npl = limbs(NP, 8)
    pro = " ".join("const uint32_t y%d = T[%d];" % (i, i) for i in range(8))
    schoolbook(em, 8, 8, npl, ["y%d" % i for i in range(8)], lambda i, j: i + j <= 7, limit=7)
    merge(em, 0, 7, lambda k: "m%d" % k)
    function("fr_gen_mlow", "uint32_t* m, const uint32_t* T", em, prologue=pro,
             epilogue=None, locals_="uint32_t " + ", ".join("m%d" % k for k in range(8)) + ";",
             tail=" ".join("m[%d] = m%d;" % (k, k) for k in range(8)))
    em = Emit()
    pro = " ".join("const uint32_t y%d = m[%d];" % (i, i) for i in range(8))
    # the accumulators start as T: e_k = T[k]
    for k in range(17):
        em.touch("e%d" % k)
        em.live.add("e%d" % k)
    load = " ".join("e%d = T[%d];" % (k, k) for k in range(17))
    schoolbook(em, 8, 8, pl, ["y%d" % i for i in range(8)], lambda i, j: True)
    merge(em, 0, 16, lambda k: ("r[%d]" % (k - 8)) if k >= 8 else None)
    function("fr_gen_mpadd", "uint32_t* r, const uint32_t* m, const uint32_t* T", em, prologue=pro, preload=load)

    path = os.path.join(os.path.dirname(__file__), "..", "csrc", "fr_mulgen.inc")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    gen()
