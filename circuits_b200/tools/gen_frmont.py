"""Generate csrc/fr_montgen.inc: Montgomery-domain products for the dependency chains of the device code.

    python -m circuits_b200.tools.gen_frmont

Every witness entry has to be stored as a canonical residue, which is why fr_mul (fr.cuh) is a Barrett
multiplication: 312 SASS instructions, all of them on the critical path of whatever consumes the product.  The
long chains of the circuit (the 65 rounds of a Poseidon permutation, the 253 doublings of the EdDSA ladder) do
not need canonical values to continue, only to be STORED.  The routines here keep a chain in the Montgomery
domain with lazily reduced values,
    x~ = x * 2^256 mod p,   0 <= x~ < 2p,
so that one product costs 176 instructions and nothing else; the canonical value of an entry that has to be
stored is recovered off the chain by one reduction (fr_mont_redc, ~115 instructions, independent work that
fills issue slots the chain leaves empty).

Construction (radix 2^32, p < 2^254): the running sum S = A + 2^32 B is held in two limb arrays, A aligned at
limb 0 and B at limb 1.  A product x_j * y lands on the 64-bit window (j, j+1): even j go to A, odd j to B, so a
row of the schoolbook product is two carry chains of `mad.lo.cc / madc.hi.cc` pairs, which ptxas fuses into
IMAD.WIDE.U32 with carry.  After the reduction row (m = A[0] * (-p^-1) mod 2^32, S += m p) limb 0 is zero and the
frame moves up one limb: the old B becomes the new A, A[2..] the new B, and the one limb in between, A[1], is
folded into the new A[0] by a single add whose carry is consumed by the first chain of the next row.  No carry
is ever resolved by a ripple.  Bounds: with x < 2^255 and y, p < 2^254 the sum stays below 2^288, so A has 9
limbs, B 8, and the top limb of a chain absorbs its carry; the dot products (up to 7 terms) get one more limb.

The generator emits each routine as ONE asm block (temporaries are PTX registers local to the block) and runs the
same instruction list through a 32-bit interpreter against big-integer arithmetic before writing anything.
"""
import os
import random

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = 1 << 256
NP0 = (-pow(P, -1, 1 << 32)) % (1 << 32)
M32 = 0xFFFFFFFF


def limbs(v, n=8):
    return [(v >> (32 * i)) & M32 for i in range(n)]


class Asm:
    """Instruction list over virtual 32-bit registers; prints PTX, runs on Python ints."""

    def __init__(self):
        self.ops = []          # (opcode, dst, srcs)
        self.n = 0
        self.inputs = []       # names bound to asm input operands
        self.outputs = []      # (operand name, virtual register)

    def new(self):
        self.n += 1
        return "t%d" % self.n

    def op(self, opcode, srcs, dst=None):
        d = dst or self.new()
        self.ops.append((opcode, d, list(srcs)))
        return d

    # -- interpreter ---------------------------------------------------------------------------
    def run(self, env):
        reg = dict(env)
        cf = 0

        def val(s):
            return s if isinstance(s, int) else reg[s]
        for opcode, d, srcs in self.ops:
            v = [val(s) for s in srcs]
            base = opcode.split(".")[0]
            cc_out = ".cc" in opcode
            cin = cf if base in ("madc", "addc", "subc") else 0
            if base in ("mad", "madc"):
                prod = v[0] * v[1]
                part = (prod & M32) if ".lo" in opcode else (prod >> 32)
                s = part + v[2] + cin
            elif base in ("add", "addc"):
                s = v[0] + v[1] + cin
            elif base in ("sub", "subc"):
                s = v[0] - v[1] - cin
            elif base == "mul":
                prod = v[0] * v[1]
                s = (prod & M32) if ".lo" in opcode else (prod >> 32)
            elif base == "mov":
                s = v[0]
            elif base == "selp":           # selp d, a, b, c: c != 0 ? a : b   (c is a 0 / all-ones register here)
                s = v[0] if v[2] != 0 else v[1]
            else:
                raise ValueError(opcode)
            if base in ("sub", "subc"):
                if cc_out:
                    cf = 1 if s < 0 else 0
                reg[d] = s & M32
            else:
                if cc_out:
                    cf = (s >> 32) & 1
                    assert s >> 33 == 0
                else:
                    # an instruction that drops its carry must not have one (the bound arguments of the module doc)
                    assert base in ("mul", "mov", "selp") or s >> 32 == 0, (opcode, d, hex(s))
                reg[d] = s & M32
        return reg

    # -- PTX -----------------------------------------------------------------------------------
    def ptx(self):
        opnum = {}
        for k, (nm, _) in enumerate(self.outputs):
            opnum[nm] = k
        for k, nm in enumerate(self.inputs):
            opnum[nm] = len(self.outputs) + k
        outreg = {vr: nm for nm, vr in self.outputs}

        def ref(s):
            if isinstance(s, int):
                return "0x%08x" % s
            if s in opnum:
                return "%%%d" % opnum[s]
            if s in outreg:
                return "%%%d" % opnum[outreg[s]]
            return s
        temps = sorted({d for _, d, _ in self.ops if d not in outreg and d not in opnum}, key=lambda x: int(x[1:]))
        lines = ["{"]
        for k in range(0, len(temps), 16):
            lines.append(".reg .u32 %s;" % ", ".join(temps[k:k + 16]))
        if any(o[0].startswith("selp") for o in self.ops):
            lines.append(".reg .pred pq;")
        for opcode, d, srcs in self.ops:
            if opcode == "selp":
                lines.append("setp.ne.u32 pq, %s, 0;" % ref(srcs[2]))
                lines.append("selp.u32 %s, %s, %s, pq;" % (ref(d), ref(srcs[0]), ref(srcs[1])))
            else:
                lines.append("%s.u32 %s, %s;" % (opcode, ref(d), ", ".join(ref(s) for s in srcs)))
        lines.append("}")
        return lines


def window_chain(asm, arr, xs, y, start, carry_in=False, top_carry=1):
    """Adds x * y for the multiplicand limbs xs = [(window offset, limb)] into arr (list of registers / None = 0),
    windows (start + off, start + off + 1), as one carry chain.  The carry out of the last window ripples into
    `top_carry` further limbs (0: the bound argument says there is none)."""
    first = not carry_in
    for off, x in xs:
        k = start + off
        for part in ("lo", "hi"):
            idx = k if part == "lo" else k + 1
            while len(arr) <= idx:
                arr.append(None)
            last = (off, x) == xs[-1] and part == "hi" and top_carry == 0
            opc = ("mad" if first else "madc") + "." + part + ("" if last else ".cc")
            arr[idx] = asm.op(opc, [x, y, arr[idx] if arr[idx] is not None else 0])
            first = False
    k = start + xs[-1][0] + 2
    for t in range(top_carry):
        while len(arr) <= k + t:
            arr.append(None)
        opc = "addc" + ("" if t == top_carry - 1 else ".cc")
        arr[k + t] = asm.op(opc, [arr[k + t] if arr[k + t] is not None else 0, 0])


def build(terms, init=None, wide=False):
    """terms: list of (xs, ys) -- 8 registers / ints each; init: 8 registers to start S from (REDC of a value).
    wide: one more limb per array (dot products of up to 7 terms)."""
    asm = Asm()
    A, B = [], []
    if init is not None:
        A = list(init)
    p = limbs(P)
    topA = 2 if wide else 1            # limbs above the 8 a chain on A may carry into
    topB = 1 if wide else 0
    for i in range(8):
        if i > 0:
            # frame moves up one limb: new A = old B, old A[1] folded in; new B = old A[2..]
            a1 = A[1] if len(A) > 1 else None
            b0 = B[0] if B else None
            nB = A[2:]
            nA = list(B)
            while len(nA) < 1:
                nA.append(None)
            if a1 is not None and b0 is not None:
                nA[0] = asm.op("add.cc", [b0, a1])
                carry = True
            else:
                nA[0] = b0 if b0 is not None else a1
                carry = False
            A, B = nA, nB
        else:
            carry = False
        # product rows of limb i: odd multiplicand limbs first (their chain starts at limb 1 and takes the fold's carry)
        for k, (xs, ys) in enumerate(terms):
            window_chain(asm, B, [(0, xs[1]), (2, xs[3]), (4, xs[5]), (6, xs[7])], ys[i], 0, carry_in=carry, top_carry=topB)
            carry = False
            window_chain(asm, A, [(0, xs[0]), (2, xs[2]), (4, xs[4]), (6, xs[6])], ys[i], 0, top_carry=topA)
        if carry:
            # no product row consumed the fold's carry (pure reduction): the odd reduction chain below takes it;
            # mul.lo does not touch the carry flag
            m = asm.op("mul.lo", [A[0], NP0])
            window_chain(asm, B, [(0, p[1]), (2, p[3]), (4, p[5]), (6, p[7])], m, 0, carry_in=True, top_carry=topB)
            window_chain(asm, A, [(0, p[0]), (2, p[2]), (4, p[4]), (6, p[6])], m, 0, top_carry=topA)
        else:
            m = asm.op("mul.lo", [A[0], NP0])
            window_chain(asm, A, [(0, p[0]), (2, p[2]), (4, p[4]), (6, p[6])], m, 0, top_carry=topA)
            window_chain(asm, B, [(0, p[1]), (2, p[3]), (4, p[5]), (6, p[7])], m, 0, top_carry=topB)
    # result = (A >> 32) + B, 8 limbs (everything above is zero by the bounds)
    res = []
    for k in range(8):
        x = A[k + 1] if len(A) > k + 1 else None
        y = B[k] if len(B) > k else None
        opc = ("add" if k == 0 else "addc") + (".cc" if k < 7 else "")
        res.append(asm.op(opc, [x if x is not None else 0, y if y is not None else 0]))
    extra = [r for r in (A[9:] + B[8:]) if r is not None]
    return asm, res, extra


def cond_sub(asm, r, k):
    """r - k when that is not negative, else r (k: 8 constant limbs)."""
    d = []
    for i in range(8):
        d.append(asm.op(("sub" if i == 0 else "subc") + ".cc", [r[i], k[i]]))
    brw = asm.op("subc", [0, 0])           # 0 or 0xffffffff
    return [asm.op("selp", [r[i], d[i], brw]) for i in range(8)]


def check(name, asm, res, extra, ins, ref, trials, gen_inputs):
    rng = random.Random(1)
    for t in range(trials):
        vals = gen_inputs(rng, t)
        env = {}
        for nm, v in zip(ins, vals):
            for j, l in enumerate(limbs(v)):
                env["%s%d" % (nm, j)] = l
        reg = asm.run(env)
        got = sum(reg[r] << (32 * k) for k, r in enumerate(res)) if isinstance(res[0], str) else None
        for e in extra:
            assert reg[e] == 0, (name, "limb above the result is not zero")
        ref(vals, got)
    print("  %-14s %4d PTX instructions, checked on %d inputs" % (name, len(asm.ops), trials))


def emit(out, fname, sig, asm, res, in_names):
    for k, r in enumerate(res):
        asm.outputs.append(("r%d" % k, r))
    for nm in in_names:
        for j in range(8):
            asm.inputs.append("%s%d" % (nm, j))
    body = asm.ptx()
    out.append("__device__ __forceinline__ void %s(%s) {" % (fname, sig))
    out.append("    asm(")
    for ln in body:
        out.append('        "%s\\n\\t"' % ln)
    outs = ", ".join('"=r"(r.l[%d])' % k for k in range(8))
    inl = []
    for nm in in_names:
        for j in range(8):
            inl.append('"r"(%s)' % (("%s.l[%d]" % (nm, j)) if "[" not in nm else nm))
    out.append("        : %s" % outs)
    out.append("        : %s);" % ", ".join(inl))
    out.append("}")
    out.append("")


def gen():
    out = ["// GENERATED by circuits_b200/tools/gen_frmont.py -- do not edit.",
           "// Montgomery-domain products on lazily reduced values (0 <= x~ < 2p); see the generator for the construction.",
           "#pragma once", ""]
    two_p = 2 * P
    rinv = pow(R, -1, P)
    print("gen_frmont:")

    def edge(rng, t, hi):
        if t == 0:
            return hi - 1
        if t == 1:
            return 0
        if t == 2:
            return P
        if t == 3:
            return (1 << 255) - 1 if hi > (1 << 254) else P - 1
        return rng.randrange(hi)

    # ---- r = a b / R (mod p), a, b < 2^255 -> r < 2p ----------------------------------------------
    a = ["a%d" % j for j in range(8)]
    b = ["b%d" % j for j in range(8)]
    asm, res, extra = build([(a, b)])

    def ref_mul(vals, got):
        x, y = vals
        assert got % P == x * y * rinv % P and got < R and (got < two_p or x >= two_p or y >= two_p)
    check("fr_mont_mul", asm, res, extra, ["a", "b"], ref_mul, 3000,
          lambda rng, t: (edge(rng, t % 7, 1 << 255), edge(rng, (t // 7) % 7, 1 << 255)))
    emit(out, "fr_mont_mul", "Fr& r, const Fr& a, const Fr& b", asm, res, ["a", "b"])

    # ---- r = a / R (mod p), canonical, a < 2^256 -------------------------------------------------------
    asm, res, extra = build([], init=a)
    res = cond_sub(asm, res, limbs(P))

    def ref_redc(vals, got):
        (x,) = vals
        assert got == x * rinv % P
    check("fr_mont_redc", asm, res, extra, ["a"], ref_redc, 3000, lambda rng, t: (edge(rng, t % 7, 1 << 256) if t % 7 != 3 else 2 * P - 1,))
    emit(out, "fr_mont_redc", "Fr& r, const Fr& a", asm, res, ["a"])

    # ---- r = sum_k s[k] c[k] / R, s[k] < 2^255, c[k] < p: r < (0.38 N + 1) p ---------------------------
    for N in range(2, 8):
        terms = [(["s%d_%d" % (k, j) for j in range(8)], ["c%d_%d" % (k, j) for j in range(8)]) for k in range(N)]
        asm, res, extra = build(terms, wide=True)
        names = []
        for k in range(N):
            names += ["s%d_" % k, "c%d_" % k]

        def ref_dot(vals, got, N=N):
            want = sum(vals[2 * k] * vals[2 * k + 1] for k in range(N)) * rinv % P
            tight = all(vals[2 * k] < two_p for k in range(N))
            assert got % P == want and got < R and (not tight or got < (0.38 * N + 1.001) * P)

        def gin(rng, t, N=N):
            v = []
            for k in range(N):
                v.append(edge(rng, t % 5, 1 << 255))
                v.append(edge(rng, (t // 5) % 5, P))
            return v
        check("fr_mont_dot%d" % N, asm, res, extra, names, ref_dot, 600, gin)
        # operands: s[k].l[j], c[k].l[j]
        for k, r in enumerate(res):
            asm.outputs.append(("r%d" % k, r))
        inl = []
        for k in range(N):
            for j in range(8):
                asm.inputs.append("s%d_%d" % (k, j))
                inl.append('"r"(s[%d].l[%d])' % (k, j))
            for j in range(8):
                asm.inputs.append("c%d_%d" % (k, j))
                inl.append('"r"(c[%d].l[%d])' % (k, j))
        body = asm.ptx()
        out.append("__device__ __forceinline__ void fr_mont_dot%d(Fr& r, const Fr* s, const Fr* c) {" % N)
        out.append("    asm(")
        for ln in body:
            out.append('        "%s\\n\\t"' % ln)
        out.append("        : %s" % ", ".join('"=r"(r.l[%d])' % k for k in range(8)))
        out.append("        : %s);" % ", ".join(inl))
        out.append("}")
        out.append("")

    # ---- r = a - 2p when a >= 2p (lazy domain after an addition: a < 4p -> r < 2p) ---------------------
    asm = Asm()
    res = cond_sub(asm, a, limbs(two_p))

    def ref_fold(vals, got):
        (x,) = vals
        assert got == (x - two_p if x >= two_p else x)
    check("fr_lazy_fold", asm, res, [], ["a"], ref_fold, 2000, lambda rng, t: (edge(rng, t % 7, 1 << 256) if t % 3 else two_p + (t % 5) - 2,))
    emit(out, "fr_lazy_fold", "Fr& r, const Fr& a", asm, res, ["a"])

    path = os.path.join(os.path.dirname(__file__), "..", "csrc", "fr_montgen.inc")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    gen()
