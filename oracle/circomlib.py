"""circomlib 0.5.2 templates restated for the oracle (test infrastructure).

circomlib is an npm dependency of the reference that is not vendored under
/root/reference (package.json:27 `"circomlib": "^0.5.2"`, lock pins 0.5.2).
Each function below restates one upstream template (`circuits/<file>.circom`)
from its published source: same signal declaration order, same `<==`/`<--`
assignments, same `===` checks.  The reference call sites are cited per
function.  Inputs are `Sig` objects (wired signals) or ints (expressions).
"""
from .circuit import Sig, val
from .field import P, inv, div, is_neg, HALF
from .poseidon_constants import poseidon_params


# ---------------------------------------------------------------------------
# bitify.circom  (used throughout, e.g. src/decode-tx.circom:96)
# ---------------------------------------------------------------------------
def Num2Bits(c, n, x):
    in_ = c.inp("in", x)
    v = in_.v
    out = []
    lc = 0
    for i in range(n):
        b = (v >> i) & 1
        out.append(c.inp("out[%d]" % i, b))
        lc += b << i
    c.check(lc, v, "Num2Bits(%d)" % n)
    return out


def Bits2Num(c, bits):
    ins = c.inps("in", bits)
    lc = 0
    for i, b in enumerate(ins):
        lc += b.v << i
    return c.inp("out", lc)


def Num2Bits_strict(c, x):
    in_ = c.inp("in", x)
    out = c.sigs("out", 254)
    ac = c.sub("aliasCheck")
    n2b = c.sub("n2b")
    bits = Num2Bits(n2b, 254, in_)
    for i in range(254):
        c.set(out[i], bits[i])
    AliasCheck(ac, bits)
    return out


# ---------------------------------------------------------------------------
# aliascheck.circom / compconstant.circom
# ---------------------------------------------------------------------------
def AliasCheck(c, bits):
    ins = c.inps("in", bits)
    o = CompConstant(c.sub("compConstant"), P - 1, ins)
    c.check(o.v, 0, "AliasCheck")


def CompConstant(c, ct, bits):
    ins = c.inps("in", bits)
    out = c.sig("out")
    parts = c.sigs("parts", 127)
    sout = c.sig("sout")
    b = (1 << 128) - 1
    a = 1
    e = 1
    total = 0
    for i in range(127):
        clsb = (ct >> (i * 2)) & 1
        cmsb = (ct >> (i * 2 + 1)) & 1
        slsb = ins[i * 2].v
        smsb = ins[i * 2 + 1].v
        if cmsb == 0 and clsb == 0:
            pv = -b * smsb * slsb + b * smsb + b * slsb
        elif cmsb == 0 and clsb == 1:
            pv = a * smsb * slsb - a * slsb + b * smsb - a * smsb + a
        elif cmsb == 1 and clsb == 0:
            pv = b * smsb * slsb - a * smsb + a
        else:
            pv = -a * smsb * slsb + a
        c.set(parts[i], pv)
        total += parts[i].v
        b -= e
        a += e
        e *= 2
    c.set(sout, total)
    bits135 = Num2Bits(c.sub("num2bits"), 135, sout)
    c.set(out, bits135[127])
    return out


# ---------------------------------------------------------------------------
# comparators.circom
# ---------------------------------------------------------------------------
def IsZero(c, x):
    in_ = c.inp("in", x)
    out = c.sig("out")
    iv = inv(in_.v)
    inv_ = c.inp("inv", iv)
    c.set(out, -in_.v * inv_.v + 1)
    c.check(in_.v * out.v, 0, "IsZero")
    return out


def IsEqual(c, a, b):
    i0 = c.inp("in[0]", a)
    i1 = c.inp("in[1]", b)
    out = c.sig("out")
    o = IsZero(c.sub("isz"), i1.v - i0.v)
    c.set(out, o)
    return out


def ForceEqualIfEnabled(c, enabled, a, b):
    en = c.inp("enabled", enabled)
    i0 = c.inp("in[0]", a)
    i1 = c.inp("in[1]", b)
    o = IsZero(c.sub("isz"), i1.v - i0.v)
    c.check((1 - o.v) * en.v, 0, "ForceEqualIfEnabled")


def LessThan(c, n, a, b):
    i0 = c.inp("in[0]", a)
    i1 = c.inp("in[1]", b)
    out = c.sig("out")
    bits = Num2Bits(c.sub("n2b"), n + 1, i0.v + (1 << n) - i1.v)
    c.set(out, 1 - bits[n].v)
    return out


def GreaterEqThan(c, n, a, b):
    i0 = c.inp("in[0]", a)
    i1 = c.inp("in[1]", b)
    out = c.sig("out")
    o = LessThan(c.sub("lt"), n, i1, i0.v + 1)
    c.set(out, o)
    return out


# ---------------------------------------------------------------------------
# gates.circom
# ---------------------------------------------------------------------------
def XOR(c, a, b):
    a_ = c.inp("a", a)
    b_ = c.inp("b", b)
    return c.inp("out", a_.v + b_.v - 2 * a_.v * b_.v)


def AND(c, a, b):
    a_ = c.inp("a", a)
    b_ = c.inp("b", b)
    return c.inp("out", a_.v * b_.v)


def MultiAND(c, xs):
    n = len(xs)
    ins = c.inps("in", xs)
    out = c.sig("out")
    if n == 1:
        c.set(out, ins[0])
    elif n == 2:
        o = AND(c.sub("and1"), ins[0], ins[1])
        c.set(out, o)
    else:
        and2 = c.sub("and2")
        n1 = n // 2
        a0 = MultiAND(c.sub("ands[0]"), ins[:n1])
        a1 = MultiAND(c.sub("ands[1]"), ins[n1:])
        o = AND(and2, a0, a1)
        c.set(out, o)
    return out


# ---------------------------------------------------------------------------
# mux1..mux4.circom
# ---------------------------------------------------------------------------
def MultiMux1(c, cs, s):
    n = len(cs)
    cc = [[c.inp("c[%d][%d]" % (i, k), cs[i][k]) for k in range(2)] for i in range(n)]
    s_ = c.inp("s", s)
    out = []
    for i in range(n):
        out.append(c.inp("out[%d]" % i, (cc[i][1].v - cc[i][0].v) * s_.v + cc[i][0].v))
    return out


def Mux1(c, c0, c1, s):
    cc = [c.inp("c[0]", c0), c.inp("c[1]", c1)]
    s_ = c.inp("s", s)
    out = c.sig("out")
    o = MultiMux1(c.sub("mux"), [cc], s_)
    c.set(out, o[0])
    return out


def MultiMux2(c, cs, s):
    n = len(cs)
    cc = [[c.inp("c[%d][%d]" % (i, k), cs[i][k]) for k in range(4)] for i in range(n)]
    ss = c.inps("s", s)
    out = c.sigs("out", n)
    a10 = c.sigs("a10", n)
    a1 = c.sigs("a1", n)
    a0 = c.sigs("a0", n)
    a = c.sigs("a", n)
    s10 = c.inp("s10", ss[1].v * ss[0].v)
    for i in range(n):
        k = [x.v for x in cc[i]]
        c.set(a10[i], (k[3] - k[2] - k[1] + k[0]) * s10.v)
        c.set(a1[i], (k[2] - k[0]) * ss[1].v)
        c.set(a0[i], (k[1] - k[0]) * ss[0].v)
        c.set(a[i], cc[i][0])
        c.set(out[i], a10[i].v + a1[i].v + a0[i].v + a[i].v)
    return out


def Mux2(c, cs, s):
    cc = c.inps("c", cs)
    ss = c.inps("s", s)
    out = c.sig("out")
    o = MultiMux2(c.sub("mux"), [cc], ss)
    c.set(out, o[0])
    return out


def MultiMux3(c, cs, s):
    n = len(cs)
    cc = [[c.inp("c[%d][%d]" % (i, k), cs[i][k]) for k in range(8)] for i in range(n)]
    ss = c.inps("s", s)
    out = c.sigs("out", n)
    a210 = c.sigs("a210", n)
    a21 = c.sigs("a21", n)
    a20 = c.sigs("a20", n)
    a2 = c.sigs("a2", n)
    a10 = c.sigs("a10", n)
    a1 = c.sigs("a1", n)
    a0 = c.sigs("a0", n)
    a = c.sigs("a", n)
    s10 = c.inp("s10", ss[1].v * ss[0].v)
    for i in range(n):
        k = [x.v for x in cc[i]]
        c.set(a210[i], (k[7] - k[6] - k[5] + k[4] - k[3] + k[2] + k[1] - k[0]) * s10.v)
        c.set(a21[i], (k[6] - k[4] - k[2] + k[0]) * ss[1].v)
        c.set(a20[i], (k[5] - k[4] - k[1] + k[0]) * ss[0].v)
        c.set(a2[i], k[4] - k[0])
        c.set(a10[i], (k[3] - k[2] - k[1] + k[0]) * s10.v)
        c.set(a1[i], (k[2] - k[0]) * ss[1].v)
        c.set(a0[i], (k[1] - k[0]) * ss[0].v)
        c.set(a[i], cc[i][0])
        c.set(out[i], (a210[i].v + a21[i].v + a20[i].v + a2[i].v) * ss[2].v
              + (a10[i].v + a1[i].v + a0[i].v + a[i].v))
    return out


def Mux3(c, cs, s):
    cc = c.inps("c", cs)
    ss = c.inps("s", s)
    out = c.sig("out")
    o = MultiMux3(c.sub("mux"), [cc], ss)
    c.set(out, o[0])
    return out


def MultiMux4(c, cs, s):
    n = len(cs)
    cc = [[c.inp("c[%d][%d]" % (i, k), cs[i][k]) for k in range(16)] for i in range(n)]
    ss = c.inps("s", s)
    out = c.sigs("out", n)
    names = ["a3210", "a321", "a320", "a310", "a32", "a31", "a30", "a3",
             "a210", "a21", "a20", "a10", "a2", "a1", "a0", "a"]
    A = {nm: c.sigs(nm, n) for nm in names}
    s10 = c.inp("s10", ss[1].v * ss[0].v)
    s20 = c.inp("s20", ss[2].v * ss[0].v)
    s21 = c.inp("s21", ss[2].v * ss[1].v)
    s210 = c.inp("s210", s21.v * ss[0].v)
    for i in range(n):
        k = [x.v for x in cc[i]]
        c.set(A["a3210"][i], (k[15] - k[14] - k[13] + k[12] - k[11] + k[10] + k[9] - k[8]
                              - k[7] + k[6] + k[5] - k[4] + k[3] - k[2] - k[1] + k[0]) * s210.v)
        c.set(A["a321"][i], (k[14] - k[12] - k[10] + k[8] - k[6] + k[4] + k[2] - k[0]) * s21.v)
        c.set(A["a320"][i], (k[13] - k[12] - k[9] + k[8] - k[5] + k[4] + k[1] - k[0]) * s20.v)
        c.set(A["a310"][i], (k[11] - k[10] - k[9] + k[8] - k[3] + k[2] + k[1] - k[0]) * s10.v)
        c.set(A["a32"][i], (k[12] - k[8] - k[4] + k[0]) * ss[2].v)
        c.set(A["a31"][i], (k[10] - k[8] - k[2] + k[0]) * ss[1].v)
        c.set(A["a30"][i], (k[9] - k[8] - k[1] + k[0]) * ss[0].v)
        c.set(A["a3"][i], k[8] - k[0])
        c.set(A["a210"][i], (k[7] - k[6] - k[5] + k[4] - k[3] + k[2] + k[1] - k[0]) * s210.v)
        c.set(A["a21"][i], (k[6] - k[4] - k[2] + k[0]) * s21.v)
        c.set(A["a20"][i], (k[5] - k[4] - k[1] + k[0]) * s20.v)
        c.set(A["a10"][i], (k[3] - k[2] - k[1] + k[0]) * s10.v)
        c.set(A["a2"][i], (k[4] - k[0]) * ss[2].v)
        c.set(A["a1"][i], (k[2] - k[0]) * ss[1].v)
        c.set(A["a0"][i], (k[1] - k[0]) * ss[0].v)
        c.set(A["a"][i], cc[i][0])
        hi = sum(A[nm][i].v for nm in names[:8])
        lo = sum(A[nm][i].v for nm in names[8:])
        c.set(out[i], hi * ss[3].v + lo)
    return out


def Mux4(c, cs, s):
    cc = c.inps("c", cs)
    ss = c.inps("s", s)
    out = c.sig("out")
    o = MultiMux4(c.sub("mux"), [cc], ss)
    c.set(out, o[0])
    return out


# ---------------------------------------------------------------------------
# switcher.circom
# ---------------------------------------------------------------------------
def Switcher(c, sel, L, R):
    sel_ = c.inp("sel", sel)
    L_ = c.inp("L", L)
    R_ = c.inp("R", R)
    outL = c.sig("outL")
    outR = c.sig("outR")
    aux = c.inp("aux", (R_.v - L_.v) * sel_.v)
    c.set(outL, aux.v + L_.v)
    c.set(outR, -aux.v + R_.v)
    return outL, outR


# ---------------------------------------------------------------------------
# poseidon.circom  (src/decode-tx.circom:275, src/lib/hash-state.circom:32)
# ---------------------------------------------------------------------------
def _Sigma(c, x):
    in_ = c.inp("in", x)
    out = c.sig("out")
    in2 = c.inp("in2", in_.v * in_.v)
    in4 = c.inp("in4", in2.v * in2.v)
    c.set(out, in4.v * in_.v)
    return out


def Poseidon(c, inputs):
    n = len(inputs)
    ins = c.inps("inputs", inputs)
    out = c.sig("out")
    t = n + 1
    C, M, rf, rp = poseidon_params(t)
    state = None
    for i in range(rf + rp):
        ark = c.sub("ark[%d]" % i)
        if i == 0:
            src = [0] + ins
        else:
            src = state
        a_in = ark.inps("in", src)
        a_out = [ark.inp("out[%d]" % j, a_in[j].v + C[t * i + j]) for j in range(t)]
        if i < rf // 2 or i >= rp + rf // 2:
            k = i if i < rf // 2 else i - rp
            mix = c.sub("mix[%d]" % i)
            sg = [_Sigma(c.sub("sigmaF[%d][%d]" % (k, j)), a_out[j]) for j in range(t)]
        else:
            k = i - rf // 2
            mix = c.sub("mix[%d]" % i)
            s0 = _Sigma(c.sub("sigmaP[%d]" % k), a_out[0])
            sg = [s0] + a_out[1:]
        m_in = mix.inps("in", sg)
        state = []
        for r in range(t):
            lc = 0
            for j in range(t):
                lc += M[r][j] * m_in[j].v
            state.append(mix.inp("out[%d]" % r, lc))
    c.set(out, state[0])
    return out


# ---------------------------------------------------------------------------
# smt/*.circom  (src/rollup-tx.circom:537,559; src/fee-tx.circom:97;
#                src/withdraw.circom:47)
# ---------------------------------------------------------------------------
def SMTHash1(c, key, value):
    k = c.inp("key", key)
    v = c.inp("value", value)
    out = c.sig("out")
    o = Poseidon(c.sub("h"), [k, v, 1])
    c.set(out, o)
    return out


def SMTHash2(c, L, R):
    L_ = c.inp("L", L)
    R_ = c.inp("R", R)
    out = c.sig("out")
    o = Poseidon(c.sub("h"), [L_, R_])
    c.set(out, o)
    return out


def SMTLevIns(c, enabled, siblings):
    n = len(siblings)
    en = c.inp("enabled", enabled)
    sib = c.inps("siblings", siblings)
    levIns = c.sigs("levIns", n)
    done = c.sigs("done", n - 1)
    isz = [IsZero(c.sub("isZero[%d]" % i), sib[i]) for i in range(n)]
    c.check((isz[n - 1].v - 1) * en.v, 0, "SMTLevIns last sibling")
    c.set(levIns[n - 1], 1 - isz[n - 2].v)
    c.set(done[n - 2], levIns[n - 1])
    for i in range(n - 2, 0, -1):
        c.set(levIns[i], (1 - done[i].v) * (1 - isz[i - 1].v))
        c.set(done[i - 1], levIns[i].v + done[i].v)
    c.set(levIns[0], 1 - done[0].v)
    return levIns


def SMTProcessorSM(c, xor, is0, levIns, fnc, prev):
    """prev = (top, old0, bot, new1, na, upd); returns the same tuple of st_*."""
    x = c.inp("xor", xor)
    z = c.inp("is0", is0)
    li = c.inp("levIns", levIns)
    f = c.inps("fnc", fnc)
    p_top = c.inp("prev_top", prev[0])
    p_old0 = c.inp("prev_old0", prev[1])
    p_bot = c.inp("prev_bot", prev[2])
    p_new1 = c.inp("prev_new1", prev[3])
    p_na = c.inp("prev_na", prev[4])
    p_upd = c.inp("prev_upd", prev[5])
    st_top = c.sig("st_top")
    st_old0 = c.sig("st_old0")
    st_bot = c.sig("st_bot")
    st_new1 = c.sig("st_new1")
    st_na = c.sig("st_na")
    st_upd = c.sig("st_upd")
    aux1 = c.inp("aux1", p_top.v * li.v)
    aux2 = c.inp("aux2", aux1.v * f[0].v)
    c.set(st_top, p_top.v - aux1.v)
    c.set(st_old0, aux2.v * z.v)
    c.set(st_new1, (aux2.v - st_old0.v + p_bot.v) * x.v)
    c.set(st_bot, (1 - x.v) * (aux2.v - st_old0.v + p_bot.v))
    c.set(st_upd, aux1.v - aux2.v)
    c.set(st_na, p_new1.v + p_old0.v + p_na.v + p_upd.v)
    return (st_top, st_old0, st_bot, st_new1, st_na, st_upd)


def SMTProcessorLevel(c, st, sibling, old1leaf, new1leaf, newlrbit, oldChild, newChild):
    st_top = c.inp("st_top", st[0])
    st_old0 = c.inp("st_old0", st[1])
    st_bot = c.inp("st_bot", st[2])
    st_new1 = c.inp("st_new1", st[3])
    c.inp("st_na", st[4])
    st_upd = c.inp("st_upd", st[5])
    oldRoot = c.sig("oldRoot")
    newRoot = c.sig("newRoot")
    sib = c.inp("sibling", sibling)
    o1 = c.inp("old1leaf", old1leaf)
    n1 = c.inp("new1leaf", new1leaf)
    lr = c.inp("newlrbit", newlrbit)
    oc = c.inp("oldChild", oldChild)
    nc = c.inp("newChild", newChild)
    aux = c.sigs("aux", 4)
    oldProofHash = c.sub("oldProofHash")
    newProofHash = c.sub("newProofHash")
    oldSwitcher = c.sub("oldSwitcher")
    newSwitcher = c.sub("newSwitcher")
    oL, oR = Switcher(oldSwitcher, lr, oc, sib)
    oph = SMTHash2(oldProofHash, oL, oR)
    c.set(aux[0], o1.v * (st_bot.v + st_new1.v + st_upd.v))
    c.set(oldRoot, aux[0].v + oph.v * st_top.v)
    c.set(aux[1], nc.v * (st_top.v + st_bot.v))
    c.set(aux[2], sib.v * st_top.v)
    nL, nR = Switcher(newSwitcher, lr, aux[1].v + n1.v * st_new1.v, aux[2].v + o1.v * st_new1.v)
    nph = SMTHash2(newProofHash, nL, nR)
    c.set(aux[3], nph.v * (st_top.v + st_bot.v + st_new1.v))
    c.set(newRoot, aux[3].v + n1.v * (st_old0.v + st_upd.v))
    return oldRoot, newRoot


def SMTProcessor(c, nLevels, oldRoot, siblings, oldKey, oldValue, isOld0, newKey, newValue, fnc):
    oldRoot_ = c.inp("oldRoot", oldRoot)
    newRoot = c.sig("newRoot")
    sib = c.inps("siblings", siblings)
    oldKey_ = c.inp("oldKey", oldKey)
    oldValue_ = c.inp("oldValue", oldValue)
    isOld0_ = c.inp("isOld0", isOld0)
    newKey_ = c.inp("newKey", newKey)
    newValue_ = c.inp("newValue", newValue)
    f = c.inps("fnc", fnc)
    enabled = c.inp("enabled", f[0].v + f[1].v - f[0].v * f[1].v)
    hash1Old = SMTHash1(c.sub("hash1Old"), oldKey_, oldValue_)
    hash1New = SMTHash1(c.sub("hash1New"), newKey_, newValue_)
    n2bOld_c = c.sub("n2bOld")
    n2bNew_c = c.sub("n2bNew")
    n2bOld = Num2Bits_strict(n2bOld_c, oldKey_)
    n2bNew = Num2Bits_strict(n2bNew_c, newKey_)
    levIns = SMTLevIns(c.sub("smtLevIns"), enabled, sib)
    xors = [XOR(c.sub("xors[%d]" % i), n2bOld[i], n2bNew[i]) for i in range(nLevels)]
    sm = []
    for i in range(nLevels):
        if i == 0:
            prev = (enabled, 0, 0, 0, 1 - enabled.v, 0)
        else:
            prev = sm[i - 1]
        sm.append(SMTProcessorSM(c.sub("sm[%d]" % i), xors[i], isOld0_, levIns[i], f, prev))
    last = sm[nLevels - 1]
    c.check(last[4].v + last[3].v + last[1].v + last[5].v, 1, "SMTProcessor sm final")
    levels = [None] * nLevels
    for i in range(nLevels - 1, -1, -1):
        if i == nLevels - 1:
            oc, nc = 0, 0
        else:
            oc, nc = levels[i + 1]
        levels[i] = SMTProcessorLevel(c.sub("levels[%d]" % i), sm[i], sib[i], hash1Old, hash1New,
                                      n2bNew[i], oc, nc)
    tL, tR = Switcher(c.sub("topSwitcher"), f[0].v * f[1].v, levels[0][0], levels[0][1])
    ForceEqualIfEnabled(c.sub("checkOldInput"), enabled, oldRoot_, tL)
    c.set(newRoot, enabled.v * (tR.v - oldRoot_.v) + oldRoot_.v)
    eq = IsEqual(c.sub("areKeyEquals"), oldKey_, newKey_)
    ok = MultiAND(c.sub("keysOk"), [1 - f[0].v, f[1], 1 - eq.v])
    c.check(ok.v, 0, "SMTProcessor keysOk")
    return newRoot


def SMTVerifierSM(c, is0, levIns, fnc, prev):
    """prev = (top, i0, iold, inew, na)."""
    z = c.inp("is0", is0)
    li = c.inp("levIns", levIns)
    f = c.inp("fnc", fnc)
    p_top = c.inp("prev_top", prev[0])
    p_i0 = c.inp("prev_i0", prev[1])
    p_iold = c.inp("prev_iold", prev[2])
    p_inew = c.inp("prev_inew", prev[3])
    p_na = c.inp("prev_na", prev[4])
    st_top = c.sig("st_top")
    st_i0 = c.sig("st_i0")
    st_iold = c.sig("st_iold")
    st_inew = c.sig("st_inew")
    st_na = c.sig("st_na")
    a = c.inp("prev_top_lev_ins", p_top.v * li.v)
    b = c.inp("prev_top_lev_ins_fnc", a.v * f.v)
    c.set(st_top, p_top.v - a.v)
    c.set(st_inew, a.v - b.v)
    c.set(st_iold, b.v * (1 - z.v))
    c.set(st_i0, a.v * z.v)
    c.set(st_na, p_na.v + p_inew.v + p_iold.v + p_i0.v)
    return (st_top, st_i0, st_iold, st_inew, st_na)


def SMTVerifierLevel(c, st, sibling, old1leaf, new1leaf, lrbit, child):
    st_top = c.inp("st_top", st[0])
    c.inp("st_i0", st[1])
    st_iold = c.inp("st_iold", st[2])
    st_inew = c.inp("st_inew", st[3])
    c.inp("st_na", st[4])
    root = c.sig("root")
    sib = c.inp("sibling", sibling)
    o1 = c.inp("old1leaf", old1leaf)
    n1 = c.inp("new1leaf", new1leaf)
    lr = c.inp("lrbit", lrbit)
    ch = c.inp("child", child)
    aux = c.sigs("aux", 2)
    proofHash = c.sub("proofHash")
    switcher = c.sub("switcher")
    sL, sR = Switcher(switcher, lr, ch, sib)
    ph = SMTHash2(proofHash, sL, sR)
    c.set(aux[0], ph.v * st_top.v)
    c.set(aux[1], o1.v * st_iold.v)
    c.set(root, aux[0].v + aux[1].v + n1.v * st_inew.v)
    return root


def SMTVerifier(c, nLevels, enabled, root, siblings, oldKey, oldValue, isOld0, key, value, fnc):
    en = c.inp("enabled", enabled)
    root_ = c.inp("root", root)
    sib = c.inps("siblings", siblings)
    oldKey_ = c.inp("oldKey", oldKey)
    oldValue_ = c.inp("oldValue", oldValue)
    isOld0_ = c.inp("isOld0", isOld0)
    key_ = c.inp("key", key)
    value_ = c.inp("value", value)
    fnc_ = c.inp("fnc", fnc)
    hash1Old = SMTHash1(c.sub("hash1Old"), oldKey_, oldValue_)
    hash1New = SMTHash1(c.sub("hash1New"), key_, value_)
    n2bOld_c = c.sub("n2bOld")
    n2bNew_c = c.sub("n2bNew")
    Num2Bits_strict(n2bOld_c, oldKey_)
    n2bNew = Num2Bits_strict(n2bNew_c, key_)
    levIns = SMTLevIns(c.sub("smtLevIns"), en, sib)
    sm = []
    for i in range(nLevels):
        if i == 0:
            prev = (en, 0, 0, 0, 1 - en.v)
        else:
            prev = sm[i - 1]
        sm.append(SMTVerifierSM(c.sub("sm[%d]" % i), isOld0_, levIns[i], fnc_, prev))
    last = sm[nLevels - 1]
    c.check(last[4].v + last[2].v + last[3].v + last[1].v, 1, "SMTVerifier sm final")
    levels = [None] * nLevels
    for i in range(nLevels - 1, -1, -1):
        child = 0 if i == nLevels - 1 else levels[i + 1]
        levels[i] = SMTVerifierLevel(c.sub("levels[%d]" % i), sm[i], sib[i], hash1Old, hash1New,
                                     n2bNew[i], child)
    eq = IsEqual(c.sub("areKeyEquals"), oldKey_, key_)
    ok = MultiAND(c.sub("keysOk"), [fnc_, 1 - isOld0_.v, eq, en])
    c.check(ok.v, 0, "SMTVerifier keysOk")
    ForceEqualIfEnabled(c.sub("checkRoot"), en, levels[0], root_)


# ---------------------------------------------------------------------------
# babyjub.circom / montgomery.circom
# ---------------------------------------------------------------------------
BJ_A = 168700
BJ_D = 168696
MONT_A = 168698   # 2(a+d)/(a-d)
MONT_B = 1        # 4/(a-d)
BASE8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
         16950150798460657717958625567821834550301663161624707787222815936182638968203)
SUBORDER = 2736030358979909402780800718157159386076813972158567259200215660948447373041


def BabyAdd(c, x1, y1, x2, y2):
    x1 = c.inp("x1", x1)
    y1 = c.inp("y1", y1)
    x2 = c.inp("x2", x2)
    y2 = c.inp("y2", y2)
    xout = c.sig("xout")
    yout = c.sig("yout")
    beta = c.inp("beta", x1.v * y2.v)
    gamma = c.inp("gamma", y1.v * x2.v)
    delta = c.inp("delta", (-BJ_A * x1.v + y1.v) * (x2.v + y2.v))
    tau = c.inp("tau", beta.v * gamma.v)
    c.set(xout, div(beta.v + gamma.v, 1 + BJ_D * tau.v))
    c.check((1 + BJ_D * tau.v) * xout.v, beta.v + gamma.v, "BabyAdd x")
    c.set(yout, div(delta.v + BJ_A * beta.v - gamma.v, 1 - BJ_D * tau.v))
    c.check((1 - BJ_D * tau.v) * yout.v, delta.v + BJ_A * beta.v - gamma.v, "BabyAdd y")
    return xout, yout


def BabyDbl(c, x, y):
    x_ = c.inp("x", x)
    y_ = c.inp("y", y)
    xout = c.sig("xout")
    yout = c.sig("yout")
    xo, yo = BabyAdd(c.sub("adder"), x_, y_, x_, y_)
    c.set(xout, xo)
    c.set(yout, yo)
    return xout, yout


def BabyCheck(c, x, y):
    x_ = c.inp("x", x)
    y_ = c.inp("y", y)
    x2 = c.inp("x2", x_.v * x_.v)
    y2 = c.inp("y2", y_.v * y_.v)
    c.check(BJ_A * x2.v + y2.v, 1 + BJ_D * x2.v * y2.v, "BabyCheck")


def Edwards2Montgomery(c, x, y):
    i0 = c.inp("in[0]", x)
    i1 = c.inp("in[1]", y)
    o0 = c.inp("out[0]", div(1 + i1.v, 1 - i1.v))
    o1 = c.inp("out[1]", div(o0.v, i0.v))
    c.check(o0.v * (1 - i1.v), 1 + i1.v, "Edwards2Montgomery u")
    c.check(o1.v * i0.v, o0.v, "Edwards2Montgomery v")
    return o0, o1


def Montgomery2Edwards(c, u, v):
    i0 = c.inp("in[0]", u)
    i1 = c.inp("in[1]", v)
    o0 = c.inp("out[0]", div(i0.v, i1.v))
    o1 = c.inp("out[1]", div(i0.v - 1, i0.v + 1))
    c.check(o0.v * i1.v, i0.v, "Montgomery2Edwards x")
    c.check(o1.v * (i0.v + 1), i0.v - 1, "Montgomery2Edwards y")
    return o0, o1


def MontgomeryAdd(c, p1, p2):
    a0 = c.inp("in1[0]", p1[0])
    a1 = c.inp("in1[1]", p1[1])
    b0 = c.inp("in2[0]", p2[0])
    b1 = c.inp("in2[1]", p2[1])
    out = c.sigs("out", 2)
    lam = c.inp("lamda", div(b1.v - a1.v, b0.v - a0.v))
    c.check(lam.v * (b0.v - a0.v), b1.v - a1.v, "MontgomeryAdd lamda")
    c.set(out[0], MONT_B * lam.v * lam.v - MONT_A - a0.v - b0.v)
    c.set(out[1], lam.v * (a0.v - out[0].v) - a1.v)
    return out[0], out[1]


def MontgomeryDouble(c, p):
    i0 = c.inp("in[0]", p[0])
    i1 = c.inp("in[1]", p[1])
    out = c.sigs("out", 2)
    lam = c.sig("lamda")
    x1_2 = c.inp("x1_2", i0.v * i0.v)
    num = 3 * x1_2.v + 2 * MONT_A * i0.v + 1
    c.set(lam, div(num, 2 * MONT_B * i1.v))
    c.check(lam.v * (2 * MONT_B * i1.v), num, "MontgomeryDouble lamda")
    c.set(out[0], MONT_B * lam.v * lam.v - MONT_A - 2 * i0.v)
    c.set(out[1], lam.v * (i0.v - out[0].v) - i1.v)
    return out[0], out[1]


# ---------------------------------------------------------------------------
# escalarmulany.circom
# ---------------------------------------------------------------------------
def Multiplexor2(c, sel, in0, in1):
    s = c.inp("sel", sel)
    a0 = c.inp("in[0][0]", in0[0])
    a1 = c.inp("in[0][1]", in0[1])
    b0 = c.inp("in[1][0]", in1[0])
    b1 = c.inp("in[1][1]", in1[1])
    o0 = c.inp("out[0]", (b0.v - a0.v) * s.v + a0.v)
    o1 = c.inp("out[1]", (b1.v - a1.v) * s.v + a1.v)
    return o0, o1


def BitElementMulAny(c, sel, dblIn, addIn):
    s = c.inp("sel", sel)
    d0 = c.inp("dblIn[0]", dblIn[0])
    d1 = c.inp("dblIn[1]", dblIn[1])
    a0 = c.inp("addIn[0]", addIn[0])
    a1 = c.inp("addIn[1]", addIn[1])
    dblOut = c.sigs("dblOut", 2)
    addOut = c.sigs("addOut", 2)
    doubler = c.sub("doubler")
    adder = c.sub("adder")
    selector = c.sub("selector")
    dd = MontgomeryDouble(doubler, (d0, d1))
    ad = MontgomeryAdd(adder, dd, (a0, a1))
    so = Multiplexor2(selector, s, (a0, a1), ad)
    c.set(dblOut[0], dd[0])
    c.set(dblOut[1], dd[1])
    c.set(addOut[0], so[0])
    c.set(addOut[1], so[1])
    return (dblOut[0], dblOut[1]), (addOut[0], addOut[1])


def SegmentMulAny(c, e, p):
    n = len(e)
    e_ = c.inps("e", e)
    p_ = c.inps("p", p)
    out = c.sigs("out", 2)
    dbl = c.sigs("dbl", 2)
    e2m_c = c.sub("e2m")
    m = Edwards2Montgomery(e2m_c, p_[0], p_[1])
    d, a = BitElementMulAny(c.sub("bits[0]"), e_[1], m, m)
    for i in range(1, n - 1):
        d, a = BitElementMulAny(c.sub("bits[%d]" % i), e_[i + 1], d, a)
    c.set(dbl[0], d[0])
    c.set(dbl[1], d[1])
    ed = Montgomery2Edwards(c.sub("m2e"), a[0], a[1])
    ex, ey = BabyAdd(c.sub("eadder"), ed[0], ed[1], -p_[0].v, p_[1])
    ls = Multiplexor2(c.sub("lastSel"), e_[0], (ex, ey), ed)
    c.set(out[0], ls[0])
    c.set(out[1], ls[1])
    return (out[0], out[1]), (dbl[0], dbl[1])


def EscalarMulAny(c, e, p):
    n = len(e)
    e_ = c.inps("e", e)
    p_ = c.inps("p", p)
    out = c.sigs("out", 2)
    nsegments = (n - 1) // 148 + 1
    nlast = n - (nsegments - 1) * 148
    zp = IsZero(c.sub("zeropoint"), p_[0])
    seg_out = []
    seg_dbl = []
    add_out = []
    for s in range(nsegments):
        nseg = 148 if s < nsegments - 1 else nlast
        seg_c = c.sub("segments[%d]" % s)
        bits = e_[s * 148: s * 148 + nseg]
        if s == 0:
            base = (p_[0].v + (BASE8[0] - p_[0].v) * zp.v,
                    p_[1].v + (BASE8[1] - p_[1].v) * zp.v)
            o, d = SegmentMulAny(seg_c, bits, base)
            seg_out.append(o)
            seg_dbl.append(d)
        else:
            dbl_c = c.sub("doublers[%d]" % (s - 1))
            m2e_c = c.sub("m2e[%d]" % (s - 1))
            add_c = c.sub("adders[%d]" % (s - 1))
            dd = MontgomeryDouble(dbl_c, seg_dbl[s - 1])
            ed = Montgomery2Edwards(m2e_c, dd[0], dd[1])
            o, d = SegmentMulAny(seg_c, bits, ed)
            seg_out.append(o)
            seg_dbl.append(d)
            first = seg_out[0] if s == 1 else add_out[s - 2]
            add_out.append(BabyAdd(add_c, first[0], first[1], o[0], o[1]))
    fin = seg_out[0] if nsegments == 1 else add_out[nsegments - 2]
    c.set(out[0], fin[0].v * (1 - zp.v))
    c.set(out[1], fin[1].v + (1 - fin[1].v) * zp.v)
    return out[0], out[1]


# ---------------------------------------------------------------------------
# escalarmulfix.circom
# ---------------------------------------------------------------------------
def WindowMulFix(c, in3, base):
    ins = c.inps("in", in3)
    b = c.inps("base", base)
    out = c.sigs("out", 2)
    out8 = c.sigs("out8", 2)
    mux_c = c.sub("mux")
    dbl2 = MontgomeryDouble(c.sub("dbl2"), b)
    pts = [(b[0], b[1]), dbl2]
    prev = dbl2
    for k in range(3, 9):
        prev = MontgomeryAdd(c.sub("adr%d" % k), b, prev)
        pts.append(prev)
    mo = MultiMux3(mux_c, [[pts[k][0] for k in range(8)], [pts[k][1] for k in range(8)]], ins)
    c.set(out8[0], pts[7][0])
    c.set(out8[1], pts[7][1])
    c.set(out[0], mo[0])
    c.set(out[1], mo[1])
    return (out[0], out[1]), (out8[0], out8[1])


def SegmentMulFix(c, e, base):
    nW = len(e) // 3
    e_ = c.inps("e", e)
    b = c.inps("base", base)
    out = c.sigs("out", 2)
    dbl = c.sigs("dbl", 2)
    m = Edwards2Montgomery(c.sub("e2m"), b[0], b[1])
    dblLast_c = c.sub("dblLast")
    win_out = []
    win_out8 = []
    cad = []
    dblLast = None
    for i in range(nW):
        w_c = c.sub("windows[%d]" % i)
        cad_c = c.sub("cadders[%d]" % i)
        wbase = m if i == 0 else win_out8[i - 1]
        o, o8 = WindowMulFix(w_c, e_[3 * i: 3 * i + 3], wbase)
        win_out.append(o)
        win_out8.append(o8)
        cin1 = m if i == 0 else cad[i - 1]
        if i < nW - 1:
            cad.append(MontgomeryAdd(cad_c, cin1, o8))
        else:
            dblLast = MontgomeryDouble(dblLast_c, o8)
            cad.append(MontgomeryAdd(cad_c, cin1, dblLast))
    add = []
    for i in range(nW):
        a_in1 = dblLast if i == 0 else add[i - 1]
        add.append(MontgomeryAdd(c.sub("adders[%d]" % i), a_in1, win_out[i]))
    ed = Montgomery2Edwards(c.sub("m2e"), add[nW - 1][0], add[nW - 1][1])
    ced = Montgomery2Edwards(c.sub("cm2e"), cad[nW - 1][0], cad[nW - 1][1])
    cx, cy = BabyAdd(c.sub("cAdd"), ed[0], ed[1], -ced[0].v, ced[1])
    c.set(out[0], cx)
    c.set(out[1], cy)
    c.set(dbl[0], win_out8[nW - 1][0])
    c.set(dbl[1], win_out8[nW - 1][1])
    return (out[0], out[1]), (dbl[0], dbl[1])


def EscalarMulFix(c, e, BASE):
    n = len(e)
    e_ = c.inps("e", e)
    out = c.sigs("out", 2)
    nsegments = (n - 1) // 246 + 1
    nlast = n - (nsegments - 1) * 249
    seg_out = []
    seg_dbl = []
    add_out = []
    for s in range(nsegments):
        nseg = 249 if s < nsegments - 1 else nlast
        nW = (nseg - 1) // 3 + 1
        seg_c = c.sub("segments[%d]" % s)
        bits = list(e_[s * 249: s * 249 + nseg]) + [0] * (nW * 3 - nseg)
        if s == 0:
            o, d = SegmentMulFix(seg_c, bits, BASE)
        else:
            m2e_c = c.sub("m2e[%d]" % (s - 1))
            add_c = c.sub("adders[%d]" % (s - 1))
            ed = Montgomery2Edwards(m2e_c, seg_dbl[s - 1][0], seg_dbl[s - 1][1])
            o, d = SegmentMulFix(seg_c, bits, ed)
            first = seg_out[0] if s == 1 else add_out[s - 2]
            add_out.append(BabyAdd(add_c, first[0], first[1], o[0], o[1]))
        seg_out.append(o)
        seg_dbl.append(d)
    fin = seg_out[0] if nsegments == 1 else add_out[nsegments - 2]
    c.set(out[0], fin[0])
    c.set(out[1], fin[1])
    return out[0], out[1]


# ---------------------------------------------------------------------------
# eddsaposeidon.circom   (src/rollup-tx.circom:472)
# ---------------------------------------------------------------------------
def EdDSAPoseidonVerifier(c, enabled, Ax, Ay, S, R8x, R8y, M):
    en = c.inp("enabled", enabled)
    Ax_ = c.inp("Ax", Ax)
    Ay_ = c.inp("Ay", Ay)
    S_ = c.inp("S", S)
    R8x_ = c.inp("R8x", R8x)
    R8y_ = c.inp("R8y", R8y)
    M_ = c.inp("M", M)
    sbits = Num2Bits(c.sub("snum2bits"), 253, S_)
    cc = CompConstant(c.sub("compConstant"), SUBORDER - 1, sbits + [0])
    c.check(cc.v * en.v, 0, "EdDSA S range")
    h = Poseidon(c.sub("hash"), [R8x_, R8y_, Ax_, Ay_, M_])
    hbits = Num2Bits_strict(c.sub("h2bits"), h)
    d1 = BabyDbl(c.sub("dbl1"), Ax_, Ay_)
    d2 = BabyDbl(c.sub("dbl2"), d1[0], d1[1])
    d3_c = c.sub("dbl3")
    d3 = BabyDbl(d3_c, d2[0], d2[1])
    isz = IsZero(c.sub("isZero"), d2[0])   # dbl3.x == dbl2.xout
    c.check(isz.v * en.v, 0, "EdDSA A is zero")
    mulAny = EscalarMulAny(c.sub("mulAny"), hbits, d3)
    right = BabyAdd(c.sub("addRight"), R8x_, R8y_, mulAny[0], mulAny[1])
    left = EscalarMulFix(c.sub("mulFix"), sbits, BASE8)
    ForceEqualIfEnabled(c.sub("eqCheckX"), en, left[0], right[0])
    ForceEqualIfEnabled(c.sub("eqCheckY"), en, left[1], right[1])


# ---------------------------------------------------------------------------
# pointbits.circom   (src/lib/utils-bjj.circom:46)
# ---------------------------------------------------------------------------
def fr_sqrt(n):
    """circomlib `sqrt()` function: Tonelli-Shanks, 0 for non-residues,
    result normalised to the non-negative (<= (p-1)/2) root."""
    n %= P
    if n == 0:
        return 0
    if pow(n, (P - 1) >> 1, P) != 1:
        return 0
    m = 28
    cst = 19103219067921713944291392827692070036145651957329286315305642004821462161904
    t0 = 81540058820840996586704275553141814055101440848469862132140264610111
    t = pow(n, t0, P)
    r = pow(n, (t0 + 1) >> 1, P)
    while r != 0 and t != 1:
        sq = t * t % P
        i = 1
        while sq != 1:
            i += 1
            sq = sq * sq % P
        b = cst
        for _ in range(m - i - 1):
            b = b * b % P
        m = i
        cst = b * b % P
        t = t * cst % P
        r = r * b % P
    if is_neg(r):
        r = P - r
    return r


def Bits2Point_Strict(c, bits):
    ins = c.inps("in", bits)
    out = c.sigs("out", 2)
    AliasCheck(c.sub("aliasCheckY"), ins[:254])
    c.check(ins[254].v, 0, "Bits2Point in[254]")
    y = Bits2Num(c.sub("b2nY"), ins[:254])
    c.set(out[1], y)
    y2 = y.v * y.v % P
    x = fr_sqrt(div(1 - y2, BJ_A - BJ_D * y2))
    if ins[255].v == 1:
        x = (-x) % P
    c.set(out[0], x)
    BabyCheck(c.sub("babyCheck"), out[0], out[1])
    xbits = Num2Bits(c.sub("n2bX"), 254, out[0])
    AliasCheck(c.sub("aliasCheckX"), xbits)
    sc = CompConstant(c.sub("signCalc"), HALF, xbits)
    c.check(sc.v, ins[255].v, "Bits2Point sign")
    return out[0], out[1]


# ---------------------------------------------------------------------------
# sha256/*.circom   (src/hash-inputs.circom:113, src/withdraw.circom:135)
# ---------------------------------------------------------------------------
SHA_H = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
         0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
SHA_K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]


def _const_bits(c, value):
    """H(x) / K(x): out[i] <== (const >> i) & 1."""
    return [c.inp("out[%d]" % i, (value >> i) & 1) for i in range(32)]


def nbits(a):
    n, r = 1, 0
    while n - 1 < a:
        r += 1
        n *= 2
    return r


def BinSum(c, ops):
    n = len(ops[0])
    nout = nbits(((1 << n) - 1) * len(ops))
    ins = [[c.inp("in[%d][%d]" % (j, k), ops[j][k]) for k in range(n)] for j in range(len(ops))]
    lin = 0
    for k in range(n):
        for j in range(len(ops)):
            lin += ins[j][k].v << k
    out = []
    lout = 0
    for k in range(nout):
        b = (lin >> k) & 1
        out.append(c.inp("out[%d]" % k, b))
        lout += b << k
    c.check(lin, lout, "BinSum")
    return out


def Xor3(c, a, b, cc):
    n = len(a)
    a_ = c.inps("a", a)
    b_ = c.inps("b", b)
    c_ = c.inps("c", cc)
    out = c.sigs("out", n)
    mid = c.sigs("mid", n)
    for k in range(n):
        c.set(mid[k], b_[k].v * c_[k].v)
        c.set(out[k], a_[k].v * (1 - 2 * b_[k].v - 2 * c_[k].v + 4 * mid[k].v)
              + b_[k].v + c_[k].v - 2 * mid[k].v)
    return out


def RotR(c, x, r):
    n = len(x)
    ins = c.inps("in", x)
    return [c.inp("out[%d]" % i, ins[(i + r) % n]) for i in range(n)]


def ShR(c, x, r):
    n = len(x)
    ins = c.inps("in", x)
    return [c.inp("out[%d]" % i, 0 if i + r >= n else ins[i + r]) for i in range(n)]


def SmallSigma(c, x, ra, rb, rc):
    ins = c.inps("in", x)
    out = c.sigs("out", 32)
    rota = RotR(c.sub("rota"), ins, ra)
    rotb = RotR(c.sub("rotb"), ins, rb)
    shrc = ShR(c.sub("shrc"), ins, rc)
    x3 = Xor3(c.sub("xor3"), rota, rotb, shrc)
    for k in range(32):
        c.set(out[k], x3[k])
    return out


def BigSigma(c, x, ra, rb, rc):
    ins = c.inps("in", x)
    out = c.sigs("out", 32)
    rota = RotR(c.sub("rota"), ins, ra)
    rotb = RotR(c.sub("rotb"), ins, rb)
    rotc = RotR(c.sub("rotc"), ins, rc)
    x3 = Xor3(c.sub("xor3"), rota, rotb, rotc)
    for k in range(32):
        c.set(out[k], x3[k])
    return out


def Ch_t(c, a, b, cc):
    n = len(a)
    a_ = c.inps("a", a)
    b_ = c.inps("b", b)
    c_ = c.inps("c", cc)
    return [c.inp("out[%d]" % k, a_[k].v * (b_[k].v - c_[k].v) + c_[k].v) for k in range(n)]


def Maj_t(c, a, b, cc):
    n = len(a)
    a_ = c.inps("a", a)
    b_ = c.inps("b", b)
    c_ = c.inps("c", cc)
    out = c.sigs("out", n)
    mid = c.sigs("mid", n)
    for k in range(n):
        c.set(mid[k], b_[k].v * c_[k].v)
        c.set(out[k], a_[k].v * (b_[k].v + c_[k].v - 2 * mid[k].v) + mid[k].v)
    return out


def SigmaPlus(c, in2, in7, in15, in16):
    i2 = c.inps("in2", in2)
    i7 = c.inps("in7", in7)
    i15 = c.inps("in15", in15)
    i16 = c.inps("in16", in16)
    out = c.sigs("out", 32)
    s1 = SmallSigma(c.sub("sigma1"), i2, 17, 19, 10)
    s0 = SmallSigma(c.sub("sigma0"), i15, 7, 18, 3)
    sm = BinSum(c.sub("sum"), [s1, i7, s0, i16])
    for k in range(32):
        c.set(out[k], sm[k])
    return out


def T1(c, h, e, f, g, k, w):
    h_ = c.inps("h", h)
    e_ = c.inps("e", e)
    f_ = c.inps("f", f)
    g_ = c.inps("g", g)
    k_ = c.inps("k", k)
    w_ = c.inps("w", w)
    out = c.sigs("out", 32)
    ch_c = c.sub("ch")
    bs_c = c.sub("bigsigma1")
    bs = BigSigma(bs_c, e_, 6, 11, 25)
    ch = Ch_t(ch_c, e_, f_, g_)
    sm = BinSum(c.sub("sum"), [h_, bs, ch, k_, w_])
    for i in range(32):
        c.set(out[i], sm[i])
    return out


def T2(c, a, b, cc):
    a_ = c.inps("a", a)
    b_ = c.inps("b", b)
    c_ = c.inps("c", cc)
    out = c.sigs("out", 32)
    bs = BigSigma(c.sub("bigsigma0"), a_, 2, 13, 22)
    mj = Maj_t(c.sub("maj"), a_, b_, c_)
    sm = BinSum(c.sub("sum"), [bs, mj])
    for i in range(32):
        c.set(out[i], sm[i])
    return out


def _sha256_compress_words(hwords, block_words):
    """Plain SHA-256 compression on 32-bit words (the `sha256compression()`
    helper function of sha256compression_function.circom)."""
    def rotr(x, r):
        return ((x >> r) | (x << (32 - r))) & 0xFFFFFFFF
    w = list(block_words)
    for t in range(16, 64):
        s0 = rotr(w[t - 15], 7) ^ rotr(w[t - 15], 18) ^ (w[t - 15] >> 3)
        s1 = rotr(w[t - 2], 17) ^ rotr(w[t - 2], 19) ^ (w[t - 2] >> 10)
        w.append((w[t - 16] + s0 + w[t - 7] + s1) & 0xFFFFFFFF)
    a, b, c_, d, e, f, g, h = hwords
    for t in range(64):
        S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)
        ch = (e & f) ^ ((~e) & g)
        t1 = (h + S1 + ch + SHA_K[t] + w[t]) & 0xFFFFFFFF
        S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)
        mj = (a & b) ^ (a & c_) ^ (b & c_)
        t2 = (S0 + mj) & 0xFFFFFFFF
        h, g, f, e, d, c_, b, a = g, f, e, (d + t1) & 0xFFFFFFFF, c_, b, a, (t1 + t2) & 0xFFFFFFFF
    return [(x + y) & 0xFFFFFFFF for x, y in zip(hwords, [a, b, c_, d, e, f, g, h])]


def Sha256compression(c, hin, inp):
    """hin: 8 words x 32 bits, LSB-first inside each word (hin[32*j+k] = bit k
    of word j); inp: 512 message bits MSB-first; out: 256 bits MSB-first."""
    hin_ = c.inps("hin", hin)
    inp_ = c.inps("inp", inp)
    out = c.sigs("out", 256)
    hwords = [sum(hin_[32 * j + k].v << k for k in range(32)) for j in range(8)]
    mwords = [sum(inp_[32 * t + k].v << (31 - k) for k in range(32)) for t in range(16)]
    res = _sha256_compress_words(hwords, mwords)
    for j in range(8):
        for k in range(32):
            c.set(out[32 * j + k], (res[j] >> (31 - k)) & 1)
    sigmaPlus_c = [c.sub("sigmaPlus[%d]" % i) for i in range(48)]
    ct_k = [_const_bits(c.sub("ct_k[%d]" % i), SHA_K[i]) for i in range(64)]
    t1_c = [c.sub("t1[%d]" % i) for i in range(64)]
    t2_c = [c.sub("t2[%d]" % i) for i in range(64)]
    suma_c = [c.sub("suma[%d]" % i) for i in range(64)]
    sume_c = [c.sub("sume[%d]" % i) for i in range(64)]
    fsum_c = [c.sub("fsum[%d]" % i) for i in range(8)]
    w = []
    for t in range(64):
        if t < 16:
            w.append([inp_[t * 32 + 31 - k] for k in range(32)])
        else:
            w.append(SigmaPlus(sigmaPlus_c[t - 16], w[t - 2], w[t - 7], w[t - 15], w[t - 16]))
    a, b, cc, d, e, f, g, h = [[hin_[32 * j + k] for k in range(32)] for j in range(8)]
    for t in range(64):
        t1 = T1(t1_c[t], h, e, f, g, ct_k[t], w[t])
        t2 = T2(t2_c[t], a, b, cc)
        se = BinSum(sume_c[t], [d, t1])
        sa = BinSum(suma_c[t], [t1, t2])
        h, g, f, e, d, cc, b, a = g, f, e, se[:32], cc, b, a, sa[:32]
    fin = [a, b, cc, d, e, f, g, h]
    for j in range(8):
        fs = BinSum(fsum_c[j], [[hin_[32 * j + k] for k in range(32)], fin[j]])
        for k in range(32):
            c.check(out[32 * j + 31 - k].v, fs[k].v, "Sha256compression out")
    return out


def Sha256(c, bits):
    nBits = len(bits)
    ins = c.inps("in", bits)
    out = c.sigs("out", 256)
    nBlocks = (nBits + 64) // 512 + 1
    padded = []
    for k in range(nBits):
        padded.append(c.inp("paddedIn[%d]" % k, ins[k]))
    padded.append(c.inp("paddedIn[%d]" % nBits, 1))
    for k in range(nBits + 1, nBlocks * 512 - 64):
        padded.append(c.inp("paddedIn[%d]" % k, 0))
    for k in range(nBlocks * 512 - 64, nBlocks * 512):
        kk = nBlocks * 512 - k - 1
        padded.append(c.inp("paddedIn[%d]" % k, (nBits >> kk) & 1))
    ha = [_const_bits(c.sub("ha%d" % j), SHA_H[j]) for j in range(8)]
    prev = None
    for i in range(nBlocks):
        cmp_c = c.sub("sha256compression[%d]" % i)
        if i == 0:
            hin = [ha[j][k] for j in range(8) for k in range(32)]
        else:
            hin = [prev[32 * j + 31 - k] for j in range(8) for k in range(32)]
        prev = Sha256compression(cmp_c, hin, padded[i * 512:(i + 1) * 512])
    for k in range(256):
        c.set(out[k], prev[k])
    return out
