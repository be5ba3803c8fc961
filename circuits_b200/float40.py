"""float40 codec: 5-bit decimal exponent | 35-bit mantissa
(/root/reference/src/lib/decode-float.circom:4-10; host mirror of commonjs
`float40` [un-vendored] used at tools/helpers/gen-inputs-utils.js:16,58)."""


def float2fix(fl):
    m = fl & 0x7FFFFFFFF
    e = fl >> 35
    return m * 10 ** e


def fix2float(v):
    """Exact encoding; raises if `v` is not representable."""
    m, e = v, 0
    while m % 10 == 0 and m // 10 >= 0x800000000 and e < 31:
        m //= 10
        e += 1
    while m >= 0x800000000:
        if m % 10 or e >= 31:
            raise ValueError("not representable as float40: %d" % v)
        m //= 10
        e += 1
    return (e << 35) | m


def floor_fix2float(v):
    """Largest representable value <= v."""
    m, e = v, 0
    while m >= 0x800000000:
        m //= 10
        e += 1
    if e > 31:
        raise ValueError("float40 overflow")
    return (e << 35) | m


def round_fix(v):
    """commonjs float40.round: nearest representable value (ties down)."""
    lo = float2fix(floor_fix2float(v))
    m, e = v, 0
    while m >= 0x800000000:
        m //= 10
        e += 1
    hi = (m + 1) * 10 ** e
    if hi >> 0 and (m + 1) < 0x800000000 and (hi - v) < (v - lo):
        return hi
    return lo
