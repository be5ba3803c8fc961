"""Pin the oracle (oracle/) against upstream known-answer vectors and every
literal golden vector the reference's own tests hold."""
import hashlib

import pytest

from oracle import circomlib as cl
from oracle import hermez as hz
from oracle.circuit import ConstraintError
from oracle.field import P
from oracle.poseidon_constants import poseidon_hash


def I(d):
    return {k: ([int(x) for x in v] if isinstance(v, list) else int(v)) for k, v in d.items()}


# -- Poseidon: circomlib test/poseidonjs.js + go-iden3-crypto vectors ---------------
POSEIDON_KAT = [
    ([1, 2], 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a),
    ([1, 2, 3, 4], 0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465),
    ([1], 18586133768512220936620570745912940619677854269274689475585506675881198879027),
    ([1, 2, 0, 0, 0], 1018317224307729531995786483840663576608797660851238720571059489595066344487),
    ([1, 2, 0, 0, 0, 0], 15336558801450556532856248569924170992202208561737609669134139141992924267169),
    ([3, 4, 0, 0, 0], 5811595552068139067952687508729883632420015185677766880877743348592482390548),
    ([3, 4, 0, 0, 0, 0], 12263118664590987767234828103155242843640892839966517009184493198782366909018),
    ([1, 2, 3, 4, 5, 6], 20400040500897583745843009878988256314335038853985262692600694741116813247201),
]


@pytest.mark.parametrize("inputs,expected", POSEIDON_KAT)
def test_poseidon_kat(inputs, expected):
    assert poseidon_hash(inputs) == expected
    main, _, out = hz.run_template(cl.Poseidon, inputs)
    assert out.v == expected


def test_poseidon_signal_count():
    # unreduced signals per permutation (aliases collapsed)
    from oracle.circuit import Comp, Ctx, count
    for n, expect in [(2, 633), (3, 776), (4, 980), (5, 1140), (6, 1351)]:
        main = Comp(Ctx(), "main", is_main=True)
        ins = [main.inp("i%d" % k, k + 1) for k in range(n)]
        sub = main.sub("h")
        cl.Poseidon(sub, ins)
        assert count(sub) == expect + 1      # + the constant ark[0].in[0]


# -- float40: test/lib/decode-float.test.js:28-38 ----------------------------------------
def test_decode_float_reference_vectors(golden):
    assert len(golden["decode_float"]) == 9
    for fl, expected in golden["decode_float"]:
        _, _, out = hz.run_template(hz.DecodeFloat, int(fl))
        assert out.v == int(expected)


# -- RollupTxStates: test/rollup-tx-states.test.js:38-625 ---------------------------------
def test_rollup_tx_states_reference_tables(golden):
    cases = golden["rollup_tx_states"]
    assert len(cases) == 22
    order = ["fromIdx", "toIdx", "toEthAddr", "auxFromIdx", "auxToIdx", "amount", "newExit",
             "loadAmount", "newAccount", "onChain", "fromEthAddr", "ethAddr1", "tokenID", "tokenID1",
             "tokenID2"]
    for case in cases:
        inp = I(case["input"])
        _, _, o = hz.run_template(hz.RollupTxStates, *[inp[k] for k in order])
        for k, v in case["output"].items():
            assert o[k].v == int(v), (case["title"], k)


# -- RqTxVerifier: test/rq-tx-verifier.test.js:43-95 ---------------------------------------
def _rq(future, past, rq, off):
    return hz.run_template(hz.RqTxVerifier, future[0], past[0], future[1], past[1], future[2], past[2],
                           rq[0], rq[1], rq[2], off)


def test_rq_tx_verifier_reference_cases():
    z3, z4 = [0, 0, 0], [0, 0, 0, 0]
    _rq([z3, z3, z3], [z4, z4, z4], [0, 0, 0], 0)                      # :43-46 empty
    with pytest.raises(ConstraintError) as e:                          # :48-68 mismatch
        _rq([[1, 0, 0], z3, z3], [z4, z4, z4], [0, 0, 0], 1)
    assert "Constraint doesn't match 1 != 0" in str(e.value)
    fut = [[0] * 3 for _ in range(3)]
    past = [[0] * 4 for _ in range(3)]
    for i in range(1, 8):                                              # :70-94 all offsets
        for k in range(3):
            if i < 4:
                fut[k][i - 1] = i
            else:
                past[k][3 - (i - 4)] = i
        _rq(fut, past, [i, i, i], i)


# -- FeeAccumulator(16): test/fee-accumulator.test.js:28-113 (all 7 vectors) ------------------
def test_fee_accumulator_reference_vectors(golden):
    assert len(golden["fee_accumulator"]) == 7
    for v in golden["fee_accumulator"]:
        inp = I(v["input"])
        _, _, out = hz.run_template(hz.FeeAccumulator, inp["tokenID"], inp["fee2Charge"],
                                    inp["feePlanTokenID"], inp["accFeeIn"])
        assert [o.v for o in out] == [int(x) for x in v["accFeeOut"]]


# -- Mux256 identity: test/lib/mux256.test.js:28-57 --------------------------------------------
def test_mux256_identity():
    for sel in range(256):
        _, _, out = hz.run_template(hz.Mux256, [(sel >> i) & 1 for i in range(8)], list(range(256)))
        assert out.v == sel


# -- BalanceUpdater: test/balance-updater.test.js:31-190 ----------------------------------------
def test_balance_updater_reference_cases(golden):
    order = ["oldStBalanceSender", "oldStBalanceReceiver", "amount", "loadAmount", "feeSelector",
             "onChain", "nop", "nullifyLoadAmount", "nullifyAmount"]
    names = ["newStBalanceSender", "newStBalanceReceiver", "isP2Nop", "fee2Charge", "isAmountNullified"]
    assert len(golden["balance_updater"]) == 6
    for case in golden["balance_updater"]:
        inp = I(case["input"])
        if "error" in case:
            with pytest.raises(ConstraintError) as e:
                hz.run_template(hz.BalanceUpdater, *[inp[k] for k in order])
            assert "Constraint doesn't match 1 != 0" in str(e.value)
            continue
        _, _, out = hz.run_template(hz.BalanceUpdater, *[inp[k] for k in order])
        got = dict(zip(names, [o.v for o in out]))
        for k, v in case["output"].items():
            assert got[k] == int(v) % P, (case["title"], k)


# -- ComputeFee: literal balances of test/rollup-main.test.js:555,616 and the
#    128-bit overflow boundary of test/compute-fee.test.js:94-129 -------------------------------
def test_compute_fee_reference_literals():
    for amount, sel, fee in [(150, 126, 15), (100, 68, 1), (150, 184, 113), (500, 184, 378)]:
        _, _, out = hz.run_template(hz.ComputeFee, sel, amount, 1)
        assert out.v == fee
        _, _, out = hz.run_template(hz.ComputeFee, sel, amount, 0)
        assert out.v == 0
    amount = 0x2FF * 10 ** 31          # float2Fix(0xF8000002FF)
    hz.run_template(hz.ComputeFee, 207, amount, 1)
    with pytest.raises(ConstraintError) as e:
        hz.run_template(hz.ComputeFee, 208, amount, 1)
    assert "Constraint doesn't match" in str(e.value)


# -- BabyJubJub / pointbits --------------------------------------------------------------------
def test_babyjub_identities_and_sqrt():
    def add(p1, p2):
        _, _, o = hz.run_template(cl.BabyAdd, p1[0], p1[1], p2[0], p2[1])
        return (o[0].v, o[1].v)
    B8 = cl.BASE8
    assert (cl.BJ_A * B8[0] ** 2 + B8[1] ** 2 - 1 - cl.BJ_D * B8[0] ** 2 * B8[1] ** 2) % P == 0
    acc, pt, k = (0, 1), B8, cl.SUBORDER
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    assert acc == (0, 1)
    assert pow(5, 81540058820840996586704275553141814055101440848469862132140264610111, P) == \
        19103219067921713944291392827692070036145651957329286315305642004821462161904
    _, _, ax = hz.run_template(hz.AySign2Ax, 0, 0)     # dummy path, SURVEY.md A.4
    assert ax.v == 2957874849018779266517920829765869116077630550401372566248359756137677864698
    sign = 1 if B8[0] > (P - 1) // 2 else 0
    _, _, ax = hz.run_template(hz.AySign2Ax, B8[1], sign)
    assert ax.v == B8[0]


# -- SHA-256 bit circuit vs hashlib --------------------------------------------------------------
@pytest.mark.parametrize("nbits", [8, 440, 448, 688])
def test_sha256_bits_vs_hashlib(nbits):
    data = bytes((i * 37 + 11) & 0xFF for i in range(nbits // 8))
    bits = [(data[i // 8] >> (7 - i % 8)) & 1 for i in range(nbits)]
    _, _, out = hz.run_template(cl.Sha256, bits)
    got = bytes(sum(out[8 * i + k].v << (7 - k) for k in range(8)) for i in range(32))
    assert got == hashlib.sha256(data).digest()


# -- EdDSA verifier accepts honest signatures, rejects tampered ones ------------------------------
def test_eddsa_verifier_self_consistency():
    from circuits_b200 import babyjub
    kp = babyjub.KeyPair(0x1234567890ABCDEF)
    R8, S = kp.sign(987654321, 4242)
    hz.run_template(cl.EdDSAPoseidonVerifier, 1, kp.pub[0], kp.pub[1], S, R8[0], R8[1], 987654321)
    with pytest.raises(ConstraintError):
        hz.run_template(cl.EdDSAPoseidonVerifier, 1, kp.pub[0], kp.pub[1], S, R8[0], R8[1], 987654322)
    # disabled: no constraint fires even on a wrong message
    hz.run_template(cl.EdDSAPoseidonVerifier, 0, kp.pub[0], kp.pub[1], S, R8[0], R8[1], 987654322)
