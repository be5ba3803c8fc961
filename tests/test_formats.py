"""The circom toolchain's files around the witness path: a `.sym` applied as a layout map
(`hzw_create(layout_map_path)`: witness delivered in the COMPILER's numbering, reduced layouts included) and an
`.r1cs` checked over the witness on the GPU (`hzw_r1cs_check`, the tester's `checkConstraints`).  No circom here, so
the files are written by tests/formats_util.py from this circuit's own symbol table: a seeded renumbering with 30 % of
the internal signals eliminated, and the constraints of every Num2Bits / IsZero found by name."""
import numpy as np
import pytest

from conftest import cpu_port_calculator
from formats_util import gadget_constraints, write_r1cs, write_sym

PRM = (3, 16, 2, 2)


def _sym_case(tmp_path, small_oracle, wc_plain):
    w_ref, names = small_oracle
    path = str(tmp_path / "circuit.sym")
    var_of_own, n_vars = write_sym(path, names, wc_plain.NInputs)
    return path, var_of_own, n_vars, w_ref, names


def _check_mapped(wc, inp, var_of_own, n_vars, w_ref, names, tmp_path):
    assert wc.NVars == n_vars
    w = wc.calculateWitness(inp, True)
    assert w[0] == 1 and len(w) == n_vars
    for own, var in var_of_own.items():
        assert w[var] == w_ref[own], (own, var)
    kept = [n for n, i in names.items() if i in var_of_own]
    for n in kept[:: max(len(kept) // 200, 1)]:
        assert wc.getSignal(w, n) == w_ref[names[n]]
        assert wc.signalIndex(n) == var_of_own[names[n]]
    gone = next(n for n, i in names.items() if i not in var_of_own)
    from circuits_b200.witness_calculator import WitnessError
    with pytest.raises(WitnessError):
        wc.signalIndex(gone)
    out = str(tmp_path / "w.wtns")
    wc.writeWtns(b"".join(x.to_bytes(32, "little") for x in w), out)
    from oracle.wtns import parse_wtns
    assert parse_wtns(open(out, "rb").read()) == w
    return w


def test_sym_layout_map_cpu_port(built, small_bb, small_oracle, tmp_path):
    plain = cpu_port_calculator(*PRM)
    path, var_of_own, n_vars, w_ref, names = _sym_case(tmp_path, small_oracle, plain)
    assert n_vars < plain.NVars                      # a reduced layout: fewer entries than this library evaluates
    wc = cpu_port_calculator(*PRM, sym=path)
    _check_mapped(wc, small_bb.getInput(), var_of_own, n_vars, w_ref, names, tmp_path)


def test_sym_with_a_foreign_name_or_a_hole_is_refused(built, small_oracle, tmp_path):
    from circuits_b200.witness_calculator import WitnessError
    plain = cpu_port_calculator(*PRM)
    path, *_ = _sym_case(tmp_path, small_oracle, plain)
    lines = open(path).read().splitlines()
    bad = str(tmp_path / "foreign.sym")
    open(bad, "w").write("\n".join(lines + ["999999,5,0,main.noSuchComponent.out"]) + "\n")
    with pytest.raises(WitnessError):
        cpu_port_calculator(*PRM, sym=bad)
    hole = str(tmp_path / "hole.sym")
    open(hole, "w").write("\n".join(ln for ln in lines if ln.split(",")[1] != "7") + "\n")
    with pytest.raises(WitnessError):
        cpu_port_calculator(*PRM, sym=hole)
    with pytest.raises(WitnessError):
        cpu_port_calculator(*PRM, sym=str(tmp_path / "missing.sym"))


def _r1cs_case(tmp_path, small_oracle, index_of=lambda i: i, n_wires=None, at_least=3000):
    w_ref, names = small_oracle
    cons = gadget_constraints(names, w_ref, index_of)
    assert len(cons) > at_least
    path = str(tmp_path / "circuit.r1cs")
    write_r1cs(path, cons, n_wires or len(w_ref))
    return path, cons


def test_r1cs_check_cpu_port(built, small_bb, small_oracle, tmp_path):
    wc = cpu_port_calculator(*PRM)
    path, cons = _r1cs_case(tmp_path, small_oracle)
    w = bytearray(wc.calculateBinWitness(small_bb.getInput(), True))
    n, bad, first, _ = wc.checkConstraints(path, w)
    assert (n, bad, first) == (len(cons), 0, None)
    # flip the bit of the 11th constraint's wire: that constraint (a booleanity check) must be the first to fail
    wire = next(iter(cons[10][0]))
    w[32 * wire] ^= 2
    n, bad, first, _ = wc.checkConstraints(path, w)
    assert bad >= 1 and first == 10


@pytest.mark.gpu
def test_sym_layout_map_and_r1cs_on_gpu(built, small_bb, small_oracle, tmp_path):
    from circuits_b200.witness_calculator import WitnessCalculator
    plain = WitnessCalculator(*PRM)
    path, var_of_own, n_vars, w_ref, names = _sym_case(tmp_path, small_oracle, plain)
    # own layout: constraints over the resident witness and over a host copy
    r1, cons = _r1cs_case(tmp_path, small_oracle)
    plain.calculateResident(plain.packInput(small_bb.getInput()), True)
    assert plain.checkConstraints(r1)[:3] == (len(cons), 0, None)
    w = bytearray(plain.readWitness())
    wire = next(iter(cons[10][0]))
    w[32 * wire] ^= 2
    n, bad, first, _ = plain.checkConstraints(r1, w)
    assert bad >= 1 and first == 10
    plain.close()
    # the compiler's layout: same witness values at the compiler's indexes, constraints numbered in it
    wc = WitnessCalculator(*PRM, sym=path)
    wm = _check_mapped(wc, small_bb.getInput(), var_of_own, n_vars, w_ref, names, tmp_path)
    (tmp_path / "mapped").mkdir()
    r2, cons2 = _r1cs_case(tmp_path / "mapped", small_oracle, index_of=lambda i: 0 if i == 0 else var_of_own.get(i), n_wires=n_vars,
                           at_least=100)        # a gadget with an eliminated signal has no constraint in the reduced system
    wc.calculateResident(wc.packInput(small_bb.getInput()), False)
    assert wc.checkConstraints(r2)[:3] == (len(cons2), 0, None)
    assert bytes(wc.readWitness()) == b"".join(x.to_bytes(32, "little") for x in wm)
    wc.close()
