"""The command line in the place of the reference's `./circuit input.json witness.wtns`
(/root/reference/tools/helpers/actions.js:139-143): input.json as tools/generate-input.js writes it (decimal
strings), `.wtns` in the snarkjs layout, witness.json as `snarkjs wtns export json` prints it."""
import json

import pytest

from conftest import CPU_PORT


def _stringify(v):
    return [_stringify(x) for x in v] if isinstance(v, (list, tuple)) else str(int(v))


def _run(tmp_path, small_bb, lib, sanity=True, mutate=None, out_name="witness.wtns"):
    from circuits_b200 import cli
    inp = {k: _stringify(v) for k, v in small_bb.getInput().items()}
    if mutate:
        mutate(inp)
    src, dst = tmp_path / "input.json", tmp_path / out_name
    src.write_text(json.dumps(inp))
    argv = ["--nTx", "3", "--nLevels", "16", "--maxL1Tx", "2", "--maxFeeTx", "2", str(src), str(dst)]
    if sanity:
        argv.insert(0, "--sanity")
    if lib:
        argv[0:0] = ["--lib", lib]
    return cli.main(argv), dst


def test_cli_writes_the_wtns_layout(built, small_bb, small_oracle, tmp_path):
    from oracle.wtns import parse_wtns
    w_ref, _ = small_oracle
    rc, dst = _run(tmp_path, small_bb, CPU_PORT)
    assert rc == 0
    assert parse_wtns(dst.read_bytes()) == list(w_ref)


def test_cli_witness_json(built, small_bb, small_oracle, tmp_path):
    w_ref, _ = small_oracle
    rc, dst = _run(tmp_path, small_bb, CPU_PORT, out_name="witness.json")
    assert rc == 0
    assert json.loads(dst.read_text()) == [str(int(x)) for x in w_ref]


def test_cli_reports_a_violated_constraint(built, small_bb, tmp_path, capsys):
    def bump(inp):
        inp["balance1"][0] = str(int(inp["balance1"][0]) + 1)
    rc, dst = _run(tmp_path, small_bb, CPU_PORT, mutate=bump)
    assert rc == 1 and not dst.exists()
    assert "Constraint doesn't match 1 != 0" in capsys.readouterr().err


@pytest.mark.gpu
def test_cli_on_gpu(built, small_bb, small_oracle, tmp_path):
    from oracle.wtns import parse_wtns
    w_ref, _ = small_oracle
    rc, dst = _run(tmp_path, small_bb, None)
    assert rc == 0
    assert parse_wtns(dst.read_bytes()) == list(w_ref)
