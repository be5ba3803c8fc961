"""Hosts on the C ABI other than the Python binding: a plain C program (bindings/c/hzw_cli.c, the shape of the
reference's `./circuit input.json witness.wtns`, /root/reference/tools/helpers/actions.js:139-143) compiled with gcc
against the shared library and run as a process; and the N-API addon (bindings/node), type-checked against a
declaration stub of the N-API subset it uses (there is no Node in this image)."""
import os
import shutil
import subprocess

import pytest

from conftest import CPU_PORT, CUDA_LIB, ROOT, cpu_port_calculator

PRM = (3, 16, 2, 2)


def _build_cli(tmp_path, lib):
    exe = str(tmp_path / "hzw_cli")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "bindings", "c", "hzw_cli.c"), lib, "-Wl,-rpath," + os.path.dirname(lib), "-o", exe], check=True)
    return exe


def _drive(tmp_path, lib, small_bb, small_oracle):
    from oracle.wtns import parse_wtns
    exe = _build_cli(tmp_path, lib)
    wc = cpu_port_calculator(*PRM)
    inp = small_bb.getInput()
    (tmp_path / "in.bin").write_bytes(bytes(wc.packInput(inp)))
    args = [exe, "rollup-main"] + [str(x) for x in PRM] + [str(tmp_path / "in.bin"), str(tmp_path / "w.wtns"), "1"]
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "w1=%064x" % small_bb.getHashInputs() in r.stdout
    assert parse_wtns((tmp_path / "w.wtns").read_bytes()) == small_oracle[0]
    # a violated constraint: the runtime's message on stderr, exit status 1
    bad = dict(inp)
    bad["oldStateRoot"] = int(inp["oldStateRoot"]) + 1
    (tmp_path / "bad.bin").write_bytes(bytes(wc.packInput(bad)))
    args[6] = str(tmp_path / "bad.bin")
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Constraint doesn't match ")


def test_c_host_on_the_c_abi(built, small_bb, small_oracle, tmp_path):
    _drive(tmp_path, CPU_PORT, small_bb, small_oracle)          # same C ABI, the checker's library (no GPU here)


@pytest.mark.gpu
def test_c_host_on_the_cuda_library(built, small_bb, small_oracle, tmp_path):
    _drive(tmp_path, CUDA_LIB, small_bb, small_oracle)


def test_napi_addon_type_checks(tmp_path):
    node = os.path.join(ROOT, "bindings", "node")
    inc = tmp_path / "inc"
    inc.mkdir()
    shutil.copy(os.path.join(node, "test", "node_api_stub.h"), inc / "node_api.h")
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", str(inc), os.path.join(node, "hzwitness_napi.cc")], check=True)
    for f in ("witness_calculator.js", "binding.gyp", "package.json"):
        assert os.path.getsize(os.path.join(node, f)) > 100
    # every C ABI symbol the addon calls is declared in the header and exported by the library
    import re
    src = open(os.path.join(node, "hzwitness_napi.cc")).read()
    hdr = open(os.path.join(ROOT, "include", "hzwitness.h")).read()
    for sym in sorted(set(re.findall(r"\b(hzw_[a-z0-9_]+)\(", src))):
        assert re.search(r"\b%s\(" % sym, hdr), sym
