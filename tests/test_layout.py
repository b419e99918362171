"""Repository contract checks that need no GPU: the product never touches the oracle, the C ABI library
loads and exports every symbol the header declares, and creating a context without a device fails loudly."""
import ctypes
import os
import re

import pytest

from smcpp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "smcpp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "oracle/" in txt or "libsmcb_oracle" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference the oracle: {bad}"


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "smcpp_b200.h")).read()
    declared = set(re.findall(r"\b(smcpp_b200_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"smcpp_b200_ctx", "smcpp_b200_stats_t"}
    assert declared, "no declarations parsed"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/smcpp_b200.h but not exported"
    assert lib.smcpp_b200_abi_version() == 2


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.Context(0)


def test_built_for_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


_SASS_DUMP = []


def _sass(function_substring):
    import subprocess
    if not _SASS_DUMP:
        _SASS_DUMP.append(subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout)
    body, keep = [], False
    for line in _SASS_DUMP[0].splitlines():
        if "Function :" in line:
            keep = function_substring in line
        elif keep:
            body.append(line)
    return "\n".join(body)


@pytest.mark.parametrize("kernel,mnemonics", [
    ("k_stats32ENS", ["DMMA.8x8x4", "LDGSTS", "MUFU.RCP64H"]),      # FP64 tensor path, cp.async operand ring, branch-free reciprocal
    ("k_stats32eENS", ["DMMA.8x8x4", "LDGSTS"]),
    ("k_forward_mmaILi1ELi1E", ["DMMA.8x8x4", "FFMA2"]),           # tensor-path recursion with the packed float step
    ("k_backward_mmaILi1ELi1E", ["DMMA.8x8x4"]),
])
def test_hot_kernels_contain_the_instructions_the_design_names(kernel, mnemonics):
    """The shipped cubin is what DESIGN.md describes (guards against a build that silently lost a code path)."""
    sass = _sass(kernel)
    assert sass, f"{kernel} not found in {capi.LIB_PATH}"
    for mn in mnemonics:
        assert mn in sass, f"{mn} missing from {kernel}"


def test_header_is_valid_c99_and_a_plain_c_consumer_links(tmp_path):
    """examples/cabi_estep.c: strict C99 against include/smcpp_b200.h, linked to the in-tree library."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "cabi_estep"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "examples", "cabi_estep.c"), "-L" + os.path.join(root, "smcpp_b200"), "-lsmcpp_b200",
                           "-Wl,-rpath," + os.path.join(root, "smcpp_b200"), "-lm", "-o", str(exe)])
    assert exe.exists()
