"""r-pkg/src/rshim.c (the R .Call shim) is compile-checked against a stub of R's C API and the real
include/cellula_b200.h: every C-ABI call it makes exists with that arity and those pointer types, every .Call
entry is registered with the right argument count, and every C-ABI symbol it uses is exported by the built
library.  (No R exists in the build image or on the GPU box; R CMD INSTALL does the real build on an R host.)"""
import ctypes
import os
import re
import subprocess

from cellula_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "r-pkg", "src", "rshim.c")


def test_shim_compiles_against_the_header_and_an_r_api_stub(tmp_path):
    obj = tmp_path / "rshim.o"
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-Wno-cast-function-type", "-c", SHIM, "-o", str(obj),
                        "-I", os.path.join(ROOT, "tests", "helpers", "r_stub"), "-I", os.path.join(ROOT, "include")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # undefined symbols of the object: the R API (stubbed) and the C ABI -- the latter must all be exported
    syms = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True).stdout.split()
    used = sorted(s for s in syms if s.startswith("cb200_"))
    assert len(used) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in used:
        assert hasattr(handle, s), f"rshim.c calls {s}, which libcellula_b200.so does not export"


def test_registered_call_entries_match_their_definitions():
    src = open(SHIM).read()
    defs = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
            for m in re.finditer(r"^SEXP (cb200r_\w+)\(([^)]*)\)", src, flags=re.M)}
    regs = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(cb200r_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', src)}
    assert defs == regs
    # and every .Call in the R wrappers names a registered entry with that many arguments
    seen = set()
    for fn in os.listdir(os.path.join(ROOT, "r-pkg", "R")):
        text = open(os.path.join(ROOT, "r-pkg", "R", fn)).read()
        for m in re.finditer(r'\.Call\("(cb200r_\w+)"((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)PACKAGE', text):
            name, args = m.group(1), m.group(2)
            assert name in regs, (fn, name)
            depth, n = 0, 0
            for ch in args:
                depth += ch in "([" 
                depth -= ch in ")]"
                n += ch == "," and depth == 0
            assert n - 1 == regs[name], (fn, name, n - 1, regs[name])
            seen.add(name)
    assert len(seen) >= 8, seen
