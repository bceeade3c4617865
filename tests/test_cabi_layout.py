"""ABI drift guard: a C program that includes include/fourierflows_b200.h prints every enum constant and the size / field
offsets of every struct that crosses the boundary; the ctypes mirror (`_lib.py`) must agree byte for byte."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fourierflows_b200.h")
STRUCTS = ["ffb_coef", "ffb_desc", "ffb_fuse", "ffb_problem_config"]


@pytest.fixture(scope="module")
def L():
    import fourierflows_jl_b200 as ff
    return ff._lib


def _header_constants():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(FFB_[A-Z0-9_]+)\s*=\s*-?\d+", txt)))


def _c_report(tmp_path, L):
    names = _header_constants()
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    lines += [f'  printf("const {n} %d\\n", (int){n});' for n in names]
    for s in STRUCTS:
        lines.append(f'  printf("sizeof {s} %zu\\n", sizeof({s}));')
        for fld, _ in getattr(L, s)._fields_:
            lines.append(f'  printf("offset {s}.{fld} %zu\\n", offsetof({s}, {fld}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "abi.c", tmp_path / "abi"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return [l.split() for l in out.strip().splitlines()]


def test_enum_constants_match_ctypes_mirror(tmp_path, L):
    rep = _c_report(tmp_path, L)
    consts = {r[1]: int(r[2]) for r in rep if r[0] == "const"}
    assert len(consts) >= 25
    from fourierflows_jl_b200 import dist as D
    mirror = dict(vars(L))
    mirror.update({"FFB_EXCHANGE_NCCL": D.EXCHANGE["nccl"], "FFB_EXCHANGE_PEER_STORE": D.EXCHANGE["peer-store"],
                   "FFB_EXCHANGE_COPY_ENGINE": D.EXCHANGE["copy-engine"]})
    for name, val in consts.items():
        assert name in mirror, f"{name} has no Python mirror"
        assert mirror[name] == val, name


def test_struct_layouts_match_ctypes_mirror(tmp_path, L):
    rep = _c_report(tmp_path, L)
    sizes = {r[1]: int(r[2]) for r in rep if r[0] == "sizeof"}
    offsets = {r[1]: int(r[2]) for r in rep if r[0] == "offset"}
    for s in STRUCTS:
        cls = getattr(L, s)
        assert C.sizeof(cls) == sizes[s], s
        for fld, _ in cls._fields_:
            assert getattr(cls, fld).offset == offsets[f"{s}.{fld}"], f"{s}.{fld}"
    # every field of the C structs is mirrored (field count from the header text)
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for s in STRUCTS:
        end = re.search(r"\}\s*" + s + r"\s*;", txt).start()
        body = txt[txt.rfind("typedef struct {", 0, end) + len("typedef struct {"):end]
        nfields = sum(len(decl.split(",")) for decl in body.split(";") if decl.strip())
        assert nfields == len(getattr(L, s)._fields_), s
