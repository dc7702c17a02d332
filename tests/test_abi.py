"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/lbm_b200.h declares, carries the same constant tables as the oracle, and the product
path has no CPU fallback."""
import os
import re

import numpy as np
import pytest

import lbm
from lbm import _abi
import oracle.lbm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lbm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lbm_b200.h but not exported"
    assert sorted(_abi.EXPORTS) == names, "lbm/_abi.py EXPORTS out of sync with the header"
    assert lib.lbm_abi_version() == _abi.LBM_ABI_VERSION


def test_desc_struct_layout_matches_header():
    # sizes implied by the C declarations (natural alignment)
    assert _abi.C.sizeof(_abi.lbm_bc) == 6 * 4 + 4 * 8
    assert _abi.C.sizeof(_abi.lbm_desc) == 8 * 4 + 16 * 8 + 8 + 8 * _abi.C.sizeof(_abi.lbm_bc) + 3 * 4 + 128 + 4


@pytest.mark.parametrize("name", list(_abi.LATTICE_IDS))
def test_builtin_tables_equal_oracle_and_host(name):
    info = _abi.lattice_info(_abi.LATTICE_IDS[name])
    qo = O.L.BY_NAME[name]()
    assert info["Q"] == qo.Q
    assert np.array_equal(info["cx"], qo.cx) and np.array_equal(info["cy"], qo.cy)
    assert np.array_equal(info["w"], qo.w), "weights must be bit-identical"
    assert info["css"] == qo.css
    assert np.array_equal(info["opposite"], qo.opp)
    assert info["eq_order"] == qo.eq_order and info["hermite_order"] == qo.N and info["halo"] == qo.h
    q = getattr(lbm.Quadratures, name)
    q.check_against_library()
    assert lbm.order(q) == qo.order
    for i in range(q.Q):
        assert lbm.opposite(q, i + 1) == qo.opp[i] + 1


def test_bad_arguments_are_reported_not_crashed():
    with pytest.raises(lbm.LbmError):
        _abi.lattice_info(99)
    d = _abi.lbm_desc()
    h = _abi.C.c_void_p()
    assert _abi.lib().lbm_create(_abi.C.byref(d), _abi.C.byref(h)) == -1  # abi_version 0
    assert b"abi_version" in _abi.lib().lbm_last_error()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(lbm.LbmError):
        _abi.Context(8, 8, "D2Q9", _abi.SRT, [1.0])
    with pytest.raises(lbm.LbmError):
        lbm.collide_(lbm.SRT(1.0), lbm.D2Q9(), np.asfortranarray(np.ones((2, 2, 9))))
    with pytest.raises(lbm.LbmError):
        lbm.simulate(lbm.TGV(lbm.D2Q9(), 0.8, 1), lbm.D2Q9(), t_end=2.0)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "latticeboltzmann.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cuh", ".jl")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f"{fn} imports oracle/"
                assert "lbm_oracle" not in text, f"{fn} references the oracle"


def test_julia_shim_matches_the_c_abi():
    """The `ccall` shim (unexecuted here: no Julia in the image) is kept honest statically: its structs mirror lbm_bc / lbm_desc
    field for field, and every symbol it calls is declared in include/lbm_b200.h with the same number of arguments."""
    jl = open(os.path.join(ROOT, "latticeboltzmann.jl_b200", "julia", "LatticeBoltzmannB200.jl")).read()
    jtypes = {"Int32": _abi.C.c_int32, "Float64": _abi.C.c_double, "NTuple{2, Float64}": _abi.C.c_double * 2,
              "NTuple{LBM_MAX_TAU, Float64}": _abi.C.c_double * _abi.LBM_MAX_TAU, "NTuple{LBM_MAX_BCS, LbmBc}": _abi.lbm_bc * _abi.LBM_MAX_BCS,
              "NTuple{128, UInt8}": _abi.C.c_uint8 * 128}
    for jname, ctype in (("LbmBc", _abi.lbm_bc), ("LbmDesc", _abi.lbm_desc)):
        body = re.search(rf"struct {jname}\b[^\n]*\n(.*?)\nend", jl, re.S).group(1)
        fields = re.findall(r"(\w+)::((?:NTuple\{[^}]*\})|\w+)", body)
        assert [f for f, _ in fields] == [f for f, _ in ctype._fields_], jname
        for (fname, jt), (_, ct) in zip(fields, ctype._fields_):
            assert _abi.C.sizeof(jtypes[jt]) == _abi.C.sizeof(ct), (jname, fname, jt)
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lbm_b200.h")).read(), flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(lbm_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, re.S)}
    calls = re.findall(r"ccall\(\(:(lbm_\w+), LIB\),\s*\w+,\s*\(([^)]*)\)", jl)
    assert len(calls) >= 15
    for name, argtypes in calls:
        assert name in decl, f"{name} is not declared in the header"
        n_julia = len([a for a in argtypes.split(",") if a.strip()])
        params = decl[name].strip()
        n_c = 0 if params in ("", "void") else len(params.split(","))
        assert n_julia == n_c, (name, argtypes, params)
    assert re.search(r"LBM_MAX_TAU, LBM_MAX_BCS = (\d+), (\d+)", jl).groups() == (str(_abi.LBM_MAX_TAU), str(_abi.LBM_MAX_BCS))
