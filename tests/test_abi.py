"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/lbm_b200.h declares, carries the same constant tables as the oracle, and the product
path has no CPU fallback."""
import os
import re
import subprocess

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


def _build_c_consumer(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-ldl"])
    return exe


def _c_layout(tmp_path):
    out = subprocess.check_output([_build_c_consumer(tmp_path), "layout"], text=True)
    fields, sizes, consts = {}, {}, {}
    for line in out.splitlines():
        w = line.split()
        if w[0] == "sizeof":
            sizes[w[1]] = int(w[2])
        elif w[0] in ("const", "enum"):
            consts[w[1]] = int(w[2])
        else:
            st, f = w[0].split(".")
            fields.setdefault(st, []).append((f, int(w[1]), int(w[2])))
    return fields, sizes, consts


def test_c_consumer_layout_equals_ctypes_mirror(tmp_path):
    """include/lbm_b200.h compiled by gcc (tests/c_abi_smoke.c) against the hand-written ctypes structs of lbm/_abi.py:
    every field's name, offset and size, the struct sizes, the constants and the enum values the binding hard-codes."""
    fields, sizes, consts = _c_layout(tmp_path)
    mirror = {"lbm_bc": _abi.lbm_bc, "lbm_desc": _abi.lbm_desc, "lbm_sep_field": _abi.lbm_sep_field,
              "lbm_batch_stop": _abi.lbm_batch_stop}
    assert set(fields) == set(mirror)
    for name, ct in mirror.items():
        assert sizes[name] == _abi.C.sizeof(ct), name
        assert [(f, getattr(ct, f).offset, getattr(ct, f).size) for f, _ in ct._fields_] == fields[name], name
    assert consts["LBM_ABI_VERSION"] == _abi.LBM_ABI_VERSION and consts["LBM_MAX_Q"] == _abi.LBM_MAX_Q
    assert consts["LBM_MAX_TAU"] == _abi.LBM_MAX_TAU and consts["LBM_MAX_BCS"] == _abi.LBM_MAX_BCS
    assert consts["LBM_NCCL_ID_BYTES"] == _abi.LBM_NCCL_ID_BYTES
    assert consts["LBM_D2Q37"] == _abi.LATTICE_IDS["D2Q37"] and consts["LBM_F32"] == _abi.F32
    assert consts["LBM_ITERATIVE_INIT"] == _abi.ITERATIVE_INIT and consts["LBM_ARITH_FAST"] == _abi.ARITH_FAST
    assert consts["LBM_BC_MOVING_WALL"] == _abi.BC_MOVING_WALL and consts["LBM_WEST"] == _abi.WEST
    assert consts["LBM_REDUCE_DENSITY_CHANGE"] == _abi.REDUCE_DENSITY_CHANGE
    assert consts["LBM_BATCH_STOP_VELOCITY_CONVERGENCE"] == _abi.BATCH_STOP_VELOCITY_CONVERGENCE


@pytest.mark.gpu
def test_c_consumer_steps_config_c1_and_a_walled_case(tmp_path):
    """A C host (no Python, no ctypes between it and the library) fills lbm_desc, steps BASELINE config 1 (D2Q9 SRT shear
    wave, 64 x 64 periodic, tau = 1, 1000 steps) and a D2Q9 TRT + force Poiseuille channel, and compares bit for bit with
    dumps written by the oracle -- the single-context path and a 3-problem batch."""
    import oracle.lbm_oracle as O
    from oracle.c_oracle import COracle
    exe = _build_c_consumer(tmp_path)
    qo = O.L.D2Q9()
    pr = O.DecayingShearFlow(1 / 6, u_max=0.02 / 8, NX=64, NY=64, static=False, convenience=False)
    f0 = O.initialize("AnalyticalEquilibrium", qo, pr)
    cases = [("c1", f0, O.collision_model("SRT", qo, pr), [], 64, 64, 1000, _abi.SRT, (1.0, 0.0), (0.0, 0.0), 0)]
    nx, ny = 24, 13
    rng = np.random.default_rng(3)
    f1 = np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)])
    walls = [O.BounceBack("N", (1, nx), (1, ny)), O.BounceBack("S", (1, nx), (1, ny))]
    cases.append(("poiseuille", f1, O.TRT(0.8, 1.1, (1e-6, 0.0)), walls, nx, ny, 57, _abi.TRT, (0.8, 1.1), (1e-6, 0.0), 1))
    for name, f, cm, bcs, nx, ny, nsteps, code, taus, force, has_walls in cases:
        want, _ = COracle(qo, cm, bcs).steps(f, nsteps)
        f.astype(np.float64).tofile(tmp_path / f"{name}_f0.bin")
        want.tofile(tmp_path / f"{name}_want.bin")
        r = subprocess.run([exe, "run", _abi.LIB_PATH, str(tmp_path / f"{name}_f0.bin"), str(tmp_path / f"{name}_want.bin"),
                            str(nx), str(ny), str(nsteps), str(_abi.LATTICE_IDS["D2Q9"]), str(code), repr(taus[0]),
                            repr(taus[1]), repr(force[0]), repr(force[1]), str(has_walls)], capture_output=True, text=True)
        assert r.returncode == 0, (name, r.returncode, r.stdout, r.stderr)
        assert "0 of" in r.stdout


def _c_kind(ctype):
    """coarse class of a C parameter type as the header spells it"""
    t = ctype.strip()
    if "*" in t or "[" in t:
        return "ptr"
    for k in ("int64_t", "int32_t", "double", "float", "size_t"):
        if k in t.split():
            return k
    if "int" in t.split():
        return "int32_t"
    raise AssertionError(f"unclassified C type {ctype!r}")


_JULIA_KIND = {"Int64": "int64_t", "Int32": "int32_t", "Cint": "int32_t", "Cdouble": "double", "Float64": "double",
               "Cfloat": "float", "Cstring": "ptr", "Csize_t": "size_t"}


def _julia_kind(jt):
    jt = jt.strip()
    if jt.startswith(("Ptr{", "Ref{")):
        return "ptr"
    return _JULIA_KIND[jt]


def test_julia_shim_matches_the_c_abi(tmp_path):
    """The `ccall` shim (unexecuted here: no Julia in the image) is kept honest statically: its structs mirror the C
    structs field for field (names, sizes, and -- through the compiled header -- offsets), and every ccall names a
    symbol the header declares with the same number AND kinds of arguments (pointer / int32 / int64 / double) and a
    matching return type."""
    jl = open(os.path.join(ROOT, "latticeboltzmann.jl_b200", "julia", "LatticeBoltzmannB200.jl")).read()
    jsize = {"Int32": 4, "Float64": 8, "NTuple{2, Float64}": 16, "NTuple{2, Ptr{Float64}}": 16,
             "NTuple{LBM_MAX_TAU, Float64}": 8 * _abi.LBM_MAX_TAU, "NTuple{LBM_MAX_BCS, LbmBc}": 56 * _abi.LBM_MAX_BCS,
             "NTuple{128, UInt8}": 128}
    jalign = {"Int32": 4, "Float64": 8, "NTuple{2, Float64}": 8, "NTuple{2, Ptr{Float64}}": 8, "NTuple{LBM_MAX_TAU, Float64}": 8,
              "NTuple{LBM_MAX_BCS, LbmBc}": 8, "NTuple{128, UInt8}": 1}
    fields, sizes, _ = _c_layout(tmp_path)
    for jname, cname in (("LbmBc", "lbm_bc"), ("LbmDesc", "lbm_desc"), ("LbmSepField", "lbm_sep_field"), ("LbmBatchStop", "lbm_batch_stop")):
        body = re.search(rf"struct {jname}\b[^\n]*\n(.*?)\nend", jl, re.S).group(1)
        jf = re.findall(r"(\w+)::(NTuple\{\w+, (?:\w+\{\w+\}|\w+)\}|\w+)", body)
        assert [f for f, _ in jf] == [f for f, _, _ in fields[cname]], jname
        off = 0
        for (fname, jt), (_, coff, csize) in zip(jf, fields[cname]):   # Julia lays isbits structs out like C
            off = (off + jalign[jt] - 1) // jalign[jt] * jalign[jt]
            assert (off, jsize[jt]) == (coff, csize), (jname, fname, jt)
            off += jsize[jt]
        assert (off + 7) // 8 * 8 == sizes[cname], jname
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lbm_b200.h")).read(), flags=re.S)
    decl = {m.group(2): (m.group(1), m.group(3)) for m in
            re.finditer(r"\b(int|void|int64_t|const char \*)\s*\b(lbm_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, re.S)}
    calls = re.findall(r"ccall\(\(:(lbm_\w+), LIB\),\s*(\w+),\s*\(((?:[^()]|\([^()]*\))*?)\)\s*(?:,|\))", jl)
    assert len(calls) >= 40 and len({c[0] for c in calls}) >= 25
    for name, ret, argtypes in calls:
        assert name in decl, f"{name} is not declared in the header"
        cret, params = decl[name]
        assert {"int": "Cint", "void": "Cvoid", "int64_t": "Int64", "const char *": "Cstring"}[cret] == ret, (name, ret, cret)
        jargs = [a for a in re.split(r",\s*(?![^{}]*\})", argtypes) if a.strip()]
        params = params.strip()
        cargs = [] if params in ("", "void") else [a for a in params.split(",")]
        assert len(jargs) == len(cargs), (name, argtypes, params)
        for ja, ca in zip(jargs, cargs):
            assert _julia_kind(ja) == _c_kind(ca), (name, ja, ca)
    assert re.search(r"LBM_MAX_TAU, LBM_MAX_BCS = (\d+), (\d+)", jl).groups() == (str(_abi.LBM_MAX_TAU), str(_abi.LBM_MAX_BCS))
    # behaviour the round-1 review found missing: time-dependent forces refreshed per batch, device-side norms in next!,
    # the package entry point and the array-level operators routed to the library
    for needle in ("lbm_set_force_separable", "lbm_reduce_errors", "lbm_reduce,", "@eval LatticeBoltzmann function simulate(problem::FluidFlowProblem",
                   "b200_collide!", "b200_stream!", "b200_apply!", "lbm_batch_run"):
        assert needle in jl, needle
