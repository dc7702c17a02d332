"""GPU parity at PRODUCTION launch geometry against the C oracle (oracle/lbm_oracle.c, OpenMP).

The small-grid tests of test_gpu_parity.py launch 32x8 CTAs; the kernels that carry the bench numbers run 256x1 / 128x1
CTAs with per-<lattice, dtype> register budgets (`step_cfg`), the packed two-nodes-per-thread Float32 kernel
(`k_step_x2`) at real widths, and -- beyond 65 535 launched CTA rows -- a grid-stride loop.  Every one of those paths is
compared here with the oracle on the same inputs, through the C ABI:

  * every lattice x {SRT, TRT, MRT} x {Float64 exact, Float64 fast, Float32 fast} on 512 x 64 with a moving wall (North)
    and a bounce-back wall (South), uniform force, 10 steps;
  * BASELINE configs at their stated sizes: C2 (D2Q9 TRT TGV 4096^2, 20 steps, Float64 exact/fast + Float32),
    C3 (D2Q9 SRT + force Poiseuille 1024 x 8192, 10 steps), a C4 slab (D2Q37 TRT Couette 8192 x 96, 10 steps);
  * 256 x 70 000 rows (> 65 535 CTA rows): the grid-stride path.

Bar (BASELINE.json north_star): Float64 exact SRT/TRT bit-identical; Float64 MRT / fast 1e-12 relative (max norm);
Float32 1e-5 relative on populations AND moments (lattice/model cases) resp. on the velocities (BASELINE configs).
"""
import numpy as np
import pytest

import lbm
from lbm import _abi
from conftest import rel_max, to_host_layout, to_oracle_layout

pytestmark = pytest.mark.gpu

LATTICES = ["D2Q4", "D2Q5", "D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"]
TOL64, TOL32 = 1e-12, 1e-5
MODES = {"f64-exact": (_abi.F64, _abi.ARITH_EXACT), "f64-fast": (_abi.F64, _abi.ARITH_FAST),
         "f32-fast": (_abi.F32, _abi.ARITH_FAST)}


def _smooth_start(O, qo, nx, ny, u_max=2e-3):
    """Taylor-Green-like equilibrium populations, |u| ~ u_max (Float32 keeps ~7 digits of it in deviation storage)."""
    pr = O.TGV(qo, 0.8, 1, nx, ny, u_max=u_max)
    return O.initialize("AnalyticalEquilibrium", qo, pr)


def _model(O, qo, kind, force):
    taus = [0.8, 0.9, 1.1, 1.3][:max(qo.N, 2)]
    return {"SRT": (O.SRT(0.8, force), _abi.SRT, [0.8]), "TRT": (O.TRT(0.8, 1.1, force), _abi.TRT, [0.8, 1.1]),
            "MRT": (O.MRT(qo, taus, force), _abi.MRT, taus)}[kind]


def _velocities(O, qo, f):
    fl = [f[i] for i in range(qo.Q)]
    rho = O.density(qo, fl)
    ux, uy = O.velocity(qo, fl, rho)
    return rho, ux, uy


def _check(O, qo, got, want, mode, model, moments=None):
    dtype, arith = MODES[mode]
    if dtype == _abi.F64:
        if arith == _abi.ARITH_EXACT and model != "MRT":
            assert np.array_equal(got, want), f"not bit-identical: max abs diff {np.abs(got - want).max():.3e}"
        assert rel_max(got, want) < TOL64
    else:
        assert rel_max(got, want) < TOL32
    if moments is not None:
        rho, ux, uy = _velocities(O, qo, want)
        tol = TOL64 if dtype == _abi.F64 else TOL32
        scale = max(np.abs(ux).max(), np.abs(uy).max())
        assert rel_max(moments["rho"].T, rho) < tol
        assert np.abs(moments["ux"].T - ux).max() < tol * scale and np.abs(moments["uy"].T - uy).max() < tol * scale


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("model", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("mode", list(MODES))
def test_production_ctas_all_lattices_models_dtypes(oracle, name, model, mode):
    """512 x 64, walls: 256x1 / 128x1 CTAs, the per-<lattice, dtype> `step_cfg` register budgets, `k_step_x2` (Float32
    fast, Q <= 13, SRT/TRT) -- vs the C oracle, 10 steps + the resume path."""
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.BY_NAME[name]()
    nx, ny, nsteps = 512, 64, 10
    f0 = _smooth_start(O, qo, nx, ny)
    force = (1e-6, -1e-6)
    cm, code, taus = _model(O, qo, model, force)
    ob = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.001, 0.0002])]
    hb = [lbm.BounceBack(lbm.South(), (1, nx), (1, ny)).to_abi(),
          lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.001, 0.0002]).to_abi()]
    co = COracle(qo, cm, ob)
    want, want_c = co.steps(f0, nsteps)
    want2, _ = co.steps(want, 2)
    dtype, arith = MODES[mode]
    with _abi.Context(nx, ny, name, code, taus, hb, dtype=dtype, arith=arith) as c:
        c.set_force_uniform(*force)
        c.upload_f(to_host_layout(f0))
        c.step(0, nsteps)
        mo = c.moments(0.3, ("rho", "ux", "uy"))   # pulled from f_collision (diagnostics path)
        got = to_oracle_layout(c.download_f())
        got_c = to_oracle_layout(c.download_f_collision())
        c.step(nsteps, 2)                           # resume from the retained f_collision
        got2 = to_oracle_layout(c.download_f())
    _check(O, qo, got, want, mode, model, mo)
    _check(O, qo, got_c, want_c, mode, model)
    _check(O, qo, got2, want2, mode, model)


def _tgv_case(O, n):
    qo = O.L.D2Q9()
    pr = O.TGV(qo, 0.8, n // 16)
    f0 = O.initialize("AnalyticalEquilibrium", qo, pr)
    return qo, pr, f0, O.collision_model("TRT", qo, pr)


def test_config_c2_full_size_4096_20_steps(oracle):
    """BASELINE configs[1] at its real size: D2Q9 TRT Taylor-Green vortex, 4096 x 4096 periodic, 20 steps from the
    analytic initial condition -- the bench workload itself -- against the C oracle: Float64 exact bit-identical,
    Float64 fast 1e-12, Float32 fast 1e-5 on populations and velocities."""
    O = oracle
    from oracle.c_oracle import COracle
    n, nsteps = 4096, 20
    qo, pr, f0, cm = _tgv_case(O, n)
    want, _ = COracle(qo, cm).steps(f0, nsteps)
    rho, ux, uy = _velocities(O, qo, want)
    q = lbm.D2Q9()
    hp = lbm.TGV(q, 0.8, n // 16)
    host0 = to_host_layout(f0)
    for mode, (dtype, arith) in MODES.items():
        hcm = lbm.CollisionModel(lbm.TRT, q, hp)
        ctx = lbm.model.make_context(q, hcm, hp.boundary_conditions(), n, n, {_abi.F64: "f64", _abi.F32: "f32"}[dtype], arith)
        st = lbm.DeviceState(ctx, q, hcm)
        ctx.upload_f(host0)
        st.step(0, nsteps, hp.delta_t())
        mo = ctx.moments(1.0, ("ux", "uy"))
        got = to_oracle_layout(ctx.download_f())
        ctx.close()
        tol = TOL64 if dtype == _abi.F64 else TOL32
        if mode == "f64-exact":
            assert np.array_equal(got, want)
        assert rel_max(got, want) < tol, mode
        # velocities (|u| <= 4e-4 here): Float64 in lattice units against the unit lattice speed -- an error of 1e-12 in
        # u is what a relative 1e-12 on the populations (f ~ w ~ 0.1 .. 0.4) can leave; bit-identical in exact mode --
        # Float32 relative to max |u| (deviation storage keeps the small velocities resolved)
        eu = max(np.abs(mo["ux"].T - ux).max(), np.abs(mo["uy"].T - uy).max())
        umax = max(np.abs(ux).max(), np.abs(uy).max())
        if mode == "f64-exact":
            assert eu <= 1e-15 * umax, (mode, eu, umax)
        elif dtype == _abi.F64:
            assert eu < TOL64, (mode, eu, umax)
        else:
            assert eu < TOL32 * umax, (mode, eu, umax)
        del got, mo


def test_config_c3_full_size_poiseuille_1024x8192(oracle):
    """BASELINE configs[2]: D2Q9 SRT + uniform force, bounce-back North + South, 1024 x 8192, 10 steps (Float64 exact
    bit-identical, fast 1e-12, Float32 1e-5)."""
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.D2Q9()
    nx, ny, nsteps = 1024, 8192, 10
    f0 = _smooth_start(O, qo, nx, ny, u_max=1e-3)
    force = (1.3e-7, 0.0)
    cm = O.SRT(0.9, force)
    ob = [O.BounceBack("N", (1, nx), (1, ny)), O.BounceBack("S", (1, nx), (1, ny))]
    hb = [lbm.BounceBack(lbm.North(), (1, nx), (1, ny)).to_abi(), lbm.BounceBack(lbm.South(), (1, nx), (1, ny)).to_abi()]
    want, _ = COracle(qo, cm, ob).steps(f0, nsteps)
    for mode, (dtype, arith) in MODES.items():
        with _abi.Context(nx, ny, "D2Q9", _abi.SRT, [0.9], hb, dtype=dtype, arith=arith) as c:
            c.set_force_uniform(*force)
            c.upload_f(to_host_layout(f0))
            c.step(0, nsteps)
            got = to_oracle_layout(c.download_f())
        _check(O, qo, got, want, mode, "SRT")


def test_config_c4_slab_d2q37_couette_8192x96(oracle):
    """BASELINE configs[3] at its real width: D2Q37 TRT Couette (halo 3, moving wall North + bounce-back South), an
    8192 x 96 strip (both walls inside), 10 steps."""
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.D2Q37()
    nx, ny, nsteps = 8192, 96, 10
    f0 = _smooth_start(O, qo, nx, ny, u_max=1e-3)
    cm = O.TRT(0.8, 0.5 + 0.25 / 0.3, None)
    ob = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.004, 0.0])]
    hb = [lbm.BounceBack(lbm.South(), (1, nx), (1, ny)).to_abi(), lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.004, 0.0]).to_abi()]
    want, _ = COracle(qo, cm, ob).steps(f0, nsteps)
    for mode, (dtype, arith) in MODES.items():
        with _abi.Context(nx, ny, "D2Q37", _abi.TRT, [0.8, 0.5 + 0.25 / 0.3], hb, dtype=dtype, arith=arith) as c:
            c.set_force_none()
            c.upload_f(to_host_layout(f0))
            c.step(0, nsteps)
            got = to_oracle_layout(c.download_f())
        _check(O, qo, got, want, mode, "TRT")


@pytest.mark.parametrize("walls", [False, True])
def test_grid_stride_path_more_than_65535_rows(oracle, walls):
    """256 x 70 000: more launched rows than gridDim.y allows -> the kernels' grid-stride loop (D2Q9 TRT, Float64 exact
    bit-identical; Float32 packed kernel 1e-5)."""
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.D2Q9()
    nx, ny, nsteps = 256, 70000, 3
    rng = np.random.default_rng(5)
    # smooth in x, rough in y: every row differs, so a skipped or doubled row cannot cancel
    amp = 1e-3 * rng.uniform(-1, 1, (qo.Q, ny, 1)) * (1 + 0.5 * np.cos(2 * np.pi * np.arange(nx) / nx))[None, None, :]
    f0 = np.ascontiguousarray(qo.w[:, None, None] * (1 + amp))
    cm = O.TRT(0.8, 1.1, (1e-7, 0.0))
    ob, hb = [], []
    if walls:
        ob = [O.BounceBack("N", (1, nx), (1, ny)), O.BounceBack("S", (1, nx), (1, ny))]
        hb = [lbm.BounceBack(lbm.North(), (1, nx), (1, ny)).to_abi(), lbm.BounceBack(lbm.South(), (1, nx), (1, ny)).to_abi()]
    want, _ = COracle(qo, cm, ob).steps(f0, nsteps)
    for mode in ("f64-exact", "f32-fast"):
        dtype, arith = MODES[mode]
        with _abi.Context(nx, ny, "D2Q9", _abi.TRT, [0.8, 1.1], hb, dtype=dtype, arith=arith) as c:
            c.set_force_uniform(1e-7, 0.0)
            c.upload_f(to_host_layout(f0))
            c.step(0, nsteps)
            got = to_oracle_layout(c.download_f())
            red = c.reduce(_abi.REDUCE_CONSERVED)
        _check(O, qo, got, want, mode, "TRT")
        assert abs(red[0] - want.sum()) < 1e-9 * want.sum()
