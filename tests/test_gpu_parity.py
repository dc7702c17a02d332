"""GPU parity: liblbm_b200.so (through the C ABI / host mirror) against the oracle.

Tolerances (BASELINE.json north_star): Float64 1e-12 relative (max-norm) on populations and
moments.  In `exact` arithmetic mode SRT/TRT are additionally required to be BIT-IDENTICAL to
the oracle.  Float32 (stored as f - w) 1e-5 against the Float64 oracle.
"""
import numpy as np
import pytest

import lbm
from lbm import _abi
from conftest import random_populations, rel_max, to_host_layout, to_oracle_layout

pytestmark = pytest.mark.gpu

LATTICES = ["D2Q4", "D2Q5", "D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"]
TOL64 = 1e-12


def _models(O, qo, force):
    taus = [0.8, 0.9, 1.1, 1.3]
    return {
        "SRT": (O.SRT(0.8, force), _abi.SRT, [0.8]),
        "TRT": (O.TRT(0.8, 1.1, force), _abi.TRT, [0.8, 1.1]),
        "MRT": (O.MRT(qo, taus[:max(qo.N, 2)], force), _abi.MRT, taus[:max(qo.N, 2)]),
    }


def _bcs_pair(O, kind, nx, ny):
    """(oracle bcs, abi bcs)"""
    if kind == "none":
        return [], []
    if kind == "poiseuille":
        ob = [O.BounceBack("N", (1, nx), (1, ny)), O.BounceBack("S", (1, nx), (1, ny))]
        hb = [lbm.BounceBack(lbm.North(), (1, nx), (1, ny)), lbm.BounceBack(lbm.South(), (1, nx), (1, ny))]
    elif kind == "couette":
        ob = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0.002])]
        hb = [lbm.BounceBack(lbm.South(), (1, nx), (1, ny)), lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.01, 0.002])]
    elif kind == "cavity":
        ob = [O.BounceBack("E", (1, nx), (1, ny)), O.BounceBack("S", (1, nx), (1, ny)),
              O.BounceBack("W", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0])]
        hb = [lbm.BounceBack(lbm.East(), (1, nx), (1, ny)), lbm.BounceBack(lbm.South(), (1, nx), (1, ny)),
              lbm.BounceBack(lbm.West(), (1, nx), (1, ny)), lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.01, 0])]
    elif kind == "partial":  # walls over sub-ranges: the wrapped populations survive elsewhere
        ob = [O.BounceBack("N", (2, nx - 1), (1, ny)), O.BounceBack("W", (1, nx), (2, ny - 2)),
              O.BounceBack("E", (1, nx), (3, ny))]
        hb = [lbm.BounceBack(lbm.North(), (2, nx - 1), (1, ny)), lbm.BounceBack(lbm.West(), (1, nx), (2, ny - 2)),
              lbm.BounceBack(lbm.East(), (1, nx), (3, ny))]
    else:
        raise ValueError(kind)
    return ob, [b.to_abi() for b in hb]


def _ctx(name, model_code, taus, bcs, nx, ny, arith, dtype=_abi.F64):
    return _abi.Context(nx, ny, name, model_code, taus, bcs, dtype=dtype, arith=arith)


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("model", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("arith", [_abi.ARITH_EXACT, _abi.ARITH_FAST])
def test_collide_matches_oracle(oracle, name, model, arith):
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = 37, 11
    f0 = random_populations(qo, nx, ny)
    force = (1e-5, -2e-5)
    cm, code, taus = _models(O, qo, force)[model]
    want = O.collide(cm, qo, f0)
    with _ctx(name, code, taus, [], nx, ny, arith) as c:
        c.set_force_uniform(*force)
        c.upload_f(to_host_layout(f0))
        c.collide(0, 0.0)
        got = to_oracle_layout(c.download_f_collision())
    if arith == _abi.ARITH_EXACT and model != "MRT":
        assert np.array_equal(got, want), f"not bit-identical: max abs diff {np.abs(got - want).max()}"
    assert rel_max(got, want) < 1e-13


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (2, 7), (40, 9), (133, 6)])
def test_stream_matches_oracle_incl_tiny_grids(oracle, name, shape):
    """stream! with mod1 wrap, including grids smaller than the largest |c| (test/quadrature.jl:135-147)."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = shape
    f0 = random_populations(qo, nx, ny, seed=7)
    want = O.stream(qo, f0)
    with _ctx(name, _abi.SRT, [1.0], [], nx, ny, _abi.ARITH_EXACT) as c:
        c.upload_f(to_host_layout(f0))
        c.upload_f_collision(to_host_layout(f0))
        c.stream()
        got = to_oracle_layout(c.download_f())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("bck", ["poiseuille", "couette", "cavity", "partial"])
def test_apply_bcs_matches_oracle(oracle, name, bck):
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = 12, 9
    f_old = random_populations(qo, nx, ny, seed=3)
    f_new = O.stream(qo, f_old)
    ob, hb = _bcs_pair(O, bck, nx, ny)
    want = O.apply_bcs(ob, qo, f_new.copy(), f_old)
    with _ctx(name, _abi.SRT, [1.0], hb, nx, ny, _abi.ARITH_EXACT) as c:
        c.upload_f(to_host_layout(f_new))
        c.upload_f_collision(to_host_layout(f_old))
        c.apply_bcs(0.0)
        got = to_oracle_layout(c.download_f())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("model", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("bck", ["none", "poiseuille", "couette", "cavity", "partial"])
@pytest.mark.parametrize("arith", [_abi.ARITH_EXACT, _abi.ARITH_FAST])
def test_fused_steps_match_oracle(oracle, name, model, bck, arith):
    """lbm_step (fused pull collide+stream+BC kernels) == nsteps x (collide!, stream!, apply!)."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = 21, 10
    f0 = random_populations(qo, nx, ny, seed=11)
    force = (2e-6, 1e-6)
    cm, code, taus = _models(O, qo, force)[model]
    ob, hb = _bcs_pair(O, bck, nx, ny)
    nsteps = 7
    f = f0.copy()
    fc = None
    for s in range(nsteps):
        f, fc = O.step(cm, qo, ob, f)
    with _ctx(name, code, taus, hb, nx, ny, arith) as c:
        c.set_force_uniform(*force)
        c.upload_f(to_host_layout(f0))
        c.step(0, 3)
        c.step(3, nsteps - 3)
        got = to_oracle_layout(c.download_f())
        got_c = to_oracle_layout(c.download_f_collision())
        # and continue after a download (resume path) for one more step
        c.step(nsteps, 1)
        got2 = to_oracle_layout(c.download_f())
    f2, _ = O.step(cm, qo, ob, f)
    if arith == _abi.ARITH_EXACT and model != "MRT":
        assert np.array_equal(got, f)
        assert np.array_equal(got_c, fc)
        assert np.array_equal(got2, f2)
    assert rel_max(got, f) < TOL64
    assert rel_max(got_c, fc) < TOL64
    assert rel_max(got2, f2) < TOL64


@pytest.mark.parametrize("name,model,bck,dtype", [
    ("D2Q9", "TRT", "poiseuille", _abi.F64), ("D2Q9", "SRT", "cavity", _abi.F32), ("D2Q37", "TRT", "couette", _abi.F64),
    ("D2Q17", "MRT", "none", _abi.F64), ("D2Q13", "TRT", "partial", _abi.F32),
])
def test_graph_replay_equals_plain_launches(oracle, name, model, bck, dtype):
    """lbm_step replays captured CUDA graphs of 16 fused steps when a batch is long enough; the result must be
    bit-identical to plain launches (option graph = 0), to the oracle, and survive force / option changes."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny, nsteps = 36, 14, 53
    f0 = random_populations(qo, nx, ny, seed=3)
    force = (2e-6, 1e-6)
    cm, code, taus = _models(O, qo, force)[model]
    ob, hb = _bcs_pair(O, bck, nx, ny)
    outs = []
    for graph in (1, 0):
        with _ctx(name, code, taus, hb, nx, ny, _abi.ARITH_EXACT, dtype) as c:
            c.set_option("graph", graph)
            c.set_force_uniform(*force)
            c.upload_f(to_host_layout(f0))
            l0 = c.kernel_launches
            c.step(0, nsteps)
            assert c.kernel_launches - l0 == nsteps  # replayed launches are counted like plain ones
            a = to_oracle_layout(c.download_f())
            c.set_force_uniform(2 * force[0], force[1])  # invalidates the captured parameters
            c.step(nsteps, 40)
            b = to_oracle_layout(c.download_f())
            outs.append((a, b))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    if dtype == _abi.F64:
        f = f0
        for _ in range(nsteps):
            f, _ = O.step(cm, qo, ob, f)
        if model != "MRT":
            assert np.array_equal(outs[0][0], f)
        assert rel_max(outs[0][0], f) < TOL64



class _PinnedVelocity:
    """Minimal problem for IterativeInitializationCollisionModel: a given lattice velocity field."""

    def __init__(self, ux, uy):
        self.NY, self.NX = ux.shape
        self.u_max = 1.0
        self._u = (ux, uy)

    def grid(self):
        return np.meshgrid(np.arange(self.NX, dtype=float), np.arange(self.NY, dtype=float))

    def velocity(self, X, Y, t=0.0):
        return self._u


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("dtype", [_abi.F64, _abi.F32])
def test_iterative_initialization_collision_matches_oracle(oracle, name, dtype):
    """LBM_ITERATIVE_INIT == IterativeInitializationCollisionModel (iterative_initialization.jl:42-60): collide alone and
    fused multi-step; Float64 exact mode bit-identical."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny, nsteps = 19, 11, 21
    rng = np.random.default_rng(8)
    u0 = (0.03 * rng.uniform(-1, 1, (ny, nx)), 0.03 * rng.uniform(-1, 1, (ny, nx)))
    cm = O.IterativeInitializationCollisionModel(qo, 0.9, _PinnedVelocity(*u0))
    f0 = random_populations(qo, nx, ny, seed=4)
    want_c = O.collide(cm, qo, f0)
    f = f0
    for _ in range(nsteps):
        f, _ = O.step(cm, qo, [], f)
    with _abi.Context(nx, ny, name, _abi.ITERATIVE_INIT, [0.9], dtype=dtype, arith=_abi.ARITH_EXACT) as c:
        c.upload_f(to_host_layout(f0))
        with pytest.raises(lbm.LbmError):
            c.step(0, 1)  # no velocity field yet
        with pytest.raises(lbm.LbmError):
            c.set_force_uniform(1e-6, 0.0)  # the operator has no force
        c.set_velocity_field(u0[0].T, u0[1].T)
        c.collide()
        got_c = to_oracle_layout(c.download_f_collision())
        c.upload_f(to_host_layout(f0))
        c.step(0, nsteps)
        got = to_oracle_layout(c.download_f())
    if dtype == _abi.F64:
        assert np.array_equal(got_c, want_c) and np.array_equal(got, f)
    else:
        assert rel_max(got_c, want_c) < 1e-6 and rel_max(got, f) < 1e-5


@pytest.mark.parametrize("name,problem_kind,whole_field", [
    ("D2Q9", "tgv", False), ("D2Q9", "tgv", True), ("D2Q17", "tgv", False), ("D2Q9", "couette", False), ("D2Q37", "shear", True),
])
def test_mei_et_al_initialisation_matches_oracle(oracle, name, problem_kind, whole_field):
    """initialize(IterativeInitializationMeiEtAl(tau, eps), q, problem) (mei_et_al.jl:11-40) through the host mirror on
    the device: same number of iterations as the oracle's literal restatement and bit-identical populations."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    q = getattr(lbm.Quadratures, name)
    if problem_kind == "tgv":
        po, ph = O.TGV(qo, 0.8, 1), lbm.TGV(q, 0.8, 1)
    elif problem_kind == "couette":
        po, ph = O.CouetteFlow(1 / 6, 2), lbm.CouetteFlow(1 / 6, 2)
    else:
        po, ph = O.DecayingShearFlow(1 / 6, 2), lbm.DecayingShearFlow(1 / 6, 2)
    want, n_want = O.initialize_mei_et_al(qo, po, tau=1.0, eps=1e-9, whole_field=whole_field)
    strategy = lbm.IterativeInitializationMeiEtAl(1.0, 1e-9, whole_field=whole_field)
    got = to_oracle_layout(lbm.initialize(strategy, q, ph))
    assert 1 <= n_want < 10000 and strategy.steps_taken == n_want
    assert np.array_equal(got, want)
    # the point of the scheme: momentum pinned to the prescribed velocity, density relaxed to a consistent pressure
    fl = [got[i] for i in range(qo.Q)]
    rho = O.density(qo, fl)
    ux, uy = O.velocity(qo, fl, rho)
    X, Y = po.grid()
    vx, vy = po.velocity(X, Y)
    if problem_kind != "couette":  # (walls: the bounce-back rows do not hold the prescribed momentum)
        assert np.abs(rho * ux - po.u_max * vx).max() < 5e-3 * po.u_max
    assert abs(rho.sum() - rho.size) < 1e-9



@pytest.mark.parametrize("name", ["D2Q9", "D2Q17", "D2Q37"])
def test_force_field_and_separable(oracle, name):
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = 16, 12
    rng = np.random.default_rng(5)
    f0 = random_populations(qo, nx, ny, seed=13)
    Fx, Fy = 1e-5 * rng.standard_normal((ny, nx)), 1e-5 * rng.standard_normal((ny, nx))
    cm = O.TRT(0.7, 0.9, (Fx, Fy))
    want = f0.copy()
    for s in range(4):
        want, _ = O.step(cm, qo, [], want)
    with _ctx(name, _abi.TRT, [0.7, 0.9], [], nx, ny, _abi.ARITH_EXACT) as c:
        c.set_force_field(Fx.T, Fy.T)
        c.upload_f(to_host_layout(f0))
        c.step(0, 4)
        got = to_oracle_layout(c.download_f())
    assert np.array_equal(got, want)
    # separable, time dependent: F = (gx[t][y], gy[t][x])
    nst = 5
    gx, gy = 1e-5 * rng.standard_normal((nst, ny)), 1e-5 * rng.standard_normal((nst, nx))
    cm = O.SRT(0.9, lambda t: (np.repeat(gx[int(round(t))][:, None], nx, 1), np.repeat(gy[int(round(t))][None, :], ny, 0)))
    want = f0.copy()
    for s in range(nst):
        want, _ = O.step(cm, qo, [], want, time=float(s))
    with _ctx(name, _abi.SRT, [0.9], [], nx, ny, _abi.ARITH_EXACT) as c:
        c.set_force_separable(0, gx, gy)
        c.upload_f(to_host_layout(f0))
        c.step(0, nst, 1.0)
        got = to_oracle_layout(c.download_f())
        with pytest.raises(lbm.LbmError):
            c.step(nst, 1, 1.0)  # outside the force table
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", LATTICES)
def test_moments_and_reductions(oracle, name):
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = 19, 8
    f0 = random_populations(qo, nx, ny, seed=17)
    fl = [f0[i] for i in range(qo.Q)]
    rho = O.density(qo, fl)
    ux, uy = O.velocity(qo, fl, rho)
    p = O.pressure(qo, fl, rho, ux, uy)
    tau = 0.3
    sig = O.deviatoric_tensor(qo, tau, fl, rho, ux, uy)

    class P:  # minimal problem for hydrodynamic_fields
        u_max = 1.0

        def lattice_viscosity(self):
            return tau / qo.css
    h = O.hydrodynamic_fields(qo, P(), f0)
    with _ctx(name, _abi.SRT, [0.8], [], nx, ny, _abi.ARITH_EXACT) as c:
        c.upload_f(to_host_layout(f0))
        m = c.moments(tau, _abi.Context.FIELDS)
        mean = c.reduce(_abi.REDUCE_MEAN_UX)
        cons = c.reduce(_abi.REDUCE_CONSERVED)
        vc1 = c.reduce(_abi.REDUCE_VELOCITY_CHANGE)
        vc2 = c.reduce(_abi.REDUCE_VELOCITY_CHANGE)
        # same fields straight from the post-collision state (pull path of the diagnostics)
        c.step(0, 1)
        m_pull = c.moments(tau, ("rho", "ux", "uy"))
        f1 = to_oracle_layout(c.download_f())
    assert np.array_equal(m["rho"].T, rho)
    assert np.array_equal(m["ux"].T, ux) and np.array_equal(m["uy"].T, uy)
    assert rel_max(m["p"].T, p) < 1e-14
    assert rel_max(m["p_track"].T, h["p"]) < 1e-14
    for k, key in (("sxx", (0, 0)), ("sxy", (0, 1)), ("syy", (1, 1))):
        assert np.max(np.abs(m[k].T - sig[key])) < 1e-15 * max(1.0, np.max(np.abs(sig[key])) / 1e-3)
    assert abs(mean[0] - ux.sum()) <= 1e-12 * np.abs(ux).sum() and mean[1] == nx * ny and mean[2] == 0
    assert abs(cons[0] - rho.sum()) < 1e-12 * rho.sum()
    assert abs(vc1[0] - (ux ** 2 + uy ** 2).sum()) < 1e-12 * (ux ** 2 + uy ** 2).sum() and vc1[1] == 0
    assert vc2[0] == 0 and abs(vc2[1] - (ux ** 2 + uy ** 2).sum()) < 1e-12 * (ux ** 2 + uy ** 2).sum()
    fl1 = [f1[i] for i in range(qo.Q)]
    rho1 = O.density(qo, fl1)
    assert np.array_equal(m_pull["rho"].T, rho1)
    u1 = O.velocity(qo, fl1, rho1)
    assert np.array_equal(m_pull["ux"].T, u1[0])


def test_config_c1_shear_wave_1000_steps(oracle):
    """BASELINE config 1: D2Q9 SRT decaying shear wave, 64x64 periodic, tau = 1, 1000 steps."""
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.D2Q9()
    pr = O.DecayingShearFlow(1 / 6, u_max=0.02 / 8, NX=64, NY=64, static=False, convenience=False)
    f0 = O.initialize("AnalyticalEquilibrium", qo, pr)
    cm = O.collision_model("SRT", qo, pr)
    assert cm.tau == 1.0 and cm.force is None
    want, _ = COracle(qo, cm).steps(f0, 1000)
    q = lbm.D2Q9()
    hp = lbm.DecayingShearFlow.fields(1.0, 0.02 / 8, 1 / 6, 64, 64, (2 * np.pi, 2 * np.pi), False, 1.0, 1.0, 1.0, 0.0)
    f0h = lbm.initialize(lbm.AnalyticalEquilibrium(), q, hp)
    assert rel_max(to_oracle_layout(f0h), f0) < 1e-15
    model = lbm.LatticeBoltzmannModel(hp, q, collision_model=lbm.SRT, process_method=lbm.TrackHydrodynamicErrors(hp, False, 1000))
    model.f_stream = to_host_layout(f0)
    lbm.simulate(model, range(0, 1000))
    got = to_oracle_layout(model.f_stream)
    assert np.array_equal(got, want)
    # diagnostics row recorded at t == n_steps == 1000 (after 1000 steps)
    pm = O.TrackHydrodynamicErrors(pr, False, 1000)
    pm.next(qo, want, 1000)
    row, ref = model.processing_method.df[-1], pm.df[-1]
    for k in ("error_rho", "error_u", "error_p", "error_sxy", "mass", "momentum", "energy"):
        assert abs(row[k] - ref[k]) <= 1e-9 * abs(ref[k]) + 1e-300, (k, row[k], ref[k])
    model.close()


@pytest.mark.parametrize("name,every", [("D2Q9", 10), ("D2Q13", [1, 7, 33, 40])])
def test_lid_driven_cavity_with_snapshots(oracle, name, every):
    """LidDrivenCavityFlow (East/South/West bounce-back + MovingWall North, last writer wins at
    the top corners, lid_driven_cavity.jl:61-66) driven through simulate(model, time) with TakeSnapshots
    (take_snapshots.jl:12-29); every snapshot bit-identical to the oracle's."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    q = getattr(lbm.Quadratures, name)
    po, ph = O.LidDrivenCavityFlow(1 / 6, 1), lbm.LidDrivenCavityFlow(1 / 6, 1)
    pmo = O.TakeSnapshots(po, every)
    mo = O.make_model(po, qo, "TRT", "ZeroVelocityInitialCondition", pmo)
    O.simulate_model(mo, range(0, 41))
    pm = lbm.TakeSnapshots(ph, every)
    model = lbm.LatticeBoltzmannModel(ph, q, collision_model=lbm.TRT, initialization_strategy=lbm.ZeroVelocityInitialCondition(),
                                      process_method=pm)
    lbm.simulate(model, range(0, 41))
    assert pm.timesteps == pmo.timesteps and len(pm.snapshots) == len(pmo.snapshots) >= 4
    for a, b in zip(pm.snapshots, pmo.snapshots):
        assert np.array_equal(to_oracle_layout(a), b)
    assert np.array_equal(to_oracle_layout(model.f_stream), mo.f_stream)
    # the lid drags the fluid: x-momentum appears under the moving wall
    fl = [mo.f_stream[i] for i in range(qo.Q)]
    rho = O.density(qo, fl)
    ux, _ = O.velocity(qo, fl, rho)
    assert ux[-1].mean() > 0
    model.close()


@pytest.mark.parametrize("row", [0, 1, 9, 29, 30, 40])
def test_golden_table_rows_through_gpu(row):
    """The reference's own golden vectors (examples/notebooks/trt_magic_parameter.ipynb:109-176)
    reproduced end to end through the host mirror + CUDA library."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "trt_magic_parameter.json")))["rows"][row]
    q = lbm.D2Q9()
    ts, ta = g["tau_s"], g["tau_a"]
    problem = lbm.PoiseuilleFlow((ts - 0.5) / q.speed_of_sound_squared, 1)
    n_steps = round(100.0 / problem.delta_t())
    pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.VelocityConvergenceStoppingCriteria(1e-7, problem))
    cm = lbm.TRT(ts, ta, lambda x, y, t: lbm.lattice_force(problem, x, y, t))
    cm.force = lbm.LatticeForce(problem)  # same closure, with the uniform fast path
    res = lbm.simulate(problem, q, t_end=100.0, should_process=False, collision_model=cm, process_method=pm,
                       initialization_strategy=lbm.ZeroVelocityInitialCondition())
    e = res.processing_method.df[-1]
    for k in ("error_u", "error_p", "error_sxy"):
        if k in g:
            assert abs(e[k] - g[k]) <= 1.01 * 10 ** (np.floor(np.log10(abs(g[k]))) - 5), (k, e[k], g[k])
    if "error_sxx" in g:
        assert np.isinf(e["error_sxx"])
    res.close()


def test_array_level_operators():
    """collide!/stream!/apply! called on plain arrays as the reference's tests do (test/quadrature.jl:64-147)."""
    for q in lbm.Quadratures:
        f = lbm.equilibrium(q, 1.0, np.array([0.1, 0.1]), 1.0).reshape(1, 1, q.Q)
        f_in, f_out = np.asfortranarray(f.copy()), np.asfortranarray(f.copy())
        lbm.collide_(lbm.SRT(1.0), q, f_old=f_in, f_new=f_out, time=0.0)
        assert np.allclose(f_in, f_out, atol=1e-4)
        f_inn, f_o = np.asfortranarray(f.copy()), np.asfortranarray(f.copy())
        for t in range(4):
            lbm.collide_(lbm.SRT(1.0), q, f_old=f_inn, f_new=f_o, time=0.0)
            lbm.stream_(q, f_new=f_inn, f_old=f_o)
        assert np.allclose(f, f_inn, atol=1e-5)
        feq = lbm.equilibrium(q, 1.0, np.array([0.0, 0.0]), 1.0).reshape(1, 1, q.Q)
        f_new = np.asfortranarray(feq.copy())
        lbm.stream_(q, f_new=f_new, f_old=np.asfortranarray(feq))
        assert np.allclose(feq, f_new)


@pytest.mark.parametrize("name,shape", [("D2Q9", (5, 4)), ("D2Q9", (1, 1)), ("D2Q13", (2, 6)), ("D2Q37", (3, 3)), ("D2Q37", (9, 4))])
def test_push_stream_variant(oracle, name, shape):
    """stream(q, f, f_new) -- the scatter variant with single wrap (stream.jl:6-16, 44-61) -- equals the oracle's literal
    restatement wherever the reference is defined (grid >= widest velocity); smaller grids are rejected."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    q = getattr(lbm.Quadratures, name)
    nx, ny = shape
    f = random_populations(qo, nx, ny, seed=2)
    got = to_oracle_layout(lbm.stream(q, to_host_layout(f)))
    assert np.array_equal(got, O.stream_push(qo, f))
    h = int(np.abs(q.abscissae).max())
    if h > 1:
        with pytest.raises(ValueError):
            lbm.stream(q, to_host_layout(random_populations(qo, h - 1, 4)))


def test_trt_equals_srt_and_mrt_only_at_tau_1():
    """test/collision_models.jl:3-114."""
    q = lbm.D2Q9()
    f_in = np.asfortranarray(np.ones((10, 10, 1)) * q.weights)
    f_srt = lbm.collide_(lbm.SRT(0.8), q, f_in)
    f_trt = lbm.collide_(lbm.TRT(0.8, 0.8), q, f_in)
    f_mrt = lbm.collide_(lbm.MRT(q, 0.8), q, f_in)
    assert np.allclose(f_srt, f_trt, rtol=1e-12) and np.allclose(f_srt, f_mrt, rtol=1e-12)
    for tau in [0.51, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1]:
        nu = (tau - 0.5) / q.speed_of_sound_squared
        res = {}
        for key, cm in (("srt", lbm.SRT(tau)), ("trt", lbm.TRT(tau, tau)), ("mrt", lbm.MRT(q, tau))):
            problem = lbm.PoiseuilleFlow(nu, 1, static=False)
            m = lbm.LatticeBoltzmannModel(problem, q, collision_model=cm, process_method=lbm.ProcessingMethod(problem, False, 10))
            lbm.simulate(m, range(0, 11))
            res[key] = m.f_stream
            m.close()
        assert np.allclose(res["srt"], res["trt"], rtol=1e-10)
        if tau == 1.0:
            assert np.allclose(res["srt"], res["mrt"], rtol=1e-10)
        # with bounce-back walls and no force the state stays at rest, where MRT == SRT too;
        # the reference marks tau != 1 as @test_broken only because of round-off level differences.


def test_mrt_differs_from_srt_off_equilibrium(oracle):
    """Regularisation != BGK for tau != 1 on non-equilibrium input: the GPU must reproduce the
    oracle's MRT (not collapse to SRT)."""
    O = oracle
    qo = O.L.D2Q9()
    f0 = random_populations(qo, 8, 8, seed=21, amp=0.05)
    want = O.collide(O.MRT(qo, 0.6), qo, f0)
    srt = O.collide(O.SRT(0.6), qo, f0)
    assert rel_max(want, srt) > 1e-4
    got = to_oracle_layout(lbm.collide_(lbm.MRT(lbm.D2Q9(), 0.6), lbm.D2Q9(), to_host_layout(f0)))
    assert rel_max(got, want) < 1e-13


def test_large_grid_properties():
    """Size-independent properties at a bench-size grid (4096 x 4096 TGV, D2Q9 TRT): mass is
    conserved to round-off, momentum decays, and exact == fast within tolerance."""
    q = lbm.D2Q9()
    n = 4096
    problem = lbm.TGV(q, 0.8, n // 16)
    out = {}
    for arith in ("exact", "fast"):
        m = lbm.LatticeBoltzmannModel(problem, q, collision_model=lbm.TRT, process_method=None, arith=arith)
        c0 = m.state.reduce(_abi.REDUCE_CONSERVED)
        m.state.step(0, 50, 1.0)
        c1 = m.state.reduce(_abi.REDUCE_CONSERVED)
        assert abs(c1[0] - c0[0]) < 1e-12 * c0[0]
        assert c1[2] < c0[2]
        out[arith] = m.ctx.moments(1.0, ("ux",))["ux"]
        m.close()
    assert rel_max(out["fast"], out["exact"]) < 1e-9


def test_error_paths():
    with pytest.raises(lbm.LbmError):  # MovingWall only exists for North (moving_wall.jl:17)
        _abi.Context(8, 8, "D2Q9", _abi.SRT, [1.0], [lbm.MovingWall(lbm.South(), (1, 8), (1, 8), [0.1, 0]).to_abi()])
    with pytest.raises(lbm.LbmError):
        _abi.Context(8, 8, "D2Q9", _abi.TRT, [1.0])  # TRT needs two relaxation times
    with _abi.Context(8, 8, "D2Q9", _abi.SRT, [1.0]) as c:
        with pytest.raises(lbm.LbmError):
            c.stream()  # stream! before collide!
        with pytest.raises(ValueError):
            c.upload_f(np.zeros((8, 8, 5)))


# ---------------------------------------------------------------------------------------------
# Float32 (new capability: the reference is Float64-only).  Storage = f - w, arithmetic on deviations.
# Tolerance from BASELINE.json north_star: 1e-5 relative (max-norm) on populations and moments
# against the Float64 oracle.
# ---------------------------------------------------------------------------------------------
TOL32 = 1e-5


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("model", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("arith", [_abi.ARITH_EXACT, _abi.ARITH_FAST])
def test_f32_steps_match_f64_oracle(oracle, name, model, arith):
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny, nsteps = 24, 16, 20
    # smooth Taylor-Green-like start, u ~ 1e-3
    pr = O.TGV(qo, 0.8, 1, nx, ny, u_max=2e-3)
    f0 = O.initialize("AnalyticalEquilibrium", qo, pr)
    force = (1e-6, -1e-6)
    cm, code, taus = _models(O, qo, force)[model]
    ob, hb = _bcs_pair(O, "couette" if name in ("D2Q9", "D2Q37") else "none", nx, ny)
    want = f0.copy()
    for _ in range(nsteps):
        want, _ = O.step(cm, qo, ob, want)
    with _ctx(name, code, taus, hb, nx, ny, arith, dtype=_abi.F32) as c:
        c.set_force_uniform(*force)
        c.upload_f(to_host_layout(f0))
        back = to_oracle_layout(c.download_f())
        c.step(0, nsteps)
        got = to_oracle_layout(c.download_f())
        m = c.moments(0.3, ("rho", "ux", "uy"))
    # storage round trip keeps the deviation to Float32 precision
    dev = np.abs(f0 - qo.w[:, None, None]).max()
    assert np.abs(back - f0).max() <= 1.2e-7 * dev + 1e-12
    assert rel_max(got, want) < TOL32
    fl = [want[i] for i in range(qo.Q)]
    rho = O.density(qo, fl)
    ux, uy = O.velocity(qo, fl, rho)
    assert rel_max(m["rho"].T, rho) < TOL32
    scale = max(np.abs(ux).max(), np.abs(uy).max())
    assert np.abs(m["ux"].T - ux).max() < TOL32 * scale and np.abs(m["uy"].T - uy).max() < TOL32 * scale


def test_f32_config_c2_crop_100_steps(oracle):
    """BASELINE config 2 physics (D2Q9 TRT TGV, tau = 0.8, Lambda = 1/4) on a 256^2 crop, 100 steps:
    Float32 velocities within 1e-5 of the Float64 oracle; Float64 within 1e-12 (bit-identical here)."""
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.D2Q9()
    n = 256
    pr = O.TGV(qo, 0.8, n // 16)
    f0 = O.initialize("AnalyticalEquilibrium", qo, pr)
    cm = O.collision_model("TRT", qo, pr)
    want, _ = COracle(qo, cm).steps(f0, 100)
    fl = [want[i] for i in range(qo.Q)]
    rho = O.density(qo, fl)
    ux, uy = O.velocity(qo, fl, rho)
    q = lbm.D2Q9()
    hp = lbm.TGV(q, 0.8, n // 16)
    for dtype, tol in (("f64", 1e-12), ("f32", TOL32)):
        for arith in ("exact", "fast"):
            m = lbm.LatticeBoltzmannModel(hp, q, collision_model=lbm.TRT, dtype=dtype, arith=arith)
            m.f_stream = to_host_layout(f0)
            m.state.step(0, 100, 1.0)
            got = to_oracle_layout(m.f_stream)
            mo = m.ctx.moments(1.0, ("ux", "uy"))
            m.close()
            assert rel_max(got, want) < tol, (dtype, arith)
            assert rel_max(mo["ux"].T, ux) < tol and rel_max(mo["uy"].T, uy) < tol, (dtype, arith)
            if dtype == "f64" and arith == "exact":
                assert np.array_equal(got, want)


def test_row_chunked_upload_download(oracle):
    """lbm_upload_f_rows / lbm_download_f_rows == the whole-array calls (both dtypes)."""
    O = oracle
    qo = O.L.D2Q13()
    nx, ny = 20, 17
    f0 = random_populations(qo, nx, ny, seed=31)
    host = to_host_layout(f0)
    for dtype in (_abi.F64, _abi.F32):
        with _ctx("D2Q13", _abi.TRT, [0.8, 1.1], [], nx, ny, _abi.ARITH_EXACT, dtype=dtype) as a, \
                _ctx("D2Q13", _abi.TRT, [0.8, 1.1], [], nx, ny, _abi.ARITH_EXACT, dtype=dtype) as b:
            a.upload_f(host)
            for y0, n in ((0, 5), (5, 1), (6, 11)):
                b.upload_f_rows(y0, host[:, y0:y0 + n, :])
            a.step(0, 4)
            b.step(0, 4)
            fa = a.download_f()
            assert np.array_equal(fa, b.download_f())
            parts = [b.download_f_rows(y0, n) for y0, n in ((0, 9), (9, 8))]
            assert np.array_equal(np.concatenate(parts, axis=1), fa)
            with pytest.raises(lbm.LbmError):
                b.download_f_rows(10, 8)


INIT_STRATEGIES = ["ZeroVelocityInitialCondition", "AnalyticalEquilibrium", "ConstantDensity", "AnalyticalVelocityAndStress",
                   "AnalyticalEquilibriumAndOffEquilibrium"]


def _init_problems(q, qo, O):
    sh = (1.0, 0.05, 1 / 6, 12, 10, (2 * np.pi, 2 * np.pi), False, 1.0, 1.0, 1.0, 1.0)
    return [
        (lbm.TGV(q, 0.8, 1, 12, 10), O.TGV(qo, 0.8, 1, 12, 10)),
        (lbm.DecayingShearFlow.fields(*sh),
         O.DecayingShearFlow(1 / 6, NX=12, NY=10, static=False, A=1.0, B=1.0, k_x=1.0, k_y=1.0, u_max=0.05, convenience=False)),
        (lbm.TaylorGreenVortex(1 / 6, 1, 16, 16, static=False), O.TaylorGreenVortex(1 / 6, 1, 16, 16, static=False)),
        (lbm.PoiseuilleFlow(1 / 6, 2), O.PoiseuilleFlow(1 / 6, 2)),
        (lbm.CouetteFlow(1 / 6, 2), O.CouetteFlow(1 / 6, 2)),
        (lbm.LinearizedTransverseShearWave(1 / 6, 1 / 6, 2), O.LinearizedTransverseShearWave(1 / 6, 1 / 6, 2)),
    ]


@pytest.mark.parametrize("name", LATTICES)
def test_device_side_initialisation(oracle, name):
    """initialize(strategy, q, problem) evaluated on the device == the ORACLE's initialize (initial_conditions.jl:7-22,
    analytical_offequilibrium.jl:10-87, analytical_velocity_stress.jl:5-31; hermite_based_equilibrium! incl. T != 1 and
    the Val{4} quirk on D2Q37) for every closed-form strategy, through both device paths: lbm_init_analytic (separable
    tables, the kernel evaluates rho, u, T and the off-equilibrium part) and lbm_init_equilibrium_rows (host rows)."""
    O = oracle
    q, qo = getattr(lbm.Quadratures, name), O.L.BY_NAME[name]()
    from lbm.initial_conditions import initialize_on_device
    for ph, po in _init_problems(q, qo, O):
        for sname in INIT_STRATEGIES:
            strategy = getattr(lbm, sname)()
            want = np.asfortranarray(np.transpose(O.initialize(sname, qo, po), (2, 1, 0)))
            dev = np.abs(want - q.weights).max() + 1e-30
            for dtype in ("f64", "f32"):
                for analytic in (True, False):
                    if not analytic and sname.startswith("AnalyticalVelocityAndStress") or (not analytic and "OffEq" in sname):
                        continue
                    m = lbm.LatticeBoltzmannModel(ph, q, collision_model=lbm.SRT(0.8), initialization_strategy=lbm.ZeroVelocityInitialCondition(),
                                                  dtype=dtype)
                    launches = m.ctx.kernel_launches
                    assert initialize_on_device(strategy, q, ph, m.ctx, analytic=analytic)
                    assert m.ctx.kernel_launches > launches
                    got = m.f_stream
                    m.close()
                    tag = (name, type(ph).__name__, sname, dtype, analytic)
                    if dtype == "f64":
                        assert np.abs(got - want).max() <= 1e-14 * np.abs(want).max(), tag
                    else:
                        assert np.abs(got - want).max() <= 2e-7 * dev + 1e-12, tag


def test_device_side_initialisation_large_grid_and_model_flag(oracle):
    """device_init=True on the model takes the analytic path (no host evaluation of the grid); 2048 x 1024 TGV with the
    off-equilibrium part against the oracle."""
    O = oracle
    q, qo = lbm.D2Q9(), O.L.D2Q9()
    ph, po = lbm.TGV(q, 0.8, 128, 2048, 1024), O.TGV(qo, 0.8, 128, 2048, 1024)
    want = np.transpose(O.initialize("AnalyticalEquilibriumAndOffEquilibrium", qo, po), (2, 1, 0))
    m = lbm.LatticeBoltzmannModel(ph, q, collision_model=lbm.TRT, initialization_strategy=lbm.AnalyticalEquilibriumAndOffEquilibrium(),
                                  device_init=True)
    got = m.f_stream
    m.close()
    assert np.abs(got - want).max() <= 1e-14 * np.abs(want).max()


# ---------------------------------------------------------------------------------------------
# The reference's convergence studies (examples/notebooks/notebook_examples.jl:34-69, shear_wave.ipynb,
# taylor_green_vortex.ipynb, couette.ipynb) run end to end through the host mirror + CUDA library:
# the quadratic convergence of the method must survive (north_star), and the error values must equal
# the oracle's.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", LATTICES)
def test_shear_wave_convergence_is_second_order(oracle, name):
    O = oracle
    q = getattr(lbm.Quadratures, name)
    qo = O.L.BY_NAME[name]()
    tau_lb = 0.8
    errs, errs_o = [], []
    scales = [1, 2, 4, 8] if name == "D2Q9" else [1, 2, 4]
    for scale in scales:
        nu = tau_lb / (2.0 * q.speed_of_sound_squared)
        problem = lbm.DecayingShearFlow(nu, scale, static=True)
        n_steps = round(1.0 / problem.delta_t())
        pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.NoStoppingCriteria())
        res = lbm.simulate(problem, q, process_method=pm, initialization_strategy=lbm.AnalyticalEquilibrium(), t_end=1.0)
        errs.append(res.processing_method.df[-1]["error_u"])
        res.close()
        if scale <= 2:
            po = O.DecayingShearFlow(nu, scale, static=True)
            mo = O.simulate(po, qo, pm=O.TrackHydrodynamicErrors(po, False, n_steps, O.NoStoppingCriteria()), t_end=1.0)
            errs_o.append(mo.pm.df[-1]["error_u"])
    for a, b in zip(errs, errs_o):
        assert abs(a - b) <= 1e-9 * abs(b), (errs, errs_o)
    slope = np.polyfit(np.log([8.0 * s for s in scales]), np.log(errs), 1)[0]
    if name in ("D2Q4", "D2Q5"):
        # first-order lattices: the reference algorithm itself does not converge on this problem
        # (oracle: error_u stays at 4.4 / 1.55); what must hold is agreement with the oracle (above)
        assert abs(slope) < 0.1
    else:
        assert slope <= -1.8, (name, errs, slope)


@pytest.mark.parametrize("model", ["SRT", "TRT", "MRT"])
def test_tgv_decay_convergence_is_second_order(model):
    q = lbm.D2Q9()
    cm = {"SRT": lbm.SRT, "TRT": lbm.TRT, "MRT": lbm.MRT}[model]
    errs = []
    for scale in (1, 2, 4):
        problem = lbm.TGV(q, 0.8, scale, 8 * scale, 8 * scale)
        t_end = round(lbm.decay_time(problem))
        pm = lbm.TrackHydrodynamicErrors(problem, False, t_end, lbm.NoStoppingCriteria())
        model_ = lbm.LatticeBoltzmannModel(problem, q, collision_model=cm, process_method=pm)
        lbm.simulate(model_, range(1, t_end + 1))
        errs.append(pm.df[-1]["error_u"])
        model_.close()
    slope = np.polyfit(np.log([1.0, 2.0, 4.0]), np.log(errs), 1)[0]
    assert slope <= -1.8, (model, errs, slope)


@pytest.mark.parametrize("name", ["D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"])
def test_couette_moving_wall_steady_state(oracle, name):
    """couette.ipynb: CouetteFlow (MovingWall North + BounceBack South) run to the stop criterion;
    same stopping step and same error as the oracle."""
    O = oracle
    q = getattr(lbm.Quadratures, name)
    qo = O.L.BY_NAME[name]()
    nu = 0.3 / q.speed_of_sound_squared
    problem = lbm.CouetteFlow(nu, 2)
    n_steps = 3000
    pm = lbm.ProcessingMethod(problem, False, n_steps)
    model = lbm.LatticeBoltzmannModel(problem, q, collision_model=lbm.TRT, process_method=pm,
                                      initialization_strategy=lbm.ZeroVelocityInitialCondition())
    lbm.simulate(model, range(0, n_steps + 1))
    po = O.CouetteFlow(nu, 2)
    mo = O.make_model(po, qo, "TRT", strategy="ZeroVelocityInitialCondition", pm=O.processing_method(po, False, n_steps))
    O.simulate_model(mo, range(0, n_steps + 1))
    got, want = pm.df[-1], mo.pm.df[-1]
    assert len(pm.df) == len(mo.pm.df)
    for k in want:  # all 16 columns of process! (processing_methods.jl:241-262); the 12 sums come from lbm_reduce_process
        assert abs(got[k] - want[k]) <= 1e-9 * abs(want[k]) + 1e-14, (name, k, got[k], want[k])
    assert rel_max(to_oracle_layout(model.f_stream), mo.f_stream) < 1e-12
    model.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_asynchronous_snapshots_do_not_disturb_the_run(oracle, dtype):
    """lbm_snapshot_begin / _end (TakeSnapshots, take_snapshots.jl:12-29): every snapshot equals the oracle's f_stream at
    that step, the copy overlaps the following steps, and the final state equals the one of a run that took no snapshots
    (the snapshot kernel reads the post-collision state through the pull, it does not materialise f_stream)."""
    O = oracle
    q, qo = lbm.D2Q13(), O.L.D2Q13()
    nx, ny = 96, 40
    ph, po = lbm.CouetteFlow.fields(1.0, 0.01, 1 / 6, nx, ny, (1.0, 1.0)), O.CouetteFlow(1 / 6, NX=nx, NY=ny, u_max=0.01, convenience=False)
    f0 = O.initialize("ZeroVelocityInitialCondition", qo, po)
    cmo = O.collision_model("TRT", qo, po)
    want, f = {}, f0
    for t in range(1, 31):
        f, _ = O.step(cmo, qo, po.boundary_conditions(), f)
        want[t] = f
    tol = 0 if dtype == "f64" else 1e-6
    m = lbm.LatticeBoltzmannModel(ph, q, collision_model=lbm.TRT, initialization_strategy=lbm.ZeroVelocityInitialCondition(), dtype=dtype)
    ctx = m.ctx
    ctx.step(0, 7)
    l0 = ctx.kernel_launches
    a = ctx.snapshot_begin()                      # page-locked array from the library
    assert ctx.kernel_launches == l0 + 1           # one kernel, no materialisation
    ctx.step(7, 13)                                # enqueued while the copy is in flight
    ctx.snapshot_end()
    assert np.abs(to_oracle_layout(a) - want[7]).max() <= tol
    b = ctx.new_f()                                # pageable array of the caller: staged through the library's buffer
    ctx.snapshot_begin(b)
    c = ctx.snapshot_begin()                       # a second begin completes the first
    assert np.abs(to_oracle_layout(b) - want[20]).max() <= tol
    ctx.step(20, 10)
    ctx.snapshot_end()
    assert np.array_equal(b, c)
    got = to_oracle_layout(ctx.download_f())
    assert np.abs(got - want[30]).max() <= tol
    m.close()
    assert np.abs(to_oracle_layout(a) - want[7]).max() <= tol  # page-locked arrays outlive the context


@pytest.mark.parametrize("kind", ["LinearizedThermalDiffusion", "LinearizedTransverseShearWave"])
@pytest.mark.parametrize("name,model", [("D2Q9", "SRT"), ("D2Q9", "MRT"), ("D2Q17", "TRT"), ("D2Q37", "MRT")])
def test_linearized_hydrodynamic_modes(oracle, kind, name, model):
    """The Linearized* problems (src/problems/linear_hydrodynamics_modes.jl; Shan & Chen's linear modes: a density /
    temperature wave with T = rho_0 theta_0 / rho != 1, and a transverse shear wave) through simulate(problem, q) on the
    device against the oracle: populations after every batch boundary and every row of the statistics."""
    O = oracle
    q, qo = getattr(lbm.Quadratures, name), O.L.BY_NAME[name]()
    cm = {"SRT": lbm.SRT, "TRT": lbm.TRT, "MRT": lbm.MRT}[model]
    nu = 0.4 / q.speed_of_sound_squared
    ph, po = getattr(lbm, kind)(nu, nu, 2), getattr(O, kind)(nu, nu, 2)
    n_steps = 120
    pm = lbm.ProcessingMethod(ph, True, n_steps)
    assert isinstance(pm, lbm.CompareWithAnalyticalSolution)
    m = lbm.LatticeBoltzmannModel(ph, q, collision_model=cm, process_method=pm)
    lbm.simulate(m, range(0, n_steps + 1))
    mo = O.make_model(po, qo, model, pm=O.processing_method(po, True, n_steps))
    O.simulate_model(mo, range(0, n_steps + 1))
    assert rel_max(to_oracle_layout(m.f_stream), mo.f_stream) < 1e-12
    assert len(pm.df) == len(mo.pm.df)
    for got, want in zip(pm.df, mo.pm.df):
        for k in want:
            # sums that vanish analytically (the momentum of a standing wave) are round-off on either side: O(1e-12) of
            # the O(100) total density
            assert abs(got[k] - want[k]) <= 1e-10 * abs(want[k]) + 1e-10, (kind, name, model, k, got[k], want[k])
    m.close()
    # Float32 storage: populations within 1e-5 of the oracle
    m = lbm.LatticeBoltzmannModel(ph, q, collision_model=cm, process_method=lbm.ProcessingMethod(ph, False, n_steps), dtype="f32")
    lbm.simulate(m, range(0, n_steps + 1))
    assert rel_max(to_oracle_layout(m.f_stream), mo.f_stream) < 1e-5
    m.close()


@pytest.mark.parametrize("name", ["D2Q4", "D2Q9", "D2Q21", "D2Q37"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_process_sums_on_device_match_oracle(oracle, name, dtype):
    """process! with should_process = true (CompareWithAnalyticalSolution, processing_methods.jl:177-262): every row of
    the statistics comes from ONE device reduction (lbm_reduce_process) and equals the oracle's row; the launch count
    shows no field download (k_moments) happened."""
    O = oracle
    q, qo = getattr(lbm.Quadratures, name), O.L.BY_NAME[name]()
    nu = 0.3 / q.speed_of_sound_squared
    for mk_h, mk_o in ((lambda: lbm.PoiseuilleFlow(nu, 2), lambda: O.PoiseuilleFlow(nu, 2)),
                       (lambda: lbm.CouetteFlow(nu, 2), lambda: O.CouetteFlow(nu, 2))):
        problem, po = mk_h(), mk_o()
        n_steps = 40
        pm = lbm.CompareWithAnalyticalSolution(problem, True, n_steps, lbm.NoStoppingCriteria())
        model = lbm.LatticeBoltzmannModel(problem, q, collision_model=lbm.TRT, process_method=pm, dtype=dtype,
                                          initialization_strategy=lbm.ZeroVelocityInitialCondition())
        lbm.simulate(model, range(0, n_steps + 1))
        mo = O.make_model(po, qo, "TRT", strategy="ZeroVelocityInitialCondition",
                          pm=O.CompareWithAnalyticalSolution(po, True, n_steps, O.NoStoppingCriteria()))
        O.simulate_model(mo, range(0, n_steps + 1))
        assert len(pm.df) == len(mo.pm.df) == n_steps + 2  # one row per loop pass + the final next!(last(time) + 1)
        tol = 1e-10 if dtype == "f64" else 2e-5
        for got, want in zip(pm.df, mo.pm.df):
            for k in want:
                assert abs(got[k] - want[k]) <= tol * abs(want[k]) + (1e-14 if dtype == "f64" else 1e-9), (name, k, got[k], want[k])
        # host-path rows (fields downloaded, numpy sums) agree as well
        pm2 = lbm.CompareWithAnalyticalSolution(problem, True, 3, lbm.NoStoppingCriteria())
        from lbm.processing_methods import process_
        rows = []
        process_(problem, q, model.state, 0.0, rows, device_sums=True)
        process_(problem, q, model.state, 0.0, rows, device_sums=False)
        for k in rows[0]:
            assert abs(rows[0][k] - rows[1][k]) <= 1e-11 * abs(rows[1][k]) + 1e-14, (k, rows[0][k], rows[1][k])
        model.close()


@pytest.mark.parametrize("name", ["D2Q9", "D2Q17", "D2Q37"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_device_error_norms_match_host_path(name, dtype):
    """lbm_reduce_errors (16 sums on the device, separable analytic fields) == the field-download path."""
    q = getattr(lbm.Quadratures, name)
    cases = [(lbm.TGV(q, 0.8, 1, 24, 16), lbm.TRT), (lbm.TaylorGreenVortex(1 / 6, 1, 16, 16), lbm.SRT),
             (lbm.DecayingShearFlow.fields(1.0, 0.01, 1 / 6, 16, 12, (2 * np.pi, 2 * np.pi), True, 1.0, 1.0, 1.0, 1.0), lbm.SRT),
             (lbm.PoiseuilleFlow.fields(1.0, 0.02, 1 / 6, 6, 12, 1.0, (1.0, 1.0), 1.0), lbm.TRT),
             (lbm.CouetteFlow.fields(1.0, 0.01, 1 / 6, 4, 10, (1.0, 1.0)), lbm.SRT)]
    for problem, cm in cases:
        rows = []
        for device_norms in (True, False):
            pm = lbm.TrackHydrodynamicErrors(problem, False, 25, lbm.NoStoppingCriteria(), device_norms=device_norms)
            m = lbm.LatticeBoltzmannModel(problem, q, collision_model=cm, process_method=pm, dtype=dtype)
            lbm.simulate(m, range(0, 25))
            m.close()
            rows.append(pm.df[-1])
        a, b = rows
        for k, v in b.items():
            if np.isfinite(v) and v != 0:
                # (sums that vanish analytically, e.g. the TGV momentum, are pure round-off ~1e-18)
                assert abs(a[k] - v) <= 1e-9 * abs(v) + 1e-13, (type(problem).__name__, k, a[k], v)
            else:
                assert (np.isnan(a[k]) and np.isnan(v)) or a[k] == v or (np.isinf(a[k]) and np.isinf(v)), (k, a[k], v)


def test_bench_contract_small():
    """bench.py prints exactly one JSON line on stdout with the contract's keys (small grid, seconds)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--nx", "256", "--ny", "256", "--steps", "2",
                          "--warmup", "3", "--inner", "20", "--cpu-n", "128", "--cpu-steps", "10", "--also-shrink", "16"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "MLUPS" and d["value"] > 0 and d["gpu_launches"] == 2 * 20 and d["n_gpus"] == 1
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1.5
    assert d["e2e"]["h2d_bytes_per_step"] == 256 * 256 * 9 * 8 == d["e2e"]["d2h_bytes_per_step"] and d["e2e"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
    # the strong-scaling configurations reported next to the headline (here shrunk 16x): all three measured
    assert [e["preset"] for e in d["also"]] == ["C5s", "C3", "C4"]
    for e in d["also"]:
        assert "skipped" not in e, e
        assert e["value"] > 0 and e["finite"] and e["scaling"] == "strong" and e["gbs_per_gpu"] > 0


def test_asynchronous_copies_with_two_contexts_in_flight(oracle):
    """lbm_upload_f_async / lbm_download_f_async (page-locked arrays): a stream of independent jobs alternating between two
    contexts -- the bench's e2e pipeline -- returns for every job exactly what the synchronous calls return, and the
    oracle's result; pageable arrays are refused."""
    O = oracle
    qo = O.L.D2Q9()
    nx, ny, nsteps = 96, 50, 23
    cm = O.TRT(0.8, 1.1, (1e-6, 0.0))
    jobs = [random_populations(qo, nx, ny, seed=10 + k) for k in range(5)]
    want = [None] * len(jobs)
    for k, f in enumerate(jobs):
        g = f
        for _ in range(nsteps):
            g, _ = O.step(cm, qo, [], g)
        want[k] = g
    ins = [lbm.pinned_empty((nx, ny, 9)) for _ in jobs]
    outs = [lbm.pinned_empty((nx, ny, 9)) for _ in jobs]
    for a, f in zip(ins, jobs):
        a[...] = to_host_layout(f)
    with _abi.Context(nx, ny, "D2Q9", _abi.TRT, [0.8, 1.1]) as c0, _abi.Context(nx, ny, "D2Q9", _abi.TRT, [0.8, 1.1]) as c1:
        for c in (c0, c1):
            c.set_force_uniform(1e-6, 0.0)
        for k in range(len(jobs)):
            c = (c0, c1)[k % 2]
            c.upload_f_async(ins[k])
            c.step(0, nsteps)
            c.download_f_async(outs[k])
        c0.sync(); c1.sync()
        for k in range(len(jobs)):
            assert np.array_equal(to_oracle_layout(outs[k]), want[k]), k
        with pytest.raises(lbm.LbmError):
            c0.upload_f_async(np.asfortranarray(ins[0].copy()))
        with pytest.raises(lbm.LbmError):
            c0.download_f_async(c0.new_f())


@pytest.mark.parametrize("name,model,bck,dtype,arith", [
    ("D2Q9", "TRT", "poiseuille", _abi.F64, _abi.ARITH_EXACT), ("D2Q37", "TRT", "couette", _abi.F64, _abi.ARITH_EXACT),
    ("D2Q17", "MRT", "none", _abi.F64, _abi.ARITH_FAST), ("D2Q21", "SRT", "cavity", _abi.F32, _abi.ARITH_EXACT),
    ("D2Q13", "TRT", "partial", _abi.F64, _abi.ARITH_EXACT),
])
def test_tma_staged_kernel_equals_the_register_kernel(oracle, name, model, bck, dtype, arith):
    """Option tma = 1: the fused pull step with its loads staged through shared memory by cp.async.bulk.tensor (tma.cuh;
    opt-in, measured slower).  Same device functions after the load, so the result is bit-identical to the register kernel
    and -- in Float64 exact SRT / TRT -- to the oracle; grids whose rows end inside a box (nx not a multiple of 128, nx <
    128) exercise the zero-filled part of the boxes."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    for nx, ny, nsteps in ((300, 21, 19), (40, 37, 9)):
        f0 = random_populations(qo, nx, ny, seed=5)
        force = (2e-6, 1e-6)
        cm, code, taus = _models(O, qo, force)[model]
        ob, hb = _bcs_pair(O, bck, nx, ny)
        outs = []
        for tma in (1, 0):
            with _ctx(name, code, taus, hb, nx, ny, arith, dtype) as c:
                c.set_option("tma", tma)
                c.set_force_uniform(*force)
                c.upload_f(to_host_layout(f0))
                c.step(0, nsteps)
                outs.append(to_oracle_layout(c.download_f()))
        assert np.array_equal(outs[0], outs[1])
        if dtype == _abi.F64 and arith == _abi.ARITH_EXACT and model != "MRT":
            want = f0
            for _ in range(nsteps):
                want, _ = O.step(cm, qo, ob, want)
            assert np.array_equal(outs[0], want)


def test_l2_prefetch_does_not_change_results(oracle):
    """Option prefetch (L2 prefetch distance of the fused pull, automatic for Float64 on grids of 2 Mi nodes and more): a
    hint only -- populations are bit-identical with it forced on (several distances, one beyond the grid) and off."""
    O = oracle
    qo = O.L.D2Q13()
    nx, ny, nsteps = 300, 70, 11
    f0 = random_populations(qo, nx, ny, seed=8)
    cm, code, taus = _models(O, qo, (1e-6, 0.0))["TRT"]
    ob, hb = _bcs_pair(O, "poiseuille", nx, ny)
    outs = []
    for pf in (0, 1, 16, 200):
        with _ctx("D2Q13", code, taus, hb, nx, ny, _abi.ARITH_EXACT, _abi.F64) as c:
            c.set_option("prefetch", pf)
            c.set_force_uniform(1e-6, 0.0)
            c.upload_f(to_host_layout(f0))
            c.step(0, nsteps)
            outs.append(to_oracle_layout(c.download_f()))
    want = f0
    for _ in range(nsteps):
        want, _ = O.step(cm, qo, ob, want)
    for o in outs:
        assert np.array_equal(o, want)


def test_timer_and_options():
    with _abi.Context(64, 64, "D2Q9", _abi.SRT, [0.9]) as c:
        c.upload_f(np.asfortranarray(np.ones((64, 64, 9)) * lbm.D2Q9().weights))
        c.timer_start()
        c.step(0, 10)
        assert c.timer_stop() > 0 and c.last_step_ms() > 0
        assert c.kernel_launches >= 10
        with pytest.raises(lbm.LbmError):
            c.set_option("no_such_option", 1)
        c.set_option("overlap", 0)
