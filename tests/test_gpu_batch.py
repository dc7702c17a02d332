"""GPU parity of the batched small-problem path (lbm_batch_*, csrc/batch.cuh) against the oracle.

The batch kernel keeps every problem's populations in shared memory for a whole run and folds stream! + apply! into a
source-index table; the checks are the same as for the fused HBM kernel: every problem of a batch, with its own
relaxation times / force / initial state, must equal the oracle stepping that problem alone (Float64 exact SRT/TRT
bit-identical, MRT and fast 1e-12, Float32 1e-5), the on-chip stop criteria must fire at the oracle's step, and the
reference's golden table (examples/notebooks/trt_magic_parameter.ipynb:109-176) must come out of ONE batch.
"""
import json
import os

import numpy as np
import pytest

import lbm
from lbm import _abi
from conftest import rel_max, to_host_layout, to_oracle_layout

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "trt_magic_parameter.json")))

# (lattice, model, nx, ny, bcs) -- warp-resident problems (<= 32 nodes, several per warp) and CTA-resident ones
SHAPES = [
    ("D2Q9", "TRT", 3, 5, "poiseuille"),    # the reference's sweep shape: 2 problems per warp
    ("D2Q9", "SRT", 1, 10, "couette"),      # bench_simulation.jl's CouetteFlow(scale = 2): 3 per warp
    ("D2Q13", "SRT", 4, 7, "couette"),      # halo 2, moving wall, 1 per warp
    ("D2Q37", "TRT", 2, 9, "couette"),      # halo 3 on a grid narrower than the stencil (multiple wraps)
    ("D2Q9", "MRT", 6, 6, "none"),          # 36 nodes: one CTA per problem
    ("D2Q17", "MRT", 5, 3, "none"),
    ("D2Q9", "TRT", 40, 9, "cavity"),       # 360 nodes: threads stride over nodes, four walls
    ("D2Q21", "TRT", 7, 11, "partial"),
    ("D2Q4", "SRT", 2, 2, "none"),
    ("D2Q5", "TRT", 1, 1, "none"),          # 32 problems per warp
]


def _bcs(O, kind, nx, ny):
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
    else:  # partial walls: wrapped populations survive elsewhere
        ob = [O.BounceBack("N", (2, nx - 1), (1, ny)), O.BounceBack("W", (1, nx), (2, ny - 2)), O.BounceBack("E", (1, nx), (3, ny))]
        hb = [lbm.BounceBack(lbm.North(), (2, nx - 1), (1, ny)), lbm.BounceBack(lbm.West(), (1, nx), (2, ny - 2)),
              lbm.BounceBack(lbm.East(), (1, nx), (3, ny))]
    return ob, [b.to_abi() for b in hb]


def _oracle_model(O, qo, model, taus, force):
    if model == "SRT":
        return O.SRT(taus[0], force)
    if model == "TRT":
        return O.TRT(taus[0], taus[1], force)
    return O.MRT(qo, list(taus), force)


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{s[0]}-{s[1]}-{s[2]}x{s[3]}-{s[4]}" for s in SHAPES])
@pytest.mark.parametrize("mode", ["f64-exact", "f64-fast", "f32-fast"])
def test_every_problem_of_a_batch_matches_the_oracle(oracle, shape, mode):
    O = oracle
    name, model, nx, ny, bck = shape
    qo = O.L.BY_NAME[name]()
    B, nsteps = 37, 23
    rng = np.random.default_rng(100 + SHAPES.index(shape))
    ntau = {"SRT": 1, "TRT": 2, "MRT": max(qo.N, 2)}[model]
    taus = rng.uniform(0.55, 2.0, (B, ntau))
    forces = rng.uniform(-1e-5, 1e-5, (B, 2))
    f0 = np.stack([np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)]) for _ in range(B)])
    ob, hb = _bcs(O, bck, nx, ny)
    dtype = _abi.F32 if mode.startswith("f32") else _abi.F64
    arith = _abi.ARITH_EXACT if mode.endswith("exact") else _abi.ARITH_FAST
    code = {"SRT": _abi.SRT, "TRT": _abi.TRT, "MRT": _abi.MRT}[model]
    with _abi.Batch(B, nx, ny, name, code, taus[0], hb, dtype=dtype, arith=arith) as b:
        b.set_tau(taus)
        b.set_force_uniform(forces)
        b.upload_f(np.stack([to_host_layout(f0[k]) for k in range(B)], axis=3))
        b.run(10)
        b.run(nsteps - 10)   # resume from f_stream in HBM
        got = b.download_f()
        steps, stopped = b.status()
        launches = b.kernel_launches
    assert launches == 2 and (steps == nsteps).all() and not stopped.any()
    for k in range(B):
        cm = _oracle_model(O, qo, model, taus[k], tuple(forces[k]))
        want = f0[k]
        for _ in range(nsteps):
            want, _ = O.step(cm, qo, ob, want)
        g = to_oracle_layout(got[..., k])
        if mode == "f64-exact" and model != "MRT":
            assert np.array_equal(g, want), (k, np.abs(g - want).max())
        assert rel_max(g, want) < (1e-5 if dtype == _abi.F32 else 1e-12), k


def _oracle_solve(O, tau_s, tau_a, stop="velocity"):
    q = O.L.D2Q9()
    problem = O.PoiseuilleFlow((tau_s - 0.5) / q.css, 1)
    n_steps = round(100.0 / problem.delta_t())
    sc = (O.VelocityConvergenceStoppingCriteria(1e-7, problem) if stop == "velocity"
          else O.MeanVelocityStoppingCriteria(1e-9))
    pm = O.TrackHydrodynamicErrors(problem, False, n_steps, sc)
    cm = O.TRT(tau_s, tau_a, O._problem_force(problem))
    m = O.simulate(problem, q, pm=pm, should_process=False, strategy="ZeroVelocityInitialCondition", t_end=100.0, collision=cm)
    return m.pm.df[-1]


def _sweep(pairs, stop="velocity", **kw):
    q = lbm.D2Q9()
    ts = sorted({p[0] for p in pairs})
    problems = [lbm.PoiseuilleFlow((t - 0.5) / q.speed_of_sound_squared, 1) for t in ts]
    pidx = [ts.index(p[0]) for p in pairs]
    sc = (lbm.VelocityConvergenceStoppingCriteria(1e-7, problems[0]) if stop == "velocity"
          else lbm.MeanVelocityStoppingCriteria(0.0, 1e-9, problems[0]))
    return lbm.simulate_many(problems, q, np.array(pairs), lbm.TRT, problem_index=pidx, t_end=100.0, stop_criteria=sc,
                             initialization_strategy=lbm.ZeroVelocityInitialCondition(), **kw)


def _close_to_printed(value, printed, digits=6):
    if np.isinf(printed):
        return np.isinf(value)
    ulp = 10.0 ** (np.floor(np.log10(abs(printed))) - (digits - 1))
    return abs(value - printed) <= 0.51 * ulp


def test_golden_table_from_one_batch():
    """All 41 rows the reference prints of its 902 500-row sweep (D2Q9 TRT + force Poiseuille, 3 x 5, velocity
    convergence 1e-7 checked every 100 steps, <= 5000 steps) from ONE batch launch."""
    rows = G["rows"]
    res = _sweep([(r["tau_s"], r["tau_a"]) for r in rows])
    assert len(res) == len(rows)
    for k, r in enumerate(rows):
        assert _close_to_printed(res.error_u[k], r["error_u"]), (r, res.row(k))
        assert _close_to_printed(res.error_p[k], r["error_p"]), (r, res.row(k))
        if "error_sxx" in r:
            assert _close_to_printed(res.error_sxx[k], r["error_sxx"]), (r, res.row(k))
            assert _close_to_printed(res.error_sxy[k], r["error_sxy"]), (r, res.row(k))


@pytest.mark.parametrize("stop", ["velocity", "mean"])
def test_on_chip_stop_criteria_fire_at_the_oracles_step(oracle, stop):
    """Per-problem stop criteria (velocity convergence / mean velocity) against the oracle's simulate: same stop step,
    same recorded row -- early stoppers (large tau), late ones, and solves that run to n_steps, mixed in one batch so that
    warp-mates retire at different times."""
    pairs = [(10.0, 10.0), (0.51, 0.51), (5.0, 0.7), (0.8, 1.1), (10.0, 0.51), (2.5, 2.5), (0.6, 9.0), (1.0, 1.0), (7.3, 3.1)]
    res = _sweep(pairs, stop)
    seen = set()
    for k, (ts, ta) in enumerate(pairs):
        e = _oracle_solve(oracle, ts, ta, stop)
        assert res.timestep[k] == e["timestep"], (ts, ta, res.timestep[k], e["timestep"])
        assert bool(res.stopped[k]) == (e["timestep"] < 5000) or e["timestep"] == 5000
        for col in ("error_u", "error_p", "error_sxy", "mass", "momentum", "energy"):
            a, b = getattr(res, col)[k], e[col]
            # error_p is the norm of p - 1 ~ 1e-9: round-off of the summation order shows at 1e-7 relative
            assert abs(a - b) <= (1e-5 if col == "error_p" else 1e-9) * abs(b) + 1e-300, (ts, ta, col, a, b)
        seen.add(int(res.timestep[k]))
    assert len(seen) > 2  # the batch really mixed stop times


def test_batch_equals_single_context_path():
    """A batch of one problem == the lbm_ctx path (same kernels' arithmetic, different residency), incl. Float32."""
    q = lbm.D2Q9()
    pr = lbm.PoiseuilleFlow(0.3, 1)
    cm = lbm.TRT(1.4, 0.8, lbm.LatticeForce(pr))
    f0 = lbm.initialize(lbm.ZeroVelocityInitialCondition(), q, pr)
    for dtype in ("f64", "f32"):
        m = lbm.LatticeBoltzmannModel(pr, q, collision_model=cm, initialization_strategy=lbm.ZeroVelocityInitialCondition(), dtype=dtype)
        m.state.step(0, 77, pr.delta_t())
        want = m.f_stream
        m.close()
        with _abi.Batch(3, pr.NX, pr.NY, "D2Q9", _abi.TRT, cm.taus(), [b.to_abi() for b in pr.boundary_conditions()],
                        dtype=_abi.F64 if dtype == "f64" else _abi.F32) as b:
            b.set_force_uniform(np.tile(lbm.LatticeForce(pr).uniform(), (3, 1)))
            b.broadcast_f(f0)
            b.run(77)
            got = b.download_f()
        for k in range(3):
            assert np.array_equal(got[..., k], want), dtype


def test_batch_error_paths():
    with pytest.raises(lbm.LbmError):  # does not fit on chip
        _abi.Batch(2, 256, 256, "D2Q9", _abi.SRT, [1.0])
    with pytest.raises(lbm.LbmError):
        _abi.Batch(0, 3, 5, "D2Q9", _abi.SRT, [1.0])
    with _abi.Batch(4, 3, 5, "D2Q9", _abi.TRT, [0.8, 1.1]) as b:
        with pytest.raises(ValueError):
            b.set_tau(np.ones((4, 3)))
        with pytest.raises(lbm.LbmError):
            b.download_f(3, 2)
        with pytest.raises(lbm.LbmError):
            b.set_tau(np.zeros((4, 2)))
