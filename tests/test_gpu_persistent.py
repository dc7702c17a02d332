"""The persistent multi-step kernel (csrc/persist.cuh: one cooperative launch per lbm_step batch, CTAs synchronise only
with the owners of the rows within the stencil's reach) against the oracle and against the launch-per-step path.

Bar as everywhere: Float64 exact SRT/TRT bit-identical to the oracle, MRT / fast 1e-12, Float32 1e-5.
"""
import numpy as np
import pytest

import lbm
from lbm import _abi
from conftest import random_populations, rel_max, to_host_layout, to_oracle_layout

pytestmark = pytest.mark.gpu

LATTICES = ["D2Q4", "D2Q5", "D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"]


def _models(O, qo, force):
    taus = [0.8, 0.9, 1.1, 1.3][:max(qo.N, 2)]
    return {"SRT": (O.SRT(0.8, force), _abi.SRT, [0.8]), "TRT": (O.TRT(0.8, 1.1, force), _abi.TRT, [0.8, 1.1]),
            "MRT": (O.MRT(qo, taus, force), _abi.MRT, taus)}


def _walls(O, kind, nx, ny):
    if kind == "none":
        return [], []
    if kind == "couette":
        ob = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0.002])]
        hb = [lbm.BounceBack(lbm.South(), (1, nx), (1, ny)), lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.01, 0.002])]
    else:  # cavity
        ob = [O.BounceBack("E", (1, nx), (1, ny)), O.BounceBack("S", (1, nx), (1, ny)),
              O.BounceBack("W", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0])]
        hb = [lbm.BounceBack(lbm.East(), (1, nx), (1, ny)), lbm.BounceBack(lbm.South(), (1, nx), (1, ny)),
              lbm.BounceBack(lbm.West(), (1, nx), (1, ny)), lbm.MovingWall(lbm.North(), (1, nx), (1, ny), [0.01, 0])]
    return ob, [b.to_abi() for b in hb]


def _check(got, want, dtype, arith, model):
    if dtype == _abi.F64 and arith == _abi.ARITH_EXACT and model != "MRT":
        assert np.array_equal(got, want), f"not bit-identical: max abs diff {np.abs(got - want).max():.3e}"
    assert rel_max(got, want) < (1e-12 if dtype == _abi.F64 else 1e-5)


@pytest.mark.parametrize("name", LATTICES)
@pytest.mark.parametrize("model", ["SRT", "TRT", "MRT"])
@pytest.mark.parametrize("shape,bck", [((21, 10), "couette"), ((133, 77), "none"), ((3, 5), "none"), ((1, 1), "none"),
                                       ((64, 7), "cavity")])
def test_persistent_steps_match_oracle(oracle, name, model, shape, bck):
    """Every lattice x model on grids from one node (all dependencies wrap onto the same CTA) to a few hundred CTAs' worth;
    odd and even step counts (the ping-pong role of the buffers at the end), resume after a download."""
    O = oracle
    qo = O.L.BY_NAME[name]()
    nx, ny = shape
    f0 = random_populations(qo, nx, ny, seed=21)
    force = (2e-6, 1e-6)
    cm, code, taus = _models(O, qo, force)[model]
    ob, hb = _walls(O, bck, nx, ny)
    want = [f0]
    for _ in range(20):
        want.append(O.step(cm, qo, ob, want[-1])[0])
    for dtype, arith in ((_abi.F64, _abi.ARITH_EXACT), (_abi.F64, _abi.ARITH_FAST), (_abi.F32, _abi.ARITH_FAST)):
        if dtype == _abi.F32 and bck == "none" and shape != (133, 77):
            continue
        with _abi.Context(nx, ny, name, code, taus, hb, dtype=dtype, arith=arith) as c:
            c.set_option("persistent", 1)
            c.set_force_uniform(*force)
            c.upload_f(to_host_layout(f0))
            l0 = c.kernel_launches
            c.step(0, 10)   # collide-only launch + ONE persistent launch of 9 steps (odd)
            launches = c.kernel_launches - l0
            got10 = to_oracle_layout(c.download_f())
            c.step(10, 6)   # resume: pulls from the retained f_collision, 1 + 5 (odd) steps
            c.step(16, 4)   # ... and an even one
            got20 = to_oracle_layout(c.download_f())
        assert launches <= 4, f"{launches} launches for a 10-step batch: the persistent kernel was not used"
        if dtype == _abi.F32:  # random start of amplitude 1e-2: Float32 deviations keep ~1e-7 of that
            assert rel_max(got10, want[10]) < 1e-5 and rel_max(got20, want[20]) < 1e-5
        else:
            _check(got10, want[10], dtype, arith, model)
            _check(got20, want[20], dtype, arith, model)


@pytest.mark.parametrize("name,model,nx,ny,bck", [
    ("D2Q9", "TRT", 1024, 1024, "none"),      # the launch-bound slab size of C3 on 8 GPUs: 4 CTAs per SM, ~7 nodes per thread
    ("D2Q9", "SRT", 1024, 1024, "couette"),
    ("D2Q37", "TRT", 512, 300, "couette"),    # halo 3: dependencies reach several CTAs
    ("D2Q13", "MRT", 777, 333, "none"),       # ranges that do not align with rows
    ("D2Q21", "TRT", 2048, 96, "cavity"),
])
def test_persistent_kernel_at_production_sizes_vs_c_oracle(oracle, name, model, nx, ny, bck):
    O = oracle
    from oracle.c_oracle import COracle
    qo = O.L.BY_NAME[name]()
    pr = O.TGV(qo, 0.8, 1, nx, ny, u_max=2e-3)
    f0 = O.initialize("AnalyticalEquilibrium", qo, pr)
    force = (1e-6, -1e-6)
    cm, code, taus = _models(O, qo, force)[model]
    ob, hb = _walls(O, bck, nx, ny)
    nsteps = 25
    want, _ = COracle(qo, cm, ob).steps(f0, nsteps)
    res = {}
    for persistent in (1, 0):
        for dtype, arith in ((_abi.F64, _abi.ARITH_EXACT), (_abi.F64, _abi.ARITH_FAST), (_abi.F32, _abi.ARITH_FAST)):
            with _abi.Context(nx, ny, name, code, taus, hb, dtype=dtype, arith=arith) as c:
                c.set_option("persistent", persistent)
                c.set_force_uniform(*force)
                c.upload_f(to_host_layout(f0))
                c.step(0, nsteps)
                got = to_oracle_layout(c.download_f())
                res[(persistent, dtype, arith)] = got
            _check(got, want, dtype, arith, model)
    # same arithmetic, different schedule: exact mode must agree bit for bit between the two paths
    assert np.array_equal(res[(1, _abi.F64, _abi.ARITH_EXACT)], res[(0, _abi.F64, _abi.ARITH_EXACT)])


def test_persistent_time_dependent_force_and_diagnostics(oracle):
    """The static decaying shear wave carries a separable time-dependent force table indexed by the step inside the
    launch (decaying_shear_flow.jl:131-147), and simulate() interleaves batches with next! reductions: the df rows of a
    persistent run equal those of the launch-per-step run and of the oracle."""
    O = oracle
    q = lbm.D2Q9()
    nu = 0.8 / (2.0 * q.speed_of_sound_squared)
    rows = {}
    for persistent in (1, 0):
        problem = lbm.DecayingShearFlow(nu, 2, static=True)
        n_steps = round(1.0 / problem.delta_t())
        pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.MeanVelocityStoppingCriteria(0.0, 1e-30, problem))
        m = lbm.LatticeBoltzmannModel(problem, q, collision_model=lbm.SRT, initialization_strategy=lbm.AnalyticalEquilibrium(),
                                      process_method=pm)
        m.ctx.set_option("persistent", persistent)
        lbm.simulate(m, range(0, n_steps + 1))
        rows[persistent] = (pm.df[-1], m.f_stream)
        m.close()
    assert np.array_equal(rows[1][1], rows[0][1])
    for k, v in rows[0][0].items():
        assert rows[1][0][k] == v, k
    qo = O.L.D2Q9()
    po = O.DecayingShearFlow(nu, 2, static=True)
    n_steps = round(1.0 / po.delta_t())
    mo = O.simulate(po, qo, pm=O.TrackHydrodynamicErrors(po, False, n_steps, O.NoStoppingCriteria()), t_end=1.0)
    assert abs(rows[1][0]["error_u"] - mo.pm.df[-1]["error_u"]) <= 1e-9 * abs(mo.pm.df[-1]["error_u"])
