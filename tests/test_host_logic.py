"""Host-side logic that runs without a GPU: problems / initial conditions of the host mirror
against the oracle's independent restatement, collision-model factories, the batching of
`simulate`, and the y-slab decomposition (world_size 2, gloo)."""
import os

import numpy as np
import pytest

import lbm
from lbm import _abi
import oracle.lbm_oracle as O
from conftest import to_oracle_layout

PAIRS = [
    ("TGV", lambda q: lbm.TGV(q, 0.8, 1, 8, 12), lambda q: O.TGV(q, 0.8, 1, 8, 12)),
    ("TaylorGreenVortex", lambda q: lbm.TaylorGreenVortex(1 / 6, 1, 8, 8), lambda q: O.TaylorGreenVortex(1 / 6, 1, 8, 8)),
    ("TaylorGreenVortexDecay", lambda q: lbm.TaylorGreenVortex(1 / 6, 1, 8, 8, static=False),
     lambda q: O.TaylorGreenVortex(1 / 6, 1, 8, 8, static=False)),
    ("DecayingShearFlow", lambda q: lbm.DecayingShearFlow(1 / 6, 2), lambda q: O.DecayingShearFlow(1 / 6, 2)),
    ("DecayingShearFlowDecay", lambda q: lbm.DecayingShearFlow(1 / 6, 2, static=False, k_y=1.0),
     lambda q: O.DecayingShearFlow(1 / 6, 2, static=False, k_y=1.0)),
    ("PoiseuilleFlow", lambda q: lbm.PoiseuilleFlow(1 / 6, 2), lambda q: O.PoiseuilleFlow(1 / 6, 2)),
    ("CouetteFlow", lambda q: lbm.CouetteFlow(1 / 6, 2), lambda q: O.CouetteFlow(1 / 6, 2)),
    ("LidDrivenCavityFlow", lambda q: lbm.LidDrivenCavityFlow(1 / 6, 1), lambda q: O.LidDrivenCavityFlow(1 / 6, 1)),
    ("LinearizedThermalDiffusion", lambda q: lbm.LinearizedThermalDiffusion(1 / 6, 1 / 6, 2),
     lambda q: O.LinearizedThermalDiffusion(1 / 6, 1 / 6, 2)),
    ("LinearizedTransverseShearWave", lambda q: lbm.LinearizedTransverseShearWave(1 / 6, 1 / 6, 2),
     lambda q: O.LinearizedTransverseShearWave(1 / 6, 1 / 6, 2)),
]


@pytest.mark.parametrize("name,mk_h,mk_o", PAIRS, ids=[p[0] for p in PAIRS])
def test_problem_fields_match_oracle(name, mk_h, mk_o):
    q, qo = lbm.D2Q9(), O.L.D2Q9()
    ph, po = mk_h(q), mk_o(qo)
    assert (ph.NX, ph.NY, ph.u_max, ph.nu) == (po.NX, po.NY, po.u_max, po.nu)
    assert ph.delta_x() == po.delta_x() and ph.delta_t() == po.delta_t() and ph.viscosity() == po.viscosity()
    X, Y = ph.grid()
    Xo, Yo = po.grid()
    assert np.array_equal(X.T, Xo) and np.array_equal(Y.T, Yo)
    for t in (0.0, 0.37):
        assert np.allclose(ph.density(q, X, Y, t).T, po.density(qo, Xo, Yo, t), rtol=1e-15)
        assert np.allclose(ph.pressure(q, X, Y, t).T, po.pressure(qo, Xo, Yo, t), rtol=1e-15)
        for a, b in zip(ph.velocity(X, Y, t), po.velocity(Xo, Yo, t)):
            assert np.allclose(np.asarray(a).T, b, rtol=1e-15, atol=1e-18)
        (s11, s12), (s21, s22) = ph.deviatoric_tensor(q, X, Y, t)
        so = po.deviatoric(qo, Xo, Yo, t)
        for a, b in ((s11, so[0, 0]), (s12, so[0, 1]), (s21, so[1, 0]), (s22, so[1, 1])):
            assert np.allclose(np.asarray(a).T, b, rtol=1e-15, atol=1e-18)
    assert ph.has_external_force() == po.has_external_force()
    assert len(ph.boundary_conditions()) == len(po.boundary_conditions())


@pytest.mark.parametrize("lattice", ["D2Q4", "D2Q9", "D2Q17", "D2Q37"])
@pytest.mark.parametrize("strategy", ["ZeroVelocityInitialCondition", "AnalyticalEquilibrium", "ConstantDensity",
                                      "AnalyticalVelocityAndStress", "AnalyticalEquilibriumAndOffEquilibrium"])
def test_initialize_matches_oracle(lattice, strategy):
    q, qo = getattr(lbm.Quadratures, lattice), O.L.BY_NAME[lattice]()
    for mk_h, mk_o in ((PAIRS[0][1], PAIRS[0][2]), (PAIRS[2][1], PAIRS[2][2]), (PAIRS[4][1], PAIRS[4][2])):
        ph, po = mk_h(q), mk_o(qo)
        fh = lbm.initialize(getattr(lbm, strategy)(), q, ph)
        fo = O.initialize(strategy, qo, po)
        assert fh.shape == (ph.NX, ph.NY, q.Q) and fh.flags.f_contiguous
        assert np.allclose(to_oracle_layout(fh), fo, rtol=1e-13, atol=1e-17)
        # slab initialisation == rows of the global one
        part = lbm.initialize(getattr(lbm, strategy)(), q, ph, rows=(3, 4))
        assert np.array_equal(part, fh[:, 3:7, :])


@pytest.mark.parametrize("strategy", ["ZeroVelocityInitialCondition", "AnalyticalEquilibrium", "ConstantDensity",
                                      "AnalyticalVelocityAndStress", "AnalyticalEquilibriumAndOffEquilibrium"])
@pytest.mark.parametrize("lattice", ["D2Q9", "D2Q37"])
def test_analytic_device_init_spec_matches_oracle(lattice, strategy):
    """The separable tables + coefficients the host hands to lbm_init_analytic (O(NX + NY) evaluations of the problem's
    pointwise functions) describe exactly initialize(strategy, q, problem): evaluated the way the kernel does (here by
    the oracle-backed stand-in context) they reproduce the oracle's initialize for every problem and strategy."""
    from _oracle_context import OracleContext
    from lbm.initial_conditions import analytic_init_spec
    q, qo = getattr(lbm.Quadratures, lattice), O.L.BY_NAME[lattice]()
    for name, mk_h, mk_o in PAIRS:
        ph, po = mk_h(q), mk_o(qo)
        spec = analytic_init_spec(getattr(lbm, strategy)(), q, ph)
        assert spec is not None, name
        ctx = OracleContext(ph.NX, ph.NY, lattice, _abi.SRT, [1.0])
        ctx.init_analytic(spec[0], **spec[1])
        want = O.initialize(strategy, qo, po)
        got = to_oracle_layout(ctx.download_f())
        assert np.abs(got - want).max() <= 2e-15 * np.abs(want).max(), (name, np.abs(got - want).max())
        # a slab of rows gets the same tables restricted to its rows
        part = analytic_init_spec(getattr(lbm, strategy)(), q, ph, 2, 3)
        c2 = OracleContext(ph.NX, 3, lattice, _abi.SRT, [1.0])
        c2.init_analytic(part[0], **part[1])
        assert np.abs(to_oracle_layout(c2.download_f()) - want[:, 2:5]).max() <= 2e-15 * np.abs(want).max(), name


def test_decompose_function_refuses_rank_three_and_costs_lines_only():
    from lbm.separable import decompose_function
    xs, ys = np.linspace(0, 1, 300), np.linspace(0, 2, 200)
    calls = []

    def fn(X, Y):
        calls.append(np.broadcast(X, Y).size)
        return np.sin(3 * X) * np.cos(Y) + X * Y ** 2
    c0, terms = decompose_function(fn, xs, ys)
    R = sum(a * np.outer(X, Y) for a, X, Y in terms)
    assert np.abs(R - fn(xs[:, None], ys[None, :])).max() < 1e-13
    assert sum(calls[:-1]) < 48 * 48 + 6 * (300 + 200) + 4 * 512  # never the whole 300 x 200 grid
    assert decompose_function(lambda X, Y: np.sin(X * Y) + np.exp(X + Y ** 2) + 1 / (1 + X + Y), xs, ys) is None
    assert decompose_function(lambda X, Y: 0.0 * X, xs, ys) == (0.0, [])


def test_forces_match_oracle():
    q, qo = lbm.D2Q9(), O.L.D2Q9()
    ph, po = lbm.PoiseuilleFlow(1 / 6, 2), O.PoiseuilleFlow(1 / 6, 2)
    assert np.allclose(lbm.LatticeForce(ph).uniform(), O._problem_force(po), rtol=1e-15)
    assert np.allclose(lbm.lattice_force(ph, 2, 3, 0.0), O._problem_force(po), rtol=1e-15)
    ph, po = lbm.TaylorGreenVortex(1 / 6, 1, 8, 8), O.TaylorGreenVortex(1 / 6, 1, 8, 8)
    Fx, Fy = lbm.LatticeForce(ph).field(0.0, 0, 8)
    Fo = O._problem_force(po)
    assert np.allclose(Fx.T, Fo[0], rtol=1e-15) and np.allclose(Fy.T, Fo[1], rtol=1e-15)
    assert np.allclose(lbm.lattice_force(ph, 2, 5, 0.0), [Fo[0][4, 1], Fo[1][4, 1]], rtol=1e-15)
    ph, po = lbm.DecayingShearFlow(1 / 6, 2), O.DecayingShearFlow(1 / 6, 2)
    fx, fy = lbm.LatticeForce(ph).separable(3, 4, 0, ph.NY)
    for k in range(4):
        Fo = O._problem_force(po)((3 + k) * po.delta_t())
        assert np.allclose(fx[k][:, None] + 0 * Fo[0], Fo[0], rtol=1e-15, atol=1e-20)
        assert np.allclose(fy[k][None, :] + 0 * Fo[1], Fo[1], rtol=1e-15, atol=1e-20)
    assert lbm.LatticeForce(ph).kind() == "separable"


def test_collision_model_factories():
    q = lbm.D2Q9()
    pr = lbm.PoiseuilleFlow(1 / 6, 2)
    srt = lbm.CollisionModel(lbm.SRT, q, pr)
    assert srt.tau == 3.0 * (1 / 6) + 0.5 and isinstance(srt.force, lbm.LatticeForce)
    trt = lbm.CollisionModel(lbm.TRT, q, pr)
    assert trt.tau_symmetric == srt.tau and trt.tau_asymmetric == 0.5 + 0.25 / (srt.tau - 0.5)
    trt = lbm.CollisionModel(lbm.TRT_Lambda(3 / 16), q, pr)
    assert trt.tau_asymmetric == 0.5 + (3 / 16) / (srt.tau - 0.5)
    mrt = lbm.CollisionModel(lbm.MRT, q, pr)
    assert mrt.taus() == [srt.tau] * 5 and mrt.force is not None  # fill(tau, order(q)), mrt.jl:38
    inst = lbm.SRT(0.7)
    assert lbm.CollisionModel(inst, q, pr) is inst and inst.force is None  # instance => no force
    t2 = lbm.TRT(0.9, 0.6)  # 2-arg convenience ctor is TRT(tau_a, tau_s): trt.jl:6
    assert (t2.tau_symmetric, t2.tau_asymmetric) == (0.6, 0.9)
    t3 = lbm.TRT(0.6, 0.9, None)
    assert (t3.tau_symmetric, t3.tau_asymmetric) == (0.6, 0.9)
    assert lbm.MRT(lbm.D2Q37(), 0.8).taus() == [0.8] * 4
    # N = round(Int, order(q) / 2) rounds half to even: order 7 -> 4 relaxation times (mrt.jl:20,25)
    assert lbm.MRT(lbm.D2Q17(), 0.8, 0.9).taus() == [0.8, 0.9, 0.8, 0.9]
    assert lbm.MRT(lbm.D2Q4(), 0.8).taus() == [0.8, 0.8]
    assert lbm.MRT(q, 0.8, lambda x, y, t: [1, 1]).force is None  # scalar form drops the force (mrt.jl:19-22)
    assert lbm.MRT(q, [0.8, 0.9], "F").force == "F"
    default = lbm.CollisionModel(lbm.collision_models.CollisionModelBase, q, pr)
    assert isinstance(default, lbm.SRT)


class FakeCtx:
    def __init__(self):
        self.log = []

    def set_force_none(self):
        self.log.append(("force_none",))

    def step(self, t0, n, dt):
        self.log.append(("step", t0, n))


class FakeModel(lbm.LatticeBoltzmannModel):
    def __init__(self, pm):
        self.ctx = FakeCtx()
        self.quadrature = lbm.D2Q9()
        self.collision_model = lbm.SRT(1.0)
        self.processing_method = pm
        self.state = lbm.DeviceState.__new__(lbm.DeviceState)
        self.state.ctx, self.state.cm, self.state.comm = self.ctx, self.collision_model, None
        self.state._force_window, self.state._static_force_set = None, False


class RecordingPM(lbm.processing_methods.ProcessingMethodBase):
    def __init__(self, every, stop_at=None):
        self.every, self.stop_at, self.calls = every, stop_at, []
        self.problem = lbm.TGV(lbm.D2Q9(), 0.8, 1)

    def noop(self, t):
        return t % self.every != 0

    def next_(self, q, state, t):
        self.calls.append(t)
        return t == self.stop_at


def test_simulate_batches_steps_between_host_visible_points():
    # reference loop: for t in time: step(t); if next!(t+1) return; end; next!(last+1)
    m = FakeModel(RecordingPM(100))
    lbm.simulate(m, range(0, 251))
    assert [e for e in m.ctx.log if e[0] == "step"] == [("step", 0, 100), ("step", 100, 100), ("step", 200, 51)]
    assert m.processing_method.calls == [100, 200, 251]
    m = FakeModel(RecordingPM(100, stop_at=200))
    lbm.simulate(m, range(0, 1000))
    assert [e for e in m.ctx.log if e[0] == "step"] == [("step", 0, 100), ("step", 100, 100)]
    assert m.processing_method.calls == [100, 200]  # early return: no trailing next!
    m = FakeModel(RecordingPM(1))
    lbm.simulate(m, range(1, 4))  # simulate(model, 1:3): next!(2), next!(3), next!(4), trailing next!(4)
    assert [e for e in m.ctx.log if e[0] == "step"] == [("step", 1, 1), ("step", 2, 1), ("step", 3, 1)]
    assert m.processing_method.calls == [2, 3, 4, 4]
    m = FakeModel(None)
    lbm.simulate(m, range(0, 11))
    assert [e for e in m.ctx.log if e[0] == "step"] == [("step", 0, 11)]


def test_processing_method_noop_matches_next_semantics():
    pr = lbm.PoiseuilleFlow(1 / 6, 1)
    pm = lbm.TrackHydrodynamicErrors(pr, False, 5000, lbm.VelocityConvergenceStoppingCriteria(1e-7, pr))
    assert [t for t in range(1, 301) if not pm.noop(t)] == [100, 200, 300]
    assert not pm.noop(5000)
    pm = lbm.TrackHydrodynamicErrors(lbm.TGV(lbm.D2Q9(), 0.8, 1), False, 250)
    assert [t for t in range(1, 301) if not pm.noop(t)] == [250]  # NoStoppingCriteria: only t == n_steps
    pm = lbm.TrackHydrodynamicErrors(pr, True, 10)
    assert not any(pm.noop(t) for t in range(1, 5))
    snap = lbm.TakeSnapshots(pr, [3, 7])
    assert [t for t in range(1, 10) if not snap.noop(t)] == [3, 7]
    assert isinstance(lbm.ProcessingMethod(pr, False, 10), lbm.CompareWithAnalyticalSolution)
    assert isinstance(lbm.ProcessingMethod(lbm.TGV(lbm.D2Q9(), 0.8, 1), False, 10), lbm.TrackHydrodynamicErrors)
    assert isinstance(lbm.StopCriteria(pr), lbm.MeanVelocityStoppingCriteria) and lbm.StopCriteria(pr).tolerance == 1e-12
    assert isinstance(lbm.StopCriteria(lbm.TGV(lbm.D2Q9(), 0.8, 1)), lbm.NoStoppingCriteria)


def test_density_convergence_literal_and_whole_field():
    """DensityConvergence (density_convergence.jl:6-17): the literal restatement watches node (NX, NY) only (first call
    compares with the initial 0), the whole-field variant uses the norm over all nodes; both stop above 100."""

    class St:
        def __init__(self, seq):
            self.seq = list(seq)

        def reduce(self, kind):
            assert kind == _abi.REDUCE_DENSITY_CHANGE
            return np.array(self.seq.pop(0) + [0.0, 0.0])

    sc = lbm.DensityConvergence(1e-3, None)
    st = St([[4.0, 1.0], [4.0, 1.5], [4.0, 1.5004], [0.0, 300.0]])
    assert [lbm.processing_methods.should_stop_(sc, None, st) for _ in range(4)] == [False, False, True, True]
    sc = lbm.DensityConvergence(1e-3, None, whole_field=True)
    st = St([[4.0, 1.0], [1e-8, 1.0], [1e6, 1.0]])
    assert [lbm.processing_methods.should_stop_(sc, None, st) for _ in range(3)] == [False, True, True]
    pm = lbm.ProcessIterativeInitialization(1e-3, lbm.TGV(lbm.D2Q9(), 0.8, 1))
    assert pm.n_steps == 100 and not pm.noop(1) and isinstance(pm.stop_criteria, lbm.DensityConvergence)
    s = lbm.IterativeInitialization()
    assert (s.tau, s.eps, s.max_steps, s.check_every) == (1.0, 1e-7, 10000, 1)
    cm = lbm.IterativeInitializationCollisionModel(lbm.D2Q9(), 0.9, lbm.TGV(lbm.D2Q9(), 0.8, 1))
    ux, uy = cm.velocity_field(4, 3)
    assert cm.taus() == [0.9] and cm.force is None and ux.shape == uy.shape == (16, 3)


def test_slab_rows_partition():
    for ny in (1, 7, 64, 4097):
        for world in (1, 2, 3, 8):
            rows = [lbm.slab_rows(ny, r, world) for r in range(world)]
            assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
            for (a, n), (b, _) in zip(rows, rows[1:]):
                assert a + n == b
            assert max(n for _, n in rows) - min(n for _, n in rows) <= 1
    assert [lbm.halo_rows_per_direction(q) for q in lbm.Quadratures] == [1, 1, 3, 5, 10, 18, 26]


# ---- world_size 2 over gloo: the decomposition the CUDA library implements ---------------------
def _slab_worker(rank, world, port, lattice, ny, nx, nsteps, with_walls, out):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = lbm.SlabComm()
    qo = O.L.BY_NAME[lattice]()
    q = getattr(lbm.Quadratures, lattice)
    H = qo.h
    y0, nyl = lbm.slab_rows(ny, rank, world)
    rng = np.random.default_rng(99)
    f_global = np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)])
    f = f_global[:, y0:y0 + nyl].copy()
    cm = O.TRT(0.8, 1.1, (1e-6, 2e-6))
    bcs = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0.0])] if with_walls else []
    up, down = (rank + 1) % world, (rank - 1) % world
    for _ in range(nsteps):
        fc = O.collide(cm, qo, f)
        ext = np.zeros((qo.Q, nyl + 2 * H, nx))
        ext[:, H:H + nyl] = fc
        # the same messages liblbm_b200 posts: rows of populations moving up go to `up`, ...
        send_up = torch.from_numpy(np.ascontiguousarray(fc[:, nyl - H:]))
        send_dn = torch.from_numpy(np.ascontiguousarray(fc[:, :H]))
        recv_dn, recv_up = torch.empty_like(send_up), torch.empty_like(send_dn)
        reqs = [dist.isend(send_up, up), dist.isend(send_dn, down), dist.irecv(recv_dn, down), dist.irecv(recv_up, up)]
        if world == 2:  # both neighbours are the same rank: receives match sends in posting order
            pass
        for r in reqs:
            r.wait()
        ext[:, :H] = recv_dn.numpy()
        ext[:, H + nyl:] = recv_up.numpy()
        fs = np.empty_like(fc)
        for i in range(qo.Q):
            cy, cx = int(qo.cy[i]), int(qo.cx[i])
            fs[i] = np.roll(ext[i, H - cy:H - cy + nyl], cx, axis=1)
        if bcs:  # BCs are node-local: apply them on the slab embedded at its global rows
            full_new = np.zeros((qo.Q, ny, nx))
            full_old = np.zeros((qo.Q, ny, nx))
            full_new[:, y0:y0 + nyl], full_old[:, y0:y0 + nyl] = fs, fc
            O.apply_bcs(bcs, qo, full_new, full_old)
            fs = full_new[:, y0:y0 + nyl]
        f = fs
    # distributed diagnostics: per-rank partial sums added by the host (lbm_reduce contract)
    rho = O.density(qo, [f[i] for i in range(qo.Q)])
    total = comm.allreduce_sum([rho.sum(), float(rho.size)])
    out.put((rank, y0, f, total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lattice,with_walls", [("D2Q9", False), ("D2Q9", True), ("D2Q37", False), ("D2Q13", True)])
def test_slab_decomposition_world2_gloo(lattice, with_walls):
    import torch.multiprocessing as mp
    ny, nx, nsteps, world = 13, 6, 4, 2
    qo = O.L.BY_NAME[lattice]()
    rng = np.random.default_rng(99)
    f = np.stack([qo.w[i] * (1 + 0.01 * rng.uniform(-1, 1, (ny, nx))) for i in range(qo.Q)])
    cm = O.TRT(0.8, 1.1, (1e-6, 2e-6))
    bcs = [O.BounceBack("S", (1, nx), (1, ny)), O.MovingWall("N", (1, nx), (1, ny), [0.01, 0.0])] if with_walls else []
    for _ in range(nsteps):
        f, _ = O.step(cm, qo, bcs, f)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, lattice, ny, nx, nsteps, with_walls, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, y0, slab, total in results:
        assert np.array_equal(slab, f[:, y0:y0 + slab.shape[1]]), f"rank {rank} slab differs from the single-domain run"
        assert total[1] == ny * nx and np.isclose(total[0], O.density(qo, [f[i] for i in range(qo.Q)]).sum(), rtol=1e-14)


def test_separable_expected_fields_reproduce_the_analytic_fields():
    """problem.expected_separable (input of lbm_reduce_errors) == density/velocity/pressure/deviatoric_tensor."""
    q = lbm.D2Q9()
    probs = [lbm.TGV(q, 0.8, 1, 8, 12), lbm.TaylorGreenVortex(1 / 6, 1, 8, 8), lbm.TaylorGreenVortex(1 / 6, 1, 8, 8, static=False),
             lbm.DecayingShearFlow(1 / 6, 2), lbm.DecayingShearFlow(1 / 6, 2, static=False, k_y=1.0),
             lbm.PoiseuilleFlow(1 / 6, 2), lbm.CouetteFlow(1 / 6, 2), lbm.LinearizedThermalDiffusion(1 / 6, 1 / 6, 2),
             lbm.LinearizedTransverseShearWave(1 / 6, 1 / 6, 2)]
    for pr in probs:
        for t in (0.0, 0.37):
            for y0, ny in ((0, None), (1, 2)):
                sep = pr.expected_separable(q, t, y0, ny)
                X, Y = pr.grid(y0, ny)
                (sxx, sxy), (syx, syy) = pr.deviatoric_tensor(q, X, Y, t)
                ux, uy = pr.velocity(X, Y, t)
                ref = [pr.density(q, X, Y, t), ux, uy, pr.pressure(q, X, Y, t), sxx, sxy, syx, syy]
                for f, (c0, terms) in enumerate(sep):
                    E = c0 + 0 * X
                    for a, Xt, Yt in terms:
                        xv = np.ones(X.shape[0]) if Xt is None else Xt
                        yv = np.ones(X.shape[1]) if Yt is None else Yt
                        E = E + a * xv[:, None] * yv[None, :]
                    assert np.abs(E - ref[f]).max() <= 1e-13 * max(np.abs(ref[f]).max(), 1e-300), (type(pr).__name__, t, f)
    assert lbm.LidDrivenCavityFlow(1 / 6, 1).expected_separable(q, 0.0) is None


def test_bench_reference_arm_contract():
    """`bench.py --impl reference`: the CPU restatement timed alone -- one JSON line with impl/cpu_baseline/e2e keys; under
    torchrun only rank 0 works and prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--cpu-n", "128", "--cpu-steps", "20"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "128x128" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_api_surface_of_the_reference_resolves():
    """Every name the reference exports or its tests / benchmarks / notebooks use (Julia's `f!` is
    spelled `f_`, `range(problem)` is `range_`) resolves in the host mirror -- except the plotting ones (visualize,
    ShowVelocityError: out of scope) and the names the reference exports without defining them (analayze_convergence,
    GenericFluidFlowProblem, Lattice, momentum, total_energy, kinetic_energy, internal_energy, hermite_equilibrium,
    hermite_first_nonequilibrium)."""
    names = """CollisionModel SRT TRT MRT Quadrature D2Q4 D2Q5 D2Q9 D2Q13 D2Q17 D2Q21 D2Q37 opposite DecayingShearFlow
    LidDrivenCavityFlow TaylorGreenVortex CouetteFlow LinearizedThermalDiffusion LinearizedTransverseShearWave PoiseuilleFlow
    process_ apply_boundary_conditions_ density velocity pressure temperature decay force FluidFlowProblem viscosity delta_t
    initialize AnalyticalEquilibriumAndOffEquilibrium AnalyticalEquilibrium AnalyticalVelocity IterativeInitialization
    dimension equilibrium equilibrium_ hermite stream_ stream collide_ simulate LatticeBoltzmannModel ProcessingMethod
    TrackHydrodynamicErrors CompareWithAnalyticalSolution TakeSnapshots StopCriteria NoStoppingCriteria
    MeanVelocityStoppingCriteria VelocityConvergenceStoppingCriteria DensityConvergence ProcessIterativeInitialization
    IterativeInitializationCollisionModel TGV decay_time ZeroVelocityInitialCondition ConstantDensity
    AnalyticalVelocityAndStress IterativeInitializationMeiEtAl InitializationStrategy TRT_Λ TRT_Lambda lattice_force
    lattice_viscosity lattice_velocity lattice_density lattice_pressure lattice_temperature has_external_force Quadratures
    order velocity_ momentum_flux deviatoric_tensor hermite_based_equilibrium equilibrium_coefficient dimensionless_velocity
    dimensionless_density dimensionless_pressure dimensionless_temperature dimensionless_stress dimensionless_force
    dimensionless_viscosity apply_ BounceBack MovingWall North East South West next_model_ range_ delta_x boundary_conditions
    collide_model_ stream_model_""".split()
    missing = [n for n in names if not hasattr(lbm, n)]
    assert not missing, missing
    # the free functions follow the reference's argument order and agree with the oracle's problem methods
    q, qo = lbm.D2Q9(), O.L.D2Q9()
    pr, po = lbm.TaylorGreenVortex(1 / 6, 1), O.TaylorGreenVortex(1 / 6, 1)
    X, Y = pr.grid()
    Xo, Yo = po.grid()
    ux, uy = lbm.lattice_velocity(q, pr, X, Y, 0.3)
    vx, vy = po.velocity(Xo, Yo, 0.3)
    assert np.array_equal(ux.T, po.u_max * vx) and np.array_equal(uy.T, po.u_max * vy)
    assert np.array_equal(lbm.lattice_pressure(q, pr, X, Y).T, po.u_max ** 2 * po.pressure(qo, Xo, Yo))
    assert lbm.dimensionless_velocity(pr, 0.5) == 0.5 / pr.u_max and lbm.dimensionless_stress(pr, 2.0) == 2.0 / pr.u_max ** 2
    assert lbm.dimensionless_force(pr, 1.0) == 1.0 / (pr.u_max * pr.delta_t()) and lbm.dimensionless_density(pr, 1.5) == 1.5
    assert np.array_equal(np.asarray(lbm.force(pr, X, Y))[0].T, np.asarray(po.force(Xo, Yo))[0])
    assert float(np.abs(np.asarray(lbm.force(lbm.CouetteFlow(1 / 6, 1), X, Y))).max()) == 0.0
    assert np.array_equal(lbm.range_(pr)[0], po.range()[0])
