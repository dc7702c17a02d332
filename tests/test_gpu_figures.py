"""GPU path against the numbers behind the reference's notebook FIGURES (tests/golden/notebook_figures.json, recovered from
the SVG the notebooks store by tests/golden/extract_notebook_plots.py): the same call sequences the notebooks make, through
the host mirror of the Julia API and the CUDA library.  This is parity with outputs of the REFERENCE ITSELF -- not with our
restatement -- for every lattice, the moving wall on the multi-speed lattices, every initialisation strategy including the
Mei et al. iteration, and 841-step time series of all four error norms."""
import json
import os

import numpy as np
import pytest

import lbm

pytestmark = pytest.mark.gpu

FIG = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_figures.json")))
RTOL = 2e-4  # the SVG coordinates are good to ~5e-5
F32_RTOL = 1e-2  # measured on B200: <= 2e-4 at N = 8, 16 and 2.2e-3 / 3.6e-3 (u / sigma_xy) at N = 32 for D2Q9; D2Q37 <= 1.2e-3
LATTICES = ["D2Q4", "D2Q5", "D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"]
POISEUILLE_INDICES = list(range(0, 950, 10)) + [949]
N_SNAPSHOTS = {"decaying": 4, "static": 4, "poiseuille": 4, "couette": 8}


def close(value, ref, rtol=RTOL):
    return abs(value - ref) <= rtol * abs(ref)


@pytest.mark.parametrize("name", LATTICES)
def test_shear_wave_convergence_figure(name):
    """shear_wave.ipynb cell 12 / notebook_examples.jl:34-69: static shear wave, SRT + force, tau = 0.8, N = 8 ... 64."""
    ref = FIG["shear_wave_convergence"]
    q = getattr(lbm.Quadratures, name)
    for i, scale in enumerate(ref["scales"]):
        problem = lbm.DecayingShearFlow(0.8 / (2.0 * q.speed_of_sound_squared), scale, static=True)
        n_steps = round(1.0 / problem.delta_t())
        pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.NoStoppingCriteria())
        res = lbm.simulate(problem, q, process_method=pm, initialization_strategy=lbm.AnalyticalEquilibrium(), t_end=1.0)
        row = res.processing_method.df[-1]
        res.close()
        for k in ("error_u", "error_p", "error_sxy"):
            want = ref["errors"][k][name]["value"][i]
            assert close(row[k], want), (name, scale, k, row[k], want)


@pytest.mark.parametrize("tau", [3.0, 2.0, 1.0, 0.8])
def test_tgv_convergence_figure(tau):
    """taylor_green_vortex.ipynb cells 3-5: TGV decay on 31 s x 17 s, s = 1, 2, 4; u, p, sigma_xy, sigma_xx errors."""
    ref = FIG["tgv_convergence"]
    q = lbm.D2Q9()
    i = ref["taus"].index(tau)
    for j, scale in enumerate(ref["scales"]):
        problem = lbm.TGV(q, tau, scale, 31 * scale, 17 * scale, np.sqrt(0.01) / scale)
        t_end = round(lbm.decay_time(problem))
        model = lbm.LatticeBoltzmannModel(problem, q, initialization_strategy=lbm.AnalyticalEquilibrium(),
                                          process_method=lbm.ProcessingMethod(problem, False, t_end))
        lbm.simulate(model, range(1, t_end + 1))
        row = model.processing_method.df[-1]
        model.close()
        for k in ("error_u", "error_p", "error_sxy", "error_sxx"):
            assert close(row[k], ref["errors"][k][i][j]), (tau, scale, k, row[k], ref["errors"][k][i][j])


def _strategy(name):
    if name.startswith("IterativeInitializationMeiEtAl"):
        tau, eps = name.split("(")[1].rstrip(")").split(",")
        return lbm.IterativeInitializationMeiEtAl(float(tau), float(eps))
    return getattr(lbm, name)()


@pytest.mark.parametrize("index", range(6))
def test_tgv_initialisation_strategies_figure(index):
    """taylor_green_vortex.ipynb cells 7-9: TGV(D2Q9(), 0.8, 2, 96, 72, 0.03), one run per initialisation strategy with
    ProcessingMethod(problem, true, t_end): a TrackHydrodynamicErrors row after every step (841 rows), four norms.  The two
    Mei et al. runs initialise on the device (LBM_ITERATIVE_INIT + the literal single-node DensityConvergence)."""
    ref = FIG["tgv_init_strategies"]
    q = lbm.D2Q9()
    problem = lbm.TGV(q, 0.8, 2, 96, 72, 0.03)
    t_end = round(lbm.decay_time(problem))
    assert t_end == 840
    strategy = _strategy(ref["strategies"][index])
    model = lbm.LatticeBoltzmannModel(problem, q, initialization_strategy=strategy,
                                      process_method=lbm.ProcessingMethod(problem, True, t_end))
    lbm.simulate(model, range(1, t_end + 1))
    df = model.processing_method.df
    model.close()
    assert len(df) == ref["n_rows"]
    for k in ("error_u", "error_p", "error_sxx", "error_sxy"):
        got = np.array([df[r - 1][k] for r in ref["rows"]])
        want = np.array(ref["errors"][k][index])
        rel = np.abs(got - want) / np.abs(want)
        assert rel.max() < RTOL, (ref["strategies"][index], k, float(rel.max()), ref["rows"][int(rel.argmax())])


@pytest.mark.parametrize("name", ["D2Q9", "D2Q13", "D2Q17", "D2Q21", "D2Q37"])
@pytest.mark.parametrize("u0", [0.01, 0.03, 0.12])
def test_couette_moving_wall_figure(name, u0):
    """couette.ipynb cells 6-7: CouetteFlow(1.0, u_0 / scale, nu, 1, 5 scale, (1.0, 1.0)) -- MovingWall North + BounceBack
    South -- run to VelocityConvergenceStoppingCriteria(1e-7) or t_end = 1; error_u over scale = 1, 2, 4, 8."""
    ref = FIG["couette_convergence"]
    q = getattr(lbm.Quadratures, name)
    for i, scale in enumerate(ref["scales"]):
        problem = lbm.CouetteFlow.fields(1.0, u0 / scale, 0.8 / (2.0 * q.speed_of_sound_squared), 1, 5 * scale, (1.0, 1.0))
        n_steps = round(1.0 / problem.delta_t())
        pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.VelocityConvergenceStoppingCriteria(1e-7, problem))
        res = lbm.simulate(problem, q, process_method=pm, initialization_strategy=lbm.ZeroVelocityInitialCondition(), t_end=1.0)
        got = res.processing_method.df[-1]["error_u"]
        res.close()
        want = ref["error_u"][str(u0)][name][i]
        # (errors of 1e-9 and below are differences of nearly equal numbers: a looser bar there)
        assert close(got, want, RTOL if want > 1e-8 else 2e-2), (name, u0, scale, got, want)


def test_poiseuille_tau_sweep_figure():
    """poiseuille.ipynb cells 6-9: error_u of the D2Q9 TRT(tau, tau, force) Poiseuille solve, every 10th of the 950 values."""
    ref = FIG["poiseuille_tau_sweep"]
    q = lbm.D2Q9()
    for index in POISEUILLE_INDICES:
        tau = ref["tau"][index]
        problem = lbm.PoiseuilleFlow((tau - 0.5) / q.speed_of_sound_squared, 1)
        n_steps = round(100.0 / problem.delta_t())
        pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.VelocityConvergenceStoppingCriteria(1e-7, problem))
        cm = lbm.TRT(tau, tau, lbm.LatticeForce(problem))
        res = lbm.simulate(problem, q, t_end=100.0, should_process=False, collision_model=cm, process_method=pm,
                           initialization_strategy=lbm.ZeroVelocityInitialCondition())
        got = res.processing_method.df[-1]["error_u"]
        res.close()
        assert close(got, ref["error_u"][index]), (tau, got, ref["error_u"][index])


@pytest.mark.parametrize("kind", ["decaying", "static"])
def test_shear_wave_snapshot_profiles_figure(kind):
    """shear_wave.ipynb cells 3-4 / 6-7: TakeSnapshots + plot_snapshots (notebook_examples.jl:71-240): dimensionless
    sigma_xx and sigma_xy along x at y_pos = round(Int, NY / 2).  decaying = a travelling, decaying wave without force;
    static = spin-up from rest under the time-dependent force (lbm_set_force_separable), 24 901 steps."""
    ref = FIG["shear_wave_snapshots"][kind]
    q = lbm.D2Q9()
    if kind == "decaying":
        problem = lbm.DecayingShearFlow(1 / 6, 4, static=False, A=3.0)
        every = [round(s / problem.delta_t()) + 1 for s in (0.0, 0.05, 0.15, 0.25)]
        strategy = lbm.AnalyticalEquilibrium()
    else:
        problem = lbm.DecayingShearFlow(1 / 6, 16, static=True, A=0.5)
        nu = problem.viscosity()
        every = [round(s / (nu * problem.delta_t())) for s in (0.01, 0.1, 1.0, 10.0)]
        strategy = lbm.ZeroVelocityInitialCondition()
    every = every[:N_SNAPSHOTS[kind]]
    pm = lbm.TakeSnapshots(problem, every)
    model = lbm.LatticeBoltzmannModel(problem, q, initialization_strategy=strategy, process_method=pm)
    lbm.simulate(model, range(0, every[-1]))
    model.close()
    assert pm.timesteps[:len(every)] == every
    tau = q.speed_of_sound_squared * problem.lattice_viscosity()
    y_pos = round(problem.NY / 2) - 1
    for a, b, name in ((0, 0, "sigma_xx"), (0, 1, "sigma_xy")):
        scale = np.abs(np.array(ref[name])).max()  # the figure's resolution is a fraction of its axis range
        for k in range(len(every)):
            f = pm.snapshots[k]
            rho = lbm.density(q, f)
            u = lbm.velocity(q, f, rho)
            sigma = problem.dimensionless_stress(lbm.deviatoric_tensor(q, tau, f, rho, u))
            got = sigma[:, y_pos, a, b]
            assert np.abs(got - np.array(ref[name][k])).max() < 5e-5 * scale, (kind, name, k)


@pytest.mark.parametrize("kind", ["poiseuille", "couette"])
def test_wall_bounded_snapshot_profiles_figure(kind):
    """poiseuille.ipynb cells 3-4 / couette.ipynb cells 2-4: spin-up from rest between walls (bounce-back N+S with a force;
    moving wall N + bounce-back S), TakeSnapshots, sigma_xx and sigma_xy along y at x_pos = max(round(Int, NX / 2), 1);
    the last Couette snapshot is 192 001 steps in."""
    ref = FIG["wall_snapshots"][kind]
    q = lbm.D2Q9()
    if kind == "poiseuille":
        problem, snap, plus = lbm.PoiseuilleFlow(1 / 6, 4), (0.01, 0.05, 0.1, 1.0), 0
    else:
        problem, snap, plus = lbm.CouetteFlow(1 / 6, 16), (0, 0.005, 0.01, 0.05, 0.1, 0.5, 1.0, 5.0), 1
    nu = problem.viscosity()
    every = [round(s / (nu * problem.delta_t())) + plus for s in snap][:N_SNAPSHOTS[kind]]
    pm = lbm.TakeSnapshots(problem, every)
    model = lbm.LatticeBoltzmannModel(problem, q, initialization_strategy=lbm.ZeroVelocityInitialCondition(), process_method=pm)
    lbm.simulate(model, range(0, every[-1]))
    model.close()
    assert pm.timesteps[:len(every)] == every
    tau = q.speed_of_sound_squared * problem.lattice_viscosity()
    x_pos = max(round(problem.NX / 2), 1) - 1
    for a, b, name in ((0, 0, "sigma_xx"), (0, 1, "sigma_xy")):
        scale = np.abs(np.array(ref[name])).max()
        for k in range(len(every)):
            f = pm.snapshots[k]
            rho = lbm.density(q, f)
            u = lbm.velocity(q, f, rho)
            sigma = problem.dimensionless_stress(lbm.deviatoric_tensor(q, tau, f, rho, u))
            assert np.abs(sigma[x_pos, :, a, b] - np.array(ref[name][k])).max() < 5e-5 * scale, (kind, name, k)


@pytest.mark.parametrize("name", ["D2Q9", "D2Q37"])
def test_float32_reproduces_the_shear_wave_figure(name):
    """Float32 contexts (deviation storage f - w; a capability the Float64-only reference does not have) against the
    reference's own numbers: the velocity and shear-stress errors of the shear-wave study, N = 8, 16, 32 (1019 steps at
    N = 32).  The error norms are differences between the numerical and the analytic field, so Float32 round-off shows up
    relative to the error itself; error_p (1e-5 ... 1e-7) is below that floor and not compared."""
    ref = FIG["shear_wave_convergence"]
    q = getattr(lbm.Quadratures, name)
    worst = {}
    for i, scale in enumerate(ref["scales"][:3]):
        problem = lbm.DecayingShearFlow(0.8 / (2.0 * q.speed_of_sound_squared), scale, static=True)
        n_steps = round(1.0 / problem.delta_t())
        pm = lbm.TrackHydrodynamicErrors(problem, False, n_steps, lbm.NoStoppingCriteria())
        res = lbm.simulate(problem, q, process_method=pm, initialization_strategy=lbm.AnalyticalEquilibrium(), t_end=1.0,
                           dtype="f32")
        row = res.processing_method.df[-1]
        res.close()
        for k in ("error_u", "error_sxy"):
            want = ref["errors"][k][name]["value"][i]
            worst[(k, scale)] = abs(row[k] - want) / want
    assert max(worst.values()) < F32_RTOL, worst
