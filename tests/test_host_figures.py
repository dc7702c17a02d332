"""The host mirror of the Julia API (latticeboltzmann.jl_b200/lbm: problems, initialisation strategies, force data, the
batching `simulate`, processing methods, stop criteria, unit scaling) run END TO END on the CPU against the reference's
notebook figures: the device context is replaced by an oracle-backed stand-in (tests/_oracle_context.py, test
infrastructure), everything above the C ABI is the product code.  The same test bodies run on the real CUDA library in
tests/test_gpu_figures.py; here a subset sized for the CPU suite."""
import copy

import pytest

import test_gpu_figures as G
from _oracle_context import emulated_backend


@pytest.fixture
def host(monkeypatch):
    with emulated_backend(monkeypatch) as lbm:
        yield lbm


def _trim_scales(monkeypatch, key, n):
    fig = copy.deepcopy(G.FIG)
    fig[key]["scales"] = fig[key]["scales"][:n]
    monkeypatch.setattr(G, "FIG", fig)


@pytest.mark.parametrize("name", ["D2Q9", "D2Q17"])
def test_shear_wave_convergence_through_the_host_mirror(host, monkeypatch, name):
    _trim_scales(monkeypatch, "shear_wave_convergence", 3)
    G.test_shear_wave_convergence_figure(name)


def test_tgv_convergence_through_the_host_mirror(host):
    G.test_tgv_convergence_figure(2.0)


@pytest.mark.parametrize("index", [3, 5])
def test_tgv_initialisation_strategies_through_the_host_mirror(host, index):
    """index 5: IterativeInitializationMeiEtAl(1.0, 1e-10) -- lbm.initialize drives the LBM_ITERATIVE_INIT context and the
    DensityConvergence reduce step by step, then the run proper starts from the downloaded populations."""
    G.test_tgv_initialisation_strategies_figure(index)


@pytest.mark.parametrize("name,u0", [("D2Q13", 0.12), ("D2Q37", 0.12), ("D2Q9", 0.03)])
def test_couette_moving_wall_through_the_host_mirror(host, monkeypatch, name, u0):
    _trim_scales(monkeypatch, "couette_convergence", 2 if name == "D2Q9" else 3)
    G.test_couette_moving_wall_figure(name, u0)


def test_poiseuille_tau_sweep_through_the_host_mirror(host, monkeypatch):
    monkeypatch.setattr(G, "POISEUILLE_INDICES", [0, 5, 49, 149, 449, 949])
    G.test_poiseuille_tau_sweep_figure()


@pytest.mark.parametrize("kind", ["decaying", "static"])
def test_shear_wave_snapshots_through_the_host_mirror(host, monkeypatch, kind):
    """TakeSnapshots + the host-side moment functions; static: the separable time-dependent force tables"""
    monkeypatch.setattr(G, "N_SNAPSHOTS", {"decaying": 4, "static": 3})
    G.test_shear_wave_snapshot_profiles_figure(kind)


def test_array_level_operators_reuse_their_scratch_context(host, monkeypatch):
    """collide!/stream!/apply! on plain arrays (the reference's tests and benchmarks call them in loops): one device context
    per (lattice, model, relaxation times, BCs, shape, dtype), kept alive between calls; results as the oracle's."""
    import numpy as np
    import oracle.lbm_oracle as O
    from conftest import random_populations, to_host_layout, to_oracle_layout
    from lbm import model
    model.clear_scratch_contexts()
    made = []
    inner = model.make_context
    monkeypatch.setattr(model, "make_context", lambda *a, **k: (made.append(a[:2]), inner(*a, **k))[1])
    q, qo = host.D2Q9(), O.L.D2Q9()
    f = random_populations(qo, 12, 7, seed=1)
    for _ in range(3):
        out = host.collide_(host.TRT(0.8, 1.1, None), q, to_host_layout(f))  # 3-argument form: (tau_s, tau_a, force)
    assert len(made) == 1
    assert np.array_equal(to_oracle_layout(out), O.collide(O.TRT(0.8, 1.1), qo, f))
    host.collide_(host.TRT(0.8, 1.2, None), q, to_host_layout(f))  # other relaxation times: another context
    assert len(made) == 2
    for _ in range(2):
        s = host.stream_(q, to_host_layout(f))
    assert len(made) == 3 and np.array_equal(to_oracle_layout(s), O.stream(qo, f))
    bcs = [host.BounceBack(host.North(), (1, 12), (1, 7)), host.MovingWall(host.North(), (1, 12), (1, 7), [0.01, 0.0])]
    ob = [O.BounceBack("N", (1, 12), (1, 7)), O.MovingWall("N", (1, 12), (1, 7), [0.01, 0.0])]
    want = O.stream(qo, f)
    O.apply_bcs(ob, qo, want, f)
    for _ in range(2):
        fn = to_host_layout(O.stream(qo, f))
        host.apply_(bcs, q, fn, to_host_layout(f))
    assert len(made) == 4 and np.array_equal(to_oracle_layout(fn), want)
    for k in range(model._SCRATCH_CACHE_SIZE + 2):  # the cache is bounded
        host.collide_(host.SRT(0.6 + 0.01 * k), q, to_host_layout(f))
    assert len(model._SCRATCH_CACHE) == model._SCRATCH_CACHE_SIZE
    model.clear_scratch_contexts()
    assert not model._SCRATCH_CACHE


@pytest.mark.parametrize("kind,n", [("poiseuille", 4), ("couette", 4)])
def test_wall_bounded_snapshots_through_the_host_mirror(host, monkeypatch, kind, n):
    monkeypatch.setattr(G, "N_SNAPSHOTS", dict(G.N_SNAPSHOTS, **{kind: n}))
    G.test_wall_bounded_snapshot_profiles_figure(kind)
