#!/usr/bin/env python
"""The reference's own BenchmarkTools suite (benchmark/bench_*.jl), case for case, through the drop-in.

    python tools/bench_suite.py [--scale 2] [--reps 30] [--suites simulation,collision_models,...] [--cpu]

Each case of /root/reference/benchmark is rebuilt with the same problem, lattice and arguments (names as in the Julia suite)
and timed through the host mirror of the Julia API on the CUDA library.  `--cpu` adds the CPU restatement (the oracle's C
implementation, OpenMP) of the array-level cases for comparison.  `--scale` is the `scale` argument of the suite's problems:
2 is what the reference benchmarks (32 x 32 nodes for TGV -- tiny: on a GPU those calls are pure launch + PCIe latency);
larger scales show where the device pays off.  One JSON line per case:
    {"suite", "lattice", "case", "impl", "nodes", "median_us", "min_us", "reps"}
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))


def timeit(fn, reps, setup=None):
    ts = []
    for _ in range(reps):
        arg = setup() if setup else None
        t0 = time.perf_counter()
        fn(arg) if setup else fn()
        ts.append(time.perf_counter() - t0)
    return 1e6 * statistics.median(ts), 1e6 * min(ts)


def emit(suite, lattice, case, impl, nodes, med, mn, reps):
    print(json.dumps({"suite": suite, "lattice": lattice, "case": case, "impl": impl, "nodes": nodes,
                      "median_us": round(med, 2), "min_us": round(mn, 2), "reps": reps}), flush=True)


def bench_simulation(lbm, a):
    """benchmark/bench_simulation.jl:11-36: CouetteFlow(cs (tau - 0.5), scale), tau = 1; collide!, stream!,
    apply_boundary_conditions!, next!(model, 0), simulate(model, 0:10) on a LatticeBoltzmannModel."""
    for q in lbm.Quadratures:
        problem = lbm.CouetteFlow((1 / q.speed_of_sound_squared) * (1.0 - 0.5), a.scale)
        nodes = problem.NX * problem.NY

        def new_model():
            return lbm.LatticeBoltzmannModel(problem, q, process_method=lbm.ProcessingMethod(problem, False, 0))

        model = new_model()
        cases = {
            "collision": lambda: lbm.collide_model_(model, 0.0),
            "stream": lambda: lbm.stream_model_(model),
            "boundary conditions": lambda: lbm.apply_boundary_conditions_(model, 0.0),
            "processing": lambda: lbm.next_model_(model, 0),
            "simulate": lambda: (lbm.simulate(model, range(0, 11)), model.ctx.sync()),
        }
        lbm.collide_model_(model, 0.0)  # stream!/apply! need an f_collision
        for name, fn in cases.items():
            fn()
            model.ctx.sync()
            med, mn = timeit(lambda: (fn(), model.ctx.sync()), a.reps)
            emit("simulation", q.name, name, "b200", nodes, med, mn, a.reps)
        model.close()


def _zero_velocity_f(lbm, q, problem):
    return lbm.initialize(lbm.ZeroVelocityInitialCondition(), q, problem)


def bench_collision_models(lbm, a):
    """benchmark/bench_collision_models.jl:18-82: collide!(cm, q, f_in, f_out) on TGV(q, 1.0, scale) arrays."""
    cpu = a.cpu
    if cpu:
        import oracle.lbm_oracle as O
        from oracle.c_oracle import COracle
    for q in (lbm.D2Q9(), lbm.D2Q13(), lbm.D2Q17(), lbm.D2Q21(), lbm.D2Q37()):
        f = _zero_velocity_f(lbm, q, lbm.TGV(q, 1.0, a.scale))
        nodes = f.shape[0] * f.shape[1]
        shear = lbm.DecayingShearFlow(1.0 / (2 * q.speed_of_sound_squared), a.scale, static=True)
        f_force = _zero_velocity_f(lbm, q, shear)
        cases = [
            ("srt 1.0", lbm.SRT(1.0), f), ("srt 1.3", lbm.SRT(1.3), f), ("srt 0.8", lbm.SRT(0.8), f),
            ("trt (1.0, 1.0)", lbm.TRT(1.0, 1.0), f), ("trt (0.9, 1.1)", lbm.TRT(0.9, 1.1), f),
            ("mrt-equilibrium 1.0", lbm.MRT(q, 1.0), f),
            ("srt force", lbm.CollisionModel(lbm.SRT, q, shear), f_force),
            ("trt force", lbm.CollisionModel(lbm.TRT, q, shear), f_force),
        ]
        for name, cm, f_in in cases:
            f_out = f_in.copy(order="F")
            lbm.collide_(cm, q, f_in, f_out)
            med, mn = timeit(lambda: lbm.collide_(cm, q, f_in, f_out), a.reps)
            emit("collision_models", q.name, name, "b200", f_in.shape[0] * f_in.shape[1], med, mn, a.reps)
        if cpu:
            qo = O.L.BY_NAME[q.name]()
            fo = np.ascontiguousarray(np.transpose(f, (2, 1, 0)))
            for name, cmo in (("srt 0.8", O.SRT(0.8)), ("trt (0.9, 1.1)", O.TRT(1.1, 0.9)), ("mrt-equilibrium 1.0", O.MRT(qo, 1.0))):
                co = COracle(qo, cmo)
                co.collide(fo)
                med, mn = timeit(lambda: co.collide(fo), a.reps)
                emit("collision_models", q.name, name, "cpu-restatement", nodes, med, mn, a.reps)


def bench_boundary_conditions(lbm, a):
    """benchmark/bench_boundary_conditions.jl:20-34: apply!(BounceBack(direction, 1:NX, 1:NY), q, f_new, f_old)."""
    for q in lbm.Quadratures:
        f = _zero_velocity_f(lbm, q, lbm.TGV(q, 1.0, a.scale))
        nx, ny, _ = f.shape
        for direction in (lbm.North(), lbm.East(), lbm.South(), lbm.West()):
            bc = lbm.BounceBack(direction, (1, nx), (1, ny))
            f_new, f_old = f.copy(order="F"), f.copy(order="F")
            lbm.apply_(bc, q, f_new, f_old)
            med, mn = timeit(lambda: lbm.apply_(bc, q, f_new, f_old), a.reps)
            emit("boundary_conditions", q.name, f"bounce back {type(direction).__name__}", "b200", nx * ny, med, mn, a.reps)


def bench_equilibria(lbm, a):
    """benchmark/bench_equilibria.jl:20-37: single-node equilibrium(q, rho, u, T) / equilibrium!(q, rho, u, T, f)
    (host-side functions of the mirror: one node is not device work)."""
    for q in lbm.Quadratures:
        u = np.zeros(2)
        f = np.zeros(q.Q)
        med, mn = timeit(lambda: lbm.equilibrium(q, 1.0, u, 1.0), a.reps)
        emit("equilibria", q.name, "initial equilibrium", "host", 1, med, mn, a.reps)
        med, mn = timeit(lambda: lbm.equilibrium_(q, 1.0, u, 1.0, f), a.reps)
        emit("equilibria", q.name, "compute equilibrium", "host", 1, med, mn, a.reps)


def bench_moments(lbm, a):
    """benchmark/bench_moments.jl:12-57: single-node density, velocity!, pressure, temperature, momentum_flux,
    deviatoric_tensor of f = q.weights (host-side functions of the mirror)."""
    for q in lbm.Quadratures:
        f = np.array(q.weights, dtype=np.float64)
        rho = lbm.density(q, f)
        u = lbm.velocity(q, f, rho)
        cases = {
            "density": lambda: lbm.density(q, f),
            "velocity": lambda: lbm.velocity_(q, f, rho, u),
            "pressure": lambda: lbm.pressure(q, f, rho, u),
            "temperature": lambda: lbm.temperature(q, f, rho, u),
            "momentum_flux": lambda: lbm.momentum_flux(q, f, rho, u),
            "deviatoric_tensor": lambda: lbm.deviatoric_tensor(q, 1.0, f, rho, u),
        }
        for name, fn in cases.items():
            med, mn = timeit(fn, a.reps)
            emit("moments", q.name, name, "host", 1, med, mn, a.reps)


SUITES = {"simulation": bench_simulation, "collision_models": bench_collision_models,
          "boundary_conditions": bench_boundary_conditions, "equilibria": bench_equilibria, "moments": bench_moments}


def main(argv=None, lbm=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=2)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--suites", default=",".join(SUITES))
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args(argv)
    if lbm is None:
        import lbm
    for name in a.suites.split(","):
        SUITES[name](lbm, a)
    return 0


if __name__ == "__main__":
    sys.exit(main())
