#!/usr/bin/env python
"""The reference's largest recorded job through the batched path: `main([D2Q9()], 1, 0.51:0.01:10.0, 0.51:0.01:10.0)` of
examples/notebooks/trt_magic_parameter.ipynb:30-103 -- 950 x 950 = 902 500 solves of a 3 x 5 D2Q9 TRT + force Poiseuille
flow, velocity-convergence stop (1e-7, every 100 steps), at most 5000 steps each; 3 h 00 min on one thread in the notebook.

    python tools/sweep_trt_magic.py [--step 0.01] [--out profiles/r02_trt_magic_sweep.json]

Prints one JSON line: wall seconds of the whole `simulate_many` call (host set-up, H2D, the batch launch, on-device error
norms, D2H), device milliseconds of the batch launch, lattice updates actually performed, and the check of the 41 rows
the notebook prints (and of the 950-point diagonal tau_s == tau_a that poiseuille.ipynb plots) against the golden table.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))
import lbm  # noqa: E402


def close_to_printed(value, printed, digits=6):
    if np.isinf(printed):
        return bool(np.isinf(value))
    ulp = 10.0 ** (np.floor(np.log10(abs(printed))) - (digits - 1))
    return bool(abs(value - printed) <= 0.51 * ulp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--step", type=float, default=0.01)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--arith", default="exact")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    q = lbm.D2Q9()
    n = int(round((10.0 - 0.51) / a.step)) + 1
    taus = 0.51 + a.step * np.arange(n)          # Julia range(0.51, stop = 10.0, step = 0.01)
    taus = np.round(taus, 10)
    t0 = time.perf_counter()
    problems = [lbm.PoiseuilleFlow((ts - 0.5) / q.speed_of_sound_squared, 1) for ts in taus]
    # for tau_s in range, tau_a in range (tau_s outer, as Julia's `for tau_s in r, tau_a in r` ... push!)
    ts_idx, ta_idx = np.divmod(np.arange(n * n), n)
    pairs = np.stack([taus[ts_idx], taus[ta_idx]], axis=1)
    sc = lbm.VelocityConvergenceStoppingCriteria(1e-7, problems[0])
    res = lbm.simulate_many(problems, q, pairs, lbm.TRT, problem_index=ts_idx, t_end=100.0, stop_criteria=sc,
                            initialization_strategy=lbm.ZeroVelocityInitialCondition(), dtype=a.dtype, arith=a.arith)
    wall = time.perf_counter() - t0
    updates = int(res.timestep.sum()) * problems[0].NX * problems[0].NY
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "trt_magic_parameter.json")))["rows"]
    ok, checked = 0, 0
    for r in golden:
        i = int(round((r["tau_s"] - 0.51) / a.step))
        j = int(round((r["tau_a"] - 0.51) / a.step))
        if i >= n or j >= n or abs(taus[i] - r["tau_s"]) > 1e-9 or abs(taus[j] - r["tau_a"]) > 1e-9:
            continue
        k = i * n + j
        checked += 1
        ok += close_to_printed(res.error_u[k], r["error_u"]) and close_to_printed(res.error_p[k], r["error_p"])
    diag = res.error_u[np.arange(n) * n + np.arange(n)]
    out = dict(job="trt_magic_parameter.ipynb main([D2Q9()], 1, taus, taus)", solves=n * n, tau_step=a.step, dtype=a.dtype,
               arith=a.arith, wall_s=round(wall, 3), device_ms=round(res.device_ms, 2), lattice_updates=updates,
               glups_device=round(updates / (res.device_ms * 1e-3) / 1e9, 2), stopped_early=int(res.stopped.sum()),
               mean_steps=float(res.timestep.mean()), golden_rows_checked=checked, golden_rows_ok=int(ok),
               diagonal_error_u_min=float(np.nanmin(diag)), diagonal_argmin_tau=float(taus[int(np.nanargmin(diag))]),
               reference="3 h 00 min 05 s on one CPU thread (trt_magic_parameter.ipynb:103)",
               speedup_vs_reference_notebook=round(10805.0 / wall, 1))
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
