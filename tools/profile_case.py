#!/usr/bin/env python
"""One <lattice, model, dtype, arith> case of the fused kernel (and optionally the diagnostics kernels) for ncu captures
and quick event-timed checks:

    ncu --set full --clock-control none -k regex:k_step -s 6 -c 1 -o out python tools/profile_case.py --lattice D2Q37 --model MRT --dtype f32
    python tools/profile_case.py --lattice D2Q9 --diag          # k_moments / k_reduce / k_errors timings

Prints one JSON line: ms per step, MLUPS, GB/s against B_alg = 2 Q sizeof(T), fraction of the measured HBM peak
(diagnostics: against Q sizeof(T) bytes per node).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))
import lbm  # noqa: E402
from lbm import _abi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", default="D2Q9")
    ap.add_argument("--model", default="TRT")
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--arith", default="fast")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--prefetch", type=int, default=-1, help="L2-prefetch distance in rows (0 = off, -1 = automatic)")
    ap.add_argument("--store", type=int, default=0, help="tuning builds: 1 st.cs, 2 st.cg, 3 st.wt for the population stores")
    ap.add_argument("--tma", type=int, default=2, help="TMA-staged fused kernel: 0 off, 1 on, 2 automatic")
    ap.add_argument("--tma-cfg", type=int, default=0, help="100 * (CTA width / 128) + 10 * stages + CTAs per SM; 0 = default")
    ap.add_argument("--check", action="store_true", help="compare the populations after the run with a run of the register kernel")
    ap.add_argument("--walls", action="store_true")
    ap.add_argument("--persistent", type=int, default=2, help="0: launch per step / graphs, 1: persistent kernel, 2: automatic")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--diag", action="store_true", help="time lbm_moments / lbm_reduce / lbm_reduce_errors instead")
    ap.add_argument("--sustain", type=float, default=0.0, help="repeat the timed batch for at least this many seconds")
    a = ap.parse_args()
    peak = 6548.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    q = getattr(lbm.Quadratures, a.lattice)
    n = a.n or (4096 if q.Q <= 13 else 2048)
    ny = a.ny or n
    es = 8 if a.dtype == "f64" else 4
    x = (np.arange(n) + 0.5) * (2 * np.pi / n)
    y = (np.arange(ny) + 0.5) * (2 * np.pi / ny)
    pert = 1e-3 * np.sin(x)[:, None] * np.cos(y)[None, :]
    f0 = np.empty((n, ny, q.Q), order="F")
    for i in range(q.Q):
        f0[:, :, i] = q.weights[i] * (1 + pert * (1 + 0.1 * i))
    code = {"SRT": _abi.SRT, "TRT": _abi.TRT, "MRT": _abi.MRT}[a.model]
    taus = {"SRT": [0.8], "TRT": [0.8, 1.3333333333333333], "MRT": [0.8, 0.8, 0.8, 0.8]}[a.model]
    bcs = []
    if a.walls:
        bcs = [lbm.BounceBack(lbm.South(), (1, n), (1, ny)).to_abi(), lbm.MovingWall(lbm.North(), (1, n), (1, ny), [0.001, 0.0]).to_abi()]
    with _abi.Context(n, ny, a.lattice, code, taus, bcs, dtype=_abi.F64 if a.dtype == "f64" else _abi.F32,
                      arith=_abi.ARITH_FAST if a.arith == "fast" else _abi.ARITH_EXACT) as c:
        c.set_option("variant", a.variant)
        c.set_option("tma", a.tma)
        c.set_option("prefetch", a.prefetch)
        c.set_option("store", a.store)
        c.set_option("tma_cfg", a.tma_cfg)
        c.set_option("persistent", a.persistent)
        c.set_option("graph", a.graph)
        c.upload_f(f0)
        c.step(0, 4)
        c.sync()
        out = dict(lattice=a.lattice, model=a.model, dtype=a.dtype, arith=a.arith, variant=a.variant, tma=a.tma, tma_cfg=a.tma_cfg, prefetch=a.prefetch, store=a.store, n=n, ny=ny,
                   persistent=a.persistent, graph=a.graph)
        if a.diag:
            import time
            one = np.ones(n)
            exp = [(1.0, [])] + [(0.0, [(1.0, np.sin(x), np.cos(y))])] * 2 + [(1.0, [])] + [(0.0, [(1.0, one, np.ones(ny))])] * 4
            res = {}
            for name, fn in (("reduce_mean_ux", lambda: c.reduce(_abi.REDUCE_MEAN_UX)),
                             ("reduce_velocity_change", lambda: c.reduce(_abi.REDUCE_VELOCITY_CHANGE)),
                             ("reduce_conserved", lambda: c.reduce(_abi.REDUCE_CONSERVED)),
                             ("reduce_errors", lambda: c.reduce_errors(0.3, 0.01, exp)),
                             ("reduce_process", lambda: c.reduce_process(0.01, exp)),
                             ("moments_rho_u", lambda: c.moments(0.3, ("rho", "ux", "uy")))):
                fn()
                t0 = time.perf_counter()
                dev_ms = 0.0
                for _ in range(5):
                    c.timer_start()   # CUDA events on the library's stream: kernels + table upload + result copy
                    fn()
                    dev_ms += c.timer_stop()
                dt = (time.perf_counter() - t0) / 5
                dev = dev_ms / 5 * 1e-3
                res[name] = dict(host_ms=round(dt * 1e3, 4), device_ms=round(dev * 1e3, 4),
                                 gbs_alg=round(q.Q * es * n * ny / dev / 1e9, 1),
                                 frac=round(q.Q * es * n * ny / dev / 1e9 / peak, 4))
            out["diag"] = res
        else:
            import time
            t_end = time.perf_counter() + a.sustain
            tot_ms, tot_steps = 0.0, 0
            while True:
                l0 = c.kernel_launches
                c.timer_start()
                c.step(0, a.steps)
                tot_ms += c.timer_stop()
                out["launches_per_batch"] = c.kernel_launches - l0
                tot_steps += a.steps
                if time.perf_counter() >= t_end:
                    break
            ms = tot_ms / tot_steps
            mlups = n * ny / (ms * 1e-3) / 1e6
            gbs = mlups * 1e6 * 2 * q.Q * es / 1e9
            out.update(ms_per_step=round(ms, 5), steps=tot_steps, mlups=round(mlups, 1), gbs=round(gbs, 1), frac=round(gbs / peak, 4))
            if a.check:
                res = []
                for tma in (a.tma, 0):
                    c.set_option("tma", tma)
                    c.upload_f(f0)
                    c.step(0, 7)
                    res.append(c.download_f())
                out["check_max_abs_diff_vs_register_kernel"] = float(np.abs(res[0] - res[1]).max())
                out["check_identical"] = bool(np.array_equal(res[0], res[1]))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
