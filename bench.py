#!/usr/bin/env python
"""bench.py -- MLUPS / HBM-roofline benchmark of the fused collide-stream path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N = 1): BASELINE.json configs[1] -- D2Q9 TRT Taylor-Green vortex decay, 4096 x 4096
periodic, Float64 (TGV(D2Q9(), 0.8, 256); CollisionModel(TRT, ...): tau_s = 0.8, Lambda = 1/4).
N > 1: y-slab weak scaling, one 4096 x 4096 slab per GPU (global NY = 4096 N); the launch that computes a
slab's boundary rows stores them straight into the neighbours' ghost rows over NVLink (peer memory, device-side
flags); `--p2p 0` selects the NCCL send/recv path on a side stream instead.

One bench "step" = one device batch of `--inner` (default 100) lattice time steps: 100 is the
reference's host-visible cadence (`next!` checks its stop criterion every 100 steps,
src/processing_methods/track_hydrodynamic_errors.jl:55).
  value : MLUPS with populations resident in HBM (K batches, CUDA events on the library's stream,
          max over ranks).
  e2e   : the same through the C ABI with HOST buffers: every bench step uploads f from pinned host
          memory (lbm_upload_f), runs the batch (lbm_step) and downloads f (lbm_download_f).
  roofline : achieved = B_alg * nodes / mean kernel time; B_alg = 2 Q sizeof(T) = 144 B per
          lattice update for D2Q9 Float64; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the C restatement of the reference algorithm (oracle/lbm_oracle.c, OpenMP over rows,
          all host cores) on the same 4096 x 4096 grid for a bounded number of steps (40).
`--impl reference` times that CPU restatement alone (the reference is Julia-only and cannot run
in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))

METRIC = "MLUPS (D2Q9 TRT collide-stream, million lattice updates per second)"
Q, BYTES = 9, {"f64": 8, "f32": 4}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--inner", type=int, default=100, help="lattice time steps per bench step")
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=4096, help="rows PER GPU")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--arith", default=os.environ.get("LBM_BENCH_ARITH", "fast"), choices=["exact", "fast"],
                    help="fast (default): FMA contraction, within the 1e-12 parity tolerance; exact: reference operation "
                         "order, bit-identical to the oracle")
    ap.add_argument("--lattice", default="D2Q9")
    ap.add_argument("--collision", default="TRT", choices=["SRT", "TRT", "MRT"])
    ap.add_argument("--variant", type=int, default=int(os.environ.get("LBM_BENCH_VARIANT", "0")))
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5s", "C5w"],
                    help="BASELINE.json config preset (C2 = default bench workload; see build_case)")
    ap.add_argument("--p2p", type=int, default=1, help="N > 1: 1 = peer-memory halo stores (default), 0 = NCCL send/recv")
    ap.add_argument("--graph", type=int, default=1, help="1 = replay CUDA graphs of 16 fused steps (default), 0 = plain launches")
    ap.add_argument("--persistent", type=int, default=2, help="0 = launches per step (graphs), 1 = persistent multi-step kernel wherever possible, 2 = automatic")
    ap.add_argument("--overlap", type=int, default=-1, help="N > 1, peer memory: 1 = boundary-row launch + interior launch per step, 2 = one merged launch, 0 = whole slab in one launch after the exchange; -1 = the library's default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e with one job at a time only (no further contexts in flight)")
    ap.add_argument("--e2e-jobs", type=int, default=3, help="independent jobs in flight for the e2e figure (N = 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing multi-slab parity check against the oracle")
    ap.add_argument("--no-also", action="store_true", help="skip the strong-scaling configs reported under `also` (C5s, C3, C4)")
    ap.add_argument("--also-shrink", type=int, default=1, help="divide the grids of the `also` configs by this (tests)")
    ap.add_argument("--cpu-n", type=int, default=4096, help="CPU arm grid (default: the full 4096 x 4096 workload grid)")
    ap.add_argument("--cpu-steps", type=int, default=40, help="lattice steps of the cpu_baseline sample (reference arm: a tenth per bench step)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle's C restatement (bench.py may execute oracle/ only here)
# ---------------------------------------------------------------------------------------------
def host_cores():
    """Cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not inherit
    that (round-1 SCALE records at N >= 2 ran the reference arm on one core), so the thread count is set explicitly."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuArm:
    """The C restatement of the reference algorithm on the host cores: state prepared once, then timed in place."""

    def __init__(self, n, lattice="D2Q9", collision="TRT"):
        import oracle.lbm_oracle as O
        from oracle.c_oracle import COracle, lib as oracle_lib, num_threads
        oracle_lib().oracle_set_threads(host_cores())
        q = O.L.BY_NAME[lattice]()
        pr = O.TGV(q, 0.8, max(n // 16, 1), NX=n, NY=n)
        cm = O.collision_model(collision, q, pr)
        X, Y = pr.grid()
        ux, uy = pr.velocity(X, Y)
        self.fs = np.ascontiguousarray(np.stack(O.equilibrium_collision(q, pr.density(q, X, Y), ux, uy)))
        self.fc = np.empty_like(self.fs)
        self.co = COracle(q, cm)
        self.n, self.threads = n, num_threads()
        self.co.steps_inplace(self.fs, self.fc, 2)  # warm the pages

    def run(self, steps):
        """-> (MLUPS, seconds) of `steps` collide-stream steps on the n x n grid"""
        t0 = time.perf_counter()
        self.co.steps_inplace(self.fs, self.fc, steps)
        dt = time.perf_counter() - t0
        return self.n * self.n * steps / dt / 1e6, dt


def cpu_restatement_mlups(n, steps, lattice="D2Q9", collision="TRT"):
    arm = CpuArm(n, lattice, collision)
    v, secs = arm.run(steps)
    return v, arm.threads, secs


def config_keys(workload, preset, grid_per_gpu, world, inner, arith, dtype, nq=Q, scaling="weak"):
    """The `config` object both arms print: same keys, same values for the same workload (the driver compares them)."""
    nx, nyl = int(grid_per_gpu[0]), int(grid_per_gpu[1])
    return {"workload": workload, "preset": preset, "grid_per_gpu": [nx, nyl],
            "grid_global": [nx, nyl * world if scaling == "weak" else nyl], "lattice_steps_per_bench_step": int(inner),
            "arith": arith, "parallelism": f"y-slabs x{world}",
            "l2": "working set (2 x %.2f GB per GPU) >> 126 MB L2; no flush needed" % (nx * nyl * nq * BYTES[dtype] / 1e9)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, inner = a.cpu_n, max(1, a.cpu_steps // 10)
    arm = CpuArm(n, a.lattice, a.collision)
    for _ in range(a.warmup):
        arm.run(inner)
    secs = 0.0
    for _ in range(a.steps):
        secs += arm.run(inner)[1]
    value = n * n * inner * a.steps / secs / 1e6
    threads = arm.threads
    wall = secs
    sample = (f"{n}x{n} grid" + (" (the full workload grid)" if (n, n) == (a.nx, a.ny) else f" crop of the {a.nx}x{a.ny} workload")
              + f", {inner} lattice steps per bench step, Float64")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * wall / max(a.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_keys(f"{a.lattice} {a.collision} Taylor-Green vortex decay, periodic (BASELINE configs[1])", "C2",
                              [a.nx, a.ny], a.gpus, a.inner, a.arith, "f64"),
        "note": "Julia is not installed; CPU arm = C restatement of the reference algorithm (oracle/lbm_oracle.c, -O2 "
                "-ffp-contract=off, OpenMP over rows); it steps a bounded sample of the config's workload: " + sample,
        "cpu_baseline": {"value": value, "unit": "MLUPS", "cores": threads, "nproc": os.cpu_count(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip().isdigit()]
        if device < len(ids):
            device = int(ids[device])  # nvidia-smi wants the physical index
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def build_case(a, world, lbm):
    """BASELINE.json configs as concrete problems (inputs: SURVEY.md section 8d)."""
    Qs = {"D2Q4": lbm.D2Q4, "D2Q5": lbm.D2Q5, "D2Q9": lbm.D2Q9, "D2Q13": lbm.D2Q13, "D2Q17": lbm.D2Q17,
          "D2Q21": lbm.D2Q21, "D2Q37": lbm.D2Q37}
    CMs = {"SRT": lbm.SRT, "TRT": lbm.TRT, "MRT": lbm.MRT}
    if a.config in ("C2", "C5w", "C5s"):
        # TGV(q, tau = 0.8, scale) Taylor-Green vortex decay, periodic; CollisionModel(TRT): Lambda = 1/4
        q = Qs[a.lattice]()
        if a.config == "C2":
            nx, ny, scaling = a.nx, a.ny * world, "weak"
        elif a.config == "C5w":
            nx, ny, scaling = 16384, 16384 * world, "weak"
        else:
            nx, ny, scaling = 32768 // a.also_shrink, 32768 // a.also_shrink, "strong"
        problem = lbm.TGV(q, 0.8, max(nx // 16, 1), nx, ny)
        cm = lbm.CollisionModel(CMs[a.collision], q, problem)
        name = f"{a.lattice} {a.collision} Taylor-Green vortex decay, periodic (BASELINE configs[{1 if a.config == 'C2' else 4}])"
        return dict(q=q, problem=problem, cm=cm, nx=nx, ny=ny, scaling=scaling, init=lbm.AnalyticalEquilibrium(),
                    workload=name)
    if a.config == "C3":
        # D2Q9 SRT + uniform force Poiseuille channel 1024 x 8192, bounce-back North + South (strong scaling)
        q = lbm.D2Q9()
        nx, ny = 1024, 8192 // a.also_shrink
        problem = lbm.PoiseuilleFlow.fields(1.0, 0.1 / (ny / 5), 1 / 6, nx, ny, 1.0, (1.0, 1.0), 1.0)
        cm = lbm.CollisionModel(lbm.SRT, q, problem)
        return dict(q=q, problem=problem, cm=cm, nx=nx, ny=ny, scaling="strong",
                    init=lbm.ZeroVelocityInitialCondition(),
                    workload="D2Q9 SRT+force Poiseuille, bounce-back walls, 1024x8192 (BASELINE configs[2])")
    # C4: D2Q37 TRT Couette 8192^2, bounce-back South + moving wall North, halo width 3 (strong scaling)
    q = lbm.D2Q37()
    nx = ny = 8192 // a.also_shrink
    problem = lbm.CouetteFlow.fields(1.0, 0.01 / (ny / 5), 0.3 / q.speed_of_sound_squared, nx, ny, (1.0, 1.0))
    cm = lbm.CollisionModel(lbm.TRT, q, problem)
    return dict(q=q, problem=problem, cm=cm, nx=nx, ny=ny, scaling="strong", init=lbm.ZeroVelocityInitialCondition(),
                workload="D2Q37 TRT Couette, moving wall North + bounce-back South, 8192x8192 (BASELINE configs[3])")


def parity_check(world, rank, local, comm, lbm, torch, dist):
    """OUTSIDE the timed region: multi-GPU parity made visible in the bench record (the GPU test box has one GPU).
    20 steps of D2Q9 TRT Taylor-Green on a 4096 x (128 N) domain split into N y-slabs, Float64 exact and Float32 fast,
    on both halo paths (peer-memory stores / NCCL send-recv); the slabs are gathered on rank 0 and compared with the C
    oracle (oracle/, the checker -- never on the measured path) stepping the same initial state."""
    nx, ny, nsteps = 4096, 128 * world, 20
    q = lbm.D2Q9()
    problem = lbm.TGV(q, 0.8, nx // 16, nx, ny)
    cm = lbm.CollisionModel(lbm.TRT, q, problem)
    cases, want, f0_full, f0_gathered = [], None, None, False
    for dtype, arith in (("f64", "exact"), ("f32", "fast")):
        for p2p in ((1, 0) if world > 1 else (1,)):
            ctx = lbm.model.make_context(q, cm, problem.boundary_conditions(), nx, ny, dtype, arith, comm, local)
            ctx.set_option("p2p", p2p)
            st = lbm.DeviceState(ctx, q, cm, comm)
            f0 = lbm.initialize(lbm.AnalyticalEquilibrium(), q, problem, rows=(ctx.y0, ctx.ny_local))
            ctx.upload_f(f0)
            st.step(0, nsteps, problem.delta_t())
            got = ctx.download_f()
            path = ctx.halo_path
            ctx.close()
            # gather [Q][nyl][nx] slabs on rank 0 (equal slabs: ny = 128 N)
            def gather(a):
                t = torch.from_numpy(np.ascontiguousarray(np.transpose(a, (2, 1, 0)))).cuda()
                if world == 1:
                    return t.cpu().numpy()
                parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
                dist.gather(t, parts, dst=0)
                return np.concatenate([x.cpu().numpy() for x in parts], axis=1) if rank == 0 else None
            got_full = gather(got)
            if not f0_gathered:  # a collective: every rank must take this branch the same number of times
                f0_full, f0_gathered = gather(f0), True
            if rank == 0:
                if want is None:
                    import oracle.lbm_oracle as O
                    from oracle.c_oracle import COracle, lib as oracle_lib
                    oracle_lib().oracle_set_threads(host_cores())
                    qo = O.L.D2Q9()
                    po = O.TGV(qo, 0.8, nx // 16, NX=nx, NY=ny)
                    want, _ = COracle(qo, O.collision_model("TRT", qo, po)).steps(f0_full, nsteps)
                err = float(np.max(np.abs(got_full - want)) / np.max(np.abs(want)))
                cases.append({"dtype": dtype, "arith": arith, "halo_path": path, "max_rel_err": err,
                              "bit_identical": bool(np.array_equal(got_full, want)),
                              "tolerance": 1e-12 if dtype == "f64" else 1e-5})
    if rank != 0:
        return None
    ok = all(c["max_rel_err"] < c["tolerance"] and (c["bit_identical"] or c["dtype"] != "f64") for c in cases)
    return {"grid": [nx, ny], "slabs": world, "steps": nsteps, "workload": "D2Q9 TRT Taylor-Green vortex vs the C oracle",
            "halo_paths": {"0": "single GPU", "1": "NCCL send/recv", "2": "peer-memory stores"}, "cases": cases, "ok": ok,
            "max_rel_err": max(c["max_rel_err"] for c in cases),
            "bit_identical": all(c["bit_identical"] for c in cases if c["dtype"] == "f64")}


ALSO_PRESETS = (("C5s", 10), ("C3", 100), ("C4", 20))  # preset, lattice steps per timed batch
ALSO_FILE = os.path.join(ROOT, ".bench_also_n1.json")  # N = 1 figures of the same scaling run (efficiency denominator)


def also_cases(a, world, rank, local, comm, lbm, torch, dist, peak):
    """OUTSIDE the headline's timed region: the north_star's STRONG-scaling configurations on the same N ranks -- C5s (D2Q9
    TRT TGV 32768^2), C3 (D2Q9 SRT + force Poiseuille 1024 x 8192, walls) and C4 (D2Q37 TRT Couette 8192^2, halo 3) -- a
    few device batches each, timed like the headline (CUDA events on the library's stream, max over ranks).  An N = 1 run
    leaves its figures in .bench_also_n1.json; later runs with N > 1 on the same box report their efficiency against
    them (value_N / (N value_1))."""
    import argparse
    n1 = {}
    if world > 1 and a.also_shrink == 1 and os.path.exists(ALSO_FILE) and time.time() - os.path.getmtime(ALSO_FILE) < 6 * 3600:
        try:
            n1 = json.load(open(ALSO_FILE))
        except Exception:
            n1 = {}
    out = []
    for preset, inner in ALSO_PRESETS:
        b = argparse.Namespace(**vars(a))
        b.config, b.dtype, b.arith = preset, "f64", a.arith
        entry = {"preset": preset, "scaling": "strong"}
        ctx = None
        try:
            case = build_case(b, world, lbm)
            q, problem, cm, nx, ny = case["q"], case["problem"], case["cm"], case["nx"], case["ny"]
            entry.update(workload=case["workload"], grid_global=[nx, ny], lattice_steps_per_batch=inner)
            need = 2 * nx * (ny // world + 8) * q.Q * 8
            free = torch.cuda.mem_get_info()[0]
            fits = need <= free - (2 << 30)
            if world > 1:  # lbm_create is collective: every rank takes the same decision
                tt = torch.tensor([1.0 if fits else 0.0], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MIN)
                fits = bool(tt.cpu()[0] > 0.5)
            if not fits:
                raise MemoryError(f"needs {need / 1e9:.1f} GB per GPU, {free / 1e9:.1f} GB free on rank {rank}")
            ctx = lbm.model.make_context(q, cm, problem.boundary_conditions(), nx, ny, "f64", a.arith, comm, local)
            ctx.set_option("p2p", a.p2p)
            ctx.set_option("graph", a.graph)
            ctx.set_option("persistent", a.persistent)
            if a.overlap >= 0:
                ctx.set_option("overlap", a.overlap)
            state = lbm.DeviceState(ctx, q, cm, comm)
            state.prepare_force(0, 1, problem.delta_t())
            t0 = time.perf_counter()
            if not lbm.initialize_on_device(case["init"], q, problem, ctx):
                raise RuntimeError("device-side initialisation unavailable")
            ctx.sync()
            entry["init_on_device_s"] = round(time.perf_counter() - t0, 4)
            t = 0
            for _ in range(3):
                ctx.step(t, inner, 1.0)
                t += inner
            ctx.sync()
            if world > 1:
                dist.barrier()
            nb = 5
            l0 = ctx.kernel_launches
            ctx.timer_start()
            for _ in range(nb):
                ctx.step(t, inner, 1.0)
                t += inner
            dev_ms = ctx.timer_stop()
            launches = ctx.kernel_launches - l0
            from lbm import _abi
            cons = ctx.reduce(_abi.REDUCE_CONSERVED)
            if world > 1:
                tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dev_ms = float(tt.cpu()[0])
            value = nx * ny * inner * nb / (dev_ms * 1e-3) / 1e6
            gbs = value * 1e6 * 2 * q.Q * 8 / 1e9 / world
            entry.update(value=value, unit="MLUPS", ms_per_lattice_step=dev_ms / (nb * inner), gbs_per_gpu=gbs,
                         frac_of_hbm_peak_per_gpu=gbs / peak, grid_per_gpu=[nx, ctx.ny_local], halo_path=ctx.halo_path,
                         launches_per_lattice_step=launches / (nb * inner), finite=bool(np.isfinite(cons).all()))
            if world == 1:
                n1[preset] = value
            elif preset in n1:
                entry.update(n1_value=n1[preset], efficiency_vs_n1=value / (world * n1[preset]))
        except Exception as e:  # a configuration that does not fit or fails is reported, it never takes the headline down
            entry["skipped"] = f"{type(e).__name__}: {e}"[:300]
        finally:
            if ctx is not None:
                ctx.close()
        if world > 1:
            dist.barrier()
        out.append(entry)
    if world == 1 and rank == 0 and n1 and a.also_shrink == 1:
        try:
            json.dump(n1, open(ALSO_FILE, "w"))
        except OSError:
            pass
    return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(a):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum), MEASURED for this run:
    after the timed region an `ncu` child process captures one fused launch of the same <lattice, model, dtype, arith, grid>
    through tools/profile_case.py.  -> (bytes or None, source).  Falls back to the committed capture in
    profiles/traffic.json when ncu cannot run (no binary / no permission to read the counters)."""
    import shutil
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if ncu and a.config == "C2" and not os.environ.get("LBM_BENCH_NO_NCU"):
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
               "-k", "regex:k_step", "-s", "4", "-c", "1", "--csv", sys.executable, os.path.join(ROOT, "tools", "profile_case.py"),
               "--lattice", a.lattice, "--model", a.collision, "--dtype", a.dtype, "--arith", a.arith, "--n", str(a.nx),
               "--ny", str(a.ny), "--steps", "4", "--variant", str(a.variant)]
        try:
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
                env.pop(k, None)
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=180, env=env).stdout
            total, seen = 0.0, 0
            for line in out.splitlines():
                cells = [c.strip('"') for c in line.split('","')]
                if len(cells) > 3 and cells[-3] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(cells[-1].strip('"').replace(",", ""))
                    seen += 1
            if seen == 2 and total > 0:
                return total, "ncu child process after the timed region (one fused launch, this run)"
        except Exception:
            pass
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p) and a.config == "C2":
        t = json.load(open(p))
        v = t.get(f"{a.lattice}_{a.collision}_{a.dtype}_{a.arith}_{a.nx}x{a.ny}")
        if v is not None:
            return v, "profiles/traffic.json (committed ncu capture; ncu unavailable in this run)"
    return None, "not measured"


def run_b200(a):
    import torch
    import torch.distributed as dist
    import lbm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = lbm.SlabComm()
    inner = a.inner
    case = build_case(a, world, lbm)
    q, problem, cm, nx, ny, scaling = case["q"], case["problem"], case["cm"], case["nx"], case["ny"], case["scaling"]
    ctx = lbm.model.make_context(q, cm, problem.boundary_conditions(), nx, ny, a.dtype, a.arith, comm, local)
    ctx.set_option("variant", a.variant)
    ctx.set_option("p2p", a.p2p)
    ctx.set_option("graph", a.graph)
    ctx.set_option("persistent", a.persistent)
    if a.overlap >= 0:
        ctx.set_option("overlap", a.overlap)
    halo_path = {0: "none (single GPU)", 1: "NCCL send/recv on a side stream", 2: "peer-memory stores from the boundary-row launch"}[ctx.halo_path]
    nyl = ctx.ny_local
    state = lbm.DeviceState(ctx, q, cm, comm)
    state.prepare_force(0, 1, problem.delta_t())
    # synthetic input: the problem's analytic initial condition for this rank's slab.  Small slabs live
    # in ONE pinned host array (also used by the e2e leg); huge ones are generated chunk by chunk.
    slab_bytes = nx * nyl * q.Q * 8
    do_e2e = (not a.no_e2e) and slab_bytes <= (8 << 30)
    host = None
    if do_e2e:
        pinned = torch.empty(nx * nyl * q.Q, dtype=torch.float64, pin_memory=True)
        host = pinned.numpy().reshape((nx, nyl, q.Q), order="F")
    if host is None:
        # device-side initialisation: host produces (rho, u, T) rows, the library evaluates the equilibrium
        assert lbm.initialize_on_device(case["init"], q, problem, ctx)
    else:
        chunk = max(1, min(nyl, (16 << 20) // max(nx, 1)))
        for off in range(0, nyl, chunk):
            n = min(chunk, nyl - off)
            host[:, off:off + n, :] = lbm.initialize(case["init"], q, problem, rows=(ctx.y0 + off, n))
        ctx.upload_f(host)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t = 0
    for _ in range(max(a.warmup, 3)):
        ctx.step(t, inner, 1.0)
        t += inner
    barrier()
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # EXACTLY K timed bench steps, CUDA events on the library's stream, barrier + sync on both sides
    wall0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(a.steps):
        ctx.step(t, inner, 1.0)
        t += inner
    dev_ms = ctx.timer_stop()
    barrier()
    region_wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    region_ms = region_wall * 1e3
    if world > 1:
        tt = torch.tensor([dev_ms, region_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, region_ms = [float(v) for v in tt.cpu()]
    updates = nx * ny * inner * a.steps
    value = updates / (dev_ms * 1e-3) / 1e6
    b_alg = 2 * q.Q * BYTES[a.dtype]
    kern_ms = dev_ms / (a.steps * inner)  # one fused kernel per lattice step (world == 1)
    peak, peak_src = measured_peak()
    achieved = b_alg * nx * nyl / (kern_ms * 1e-3) / 1e9

    # ---- e2e: host buffers in and out through the C ABI --------------------------------------
    e2e = None
    if do_e2e:
        out_pinned = torch.empty(nx * nyl * q.Q, dtype=torch.float64, pin_memory=True)
        out_host = out_pinned.numpy().reshape((nx, nyl, q.Q), order="F")
        n_e2e = max(2, min(a.steps, 5))
        ctx.upload_f(host); ctx.step(0, inner, 1.0); ctx.download_f(out_host)  # warm
        barrier()
        t0 = time.perf_counter()
        for k in range(n_e2e):
            ctx.upload_f(host)
            ctx.step(0, inner, 1.0)
            ctx.download_f(out_host)
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.cpu()[0])
        nbytes = nx * nyl * q.Q * 8
        serial = nx * ny * inner * n_e2e / e2e_s / 1e6
        e2e = {"value": serial, "unit": "MLUPS", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "bench_steps": n_e2e, "jobs_in_flight": 1,
               "note": "per bench step: lbm_upload_f(pinned host f) + lbm_step(inner) + lbm_download_f(host f)"}
        if world == 1 and not a.e2e_serial:
            # A stream of independent jobs, two in flight: job k + 1's upload and job k - 1's download (asynchronous forms
            # of the same calls, page-locked arrays) overlap job k's steps on a second context -- every job still moves its
            # input host -> device and its result device -> host inside the timed region.
            extra = []
            try:
                nfl = max(2, a.e2e_jobs)
                for _ in range(nfl - 1):
                    c2 = lbm.model.make_context(q, cm, problem.boundary_conditions(), nx, ny, a.dtype, a.arith, comm, local)
                    extra.append(c2)
                    for k_, v_ in (("variant", a.variant), ("graph", a.graph), ("persistent", a.persistent)):
                        c2.set_option(k_, v_)
                    lbm.DeviceState(c2, q, cm, comm).prepare_force(0, 1, problem.delta_t())
                ring = [ctx] + extra
                outs = [out_host] + [torch.empty(nx * nyl * q.Q, dtype=torch.float64, pin_memory=True).numpy().reshape((nx, nyl, q.Q), order="F")
                                     for _ in extra]
                n_jobs = max(4 * nfl, 2 * n_e2e)
                for k in range(nfl):  # warm every context (graphs, allocations)
                    ring[k].upload_f_async(host); ring[k].step(0, inner, 1.0); ring[k].download_f_async(outs[k])
                for c_ in ring:
                    c_.sync()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(n_jobs):
                    c_ = ring[k % nfl]
                    c_.upload_f_async(host)
                    c_.step(0, inner, 1.0)
                    c_.download_f_async(outs[k % nfl])
                for c_ in ring:
                    c_.sync()
                torch.cuda.synchronize()
                piped_s = time.perf_counter() - t0
                for o_ in outs[1:]:  # same input, same result from every context
                    assert np.array_equal(out_host, o_) and np.isfinite(o_).all()
                e2e.update(value=nx * ny * inner * n_jobs / piped_s / 1e6, bench_steps=n_jobs, jobs_in_flight=nfl,
                           serial_value=serial,
                           note=f"stream of independent jobs, {nfl} in flight on {nfl} contexts: per job lbm_upload_f_async(pinned "
                                "host f) + lbm_step(inner) + lbm_download_f_async(host f); the uploads and downloads of the "
                                "other jobs overlap one job's steps (PCIe full duplex; one context serialises download -> "
                                "upload -> steps on its buffers, so three keep the SMs busy).  serial_value = one job at a "
                                "time (upload, steps, download back to back)")
            except Exception as ex:
                e2e["pipelined_error"] = f"{type(ex).__name__}: {ex}"[:200]
            finally:
                for c2 in extra:
                    c2.close()
        assert np.isfinite(out_host).all()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        v, threads, secs = cpu_restatement_mlups(a.cpu_n, a.cpu_steps, q.name, type(cm).__name__)
        cpu = {"value": v, "unit": "MLUPS", "cores": threads, "kind": "port",
               "sample": f"{a.cpu_n}x{a.cpu_n} grid, {a.cpu_steps} lattice steps ({secs:.1f} s), C restatement "
                         f"(oracle/lbm_oracle.c) with OpenMP over rows"}
    ctx.close()
    traffic, traffic_src = ncu_traffic(a) if rank == 0 else (None, None)
    parity = None if a.no_parity else parity_check(world, rank, local, comm, lbm, torch, dist)
    headline = a.config == "C2" and ((a.nx, a.ny) == (4096, 4096) or a.also_shrink > 1)
    also = None if (a.no_also or not headline) else also_cases(a, world, rank, local, comm, lbm, torch, dist, peak)
    if scaling == "weak":
        cfg = config_keys(case["workload"], a.config, [nx, ny // world], world, inner, a.arith, a.dtype, q.Q, "weak")
    else:
        cfg = config_keys(case["workload"], a.config, [nx, ny], world, inner, a.arith, a.dtype, q.Q, "strong")
        cfg["grid_per_gpu"] = [nx, nyl]
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": cfg,
            "details": {"grid_local": [nx, nyl], "variant": a.variant, "cuda_graphs": bool(a.graph), "persistent": a.persistent, "overlap": a.overlap, "halo_exchange": halo_path,
                        "wall_ms_per_step": region_ms / a.steps},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_update": b_alg,
                         "kernel_ms": kern_ms, "kernel": "k_step<%s, %s, pull>" % (type(cm).__name__, a.dtype),
                         "note": "peak is the measured COPY bandwidth (MEASURED_PEAKS.json); a fused read+write kernel with L2 "
                                 "prefetch can exceed it (nominal HBM3e: 7.7-8 TB/s)" if achieved > peak else None},
            "cpu_baseline": cpu,
            "parity_check": parity,
            "also": also,
        }))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner there) get
    # stderr for the duration, the result line goes to the real stdout.
    real = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")
    try:
        return run_reference(a) if a.impl == "reference" else run_b200(a)
    finally:
        sys.stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
