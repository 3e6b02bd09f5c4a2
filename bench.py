#!/usr/bin/env python
"""bench.py - BASELINE.json metric: batched iLQR-ADMM solves/sec (car, N=100, 65,536 problems per GPU).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the numpy port of the reference, all host cores)

A "step" is one full solve of the batch: I_o=20 outer iLQR iterations x [linearise + Riccati K-pass + I_a=5 ADMM
iterations x (feed-forward pass + linear rollout, L=20-candidate nonlinear line search, winner rollout + projection /
dual update)], fixed budget (every stop test disabled, so the work is deterministic).  Prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ilqr-admm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "batched iLQR-ADMM solves/sec"
UNIT = "solves/s"

# algorithmic FP64 work of one line-search candidate-step of the car (DESIGN.md "Kernels"): u = u^ + a*du (2 FMA),
# control cost (2 FMA), ADMM penalty (2 x [sub, mul, FMA]), model (dv, 4 FMA-type updates = 9 flop), sincos counted
# as 40 flop (3-term Cody-Waite reduction + two 7-term Horner polynomials + reconstruction)
FLOP_PER_CAND_STEP_CAR = 4 + 4 + 8 + 9 + 40


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU")
    ap.add_argument("--early-exit", action="store_true", help="reference stop rules instead of the fixed budget")
    ap.add_argument("--notebook-budget", action="store_true",
                    help="I_o=30, I_a=5, L=50 (Car/Iterative LQR with control constraints.ipynb cell 20) instead of the "
                         "reference defaults I_o=20, L=20 (secondary figure of SURVEY 8d; not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work per reference step")
    return ap.parse_args()


BUDGET = dict(I_o=20, I_a=5, L=20)


def workload(B, seed=None):
    from isls_b200 import configs
    kw = {} if seed is None else {"seed": seed}
    return configs.car_batch(B, tol=1e-3, **BUDGET, **kw)


def config_dict(p, B, n_gpus, fixed):
    return {"workload": "C5 2D-car iLQR-ADMM (x_dim=4,u_dim=2,N=100, |u|<=0.5, rho_u=10), %d problems per GPU, "
                        "I_o=%d x I_a=%d x L=%d, %s" % (B, p["I_o"], p["I_a"], p["L"],
                                                         "fixed budget" if fixed else "reference stop rules"),
            "problems_per_gpu": B, "global_problems": B * n_gpus, "N": p["N"], "x_dim": 4, "u_dim": 2,
            "outer_iters": p["I_o"], "admm_iters": p["I_a"], "linesearch_candidates": p["L"],
            "fixed_budget": fixed, "parallelism": "problem-sharded x%d, no solve-path collectives" % n_gpus,
            "l2_policy": "workspace (3.2 GB per 65,536 problems) >> 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """Solve a few problems one at a time with the numpy port (B=1 calls = the reference looped over problems)."""
    seed_idx, count, fixed = args
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    from oracle import restated as R
    from isls_b200 import configs
    p = workload(max(seed_idx + count, 1))
    t0 = time.perf_counter()
    for i in range(count):
        R.ilqr_admm(configs.subset(p, [seed_idx + i]), fixed_budget=fixed)
    return time.perf_counter() - t0


def cpu_reference_step(per_core, cores, fixed, pool):
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(c * per_core, per_core, fixed) for c in range(cores)])
    return time.perf_counter() - t0, per_core * cores


def run_reference(a):
    """--impl reference: the reference's algorithm on the host cores (oracle port, kind='port': the reference itself
    is Python and does not exist on the GPU box), multiprocessing over all cores, one BLAS thread per worker."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    fixed = not a.early_exit
    cores = os.cpu_count() or 1
    t1 = _cpu_worker((0, 1, fixed))                       # calibration: one solve on one core
    per_core = max(1, int(a.cpu_seconds / max(t1, 1e-3)))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(a.warmup):
            cpu_reference_step(1, cores, fixed, pool)
        times, solved = [], 0
        for _ in range(a.steps):
            dt, cnt = cpu_reference_step(per_core, cores, fixed, pool)
            times.append(dt)
            solved += cnt
    total = sum(times)
    value = solved / total
    p = workload(1)
    sample = "%d problems per step (%d per core x %d cores), %d steps, one numpy solve per problem" % (
        per_core * cores, per_core, cores, a.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(p, a.batch, a.gpus, fixed),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "single_core_solve_s": t1, "riccati_pass_ms_one_core": round(cpu_riccati_pass_ms(), 3)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = "/tmp/isls_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


_OUT = sys.stdout


def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from isls_b200 import _lib, solver as S
    _lib.lib()                                             # fails loudly when the CUDA library is missing
    B = a.batch
    fixed = not a.early_exit
    p = workload(B)
    # every rank solves its own B problems (weak scaling): different seeds per rank
    if world > 1:
        from isls_b200 import configs
        p = workload(B, seed=1234 + 2 + 1000 * rank)
    plan = S.Plan("car", p["N"], 4, 2, p["dt"], p["Qdiag"], p["seq"], p["u_std"], p["L"], rho_u=p["rho_u"],
                  lo_u=p["lo_u"], hi_u=p["hi_u"])
    sv = S.BatchSolver(plan, B, dev, max_outer=p["I_o"], max_admm=p["I_a"], logs=False)
    launches_per_step = 2 + p["I_o"] * (2 + 2 * p["I_a"])     # init, finalize; per outer: kpass, outer_end, I_a x (ff, ls)

    # host buffers (pinned) for the end-to-end arm
    h_x0 = torch.from_numpy(p["x0"]).pin_memory()
    h_u0 = torch.from_numpy(p["u0"]).pin_memory()
    h_zs = torch.from_numpy(p["zs"]).pin_memory()
    h_out = {k: torch.empty(sv.out[k].shape, dtype=sv.out[k].dtype).pin_memory()
             for k in ("x", "u", "cost", "status", "cost_log")}
    from isls_b200.sharding import gather_scalars

    def solve():
        out = sv.ilqr_admm(tol=p["tol"], fixed_budget=fixed)
        if world > 1:                                       # NCCL only gathers per-problem result scalars
            gather_scalars(out.cost, B * world)
        return out

    def step_e2e():
        sv.h2d_bytes = 0
        sv.set_inputs(h_x0, h_u0, h_zs)
        out = solve()
        for k, h in h_out.items():
            h.copy_(out[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm
    sv.set_inputs(h_x0, h_u0, h_zs)
    for _ in range(max(a.warmup, 1)):
        solve()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(solve, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * a.steps / (ms * 1e-3)

    # ---- end-to-end arm: pinned host inputs -> device -> solve -> pinned host results, every step
    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    e2e_value = B * world * a.steps / (ms_e2e * 1e-3)
    d2h = sum(h.numel() * h.element_size() for h in h_out.values())

    # ---- per-kernel CUDA-event timing (separate pass, rank 0) and roofline of the dominant kernel
    roof = kernels = None
    if rank == 0:
        fp64_peak = S.measure_fp64_tflops(dev)
        S.profile_enable(True)
        sv.ilqr_admm(tol=p["tol"], fixed_budget=fixed)      # rank-0-only pass: no collective in here
        prof = S.profile_collect()
        S.profile_enable(False)
        tot = sum(v[0] for v in prof.values())
        kernels = {k: {"ms_total": round(v[0], 4), "launches": v[1], "share": round(v[0] / tot, 4),
                       "ms_per_launch": round(v[0] / v[1], 5)} for k, v in prof.items()}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        ls_ms, ls_n = prof["linesearch"]
        flops = float(B) * p["L"] * p["N"] * FLOP_PER_CAND_STEP_CAR
        ach = flops / (ls_ms / ls_n * 1e-3) / 1e12
        fused_update = "admm" not in prof          # the streaming ADMM z/lambda update runs as the kernel's epilogue
        # bytes the kernel has to move per launch: u^, du read once per problem (2 x N x m), x^_0 (n), the control-cost
        # polynomial (3), best index / best cost written (2)
        alg_bytes = float(B) * (2 * p["N"] * 2 + 4 + 6 + 2) * 8
        # two-phase bound of the fused kernel: FP64 phase at the DFMA peak + streaming ADMM epilogue (z, lambda read;
        # z, lambda, reg written = 5 doubles per control element; u^, du are re-read from cache) at the measured HBM peak
        epi_bytes = float(B) * p["N"] * 2 * 5 * 8 if fused_update else 0.0
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        t_bound = flops / (fp64_peak * 1e12) + epi_bytes / (hbm_peak * 1e9)
        shape = "5,4,3" if p["L"] <= 20 else "5,10,1"            # csrc/isls_b200.cu: launch_linesearch
        roof = {"kernel": "k_linesearch<CarModel,%s>" % shape + (" + fused ADMM z/lambda epilogue" if fused_update else ""),
                "bound": "fp64", "achieved": round(ach, 3),
                "peak": round(fp64_peak, 3), "unit": "TFLOP/s", "frac": round(ach / fp64_peak, 4),
                "peak_source": "DFMA throughput measured live by isls_measure_fp64_tflops (MEASURED_PEAKS.json has "
                               "no FP64 figure)",
                "algorithmic_flop_per_launch": flops, "flop_per_candidate_step": FLOP_PER_CAND_STEP_CAR,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes + epi_bytes,
                        "achieved_gbs": round((alg_bytes + epi_bytes) / (ls_ms / ls_n * 1e-3) / 1e9, 2),
                        "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at B=65,536 from the committed
                # `ncu --set full` capture (profiles/r1_ncu_full_main_kernels.csv), scaled to this batch size
                "traffic": round(780.1e6 * B / 65536.0), "traffic_source": "profiles/r1_ncu_linesearch_staged.csv",
                "two_phase": {"epilogue_bytes_per_launch": epi_bytes, "bound_ms": round(t_bound * 1e3, 4),
                              "measured_ms": round(ls_ms / ls_n, 4),
                              "frac": round(t_bound / (ls_ms / ls_n * 1e-3), 4)}}

    if rank == 0 and roof is not None and "ff" in prof:
        # second-largest kernel (HBM-bound): ff-pass + linear rollout, 44 doubles per problem-step
        ff_ms, ff_n = prof["ff"]
        ff_bytes = float(B) * p["N"] * 44 * 8
        roof["second_kernel"] = {"kernel": "k_ff<CarModel>", "bound": "hbm",
                                 "achieved": round(ff_bytes / (ff_ms / ff_n * 1e-3) / 1e9, 1), "peak": hbm_peak,
                                 "unit": "GB/s", "frac": round(ff_bytes / (ff_ms / ff_n * 1e-3) / 1e9 / hbm_peak, 4),
                                 "algorithmic_bytes_per_launch": ff_bytes, "traffic": round(2171.1e6 * B / 65536.0)}
    if rank == 0 and roof is not None and "kpass" in prof:
        # the Riccati K-pass named by BASELINE.json's metric: fused-model variant (Jacobians recomputed in-kernel), per
        # problem-step x^, u^ in (6 doubles; the car never loads x, y) and K, Qux, packed Quu, Quu^-1 out (22 doubles)
        # + the ADMM state reset (z_u read; lambda_u, reg_u written: 6 doubles) = 34 doubles; 723 flop.
        kp_ms, kp_n = prof["kpass"]
        kp_bytes = float(B) * p["N"] * 34 * 8
        kp_flops = float(B) * (p["N"] - 1) * 723.0
        t = kp_ms / kp_n * 1e-3
        roof["riccati_kernel"] = {"kernel": "k_kpass<CarModel>", "bound": "hbm",
                                  "achieved": round(kp_bytes / t / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                                  "frac": round(kp_bytes / t / 1e9 / hbm_peak, 4),
                                  "algorithmic_bytes_per_launch": kp_bytes,
                                  "fp64": {"algorithmic_flop_per_launch": kp_flops,
                                           "achieved_tflops": round(kp_flops / t / 1e12, 3),
                                           "frac_of_dfma_peak": round(kp_flops / t / 1e12 / fp64_peak, 4)},
                                  "passes_per_s": round(B / t, 1),
                                  # dram bytes read + written of one launch, profiles/r1_ncu_kpass_ff.csv
                                  "traffic": round(1671.1e6 * B / 65536.0)}
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                "--warmup", "0", "--cpu-seconds", str(a.cpu_seconds)] +
                               (["--early-exit"] if a.early_exit else []) +
                               (["--notebook-budget"] if a.notebook_budget else []), capture_output=True, text=True,
                               timeout=600)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:                                # the baseline is a report, never a blocker
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % e}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(p, B, world, fixed), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                        "h2d_bytes_per_step": sv.h2d_bytes, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches_per_step * a.steps, "roofline": roof, "kernels": kernels,
                "cpu_baseline": cpu}
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_riccati_pass_ms(reps=5):
    """One backward_pass_DP of the reference's algorithm (isls.py:229-308; oracle port) on one car problem, ms."""
    import numpy as np
    from oracle import restated as R
    p = workload(1)
    model = R._model_of(p)
    x, u = R.initial_rollout(p)
    A, Bm = model.get_AB(x, u)
    N, n, m = p["N"], p["n"], p["m"]
    Cxx = R._diag_embed(2.0 * (p["Qdiag"][p["seq"]] + 1.0))[None]
    Cuu = R._diag_embed(np.full((N, m), 2.0 * (p["u_std"] + 10.0)))[None]
    z = np.zeros((1, N, n)), np.zeros((1, N, m))
    R.backward_pass(A, Bm, z[0], z[1], Cxx, Cuu)
    t0 = time.perf_counter()
    for _ in range(reps):
        R.backward_pass(A, Bm, z[0], z[1], Cxx, Cuu)
    return (time.perf_counter() - t0) / reps * 1e3


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout
    when NCCL_DEBUG is set, as it is on the GPU boxes), so file descriptor 1 is pointed at stderr for the whole run and
    the JSON line goes to a private duplicate of the original stdout."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


if __name__ == "__main__":
    args = parse()
    _claim_stdout()
    if args.notebook_budget:
        BUDGET.update(I_o=30, I_a=5, L=50)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
