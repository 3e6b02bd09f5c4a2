#!/usr/bin/env python
"""bench.py - BASELINE.json metric: batched iLQR-ADMM solves/sec (configs[4]: 2D car, N=100, 65,536 problems, sharded
over 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the unmodified reference from baseline/_ref,
                                                              all host cores; the numpy port beside it)

A "step" is one full solve of the batch: I_o=20 outer iLQR iterations x [linearise + Riccati K-pass + I_a=5 ADMM
iterations x (feed-forward pass + linear rollout, L=20-candidate nonlinear line search, winner rollout + projection /
dual update)], fixed budget (every stop test disabled, so the work is deterministic).

Scaling: the default is the sweep BASELINE.json names - STRONG scaling, 65,536 problems in total, contiguous shards of
65,536 / N problems per GPU (`--scaling strong`).  On N > 1 GPUs the weak-scaling figure (65,536 problems per GPU) is
measured in the same run and reported under "weak_scaling".  Prints ONE JSON line.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ilqr-admm_b200")
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "batched iLQR-ADMM solves/sec"
UNIT = "solves/s"

# algorithmic FP64 work of one line-search candidate-step of the car (DESIGN.md "Kernels"): u = u^ + a*du (2 FMA),
# control cost (2 FMA), ADMM penalty (2 x [sub, mul, FMA]), model (dv, 4 FMA-type updates = 9 flop), sincos counted
# as 40 flop (3-term Cody-Waite reduction + two 7-term Horner polynomials + reconstruction)
FLOP_PER_CAND_STEP_CAR = 4 + 4 + 8 + 9 + 40
# FP64-pipe instructions the kernel actually EXECUTES per candidate-step (SASS count, profiles/r1_linesearch_sass_mix.txt:
# 23 DFMA + 3 DMUL + 2 DADD; the control cost / penalty are folded into a per-problem quadratic by k_ff)
EXEC_FLOP_PER_CAND_STEP_CAR = 23 * 2 + 3 + 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --batch problems in total, sharded over the GPUs (BASELINE configs[4]); "
                         "weak: --batch problems per GPU")
    ap.add_argument("--batch", type=int, default=65536, help="problems (total for strong scaling, per GPU for weak)")
    ap.add_argument("--early-exit", action="store_true", help="reference stop rules instead of the fixed budget")
    ap.add_argument("--notebook-budget", action="store_true",
                    help="I_o=30, I_a=5, L=50 (Car/Iterative LQR with control constraints.ipynb cell 20) instead of the "
                         "reference defaults I_o=20, L=20 (secondary figure of SURVEY 8d; not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the secondary weak-scaling measurement (N > 1)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work per reference step")
    ap.add_argument("--cpu-total-seconds", type=float, default=150.0,
                    help="bound on the timed CPU work of the whole reference run (per-step work shrinks with --steps)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"],
                    help="CPU arm: the unmodified reference in baseline/_ref (default when present) or the numpy port")
    return ap.parse_args()


BUDGET = dict(I_o=20, I_a=5, L=20)


def workload(B, seed=None):
    from isls_b200 import configs
    kw = {} if seed is None else {"seed": seed}
    return configs.car_batch(B, tol=1e-3, **BUDGET, **kw)


def config_dict(p, B_total, B_local, n_gpus, fixed, scaling):
    return {"workload": "C5 2D-car iLQR-ADMM (x_dim=4,u_dim=2,N=100, |u|<=0.5, rho_u=10), %d problems in total = %d per "
                        "GPU x %d (%s scaling), I_o=%d x I_a=%d x L=%d, %s" % (
                            B_total, B_local, n_gpus, scaling, p["I_o"], p["I_a"], p["L"],
                            "fixed budget" if fixed else "reference stop rules"),
            "problems_per_gpu": B_local, "global_problems": B_total, "N": p["N"], "x_dim": 4, "u_dim": 2,
            "outer_iters": p["I_o"], "admm_iters": p["I_a"], "linesearch_candidates": p["L"],
            "fixed_budget": fixed, "parallelism": "problem-sharded x%d, no solve-path collectives" % n_gpus,
            "l2_policy": "workspace (3.2 GB per 65,536 problems) >> 126 MB L2 at >= 8,192 problems per GPU; every step "
                         "re-initialises and streams the whole workspace, nothing is reused across steps"}


# ------------------------------------------------------------------------------------------------- CPU arm
def _single_thread_blas():
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"


def _port_worker(args):
    """Solve a few problems one at a time with the numpy port (B=1 calls = the reference looped over problems)."""
    seed_idx, count, fixed = args
    _single_thread_blas()
    from oracle import restated as R
    from isls_b200 import configs
    p = workload(max(seed_idx + count, 1))
    t0 = time.perf_counter()
    for i in range(count):
        R.ilqr_admm(configs.subset(p, [seed_idx + i]), fixed_budget=fixed)
    return time.perf_counter() - t0, count * p["I_o"]


def ref_available():
    return os.path.isdir(os.path.join(REF_DIR, "isls"))


def _ref_worker(args):
    """The UNMODIFIED reference (baseline/_ref/isls = /root/reference/isls copied by build()) driven through
    oracle/ref_shim.py (external shims S0-S4 only: matplotlib stub, C/D aliases, ADMM(threshold=), log=True, quadratic
    cost closure), one iSLS.ilqr_admm call per problem.  The reference has no fixed-budget mode: its own stop rules run
    (isls/isls.py:493-499, admm.py:72-85), i.e. it does at most the budgeted work."""
    seed_idx, count, _fixed = args
    _single_thread_blas()
    os.environ["ISLS_REFERENCE"] = REF_DIR
    import numpy as np
    from oracle import models as M, ref_shim as S
    p = workload(max(seed_idx + count, 1))
    model = M.make_model("car", dt=p["dt"])
    Qs = np.stack([np.diag(q) for q in p["Qdiag"]])
    lo, hi = p["lo_u"].flatten(), p["hi_u"].flatten()
    t0 = time.perf_counter()
    outer = 0
    for i in range(count):
        b = seed_idx + i
        s = S.make_isls(model, p["N"], p["zs"][b] if p["zs"].ndim == 3 else p["zs"], Qs, p["seq"], p["u_std"])
        S.init_nominal(s, p["x0"][b], p["u0"] if p["u0"].ndim == 2 else p["u0"][b])
        r = S.run_ilqr_admm(s, model, project_u=lambda z: np.clip(z, lo, hi), rho_u=float(p["rho_u"][0, 0]),
                            max_iter=p["I_o"], max_admm_iter=p["I_a"], max_line_search_iter=p["L"], tol=p["tol"])
        outer += len(r["cost_log"]) - 1
    return time.perf_counter() - t0, outer


def cpu_arm(kind, cpu_seconds, steps, warmup, fixed):
    """Times `kind` ('reference' | 'port') over all host cores: (solves/s, ms per step, description dict)."""
    import multiprocessing as mp
    _single_thread_blas()
    worker = _ref_worker if kind == "reference" else _port_worker
    cores = os.cpu_count() or 1
    t1, _ = worker((0, 1, fixed))                          # calibration: one solve on one core
    per_core = max(1, int(cpu_seconds / max(t1, 1e-3)))
    ctx = mp.get_context("fork")
    times, solved, outer = [], 0, 0
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(worker, [(c, 1, fixed) for c in range(cores)])
        for _ in range(steps):
            t0 = time.perf_counter()
            res = pool.map(worker, [(c * per_core, per_core, fixed) for c in range(cores)])
            times.append(time.perf_counter() - t0)
            solved += per_core * cores
            outer += sum(r[1] for r in res)
    total = sum(times)
    what = ("unmodified reference (baseline/_ref/isls through oracle/ref_shim.py), iSLS.ilqr_admm per problem, reference "
            "stop rules (mean %.1f of %d outer iterations)" % (outer / solved, BUDGET["I_o"])) if kind == "reference" \
        else "numpy port of the reference's algorithm (oracle/restated.py), one solve per problem, %s" % (
            "fixed budget" if fixed else "reference stop rules")
    desc = {"value": solved / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d problems per step (%d per core x %d cores), %d step(s); %s" % (
                per_core * cores, per_core, cores, steps, what),
            "single_core_solve_s": t1}
    return solved / total, 1e3 * total / steps, desc


def run_reference(a):
    """--impl reference: the reference's own CPU implementation on the host cores - the unmodified package from
    baseline/_ref when it is there (kind 'reference'), else the numpy port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fixed = not a.early_exit
    kind = a.ref_kind
    if kind == "auto":
        kind = "reference" if ref_available() else "port"
    if kind == "reference" and not ref_available():
        raise SystemExit("baseline/_ref/isls is missing: run __graft_entry__.build() where /root/reference exists")
    per_step = min(a.cpu_seconds, a.cpu_total_seconds / max(a.steps, 1))     # the whole run ends within a few minutes
    value, ms, desc = cpu_arm(kind, per_step, a.steps, a.warmup, fixed)
    desc["riccati_pass_ms_one_core"] = round(cpu_riccati_pass_ms(), 3)
    if kind == "reference":                                 # the port beside it, one short step
        try:
            pv, _, pdesc = cpu_arm("port", min(a.cpu_seconds, 8.0), 1, 0, fixed)
            desc["port"] = {"value": pv, "sample": pdesc["sample"]}
        except Exception as e:                              # a report, never a blocker
            desc["port"] = {"value": None, "sample": "failed: %r" % e}
    p = workload(1)
    B_local = a.batch // a.gpus if a.scaling == "strong" else a.batch
    B_total = a.batch if a.scaling == "strong" else a.batch * a.gpus
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(p, B_total, B_local, a.gpus, fixed, a.scaling), "cpu_baseline": desc,
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


class Arm:
    """One measured configuration on this rank: plan + solver for B_local problems, pinned host buffers, and the
    device-resident and end-to-end timing loops."""

    def __init__(self, p, B_local, dev, world, fixed, B_total):
        import torch
        from isls_b200 import solver as S
        self.torch, self.p, self.B, self.dev, self.world, self.fixed, self.B_total = torch, p, B_local, dev, world, fixed, B_total
        self.plan = S.Plan("car", p["N"], 4, 2, p["dt"], p["Qdiag"], p["seq"], p["u_std"], p["L"], rho_u=p["rho_u"],
                           lo_u=p["lo_u"], hi_u=p["hi_u"], device=dev)
        self.sv = S.BatchSolver(self.plan, B_local, dev, max_outer=p["I_o"], max_admm=p["I_a"], logs=False)
        self.sv.add_output_set()
        # host buffers (pinned) for the end-to-end arm: inputs, and two result sets (double-buffered D2H)
        self.h_in = [torch.from_numpy(p[k]).pin_memory() for k in ("x0", "u0", "zs")]
        keys = ("x", "u", "cost", "status", "cost_log")
        self.h_out = [{k: torch.empty(self.sv.out[k].shape, dtype=self.sv.out[k].dtype).pin_memory() for k in keys}
                      for _ in range(2)]
        self.d2h = sum(h.numel() * h.element_size() for h in self.h_out[0].values())
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.solved = [torch.cuda.Event() for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        # gather buffer of the per-problem result scalars (NCCL all_gather_into_tensor; ragged shards padded)
        self.pad = -(-B_total // world)
        if world > 1:
            self.g_in = torch.zeros(self.pad, dtype=torch.float64, device=dev)
            self.g_out = torch.empty(self.pad * world, dtype=torch.float64, device=dev)
        self.i = 0

    def solve(self):
        out = self.sv.ilqr_admm(tol=self.p["tol"], fixed_budget=self.fixed)
        if self.world > 1:                                   # NCCL only gathers per-problem result scalars
            import torch.distributed as dist
            self.g_in[:self.B].copy_(out.cost)
            dist.all_gather_into_tensor(self.g_out, self.g_in)
        return out

    def step_e2e(self):
        """Host inputs -> device, solve, results -> host.  The D2H of step i runs on a copy stream while step i+1
        solves into the other result set; every copy is issued and completed inside the timed region."""
        torch = self.torch
        s = self.i % 2
        self.i += 1
        main = torch.cuda.current_stream(self.dev)
        self.sv.use_output_set(s)
        main.wait_event(self.copied[s])                      # set s is free again (its previous D2H is done)
        self.sv.h2d_bytes = 0
        self.sv.set_inputs(*self.h_in)
        out = self.solve()
        self.solved[s].record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.solved[s])
            for k, h in self.h_out[s].items():
                h.copy_(out[k], non_blocking=True)
            self.copied[s].record(self.copy_stream)

    def drain(self):
        main = self.torch.cuda.current_stream(self.dev)
        for e in self.copied:
            main.wait_event(e)


def run_b200(a):
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
    from isls_b200 import _lib, configs, solver as S
    from isls_b200.sharding import shard_range
    _lib.lib()                                             # fails loudly when the CUDA library is missing
    fixed = not a.early_exit

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        if after:
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(arm, sample_clocks):
        arm.sv.set_inputs(*arm.h_in)
        for _ in range(max(a.warmup, 1)):
            arm.solve()
        sampler = ClockSampler(local)
        if sample_clocks:
            sampler.start()
        ms = timed(arm.solve, a.steps)
        clocks = sampler.stop() if sample_clocks else None
        for _ in range(2):                                  # warm both result sets
            arm.step_e2e()
        arm.drain()
        ms_e2e = timed(arm.step_e2e, a.steps, after=arm.drain)
        return ms, ms_e2e, clocks

    # ---- headline: the BASELINE-named sweep (strong: --batch problems in total; weak: per GPU)
    if a.scaling == "strong":
        B_total = a.batch
        lo, hi = shard_range(B_total, rank, world)
        p = configs.subset(workload(B_total), range(lo, hi)) if world > 1 else workload(B_total)
        B_local = hi - lo
    else:
        B_local, B_total = a.batch, a.batch * world
        p = workload(B_local, seed=None if world == 1 else 1234 + 2 + 1000 * rank)
    arm = Arm(p, B_local, dev, world, fixed, B_total)
    ms, ms_e2e, clocks = measure(arm, rank == 0)
    value = B_total * a.steps / (ms * 1e-3)
    e2e_value = B_total * a.steps / (ms_e2e * 1e-3)
    launches_per_step = 2 + p["I_o"] * (2 + 2 * p["I_a"])     # init, finalize; per outer: kpass, outer_end, I_a x (ff, ls)

    # ---- secondary: weak scaling (65,536 problems per GPU) in the same run
    weak = None
    if world > 1 and a.scaling == "strong" and not a.no_weak:
        pw = workload(a.batch, seed=1234 + 2 + 1000 * rank)
        del arm.sv.ws                                       # the two workspaces need not coexist
        arm_w = Arm(pw, a.batch, dev, world, fixed, a.batch * world)
        ms_w, ms_w_e2e, _ = measure(arm_w, False)
        weak = {"scaling": "weak", "problems_per_gpu": a.batch, "global_problems": a.batch * world,
                "value": a.batch * world * a.steps / (ms_w * 1e-3), "ms_per_step": ms_w / a.steps,
                "e2e_value": a.batch * world * a.steps / (ms_w_e2e * 1e-3), "unit": UNIT}
        del arm_w

    # ---- per-kernel CUDA-event timing (separate pass, rank 0, at the headline single-GPU size) and rooflines
    roof = kernels = None
    if rank == 0:
        Bp = B_local if world == 1 else a.batch
        if world == 1:
            arm_p = arm
        else:
            arm_p = Arm(workload(Bp), Bp, dev, 1, fixed, Bp)
            arm_p.sv.set_inputs(*arm_p.h_in)
        fp64_peak = S.measure_fp64_tflops(dev)
        S.profile_enable(True)
        arm_p.sv.ilqr_admm(tol=p["tol"], fixed_budget=fixed)   # rank-0-only pass: no collective in here
        prof = S.profile_collect()
        S.profile_enable(False)
        tot = sum(v[0] for v in prof.values())
        kernels = {k: {"ms_total": round(v[0], 4), "launches": v[1], "share": round(v[0] / tot, 4),
                       "ms_per_launch": round(v[0] / v[1], 5)} for k, v in prof.items()}
        roof = rooflines(prof, Bp, p, fp64_peak)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                "--warmup", "0", "--cpu-seconds", str(a.cpu_seconds), "--ref-kind", a.ref_kind] +
                               (["--early-exit"] if a.early_exit else []) +
                               (["--notebook-budget"] if a.notebook_budget else []), capture_output=True, text=True,
                               timeout=900)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:                                # the baseline is a report, never a blocker
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % e}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(p, B_total, B_local, world, fixed, a.scaling), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                        "h2d_bytes_per_step": arm.sv.h2d_bytes, "d2h_bytes_per_step": arm.d2h,
                        "pipelining": "the D2H of step i runs on a copy stream while step i+1 solves into the second "
                                      "result set; all copies start and finish inside the timed region"},
                "gpu_launches": launches_per_step * a.steps, "roofline": roof, "kernels": kernels,
                "cpu_baseline": cpu}
        if weak is not None:
            line["weak_scaling"] = weak
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def rooflines(prof, B, p, fp64_peak):
    """Roofline objects of the dominant kernels from the per-kernel CUDA-event pass at B problems on one GPU."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ls_ms, ls_n = prof["linesearch"]
    t_ls = ls_ms / ls_n * 1e-3
    flops = float(B) * p["L"] * p["N"] * FLOP_PER_CAND_STEP_CAR
    exec_flops = float(B) * p["L"] * p["N"] * EXEC_FLOP_PER_CAND_STEP_CAR
    ach = flops / t_ls / 1e12
    fused_update = "admm" not in prof          # the streaming ADMM z/lambda update runs as the kernel's epilogue
    # bytes the kernel has to move per launch: u^, du read once per problem (2 x N x m), x^_0 (n), the control-cost
    # polynomial (3), best index / best cost written (2)
    alg_bytes = float(B) * (2 * p["N"] * 2 + 4 + 6 + 2) * 8
    # two-phase bound of the fused kernel: FP64 phase at the DFMA peak + streaming ADMM epilogue (z, lambda read;
    # z, lambda, reg written = 5 doubles per control element; u^, du are re-read from cache) at the measured HBM peak
    epi_bytes = float(B) * p["N"] * 2 * 5 * 8 if fused_update else 0.0
    t_bound = flops / (fp64_peak * 1e12) + epi_bytes / (hbm_peak * 1e9)
    shape = "5,4,3" if p["L"] <= 20 else "5,10,1"            # csrc/isls_kernels.cuh: launch_linesearch
    src = os.path.join(PKG, "csrc", "isls_b200.cu")
    roof = {"kernel": "k_linesearch<CarModel,%s>" % shape + (" + fused ADMM z/lambda epilogue" if fused_update else ""),
            "bound": "fp64", "achieved": round(ach, 3),
            "peak": round(fp64_peak, 3), "unit": "TFLOP/s", "frac": round(ach / fp64_peak, 4),
            "frac_executed_flop": round(exec_flops / t_ls / 1e12 / fp64_peak, 4),
            "peak_source": "DFMA throughput measured live by isls_measure_fp64_tflops (k_fp64_peak in csrc/isls_b200.cu, "
                           "sha256 of the file %s; MEASURED_PEAKS.json has no FP64 figure; theoretical 148 SM x 64 x 2 "
                           "x 1.965 GHz = 37.2)" % hashlib.sha256(open(src, "rb").read()).hexdigest()[:16],
            "algorithmic_flop_per_launch": flops, "flop_per_candidate_step": FLOP_PER_CAND_STEP_CAR,
            "executed_flop_per_candidate_step": EXEC_FLOP_PER_CAND_STEP_CAR, "problems": B,
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes + epi_bytes,
                    "achieved_gbs": round((alg_bytes + epi_bytes) / t_ls / 1e9, 2),
                    "peak_gbs": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at B=65,536 from the committed
            # `ncu --set full` capture (profiles/r2c_ncu_c5_kernels.csv), scaled to this batch size
            "traffic": round(763.6e6 * B / 65536.0), "traffic_source": "profiles/r2c_ncu_c5_kernels.csv",
            "two_phase": {"epilogue_bytes_per_launch": epi_bytes, "bound_ms": round(t_bound * 1e3, 4),
                          "measured_ms": round(t_ls * 1e3, 4), "frac": round(t_bound / t_ls, 4)}}
    if "ff" in prof:
        # second-largest kernel (HBM-bound): ff-pass + linear rollout, 40 doubles per problem-step (DESIGN.md section 3:
        # bwd x^2 u^2 reg2 Qux8 Quu3 Quu^-1 3 + k2 st; fwd K8 k2 u^2 reg2 x^2 + du2 st - only the two state components the
        # car's Jacobian reads are staged); below 1,536 tiles the profile pass times the small-batch form of the same
        # kernel (deeper ring, Jacobian cache)
        ff_ms, ff_n = prof["ff"]
        ff_bytes = float(B) * p["N"] * 40 * 8
        roof["second_kernel"] = {"kernel": "k_ff_tma<CarModel,PX=0,JC=0,2,2> (TMA-staged ff-pass + linear rollout)", "bound": "hbm",
                                 "achieved": round(ff_bytes / (ff_ms / ff_n * 1e-3) / 1e9, 1), "peak": hbm_peak,
                                 "unit": "GB/s", "frac": round(ff_bytes / (ff_ms / ff_n * 1e-3) / 1e9 / hbm_peak, 4),
                                 "algorithmic_bytes_per_launch": ff_bytes, "doubles_per_problem_step": 40,
                                 # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 65,536 problems
                                 # (profiles/r2c_ncu_c5_kernels.csv: 1.841 + 0.205 GB; k is written and read back
                                 # through L2), scaled to this batch size
                                 "traffic": round(2045.9e6 * B / 65536.0)}
    if "kpass" in prof:
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
                                  # dram bytes read + written of one launch, profiles/r2c_ncu_c5_kernels.csv
                                  "traffic": round(1669.9e6 * B / 65536.0)}
    return roof


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
