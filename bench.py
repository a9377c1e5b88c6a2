#!/usr/bin/env python
"""Benchmark of the gpfanova Gibbs-sampling hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one Gibbs sweep (all samplers of base.py:57-84: 1 + f + 2P blocks) of every chain.  Workload: BASELINE
config 5 -- two-way 3x3 FANOVA with interaction, 5 replicates (m=45, f=9, P=4), n=256 time points, 4096 chains per GPU
(chains are independent units: every rank samples its own 4096, no collective on the sampling path -- weak scaling;
--scaling strong splits ONE set of 4096 chains over the N ranks instead).
Metric: posterior function draws/s = chains * f * sweeps / time, FP64.
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

METRIC = "posterior function draws/sec (FP64, n=256)"
UNIT = "draws/s"
WORKLOAD = "cfg5_twoway_n256"


def problem(n=256):
    from gpfanova_b200 import synthetic
    levels, reps, n_, inter, hc, chains = synthetic.CONFIGS[WORKLOAD]
    eff = synthetic.effects(levels, reps)
    x = np.linspace(-1, 1, n)[:, None]
    return x, eff, inter, hc, chains


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (py3-importable copy in oracle/_ref, emitted by oracle/build_ref.py) or, if that is
# absent, the oracle port.  One chain per process, single-threaded BLAS, one process per host core.
# ----------------------------------------------------------------------------------------------------------
def cpu_worker(args):
    """child process: time `sweeps` sweeps of one chain after `warm` untimed ones; prints one JSON line"""
    n, sweeps, warm, chain = args.n, args.cpu_sweeps, args.cpu_warm, args.cpu_chain
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    x, eff, inter, hc, _ = problem(n)
    from gpfanova_b200 import synthetic
    kind = "port"
    try:
        import build_ref
        if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "gpfanova")):
            build_ref.build(quiet=True)
        gp = build_ref.import_ref()
        kind = "reference"
    except Exception:
        gp = None
    import gpfanova_oracle as orc
    d = orc.Design(eff, interactions=inter, helmertConvert=hc)
    y, _ = synthetic.prior_draw(x, d.X, d.prior_groups())
    start = synthetic.chain_starts(1, len(d.prior_groups()), first=chain)[0]
    np.random.seed(1000 + chain)
    if gp is not None:
        import logging
        logging.disable(logging.CRITICAL)
        m = gp.fanova.FANOVA(x, y, eff, interactions=inter, helmertConvert=hc)
        names = ['y_sigma'] + [s % g for g in range(len(d.prior_groups())) for s in ('prior%d_sigma', 'prior%d_lengthscale')]
        for nm, v in zip(names, start):
            m.parameter_cache[nm] = v
        m.sample(warm, thin=0)
        t0 = time.perf_counter()
        m.sample(sweeps, thin=10)
        dt = time.perf_counter() - t0
    else:
        m = orc.Model(x, y, d.X, d.prior_groups())
        m.hyper = start.copy()
        st = orc.Stream(chain)
        for s in range(warm):
            m.sweep(st, chain, s)
        t0 = time.perf_counter()
        for s in range(sweeps):
            m.sweep(st, chain, warm + s)
        dt = time.perf_counter() - t0
    print(json.dumps({"kind": kind, "sweeps": sweeps, "seconds": dt, "f": int(d.f)}), flush=True)


def run_cpu_arm(n, sweeps, warm, procs=None):
    """aggregate draws/s of `procs` single-threaded reference processes running concurrently"""
    procs = procs or os.cpu_count() or 1
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", "--n", str(n), "--cpu-sweeps", str(sweeps),
                            "--cpu-warm", str(warm), "--cpu-chain", str(c)], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, text=True) for c in range(procs)]
    outs = []
    for p in ps:
        o, _ = p.communicate()
        for line in o.splitlines():
            if line.startswith("{"):
                outs.append(json.loads(line))
    wall = time.perf_counter() - t0
    if not outs:
        raise RuntimeError("CPU baseline workers produced no output")
    rate = sum(o["f"] * o["sweeps"] / o["seconds"] for o in outs)
    return {"value": rate, "unit": UNIT, "cores": len(outs), "kind": outs[0]["kind"],
            "sample": "%d sweeps (after %d warm-up) of 1 chain per process, %d processes, single-threaded BLAS, same "
                      "cfg5 data (n=%d, m=45, f=9, P=4); per-core %.2f draws/s; wall %.1f s" % (
                          sweeps, warm, len(outs), n, rate / len(outs), wall),
            "ms_per_sweep_per_chain": 1e3 * float(np.mean([o["seconds"] / o["sweeps"] for o in outs]))}


# ----------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []          # (arrival time, csv line)
        self.p = None
        self.t_start = self.t_end = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_start(self):
        self.t_start = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons, pw = [], [], set(), []
        # samples that arrived during the timed region; short regions fall back to everything since the sampler was
        # started (it is started before the warm-up sweeps, so the GPU is under the same load)
        timed = [l for (t, l) in self.lines if self.t_start is not None and self.t_start <= t <= (self.t_end or t)]
        window = "timed region"
        if len(timed) < 2:
            timed = [l for (_, l) in self.lines][1:] or [l for (_, l) in self.lines]
            window = "warm-up + timed region"
        for l in timed:
            parts = [q.strip() for q in l.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm), "window": window}


def fp64_peak():
    """measured FP64 tensor (DMMA) peak of this pool's B200: MEASURED_PEAKS.json has no FP64 entry, so the denominator is
    our own microbenchmark (tools/fp64_peak.cu, result committed as profiles/fp64_peak_r0*.json)"""
    try:
        for nm in ("fp64_peak_r02.json", "fp64_peak_r01.json"):
            path = os.path.join(ROOT, "profiles", nm)
            if os.path.exists(path):
                d = json.load(open(path))
                key = "dmma884_tflops_sustained" if "dmma884_tflops_sustained" in d else "dmma884_tflops"
                return float(d[key]), "profiles/%s (tools/fp64_peak.cu, DMMA m8n8k4 %s, measured on this pool)" % (
                    nm, "sustained" if key.endswith("sustained") else "burst")
        raise OSError("no FP64 peak file")
    except Exception:
        return 37.0, "fallback 37 TFLOP/s (B200 FP64 datasheet order; profiles/fp64_peak_r01.json missing)"


def ncu_traffic(kernel_file, n, chains_per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed ncu --set full
    summary (profiles/, captured on the default workload: n = 256, 4096 chains on one GPU); None for other workloads"""
    if n != 256 or chains_per_gpu != 4096:
        return None
    path = os.path.join(ROOT, "profiles", kernel_file)
    try:
        tot, seen = 0.0, 0
        for line in open(path):
            t = line.split()
            if len(t) >= 3 and t[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and seen < 2:
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(t[2], None)
                if scale is None:
                    return None
                tot += float(t[1]) * scale
                seen += 1
        return tot if seen == 2 else None
    except (OSError, ValueError):
        return None


def ncu_metric(kernel_file, metric, n, chains_per_gpu):
    """first value of `metric` in the committed ncu summary of the dominant kernel (same capture as ncu_traffic)"""
    if n != 256 or chains_per_gpu != 4096:
        return None
    try:
        for line in open(os.path.join(ROOT, "profiles", kernel_file)):
            t = line.split()
            if len(t) >= 2 and t[0] == metric:
                return float(t[1])
    except (OSError, ValueError):
        pass
    return None


def flops_model(n, f, P, ng_list, counts, schur, schurfn):
    """algorithmic FP64 flops of what was executed (counted by the kernels), DESIGN.md 5.
    Schur recursion on an n x n Toeplitz matrix in the scaled (square-root-free) form: per step and live position
    u' = u - rho v, v' = e v - rho u' = two multiply-adds and one multiply = 5 flops, n^2/2 live positions = 2.5 n^2;
    every right-hand side riding along adds n^2 (forward substitution, one multiply-add per live position).
      slice evaluation               (2.5 + |g|) n^2      (rotation + quadratic forms of the group's functions)
      function step, K_o pass        2.5 n^2 + n^2 per function of the group (prior draw f = L_B z1 accumulated)
      function step, C pass          5 n^2 (rotation of the Toeplitz generator and of the inverse-factor rows)
                                     + 2 n^2 per function (forward substitution + accumulation of v = L^-T y)
    (the per-step scalars -- reflection coefficient, reciprocal, pivots -- and the dead lanes of the SIMD layout are NOT
    counted: they are overhead of the implementation, which is why the FP64 pipe is busier than these numbers say)
    Dense paths (non-uniform grids, n > 256, large groups): factor of K_o per (group, chain, sweep) and of C per class
    n^3/3 each (centrosymmetric split: two half-size blocks, n^3/12); per draw two triangular solves and one to three
    triangular products ~ 5 n^2; dense slice factorisation n^3/3 + |g| n^2.  K_inv operator (not on the sweep path) n^3."""
    n2, n3 = float(n) ** 2, float(n) ** 3
    bfact = counts.get("ko_factor", 0)
    fchol = counts.get("function_chol", counts["draws"])       # passes over C: functions with the same precision share one
    inv = counts["inverses"] * n3
    if schurfn:
        draws = bfact * 2.5 * n2 + counts["draws"] * n2 + fchol * 5 * n2 + counts["draws"] * 2 * n2
    else:
        split = schur and n % 2 == 0 and n >= 64
        draws = (bfact + fchol) * (n3 / 12 if split else n3 / 3) + counts["draws"] * (2.5 if split else 5.0) * n2
    prior_facts = counts["chol"] - counts["inverses"] - fchol - bfact
    gbar = float(np.mean(ng_list))
    per = (2.5 + gbar) * n2 if schur else (n3 / 3 + gbar * n2)
    return {"kinv": inv, "function": draws, "prior_slice": prior_facts * per, "prior_factorisations": prior_facts,
            "total": inv + draws + prior_facts * per}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=0, help="chains (per GPU when weak, in total when strong; default: the workload's 4096)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every GPU samples the workload's 4096 chains -- chains are independent units, no "
                         "collective on the path; strong: the 4096 chains are split over the GPUs")
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--burnin", type=int, default=20, help="untimed sweeps before warm-up (chains leave their start state)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-worker", action="store_true")
    ap.add_argument("--cpu-sweeps", type=int, default=5)
    ap.add_argument("--cpu-warm", type=int, default=1)
    ap.add_argument("--cpu-chain", type=int, default=0)
    ap.add_argument("--cpu-procs", type=int, default=0)
    args = ap.parse_args()
    if args.cpu_worker:
        return cpu_worker(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        # each step = one sweep of one chain per host core (bounded sample of the workload)
        res = run_cpu_arm(args.n, max(1, args.steps), max(0, args.warmup), args.cpu_procs or None)
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_sweep_per_chain"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": WORKLOAD, "n": args.n, "m": 45, "f": 9, "P": 4, "chains": res["cores"],
                           "note": "reference CPU sampler, one chain per core"},
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: gpfanova_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gpfanova_b200 import engine, fanova, synthetic

    x, eff, inter, hc, total_chains = problem(args.n)
    if args.chains:
        total_chains = args.chains
    if args.scaling == "weak":
        C = total_chains                            # weak scaling: every rank samples a full workload of independent chains
        total_chains = C * world
    else:
        C = total_chains // world                   # strong scaling: the workload's chains are split over the ranks
    first = rank * C
    model = fanova.FANOVA(x, np.zeros((args.n, eff.shape[0])), eff, interactions=inter, helmertConvert=hc, chains=C,
                          seed=20261019, device=local, store_chains=True)
    y, Ftrue = synthetic.prior_draw(x, model.design_matrix, model.priorGroups())
    groups = [list(g) for g in model.priorGroups()]
    P, f, n = len(groups), model.f, args.n
    start = synthetic.chain_starts(C, P, first=first)

    # ---------------- kernel-level arm: inputs resident in HBM ----------------
    eng = engine.ChainEngine(x, y, model.design_matrix, groups, chains=C, seed=20261019, device=local, chain_offset=first)
    eng.set_state(hyper=start)
    eng.sweep(args.burnin)
    clocks = ClockSampler(local)
    clocks.start()
    eng.sweep(args.warmup)
    eng.sync()
    eng.status()
    eng.counters(clear=True)
    eng.set_profile(2)                        # one CUDA-event pair on the context's stream around the K sweeps
    _, launches0 = eng.profile()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.mark_start()
    t0 = time.perf_counter()
    eng.sweep(args.steps)
    eng.sync()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks.mark_end()
    clk = clocks.stop()
    prof_total, launches1 = eng.profile()
    dev_ms = prof_total["total"]
    eng.status()
    counts = eng.counters()
    # per-kernel breakdown for the roofline: the same K sweeps again with per-launch event brackets
    eng.counters(clear=True)
    eng.set_profile(1)
    eng.sweep(args.steps)
    eng.sync()
    prof, _ = eng.profile()
    counts_prof = eng.counters()
    eng.set_profile(0)
    schurfn = bool(eng.get_option(5))
    el = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    dev_ms_max = float(el.item())
    value = total_chains * f * args.steps / (dev_ms_max * 1e-3)

    # ---------------- roofline of the dominant kernel (largest share of the step) ----------------
    peak, peak_src = fp64_peak()
    schur = bool(np.allclose(np.diff(x[:, 0]), np.diff(x[:, 0])[0], rtol=0, atol=1e-13)) and x.shape[1] == 1
    fm = flops_model(n, f, P, [len(g) for g in groups], counts_prof, schur, schurfn)
    fm_timed = flops_model(n, f, P, [len(g) for g in groups], counts, schur, schurfn)
    fn_name = ("k_group_schurfn (all Function._sample draws of a sweep on a uniform grid, one warp per (prior group, chain): Schur "
               "recursions on the Toeplitz matrices K_o and I + c K_o, prior draws, embedded inverse factor, Philox normals)"
               if schurfn else
               "k_group_fused (all Function._sample draws of a sweep, one (prior group, chain) item per CTA: RBF build + "
               "jitchol(K_o) + per class jitchol(I + c K_o) with the right-hand sides riding along, solves, Philox draws; "
               "centrosymmetric split: half-size blocks)")
    kern = {"kinv": ("k_group_factor (jitchol(K + 1e-9 I) of every prior group; only launched by sequential sweeps)", args.steps),
            "function": (fn_name, args.steps),
            "prior_slice": ("k_prior_schur (sigma+lengthscale slice steps: Toeplitz/Schur GP log-likelihood, one warp per chain; "
                            "one launch per prior group)" if schur else
                            "k_prior (sigma+lengthscale slice steps: dense GP log-likelihood)", P * args.steps)}
    dom = max(("kinv", "function", "prior_slice"), key=lambda k: prof[k])
    tf = {k: (fm[k] / (prof[k] * 1e-3) / 1e12 if prof[k] > 0 else 0.0) for k in kern}
    ncu_file = {"function": "ncu_r02_final_k_group_schurfn.txt" if schurfn else "ncu_r02_k_group_fused.txt",
                "prior_slice": "ncu_r02_final_k_prior_schur.txt"}.get(dom, "")
    roofline = {"bound": "tensor", "kernel": kern[dom][0], "achieved": tf[dom], "peak": peak, "unit": "TFLOP/s",
                "frac": tf[dom] / peak,
                "bound_note": "FP64 pipe: B200 issues FP64 FMA and FP64 tensor (DMMA) work at the same rate (measured 33.98 / "
                              "37.09 TFLOP/s sustained, profiles/fp64_peak_r02.json); the Schur kernels are FMA work, the dense "
                              "kernels DMMA -- one denominator for both",
                "traffic": ncu_traffic(ncu_file, n, C),
                "fp64_pipe_active_pct_ncu": ncu_metric(ncu_file, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", n, C),
                "pipe_note": "achieved/frac count ALGORITHMIC flops only (live positions of the recursion); the per-step scalars and "
                             "the dead SIMD lanes of the register layout also occupy the FP64 pipe, which ncu sees busy for "
                             "fp64_pipe_active_pct_ncu percent of the cycles",
                "traffic_note": "bytes per launch from profiles/%s (one ncu --set full capture of this workload).  Algorithmic "
                                "HBM bytes per sweep: the function values written (%.0f MB) and read back by the slice kernels "
                                "and the y_sigma kernel; every matrix is generated and consumed in registers" % (
                                    ncu_file, C * f * n * 8 / 1e6),
                "peak_source": peak_src,
                "flops_per_launch": fm[dom] / kern[dom][1], "ms_per_launch": prof[dom] / kern[dom][1],
                "share_of_step": prof[dom] / prof["total"] if prof["total"] else None,
                "per_kernel": {k: {"ms_per_step": prof[k] / args.steps, "tflops": tf[k], "frac": tf[k] / peak,
                                   "launches_per_step": kern[k][1] / args.steps} for k in kern},
                "whole_step": {"achieved": fm_timed["total"] / (dev_ms * 1e-3) / 1e12,
                               "frac": fm_timed["total"] / (dev_ms * 1e-3) / 1e12 / peak,
                               "note": "algorithmic flops of the executed algorithm (Schur slices are O(n^2), "
                                       "split factorisations are n^3/12, no inverse is formed)",
                               # SURVEY 8(d)'s work model of the reference's dense algorithm for the same sweeps:
                               # P n^3 + f (n^3/3 + 6 n^2) + (n^3/3) * (slice factorisations), per chain and sweep
                               "reference_algorithm_tflops": (
                                   (counts.get("ko_factor", 0) * float(n) ** 3 + counts["draws"] * (float(n) ** 3 / 3 + 6.0 * n * n) +
                                    fm_timed["prior_factorisations"] * float(n) ** 3 / 3) / (dev_ms * 1e-3) / 1e12)},
                "counted": {"factorisations_per_sweep_per_chain": counts["chol"] / (C * args.steps),
                            "logp_evals_per_slice_step": counts["logp"] / max(1, counts["slice_steps"]),
                            "jitter_retries": counts["jitter"]}}
    eng.close()

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        model.y = y                                  # host arrays; uploaded when the device context is created
        model.set_chain_state(hyper=start)
        ksteps = args.steps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.sample(max(1, args.warmup), thin=1)    # warm-up call: one-time context creation, allocations, module load
        torch.cuda.synchronize()
        first_s = time.perf_counter() - t0
        model.set_chain_state(hyper=start)           # the timed call re-uploads every chain's state from the host
        try:
            pinned_bytes = model.prepare(ksteps, thin=1)  # page-locked result buffer for the timed call (allocated outside it)
        except RuntimeError:                              # (no page-locked memory to be had: pageable results, staged pipeline)
            pinned_bytes = 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.sample(ksteps, thin=1)                 # H2D: chain states; D2H: every stored sweep of every chain + final state
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        el = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e_s = float(el.item())
        cols = model.H + f * n
        e2e = {"value": total_chains * f * ksteps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(C * cols * 8 / ksteps),
               "d2h_bytes_per_step": int(C * cols * 8 + C * cols * 8 / ksteps),
               "kernel_ms_per_step": dev_ms_max / args.steps,
               "seconds": e2e_s, "first_call_seconds": first_s, "prepared_result_bytes": int(pinned_bytes),
               "api": "FANOVA(...).sample(%d, thin=1) on host arrays, timed after a %d-sweep warm-up call of the same API "
                      "(first_call_seconds: that call, with the one-time creation of the device context and upload of "
                      "y / x / design): upload of all chain states from host memory, %d sweeps, every sweep's state of "
                      "every chain read back by DMA into the result array (page-locked, allocated before the timed call by "
                      "FANOVA.prepare), chain 0 appended to parameter_history, final state read back" % (ksteps, max(1, args.warmup), ksteps)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = run_cpu_arm(n, args.cpu_sweeps, args.cpu_warm, args.cpu_procs or None)
        except Exception as e:  # noqa
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "n": n, "m": int(eff.shape[0]), "f": int(f), "P": P, "chains": total_chains,
                           "chains_per_gpu": C, "parallelism": "chains sharded over %d GPU(s), no collective" % world,
                           "burnin_sweeps": args.burnin,
                           "l2": "no flush needed: the per-chain state read and written every sweep (%.0f MB of function values + "
                                 "hyper-parameters for %d chains) is the only HBM traffic of the Schur path and is touched once "
                                 "per kernel; the matrices live in registers" % (C * f * n * 8 / 1e6, C)},
                "wall_ms_per_step": 1e3 * wall / args.steps, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": int(launches1 - launches0)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
