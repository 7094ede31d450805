#!/usr/bin/env python
"""bench.py -- throughput of the fused bootstrap-particle-filter hot path on B200 (contract: task prompt (4)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4|cfg5]

A "step" is one pass of the hot path over one batch of synthetic input: B independent PIMH proposals (bootstrap
filters) of the workload's model, every one running all T time steps through the fused CUDA kernel.
Default workload = BASELINE.json configs[1]: Levy-driven stochastic volatility (get_levy_sv), T=1000, N=4096,
1024 independent chains; theta = the reference's theta_mle, observations simulated from the model (seed 17).

One JSON line on stdout (rank 0).
  value     particle-steps/s with inputs resident in HBM (CUDA events around the K launches, max over ranks)
  e2e       the same metric through the ONE-SHOT C entry the R shim calls (cpimh_pf_batch: host observations in, log Z +
            trajectories + path indices out, every step; its HBM workspaces are cached in the library between calls)
  roofline  ALGORITHMIC bytes of the path store that actually ran (dense history or none 16d+4 B per particle-step,
            ancestor tree 24d+16 B) / kernel time, against the measured HBM copy peak.  `traffic` is the DRAM traffic of one ncu
            capture scaled to the launch ("static": not measured in this run).
  roofline_secondary  what actually bounds the kernel: warp-instruction issue.  achieved = warp instructions per particle-step of
            the committed ncu capture x measured particle-steps/s; peak = issue rate measured in this run by a synthetic kernel;
            plus the measured fp64 FMA peak.
  configs   the same three numbers for every BASELINE config shape (cfg1..cfg5), a few steps each (N = 1 only)
  cpu_baseline  the CPU oracle (a C restatement of the reference's R+Rcpp path, all host threads) on a bounded sample
`--impl reference` times that CPU path alone (the real R reference cannot run: no R toolchain).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

THETA_MLE = [0.24487179487179486, -0.27564102564102566, 0.8161538461538461, 0.09333333333333334, 0.05153846153846154]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def simulate_levy_sv(T, theta, seed=17):
    """y_t from the model's own robs (R/model_levy_sv.R:72-77) on a simulated latent path."""
    rng = np.random.default_rng(seed)
    mu, beta, xi, w2, lam = theta
    z = rng.gamma(xi * xi / w2, w2 / xi)
    y = np.empty((T, 1))
    for t in range(T + 1):
        k = rng.poisson(lam * xi * xi / w2)
        c = rng.random(k)
        e = rng.exponential(w2 / xi, k)
        znew = np.exp(-lam) * z + np.sum(np.exp(-lam * c) * e)
        v = (z - znew + np.sum(e)) / lam
        z = znew
        if t >= 1:
            y[t - 1, 0] = mu + beta * v + np.sqrt(max(v, 1e-12)) * rng.normal()
    return y


def simulate_gss(T, seed=17):
    rng = np.random.default_rng(seed)
    x = np.sqrt(2.0) * rng.normal()
    y = np.empty((T, 1))
    for t in range(1, T + 1):
        x = 0.5 * x + 25 * x / (1 + x * x) + 8 * np.cos(1.2 * (t - 1)) + rng.normal()
        y[t - 1, 0] = x * x / 20 + np.sqrt(10.0) * rng.normal()
    return y


def simulate_ar(T, d, alpha, sigma_y, seed=17):
    rng = np.random.default_rng(seed)
    A = np.array([[alpha ** (1 + abs(i - j)) for j in range(d)] for i in range(d)])
    x = rng.normal(size=d)
    y = np.empty((T, d))
    for t in range(T):
        x = A @ x + rng.normal(size=d)
        y[t] = x + sigma_y * rng.normal(size=d)
    return y


def bytes_per_particle_step(d, path_storage, spilled_weights=False):
    """gather read 8d + state write 8d + ancestor 4 (dense history or no path store: the state double buffer IS the
    history); the ancestor tree adds x_star 8d + a_star 4 + o_star read/write 8.  +16 when weights / CDF live in HBM (N = 65536)."""
    b = 24 * d + 16 if path_storage == "tree" else 16 * d + 4
    return b + (16 if spilled_weights else 0)


def workload(name):
    """(model factory name, theta, observations, N, B, oracle config kwargs)"""
    if name == "cfg2":
        T, N, B = 1000, 4096, 1024
        return dict(name="levy_sv T=1000 N=4096 B=1024 (BASELINE configs[1])", model="levy_sv", theta=THETA_MLE, obs=simulate_levy_sv(T, THETA_MLE),
                    N=N, B=B, T=T, d=2, dy=1)
    if name == "cfg1":
        T, N, B = 100, 256, 65536
        return dict(name="gss T=100 N=256 B=65536 (BASELINE configs[0], batched)", model="gss", theta=[], obs=simulate_gss(T), N=N, B=B, T=T, d=1, dy=1)
    if name == "cfg4":
        T, N, B = 100, 1024, 16384
        return dict(name="ar(10) T=100 N=1024 B=16384 (BASELINE configs[3] filter shape)", model="ar10", theta=[0.5, float(np.sqrt(10.0))],
                    obs=simulate_ar(T, 10, 0.5, float(np.sqrt(10.0))), N=N, B=B, T=T, d=10, dy=10)
    if name == "cfg3":
        g = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
        T, N, B = 100, 8192, 1184
        A = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
        return dict(name="stochastic_kinetic T=100 N=8192 B=1184 (BASELINE configs[2] filter shape)", model="sk", theta={"A_obs": A, "s2_obs": 1.0},
                    obs=g["sk_y"], N=N, B=B, T=T, d=4, dy=2)
    if name == "cfg5":
        T, N, B, d = 500, 65536, 1, 64
        rng = np.random.default_rng(17)
        return dict(name="lorenz96 d=64 rk4 T=500 N=65536 B=1 (BASELINE configs[4])", model="lorenz64", theta=[8.0, 0.5], obs=rng.normal(size=(T, d)) * 1.5 + 2.0,
                    N=N, B=B, T=T, d=d, dy=d, spilled=True)
    raise SystemExit("unknown workload %s" % name)


def make_model(up, w):
    return {"levy_sv": up.get_levy_sv, "gss": up.get_gss, "sk": up.get_stochastic_kinetic, "ar10": lambda: up.get_ar(10), "lorenz64": lambda: up.get_lorenz(1.0, 64, "rk4", 1)}[w["model"]]()


def oracle_config(O, w):
    mid = {"levy_sv": O.MODEL_LEVY_SV, "gss": O.MODEL_GSS, "sk": O.MODEL_SK, "ar10": O.MODEL_AR, "lorenz64": O.MODEL_LORENZ}[w["model"]]
    th = w["theta"]
    if w["model"] == "lorenz64":
        return O.make_config(mid, w["d"], w["dy"], w["N"], w["T"], theta=list(th) + [1.0], integrator=0, substeps=1)
    if w["model"] == "sk":
        th = [th["s2_obs"]] + list(np.ravel(th["A_obs"], order="F"))
    return O.make_config(mid, w["d"], w["dy"], w["N"], w["T"], theta=th)


def cpu_sample(w, target_s=12.0, min_filters=None):
    """Time the CPU oracle (all host threads) on a bounded sample of the workload: whole filters of the same shape."""
    from oracle import pyoracle as O
    cfg = oracle_config(O, w)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    O.cpf_batch(cfg, w["obs"], 17, 0, 1, nthreads=1, want_traj=False)
    t1 = time.perf_counter() - t0
    nf = int(max(cores, min(4096, (target_s / max(t1, 1e-6)) * cores)))
    if min_filters:
        nf = max(nf, min_filters)
    nf = max(cores, (nf // cores) * cores)
    t0 = time.perf_counter()
    O.cpf_batch(cfg, w["obs"], 17, 1000, nf, nthreads=cores, want_traj=False)
    dt = time.perf_counter() - t0
    ps = nf * w["N"] * w["T"]
    return dict(value=ps / dt, unit="particle-steps/s", cores=cores, kind="port",
                sample="%d whole filters of the workload shape (N=%d, T=%d), one filter per host thread, %.1f s" % (nf, w["N"], w["T"], dt)), dt, ps


class ClockSampler:
    def __init__(self, device_index):
        self.p = None
        self.path = "/tmp/cpimh_clocks_%d.csv" % os.getpid()
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device_index),
                                       "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                                       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The R package cannot run here (no R/Rcpp), so this
    is the oracle port with all host threads; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    per_step_s = min(20.0, 150.0 / (steps + warm))
    for _ in range(warm):
        cpu_sample(w, target_s=per_step_s / 4)
    tot_ps, tot_t, cb = 0, 0.0, None
    for _ in range(steps):
        cb, dt, ps = cpu_sample(w, target_s=per_step_s)
        tot_ps += ps; tot_t += dt
    v = tot_ps / tot_t
    cb["value"] = v
    line = {"metric": "particle-steps/sec", "value": v, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "config": {"workload": w["name"], "note": "CPU oracle port of R/CPF.R + src/*.cpp; R itself is not installable here"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def issue_counts():
    """warp instructions per particle-step of the committed ncu captures (profiles/issue_counts.json): static, per workload"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "issue_counts.json")))
    except Exception:
        return {}


def time_filters(torch, dist, up, w, steps, warm, rank, world, local, threads=0, store_paths=1, sample_clocks=False, e2e=True):
    """One workload: (1) K launches with inputs resident (CUDA events, max over ranks); (2) the same through the one-shot host entry."""
    model = make_model(up, w)
    N, T, B, d = w["N"], w["T"], w["B"], w["d"]
    pb = up.PfBatch(model, w["theta"], w["obs"], N, B, threads_per_filter=threads, store_paths=store_paths, device=local)
    stream = torch.cuda.current_stream().cuda_stream
    first = rank * B                                  # weak scaling: every rank owns B independent chains (distinct Philox streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        pb.run(B, 1000 + i, first, stream)
    pb.fetch(trajectory=False)
    l0 = pb.launch_count
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(steps):
        pb.run(B, 17 + i, first, stream)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pb.launch_count - l0
    clocks = sampler.stop() if sampler else None
    pb.fetch(trajectory=False)                        # surfaces per-filter errors
    repeats = None
    try:
        repeats = float(np.mean(pb.fetch_exact_repeats(B)))
    except Exception:
        pass
    tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    units = float(N) * T * B * steps * world
    eff = pb.effective_config()
    dev_bytes = pb.device_bytes
    pb.close()
    out = dict(value=units / (ms * 1e-3), ms_per_step=ms / steps, launches=int(launches), clocks=clocks, eff=eff, device_bytes=dev_bytes,
               exact_repeats_per_filter=repeats, units_per_step=float(N) * T * B)
    if e2e:
        # the entry the R shim calls: host observations in, log Z + trajectories + path indices out, every step
        kw = dict(threads_per_filter=threads, store_paths=store_paths)
        obs_pinned = torch.from_numpy(np.ascontiguousarray(w["obs"])).pin_memory().numpy()
        for i in range(2):
            r = up.cpf_batch(N, model, w["theta"], obs_pinned, B, seed=160 + i, first_filter=first, **kw)   # first call allocates the cached workspaces
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            r = up.cpf_batch(N, model, w["theta"], obs_pinned, B, seed=170 + i, first_filter=first, **kw)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": units / float(te.item()), "unit": "particle-steps/s", "h2d_bytes_per_step": int(w["obs"].size * 8),
                      "d2h_bytes_per_step": int(B * 8 + B * (T + 1) * d * 8 + B * 4),
                      "entry": "cpimh_pf_batch (one-shot C entry of r-pkg/src/shim.c cpimh_R_pf_batch; workspaces cached between calls)"}
        out["mean_logz"] = float(np.mean(r["logz"]))
        up.lib().cpimh_release_cached()
    return out


def roofline_of(w, res, peak, which, counts, issue_peak, name):
    eff = res["eff"]
    bps = bytes_per_particle_step(w["d"], eff["path_storage"], w.get("spilled", False))
    kernel_s = res["ms_per_step"] * 1e-3
    achieved = res["units_per_step"] * bps / kernel_s / 1e9
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get(name)
            if tj:
                traffic = tj["dram_bytes_per_particle_step"] * res["units_per_step"]
                tsrc = "static: %s, scaled by the particle-steps of one launch (not measured in this run)" % tj.get("source", "profiles/ncu_traffic.json")
        except Exception:
            pass
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
            "peak_source": which, "algorithmic_bytes_per_particle_step": bps, "path_storage": eff["path_storage"]}
    sec = None
    c = counts.get(name)
    if c and issue_peak:
        ach = c["warp_instructions_per_particle_step"] * res["units_per_step"] / kernel_s / 1e9
        sec = {"bound": "warp-instruction issue", "achieved": ach, "peak": issue_peak, "unit": "1e9 warp instructions/s", "frac": ach / issue_peak,
               "warp_instructions_per_particle_step": c["warp_instructions_per_particle_step"],
               "source": "static count from %s x particle-steps/s measured in this run; peak measured in this run (cpimh_measure_issue_peak)" % c.get("source", "profiles/")}
    return roof, sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--filters", type=int, default=0, help="override the number of filters per GPU")
    ap.add_argument("--threads", type=int, default=0, help="override threads per filter")
    ap.add_argument("--store-paths", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-chains", action="store_true", help="skip the coupled-PIMH replicates/s side measurements")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table (cfg1..cfg5)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    w = workload(args.workload)
    if args.filters:
        w["B"] = args.filters
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    import unbiased_pimh_b200 as up
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps, warm = max(args.steps, 1), max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    main_res = time_filters(torch, dist, up, w, steps, warm, rank, world, local, threads=args.threads, store_paths=args.store_paths, sample_clocks=True)

    # ---- a batch of whole waves (no 1024-thread tail launch): the kernel's own rate, for the roofline discussion
    whole = None
    if args.workload == "cfg2" and not args.filters:
        grid = main_res["eff"]["ctas_per_sm"] * torch.cuda.get_device_properties(local).multi_processor_count
        ww = dict(w); ww["B"] = 2 * grid
        rw = time_filters(torch, dist, up, ww, max(2, steps // 2), 3, rank, world, local, e2e=False)
        whole = {"filters_per_gpu": ww["B"], "value": rw["value"], "ms_per_step": rw["ms_per_step"]}

    # ---- secondary metric of BASELINE.json: coupled-PIMH estimator replicates/s (configs[3] shape: get_ar(10), N=1024,
    # T=100, K=1).  (a) weak: 50 000 replicates per GPU; (b) strong: configs[3]'s 100 000 replicates split over the ranks.
    # One all-reduce of the estimator sums; not part of `value`.
    cp_extra = cp_strong = None
    if not args.no_chains:
        arobs = simulate_ar(100, 10, 0.5, float(np.sqrt(10.0)))
        armodel, arth = up.get_ar(10), [0.5, float(np.sqrt(10.0))]

        def chains(Rrep, seed):
            lo, hi = up.shard_range(Rrep, rank, world)
            chb = up.ChainBatch(armodel, arth, arobs, 1024, hi - lo)          # HBM workspaces allocated once, outside the timed region
            up.unbiased_estimator_sharded(armodel, arth, arobs, 1024, Rrep, K=1, seed=seed, chain_batch=chb)
            barrier()
            t0 = time.perf_counter()
            est = up.unbiased_estimator_sharded(armodel, arth, arobs, 1024, Rrep, K=1, seed=seed + 1, chain_batch=chb)
            barrier()
            tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            chb.close()
            sec = float(tc.item())
            return {"metric": "coupled-PIMH estimator replicates/sec", "value": Rrep / sec, "replicates": Rrep, "seconds": sec,
                    "filters_run": est["nfilters"], "particle_steps_per_s": est["nfilters"] * 100.0 * 1024.0 / sec}
        cp_extra = chains(50000 * world, 3)
        cp_extra["workload"] = "get_ar(10) N=1024 T=100 K=1 (BASELINE configs[3] shape), 50 000 replicates per GPU (weak); host wall time of run + all-reduce, workspaces pre-allocated"
        cp_strong = chains(100000, 7)
        cp_strong["workload"] = "BASELINE configs[3]: 100 000 replicates of get_ar(10) N=1024 T=100 K=1 in total, sharded over the ranks (strong scaling)"
        cp_strong["scaling"] = "strong"

    # ---- coupled conditional-particle-filter chains (coupled_ccpf_chains + H_bar, ancestor sampling), the
    # LGSSM shape of inst/reproduce/lgssm (get_ar(1), T=100), N=128, k=3, K=10; replicates sharded, sums all-reduced.
    cc_extra = None
    if not args.no_chains:
        Rcc = 10000 * world
        obs1 = simulate_ar(100, 1, 0.5, float(np.sqrt(10.0)))
        m1, th1 = up.get_ar(1), [0.5, float(np.sqrt(10.0))]
        lo, hi = up.shard_range(Rcc, rank, world)
        up.coupled_ccpf_chains_batch(m1, th1, obs1, 128, hi - lo, K=10, k=3, with_as=True, seed=5, first_replicate=lo)   # warm-up at the timed size (the cached workspaces grow to it)
        times = []
        for rep in range(5):                 # median of five whole calls
            barrier()
            t0 = time.perf_counter()
            g = up.coupled_ccpf_chains_batch(m1, th1, obs1, 128, hi - lo, K=10, k=3, with_as=True, seed=6 + rep, first_replicate=lo)
            sums = torch.tensor(np.concatenate([g["sum_hbar"].ravel(), g["sum_hbar_sq"].ravel(), [g["nfilters"], hi - lo]]), dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(sums)
            barrier()
            tcc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tcc, op=dist.ReduceOp.MAX)
            times.append(float(tcc.item()))
        nf = float(sums[-2].item())
        tmed = sorted(times)[len(times) // 2]
        cc_extra = {"metric": "coupled-CCPF estimator replicates/sec", "value": Rcc / tmed, "replicates": Rcc, "filters_run": int(nf),
                    "mean_meeting_time": float(np.mean(g["meetingtime"])), "seconds_per_call": times,
                    "workload": "get_ar(1) N=128 T=100 k=3 K=10 ancestor sampling; median of 5 calls after one warm-up call of the same size, host wall time incl. H2D/D2H and the all-reduce"}

    # ---- every BASELINE config shape, a few steps each (N = 1 only: the driver's BENCH record then covers all five)
    cfg_table = None
    peak, which = load_peaks()
    counts = issue_counts()
    issue_peak = fp64_peak = None
    if rank == 0:
        v = C.c_double()
        if up.lib().cpimh_measure_issue_peak(C.byref(v)) == 0:
            issue_peak = v.value
        if up.lib().cpimh_measure_fp64_peak(C.byref(v)) == 0:
            fp64_peak = v.value
    if world == 1 and not args.no_configs and args.workload == "cfg2" and not args.filters:
        cfg_table = {}
        for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
            if name == "cfg2":
                r = main_res
            else:
                r = time_filters(torch, dist, up, workload(name), 3, 3, rank, world, local)
            wl = workload(name)
            roof, sec = roofline_of(wl, r, peak, which, counts, issue_peak, name)
            cfg_table[name] = {"workload": wl["name"], "value": r["value"], "ms_per_step": r["ms_per_step"], "e2e": r["e2e"]["value"], "frac": roof["frac"],
                               "algorithmic_bytes_per_particle_step": roof["algorithmic_bytes_per_particle_step"], "path_storage": roof["path_storage"],
                               "issue_frac": sec["frac"] if sec else None}
            if name == "cfg5":     # the wide filter: warp per group of G particles, 1024-thread CTAs, one launch per time step
                cfg_table[name].update({"particles_per_warp_group": r["eff"]["threads_per_filter"], "warp_groups": r["eff"]["ctas_per_sm"]})
            else:
                cfg_table[name].update({"threads_per_filter": r["eff"]["threads_per_filter"], "ctas_per_sm": r["eff"]["ctas_per_sm"]})

    if rank == 0:
        eff = main_res["eff"]
        roof, sec = roofline_of(w, main_res, peak, which, counts, issue_peak, args.workload)
        roof["note"] = "dominant (only) kernel pf_batch_kernel; bound by warp-instruction issue (fp64 log/exp, Philox), not by HBM: see roofline_secondary"
        if sec is not None:
            sec["fp64_fma_peak_tflops"] = fp64_peak
            prop = torch.cuda.get_device_properties(local)
            mhz = (main_res["clocks"] or {}).get("sm_mhz") or 1965.0
            sec["peak_theoretical"] = prop.multi_processor_count * 4 * mhz * 1e6 / 1e9      # SMs x 4 schedulers x clock
            sec["frac_of_theoretical"] = sec["achieved"] / sec["peak_theoretical"]
        line = {"metric": "particle-steps/sec", "value": main_res["value"], "unit": "particle-steps/s", "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "filters_per_gpu": w["B"], "nparticles": w["N"], "nobs": w["T"], "state_dim": w["d"],
                           "path_storage": eff["path_storage"], "threads_per_filter": eff["threads_per_filter"], "ctas_per_sm": eff["ctas_per_sm"],
                           "l2": "no flush needed: each launch streams %.1f GB of particle history through HBM (>> 126 MB L2)" % (main_res["device_bytes"] / 1e9),
                           "mean_logz": main_res.get("mean_logz"),
                           "exact_order_repeats_per_filter": main_res["exact_repeats_per_filter"]},
                "roofline": roof, "roofline_secondary": sec,
                "e2e": main_res["e2e"], "gpu_launches": main_res["launches"], "clocks": main_res["clocks"],
                "whole_waves": whole, "configs": cfg_table,
                "coupled_pimh": cp_extra, "coupled_pimh_strong": cp_strong, "coupled_ccpf": cc_extra}
        if world == 1 and not args.no_cpu:
            cb, _, _ = cpu_sample(w)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
