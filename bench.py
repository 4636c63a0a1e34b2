#!/usr/bin/env python
"""bench.py -- EMAT log-lik evals/s + SPR candidates scored/s on B200 (BASELINE.json metric).

One "step" = one log-G evaluation (calc_lambda_i + calc_log_root_prior + calc_log_G_below_root)
of every EMAT of a forest of `--chains` independent synthetic 100k-tip x 29,903-site EMATs (BASELINE.json configs[3]
shape; the forest is larger than the 126 MB L2 so the timed kernels stream from HBM).  K steps are enqueued back to
back between two CUDA events.  The metric's second figure -- SPR candidates scored/s -- is timed the same way right
after (K batches of `--spr-studies` full, unbounded regraft studies), and so is the general log-G schedule.

  value               = log-lik evals/s, inputs resident in HBM (whole job, all ranks)
  spr_*               = SPR candidate regions scored/s
  model_change_cycle  = evals/s when every evaluation follows a model change (set_evo on every table + eval + read-back)
  e2e                 = the same log-lik metric through the C ABI with HOST buffers (upload + eval + download per step)
  roofline            = algorithmic bytes of the log-G evaluation / its CUDA-event duration vs the measured HBM peak
  cpu_baseline        = the reference's own CPU code (oracle/_ref, compiled from /root/reference) on the box's host cores

Multi-GPU (torchrun): each rank owns its own forest of chains (weak scaling; the path shards over independent
EMATs with no data-path collective -- SURVEY.md section 8e); NCCL is used for the barrier and the max-over-ranks time.
"""
from __future__ import annotations

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

METRIC = "EMAT log-lik evals/s (+ spr_candidates_per_s)"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, help="synthetic shape: 1..5 == BASELINE.json configs[0..4]")
    ap.add_argument("--chains", type=int, default=16, help="independent EMATs per GPU (forest must exceed L2)")
    ap.add_argument("--spr-studies", type=int, default=128, help="full SPR studies per step (0 disables)")
    ap.add_argument("--e2e-chains", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(cfg, chains):
    shapes = {1: "200-tip x 29,903-site", 2: "1,600-tip x 18,959-site (nu_l on)", 3: "10k-tip x 29,903-site",
              4: "100k-tip x 29,903-site", 5: "50k-tip x 197,000-site (heavy missing)"}
    return f"synthetic {shapes.get(cfg, 'small')} EMAT x {chains} independent chains per GPU"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel, cfg, chains):
    """DRAM bytes (read + write) per launch of `kernel` from the committed ncu --set full capture of this workload
    (profiles/traffic.json), or None when no capture matches."""
    try:
        for row in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))):
            if row["kernel"] == kernel and row["config"] == cfg and row["chains"] == chains:
                return row["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_run(emat, sites, budget_s, threads, t_max_tip, spr_xs=None):
    """Times the reference's own CPU functions (oracle/_ref) -- or the oracle port if _ref is absent -- on a bounded
    sample of the same workload.  bench.py is one of the three places allowed to execute oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import oracle_lib as ol
    from helpers import to_oracle
    e, s = to_oracle(emat, sites)
    es, ss = e.as_struct(), s.as_struct()
    out = {}
    if ol.ref_available():
        lib = ol.ref()
        kind = "reference"
        lg = C.c_double()
        t1 = lib.ref_bench_log_G(C.byref(es), C.byref(ss), 2, threads, C.byref(lg))      # calibration
        per = max(t1 / 2, 1e-6)
        reps = max(2, int(budget_s / per))
        t = lib.ref_bench_log_G(C.byref(es), C.byref(ss), reps, threads, C.byref(lg))
        out.update(value=threads * reps / t, unit=UNIT, cores=threads, kind=kind, log_G=lg.value,
                   sample=f"{reps} evals/thread x {threads} threads of one {emat.num_nodes}-node EMAT "
                          f"(calc_lambda_i + calc_log_root_prior + calc_log_G_below_root), {t:.1f} s")
        if spr_xs is not None and len(spr_xs):
            xs = np.ascontiguousarray(spr_xs, np.int32)
            nreg = C.c_int64()
            t0 = lib.ref_bench_spr(C.byref(es), C.byref(ss), xs.ctypes.data_as(ol.i32p), min(len(xs), threads), threads, 0.8,
                                   t_max_tip, C.byref(nreg))
            per_x = max(t0 / max(1, min(len(xs), threads)) * threads, 1e-6) / threads
            n_x = int(max(threads, min(len(xs), budget_s / max(per_x, 1e-9) * 1.0)))
            n_x = min(n_x, len(xs))
            t = lib.ref_bench_spr(C.byref(es), C.byref(ss), xs.ctypes.data_as(ol.i32p), n_x, threads, 0.8, t_max_tip, C.byref(nreg))
            out["spr"] = dict(value=nreg.value / t, unit="candidates/s", cores=threads, kind=kind,
                              sample=f"{n_x} full SPR studies (reconstruct_missing_sites_at + seed_fill_from + Spr_study), "
                                     f"{nreg.value} regions, {t:.1f} s")
    else:
        o = ol.Oracle("oracle")
        kind = "port"
        cq = o.cum_Q_l(s)
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < budget_s or n < 2:
            lam = o.lambda_i(e, s, cq); o.log_root_prior(e, s); o.log_G_below_root(e, s, lam); n += 1
        t = time.perf_counter() - t0
        out.update(value=n / t, unit=UNIT, cores=1, kind=kind, sample=f"{n} evals of one {emat.num_nodes}-node EMAT, 1 thread, {t:.1f} s")
    return out


def pick_spr_nodes(emat, n, seed=1234):
    rng = np.random.default_rng(seed)
    cand = np.array([v for v in rng.permutation(emat.num_nodes)[: 8 * n + 8] if v != emat.root and emat.parent[v] != emat.root], np.int32)
    return cand[:n]


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner at communicator
    creation), so file descriptor 1 is pointed at stderr for the whole run and the line goes to a private duplicate of the
    original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import delphy_b200 as db

    if args.impl == "reference":
        # the reference's own CPU implementation of the path, all host threads, rank 0 only
        if rank != 0:
            return 0
        threads = host_cores()
        emat, sites, info = db.synth_generate(db.synth_params(args.config))
        spr_xs = pick_spr_nodes(emat, max(args.spr_studies, threads)) if args.spr_studies > 0 else None
        per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
        vals, sprs = [], []
        last = None
        for i in range(args.warmup + args.steps):
            r = cpu_baseline_run(emat, sites, per_step, threads, info["t_max_tip"], spr_xs)
            if i >= args.warmup:
                vals.append(r["value"]); sprs.append(r.get("spr", {}).get("value", 0.0))
            last = r
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.config, args.chains), "host_threads": threads},
                "spr_candidates_per_s": float(np.mean(sprs)) if sprs else None,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": last["kind"], "sample": last["sample"],
                                 "spr": last.get("spr")},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit_line(line)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = db.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    # ---- synthetic inputs: `chains` distinct EMATs per rank (different seeds) --------------------------------------
    emats, tables, infos = [], [], []
    base_seed = 20251017 + 1000 * rank
    for c in range(args.chains):
        e, s, info = db.synth_generate(db.synth_params(args.config, seed=base_seed + c))
        emats.append(e); infos.append(info)
        tables.append(db.DeviceSites(ctx, s))
    host_sites = [t.host for t in tables]
    forest = db.Forest(ctx, emats, tables, sites_index=np.arange(args.chains))
    alg_bytes = forest.log_G_algorithmic_bytes

    # SPR requests: full studies of random attached nodes of chain 0 (seeded as Subrun::spr1_move does)
    spr_reqs = None
    spr_xs = pick_spr_nodes(emats[0], args.spr_studies) if args.spr_studies > 0 else np.zeros(0, np.int32)
    if args.spr_studies > 0 and hasattr(db, "spr_requests_for_attached"):
        forest.eval_log_G()
        lam0 = forest.lambda_i(0)
        spr_reqs = db.spr_requests_for_attached(emats[0], 0, spr_xs, lam0, infos[0]["t_max_tip"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def spr_batch():
        b = forest.spr_study_batch(spr_reqs)
        b.close()                    # stream-ordered free: the memory is reused by the next batch

    for _ in range(max(args.warmup, 3)):
        forest.eval_log_G()
        if spr_reqs is not None:
            spr_batch()
    barrier()

    # ---- timed region ------------------------------------------------------------------------------------------------
    # A step = one log-G evaluation of every EMAT of the forest (the metric's unit).  The K steps are enqueued back to back
    # between two events on the launching stream, so the interval is device time, not launch latency.  The SPR sweep
    # (the metric's second figure) is timed the same way right after, K batches of `--spr-studies` full studies, and so is
    # the general (per-event) log-G schedule, which re-reads every list instead of the per-branch folded weights.
    def ev():
        return torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    barrier()
    e0.record(stream)
    h0 = time.perf_counter()
    for i in range(args.steps):
        forest.eval_log_G()
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / args.steps    # CPU cost of enqueuing one step (must stay below ms_per_step)
    e1.record(stream)
    barrier()
    launches = ctx.launches - launches0
    logg_ms = e0.elapsed_time(e1) / args.steps
    _, _, lg = forest.log_G()        # checksum of the timed work: log G of every chain

    spr_ms = None
    spr_regions = 0
    spr_launches = 0
    if spr_reqs is not None:
        l0 = ctx.launches
        barrier()
        e2.record(stream)
        for i in range(args.steps):
            spr_batch()
        e3.record(stream)
        barrier()
        spr_ms = e2.elapsed_time(e3) / args.steps
        spr_launches = ctx.launches - l0
        b = forest.spr_study_batch(spr_reqs)
        spr_regions = b.total_regions()
        b.close()

    # the general schedule (every mutation / missation / from-state list re-read per evaluation)
    ctx.set_log_G_path("general")
    for _ in range(3):
        forest.eval_log_G()
    g0, g1 = ev(), ev()
    barrier()
    g0.record(stream)
    for i in range(args.steps):
        forest.eval_log_G()
    g1.record(stream)
    barrier()
    gen_ms = g0.elapsed_time(g1) / args.steps
    _, _, lg_gen = forest.log_G()
    ctx.set_log_G_path("auto")
    assert np.allclose(lg_gen, lg, rtol=1e-11), "folded and general log-G schedules disagree"
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([logg_ms, spr_ms or 0.0, gen_ms], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    logg_ms_max, spr_ms_max, gen_ms_max = [float(x) for x in t.tolist()]

    # ---- the reference's global-move cycle: new mu on every chain's site table, re-evaluate every chain, read log G back --------
    # (what Run does after a global move, core/run.cpp:437-453; every evaluation here follows a real model change)
    mc_steps = max(3, min(args.steps, 20))
    mus = [t.host.mu.copy() for t in tables]
    def model_change_cycle(i):
        for k, t in enumerate(tables):
            t.set_evo(mu=mus[k] * (1.0 + 1e-3 * ((i % 7) + 1)))
        forest.eval_log_G()
        return forest.log_G()
    for i in range(2):
        model_change_cycle(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(mc_steps):
        mc_out = model_change_cycle(i)
    torch.cuda.synchronize()
    mc_s = time.perf_counter() - t0
    for k, t in enumerate(tables):
        t.set_evo(mu=mus[k])
    forest.eval_log_G()
    assert np.allclose(forest.log_G()[2], lg, rtol=1e-12), "restoring the model does not restore log G"
    tm = torch.tensor([mc_s], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    mc_s = float(tm.item())

    # ---- e2e: host buffers -> C ABI -> host scalars, copies inside the timed region ------------------------------
    n_e2e = min(args.e2e_chains, args.chains)
    # the caller's EMAT arrays live in page-locked host memory (dphy_host_alloc), as the contract's e2e leg asks: the upload
    # DMAs them from where they lie
    e2e_emats = [e.pinned(ctx) for e in emats[:n_e2e]]
    e2e_steps = max(3, min(args.steps, 10))
    h2d = 0
    for e in e2e_emats:
        for k in db.HostEmat.FIELDS_I32 + db.HostEmat.FIELDS_U8 + db.HostEmat.FIELDS_F64:
            h2d += getattr(e, k).nbytes
    d2h = n_e2e * 3 * 8

    def e2e_step():
        fo = db.Forest(ctx, e2e_emats, tables[:n_e2e], sites_index=np.arange(n_e2e))
        fo.eval_log_G()
        out = fo.log_G()
        fo.close()
        return out
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_out = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    assert np.allclose(e2e_out[2], lg[:n_e2e], rtol=1e-12)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = args.chains * world / (logg_ms_max * 1e-3)
        achieved = alg_bytes / (logg_ms_max * 1e-3) / 1e9
        traffic = measured_traffic("emat_log_G_folded_kernel", args.config, args.chains)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": logg_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, args.chains), "chains_per_gpu": args.chains,
                       "step": "one log-G evaluation (lambda_i + root prior + log G below root) of every EMAT of the forest",
                       "nodes_per_chain": emats[0].num_nodes, "mutations_per_chain": infos[0]["num_mutations"],
                       "missation_intervals_per_chain": infos[0]["num_intervals"], "max_depth": infos[0]["max_depth"],
                       "forest_device_bytes": forest.device_bytes, "l2": "inputs larger than L2 (forest > 126 MB)" if forest.device_bytes > 130e6 else "forest fits in L2",
                       "spr_studies_per_batch": int(args.spr_studies if spr_reqs is not None else 0), "parallelism": f"chains x{world}"},
            "roofline": {"kernel": "emat_log_G_folded_kernel (+ emat_log_G_folded_tree_kernel)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": logg_ms_max,
                         "note": "algorithmic bytes = SURVEY 8(d) per-event figure; this schedule reads per-branch folded weights "
                                 "instead of the missation / from-state lists, so its DRAM traffic is below the algorithmic bytes"},
            "loglik_general_schedule": {"value": args.chains * world / (gen_ms_max * 1e-3), "unit": UNIT, "launch_ms": gen_ms_max,
                                        "achieved": alg_bytes / (gen_ms_max * 1e-3) / 1e9, "frac": alg_bytes / (gen_ms_max * 1e-3) / 1e9 / peak,
                                        "note": "every mutation / missation / from-state list re-read per evaluation (also refreshes nsmn and the num_muts tallies)"},
            "model_change_cycle": {"value": args.chains * mc_steps * world / mc_s, "unit": UNIT, "ms_per_cycle": mc_s / mc_steps * 1e3,
                                   "note": "dphy_sites_set_evo(new mu) on every chain's site table + evaluation of every chain + log G read back, wall clock"},
            "spr_candidates_per_s": (spr_regions * world / (spr_ms_max * 1e-3)) if spr_ms else None,
            "spr_regions_per_batch": spr_regions, "spr_ms_per_batch": spr_ms_max if spr_ms else None,
            "e2e": {"value": n_e2e * e2e_steps * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "chains_per_step": n_e2e, "steps": e2e_steps},
            "host_enqueue_ms_per_step": host_enqueue_ms, "gpu_launches": int(launches), "gpu_launches_spr": int(spr_launches), "clocks": clocks, "log_G_checksum": float(np.sum(lg)),
        }
        if spr_ms:
            spr_alg = spr_regions * 60
            line["roofline_spr"] = {"bound": "hbm", "achieved": spr_alg / (spr_ms_max * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": spr_alg / (spr_ms_max * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_candidate": 60,
                                    "launch_ms": spr_ms_max}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_run(emats[0], host_sites[0], args.cpu_seconds, host_cores(), infos[0]["t_max_tip"],
                                                    spr_xs if spr_reqs is not None else None)
        emit_line(line)
    forest.close()
    for tb in tables:
        tb.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
