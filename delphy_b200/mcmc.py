"""Runs the reference's own CLI (tools/delphy.cpp, unmodified) -- either the stock build (oracle/_ref/delphy) or the build whose
hot path is substituted at link time by this repository (delphy_b200/adapter/_build/delphy_b200_cli) -- and parses its stats
line (tools/delphy.cpp:26-127): the posterior summaries named by BASELINE.json (t_MRCA, mu, n0) and the MCMC throughput.

Steps/s are measured from the arrival times of the stats lines (the CLI's own figure is printed with one decimal of M steps/s).
"""
import datetime
import os
import re
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STOCK_CLI = os.path.join(ROOT, "oracle", "_ref", "delphy")
DROPIN_CLI = os.path.join(ROOT, "delphy_b200", "adapter", "_build", "delphy_b200_cli")

_NUM = r"([-+0-9.eE]+|nan|inf|-inf)"
_PATTERNS = {
    "step": re.compile(r"Step (\d+),"),
    "log_posterior": re.compile(r"log_posterior = " + _NUM),
    "log_G": re.compile(r"log_G = " + _NUM),
    "log_coal": re.compile(r"log_coal = " + _NUM),
    "num_muts": re.compile(r"num_muts = (\d+),"),
    "T": re.compile(r" T = " + _NUM),
    "t_MRCA": re.compile(r"t_MRCA = (\d{4}-\d{2}-\d{2})"),
    "n0": re.compile(r"n0 = " + _NUM + " yr"),
    "g": re.compile(r" g = " + _NUM + " e-fold/yr"),
    "mu": re.compile(r"mu = " + _NUM + r" \* 10\^-3"),
    "kappa": re.compile(r"kappa = " + _NUM),
}
_EPOCH = datetime.date(2020, 1, 1)


def parse_stats_line(line):
    if "Step " not in line or "log_G" not in line:
        return None
    out = {}
    for k, pat in _PATTERNS.items():
        m = pat.search(line)
        if not m:
            continue
        v = m.group(1)
        if k == "t_MRCA":
            out[k] = (datetime.date.fromisoformat(v) - _EPOCH).days
        elif k in ("step", "num_muts"):
            out[k] = int(v)
        else:
            out[k] = float(v)
    return out if "step" in out else None


def run_cli(binary, maple, steps, threads=1, seed=1, log_every=None, extra_args=(), env=None, timeout=3600):
    """Returns {"samples": [parsed stats lines], "steps_per_s": ..., "wall_s": ..., "init_s": ..., "returncode": ..., "stderr_tail": ...}."""
    if log_every is None:
        log_every = max(1, steps // 10)
    cmd = [binary, "--v0-in-maple", maple, "--v0-steps", str(steps), "--v0-threads", str(threads), "--v0-seed", str(seed),
           "--v0-log-every", str(log_every)] + list(extra_args)
    e = dict(os.environ)
    if env:
        e.update({k: str(v) for k, v in env.items()})
    t_start = time.perf_counter()
    proc = subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=e, bufsize=1)
    samples, stamps, tail = [], [], []
    try:
        for line in proc.stderr:
            now = time.perf_counter()
            tail.append(line.rstrip()[:400])
            tail = tail[-12:]
            s = parse_stats_line(line)
            if s is not None:
                samples.append(s)
                stamps.append(now)
            if now - t_start > timeout:
                proc.kill()
                break
        proc.wait(timeout=60)
    finally:
        if proc.poll() is None:
            proc.kill()
    wall = time.perf_counter() - t_start
    sps = None
    if len(samples) >= 2 and stamps[-1] > stamps[0]:
        sps = (samples[-1]["step"] - samples[0]["step"]) / (stamps[-1] - stamps[0])
    return {"samples": samples, "steps_per_s": sps, "wall_s": wall, "init_s": (stamps[0] - t_start) if stamps else None,
            "mcmc_s": (stamps[-1] - stamps[0]) if len(stamps) >= 2 else None,
            "returncode": proc.returncode, "stderr_tail": tail, "cmd": cmd}


def posterior_means(samples, burnin_frac=0.3, keys=("t_MRCA", "mu", "n0", "g", "kappa", "log_G", "num_muts", "T")):
    """Mean and batch-means standard error of each summary over the post-burn-in samples."""
    import numpy as np
    n = len(samples)
    keep = samples[int(n * burnin_frac):]
    out = {}
    for k in keys:
        v = np.array([s[k] for s in keep if k in s], float)
        if len(v) < 4:
            continue
        nb = max(2, min(10, len(v) // 4))
        batches = np.array_split(v, nb)
        bm = np.array([b.mean() for b in batches])
        out[k] = {"mean": float(v.mean()), "sem": float(bm.std(ddof=1) / np.sqrt(nb)), "n": int(len(v))}
    return out
