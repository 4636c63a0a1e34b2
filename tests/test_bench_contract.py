"""bench.py's output contract, checked on CPU through the reference arm (the one leg that needs no GPU): exactly one JSON line
on stdout with the keys the driver reads; the N > 1 reference arm prints from rank 0 only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libdelphy_ref.so")


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1",
                           "--warmup", "0", "--cpu-seconds", "0.3", "--spr-studies", "2", "--mcmc-tips", "300", "--mcmc-steps", "20000"],
                          capture_output=True, text=True, timeout=300, env=env)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs the reference checkout)")
def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["ms_per_step"] > 0
    # the reference's own multithreaded scheme (tree cut into one part per thread) and its own CLI's MCMC throughput
    assert d["partitioned"]["kind"] == "reference" and d["partitioned"]["value"] > 0
    stock = os.path.join(ROOT, "oracle", "_ref", "delphy")
    if os.path.exists(stock):
        for case in ("cfg1_200_tips", "cfg3_300_tips"):
            assert d["mcmc"][case]["returncode"] == 0 and d["mcmc"][case]["steps_per_s"] > 0
    # the reference arm loads no product code: its inputs come from libdphy_synth.so
    assert "libdelphy_b200" not in r.stderr


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs the reference checkout)")
def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
