"""SM / memory clocks and power while log-G evaluations run back to back for a few seconds."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import delphy_b200 as db
ctx = db.Context(0)
ems, tabs = [], []
for c in range(16):
    e, s, info = db.synth_generate(db.synth_params(4, seed=20251017 + c))
    ems.append(e); tabs.append(db.DeviceSites(ctx, s))
fo = db.Forest(ctx, ems, tabs, sites_index=np.arange(16))
st = torch.cuda.ExternalStream(ctx.stream)
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
th = threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True); th.start()
time.sleep(0.5)
n_idle = len(rows)
for rnd in range(6):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(5000): fo.eval_log_G()
    b.record(st); ctx.synchronize()
    print(f"round {rnd}: {a.elapsed_time(b)/5000*1e3:.1f} us per evaluation; samples so far {len(rows)}")
time.sleep(0.3); p.terminate()
print("idle:", rows[:n_idle][-3:])
print("load:", rows[n_idle:][::6][:16])
