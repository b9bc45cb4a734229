"""Runs the ladder back to back for ~12 s and prints a timeline of per-launch time against nvidia-smi clocks / power /
throttle reasons, to see whether late phases of bench.py slow down because the SM clock drops."""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from curve25519_b200 import api
api.init(0)
n = 1 << 20
rng = np.random.Generator(np.random.PCG64(1))
sk = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
pk = torch.from_numpy(rng.integers(0, 256, (n, 32), dtype=np.uint8)).cuda()
out = torch.empty_like(sk)
rows = []
q = "timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,pstate,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown"
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader", "-lms", "250"], stdout=subprocess.PIPE, text=True)
def rd():
    for l in p.stdout: rows.append((time.time(), l.strip()))
threading.Thread(target=rd, daemon=True).start()
t0 = time.time(); tl = []
for phase in range(3):
    for i in range(200):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); api.x25519_shared(pk, sk, out=out, sk_inplace=True); e1.record(); torch.cuda.synchronize()
        tl.append((time.time() - t0, e0.elapsed_time(e1)))
    if phase == 0:
        print("--- 1.0 s idle gap"); time.sleep(1.0)
    if phase == 1:
        print("--- killing and restarting nvidia-smi sampler"); p.terminate(); p.wait()
        p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader", "-lms", "250"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=rd, daemon=True).start()
p.terminate()
for k in range(0, len(tl), 20):
    print("t=%6.2fs  launch %3d  %.2f ms" % (tl[k][0], k, tl[k][1]))
print("nvidia-smi samples:")
for t, l in rows[::2]:
    print("t=%6.2fs %s" % (t - t0, l))
