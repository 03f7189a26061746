"""Small driver for ncu captures of the latency-mode persistent kernel (not a test):
    python tests/prof_latency.py n_u N [solves]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    sys.path.insert(0, p)
import torch
import gpad_b200 as G

n_u, N = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
prob = G.Problem("battery", n_u=n_u, N=N)
M_G, G_L = prob.operators()
x0 = (np.sin(1.0 + 2.0 * np.arange(n_u)) * 0.4)[None, :]
g_P, p_D, _ = prob.instances(x0, want_f=False)
theta, beta = G.schedule(100)
s = G.Solver(n_u, N, prob.m, prob.L, M_G, G_L, mode=G.MODE_LATENCY)
dg, dp = torch.from_numpy(g_P[0]).cuda(), torch.from_numpy(p_D[0]).cuda()
dz = torch.empty(prob.n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve_device(1, dg, dp, theta, beta, 100, stream=st, z=dz); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(s.description)
print(f"({n_u},{N}) p50 {np.median(ts):.1f} us per 100-iteration solve; |z| {float(dz.abs().max()):.6f}")
