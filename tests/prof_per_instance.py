"""Small driver for ncu captures of the per-instance (batched GEMV) mode (not a test):
    python tests/prof_per_instance.py [B] [solves]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    sys.path.insert(0, p)
import torch
import gpad_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_u, N = 3, 4
prob = G.Problem("battery", n_u=n_u, N=N)
M_G, G_L = prob.operators()
n, m = prob.n, prob.m
rng = np.random.default_rng(5)
scale = (1.0 + 0.1 * (2 * rng.random((B, 1, 1)) - 1)).astype(np.float32)
dM = torch.from_numpy(M_G[None] * scale).cuda().contiguous(); dG = torch.from_numpy(G_L[None] / scale).cuda().contiguous()
g_P, p_D, _ = prob.instances(rng.random((B, n_u)) - 0.5, want_f=False)
theta, beta = G.schedule(100)
s = G.Solver(n_u, N, m, prob.L, dM, dG, mode=G.MODE_BATCH_PER_INSTANCE, max_batch=B, operators_mem=G.MEM_DEVICE)
dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
out = {k: torch.empty((B, m if k in ("y_next", "y", "w") else n), device="cuda") for k in ("y_next", "y", "z", "zhat", "w")}
st = torch.cuda.current_stream().cuda_stream
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve_device(B, dg, dp, theta, beta, 100, stream=st, **out); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print(s.description)
print(f"B={B}: {np.median(ts):.3f} ms per batch = {B / np.median(ts) * 1e3 / 1e6:.1f} M solves/s; |z| {float(out['z'].abs().max()):.6f}")
