"""Small driver for ncu captures of the throughput-mode kernels (not a test):
    python tests/prof_batch.py [B] [iters] [precision]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    sys.path.insert(0, p)
import torch
import gpad_b200 as G
from bench import quad_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = {"fp32": G.PREC_FP32, "tf32x3": G.PREC_TF32X3, "fp16x3": G.PREC_FP16X3}[sys.argv[3] if len(sys.argv) > 3 else "fp16x3"]
prob = G.Problem("quadrotor", N=100)
M_G, G_L = prob.operators()
g_P, p_D, _ = prob.instances(quad_params(B, 0), want_f=False)
theta, beta = G.schedule(100)
s = G.Solver(4, 100, prob.m, prob.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=prec, max_batch=B)
dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
dz = torch.empty((B, prob.n), device="cuda")
st = torch.cuda.current_stream().cuda_stream
s.profile(True)
s.solve_device(B, dg, dp, theta, beta, iters, stream=st, z=dz)
torch.cuda.synchronize()
m0, c0 = s.profile_read(0); m1, c1 = s.profile_read(1); m2, c2 = s.profile_read(2)
print(f"B={B} iters={iters}: product1 {m1/c1:.3f} ms, zhat quantisation {m0/max(c0, 1):.3f} ms, product2 {m2/c2:.3f} ms per launch; "
      f"z finite {bool(torch.isfinite(dz).all())}")
