"""Diagnostics for DESIGN.md (not a test): accuracy tables GPU / oracle / fp64 arbiter, 3xTF32
error growth with K, first timings.  Run on the GPU box:  python tests/diag_gpu.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    sys.path.insert(0, p)
import torch
import gpad_b200 as G
import problems as P
from oracle import Oracle, schedule

VECS = ("y_next", "y", "z", "zhat", "w")
o = Oracle()
theta, beta = schedule(100)


def row(label, gpu, ora, f64):
    s = f"{label:34s}"
    for k in VECS:
        s += f" | {k}: g-o {P.rel_inf(gpu[k], ora[k]):.1e} g-64 {P.rel_inf(gpu[k], f64[k]):.1e} o-64 {P.rel_inf(ora[k], f64[k]):.1e}"
    flips = int(((gpu["y_next"] > 0) != (ora["y_next"] > 0)).sum())
    print(s, "| flips", flips, flush=True)


print("== latency mode, 100 iterations")
for n_u, N in [(3, 4), (4, 3), (10, 15), (15, 10), (30, 30), (10, 100)]:
    pb = P.battery(n_u, N)
    g_P, p_D, f = pb.instance(P.battery_x0(n_u, np.random.default_rng(0)))
    s = G.Solver(n_u, N, pb.m, pb.L, pb.M_G, pb.G_L, mode=G.MODE_LATENCY)
    gpu = s.solve_host(g_P, p_D, theta, beta)
    ora = o.solve(n_u, N, pb.m, pb.M_G, pb.G_L, g_P, p_D, theta, beta)
    f64 = o.solve_f64(n_u, N, pb.m, pb.M_G, pb.G_L, g_P, p_D, theta, beta)
    row(f"latency ({n_u},{N})", gpu, ora, f64)
    # p50 latency, device-resident
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    dg, dp = dev(g_P), dev(p_D)
    outs = {k: torch.empty(pb.m if k in ("y_next", "y", "w") else pb.n, device="cuda") for k in VECS}
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        s.solve_device(1, dg, dp, theta, beta, 100, stream=st, **outs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve_device(1, dg, dp, theta, beta, 100, stream=st, **outs); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"    {s.description}\n    p50 {np.median(ts):.1f} us / solve = {np.median(ts)/100:.2f} us / iteration", flush=True)
    s.close()

print("== batch mode, 100 iterations")
for name, pb, n_u, N, par in [("battery(10,15)", P.battery(10, 15), 10, 15, None), ("quadrotor N=20", P.quadrotor(20), 4, 20, 1),
                              ("quadrotor N=100", P.quadrotor(100), 4, 100, 1)]:
    B = 130
    rng = np.random.default_rng(1)
    g_P, p_D, _ = pb.instance(rng.random((B, n_u)) - 0.5 if par is None else P.quadrotor_params(B, rng))
    nref = 6
    ora = o.solve_batch(n_u, N, pb.m, pb.M_G, pb.G_L, g_P[:nref], p_D[:nref], theta, beta)
    f64 = {k: np.stack([o.solve_f64(n_u, N, pb.m, pb.M_G, pb.G_L, g_P[b], p_D[b], theta, beta)[k] for b in range(nref)]) for k in VECS}
    for prec, code in (("fp32", G.PREC_FP32), ("tf32x3", G.PREC_TF32X3)):
        s = G.Solver(n_u, N, pb.m, pb.L, pb.M_G, pb.G_L, mode=G.MODE_BATCH_SHARED, precision=code, max_batch=B)
        gpu = s.solve_host(g_P, p_D, theta, beta)
        row(f"batch {prec} {name}", {k: gpu[k][:nref] for k in VECS}, ora, f64)
        s.close()

print("== 3xTF32 GEMM error vs K (relative to sum|a||b|), M=256 N=256")
for K in (64, 256, 1024, 2400, 4096, 8192):
    rng = np.random.default_rng(K)
    A = rng.standard_normal((256, K)).astype(np.float32); Bm = rng.standard_normal((256, K)).astype(np.float32)
    dC = torch.empty((256, 256), device="cuda")
    G.debug_gemm_tf32x3(torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda(), dC, 256, 256, K)
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ Bm.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(Bm).astype(np.float64).T
    C = dC.cpu().numpy()
    serial = np.zeros((256, 256), np.float32)
    e = (C - ref)
    print(f"   K={K:5d}  max {np.max(np.abs(e)/scale):.2e}  mean-signed(err*sign(ref))/|ref|-scale {np.mean(e*np.sign(ref))/np.mean(np.abs(ref)):.2e}"
          f"  rel-to-|ref| rms {np.sqrt(np.mean(e**2))/np.sqrt(np.mean(ref**2)):.2e}  torch-fp32 {np.max(np.abs((torch.from_numpy(A).cuda() @ torch.from_numpy(Bm).cuda().T).cpu().numpy() - ref)/scale):.2e}", flush=True)

print("== throughput timing, quadrotor N=100 (n=400, m=2400)")
pbq = G.Problem("quadrotor", N=100)
M_G, G_L = pbq.operators()
for prec, code, B, iters in (("tf32x3", G.PREC_TF32X3, 16384, 10), ("fp32", G.PREC_FP32, 4096, 5)):
    par = P.quadrotor_params(B, np.random.default_rng(2))
    g_P, p_D, _ = pbq.instances(par, want_f=False)
    s = G.Solver(4, 100, pbq.m, pbq.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=code, max_batch=B)
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    dz = torch.empty((B, 400), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    s.solve_device(B, dg, dp, theta, beta, 2, stream=st, z=dz); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve_device(B, dg, dp, theta, beta, iters, stream=st, z=dz); e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 4.0 * 400 * 2400 * B * iters
    print(f"   {prec}: B={B} {iters} iterations {ms:.1f} ms -> {ms/iters:.2f} ms/iter, {fl/ms/1e9:.1f} algorithmic TFLOP/s, "
          f"{B/(ms/iters*100/1e3):.0f} solves/s @100 it", flush=True)
    s.close()
