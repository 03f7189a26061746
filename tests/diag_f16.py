"""GPU diagnostics for GPAD_PREC_FP16X3 (not a test): GEMM hook errors, solver agreement with the tf32 / fp32 paths,
kernel timings of the 64K quadrotor batch for both tensor-core families.
    python tests/diag_f16.py [gemm] [solve] [time]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gpad_b200 as G
from bench import quad_params

sections = sys.argv[1:] or ["gemm", "solve", "time"]
B, ITERS = int(os.environ.get("DIAG_B", "65536")), 20

if "gemm" in sections:
    for kernel in (0, 1):
        for (M, N, K) in [(128, 16, 32), (128, 208, 64), (300, 200, 1000), (1000, 416, 2400), (257, 2400, 400)]:
            rng = np.random.default_rng(M + N + K)
            A = (rng.standard_normal((M, K)) * 10.0 ** rng.uniform(-6, 6, (M, 1))).astype(np.float32)
            Bm = (rng.standard_normal((N, K)) * 10.0 ** rng.uniform(-5, 3, (N, 1))).astype(np.float32)
            dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda()
            dC = torch.full((M, N), float("nan"), device="cuda")
            try:
                G.debug_gemm_f16x3(dA, dB, dC, M, N, K, kernel)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"gemm kernel {kernel} {(M, N, K)}: FAILED {e}", flush=True)
                continue
            C = dC.cpu().numpy()
            ref = A.astype(np.float64) @ Bm.astype(np.float64).T
            scale = np.abs(A).astype(np.float64) @ np.abs(Bm).astype(np.float64).T
            err = np.abs(C - ref) / scale
            print(f"gemm kernel {kernel} {(M, N, K)}: max err {np.nanmax(err):.3e} of sum|a||b| (bound {1e-6 + 1e-9 * K:.2e}), "
                  f"nan {int(np.isnan(C).sum())}, argmax {np.unravel_index(np.nanargmax(err), err.shape)}", flush=True)

if "solve" in sections:
    from problems import quadrotor, quadrotor_params, rel_inf
    N, Bs = 20, 300
    pb = quadrotor(N)
    g_P, p_D, _ = pb.instance(quadrotor_params(Bs, np.random.default_rng(11)))
    theta, beta = G.schedule(60)
    res = {}
    for name, code in (("fp32", G.PREC_FP32), ("tf32x3", G.PREC_TF32X3), ("fp16x3", G.PREC_FP16X3)):
        s = G.Solver(4, N, pb.m, pb.L, pb.M_G, pb.G_L, mode=G.MODE_BATCH_SHARED, precision=code, max_batch=Bs)
        res[name] = s.solve_host(g_P, p_D, theta, beta)
        if name == "fp16x3":
            print(s.description)
            warm = s.solve_host(g_P, p_D, theta, beta, y0=res[name]["y_next"], y_prev0=res[name]["y"])
        s.close()
    s = G.Solver(4, N, pb.m, pb.L, pb.M_G, pb.G_L, mode=G.MODE_BATCH_SHARED, precision=G.PREC_FP32, max_batch=Bs)
    warm32 = s.solve_host(g_P, p_D, theta, beta, y0=res["fp16x3"]["y_next"], y_prev0=res["fp16x3"]["y"])
    s.close()
    for k in ("y_next", "y", "z", "zhat", "w"):
        print(f"solve {k:7s}: fp16x3 vs fp32 {rel_inf(res['fp16x3'][k], res['fp32'][k]):.2e}   tf32x3 vs fp32 {rel_inf(res['tf32x3'][k], res['fp32'][k]):.2e}"
              f"   warm fp16x3 vs fp32 {rel_inf(warm[k], warm32[k]):.2e}", flush=True)

if "time" in sections:
    prob = G.Problem("quadrotor", N=100)
    M_G, G_L = prob.operators()
    par = quad_params(B, 0)
    g_P, p_D, _ = prob.instances(par, want_f=False)
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    dz = torch.empty((B, prob.n), device="cuda")
    theta, beta = G.schedule(ITERS)
    st = torch.cuda.current_stream().cuda_stream
    outs = {}
    variants = [("tf32x3", G.PREC_TF32X3, ""), ("fp16x3", G.PREC_FP16X3, "")] + \
        [("fp16x3 " + k, G.PREC_FP16X3, k) for k in os.environ.get("DIAG_KNOBS", "tc_pdl=1").split(";") if k]
    for name, code, knobs in variants:
        if knobs:
            os.environ["GPAD_DEBUG"] = knobs
        else:
            os.environ.pop("GPAD_DEBUG", None)
        s = G.Solver(4, 100, prob.m, prob.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=code, max_batch=B)
        for _ in range(2):
            s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
        e1.record(); e1.synchronize()
        per_iter = e0.elapsed_time(e1) / (3 * ITERS)
        outs[name] = dz.cpu().numpy().copy()
        s.profile(True)
        for w in range(3):
            s.profile_read(w)
        for _ in range(2):
            s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
        torch.cuda.synchronize()
        (m0, c0), (m1, c1), (m2, c2) = s.profile_read(0), s.profile_read(1), s.profile_read(2)
        s.profile(False)
        print(f"{name}: ms/iteration {per_iter:.4f}  product1 {m1 / max(c1, 1):.4f}  zsplit {m0 / max(c0, 1):.4f}  product2 {m2 / max(c2, 1):.4f}", flush=True)
        print("   ", s.description, flush=True)
        s.close()
    for name in outs:
        d = np.abs(outs[name] - outs["tf32x3"]).max() / np.abs(outs["tf32x3"]).max()
        print(f"z after {ITERS} iterations, {name} vs tf32x3 rel_inf {d:.2e}", flush=True)

if "small" in sections:
    # strong-scaling shards: one handle sized for 64K instances solving 8K .. 64K, with and without programmatic dependent launch
    prob = G.Problem("quadrotor", N=100)
    M_G, G_L = prob.operators()
    par = quad_params(65536, 0)
    g_P, p_D, _ = prob.instances(par, want_f=False)
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    dz = torch.empty((65536, prob.n), device="cuda")
    theta, beta = G.schedule(100)
    st = torch.cuda.current_stream().cuda_stream
    for knobs in os.environ.get("DIAG_KNOBS", ";tc_pdl=1").split(";"):
        if knobs:
            os.environ["GPAD_DEBUG"] = knobs
        else:
            os.environ.pop("GPAD_DEBUG", None)
        s = G.Solver(4, 100, prob.m, prob.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=G.PREC_FP16X3, max_batch=65536)
        base = None
        for b in (65536, 32768, 16384, 8192, 4096):
            for _ in range(2):
                s.solve_device(b, dg, dp, theta, beta, 100, stream=st, z=dz)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                s.solve_device(b, dg, dp, theta, beta, 100, stream=st, z=dz)
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / 3
            base = base or ms / b
            s.profile(True)
            for w in range(3):
                s.profile_read(w)
            s.solve_device(b, dg, dp, theta, beta, 100, stream=st, z=dz)
            torch.cuda.synchronize()
            (m0, c0), (m1, c1), (m2, c2) = s.profile_read(0), s.profile_read(1), s.profile_read(2)
            s.profile(False)
            print(f"[{knobs or 'default'}] B={b:6d}: {ms:8.3f} ms per solve, {ms / 100:.4f} ms/iteration, per-instance efficiency {base / (ms / b):.3f}"
                  f"   event-timed: product1 {m1 / max(c1, 1):.4f} quantise {m0 / max(c0, 1):.4f} product2 {m2 / max(c2, 1):.4f}", flush=True)
        s.close()
