"""GPU diagnostics for round 2 (not a test): kernel timings of the 64K quadrotor batch under knob variants, the
round-1 library on the same box, the instance build, and the async pipeline breakdown.
    python tests/diag_r2.py [section ...]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gpad_b200 as G
from bench import quad_params

B, ITERS = int(os.environ.get("DIAG_B", "65536")), 20
sections = sys.argv[1:] or ["kernels", "build", "async"]


def setup(knobs=""):
    if knobs:
        os.environ["GPAD_DEBUG"] = knobs
    else:
        os.environ.pop("GPAD_DEBUG", None)
    prob = G.Problem("quadrotor", N=100)
    M_G, G_L = prob.operators()
    s = G.Solver(4, 100, prob.m, prob.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=G.PREC_TF32X3, max_batch=B)
    return prob, s


def time_kernels(label, knobs, lib_path=None):
    if lib_path:
        G._lib = None; G.LIB_PATH = lib_path
    prob, s = setup(knobs)
    par = quad_params(B, 0)
    g_P, p_D, _ = prob.instances(par, want_f=False)
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    dz = torch.empty((B, prob.n), device="cuda")
    theta, beta = G.schedule(ITERS)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
    e1.record(); e1.synchronize()
    per_iter = e0.elapsed_time(e1) / (3 * ITERS)
    s.profile(True); s.profile_read(1); s.profile_read(2)
    for _ in range(2):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
    torch.cuda.synchronize()
    m1, c1 = s.profile_read(1); m2, c2 = s.profile_read(2)
    s.profile(False)
    print(f"{label:34s} ms/iteration {per_iter:.4f}  product1 {m1 / max(c1, 1):.4f}  product2 {m2 / max(c2, 1):.4f}   [{knobs}]", flush=True)
    desc = s.description
    s.close()
    if lib_path:
        G._lib = None; G.LIB_PATH = os.path.join(ROOT, "gpu-dualgradient-mpc_b200", "lib", "libgpad_b200.so")
    return desc


if sections[0] == "one":            # child process: one configuration (two libraries in one process do not mix)
    time_kernels(sections[1], sections[2], sections[3] if len(sections) > 3 else None)
    sys.exit(0)

if "kernels" in sections:
    import subprocess
    r1 = os.path.join(ROOT, "build", "r1", "libgpad_r1.so")
    runs = []
    if os.path.exists(r1):
        runs += [("round-1 library, width 160", "", r1, {"GPAD_TC_BN2": "160"}), ("round-1 library, width 192", "", r1, {"GPAD_TC_BN2": "192"})]
    for knobs in ("tc_bn2=160", "tc_bn2=160,tc_cluster_attr=1", "tc_bn2=192", "tc_bn2=192,tc_cluster_attr=1", ""):
        runs.append(("this build", knobs, None, {}))
    for label, knobs, lib, env in runs:
        cmd = [sys.executable, __file__, "one", label, knobs] + ([lib] if lib else [])
        r = subprocess.run(cmd, env={**os.environ, **env}, capture_output=True, text=True)
        print(r.stdout.strip() or f"{label} [{knobs}]: rc={r.returncode} {r.stderr[-400:]}", flush=True)

if "small" in sections:
    import subprocess
    for b in (8192, 16384, 32768, 65536):
        for knobs in ("", "tc_pdl=1"):
            r = subprocess.run([sys.executable, __file__, "one", f"B={b}", knobs], env={**os.environ, "DIAG_B": str(b)}, capture_output=True, text=True)
            print(r.stdout.strip() or f"B={b} [{knobs}]: rc={r.returncode} {r.stderr[-400:]}", flush=True)

if "latency" in sections:
    for n_u, N in ((10, 100), (10, 15), (15, 10), (4, 6)):
        for knobs in ("latency_flat=1,flat_xchg=1", "latency_flat=1,flat_xchg=0", "latency_flat=0"):
            os.environ["GPAD_DEBUG"] = knobs
            prob = G.Problem("battery", n_u=n_u, N=N)
            M_G, G_L = prob.operators()
            g_P, p_D, _ = prob.instances(np.random.default_rng(1).random((1, n_u)) - 0.5, want_f=False)
            theta, beta = G.schedule(100)
            s = G.Solver(n_u, N, prob.m, prob.L, M_G, G_L, mode=G.MODE_LATENCY)
            dg, dp = torch.from_numpy(g_P[0]).cuda(), torch.from_numpy(p_D[0]).cuda()
            dz = torch.empty(prob.n, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(10):
                s.solve_device(1, dg, dp, theta, beta, 100, stream=st, z=dz)
            torch.cuda.synchronize()
            ts = []
            for _ in range(200):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.solve_device(1, dg, dp, theta, beta, 100, stream=st, z=dz); e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            th11, be11 = G.schedule(1100)
            s.solve_device(1, dg, dp, th11, be11, 1100, stream=st, z=dz); torch.cuda.synchronize()
            tl = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.solve_device(1, dg, dp, th11, be11, 1100, stream=st, z=dz); e1.record(); e1.synchronize()
                tl.append(e0.elapsed_time(e1) * 1e3)
            print(f"latency ({n_u},{N}) [{knobs}]: p50 {np.median(ts):.1f} us, slope {(np.median(tl) - np.median(ts)) / 1000:.3f} us/iteration   {s.description[:90]}", flush=True)
            s.close()

if "plants" in sections:
    n_u, N, Bp = 3, 4, 131072
    rng = np.random.default_rng(3)
    plants = G.Plants(n_u, N, 1.0 + 0.1 * (2 * rng.random((Bp, n_u)) - 1))
    M, Gl, L = plants.operators()
    theta, beta = G.schedule(100)
    s = G.Solver(n_u, N, plants.m, float(L[0]), M, Gl, mode=G.MODE_BATCH_PER_INSTANCE, max_batch=Bp)
    x0 = rng.random((Bp, n_u)) - 0.5
    for samples in (1, 5, 9):
        for warm in (0, 2):
            plants.closed_loop(s, x0, samples, theta, beta, warm_start=warm)
            t0 = time.perf_counter(); plants.closed_loop(s, x0, samples, theta, beta, warm_start=warm); dt = time.perf_counter() - t0
            print(f"plants closed loop: samples {samples} warm {warm}: {dt * 1e3:.1f} ms total, {dt / samples * 1e3:.2f} ms per sample", flush=True)
    s.close()

if "build" in sections:
    prob, s = setup("tc_bn2=160")
    par = quad_params(B, 0)
    dpar = torch.from_numpy(par).cuda()
    dg, dp = torch.empty((B, prob.n), device="cuda"), torch.empty((B, prob.m), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    G.instances_device(prob, B, dpar, dg, dp, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        G.instances_device(prob, B, dpar, dg, dp, stream=st)
    e1.record(); e1.synchronize()
    print(f"instances_device 64K: {e0.elapsed_time(e1) / 5:.3f} ms per build", flush=True)
    s.close()

if "async" in sections:
    prob, s = setup("tc_bn2=160")
    par = quad_params(B, 0)
    theta, beta = G.schedule(100)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_par = pin(par)
    g_P, p_D, _ = prob.instances(par, want_f=False)
    h_g, h_p = pin(g_P), pin(p_D)
    names = ("y_next", "y", "z", "zhat", "w")
    sets = [{k: torch.empty((B, prob.m if k in ("y_next", "y", "w") else prob.n), pin_memory=True).numpy() for k in names} for _ in range(2)]
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    dz = torch.empty((B, prob.n), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        s.solve_device(B, dg, dp, theta, beta, 100, stream=st, z=dz)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        s.solve_device(B, dg, dp, theta, beta, 100, stream=st, z=dz)
    torch.cuda.synchronize()
    print(f"device-resident: {(time.perf_counter() - t0) / 4 * 1e3:.1f} ms per step", flush=True)

    def run(label, mk, K=8):
        a = [mk(sets[j]) for j in range(2)]
        for j in range(2):
            s.wait(s.solve_async(a[j]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tk = []
        for k in range(K):
            if k >= 2:
                s.wait(tk[k - 2])
            tk.append(s.solve_async(a[k % 2]))
        for t in tk[-2:]:
            s.wait(t)
        torch.cuda.synchronize()
        print(f"async {label:48s}: {(time.perf_counter() - t0) / K * 1e3:.1f} ms per step", flush=True)

    sub = lambda S, keys: {k: S[k] for k in keys}
    run("params in, five vectors out", lambda S: G.host_args(B, theta, beta, 100, params=h_par, problem=prob, outputs=S))
    run("params in, z only out", lambda S: G.host_args(B, theta, beta, 100, params=h_par, problem=prob, outputs=sub(S, ("z",))))
    run("params in, y_next y w out (contiguous rows)", lambda S: G.host_args(B, theta, beta, 100, params=h_par, problem=prob, outputs=sub(S, ("y_next", "y", "w"))))
    run("params in, z zhat out (strided rows)", lambda S: G.host_args(B, theta, beta, 100, params=h_par, problem=prob, outputs=sub(S, ("z", "zhat"))))
    run("g_P p_D in, z only out", lambda S: G.host_args(B, theta, beta, 100, g_P=h_g, p_D=h_p, outputs=sub(S, ("z",))))
    run("g_P p_D in, five vectors out", lambda S: G.host_args(B, theta, beta, 100, g_P=h_g, p_D=h_p, outputs=S))
    s.close()
