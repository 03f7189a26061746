#!/usr/bin/env python
"""bench.py -- GPAD QP solves/sec at batch 64K (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3]): quadrotor-style MPC, nx=12, nu=4, horizon N=100 (n=400 decision
variables, m=2400 box + polytopic constraints), a batch of 65536 independent QPs per GPU sharing
M_G / G_L, random initial states and set-points (seeded), fixed 100 GPAD iterations per solve
(main.cu:87).  A STEP is one whole solve of the batch: 100 iterations of the hot path
(main.cu:160-175 loop body) through gpad_solve() in throughput mode (tcgen05 3xTF32).

  value     device-resident: inputs already in HBM, CUDA events around K gpad_solve() calls on the
            launching stream, max over ranks;  solves/s = N_gpus * 65536 * K / T
  e2e       the same solves through the C ABI with HOST (pinned) buffers, every step: the step's inputs (the
            per-instance parameters [x0; xref], from which g_P / p_D are built on the device exactly as the host
            build does) go host -> device and the five vectors main.cu:176-180 copies back come device -> host,
            inside the timed region, through gpad_solve_async / gpad_wait (copies of neighbouring steps overlap
            the iterations).  e2e.variants also times host g_P / p_D inputs (async and synchronous).
  parity_sample  instances of the timed 64K run itself (first / last tile, tile boundaries, spread) against
            oracle/_ref (the reference's compiled seq_functions.cpp) with the parity bound of the tests
  roofline  dominant kernel = the tcgen05 3xTF32 GEMM (both products): algorithmic FLOPs per launch 2*n*m*B over the
            mean launch duration measured with CUDA events inside the library (gpad_profile_*), against the tensor
            peak for this precision scheme: MEASURED_PEAKS.json bf16 sustained / 2 (tf32) / 3 (three MMAs/product)
  cpu_baseline  the reference's own seq_functions.cpp (oracle/_ref, kind "reference") or the C restatement
            (kind "port") on all host cores, bounded sample of the same workload
  further legs (rank 0, after the timed region): tolerance mode at 64K, strong scaling of a fixed 64K batch,
            single-QP latency with the reference's CPU and GPU loops beside it, per-instance plants (config 5),
            4096 battery QPs (config 3), closed loop
  --impl reference  times the reference CPU path as the whole arm (no GPU code, no repo library involved)
Inputs (2.9 GB per GPU per solve with state) are far larger than the 126 MB L2, so no flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gpu-dualgradient-mpc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "GPAD QP solves/sec at batch 64K"
UNIT = "solves/s"
BATCH = 65536
HORIZON = 100
ITERS = 100           # N_v, main.cu:87
WORKLOAD = ("quadrotor MPC nx=12 nu=4 N=100: n=400, m=2400 (box + polytopic), 65536 QPs per GPU sharing M_G/G_L, "
            "random x0/set-points (seed 0), fixed 100 GPAD iterations per solve")


def quad_params(B, seed):
    """per-instance [x0 (12); xref (12)], same distribution as tests/problems.py:quadrotor_params"""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, 12)); xr = np.zeros((B, 12))
    x0[:, 0:3] = rng.uniform(-1.0, 1.0, (B, 3)); x0[:, 3:6] = rng.uniform(-1.0, 1.0, (B, 3))
    x0[:, 6:8] = rng.uniform(-0.2, 0.2, (B, 2)); x0[:, 8] = rng.uniform(-0.5, 0.5, B)
    x0[:, 9:12] = rng.uniform(-0.5, 0.5, (B, 3))
    xr[:, 0:3] = rng.uniform(-2.0, 2.0, (B, 3)); xr[:, 8] = rng.uniform(-0.5, 0.5, B)
    return np.hstack([x0, xr])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                smax = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    sm.append(float(f[0]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def round_up32(v):
    return (v + 31) // 32 * 32


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1388.8), d.get("hbm_gbs", 6545.3), "MEASURED_PEAKS.json"
    return 1400.0, 6650.0, "B200_PROFILING.md fallback"


# ------------------------------------------------------------------------------------ CPU arm
def cpu_solver():
    """(callable, kind): batch solve on all host cores through the reference's own compiled
    seq_functions.cpp when oracle/_ref exists, else the C restatement"""
    import oracle
    oracle.build()
    if oracle.have_ref():
        ref = oracle.RefLib()
        return (lambda pb, g, p, th, be, nt: ref.solve_batch(pb["n_u"], pb["N"], pb["m"], pb["M_G"], pb["G_L"], g, p, th, be,
                                                             nthreads=nt)), "reference"
    ora = oracle.Oracle()
    return (lambda pb, g, p, th, be, nt: ora.solve_batch(pb["n_u"], pb["N"], pb["m"], pb["M_G"], pb["G_L"], g, p, th, be,
                                                         nthreads=nt)), "port"


def host_problem():
    import gpad_b200 as G
    prob = G.Problem("quadrotor", N=HORIZON)
    M_G, G_L = prob.operators(G.LAYOUT_SEQUENTIAL)
    return prob, dict(n_u=prob.n_u, N=prob.N, m=prob.m, n=prob.n, L=prob.L, M_G=M_G, G_L=G_L)


def numpy_problem():
    """the same quadrotor problem from the numpy restatement (tests/problems.py): the reference arm loads no repo library"""
    import problems as P
    pb = P.quadrotor(HORIZON)
    return pb, dict(n_u=pb.n_u, N=pb.N, m=pb.m, n=pb.n, L=pb.L, M_G=pb.M_G, G_L=pb.G_L)


def numpy_schedule(count):
    from oracle import schedule
    return schedule(count)


def cpu_rate(budget_s, cores=None):
    """solves/s of the CPU path on `cores` threads over a sample sized to ~budget_s seconds"""
    import gpad_b200 as G
    solve, kind = cpu_solver()
    prob, pb = host_problem()
    theta, beta = G.schedule(ITERS)
    cores = cores or os.cpu_count() or 1
    g1, p1, _ = prob.instances(quad_params(cores, 99), want_f=False)
    t0 = time.perf_counter(); solve(pb, g1, p1, theta, beta, cores); t_probe = time.perf_counter() - t0   # one solve per core
    count = int(max(cores, min(65536, budget_s / max(t_probe, 1e-3) * cores)))
    g, p, _ = prob.instances(quad_params(count, 7), want_f=False)
    t0 = time.perf_counter(); out = solve(pb, g, p, theta, beta, cores); dt = time.perf_counter() - t0
    return count / dt, dict(kind=kind, cores=int(out["threads"]),
                            sample=f"{count} of the workload's QPs, 100 iterations each, {dt:.1f} s on {out['threads']} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    solve, kind = cpu_solver()
    prob, pb = numpy_problem()
    theta, beta = numpy_schedule(ITERS)
    cores = os.cpu_count() or 1
    g1, p1, _ = prob.instance(quad_params(cores, 99))
    t0 = time.perf_counter(); solve(pb, g1, p1, theta, beta, cores); t_probe = time.perf_counter() - t0
    total_steps = args.steps + args.warmup
    per_step_s = max(1.0, min(15.0, 150.0 / max(total_steps, 1)))
    count = int(max(cores, per_step_s / max(t_probe, 1e-3) * cores))
    g, p, _ = prob.instance(quad_params(count, 7))
    for _ in range(args.warmup):
        solve(pb, g, p, theta, beta, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = solve(pb, g, p, theta, beta, cores)
    dt = time.perf_counter() - t0
    value = count * args.steps / dt
    sample = f"{count} of the workload's QPs per step, 100 iterations each, on {out['threads']} host threads"
    emit_line({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": "bounded sample: " + sample,
                   "inputs": "tests/problems.py numpy restatement of the quadrotor problem (no repo library loaded)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(out["threads"]), "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------ GPU arm
def p50(xs):
    return float(np.median(np.asarray(xs, np.float64)))


def latency_probe(torch, G):
    """single-QP p50 solve latency (second half of the BASELINE metric), latency mode, with the reference beside it:
    (i) the reference's CPU step functions on one core (oracle/_ref), (ii) the reference's own GPU kernels driven by
    the loop of main.cu:160-175 (5 launches + 3 device syncs per iteration) on this same B200 (oracle/_ref, compiled
    unmodified for sm_100a), (iii) our p50 through the C ABI with host buffers, (iv) tolerance mode, (v) a model"""
    import oracle
    ref_cpu = oracle.RefLib() if oracle.have_ref() else None
    ref_gpu = oracle.RefCuda() if oracle.have_refcuda() else None
    port = oracle.Oracle()
    out = {}
    sms, clk_ghz = 148, 1.9
    for n_u, N in ((3, 4), (10, 15), (10, 100)):
        prob = G.Problem("battery", n_u=n_u, N=N)
        M_G, G_L = prob.operators()
        n, m = prob.n, prob.m
        x0 = np.array([[-0.1, 0.45, -0.09, 0.05, 0, -0.05, 0.3, 0.2, 0.25, -0.45]]) if n_u == 10 else np.array([[0.31, -0.12, 0.44]])
        g_P, p_D, f = prob.instances(x0, want_f=True)
        theta, beta = G.schedule(ITERS)
        s = G.Solver(n_u, N, m, prob.L, M_G, G_L, mode=G.MODE_LATENCY)
        dg, dp = torch.from_numpy(g_P[0]).cuda(), torch.from_numpy(p_D[0]).cuda()
        dz = torch.empty(n, device="cuda"); dy = torch.empty(m, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(10):
            s.solve_device(1, dg, dp, theta, beta, ITERS, stream=st, z=dz, y_next=dy)
        torch.cuda.synchronize()
        ts = []
        for _ in range(300):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve_device(1, dg, dp, theta, beta, ITERS, stream=st, z=dz, y_next=dy); e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        entry = {"p50_us": p50(ts), "p99_us": float(np.percentile(ts, 99)), "iterations": ITERS, "path": s.description,
                 "us_per_iteration": p50(ts) / ITERS}
        z_fixed = dz.cpu().numpy().copy()             # z_99 of the fixed 100-iteration solve
        # (iii) through the C ABI with host buffers: H2D of g_P / p_D, the solve, D2H of the five vectors, synchronised
        th = []
        for _ in range(200):
            t0 = time.perf_counter(); s.solve_host(g_P[0], p_D[0], theta, beta); th.append((time.perf_counter() - t0) * 1e6)
        entry["p50_us_host_buffers_c_abi"] = p50(th[20:])
        # (iv) tolerance mode (BASELINE config 2): eps_g = eps_V = 1e-3, checked every 5 iterations, N_max guard
        nmax, every = 2000, 5
        th2, be2 = G.schedule(nmax)
        dit = torch.zeros(1, dtype=torch.int32, device="cuda"); dst = torch.zeros(1, dtype=torch.int32, device="cuda")
        tt = []
        for _ in range(60):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.solve_device(1, dg, dp, th2, be2, nmax, stream=st, z=dz, y_next=dy, iters=dit, status=dst, check_every=every,
                                        eps_g=1e-3, eps_V=1e-3); e1.record()
            e1.synchronize()
            tt.append(e0.elapsed_time(e1) * 1e3)
        entry["tolerance_1e-3"] = {"p50_us": p50(tt[10:]), "iterations": int(dit.item()), "status": G.STATUS_NAMES.get(int(dst.item())),
                                   "n_max": nmax, "check_every": every}
        if n * m < 200000:          # the CPU oracle on the same solve (seconds for the small problems only)
            ora = port.solve(n_u, N, m, M_G, G_L, g_P[0], p_D[0], th2, be2, L=prob.L, check_every=every, eps_g=1e-3, eps_V=1e-3)
            entry["tolerance_1e-3"].update({"oracle_iterations": int(ora["iters"]), "oracle_status": G.STATUS_NAMES.get(int(ora["status"]))})
        # (i) reference CPU loop, one core
        if ref_cpu is not None:
            reps = 200 if n * m < 50000 else (20 if n * m < 500000 else 3)
            tc = []
            for _ in range(reps):
                t0 = time.perf_counter(); ref_cpu.solve(n_u, N, m, M_G, G_L, g_P[0], p_D[0], theta, beta); tc.append((time.perf_counter() - t0) * 1e6)
            entry["reference_cpu_one_core_p50_us"] = p50(tc)
        # (ii) reference GPU loop on this B200 (operators resident, the loop of main.cu:160-175 alone)
        if ref_gpu is not None:
            tg, zref = ref_gpu.loop_times(n_u, N, m, M_G, G_L, g_P[0], p_D[0], theta, beta, ITERS, 40)
            entry["reference_cuda_loop_p50_us"] = p50(tg[5:])
            entry["reference_cuda_loop_matches"] = bool(np.max(np.abs(zref - z_fixed)) <= 2e-5 * max(1e-3, float(np.abs(zref).max())))
            entry["speedup_vs_reference_cuda_loop"] = entry["reference_cuda_loop_p50_us"] / entry["p50_us"]
        # (v) SM-cycle model of SURVEY 8(d): T_iter >= operator bytes through the shared-memory ports of the C
        # cooperating SMs + 2 exchanges (+ the dependent reduction chain for one warp)
        if "one warp" in s.description.lower():
            model = 300.0 / (clk_ghz * 1e3)
            basis = "dependent chain of one warp: 5 reduction levels x (select + shuffle + add) + broadcast + 16 chained FMAs + projection ~ 300 clk"
        elif "cluster" in s.description:
            C = 16
            model = 8.0 * n * m / (C * 128.0) / (clk_ghz * 1e3) + 2 * (215 + 38 + 60) / (clk_ghz * 1e3)
            basis = ("8nm B / (16 CTAs x 128 B/clk) + 2 exchanges x (DSMEM cross-CTA 215 clk + local 38 clk + mbarrier wait ~60 clk, "
                     "B300_MICROARCH.md CGA table) at 1.9 GHz")
        else:
            model = 4.0 * n * m / (sms * 128.0) / (clk_ghz * 1e3) + 2 * 1.1
            basis = "4nm B / (148 SMs x 128 B/clk) at 1.9 GHz + 2 grid exchanges x 1.1 us"
        entry.update({"model_us_per_iteration": model, "frac_of_model": model / entry["us_per_iteration"], "model": basis})
        out[f"battery({n_u},{N}) n={n} m={m}"] = entry
        s.close()
    return out


def per_instance_probe(torch, G, B=262144):
    """BASELINE config 5 kernel (per-instance plants, batched-GEMV mode): battery (3,4) QPs each with its own
    M_G / G_L; scored against the HBM roofline with the algorithmic minimum bytes per solve (operators stay
    on chip), against the fp32 FMA peak with the algorithmic flops, and against the algorithmic instruction minimum"""
    n_u, N = 3, 4
    prob = G.Problem("battery", n_u=n_u, N=N)
    M_G, G_L = prob.operators()
    n, m = prob.n, prob.m
    rng = np.random.default_rng(5)
    scale = (1.0 + 0.1 * (2 * rng.random((B, 1, 1)) - 1)).astype(np.float32)      # +-10 % per-instance operator perturbation
    dM = torch.from_numpy(M_G[None] * scale).cuda().contiguous(); dG = torch.from_numpy(G_L[None] / scale).cuda().contiguous()
    g_P, p_D, _ = prob.instances(rng.random((B, n_u)) - 0.5, want_f=False)
    theta, beta = G.schedule(ITERS)
    s = G.Solver(n_u, N, m, prob.L, dM, dG, mode=G.MODE_BATCH_PER_INSTANCE, max_batch=B, operators_mem=G.MEM_DEVICE)
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    out = {k: torch.empty((B, m if k in ("y_next", "y", "w") else n), device="cuda") for k in ("y_next", "y", "z", "zhat", "w")}
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, **out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, **out)
    e1.record(); e1.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    bytes_min = ((2 * n * m + n + m) * 4 + (3 * m + 2 * n) * 4) * B
    _, hbm, src = measured_peaks()
    fp32_peak = 148 * 128 * 2 * 1.965e9
    flops = 4.0 * n * m * ITERS * B
    # algorithmic minimum of warp instructions per QP-iteration: 2 n m / 32 = 42 FMAs; the two-QPs-per-warp kernel issues
    # 112.3 (profiles/r2_ncu_per_instance_warp2.csv: smsp__inst_executed / (B x 100); round 1's one-warp kernel: 174.4):
    # 56 products, 13.5 shuffles, the adds of the select-free transposed reduction, the projection and momentum updates
    inst_min, inst_issued = 2.0 * n * m / 32.0, 112.3
    res = {"workload": f"battery(3,4) n={n} m={m}, {B} QPs with per-instance operators, 100 iterations", "solves_per_s": B / sec,
           "ms_per_batch": sec * 1e3, "algorithmic_bytes_per_solve": bytes_min // B, "achieved_GBps": bytes_min / sec / 1e9,
           "hbm_peak_GBps": hbm, "frac_of_hbm_roofline": bytes_min / sec / 1e9 / hbm, "path": s.description,
           "achieved_TFLOPs": flops / sec / 1e12, "frac_of_fp32_fma_peak": flops / sec / fp32_peak,
           "warp_instructions_per_iteration": inst_issued, "algorithmic_warp_instructions_per_iteration": inst_min,
           "frac_of_algorithmic_instruction_minimum": inst_min / inst_issued,
           "issue_roofline_solves_per_s": 148 * 4 * 1.965e9 / (inst_issued * ITERS),
           "frac_of_issue_roofline": (B / sec) / (148 * 4 * 1.965e9 / (inst_issued * ITERS)),
           "note": "operators stay in registers for all 100 iterations, so HBM carries a few % of its peak; the kernel is bound by "
                   "instruction issue (ncu: issue slots 80 % busy), at 37 % of the algorithmic instruction minimum (round 1: 24 %)"}
    s.close()
    return res


def plants_probe(torch, G, world, rank, dist, B=131072, samples=5, iters=100):
    """BASELINE config 5 as written: a 1 M-instance scenario sweep with per-instance plant matrices (battery (3,4), cell
    capacities +-10 % per instance) sharded over the GPUs of the box -- 131 072 plants per GPU, so 8 GPUs hold the 1 M --
    T = 5 receding-horizon samples of 100 iterations, warm-started with the shifted duals of the previous sample.
    Every rank condenses and owns its shard; nothing is exchanged inside the loop."""
    n_u, N = 3, 4
    rng = np.random.default_rng(1000 + rank)
    t0 = time.perf_counter()
    plants = G.Plants(n_u, N, 1.0 + 0.1 * (2 * rng.random((B, n_u)) - 1))
    t_cond = time.perf_counter() - t0
    M, Gl, L = plants.operators()
    theta, beta = G.schedule(iters)
    s = G.Solver(n_u, N, plants.m, float(L[0]), M, Gl, mode=G.MODE_BATCH_PER_INSTANCE, max_batch=B)
    x0 = rng.random((B, n_u)) - 0.5
    plants.closed_loop(s, x0, 1, theta, beta, warm_start=G.WARM_SHIFTED)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    xt, ut = plants.closed_loop(s, x0, samples, theta, beta, warm_start=G.WARM_SHIFTED)
    sec = time.perf_counter() - t0
    if dist is not None:
        tmax = torch.tensor([sec], device="cuda"); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); sec = float(tmax.item())
    ok = bool(np.isfinite(xt).all() and np.isfinite(ut).all())
    u_max = float(np.abs(ut).max())
    desc = s.description
    s.close(); plants.close()
    return {"workload": f"battery(3,4) per-instance plants, {B} per GPU x {world} GPUs = {B * world} plants, {samples} receding-horizon samples x "
                        f"{iters} iterations, shifted warm start", "plants_total": B * world, "solves_per_s": B * world * samples / sec,
            "plant_steps_per_s": B * world * samples / sec, "ms_per_sample": sec / samples * 1e3, "host_condensing_s_per_gpu_shard": t_cond,
            "timed": "wall clock around gpad_closed_loop_plants (uploads of the shard's maps and the trajectory copy-back included), max over ranks",
            "finite": ok, "max_abs_input": u_max, "input_box": 0.3, "path": desc}


def battery_batch_probe(torch, G, B=4096):
    """BASELINE config 3: 4096 battery-balancing QPs (10,100) sharing M_G / G_L, random initial states, fixed 100
    iterations, tensor-core GEMM mode; algorithmic flops 4 n m per instance-iteration"""
    n_u, N = 10, 100
    prob = G.Problem("battery", n_u=n_u, N=N)
    M_G, G_L = prob.operators()
    n, m = prob.n, prob.m
    rng = np.random.default_rng(3)
    g_P, p_D, _ = prob.instances(rng.random((B, n_u)) - 0.5, want_f=False)
    theta, beta = G.schedule(ITERS)
    s = G.Solver(n_u, N, m, prob.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=G.PREC_FP16X3, max_batch=B)
    dg, dp = torch.from_numpy(g_P).cuda(), torch.from_numpy(p_D).cuda()
    dz = torch.empty((B, n), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        s.solve_device(B, dg, dp, theta, beta, ITERS, stream=st, z=dz)
    e1.record(); e1.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    res = {"workload": f"battery(10,100) n={n} m={m}, {B} QPs sharing M_G/G_L, 100 iterations", "solves_per_s": B / sec,
           "ms_per_batch": sec * 1e3, "achieved_TFLOPs": 4.0 * n * m * B * ITERS / sec / 1e12,
           "z_finite": bool(torch.isfinite(dz).all()), "path": s.description}
    s.close()
    return res


def closed_loop_probe(torch, G, B=16384, samples=5, iters=20):
    """warm-started receding-horizon steps of a quadrotor batch, everything on the device
    (instance build from the states, solve, state advance); only the trajectories cross PCIe"""
    prob = G.Problem("quadrotor", N=100)
    M_G, G_L = prob.operators()
    par = quad_params(B, 1)
    nx = prob.plant()[0].shape[0]
    x0, xref = np.ascontiguousarray(par[:, :nx]), np.ascontiguousarray(par[:, nx:])
    theta, beta = G.schedule(iters)
    s = G.Solver(4, 100, prob.m, prob.L, M_G, G_L, mode=G.MODE_BATCH_SHARED, precision=G.PREC_FP16X3, max_batch=B)
    G.closed_loop(prob, s, x0, 1, theta, beta, xref=xref, warm_start=G.WARM_SHIFTED)
    sec = float("inf")
    for _ in range(2):                       # best of two (each call also allocates and frees its device buffers)
        t0 = time.perf_counter()
        xt, ut = G.closed_loop(prob, s, x0, samples, theta, beta, xref=xref, warm_start=G.WARM_SHIFTED)
        sec = min(sec, time.perf_counter() - t0)
    s.close()
    return {"workload": f"quadrotor N=100, {B} plants, {samples} receding-horizon samples x {iters} iterations, shifted warm start",
            "plant_steps_per_s": B * samples / sec, "ms_per_sample": sec / samples * 1e3,
            "finite": bool(np.isfinite(xt).all() and np.isfinite(ut).all())}


def parity_sample(solver, prob, pb, params, d_out, theta, beta, B):
    """instances of the timed run itself against oracle/_ref (the reference's compiled seq_functions.cpp): first and
    last batch tile, both sides of tile boundaries, and a spread -- same bound as tests/test_gpu_parity.py:check_parity"""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    idx = sorted(set([0, 1, 127, 128, 129, 255, 256, B // 2 - 1, B // 2, B // 2 + 127, B - 257, B - 256, B - 129, B - 128, B - 127, B - 1]
                     + [int(v) for v in np.random.default_rng(17).integers(0, B, 8)]))
    idx = [i for i in idx if 0 <= i < B]
    g_P, p_D, _ = prob.instances(params[idx], want_f=False)
    ref = oracle.RefLib() if oracle.have_ref() else None
    port = oracle.Oracle()
    cores = os.cpu_count() or 1
    if ref is not None:
        ora = ref.solve_batch(pb["n_u"], pb["N"], pb["m"], pb["M_G"], pb["G_L"], g_P, p_D, theta, beta, nthreads=cores)
        kind = "reference"
    else:
        ora = port.solve_batch(pb["n_u"], pb["N"], pb["m"], pb["M_G"], pb["G_L"], g_P, p_D, theta, beta, nthreads=cores)
        kind = "port"
    with ThreadPoolExecutor(min(cores, len(idx))) as ex:
        f64 = list(ex.map(lambda j: port.solve_f64(pb["n_u"], pb["N"], pb["m"], pb["M_G"], pb["G_L"], g_P[j], p_D[j], theta, beta), range(len(idx))))
    names = ("y_next", "y", "z", "zhat", "w")
    sel = {k: d_out[k][idx].cpu().numpy() for k in names}
    rel = lambda a, b: float(np.max(np.abs(a.astype(np.float64) - b)) / max(float(np.max(np.abs(b))), 1e-300))
    worst, worst_over, flips, ok = 0.0, 0.0, 0, True
    for j in range(len(idx)):
        for k in names:
            e_or, noise = rel(sel[k][j], ora[k][j].astype(np.float64)), rel(ora[k][j], f64[j][k])
            tol = 2e-5 if k == "zhat" else 1e-5
            bound = max(tol, noise) + noise
            worst = max(worst, e_or); worst_over = max(worst_over, e_or / bound)
            ok = ok and e_or <= bound and rel(sel[k][j], f64[j][k]) <= 1.25 * noise + 1e-5
        fl = np.flatnonzero((sel["y_next"][j] > 0) != (ora["y_next"][j] > 0))
        flips += int(sum(1 for i in fl if abs(f64[j]["y_next"][i]) >= 1e-6))
    return {"against": f"oracle/_ref ({kind})", "instances": len(idx), "indices": idx, "worst_rel_inf": worst,
            "worst_fraction_of_bound": worst_over, "active_set_flips": flips, "within_test_bound": bool(ok and flips == 0),
            "bound": "GPU vs reference <= noise + max(1e-5 (2e-5 on zhat), noise), noise = reference vs fp64 arbiter; GPU vs arbiter <= 1.25 noise + 1e-5"}


def tolerance_probe(torch, G, prob, pb, solver, d_par, B, eps=1e-3, check_every=5, max_iter=4000):
    """BASELINE config 4's tolerance variant: the same 64K quadrotor batch with eps_g = eps_V = 1e-3, N_max = 4000;
    instances stop at very different iterations, the library retires finished tiles and compacts the running instances"""
    theta, beta = G.schedule(max_iter)
    d_it = torch.zeros(B, dtype=torch.int32, device="cuda"); d_st = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_z = torch.empty((B, prob.n), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    solver.solve_device(B, None, None, theta, beta, max_iter, stream=st, params=d_par, problem=prob, z=d_z, iters=d_it, status=d_st,
                        check_every=check_every, eps_g=eps, eps_V=eps)
    e1.record(); e1.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    stats = solver.stats()
    it = d_it.cpu().numpy(); stt = d_st.cpu().numpy()
    edges = [0, 100, 250, 500, 1000, 1500, 2000, 3000, max_iter + 1]
    hist = np.histogram(it, bins=edges)[0]
    launched = stats["scheduled"]
    return {"workload": f"the 64K quadrotor batch, eps_g = eps_V = {eps}, checked every {check_every} iterations, N_max = {max_iter}",
            "solves_per_s": B / sec, "ms_per_solve_of_the_batch": sec * 1e3,
            "iterations": {"min": int(it.min()), "median": float(np.median(it)), "mean": float(it.mean()), "p99": float(np.percentile(it, 99)),
                           "max": int(it.max()), "histogram_edges": edges, "histogram": [int(v) for v in hist]},
            "status_counts": {G.STATUS_NAMES[int(k)]: int((stt == k).sum()) for k in np.unique(stt)},
            "instance_iterations_needed": stats["needed"], "instance_iterations_scheduled": launched,
            "share_of_mma_work_on_finished_rows": 1.0 - stats["needed"] / max(launched, 1.0),
            "share_without_retirement_or_compaction": 1.0 - stats["needed"] / (float(round_up(B, 128)) * float(it.max())),
            "compactions": stats["compactions"], "z_finite": bool(torch.isfinite(d_z).all())}


def round_up(v, q):
    return (v + q - 1) // q * q


def bind_to_gpu_numa_node(local):
    """multi-GPU runs: run this rank (and so allocate its pinned host buffers) on the CPUs NVML names as local to its
    GPU, so that the 2.1 GB of results per step do not cross the socket interconnect; returns a note for the JSON line"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = ideal & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"rank bound to the {len(cpus)} CPUs local to its GPU (NVML affinity)"
        return "no narrower GPU-local CPU set available"
    except Exception as e:           # NVML absent / restricted container: run unbound
        return f"unbound ({type(e).__name__})"


def run_ours(args):
    import torch
    import gpad_b200 as G

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device and no CPU fallback exists for the product path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "single GPU: unbound"
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob, pb = host_problem()
    n, m, B = prob.n, prob.m, args.batch
    theta, beta = G.schedule(ITERS)
    from gpad_b200 import sharding
    params = quad_params(B, seed=sharding.shard_seed(0, rank))
    prec = {"fp32": G.PREC_FP32, "tf32x3": G.PREC_TF32X3, "fp16x3": G.PREC_FP16X3}[args.precision]
    solver = G.Solver(prob.n_u, prob.N, m, prob.L, pb["M_G"], pb["G_L"], mode=G.MODE_BATCH_SHARED, precision=prec,
                      max_batch=B, device=local)
    desc_main = solver.description
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    # device-resident inputs: g_P / p_D built once on the device from the parameters (bit-identical to the host build)
    d_par = torch.from_numpy(params).cuda()
    d_gP, d_pD = torch.empty((B, n), device="cuda"), torch.empty((B, m), device="cuda")
    G.instances_device(prob, B, d_par, d_gP, d_pD, stream=st)
    names = ("y_next", "y", "z", "zhat", "w")
    d_out = {k: torch.empty((B, m if k in ("y_next", "y", "w") else n), device="cuda") for k in names}
    d_it = torch.zeros(B, dtype=torch.int32, device="cuda"); d_st = torch.zeros(B, dtype=torch.int32, device="cuda")

    def step_device():
        solver.solve_device(B, d_gP, d_pD, theta, beta, ITERS, stream=st, iters=d_it, status=d_st, **d_out)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sharding.gather_first_moves(d_out["z"], prob.n_u, dst=0, equal_shards=True)   # NCCL connections set up outside the timed region
    barrier()
    solver.profile(True)
    solver.profile_read(0); solver.profile_read(1); solver.profile_read(2)
    launches0 = solver.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    # the only exchange of the path: final gather of the first control move u0 = z[:, :n_u] on rank 0
    u0_all = sharding.gather_first_moves(d_out["z"], prob.n_u, dst=0, equal_shards=True)
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    elapsed_ms = sharding.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    assert rank != 0 or u0_all.shape[0] == world * B
    launches = solver.launches - launches0
    ms0, c0 = solver.profile_read(0)          # fp16x3: the zhat row-quantisation kernel between the products
    ms1, c1 = solver.profile_read(1)
    ms2, c2 = solver.profile_read(2)
    solver.profile(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region, every step ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_par = pin(params)
    sets = [dict(out={k: torch.empty((B, m if k in ("y_next", "y", "w") else n), pin_memory=True).numpy() for k in names},
                 it=pin(np.zeros(B, np.int32)), st=pin(np.full(B, -1, np.int32))) for _ in range(2)]
    e2e_steps = max(2, min(args.steps, 10))

    def run_async(make_args):
        """K back-to-back steps, at most two in flight, results of step k land in buffer set k % 2"""
        argsets = [make_args(sets[j]) for j in range(2)]
        t = solver.solve_async(argsets[0]); solver.wait(t)                     # warm-up (allocates the second state set on step 2)
        t = solver.solve_async(argsets[1]); solver.wait(t)
        barrier()
        t0 = time.perf_counter()
        tickets = []
        for k in range(e2e_steps):
            if k >= 2:
                solver.wait(tickets[k - 2])            # the buffer set of step k - 2 is reused: its results must have landed
            tickets.append(solver.solve_async(argsets[k % 2]))
        for tk in tickets[-2:]:
            solver.wait(tk)
        torch.cuda.synchronize()
        return sharding.max_over_ranks(time.perf_counter() - t0, device="cuda")

    t_e2e = run_async(lambda S: G.host_args(B, theta, beta, ITERS, params=h_par, problem=prob, outputs=S["out"], iters=S["it"], status=S["st"]))
    last = sets[(e2e_steps - 1) % 2]
    ok = bool(np.isfinite(last["out"]["z"]).all() and (last["st"] == 0).all() and (last["it"] == ITERS).all())
    same = bool(all(np.array_equal(last["out"][k], d_out[k].cpu().numpy()) for k in ("z", "y_next")))
    h2d = h_par.nbytes
    d2h = sum(v.nbytes for v in last["out"].values()) + last["it"].nbytes + last["st"].nbytes
    variants = {}
    if rank == 0 or world > 1:
        # the same pipeline with g_P / p_D handed over from the host (734 MB more per step), and the synchronous call
        h_gP, h_pD = pin(d_gP.cpu().numpy()), pin(d_pD.cpu().numpy())
        t_v = run_async(lambda S: G.host_args(B, theta, beta, ITERS, g_P=h_gP, p_D=h_pD, outputs=S["out"], iters=S["it"], status=S["st"]))
        variants["async_host_gP_pD"] = {"value": world * B * e2e_steps / t_v, "h2d_bytes_per_step": int(h_gP.nbytes + h_pD.nbytes)}
        sync_args = G.host_args(B, theta, beta, ITERS, g_P=h_gP, p_D=h_pD, outputs=sets[0]["out"], iters=sets[0]["it"], status=sets[0]["st"])
        from gpad_b200 import check, lib
        import ctypes as C
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            check(lib().gpad_solve(solver._h, C.byref(sync_args)), "gpad_solve (sync e2e)")
        t_s = sharding.max_over_ranks((time.perf_counter() - t0) / 2, device="cuda")
        variants["synchronous_gpad_solve_host_gP_pD"] = {"value": world * B / t_s, "h2d_bytes_per_step": int(h_gP.nbytes + h_pD.nbytes)}
        del h_gP, h_pD
        # what a receding-horizon controller reads back: the primal solution z (its first n_u entries are the control move),
        # iteration counts and status -- the same pipeline with 105 MB instead of 2.1 GB of results per step
        t_z = run_async(lambda S: G.host_args(B, theta, beta, ITERS, params=h_par, problem=prob, outputs={"z": S["out"]["z"]},
                                              iters=S["it"], status=S["st"]))
        variants["async_params_outputs_z_only"] = {"value": world * B * e2e_steps / t_z, "h2d_bytes_per_step": int(h2d),
                                                   "d2h_bytes_per_step": int(last["out"]["z"].nbytes + last["it"].nbytes + last["st"].nbytes)}

    # ---- strong scaling of a FIXED 64K batch (BASELINE config 4 as written): every rank solves 65536 / N instances ----
    strong = None
    if world > 1:
        Bs = BATCH // world
        barrier()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = max(3, args.steps)
        for _ in range(2):
            solver.solve_device(Bs, d_gP, d_pD, theta, beta, ITERS, stream=st, z=d_out["z"])
        barrier()
        es0.record(stream)
        for _ in range(ks):
            solver.solve_device(Bs, d_gP, d_pD, theta, beta, ITERS, stream=st, z=d_out["z"])
        sharding.gather_first_moves(d_out["z"][:Bs], prob.n_u, dst=0, equal_shards=True)
        es1.record(stream)
        barrier()
        ms = sharding.max_over_ranks(es0.elapsed_time(es1), device="cuda")
        strong = {"total_batch": BATCH, "batch_per_gpu": Bs, "steps": ks, "solves_per_s": BATCH * ks / (ms * 1e-3), "ms_per_step": ms / ks,
                  "note": "fixed 65536-instance batch cut into N shards; compare with the N = 1 value of this metric for the strong-scaling factor"}

    plants = plants_probe(torch, G, world, rank, dist) if not args.no_latency else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (elapsed_ms * 1e-3)
    e2e_value = world * B * e2e_steps / t_e2e
    bf16_sus, hbm, peak_src = measured_peaks()
    peak = bf16_sus / 2.0 / 3.0 if prec == G.PREC_TF32X3 else None
    flops_per_launch = 2.0 * n * m * B
    k_ms = (ms1 + ms2) / max(c1 + c2, 1)
    achieved = flops_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
    traffic = None
    traffic_all = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic_all = json.load(open(tpath))
        traffic = traffic_all.get("dram_bytes_per_launch")
    if prec == G.PREC_FP16X3:
        # fp16 hi/lo operands double the tensor rate and halve the operand bytes: the iteration's arithmetic intensity
        # nm / (4m + 3n) = 89 flop/B now sits BELOW the ridge (3xFP16: 732 TF/s / 6.5 TB/s = 112 flop/B), so the
        # path is HBM bound.  Dominant kernel = product 2 (tc_p2_kernel): its share of the algorithmic bytes of SURVEY 8(d)
        # is the 4m floats per instance (read y_v, y_{v-1}, p_D, write y_{v+1}); product 1 carries the 3n (read g_P, r/w z).
        p1_ms, p2_ms, q_ms = ms1 / max(c1, 1), ms2 / max(c2, 1), ms0 / max(c0, 1)
        bytes_p2 = 4.0 * m * 4.0 * B
        bytes_iter = (4.0 * m + 3.0 * n) * 4.0 * B
        ach = bytes_p2 / (p2_ms * 1e-3) / 1e9 if p2_ms > 0 else None
        iter_ms = elapsed_ms / (args.steps * ITERS)
        tf_peak = bf16_sus / 3.0
        p1_tf = flops_per_launch / (p1_ms * 1e-3) / 1e12 if p1_ms > 0 else None
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": (ach / hbm) if ach else None,
                "traffic": (traffic_all or {}).get("product2_f16", {}).get("dram_bytes_per_launch"),
                "kernel": "tc_p2_kernel (product 2, tcgen05 kind::f16 x3, TMA-streamed epilogue)",
                "algorithmic_bytes_per_launch": bytes_p2, "mean_launch_ms": p2_ms, "launches_timed": int(c2),
                "kernel_share_of_step": ms2 / elapsed_ms,
                "peak_basis": f"{peak_src}: hbm_gbs (copy bandwidth, of measured)",
                "product1": {"bound": "tensor", "kernel": "tc_p1_kernel<8,false,true> (kind::f16 x3, A operand quantised into TMEM)",
                             "mean_launch_ms": p1_ms, "achieved": p1_tf, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": (p1_tf / tf_peak) if p1_tf else None, "algorithmic_flops_per_launch": flops_per_launch,
                             "tcgen05_f16x3_peak_measured": 732.0, "frac_of_tcgen05_peak": (p1_tf / 732.0) if p1_tf else None,
                             "kernel_share_of_step": ms1 / elapsed_ms,
                             "traffic": (traffic_all or {}).get("product1_f16", {}).get("dram_bytes_per_launch"),
                             "peak_basis": f"{peak_src}: bf16 sustained {bf16_sus} TF/s / 3 MMAs per product; the tensor pipe itself "
                                           "does 4096 fp16 MAC/clk/SM = 732 TF/s of 3xFP16-effective at 1.81 GHz (ubench, twice the tf32 rate)"},
                "zhat_quantize_kernel": {"mean_launch_ms": q_ms, "bytes_per_launch": 2.0 * round_up32(n) * 4.0 * B,
                                         "GBps": 2.0 * round_up32(n) * 4.0 * B / (q_ms * 1e-3) / 1e9 if q_ms > 0 else None,
                                         "kernel_share_of_step": ms0 / elapsed_ms},
                "whole_iteration": {"algorithmic_bytes": bytes_iter, "ms": iter_ms, "GBps": bytes_iter / (iter_ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": bytes_iter / (iter_ms * 1e-3) / 1e9 / hbm,
                                    "algorithmic_TFLOPs": 2.0 * flops_per_launch / (iter_ms * 1e-3) / 1e12,
                                    "note": "(4m + 3n) * 4 B per instance-iteration (SURVEY 8d) over the device-timed iteration: launch gaps, "
                                            "the quantisation kernel and product 1 included"}}
    elif prec == G.PREC_TF32X3:
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": "tc_p1_kernel (product 1) + tc_gemm_kernel<2> (product 2), tcgen05 kind::tf32 x3", "algorithmic_flops_per_launch": flops_per_launch,
                "mean_launch_ms": k_ms, "launches_timed": int(c1 + c2), "product1_ms": ms1 / max(c1, 1), "product2_ms": ms2 / max(c2, 1),
                "kernel_share_of_step": (ms1 + ms2) / elapsed_ms,
                # the tensor pipe itself, measured on this pool's B200 with MMAs only (tests/ubench/ubench_tc.cu,
                # profiles/r1_ubench_tc_issue_and_shapes.log): 2048 tf32 MAC/clk/SM -> 366 TF/s of 3xTF32-effective at ~1.81 GHz
                "tcgen05_tf32x3_peak_measured": 366.0, "frac_of_tcgen05_peak": (achieved / 366.0) if achieved else None,
                "peak_basis": f"{peak_src}: bf16 sustained {bf16_sus} TF/s / 2 (tf32) / 3 (hi*lo + lo*hi + hi*hi MMAs per product), of measured"}
    else:
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if achieved else None,
                "traffic": traffic, "kernel": "simt_gemm_kernel (CUDA-core FFMA)", "peak_basis": "148 SMs x 128 FMA/clk x 1.965 GHz (nominal fp32)"}

    psample = parity_sample(solver, prob, pb, params, d_out, theta, beta, B)
    # the other tensor-core family on the same box, same inputs (a tolerance-mode solve of this handle runs its kernels):
    # device-resident steps only, for the comparison of DESIGN.md section 4.1d
    alt = None
    if world == 1 and prec == G.PREC_FP16X3 and not args.no_latency:
        s2 = G.Solver(prob.n_u, prob.N, m, prob.L, pb["M_G"], pb["G_L"], mode=G.MODE_BATCH_SHARED, precision=G.PREC_TF32X3,
                      max_batch=B, device=local)
        d_z2 = torch.empty((B, n), device="cuda")
        for _ in range(2):
            s2.solve_device(B, d_gP, d_pD, theta, beta, ITERS, stream=st, z=d_z2)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(5):
            s2.solve_device(B, d_gP, d_pD, theta, beta, ITERS, stream=st, z=d_z2)
        a1.record(stream); a1.synchronize()
        ms_alt = a0.elapsed_time(a1) / 5
        dz = (d_z2 - d_out["z"]).abs().max().item() / d_out["z"].abs().max().item()
        alt = {"precision": "tf32x3", "value": B / (ms_alt * 1e-3), "unit": UNIT, "ms_per_step": ms_alt, "steps": 5,
               "z_rel_inf_vs_fp16x3": dz, "path": s2.description}
        s2.close()
        del d_z2
    tol = tolerance_probe(torch, G, prob, pb, solver, d_par, B) if (not args.no_latency and prec != G.PREC_FP32) else None
    solver.close()
    cpu_val, cpu_info = cpu_rate(args.cpu_budget)
    lat = latency_probe(torch, G) if not args.no_latency else None
    per_inst = per_instance_probe(torch, G) if not args.no_latency else None
    bat3 = battery_batch_probe(torch, G) if not args.no_latency else None
    cloop = closed_loop_probe(torch, G) if not args.no_latency else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {G.PREC_TF32X3: "f32 (tf32 x3 split products, fp32 accumulate)",
                  G.PREC_FP16X3: "f32 (fp32 state; products as 3 kind::f16 MMAs on fp16 hi/lo splits of power-of-two row-scaled operands, "
                                 "11 + 11 significant bits like the tf32 split, fp32 accumulate)"}.get(prec, "f32"), "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "iters_per_solve": ITERS, "n": n, "m": m,
                   "step": "one gpad_solve() of the whole batch = 100 GPAD iterations", "l2": "inputs (>2.9 GB/GPU) exceed the 126 MB L2; no flush",
                   "precision": args.precision, "path": desc_main, "host_placement": numa_note},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                "call": "gpad_solve_async / gpad_wait, pinned host buffers, two steps in flight; inputs = per-instance parameters [x0; xref] "
                        "(g_P / p_D built on the device), outputs = the five vectors of main.cu:176-180 + iters + status",
                "fraction_of_device_resident_value": e2e_value / value,
                "results_finite_and_complete": ok, "matches_device_resident_run": same, "variants": variants},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": dict(value=cpu_val, unit=UNIT, **cpu_info),
        "clocks": clocks,
        "parity_sample": psample,
        "same_box_tf32x3": alt,
        "quadrotor_64k_eps1e-3": tol,
        "strong_scaling": strong,
        "single_qp_latency": lat,
        "per_instance_operators": per_inst,
        "per_instance_plants_1M_sweep": plants,
        "battery_batch_4096": bat3,
        "closed_loop": cloop,
    }
    emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit_line(obj):
    """exactly one JSON line on the real stdout (library banners, e.g. NCCL's, go to stderr)"""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                     # anything printed by libraries from here on lands on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "tf32x3", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
