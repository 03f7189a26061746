"""CPU tests (-m "not gpu") of the host side of the product: the C-ABI library loads and exports
every symbol include/gpad.h declares, the C++ condensing reproduces the numpy restatement of
gpad.m, the schedule matches the oracle's, the data-file reader/writer round-trips, and compute
entry points fail loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpad_b200 as G
import problems as P
from oracle import schedule as oracle_schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gpad.h")).read()
    declared = set(re.findall(r"\b(gpad_[a-z0-9_]+)\s*\(", header))
    declared -= {"gpad_status", "gpad_layout"}
    lib = C.CDLL(G.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(G.EXPORTS), declared ^ set(G.EXPORTS)
    assert G.lib().gpad_api_version() == 2


def test_schedule_matches_oracle():
    for variant, name in ((G.SCHEDULE_PAPER, "paper"), (G.SCHEDULE_MATLAB_LAG, "matlab_lag")):
        th, be = G.schedule(300, variant)
        oth, obe = oracle_schedule(300, name)
        assert np.array_equal(th, oth) and np.array_equal(be, obe)


@pytest.mark.parametrize("dims", [(3, 4), (4, 3), (10, 15), (5, 2)])
def test_battery_condensing_matches_numpy_restatement(dims):
    n_u, N = dims
    ref = P.battery(n_u, N)
    pb = G.Problem("battery", n_u=n_u, N=N)
    assert (pb.n_u, pb.N, pb.m, pb.n_par) == (n_u, N, 4 * n_u * N + 2 * N, n_u)
    assert abs(pb.L - ref.L) <= 1e-6 * ref.L
    M_G, G_L = pb.operators(G.LAYOUT_SEQUENTIAL)
    assert np.allclose(M_G, ref.M_G, rtol=2e-6, atol=1e-9)
    assert np.allclose(G_L, ref.G_L, rtol=2e-6, atol=0)
    Mf, Gf = pb.operators(G.LAYOUT_FLIPPED)
    assert np.array_equal(Mf, M_G.T) and np.array_equal(Gf, G_L.T)
    x0 = np.random.default_rng(3).random((4, n_u)) - 0.5
    g_P, p_D, f = pb.instances(x0)
    rg, rp, rf = ref.instance(x0)
    assert np.allclose(g_P, rg, rtol=2e-6, atol=1e-9)
    assert np.allclose(p_D, rp, rtol=2e-6, atol=0)
    assert np.allclose(f, rf, rtol=2e-6, atol=1e-9)
    A, Bm = pb.plant()
    assert np.array_equal(A, np.eye(n_u)) and np.allclose(np.diag(Bm), -1 / (3600 * 0.027 * 4.1))


def test_quadrotor_condensing_matches_numpy_restatement():
    ref = P.quadrotor(20)
    pb = G.Problem("quadrotor", N=20)
    assert (pb.n_u, pb.N, pb.n, pb.m, pb.n_par) == (4, 20, 80, 24 * 20, 24)
    assert abs(pb.L - ref.L) <= 1e-5 * ref.L
    M_G, G_L = pb.operators()
    scale = np.abs(ref.M_G).max()
    assert np.max(np.abs(M_G - ref.M_G)) <= 2e-5 * scale       # cond(H) is large: compare in the inf-norm
    assert np.allclose(G_L, ref.G_L, rtol=2e-5, atol=1e-9)
    par = P.quadrotor_params(3, np.random.default_rng(1))
    g_P, p_D, f = pb.instances(par)
    rg, rp, rf = ref.instance(par)
    assert np.max(np.abs(g_P - rg)) <= 2e-5 * np.abs(rg).max()
    assert np.allclose(p_D, rp, rtol=2e-5, atol=1e-9)
    assert np.max(np.abs(f - rf)) <= 2e-5 * np.abs(rf).max()


def test_full_size_quadrotor_dimensions():
    pb = G.Problem("quadrotor", N=100)
    assert (pb.n, pb.m) == (400, 2400)


def test_data_file_round_trip(tmp_path):
    """main.cu:29-67 format: written by us, parsed the way readData() parses it."""
    pb = G.Problem("battery", n_u=3, N=4)
    M_G, G_L = pb.operators(G.LAYOUT_FLIPPED)
    g_P, p_D, _ = pb.instances(np.array([0.1, -0.2, 0.3]))
    theta, beta = G.schedule(100)
    path = str(tmp_path / "input_1.txt")
    G.file_write(path, 3, 4, pb.m, pb.L, M_G, g_P, G_L, p_D, theta, beta)
    tok = open(path).read().split()
    assert [int(t) for t in tok[:4]] == [3, 4, 56, 100]                     # header "n_u N m num_iterations L"
    assert len(tok) == 5 + 2 * 12 * 56 + 12 + 56 + 200
    back = G.file_read(path)
    assert (back["n_u"], back["N"], back["m"], back["num_iterations"]) == (3, 4, 56, 100)
    assert np.array_equal(back["M_G"], M_G.ravel()) and np.array_equal(back["G_L"], G_L.ravel())
    assert np.array_equal(back["g_P"], g_P.ravel()) and np.array_equal(back["p_D"], p_D.ravel())
    assert np.array_equal(back["theta"], theta) and np.array_equal(back["beta"], beta)
    with pytest.raises(G.GpadError):
        G.file_read(str(tmp_path / "missing.txt"))
    open(str(tmp_path / "short.txt"), "w").write("3 4 56 100 12.0\n1.0 2.0\n")
    with pytest.raises(G.GpadError):
        G.file_read(str(tmp_path / "short.txt"))


def test_invalid_arguments_are_rejected_before_touching_the_device():
    M = np.zeros((12, 56), np.float32)
    with pytest.raises(G.GpadError, match="invalid"):
        G.Solver(0, 4, 56, 1.0, M, M.T)
    with pytest.raises(G.GpadError, match="invalid"):
        G.Solver(3, 4, 56, 1.0, M, M.T, mode=G.MODE_LATENCY, max_batch=4)
    with pytest.raises(G.GpadError, match="invalid"):
        G.Solver(3, 4, 56, 1.0, M, M.T, mode=G.MODE_LATENCY, precision=G.PREC_TF32X3)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert G.device_count() == 0
    M = np.zeros((12, 56), np.float32)
    with pytest.raises(G.GpadError, match="no usable CUDA device|no CUDA device"):
        G.Solver(3, 4, 56, 1.0, M, M.T)


@pytest.mark.parametrize("dims", [(3, 4), (10, 15)])
def test_flat_operators_round_trip_and_match_reference_flat_steps(dims):
    """ENABLE_FLATTEN_MATRICES data format (main.cu:39-41, seq_functions.cpp:5-43): the battery operators
    flatten without residual, expand back exactly, and the reference-restated flat step functions on the
    flat operators reproduce the dense step functions."""
    from oracle import Oracle
    n_u, N = dims
    pb = G.Problem("battery", n_u=n_u, N=N)
    M_G, G_L = pb.operators()
    Mf, Gf, resid = G.flatten_operators(n_u, N, pb.m, M_G, G_L)
    assert Mf.shape == (N, pb.m) and Gf.shape == (pb.m, N) and resid == 0.0
    M2, G2 = G.expand_operators(n_u, N, pb.m, Mf, Gf)
    assert np.array_equal(M2, M_G) and np.array_equal(G2, G_L)
    o = Oracle()
    rng = np.random.default_rng(1)
    w = rng.standard_normal(pb.m).astype(np.float32)
    g_P, p_D, _ = pb.instances(rng.random(n_u) - 0.5)
    z_dense = o.step_two(M_G, w, g_P[0], n_u, N)
    z_flat = o.step_two(Mf, w, g_P[0], n_u, N, flat=True)
    assert np.array_equal(z_dense, z_flat)                       # same non-zero terms in the same order
    y_dense = o.step_four(G_L, w, p_D[0], z_dense, n_u, N)
    y_flat = o.step_four(Gf, w, p_D[0], z_dense, n_u, N, flat=True)
    # the flat variant adds (sum + w) + p_D, the dense one sum + (w + p_D) (seq_functions.cpp:37 vs :84)
    assert np.allclose(y_dense, y_flat, rtol=2e-6, atol=1e-7) and np.array_equal(y_dense > 0, y_flat > 0)
    # a dense operator without the battery structure does not flatten
    _, _, r2 = G.flatten_operators(n_u, N, pb.m, rng.standard_normal(M_G.shape).astype(np.float32), G_L)
    assert r2 > 0.1


@pytest.mark.parametrize("kernel,max_bn", [(0, 256), (1, 208)])
def test_batch_column_tiling_covers_every_width(kernel, max_bn):
    """Host-side tile planner of the batch kernels (batch_tc.cu plan_tiles, batch_tc_p1.cu plan_tiles_p1):
    every operator width from 1 to 4800 columns (2x the quadrotor's m = 2400) is covered by the planned tiles,
    tiles are UMMA-legal (N multiple of 16, <= 256), start on 128-byte lines whenever the stride differs from
    the width, and the accumulators plus the state ring fit the 512 TMEM columns of one SM."""
    for ncols in list(range(1, 1200)) + [2399, 2400, 2401, 4799, 4800]:
        bn, nt, step, tmem = G.debug_plan_tiles(kernel, ncols)
        assert bn % 16 == 0 and 16 <= bn <= max_bn, (ncols, bn)
        assert 1 <= step <= bn and (step == bn or step % 32 == 0), (ncols, bn, step)
        assert (nt - 1) * step + bn >= ncols, (ncols, bn, nt, step)          # last tile reaches the last column
        assert nt == 1 or (nt - 1) * step < ncols, (ncols, bn, nt, step)     # no tile entirely past the end
        assert tmem <= 512, (ncols, bn, tmem)
    assert G.debug_plan_tiles(1, 400)[:3] == (208, 2, 192)                  # quadrotor product 1: aligned second tile
    with pytest.raises(G.GpadError):
        G.debug_plan_tiles(3, 400)


def test_fp16_product2_tiling_is_whole_epilogue_blocks():
    """batch_tc_p2.cu plan_tiles_p2: the TMA-fed epilogue works in 32-column blocks, so tiles are multiples of 32
    columns (<= 256, a legal UMMA N), cover every width without a tile entirely past the end, and two 256-column
    accumulator stages fit the 512 TMEM columns"""
    for ncols in list(range(1, 1200)) + [2399, 2400, 2401, 4200, 4799, 4800]:
        bn, nt, step, tmem = G.debug_plan_tiles(2, ncols)
        assert bn % 32 == 0 and 32 <= bn <= 256 and step == bn, (ncols, bn, step)
        assert nt * bn >= ncols and (nt - 1) * bn < ncols, (ncols, bn, nt)
        assert tmem <= 512
    assert G.debug_plan_tiles(2, 2400)[:2] == (256, 10)


# ------------------------------------------------------------------------------------ round 2: formats, plants, groups
def test_flat_data_file_round_trip(tmp_path):
    """the ENABLE_FLATTEN_MATRICES variant of the data file (main.cu:39-41,50-52): operators of N*m floats"""
    n_u, N = 3, 4
    pb = P.battery(n_u, N)
    Mf, Gf, resid = G.flatten_operators(n_u, N, pb.m, pb.M_G, pb.G_L)
    assert resid == 0.0 and Mf.shape == (N, pb.m) and Gf.shape == (pb.m, N)
    g_P, p_D, _ = pb.instance(np.array([0.1, -0.2, 0.3]))
    th, be = G.schedule(7)
    path = str(tmp_path / "flat.txt")
    G.file_write(path, n_u, N, pb.m, pb.L, Mf, g_P, Gf, p_D, th, be, flat=True)
    d = G.file_read(path, flat=True)
    assert (d["n_u"], d["N"], d["m"], d["num_iterations"]) == (n_u, N, pb.m, 7)
    assert np.array_equal(d["M_G"], Mf.ravel()) and np.array_equal(d["G_L"], Gf.ravel())
    assert np.array_equal(d["g_P"], g_P) and np.array_equal(d["p_D"], p_D) and np.array_equal(d["theta"], th)
    # the dense reader refuses the flat file (it promises n*m floats the file does not hold) instead of reading garbage
    with pytest.raises(G.GpadError):
        G.file_read(path, flat=False)


def test_file_writer_rejects_bad_arguments(tmp_path):
    fd = G.FileData(3, 4, 56, 5, 1.0)          # every pointer NULL
    assert G.lib().gpad_file_write(str(tmp_path / "x.txt").encode(), C.byref(fd)) == 1
    a = np.zeros(3 * 4 * 56, np.float32)
    fd = G.FileData(3, 4, 56, 5, 1.0, G._f32p(a), G._f32p(a), G._f32p(a), G._f32p(a), None, None)     # iterations without theta / beta
    assert G.lib().gpad_file_write(str(tmp_path / "x.txt").encode(), C.byref(fd)) == 1
    fd = G.FileData(3, 4, 56, -1, 1.0, G._f32p(a), G._f32p(a), G._f32p(a), G._f32p(a), G._f32p(a), G._f32p(a))
    assert G.lib().gpad_file_write(str(tmp_path / "x.txt").encode(), C.byref(fd)) == 1
    bad = tmp_path / "huge.txt"
    bad.write_text("1000 1000 100000 10 1.0\n0.5 0.5\n")                                             # header promises 2e11 floats
    out = G.FileData()
    assert G.lib().gpad_file_read(str(bad).encode(), C.byref(out)) == 6


@pytest.mark.parametrize("case", [1, 2, 3, 4, 5])
def test_step3_fixture_reader_on_reference_files(golden_dir, case, tmp_path):
    """gpad_fixture_read (step3.cu:59-81) on the reference's own files when the reference tree is present, and on the
    same vectors rewritten by gpad_fixture_write from the committed golden copy otherwise"""
    g = np.load(os.path.join(golden_dir, "step3_fixtures.npz"))
    ref_dir = f"/root/reference/Code/CUDA/FinalProject/build/step3/{case}"
    if os.path.isdir(ref_dir):
        fx = G.fixture_read(ref_dir, 3)
    else:
        n_u, N, m = (int(v) for v in g[f"dims{case}"])
        G.fixture_write(str(tmp_path), 3, n_u, N, m, theta=float(g[f"theta{case}"]), z_prev=g[f"z_prev{case}"],
                        zhat_in=g[f"zhat{case}"], z_out=g[f"z{case}"])
        fx = G.fixture_read(str(tmp_path), 3)
    assert (fx["n_u"], fx["N"], fx["m"]) == tuple(int(v) for v in g[f"dims{case}"])
    assert abs(fx["theta"] - float(g[f"theta{case}"])) < 1e-8
    for a, b in (("z_prev", "z_prev"), ("zhat_in", "zhat"), ("z_out", "z")):
        assert np.max(np.abs(fx[a] - g[f"{b}{case}"])) <= 6e-8, a          # golden holds the text values in float64


@pytest.mark.parametrize("flat", [False, True])
def test_step2_step4_fixture_round_trip(flat, tmp_path):
    n_u, N = 3, 2
    m = 4 * n_u * N + 2 * N
    n = n_u * N
    rng = np.random.default_rng(4)
    r = lambda k: rng.standard_normal(k).astype(np.float32)
    op = r((N if flat else n) * m)
    d2, d4 = tmp_path / "2", tmp_path / "4"
    d2.mkdir(); d4.mkdir()
    v2 = dict(op=op, w=r(m), g_P=r(n), prod=r(n), zhat_out=r(n))
    v4 = dict(op=op, w=r(m), zhat_in=r(n), p_D=r(m), prod=r(m), sum=r(m), y_next=r(m))
    G.fixture_write(str(d2), 2, n_u, N, m, flat=flat, **v2)
    G.fixture_write(str(d4), 4, n_u, N, m, flat=flat, **v4)
    f2, f4 = G.fixture_read(str(d2), 2, flat=flat), G.fixture_read(str(d4), 4, flat=flat)
    for k, v in v2.items():
        assert np.array_equal(f2[k], v), k
    for k, v in v4.items():
        assert np.array_equal(f4[k], v), k
    with pytest.raises(G.GpadError):
        G.fixture_read(str(tmp_path / "missing"), 4, flat=flat)
    with pytest.raises(G.GpadError):
        G.fixture_read(str(d2), 4, flat=flat)              # a step-2 fixture is too short for the step-4 layout


def test_plants_condensing_matches_numpy_restatement():
    """per-instance plants (BASELINE config 5): every pack condensed from its own capacities equals the numpy
    restatement of gpad.m with gpad.m:18 scaled; scale 1 reproduces gpad_problem_battery bit for bit"""
    n_u, N, B = 3, 4, 9
    rng = np.random.default_rng(8)
    scale = 1.0 + 0.1 * (2 * rng.random((B, n_u)) - 1)
    scale[0] = 1.0
    pl = G.Plants(n_u, N, scale, threads=3)
    assert (pl.n_u, pl.N, pl.m, pl.n_par, pl.B) == (n_u, N, 56, n_u, B)
    M, Gl, L = pl.operators()
    x0 = rng.random((B, n_u)) - 0.5
    g_P, p_D, f = pl.instances(x0, want_f=True)
    for b in range(B):
        ref = P.battery(n_u, N, cap_scale=scale[b])
        assert abs(L[b] - ref.L) <= 1e-6 * ref.L
        assert np.allclose(M[b].reshape(ref.n, ref.m), ref.M_G, rtol=2e-6, atol=1e-9)
        assert np.allclose(Gl[b].reshape(ref.m, ref.n), ref.G_L, rtol=2e-6, atol=0)
        rg, rp, rf = ref.instance(x0[b])
        assert np.allclose(g_P[b], rg, rtol=2e-6, atol=1e-9) and np.allclose(p_D[b], rp, rtol=2e-6, atol=0)
        assert np.allclose(f[b], rf, rtol=2e-6, atol=1e-9)
    base = G.Problem("battery", n_u=n_u, N=N)
    M0, G0 = base.operators()
    assert np.array_equal(M[0].reshape(M0.shape), M0) and np.array_equal(Gl[0].reshape(G0.shape), G0)
    Mf, Gf, _ = pl.operators(G.LAYOUT_FLIPPED)
    assert np.array_equal(Mf[3].reshape(pl.m, pl.n), M[3].reshape(pl.n, pl.m).T)
    with pytest.raises(G.GpadError):
        G.Plants(n_u, N, np.zeros((2, n_u)))            # non-positive capacity


def test_group_and_async_entry_points_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pb = P.battery(3, 4)
    with pytest.raises(G.GpadError, match="no CUDA device|no usable"):
        G.Group([0, 1], 3, 4, pb.m, pb.L, pb.M_G, pb.G_L, max_batch=256)


def test_lipschitz_constant_is_selectable():
    """acceldualgrad.m:11 uses L = ||H||_F^2; the paper (section 4) lambda_max(G H^-1 G'): both on request"""
    ref = P.battery(3, 4)
    pb = G.Problem("battery", n_u=3, N=4)
    assert abs(pb.L - ref.L) <= 1e-6 * ref.L                                # the reference's choice is the battery default
    M0, G0 = pb.operators()
    lam = np.linalg.eigvalsh(ref.G @ np.linalg.solve(ref.H, ref.G.T)).max()
    L1 = pb.set_lipschitz(G.L_LAMBDA_MAX)
    # power iteration approaches lambda_max from below (clustered eigenvalues: 0.1 % short after 400 steps), hence the 2 % margin
    assert lam <= L1 <= 1.021 * lam and L1 < 0.8 * ref.L                    # SURVEY: 7.9986 against 12.0379
    M1, G1 = pb.operators()
    assert np.array_equal(M0, M1)                                           # M_G does not depend on L
    assert np.allclose(G1, G0 * (ref.L / L1), rtol=2e-6)
    g0, p0, _ = G.Problem("battery", n_u=3, N=4).instances(np.array([[0.1, -0.2, 0.3]]))
    g1, p1, _ = pb.instances(np.array([[0.1, -0.2, 0.3]]))
    assert np.array_equal(g0, g1) and np.allclose(p1, p0 * (ref.L / L1), rtol=2e-6)
    assert abs(pb.set_lipschitz(G.L_REFERENCE) - ref.L) <= 1e-6 * ref.L
    quad = G.Problem("quadrotor", N=10)
    Lq = quad.L
    assert abs(quad.set_lipschitz(G.L_LAMBDA_MAX) - Lq) <= 1e-6 * Lq        # the quadrotor default already is the paper's


def test_fp16x3_split_numerics_emulated():
    """GPAD_PREC_FP16X3 in numpy (csrc/tc_ptx.cuh:f16_scale_exp / split_f16x2, csrc/batch_f16.cu): every row is scaled by
    the power of two that brings its largest magnitude into [2^14, 2^15), split into fp16 hi + lo, and the product of two
    such operands is hi*lo + lo*hi + hi*hi accumulated in fp32.  Checked here without a GPU: (a) the scale is exact and
    puts the row maximum where claimed; (b) hi + lo reproduces entries within 2^-18 of the row maximum to 22 bits and
    smaller ones to 2^-40 of the row maximum; (c) the three-term product of wide-range rows is as close to the exact
    product as the same construction with tf32 (10 explicit mantissa bits, fp32 exponent range) operands"""
    rng = np.random.default_rng(5)

    def scale_exp(row_max):
        e = np.zeros(row_max.shape, np.int64)
        ok = (row_max > 0) & np.isfinite(row_max)
        ex = np.floor(np.log2(row_max[ok])).astype(np.int64)
        e[ok] = np.clip(14 - ex, -100, 100)
        return e

    def split_f16(x):
        e = scale_exp(np.abs(x).max(axis=1))
        xs = (x.astype(np.float64) * 2.0 ** e[:, None]).astype(np.float32)      # power of two: exact in fp32
        hi = xs.astype(np.float16)
        lo = (xs - hi.astype(np.float32)).astype(np.float16)                  # the remainder is exact in fp32
        return hi, lo, e, xs

    def rn_tf32(x):
        b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
        b = (b + 0x1000) & 0xFFFFE000                                          # round to nearest (ties away), 10 mantissa bits
        return b.astype(np.uint32).view(np.float32)

    M, N, K = 40, 24, 512
    A = (rng.standard_normal((M, K)) * 10.0 ** rng.uniform(-6, 6, (M, 1)) * 10.0 ** rng.uniform(-3, 0, (M, K))).astype(np.float32)
    A[rng.random((M, K)) < 0.3] = 0.0
    A[7] = 0.0
    B = (rng.standard_normal((N, K)) * 10.0 ** rng.uniform(-5, 3, (N, 1))).astype(np.float32)
    ah, al, ea, axs = split_f16(A)
    bh, bl, eb, _ = split_f16(B)
    # (a)
    amax = np.abs(axs).max(axis=1)
    nz = amax > 0
    assert (amax[nz] >= 2.0 ** 14).all() and (amax[nz] < 2.0 ** 15).all() and ea[7] == 0
    assert np.array_equal((axs.astype(np.float64) * 2.0 ** (-ea[:, None])).astype(np.float32), A)      # undoing the scale is exact
    assert np.isfinite(ah.astype(np.float32)).all() and np.isfinite(al.astype(np.float32)).all()
    # (b)
    rec = ah.astype(np.float64) + al.astype(np.float64)
    err = np.abs(rec - axs.astype(np.float64))
    big = np.abs(axs) >= 2.0 ** -3
    assert (err[big] <= 2.0 ** -22 * np.abs(axs[big])).all()
    assert (err[~big] <= 2.0 ** -25).all()                                     # half an fp16 subnormal step = 2^-40 of 2^15
    # (c)
    def three_products(xh, xl, yh, yl):
        f = lambda v: v.astype(np.float32).astype(np.float64)
        return (f(xh) @ f(yl).T + f(xl) @ f(yh).T + f(xh) @ f(yh).T)
    c16 = three_products(ah, al, bh, bl) * 2.0 ** (-ea[:, None]) * 2.0 ** (-eb[None, :])
    th = rn_tf32(A); tl = rn_tf32(A - th); uh = rn_tf32(B); ul = rn_tf32(B - uh)
    c32 = three_products(th, tl, uh, ul)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    scale[scale == 0] = 1.0
    e16, e32 = np.max(np.abs(c16 - ref) / scale), np.max(np.abs(c32 - ref) / scale)
    assert e16 <= 3 * 2.0 ** -22 and e16 <= 1.5 * e32 + 1e-8, (e16, e32)
    assert (c16[7] == 0).all()
