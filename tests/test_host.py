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
    assert G.lib().gpad_api_version() == 1


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
        G.debug_plan_tiles(2, 400)
