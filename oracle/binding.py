"""ctypes bindings for oracle/liboracle.so (our C restatement) and oracle/_ref/libgpad_ref.so
(the reference's own seq_functions.cpp compiled where it lies).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

STATUS_NAMES = {0: "max_iter", 1: "converged_z", 2: "converged_zhat", 3: "converged_dual", 4: "nonfinite"}


def build(force=False):
    """Compile liboracle.so and, when /root/reference is present, _ref/libgpad_ref.so."""
    lib = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "gpad_oracle.c")
    ref = os.path.join(_HERE, "_ref", "libgpad_ref.so")
    refcuda = os.path.join(_HERE, "_ref", "libgpad_refcuda.so")
    stale = (not os.path.exists(lib)) or os.path.getmtime(lib) < os.path.getmtime(src)
    want_ref = os.path.isdir("/root/reference/Code/CUDA/FinalProject/src") and not (os.path.exists(ref) and os.path.exists(refcuda))
    if force or stale or want_ref:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return lib


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libgpad_ref.so"))


def have_refcuda():
    return os.path.exists(os.path.join(_HERE, "_ref", "libgpad_refcuda.so"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


class _Problem(C.Structure):
    _fields_ = [("n_u", C.c_int), ("N", C.c_int), ("m", C.c_int),
                ("M_G", _fp), ("G_L", _fp), ("g_P", _fp), ("p_D", _fp),
                ("theta", _fp), ("beta", _fp), ("max_iter", C.c_int),
                ("y0", _fp), ("y_prev0", _fp), ("z0", _fp),
                ("check_every", C.c_int), ("eps_g", C.c_float), ("eps_V", C.c_float), ("L", C.c_float),
                ("f", _fp)]


class _Result(C.Structure):
    _fields_ = [("y_next", _fp), ("y", _fp), ("z", _fp), ("zhat", _fp), ("w", _fp),
                ("iters", C.c_int), ("status", C.c_int), ("max_viol", C.c_float), ("gap", C.c_float)]


def schedule(count, variant="paper"):
    th = np.zeros(count, np.float32)
    be = np.zeros(count, np.float32)
    Oracle().lib.oracle_schedule(_p(th), _p(be), count, 0 if variant == "paper" else 1)
    return th, be


class Oracle:
    """The plain-C restatement (oracle/gpad_oracle.c)."""
    _lib = None

    def __init__(self):
        if Oracle._lib is None:
            Oracle._lib = C.CDLL(build())
            L = Oracle._lib
            L.oracle_solve.argtypes = [C.POINTER(_Problem), C.POINTER(_Result)]
            L.oracle_solve_f64.argtypes = [C.POINTER(_Problem), _dp, _dp, _dp, _dp, _dp, _ip, _ip]
            L.oracle_solve_batch.argtypes = [C.POINTER(_Problem), C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _ip, _ip, C.c_int]
            L.oracle_schedule.argtypes = [_fp, _fp, C.c_int, C.c_int]
            for name in ("oracle_step_two", "oracle_step_two_flat"):
                getattr(L, name).argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
            for name in ("oracle_step_four", "oracle_step_four_flat"):
                getattr(L, name).argtypes = [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
            L.oracle_step_one.argtypes = [_fp, _fp, _fp, C.c_float, C.c_int]
            L.oracle_step_three.argtypes = [C.c_float, C.c_int, _fp, _fp, _fp]
        self.lib = Oracle._lib

    # ---- steps ----
    def step_one(self, y, y_prev, beta):
        y, y_prev = _f32(y), _f32(y_prev)
        w = np.empty_like(y)
        self.lib.oracle_step_one(_p(y), _p(y_prev), _p(w), float(beta), y.size)
        return w

    def step_two(self, M_G, w, g_P, n_u, N, flat=False):
        M_G, w, g_P = _f32(M_G), _f32(w), _f32(g_P)
        zhat = np.empty(n_u * N, np.float32)
        fn = self.lib.oracle_step_two_flat if flat else self.lib.oracle_step_two
        fn(_p(M_G), _p(w), _p(g_P), _p(zhat), N, n_u, w.size)
        return zhat

    def step_three(self, theta, z_prev, zhat):
        z_prev, zhat = _f32(z_prev), _f32(zhat)
        z = np.empty_like(z_prev)
        self.lib.oracle_step_three(float(theta), z.size, _p(z_prev), _p(zhat), _p(z))
        return z

    def step_four(self, G_L, w, p_D, zhat, n_u, N, flat=False):
        G_L, w, p_D, zhat = _f32(G_L), _f32(w), _f32(p_D), _f32(zhat)
        y = np.empty_like(w)
        fn = self.lib.oracle_step_four_flat if flat else self.lib.oracle_step_four
        fn(_p(G_L), _p(y), _p(w), _p(p_D), _p(zhat), N, n_u, w.size)
        return y

    # ---- whole solve ----
    def _problem(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter, y0, y_prev0, z0,
                 check_every, eps_g, eps_V, L, f, keep):
        arrs = [_f32(M_G), _f32(G_L), _f32(g_P), _f32(p_D), _f32(theta), _f32(beta)]
        opt = [None if a is None else _f32(a) for a in (y0, y_prev0, z0, f)]
        keep.extend(arrs + opt)
        assert arrs[0].size == n_u * N * m and arrs[1].size == n_u * N * m
        assert arrs[4].size >= max_iter and arrs[5].size >= max_iter
        return _Problem(n_u, N, m, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]),
                        _p(arrs[5]), max_iter, _p(opt[0]), _p(opt[1]), _p(opt[2]), check_every,
                        eps_g, eps_V, L, _p(opt[3]))

    def solve(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter=None, y0=None,
              y_prev0=None, z0=None, check_every=0, eps_g=0.0, eps_V=0.0, L=1.0, f=None):
        """fp32 solve, M_G [n][m], G_L [m][n].  Returns dict of the five vectors + iters/status."""
        n = n_u * N
        max_iter = len(theta) if max_iter is None else max_iter
        keep = []
        prob = self._problem(n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter, y0, y_prev0, z0,
                             check_every, eps_g, eps_V, L, f, keep)
        out = {k: np.zeros(sz, np.float32) for k, sz in
               (("y_next", m), ("y", m), ("z", n), ("zhat", n), ("w", m))}
        res = _Result(_p(out["y_next"]), _p(out["y"]), _p(out["z"]), _p(out["zhat"]), _p(out["w"]),
                      0, 0, 0.0, 0.0)
        rc = self.lib.oracle_solve(C.byref(prob), C.byref(res))
        assert rc == 0
        out.update(iters=res.iters, status=res.status, max_viol=res.max_viol, gap=res.gap)
        return out

    def solve_f64(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter=None, y0=None,
                  y_prev0=None, z0=None, check_every=0, eps_g=0.0, eps_V=0.0, L=1.0, f=None):
        n = n_u * N
        max_iter = len(theta) if max_iter is None else max_iter
        keep = []
        prob = self._problem(n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter, y0, y_prev0, z0,
                             check_every, eps_g, eps_V, L, f, keep)
        out = {k: np.zeros(sz, np.float64) for k, sz in
               (("y_next", m), ("y", m), ("z", n), ("zhat", n), ("w", m))}
        it, st = C.c_int(0), C.c_int(0)
        d = lambda a: a.ctypes.data_as(_dp)
        rc = self.lib.oracle_solve_f64(C.byref(prob), d(out["y_next"]), d(out["y"]), d(out["z"]),
                                       d(out["zhat"]), d(out["w"]), C.byref(it), C.byref(st))
        assert rc == 0
        out.update(iters=it.value, status=st.value)
        return out

    def solve_batch(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter=None, y0=None,
                    y_prev0=None, check_every=0, eps_g=0.0, eps_V=0.0, L=1.0, f=None, nthreads=0):
        """g_P [B][n], p_D [B][m] instance-major; outputs instance-major."""
        n = n_u * N
        g_P, p_D = _f32(g_P).reshape(-1, n), _f32(p_D).reshape(-1, m)
        B = g_P.shape[0]
        max_iter = len(theta) if max_iter is None else max_iter
        keep = []
        prob = self._problem(n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter, y0, y_prev0, None,
                             check_every, eps_g, eps_V, L, f, keep)
        out = {k: np.zeros((B, sz), np.float32) for k, sz in
               (("y_next", m), ("y", m), ("z", n), ("zhat", n), ("w", m))}
        iters = np.zeros(B, np.int32)
        status = np.zeros(B, np.int32)
        used = self.lib.oracle_solve_batch(C.byref(prob), B, _p(g_P), _p(p_D), _p(out["y_next"]),
                                           _p(out["y"]), _p(out["z"]), _p(out["zhat"]), _p(out["w"]),
                                           iters.ctypes.data_as(_ip), status.ctypes.data_as(_ip), nthreads)
        out.update(iters=iters, status=status, threads=used)
        return out


class RefLib:
    """The reference's own compiled step functions + our composition shim (oracle/_ref)."""

    def __init__(self):
        build()
        path = os.path.join(_HERE, "_ref", "libgpad_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.StepOneGPADSequential.argtypes = [_fp, _fp, _fp, C.c_float, C.c_int]
        for name in ("StepTwoGPADSequential", "StepTwoGPADFlatSequential"):
            getattr(L, name).argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.StepThreeGPADSequential.argtypes = [C.c_float, C.c_int, _fp, _fp, _fp]
        for name in ("StepFourGPADSequential", "StepFourGPADFlatSequential"):
            getattr(L, name).argtypes = [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.ref_solve_fixed.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int,
                                      _fp, _fp, _fp, _fp, _fp, _fp, _fp]
        L.ref_solve_fixed_batch.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int,
                                            _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int]
        L.ref_solve_fixed_batch.restype = C.c_int

    def step_one(self, y, y_prev, beta):
        y, y_prev = _f32(y), _f32(y_prev)
        w = np.empty_like(y)
        self.lib.StepOneGPADSequential(_p(y), _p(y_prev), _p(w), float(beta), y.size)
        return w

    def step_two(self, M_G, w, g_P, n_u, N, flat=False):
        M_G, w, g_P = _f32(M_G), _f32(w), _f32(g_P)
        zhat = np.empty(n_u * N, np.float32)
        fn = self.lib.StepTwoGPADFlatSequential if flat else self.lib.StepTwoGPADSequential
        fn(_p(M_G), _p(w), _p(g_P), _p(zhat), N, n_u, w.size)
        return zhat

    def step_three(self, theta, z_prev, zhat):
        z_prev, zhat = _f32(z_prev), _f32(zhat)
        z = np.empty_like(z_prev)
        self.lib.StepThreeGPADSequential(float(theta), z.size, _p(z_prev), _p(zhat), _p(z))
        return z

    def step_four(self, G_L, w, p_D, zhat, n_u, N, flat=False):
        G_L, w, p_D, zhat = _f32(G_L), _f32(w), _f32(p_D), _f32(zhat)
        y = np.empty_like(w)
        fn = self.lib.StepFourGPADFlatSequential if flat else self.lib.StepFourGPADSequential
        fn(_p(G_L), _p(y), _p(w), _p(p_D), _p(zhat), N, n_u, w.size)
        return y

    def solve(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter=None, y0=None, y_prev0=None):
        n = n_u * N
        max_iter = len(theta) if max_iter is None else max_iter
        a = [_f32(x) for x in (M_G, g_P, G_L, p_D, theta, beta)]
        o = [None if x is None else _f32(x) for x in (y0, y_prev0)]
        out = {k: np.zeros(sz, np.float32) for k, sz in
               (("y_next", m), ("y", m), ("z", n), ("zhat", n), ("w", m))}
        self.lib.ref_solve_fixed(n_u, N, m, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]),
                                 max_iter, _p(o[0]), _p(o[1]), _p(out["y_next"]), _p(out["y"]),
                                 _p(out["z"]), _p(out["zhat"]), _p(out["w"]))
        out.update(iters=max_iter, status=0)
        return out

    def solve_batch(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter=None, nthreads=0):
        n = n_u * N
        g_P, p_D = _f32(g_P).reshape(-1, n), _f32(p_D).reshape(-1, m)
        B = g_P.shape[0]
        max_iter = len(theta) if max_iter is None else max_iter
        a = [_f32(x) for x in (M_G, G_L, theta, beta)]
        out = {k: np.zeros((B, sz), np.float32) for k, sz in
               (("y_next", m), ("y", m), ("z", n), ("zhat", n), ("w", m))}
        used = self.lib.ref_solve_fixed_batch(n_u, N, m, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), max_iter, B,
                                              _p(g_P), _p(p_D), _p(out["y_next"]), _p(out["y"]), _p(out["z"]),
                                              _p(out["zhat"]), _p(out["w"]), nthreads)
        out.update(threads=used)
        return out


class RefCuda:
    """The reference's own GPU kernels (kernel_functions.cu, unmodified, compiled for sm_100a) driven by the loop of
    main.cu:117-180 (oracle/ref_cuda_compose.cu).  Needs a GPU; operators in the sequential layout are flipped here
    into what the kernels read."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libgpad_refcuda.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.ref_cuda_solve.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int,
                                     _fp, _fp, _fp, _fp, _fp, _dp, _dp]
        L.ref_cuda_loop_resident.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _dp, _fp]

    def solve(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter=None):
        n = n_u * N
        max_iter = len(theta) if max_iter is None else max_iter
        MGf = _f32(np.asarray(M_G, np.float32).reshape(n, m).T); GLf = _f32(np.asarray(G_L, np.float32).reshape(m, n).T)
        a = [_f32(x) for x in (g_P, p_D, theta, beta)]
        out = {k: np.zeros(sz, np.float32) for k, sz in (("y_next", m), ("y", m), ("z", n), ("zhat", n), ("w", m))}
        loop, total = C.c_double(), C.c_double()
        rc = self.lib.ref_cuda_solve(n_u, N, m, _p(MGf), _p(a[0]), _p(GLf), _p(a[1]), _p(a[2]), _p(a[3]), max_iter,
                                     _p(out["y_next"]), _p(out["y"]), _p(out["z"]), _p(out["zhat"]), _p(out["w"]),
                                     C.byref(loop), C.byref(total))
        assert rc == 0, f"CUDA error {rc} in the reference kernels"
        out.update(iters=max_iter, status=0, loop_us=loop.value, total_us=total.value)
        return out

    def loop_times(self, n_u, N, m, M_G, G_L, g_P, p_D, theta, beta, max_iter, reps):
        """per-solve wall time (us) of the reference loop on resident operators, `reps` solves back to back"""
        n = n_u * N
        MGf = _f32(np.asarray(M_G, np.float32).reshape(n, m).T); GLf = _f32(np.asarray(G_L, np.float32).reshape(m, n).T)
        a = [_f32(x) for x in (g_P, p_D, theta, beta)]
        t = np.zeros(reps, np.float64); z = np.zeros(n, np.float32)
        rc = self.lib.ref_cuda_loop_resident(n_u, N, m, _p(MGf), _p(a[0]), _p(GLf), _p(a[1]), _p(a[2]), _p(a[3]), max_iter,
                                             reps, t.ctypes.data_as(_dp), _p(z))
        assert rc == 0, f"CUDA error {rc} in the reference kernels"
        return t, z
