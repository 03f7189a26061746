"""gpad_b200 -- thin ctypes binding of libgpad_b200.so (include/gpad.h).

Python is plumbing only: every solve goes through the C ABI into the hand-written sm_100a
kernels.  There is no CPU or PyTorch fallback: if the shared library is missing, or no B200 is
present, the calls raise.  Device buffers are passed as raw pointers (torch tensors'
``data_ptr()``), host buffers as numpy arrays.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libgpad_b200.so")

# enums of include/gpad.h
LAYOUT_FLIPPED, LAYOUT_SEQUENTIAL, LAYOUT_FLAT = 0, 1, 2
MODE_LATENCY, MODE_BATCH_SHARED, MODE_BATCH_PER_INSTANCE = 1, 2, 3
PREC_FP32, PREC_TF32X3, PREC_FP16X3 = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
SCHEDULE_PAPER, SCHEDULE_MATLAB_LAG = 0, 1
WARM_COLD, WARM_PREVIOUS, WARM_SHIFTED = 0, 1, 2
L_REFERENCE, L_LAMBDA_MAX = 0, 1
STATUS_NAMES = {0: "max_iter", 1: "converged_z", 2: "converged_zhat", 3: "converged_dual", 4: "nonfinite"}

EXPORTS = [
    "gpad_status_string", "gpad_last_error", "gpad_api_version", "gpad_device_count",
    "gpad_step_one", "gpad_step_two", "gpad_array_copy", "gpad_step_three", "gpad_step_four",
    "gpad_setup", "gpad_destroy", "gpad_solve", "gpad_launch_count", "gpad_describe",
    "gpad_profile_enable", "gpad_profile_read",
    "gpad_problem_battery", "gpad_problem_quadrotor", "gpad_problem_destroy", "gpad_problem_dims",
    "gpad_problem_operators", "gpad_problem_instances", "gpad_problem_plant", "gpad_schedule",
    "gpad_file_read", "gpad_file_write", "gpad_file_free", "gpad_debug_gemm_tf32x3", "gpad_debug_gemm_f16x3", "gpad_debug_plan_tiles",
    "gpad_flatten_operators", "gpad_expand_operators", "gpad_closed_loop",
    "gpad_solve_async", "gpad_wait", "gpad_handle_dims", "gpad_solve_stats", "gpad_instances_device",
    "gpad_plants_battery", "gpad_plants_destroy", "gpad_plants_dims", "gpad_plants_operators", "gpad_plants_instances",
    "gpad_closed_loop_plants",
    "gpad_group_setup", "gpad_group_destroy", "gpad_group_solve", "gpad_group_size", "gpad_group_shard",
    "gpad_file_read_flat", "gpad_file_write_flat", "gpad_fixture_read", "gpad_fixture_write", "gpad_fixture_free",
    "gpad_problem_set_lipschitz",
]

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


class GpadError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("n_u", C.c_int), ("N", C.c_int), ("m", C.c_int), ("L", C.c_float), ("layout", C.c_int),
                ("mode", C.c_int), ("precision", C.c_int), ("max_batch", C.c_int), ("device", C.c_int),
                ("operators_mem", C.c_int), ("reserved", C.c_int * 6)]


class SolveArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("mem", C.c_int),
                ("g_P", C.c_void_p), ("p_D", C.c_void_p), ("f", C.c_void_p), ("y0", C.c_void_p), ("y_prev0", C.c_void_p),
                ("theta", _fp), ("beta", _fp), ("max_iter", C.c_int), ("check_every", C.c_int),
                ("eps_g", C.c_float), ("eps_V", C.c_float),
                ("y_next", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p), ("zhat", C.c_void_p), ("w", C.c_void_p),
                ("iters", C.c_void_p), ("status", C.c_void_p), ("max_viol", C.c_void_p), ("gap", C.c_void_p),
                ("stream", C.c_void_p), ("params", C.c_void_p), ("problem", C.c_void_p), ("build_f", C.c_int),
                ("reserved", C.c_int * 3)]


class SolveStats(C.Structure):
    _fields_ = [("instance_iterations_scheduled", C.c_double), ("instance_iterations_needed", C.c_double),
                ("compactions", C.c_int), ("reserved", C.c_int)]


class Fixture(C.Structure):
    _fields_ = [("step", C.c_int), ("n_u", C.c_int), ("N", C.c_int), ("m", C.c_int), ("flat", C.c_int), ("theta", C.c_float),
                ("op", _fp), ("w", _fp), ("g_P", _fp), ("p_D", _fp), ("zhat_in", _fp), ("z_prev", _fp),
                ("prod", _fp), ("sum", _fp), ("zhat_out", _fp), ("z_out", _fp), ("y_next", _fp)]


class FileData(C.Structure):
    _fields_ = [("n_u", C.c_int), ("N", C.c_int), ("m", C.c_int), ("num_iterations", C.c_int), ("L", C.c_float),
                ("M_G", _fp), ("g_P", _fp), ("G_L", _fp), ("p_D", _fp), ("theta", _fp), ("beta", _fp)]


_lib = None


def lib():
    """Loads libgpad_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpadError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.gpad_status_string.restype = C.c_char_p
        L.gpad_last_error.restype = C.c_char_p
        L.gpad_describe.restype = C.c_char_p
        L.gpad_describe.argtypes = [C.c_void_p]
        L.gpad_launch_count.restype = C.c_longlong
        L.gpad_launch_count.argtypes = [C.c_void_p]
        L.gpad_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.gpad_profile_read.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.gpad_setup.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.gpad_destroy.argtypes = [C.c_void_p]
        L.gpad_solve.argtypes = [C.c_void_p, C.POINTER(SolveArgs)]
        L.gpad_step_one.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.gpad_step_two.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gpad_array_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gpad_step_three.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gpad_step_four.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p]
        L.gpad_problem_battery.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.gpad_problem_quadrotor.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.gpad_problem_destroy.argtypes = [C.c_void_p]
        L.gpad_problem_dims.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _fp]
        L.gpad_problem_operators.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
        L.gpad_problem_instances.argtypes = [C.c_void_p, C.c_int, _dp, _fp, _fp, _fp]
        L.gpad_problem_plant.argtypes = [C.c_void_p, _ip, _dp, _dp]
        L.gpad_schedule.argtypes = [_fp, _fp, C.c_int, C.c_int]
        L.gpad_flatten_operators.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp]
        L.gpad_expand_operators.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp]
        L.gpad_closed_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, _dp, C.c_int, _fp, _fp, C.c_int, C.c_int, _dp, _dp]
        L.gpad_file_read.argtypes = [C.c_char_p, C.POINTER(FileData)]
        L.gpad_file_write.argtypes = [C.c_char_p, C.POINTER(FileData)]
        L.gpad_file_free.argtypes = [C.POINTER(FileData)]
        v2 = {
            "gpad_solve_async": [C.c_void_p, C.POINTER(SolveArgs), C.POINTER(C.c_longlong)],
            "gpad_wait": [C.c_void_p, C.c_longlong],
            "gpad_handle_dims": [C.c_void_p] + [_ip] * 6,
            "gpad_solve_stats": [C.c_void_p, C.POINTER(SolveStats)],
            "gpad_instances_device": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "gpad_plants_battery": [C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.POINTER(C.c_void_p)],
            "gpad_plants_destroy": [C.c_void_p],
            "gpad_plants_dims": [C.c_void_p] + [_ip] * 5,
            "gpad_plants_operators": [C.c_void_p, C.c_int, _fp, _fp, _fp],
            "gpad_plants_instances": [C.c_void_p, _dp, _fp, _fp, _fp],
            "gpad_closed_loop_plants": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _dp, C.c_int, _fp, _fp, C.c_int, C.c_int, _dp, _dp],
            "gpad_group_setup": [C.POINTER(Config), _ip, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)],
            "gpad_group_destroy": [C.c_void_p],
            "gpad_group_solve": [C.c_void_p, C.POINTER(SolveArgs)],
            "gpad_group_size": [C.c_void_p],
            "gpad_group_shard": [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), _ip, _ip],
            "gpad_file_read_flat": [C.c_char_p, C.POINTER(FileData)],
            "gpad_file_write_flat": [C.c_char_p, C.POINTER(FileData)],
            "gpad_fixture_read": [C.c_char_p, C.c_int, C.c_int, C.POINTER(Fixture)],
            "gpad_fixture_write": [C.c_char_p, C.POINTER(Fixture)],
            "gpad_fixture_free": [C.POINTER(Fixture)],
            "gpad_problem_set_lipschitz": [C.c_void_p, C.c_int, _fp],
        }
        for name, sig in v2.items():          # API version 2 entry points (an older library simply lacks them)
            if hasattr(L, name):
                getattr(L, name).argtypes = sig
        L.gpad_debug_gemm_tf32x3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gpad_debug_gemm_f16x3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gpad_debug_plan_tiles.argtypes = [C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4
        _lib = L
    return _lib


def check(rc, what="gpad call"):
    if rc != 0:
        L = lib()
        raise GpadError(f"{what}: {L.gpad_status_string(rc).decode()} -- {L.gpad_last_error().decode()}")


def device_count():
    return lib().gpad_device_count()


def _f32p(a):
    return a.ctypes.data_as(_fp)


def _ptr(x):
    """raw address of a numpy array / torch tensor / int / None"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if isinstance(x, int):
        return x
    return x.data_ptr()


def schedule(count, variant=SCHEDULE_PAPER):
    th = np.zeros(count, np.float32)
    be = np.zeros(count, np.float32)
    check(lib().gpad_schedule(_f32p(th), _f32p(be), count, variant), "gpad_schedule")
    return th, be


class Problem:
    """Host-side condensed MPC problem (C++ restatement of gpad.m / acceldualgrad.m precompute)."""

    def __init__(self, kind, **kw):
        self._h = C.c_void_p()
        L = lib()
        if kind == "battery":
            check(L.gpad_problem_battery(kw["n_u"], kw["N"], C.byref(self._h)), "gpad_problem_battery")
        elif kind == "quadrotor":
            check(L.gpad_problem_quadrotor(kw.get("N", 100), C.byref(self._h)), "gpad_problem_quadrotor")
        else:
            raise ValueError(kind)
        nu, N, m, npar, Lc = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_float()
        check(L.gpad_problem_dims(self._h, C.byref(nu), C.byref(N), C.byref(m), C.byref(npar), C.byref(Lc)))
        self.kind, self.n_u, self.N, self.m, self.n_par, self.L = kind, nu.value, N.value, m.value, npar.value, Lc.value
        self.n = self.n_u * self.N

    def set_lipschitz(self, which):
        """L_REFERENCE: ||H||_F^2 (acceldualgrad.m:11); L_LAMBDA_MAX: 1.02 lambda_max(G H^-1 G') (paper section 4)"""
        Lc = C.c_float()
        check(lib().gpad_problem_set_lipschitz(self._h, which, C.byref(Lc)), "gpad_problem_set_lipschitz")
        self.L = Lc.value
        return self.L

    def operators(self, layout=LAYOUT_SEQUENTIAL):
        M_G = np.empty(self.n * self.m, np.float32)
        G_L = np.empty(self.n * self.m, np.float32)
        check(lib().gpad_problem_operators(self._h, layout, _f32p(M_G), _f32p(G_L)), "gpad_problem_operators")
        if layout == LAYOUT_SEQUENTIAL:
            return M_G.reshape(self.n, self.m), G_L.reshape(self.m, self.n)
        return M_G.reshape(self.m, self.n), G_L.reshape(self.n, self.m)

    def instances(self, params, want_f=True):
        params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
        B = params.shape[0]
        assert params.shape[1] == self.n_par
        g_P = np.empty((B, self.n), np.float32)
        p_D = np.empty((B, self.m), np.float32)
        f = np.empty((B, self.n), np.float32) if want_f else None
        check(lib().gpad_problem_instances(self._h, B, params.ctypes.data_as(_dp), _f32p(g_P), _f32p(p_D),
                                           _f32p(f) if want_f else None), "gpad_problem_instances")
        return g_P, p_D, f

    def plant(self):
        nx = C.c_int()
        check(lib().gpad_problem_plant(self._h, C.byref(nx), None, None))
        A = np.empty((nx.value, nx.value)); Bm = np.empty((nx.value, self.n_u))
        check(lib().gpad_problem_plant(self._h, C.byref(nx), A.ctypes.data_as(_dp), Bm.ctypes.data_as(_dp)))
        return A, Bm

    def close(self):
        if self._h:
            lib().gpad_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solver:
    """gpad_setup / gpad_solve / gpad_destroy."""

    def __init__(self, n_u, N, m, L, M_G, G_L, layout=LAYOUT_SEQUENTIAL, mode=MODE_LATENCY, precision=PREC_FP32,
                 max_batch=1, device=-1, operators_mem=MEM_HOST):
        cfg = Config(n_u, N, m, float(L), layout, mode, precision, max_batch, device, operators_mem)
        if operators_mem == MEM_HOST:
            M_G = np.ascontiguousarray(M_G, np.float32)
            G_L = np.ascontiguousarray(G_L, np.float32)
            if n_u > 0 and N > 0 and m > 0:
                copies = max_batch if mode == MODE_BATCH_PER_INSTANCE else 1
                each = N * m if layout == LAYOUT_FLAT else n_u * N * m
                assert M_G.size == copies * each and G_L.size == copies * each
        self._keep = (M_G, G_L)
        self._h = C.c_void_p()
        check(lib().gpad_setup(C.byref(cfg), _ptr(M_G), _ptr(G_L), C.byref(self._h)), "gpad_setup")
        self.n_u, self.N, self.m, self.n, self.mode, self.max_batch = n_u, N, m, n_u * N, mode, max_batch

    @property
    def description(self):
        return lib().gpad_describe(self._h).decode()

    @property
    def launches(self):
        return lib().gpad_launch_count(self._h)

    def profile(self, enable=True):
        check(lib().gpad_profile_enable(self._h, 1 if enable else 0), "gpad_profile_enable")

    def profile_read(self, which):
        """(total_ms, launches) of kernel `which` (0 latency, 1 product 1, 2 product 2) since last read"""
        ms, cnt = C.c_double(), C.c_longlong()
        check(lib().gpad_profile_read(self._h, which, C.byref(ms), C.byref(cnt)), "gpad_profile_read")
        return ms.value, cnt.value

    def solve_host(self, g_P, p_D, theta, beta, max_iter=None, f=None, y0=None, y_prev0=None, check_every=0,
                   eps_g=0.0, eps_V=0.0, outputs=("y_next", "y", "z", "zhat", "w"), params=None, problem=None, build_f=False):
        """numpy in, numpy out (GPAD_MEM_HOST): H2D + solve + D2H + sync inside the call.
        params/problem: build g_P / p_D (and f) on the device from the parameter rows instead."""
        n, m = self.n, self.m
        if params is not None:
            params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
            B = params.shape[0]
            g_P = p_D = None
        else:
            g_P = np.ascontiguousarray(g_P, np.float32).reshape(-1, n)
            B = g_P.shape[0]
            p_D = np.ascontiguousarray(p_D, np.float32).reshape(B, m)
        opt = {k: (None if v is None else np.ascontiguousarray(v, np.float32).reshape(B, -1))
               for k, v in (("f", f), ("y0", y0), ("y_prev0", y_prev0))}
        theta = np.ascontiguousarray(theta, np.float32); beta = np.ascontiguousarray(beta, np.float32)
        max_iter = len(theta) if max_iter is None else max_iter
        out = {k: np.empty((B, m if k in ("y_next", "y", "w") else n), np.float32) for k in outputs}
        iters = np.zeros(B, np.int32); status = np.zeros(B, np.int32)
        viol = np.zeros(B, np.float32); gap = np.zeros(B, np.float32)
        a = SolveArgs(B, MEM_HOST, _ptr(g_P), _ptr(p_D), _ptr(opt["f"]), _ptr(opt["y0"]), _ptr(opt["y_prev0"]),
                      _f32p(theta), _f32p(beta), max_iter, check_every, eps_g, eps_V,
                      _ptr(out.get("y_next")), _ptr(out.get("y")), _ptr(out.get("z")), _ptr(out.get("zhat")),
                      _ptr(out.get("w")), _ptr(iters), _ptr(status), _ptr(viol), _ptr(gap), None,
                      _ptr(params), problem._h if problem is not None else None, 1 if build_f else 0)
        check(lib().gpad_solve(self._h, C.byref(a)), "gpad_solve")
        if B == 1 and self.mode == MODE_LATENCY:
            out = {k: v[0] for k, v in out.items()}
            out.update(iters=int(iters[0]), status=int(status[0]), max_viol=float(viol[0]), gap=float(gap[0]))
        else:
            out.update(iters=iters, status=status, max_viol=viol, gap=gap)
        return out

    def solve_device(self, batch, g_P, p_D, theta, beta, max_iter, stream=None, f=None, y0=None, y_prev0=None,
                     check_every=0, eps_g=0.0, eps_V=0.0, y_next=None, y=None, z=None, zhat=None, w=None,
                     iters=None, status=None, max_viol=None, gap=None, params=None, problem=None, build_f=False):
        """device pointers (torch tensors or ints) in and out (GPAD_MEM_DEVICE): enqueue only."""
        theta = np.ascontiguousarray(theta, np.float32); beta = np.ascontiguousarray(beta, np.float32)
        a = SolveArgs(batch, MEM_DEVICE, _ptr(g_P), _ptr(p_D), _ptr(f), _ptr(y0), _ptr(y_prev0),
                      _f32p(theta), _f32p(beta), max_iter, check_every, eps_g, eps_V,
                      _ptr(y_next), _ptr(y), _ptr(z), _ptr(zhat), _ptr(w), _ptr(iters), _ptr(status),
                      _ptr(max_viol), _ptr(gap), stream, _ptr(params), problem._h if problem is not None else None,
                      1 if build_f else 0)
        check(lib().gpad_solve(self._h, C.byref(a)), "gpad_solve")

    def solve_async(self, args):
        """gpad_solve_async on a prepared SolveArgs (host buffers, ideally pinned) -> ticket"""
        t = C.c_longlong()
        check(lib().gpad_solve_async(self._h, C.byref(args), C.byref(t)), "gpad_solve_async")
        return t.value

    def wait(self, ticket):
        check(lib().gpad_wait(self._h, ticket), "gpad_wait")

    def stats(self):
        """tolerance-mode bookkeeping of the last shared-operator batch solve"""
        st = SolveStats()
        check(lib().gpad_solve_stats(self._h, C.byref(st)), "gpad_solve_stats")
        return {"scheduled": st.instance_iterations_scheduled, "needed": st.instance_iterations_needed,
                "compactions": st.compactions}

    def close(self):
        if self._h:
            lib().gpad_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flatten_operators(n_u, N, m, M_G, G_L):
    """dense sequential -> flat ([N][m], [m][N]); returns (M_G_flat, G_L_flat, max_residual)"""
    M_G = np.ascontiguousarray(M_G, np.float32); G_L = np.ascontiguousarray(G_L, np.float32)
    Mf = np.empty((N, m), np.float32); Gf = np.empty((m, N), np.float32)
    res = C.c_float()
    check(lib().gpad_flatten_operators(n_u, N, m, _f32p(M_G), _f32p(G_L), _f32p(Mf), _f32p(Gf), C.byref(res)), "gpad_flatten_operators")
    return Mf, Gf, res.value


def expand_operators(n_u, N, m, M_G_flat, G_L_flat):
    Mf = np.ascontiguousarray(M_G_flat, np.float32); Gf = np.ascontiguousarray(G_L_flat, np.float32)
    M_G = np.empty((n_u * N, m), np.float32); G_L = np.empty((m, n_u * N), np.float32)
    check(lib().gpad_expand_operators(n_u, N, m, _f32p(Mf), _f32p(Gf), _f32p(M_G), _f32p(G_L)), "gpad_expand_operators")
    return M_G, G_L


def host_args(batch, theta, beta, max_iter, g_P=None, p_D=None, params=None, problem=None, outputs=None, iters=None, status=None,
              f=None, y0=None, y_prev0=None, check_every=0, eps_g=0.0, eps_V=0.0, build_f=False):
    """SolveArgs over caller-owned host arrays (keep them and theta/beta alive while the solve is in flight)"""
    o = outputs or {}
    return SolveArgs(batch, MEM_HOST, _ptr(g_P), _ptr(p_D), _ptr(f), _ptr(y0), _ptr(y_prev0), _f32p(theta), _f32p(beta), max_iter,
                     check_every, eps_g, eps_V, _ptr(o.get("y_next")), _ptr(o.get("y")), _ptr(o.get("z")), _ptr(o.get("zhat")),
                     _ptr(o.get("w")), _ptr(iters), _ptr(status), None, None, None, _ptr(params),
                     problem._h if problem is not None else None, 1 if build_f else 0)


class Plants:
    """B battery plants with per-instance cell capacities (BASELINE config 5), condensed on host threads."""

    def __init__(self, n_u, N, capacity_scale, threads=0):
        cs = np.ascontiguousarray(capacity_scale, np.float64).reshape(-1, n_u)
        self._h = C.c_void_p()
        check(lib().gpad_plants_battery(n_u, N, cs.shape[0], cs.ctypes.data_as(_dp), threads, C.byref(self._h)), "gpad_plants_battery")
        v = [C.c_int() for _ in range(5)]
        check(lib().gpad_plants_dims(self._h, *[C.byref(x) for x in v]))
        self.n_u, self.N, self.m, self.n_par, self.B = (x.value for x in v)
        self.n = self.n_u * self.N

    def operators(self, layout=LAYOUT_SEQUENTIAL):
        M_G = np.empty((self.B, self.n * self.m), np.float32); G_L = np.empty((self.B, self.n * self.m), np.float32)
        L = np.empty(self.B, np.float32)
        check(lib().gpad_plants_operators(self._h, layout, _f32p(M_G), _f32p(G_L), _f32p(L)), "gpad_plants_operators")
        return M_G, G_L, L

    def instances(self, params, want_f=False):
        params = np.ascontiguousarray(params, np.float64).reshape(self.B, self.n_par)
        g_P = np.empty((self.B, self.n), np.float32); p_D = np.empty((self.B, self.m), np.float32)
        f = np.empty((self.B, self.n), np.float32) if want_f else None
        check(lib().gpad_plants_instances(self._h, params.ctypes.data_as(_dp), _f32p(g_P), _f32p(p_D), _f32p(f) if want_f else None),
              "gpad_plants_instances")
        return g_P, p_D, f

    def closed_loop(self, solver, x0, samples, theta, beta, max_iter=None, warm_start=WARM_COLD, first=0, count=0):
        x0 = np.ascontiguousarray(np.atleast_2d(x0), np.float64)
        B = x0.shape[0]
        theta = np.ascontiguousarray(theta, np.float32); beta = np.ascontiguousarray(beta, np.float32)
        max_iter = len(theta) if max_iter is None else max_iter
        xt = np.empty((samples + 1, B, self.n_u)); ut = np.empty((samples, B, self.n_u))
        check(lib().gpad_closed_loop_plants(self._h, solver._h, first, count, x0.ctypes.data_as(_dp), samples, _f32p(theta), _f32p(beta),
                                            max_iter, int(warm_start), xt.ctypes.data_as(_dp), ut.ctypes.data_as(_dp)),
              "gpad_closed_loop_plants")
        return xt, ut

    def close(self):
        if self._h:
            lib().gpad_plants_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """gpad_group_*: one logical solver over several devices of one box (host-memory solves)."""

    def __init__(self, devices, n_u, N, m, L, M_G, G_L, layout=LAYOUT_SEQUENTIAL, mode=MODE_BATCH_SHARED, precision=PREC_TF32X3,
                 max_batch=1):
        cfg = Config(n_u, N, m, float(L), layout, mode, precision, max_batch, -1, MEM_HOST)
        M_G = np.ascontiguousarray(M_G, np.float32); G_L = np.ascontiguousarray(G_L, np.float32)
        dev = np.ascontiguousarray(devices, np.int32)
        self._keep = (M_G, G_L)
        self._h = C.c_void_p()
        check(lib().gpad_group_setup(C.byref(cfg), dev.ctypes.data_as(_ip), dev.size, _ptr(M_G), _ptr(G_L), C.byref(self._h)),
              "gpad_group_setup")
        self.n, self.m, self.size = n_u * N, m, dev.size

    def solve(self, args):
        check(lib().gpad_group_solve(self._h, C.byref(args)), "gpad_group_solve")

    def shard(self, i, batch):
        h, first, count = C.c_void_p(), C.c_int(), C.c_int()
        check(lib().gpad_group_shard(self._h, i, batch, C.byref(h), C.byref(first), C.byref(count)), "gpad_group_shard")
        return first.value, count.value

    def close(self):
        if self._h:
            lib().gpad_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def instances_device(problem, batch, params, g_P=None, p_D=None, f=None, stream=None):
    check(lib().gpad_instances_device(problem._h, batch, _ptr(params), _ptr(g_P), _ptr(p_D), _ptr(f), stream), "gpad_instances_device")


def closed_loop(problem, solver, x0, samples, theta, beta, max_iter=None, xref=None, warm_start=False):
    """receding-horizon simulation (gpad.m:79-95) -> (x_traj [samples+1][B][nx], u_traj [samples][B][n_u])"""
    x0 = np.ascontiguousarray(np.atleast_2d(x0), np.float64)
    B, nx = x0.shape
    theta = np.ascontiguousarray(theta, np.float32); beta = np.ascontiguousarray(beta, np.float32)
    max_iter = len(theta) if max_iter is None else max_iter
    xr = None if xref is None else np.ascontiguousarray(np.atleast_2d(xref), np.float64)
    xt = np.empty((samples + 1, B, nx)); ut = np.empty((samples, B, problem.n_u))
    check(lib().gpad_closed_loop(problem._h, solver._h, B, x0.ctypes.data_as(_dp), None if xr is None else xr.ctypes.data_as(_dp),
                                 samples, _f32p(theta), _f32p(beta), max_iter, int(warm_start),
                                 xt.ctypes.data_as(_dp), ut.ctypes.data_as(_dp)), "gpad_closed_loop")
    return xt, ut


# ---- step shims (device pointers) ----
def step_one(y, y_prev, w, beta, m, stream=None):
    check(lib().gpad_step_one(_ptr(y), _ptr(y_prev), _ptr(w), beta, m, stream), "gpad_step_one")


def step_two(M_G, w, g_P, zhat, N, n_u, m, stream=None):
    check(lib().gpad_step_two(_ptr(M_G), _ptr(w), _ptr(g_P), _ptr(zhat), N, n_u, m, stream), "gpad_step_two")


def array_copy(dst, src, size, stream=None):
    check(lib().gpad_array_copy(_ptr(dst), _ptr(src), size, stream), "gpad_array_copy")


def step_three(theta, zhat, z, length, stream=None):
    check(lib().gpad_step_three(theta, _ptr(zhat), _ptr(z), length, stream), "gpad_step_three")


def step_four(G_L, y_vp1, w, p_D, zhat, N, n_u, m, max_threads=0, stream=None):
    check(lib().gpad_step_four(_ptr(G_L), _ptr(y_vp1), _ptr(w), _ptr(p_D), _ptr(zhat), N, n_u, m, max_threads, stream),
          "gpad_step_four")


def debug_gemm_tf32x3(A, B, Cout, M, N, K, stream=None):
    check(lib().gpad_debug_gemm_tf32x3(_ptr(A), _ptr(B), _ptr(Cout), M, N, K, stream), "gpad_debug_gemm_tf32x3")


def debug_gemm_f16x3(A, B, Cout, M, N, K, kernel=0, stream=None):
    check(lib().gpad_debug_gemm_f16x3(_ptr(A), _ptr(B), _ptr(Cout), M, N, K, kernel, stream), "gpad_debug_gemm_f16x3")


def debug_plan_tiles(kernel, ncols):
    """(bn, n_tiles, step, tmem_cols) of the batch kernels' column tiling; host only."""
    v = [C.c_int() for _ in range(4)]
    check(lib().gpad_debug_plan_tiles(kernel, ncols, *[C.byref(x) for x in v]), "gpad_debug_plan_tiles")
    return tuple(x.value for x in v)


# ---- reference data file (main.cu:29-67) ----
def file_write(path, n_u, N, m, L, M_G, g_P, G_L, p_D, theta, beta, flat=False):
    arrs = [np.ascontiguousarray(a, np.float32).ravel() for a in (M_G, g_P, G_L, p_D, theta, beta)]
    assert arrs[0].size == (N if flat else n_u * N) * m and arrs[2].size == arrs[0].size
    fd = FileData(n_u, N, m, arrs[4].size, float(L), *[_f32p(a) for a in arrs])
    fn = lib().gpad_file_write_flat if flat else lib().gpad_file_write
    check(fn(path.encode(), C.byref(fd)), "gpad_file_write")


def file_read(path, flat=False):
    fd = FileData()
    fn = lib().gpad_file_read_flat if flat else lib().gpad_file_read
    check(fn(path.encode(), C.byref(fd)), "gpad_file_read")
    n, m, it = fd.n_u * fd.N, fd.m, fd.num_iterations
    op = (fd.N if flat else n) * m
    take = lambda p, cnt: np.ctypeslib.as_array(p, shape=(cnt,)).copy()
    out = dict(n_u=fd.n_u, N=fd.N, m=m, num_iterations=it, L=fd.L, M_G=take(fd.M_G, op), g_P=take(fd.g_P, n),
               G_L=take(fd.G_L, op), p_D=take(fd.p_D, m), theta=take(fd.theta, max(it, 1))[:it],
               beta=take(fd.beta, max(it, 1))[:it])
    lib().gpad_file_free(C.byref(fd))
    return out


_FIXTURE_FIELDS = {"op": None, "w": "m", "g_P": "n", "p_D": "m", "zhat_in": "n", "z_prev": "n",
                   "prod": None, "sum": "m", "zhat_out": "n", "z_out": "n", "y_next": "m"}


def fixture_read(directory, step, flat=False):
    """step-2/3/4 fixture of the reference's harnesses (main_prof.cu:117-156,198-239; step3.cu:59-81) -> dict"""
    fx = Fixture()
    check(lib().gpad_fixture_read(directory.encode(), step, 1 if flat else 0, C.byref(fx)), "gpad_fixture_read")
    n, m = fx.n_u * fx.N, fx.m
    size = {"n": n, "m": m}
    out = dict(step=fx.step, n_u=fx.n_u, N=fx.N, m=m, flat=bool(fx.flat), theta=fx.theta)
    for name, kind in _FIXTURE_FIELDS.items():
        p = getattr(fx, name)
        if not p:
            continue
        cnt = size[kind] if kind else ((fx.N if fx.flat else n) * m if name == "op" else (n if step == 2 else m))
        out[name] = np.ctypeslib.as_array(p, shape=(cnt,)).copy()
    lib().gpad_fixture_free(C.byref(fx))
    return out


def fixture_write(directory, step, n_u, N, m, flat=False, theta=0.0, **vectors):
    keep = {k: np.ascontiguousarray(v, np.float32).ravel() for k, v in vectors.items()}
    fx = Fixture(step, n_u, N, m, 1 if flat else 0, float(theta))
    for k, v in keep.items():
        setattr(fx, k, _f32p(v))
    check(lib().gpad_fixture_write(directory.encode(), C.byref(fx)), "gpad_fixture_write")
