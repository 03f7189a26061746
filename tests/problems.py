"""numpy (float64) restatement of the reference's offline problem setup -- test-side mirror of the
C++ host condensing in gpu-dualgradient-mpc_b200/host/problem.cpp.

battery():   Code/MATLAB/gpad.m:4-85 + acceldualgrad.m:9-23
quadrotor(): general condensed form, paper section 3 eq. 1-4; builder-defined
             hover-linearised 12-state model, documented in DESIGN.md.

Operators follow the C convention: M_G = -H^-1 G' so that zhat = M_G w - g_P.
Instance data are affine in a parameter vector p:  f = Ff p,  b = b0 + Bb p,
g_P = H^-1 f,  p_D = -b/L.
"""
import numpy as np


class Problem:
    def __init__(self, name, n_u, N, H, G, L, Ff, b0, Bb):
        self.name, self.n_u, self.N = name, n_u, N
        self.n, self.m = H.shape[0], G.shape[0]
        self.H, self.G, self.L = H, G, float(L)
        self.Ff, self.b0, self.Bb = Ff, b0, Bb
        self.n_par = Ff.shape[1]
        Hinv = np.linalg.inv(H)
        self.M_G64 = -Hinv @ G.T            # [n][m] sequential layout (seq_functions.cpp:61)
        self.G_L64 = G / L                  # [m][n]                    (seq_functions.cpp:82)
        self.Kg = Hinv @ Ff                 # g_P = Kg p
        self.M_G = self.M_G64.astype(np.float32)
        self.G_L = self.G_L64.astype(np.float32)

    def instance(self, p):
        """p [n_par] or [B][n_par] -> (g_P, p_D, f) fp32, instance-major."""
        p = np.atleast_2d(np.asarray(p, np.float64))
        f = p @ self.Ff.T
        g_P = p @ self.Kg.T
        b = self.b0[None, :] + p @ self.Bb.T
        p_D = -b / self.L
        sq = (lambda a: a[0]) if p.shape[0] == 1 else (lambda a: a)
        return sq(g_P.astype(np.float32)), sq(p_D.astype(np.float32)), sq(f.astype(np.float32))

    # flipped layouts the reference kernels read (kernel_functions.cu:50,180)
    def M_G_flipped(self):
        return np.ascontiguousarray(self.M_G.T)   # [m][n]

    def G_L_flipped(self):
        return np.ascontiguousarray(self.G_L.T)   # [n][m]


def battery(n_u=3, N=4, cap_scale=None):
    n, p = n_u, N
    cap = 0.027 * 4.1 * np.ones(n)                         # gpad.m:18
    if cap_scale is not None:                              # per-instance plants (BASELINE config 5)
        cap = cap * np.asarray(cap_scale, float)
    A = np.eye(n)                                          # gpad.m:34
    Bm = np.diag(-1.0 / (3600.0 * cap))                    # gpad.m:47-49
    M_ak = np.vstack([np.linalg.matrix_power(A, i) for i in range(1, p + 1)])   # gpad.m:50-52
    M_ab = np.zeros((n * p, n * p))
    for i in range(p):
        for j in range(i + 1):
            M_ab[i * n:(i + 1) * n, j * n:(j + 1) * n] = np.linalg.matrix_power(A, i - j) @ Bm  # gpad.m:55-63
    K = np.kron(np.eye(p), np.ones((1, n)))                # gpad.m:65-73
    Mx, Mu = 100.0 * np.eye(n * p), np.eye(n * p)          # gpad.m:36-43
    H = M_ab.T @ Mx @ M_ab + Mu                            # gpad.m:76
    F = M_ak.T @ Mx @ M_ab                                 # gpad.m:77  (n x np); f = x0' F
    I = np.eye(n * p)
    G = np.vstack([M_ab, -M_ab, I, -I, K, -K])             # gpad.m:84
    xmax, xmin, umax, umin = 0.5, -0.5, 0.3, -0.3          # gpad.m:30-33
    b0 = np.concatenate([xmax * np.ones(n * p), -xmin * np.ones(n * p), umax * np.ones(n * p),
                         -umin * np.ones(n * p), np.zeros(p), np.zeros(p)])     # gpad.m:85
    Z = np.zeros((n * p, n))
    Bb = np.vstack([-M_ak, M_ak, Z, Z, np.zeros((p, n)), np.zeros((p, n))])
    L = np.linalg.norm(H, 'fro') ** 2                      # acceldualgrad.m:11
    return Problem("battery", n_u, N, H, G, L, F.T, b0, Bb)


BATTERY_X0_10 = np.array([-0.1, 0.45, -0.09, 0.05, 0, -0.05, 0.3, 0.2, 0.25, -0.45])  # gpad.m:10
BATTERY_X0_5 = np.array([-0.1, 0.05, 0, -0.05, 0.1])                                   # gpad.m:12


def battery_x0(n_u, rng):
    """gpad.m:9-15: fixed vectors for 10 and 5 cells, else U(-0.5, 0.5)."""
    if n_u == 10:
        return BATTERY_X0_10.copy()
    if n_u == 5:
        return BATTERY_X0_5.copy()
    return rng.random(n_u) - 0.5


# ---- quadrotor ------------------------------------------------------------------------------
QUAD = dict(nx=12, nu=4, dt=0.05, g=9.81, mass=1.0, J=(0.01, 0.01, 0.02),
            q=(10, 10, 10, 1, 1, 1, 5, 5, 1, 0.1, 0.1, 0.1), r=(0.1, 10.0, 10.0, 10.0),
            u_max=(6.0, 0.3, 0.3, 0.15), vel_max=2.0, tilt_max=0.35, rate_max=3.0,
            kappa=0.02, poly_c=0.35, power_iters=400, L_margin=1.02)


def quadrotor_dynamics():
    """ZOH discretisation of the hover-linearised model (nilpotent A: the series is exact)."""
    c = QUAD
    A = np.zeros((12, 12)); B = np.zeros((12, 4))
    A[0:3, 3:6] = np.eye(3)                 # p' = v
    A[3, 7] = c["g"]                        # vx' =  g * pitch
    A[4, 6] = -c["g"]                       # vy' = -g * roll
    A[6:9, 9:12] = np.eye(3)                # angles' = rates
    B[5, 0] = 1.0 / c["mass"]               # vz' = dT / m
    for i in range(3):
        B[9 + i, 1 + i] = 1.0 / c["J"][i]   # rates' = J^-1 tau
    dt = c["dt"]
    Ad = np.eye(12); Bd = np.zeros((12, 4))
    term = np.eye(12)                       # A^k dt^k / k!
    for k in range(1, 8):
        Bd += term @ B * (dt / k)
        term = term @ A * (dt / k)
        Ad += term
    return Ad, Bd


def quadrotor(N=100):
    c = QUAD
    nx, nu = 12, 4
    Ad, Bd = quadrotor_dynamics()
    Sx = np.zeros((nx * N, nx)); Su = np.zeros((nx * N, nu * N))
    Ap = np.eye(nx)
    powers = [np.eye(nx)]
    for i in range(N):
        powers.append(powers[-1] @ Ad)
    for i in range(N):                       # x_{i+1} = A^{i+1} x0 + sum_j A^{i-j} B u_j
        Sx[i * nx:(i + 1) * nx] = powers[i + 1]
        for j in range(i + 1):
            Su[i * nx:(i + 1) * nx, j * nu:(j + 1) * nu] = powers[i - j] @ Bd
    Qb = np.kron(np.eye(N), np.diag(c["q"])); Rb = np.kron(np.eye(N), np.diag(c["r"]))
    H = Su.T @ Qb @ Su + Rb
    H = 0.5 * (H + H.T)
    ones = np.kron(np.ones((N, 1)), np.eye(nx))
    Ff = np.hstack([Su.T @ Qb @ Sx, -Su.T @ Qb @ ones])        # f = Ff [x0; xref]
    # constrained states: velocities (3,4,5), roll/pitch (6,7), yaw rate (11)
    sel = [3, 4, 5, 6, 7, 11]
    smax = np.array([c["vel_max"]] * 3 + [c["tilt_max"]] * 2 + [c["rate_max"]])
    Es = np.kron(np.eye(N), np.eye(nx)[sel])
    I = np.eye(nu * N)
    k = c["kappa"]
    Pst = np.array([[k, 1, 1, 0], [k, 1, -1, 0], [k, -1, 1, 0], [k, -1, -1, 0]], float)
    Pp = np.kron(np.eye(N), Pst)
    G = np.vstack([Es @ Su, -Es @ Su, I, -I, Pp])
    umax = np.tile(np.array(c["u_max"]), N)
    b0 = np.concatenate([np.tile(smax, N), np.tile(smax, N), umax, umax, c["poly_c"] * np.ones(4 * N)])
    ns = len(sel) * N
    Bx = Es @ Sx
    Bb = np.zeros((G.shape[0], 2 * nx))
    Bb[:ns, :nx] = -Bx
    Bb[ns:2 * ns, :nx] = Bx
    # L = margin * lambda_max(G H^-1 G') by power iteration on H^-1 G'G (n x n), fixed start
    T = np.linalg.solve(H, G.T @ G)
    v = np.ones(H.shape[0]) / np.sqrt(H.shape[0])
    lam = 0.0
    for _ in range(c["power_iters"]):
        u = T @ v
        lam = np.linalg.norm(u)
        v = u / lam
    L = c["L_margin"] * lam
    return Problem("quadrotor", nu, N, H, G, L, Ff, b0, Bb)


def quadrotor_params(B, rng):
    """per-instance [x0 (12); xref (12)]: random start near hover, random position/yaw setpoint."""
    x0 = np.zeros((B, 12)); xr = np.zeros((B, 12))
    x0[:, 0:3] = rng.uniform(-1.0, 1.0, (B, 3))
    x0[:, 3:6] = rng.uniform(-1.0, 1.0, (B, 3))
    x0[:, 6:8] = rng.uniform(-0.2, 0.2, (B, 2))
    x0[:, 8] = rng.uniform(-0.5, 0.5, B)
    x0[:, 9:12] = rng.uniform(-0.5, 0.5, (B, 3))
    xr[:, 0:3] = rng.uniform(-2.0, 2.0, (B, 3))
    xr[:, 8] = rng.uniform(-0.5, 0.5, B)
    return np.hstack([x0, xr])


def rel_inf(a, b):
    """infinity-norm relative difference ||a-b||_inf / ||b||_inf (the tolerance metric of DESIGN.md)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
