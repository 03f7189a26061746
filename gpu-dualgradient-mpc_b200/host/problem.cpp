// problem.cpp -- host-side condensing and precompute (C++ restatement of the reference's
// MATLAB offline stage; nothing here runs on the GPU and nothing here is on the hot path).
//
//   battery:   Code/MATLAB/gpad.m:4-85 (plant, weights, M_ab, M_ak, K, H, F, A_i, b_i) and
//              acceldualgrad.m:9-23 (L = ||H||_F^2, M_G, g_P, G_L, p_D), with a Cholesky solve
//              in place of inv(H).
//   quadrotor: general condensed form (paper section 3, eq. 1-4) of a
//              hover-linearised 12-state / 4-input model -- builder-defined, see DESIGN.md.
//   schedule:  theta/beta recursion, acceldualgrad.m:55-56 / paper eq. (8e).
//
// Sign convention of the C code: M_G = -H^-1 G' so that zhat = M_G w - g_P.
// Instance data are affine in a parameter vector p: f = Ff p, b = b0 + Bb p,
// g_P = H^-1 f = Kg p, p_D = -b / L.
#include <cmath>
#include <cstring>
#include <vector>

#include "gpad.h"

#include "problem_internal.h"

namespace {

using gpad::Mat;

Mat eye(int n) { Mat m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }

Mat matmul(const Mat& A, const Mat& B) {
    Mat C(A.r, B.c);
    for (int i = 0; i < A.r; ++i)
        for (int k = 0; k < A.c; ++k) {
            const double aik = A(i, k);
            if (aik == 0.0) continue;
            const double* brow = &B.a[(size_t)k * B.c];
            double* crow = &C.a[(size_t)i * C.c];
            for (int j = 0; j < B.c; ++j) crow[j] += aik * brow[j];
        }
    return C;
}

Mat transpose(const Mat& A) {
    Mat T(A.c, A.r);
    for (int i = 0; i < A.r; ++i)
        for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
    return T;
}

// in-place lower Cholesky factor of SPD H; returns false if not positive definite
bool cholesky(Mat& H) {
    const int n = H.r;
    for (int j = 0; j < n; ++j) {
        double d = H(j, j);
        for (int k = 0; k < j; ++k) d -= H(j, k) * H(j, k);
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        H(j, j) = d;
        for (int i = j + 1; i < n; ++i) {
            double s = H(i, j);
            const double* ri = &H.a[(size_t)i * n];
            const double* rj = &H.a[(size_t)j * n];
            for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
            H(i, j) = s / d;
        }
        for (int i = 0; i < j; ++i) H(i, j) = 0.0;
    }
    return true;
}

// X <- H^-1 X given the Cholesky factor Lc (X is n x k, solved column-block-wise, row-major)
void chol_solve(const Mat& Lc, Mat& X) {
    const int n = Lc.r, k = X.c;
    for (int i = 0; i < n; ++i) {            // forward: Lc Y = X
        double* xi = &X.a[(size_t)i * k];
        for (int p = 0; p < i; ++p) {
            const double l = Lc(i, p);
            if (l == 0.0) continue;
            const double* xp = &X.a[(size_t)p * k];
            for (int j = 0; j < k; ++j) xi[j] -= l * xp[j];
        }
        const double inv = 1.0 / Lc(i, i);
        for (int j = 0; j < k; ++j) xi[j] *= inv;
    }
    for (int i = n - 1; i >= 0; --i) {       // backward: Lc' Z = Y
        double* xi = &X.a[(size_t)i * k];
        for (int p = i + 1; p < n; ++p) {
            const double l = Lc(p, i);
            if (l == 0.0) continue;
            const double* xp = &X.a[(size_t)p * k];
            for (int j = 0; j < k; ++j) xi[j] -= l * xp[j];
        }
        const double inv = 1.0 / Lc(i, i);
        for (int j = 0; j < k; ++j) xi[j] *= inv;
    }
}

}  // namespace

namespace {

// shared tail: Cholesky, M_G, Kg
int finish(gpad_problem_s* P) {
    Mat Lc = P->H;
    if (!cholesky(Lc)) return GPAD_ERR_INVALID_ARG;
    Mat Gt = transpose(P->G);          // n x m
    chol_solve(Lc, Gt);                // H^-1 G'
    for (double& v : Gt.a) v = -v;
    P->MG = Gt;
    P->Kg = P->Ff;
    chol_solve(Lc, P->Kg);
    return GPAD_OK;
}

// S_x (nx N x nx): block i = A^(i+1); S_u (nx N x nu N): block (i,j) = A^(i-j) B, j <= i
void prediction_matrices(const Mat& A, const Mat& B, int N, Mat& Sx, Mat& Su) {
    const int nx = A.r, nu = B.c;
    std::vector<Mat> pw(N + 1);
    pw[0] = eye(nx);
    for (int i = 1; i <= N; ++i) pw[i] = matmul(pw[i - 1], A);
    std::vector<Mat> pb(N);
    for (int i = 0; i < N; ++i) pb[i] = matmul(pw[i], B);
    Sx = Mat(nx * N, nx);
    Su = Mat(nx * N, nu * N);
    for (int i = 0; i < N; ++i) {
        for (int r = 0; r < nx; ++r)
            for (int c = 0; c < nx; ++c) Sx(i * nx + r, c) = pw[i + 1](r, c);
        for (int j = 0; j <= i; ++j)
            for (int r = 0; r < nx; ++r)
                for (int c = 0; c < nu; ++c) Su(i * nx + r, j * nu + c) = pb[i - j](r, c);
    }
}

// H = Su' Qbar Su + Rbar with diagonal stage weights q (nx), r (nu); also SuTQ = Su' Qbar
void hessian(const Mat& Su, const std::vector<double>& q, const std::vector<double>& r, int N, Mat& H, Mat& SuTQ) {
    const int nx = (int)q.size(), nu = (int)r.size();
    SuTQ = transpose(Su);
    for (int i = 0; i < SuTQ.r; ++i)
        for (int j = 0; j < SuTQ.c; ++j) SuTQ(i, j) *= q[j % nx];
    H = matmul(SuTQ, Su);
    for (int i = 0; i < nu * N; ++i) H(i, i) += r[i % nu];
    for (int i = 0; i < H.r; ++i)
        for (int j = i + 1; j < H.c; ++j) { const double s = 0.5 * (H(i, j) + H(j, i)); H(i, j) = s; H(j, i) = s; }
}

// lambda_max(G H^-1 G') = lambda_max(H^-1 G'G) by 400 power iterations from a fixed start vector (paper section 4)
bool lambda_max_dual(const gpad_problem_s* P, double* out) {
    const int n = P->n;
    Mat Lc = P->H;
    if (!cholesky(Lc)) return false;
    Mat T = matmul(transpose(P->G), P->G);
    chol_solve(Lc, T);
    std::vector<double> v(n, 1.0 / std::sqrt((double)n)), u(n);
    double lam = 0.0;
    for (int it = 0; it < 400; ++it) {
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            const double* row = &T.a[(size_t)i * n];
            for (int j = 0; j < n; ++j) s += row[j] * v[j];
            u[i] = s;
        }
        double nrm = 0.0;
        for (double x : u) nrm += x * x;
        lam = std::sqrt(nrm);
        if (!(lam > 0.0)) return false;
        for (int i = 0; i < n; ++i) v[i] = u[i] / lam;
    }
    *out = lam;
    return true;
}

}  // namespace

namespace gpad {

// battery balancing with cell capacities c_i = 0.027 * 4.1 Ah * cap_scale[i] (cap_scale == nullptr: the reference's pack)
int build_battery(int n_u, int N, const double* cap_scale, gpad_problem_s** out) {
    if (!out || n_u < 1 || N < 1) return GPAD_ERR_INVALID_ARG;
    gpad_problem_s* P = new gpad_problem_s;
    const int n = n_u * N;
    P->n_u = n_u; P->N = N; P->n = n; P->m = 4 * n + 2 * N; P->n_par = n_u; P->nx = n_u;
    // plant: A = I, B = diag(-1/(3600 c_i)), c_i = 0.027*4.1 Ah         gpad.m:18,34-35,47-49
    P->A = eye(n_u);
    P->B = Mat(n_u, n_u);
    for (int i = 0; i < n_u; ++i) {
        const double scale = cap_scale ? cap_scale[i] : 1.0;
        if (!(scale > 0.0)) { delete P; return GPAD_ERR_INVALID_ARG; }
        P->B(i, i) = -1.0 / (3600.0 * (0.027 * 4.1 * scale));
    }
    Mat Sx, Su;                                                        // M_ak, M_ab gpad.m:50-63
    prediction_matrices(P->A, P->B, N, Sx, Su);
    std::vector<double> q(n_u, 100.0), r(n_u, 1.0);                    // Qx, Qu     gpad.m:36-43
    Mat SuTQ;
    hessian(Su, q, r, N, P->H, SuTQ);                                  // H          gpad.m:76
    P->Ff = matmul(SuTQ, Sx);                                          // F' : f = (x0' F)'  gpad.m:77,81
    // A_i = [M_ab; -M_ab; I; -I; K; -K], K(i,j) = 1 iff stage(j) == i  gpad.m:65-73,84
    const int m = P->m;
    P->G = Mat(m, n);
    P->Bb = Mat(m, n_u);
    P->b0.assign(m, 0.0);
    const double xmax = 0.5, xmin = -0.5, umax = 0.3, umin = -0.3;     // gpad.m:30-33
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) { P->G(i, j) = Su(i, j); P->G(n + i, j) = -Su(i, j); }
        P->G(2 * n + i, i) = 1.0;
        P->G(3 * n + i, i) = -1.0;
        P->b0[i] = xmax; P->b0[n + i] = -xmin; P->b0[2 * n + i] = umax; P->b0[3 * n + i] = -umin;   // gpad.m:85
        for (int c = 0; c < n_u; ++c) { P->Bb(i, c) = -Sx(i, c); P->Bb(n + i, c) = Sx(i, c); }
    }
    for (int s = 0; s < N; ++s)
        for (int u = 0; u < n_u; ++u) { P->G(4 * n + s, s * n_u + u) = 1.0; P->G(4 * n + N + s, s * n_u + u) = -1.0; }
    double fro = 0.0;                                                  // L = ||H||_F^2  acceldualgrad.m:11
    for (double v : P->H.a) fro += v * v;
    P->L = fro;
    // dual blocks for the receding-horizon shift: {offset, rows per stage}
    P->blocks = {{0, n_u}, {n, n_u}, {2 * n, n_u}, {3 * n, n_u}, {4 * n, 1}, {4 * n + N, 1}};
    const int rc = finish(P);
    if (rc != GPAD_OK) { delete P; return rc; }
    *out = P;
    return GPAD_OK;
}

}  // namespace gpad

extern "C" {

int gpad_problem_battery(int n_u, int N, gpad_problem_t* out) { return gpad::build_battery(n_u, N, nullptr, out); }

int gpad_problem_quadrotor(int N, gpad_problem_t* out) {
    if (!out || N < 1) return GPAD_ERR_INVALID_ARG;
    gpad_problem_s* P = new gpad_problem_s;
    const int nx = 12, nu = 4;
    const double dt = 0.05, grav = 9.81, mass = 1.0, J[3] = {0.01, 0.01, 0.02};
    const std::vector<double> q = {10, 10, 10, 1, 1, 1, 5, 5, 1, 0.1, 0.1, 0.1};
    const std::vector<double> r = {0.1, 10.0, 10.0, 10.0};
    const double u_max[4] = {6.0, 0.3, 0.3, 0.15};
    const double vel_max = 2.0, tilt_max = 0.35, rate_max = 3.0, kappa = 0.02, poly_c = 0.35;
    const int sel[6] = {3, 4, 5, 6, 7, 11};
    const double smax[6] = {vel_max, vel_max, vel_max, tilt_max, tilt_max, rate_max};
    // continuous hover linearisation, exact ZOH (A is nilpotent, the series terminates)
    Mat Ac(nx, nx), Bc(nx, nu);
    for (int i = 0; i < 3; ++i) { Ac(i, 3 + i) = 1.0; Ac(6 + i, 9 + i) = 1.0; Bc(9 + i, 1 + i) = 1.0 / J[i]; }
    Ac(3, 7) = grav; Ac(4, 6) = -grav; Bc(5, 0) = 1.0 / mass;
    Mat Ad = eye(nx), Bd(nx, nu), term = eye(nx);
    for (int k = 1; k < 8; ++k) {
        Mat tb = matmul(term, Bc);
        for (size_t i = 0; i < Bd.a.size(); ++i) Bd.a[i] += tb.a[i] * (dt / k);
        term = matmul(term, Ac);
        for (double& v : term.a) v *= dt / k;
        for (size_t i = 0; i < Ad.a.size(); ++i) Ad.a[i] += term.a[i];
    }
    P->A = Ad; P->B = Bd;
    const int n = nu * N, ns = 6 * N, m = 2 * ns + 2 * n + 4 * N;
    P->n_u = nu; P->N = N; P->n = n; P->m = m; P->n_par = 2 * nx; P->nx = nx;
    Mat Sx, Su, SuTQ;
    prediction_matrices(Ad, Bd, N, Sx, Su);
    hessian(Su, q, r, N, P->H, SuTQ);
    // f = Su' Qbar (Sx x0 - 1 (x) xref) = Ff [x0; xref]
    Mat F1 = matmul(SuTQ, Sx);
    P->Ff = Mat(n, 2 * nx);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < nx; ++c) {
            P->Ff(i, c) = F1(i, c);
            double s = 0.0;
            for (int k = 0; k < N; ++k) s += SuTQ(i, k * nx + c);
            P->Ff(i, nx + c) = -s;
        }
    // G = [Es Su; -Es Su; I; -I; Ppoly], b = [smax - Es Sx x0; smax + Es Sx x0; umax; umax; c]
    P->G = Mat(m, n);
    P->Bb = Mat(m, 2 * nx);
    P->b0.assign(m, 0.0);
    for (int k = 0; k < N; ++k)
        for (int s = 0; s < 6; ++s) {
            const int row = k * 6 + s, src = k * nx + sel[s];
            for (int j = 0; j < n; ++j) { P->G(row, j) = Su(src, j); P->G(ns + row, j) = -Su(src, j); }
            for (int c = 0; c < nx; ++c) { P->Bb(row, c) = -Sx(src, c); P->Bb(ns + row, c) = Sx(src, c); }
            P->b0[row] = smax[s]; P->b0[ns + row] = smax[s];
        }
    for (int i = 0; i < n; ++i) {
        P->G(2 * ns + i, i) = 1.0; P->G(2 * ns + n + i, i) = -1.0;
        P->b0[2 * ns + i] = u_max[i % nu]; P->b0[2 * ns + n + i] = u_max[i % nu];
    }
    const double pst[4][4] = {{kappa, 1, 1, 0}, {kappa, 1, -1, 0}, {kappa, -1, 1, 0}, {kappa, -1, -1, 0}};
    for (int k = 0; k < N; ++k)
        for (int rr = 0; rr < 4; ++rr) {
            const int row = 2 * ns + 2 * n + k * 4 + rr;
            for (int c = 0; c < nu; ++c) P->G(row, k * nu + c) = pst[rr][c];
            P->b0[row] = poly_c;
        }
    // L = 1.02 lambda_max(G H^-1 G') by power iteration on T = H^-1 (G'G), fixed start vector
    {
        double lam = 0.0;
        if (!lambda_max_dual(P, &lam)) { delete P; return GPAD_ERR_INVALID_ARG; }
        P->L = 1.02 * lam;
    }
    P->blocks = {{0, 6}, {ns, 6}, {2 * ns, nu}, {2 * ns + n, nu}, {2 * ns + 2 * n, 4}};
    const int rc = finish(P);
    if (rc != GPAD_OK) { delete P; return rc; }
    *out = P;
    return GPAD_OK;
}

int gpad_problem_destroy(gpad_problem_t p) {
    if (!p) return GPAD_OK;
    for (void* d : p->dev_cache) gpad::problem_dev_free(d);
    delete p;
    return GPAD_OK;
}

int gpad_problem_set_lipschitz(gpad_problem_t p, int which, float* L_out) {
    if (!p) return GPAD_ERR_INVALID_ARG;
    double L = 0.0;
    if (which == GPAD_L_REFERENCE) {                  // L = ||H||_F^2, acceldualgrad.m:11
        for (double v : p->H.a) L += v * v;
    } else if (which == GPAD_L_LAMBDA_MAX) {          // L = 1.02 lambda_max(G H^-1 G'), paper section 4
        double lam = 0.0;
        if (!lambda_max_dual(p, &lam)) return GPAD_ERR_INVALID_ARG;
        L = 1.02 * lam;
    } else {
        return GPAD_ERR_INVALID_ARG;
    }
    p->L = L;
    if (L_out) *L_out = (float)L;
    return GPAD_OK;
}

int gpad_problem_dims(gpad_problem_t p, int* n_u, int* N, int* m, int* n_par, float* L) {
    if (!p) return GPAD_ERR_INVALID_ARG;
    if (n_u) *n_u = p->n_u;
    if (N) *N = p->N;
    if (m) *m = p->m;
    if (n_par) *n_par = p->n_par;
    if (L) *L = (float)p->L;
    return GPAD_OK;
}

int gpad_problem_operators(gpad_problem_t p, int layout, float* M_G, float* G_L) {
    if (!p || !M_G || !G_L) return GPAD_ERR_INVALID_ARG;
    const int n = p->n, m = p->m;
    const double invL = 1.0 / p->L;
    if (layout == GPAD_LAYOUT_SEQUENTIAL) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) M_G[(size_t)i * m + j] = (float)p->MG(i, j);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) G_L[(size_t)i * n + j] = (float)(p->G(i, j) * invL);     // acceldualgrad.m:22
    } else if (layout == GPAD_LAYOUT_FLIPPED) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) M_G[(size_t)j * n + i] = (float)p->MG(i, j);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) G_L[(size_t)j * m + i] = (float)(p->G(i, j) * invL);
    } else {
        return GPAD_ERR_INVALID_ARG;
    }
    return GPAD_OK;
}

int gpad_problem_instances(gpad_problem_t p, int B, const double* params, float* g_P, float* p_D, float* f) {
    if (!p || !params || B < 1) return GPAD_ERR_INVALID_ARG;
    const int n = p->n, m = p->m, np = p->n_par;
    const double invL = 1.0 / p->L;
    for (int b = 0; b < B; ++b) {
        const double* par = params + (size_t)b * np;
        if (g_P || f)
            for (int i = 0; i < n; ++i) {
                double sg = 0.0, sf = 0.0;
                for (int c = 0; c < np; ++c) { sg += p->Kg(i, c) * par[c]; sf += p->Ff(i, c) * par[c]; }
                if (g_P) g_P[(size_t)b * n + i] = (float)sg;                                     // acceldualgrad.m:21
                if (f) f[(size_t)b * n + i] = (float)sf;
            }
        if (p_D)
            for (int i = 0; i < m; ++i) {
                double s = p->b0[i];
                for (int c = 0; c < np; ++c) s += p->Bb(i, c) * par[c];
                p_D[(size_t)b * m + i] = (float)(-s * invL);                                     // acceldualgrad.m:23
            }
    }
    return GPAD_OK;
}

int gpad_problem_plant(gpad_problem_t p, int* nx, double* A, double* B) {
    if (!p) return GPAD_ERR_INVALID_ARG;
    if (nx) *nx = p->nx;
    if (A) std::memcpy(A, p->A.a.data(), sizeof(double) * p->A.a.size());
    if (B) std::memcpy(B, p->B.a.data(), sizeof(double) * p->B.a.size());
    return GPAD_OK;
}

int gpad_flatten_operators(int n_u, int N, int m, const float* MG, const float* GL, float* MGf, float* GLf, float* max_residual) {
    if (!MG || !GL || !MGf || !GLf || n_u < 1 || N < 1 || m < 4 * n_u * N) return GPAD_ERR_INVALID_ARG;
    const int n = n_u * N, box = 4 * n_u * N;
    float resid = 0.f;
    // M_G flat [N][m]: row s holds, for column k of the box part, the entry of dense row (s n_u + k % n_u);
    // the sum-constraint columns are shared by the n_u dense rows of stage s (seq_functions.cpp:5-20)
    for (int s = 0; s < N; ++s)
        for (int k = 0; k < m; ++k) {
            const int u_of_k = k < box ? k % n_u : 0;
            MGf[(size_t)s * m + k] = MG[(size_t)(s * n_u + u_of_k) * m + k];
            for (int u = 0; u < n_u; ++u) {
                const float dense = MG[(size_t)(s * n_u + u) * m + k];
                const float flat = (k >= box || u == k % n_u) ? MGf[(size_t)s * m + k] : 0.f;
                resid = std::fmax(resid, std::fabs(dense - flat));
            }
        }
    // G_L flat [m][N]: row i, stage s holds dense G_L[i][s n_u + i % n_u] (box rows) or the common value of
    // the n_u entries of stage s (sum-constraint rows)  (seq_functions.cpp:23-43)
    for (int i = 0; i < m; ++i)
        for (int s = 0; s < N; ++s) {
            const int u_of_i = i < box ? i % n_u : 0;
            GLf[(size_t)i * N + s] = GL[(size_t)i * n + s * n_u + u_of_i];
            for (int u = 0; u < n_u; ++u) {
                const float dense = GL[(size_t)i * n + s * n_u + u];
                const float flat = (i >= box || u == i % n_u) ? GLf[(size_t)i * N + s] : 0.f;
                resid = std::fmax(resid, std::fabs(dense - flat));
            }
        }
    if (max_residual) *max_residual = resid;
    return GPAD_OK;
}

int gpad_expand_operators(int n_u, int N, int m, const float* MGf, const float* GLf, float* MG, float* GL) {
    if (!MG || !GL || !MGf || !GLf || n_u < 1 || N < 1 || m < 4 * n_u * N) return GPAD_ERR_INVALID_ARG;
    const int n = n_u * N, box = 4 * n_u * N;
    for (int s = 0; s < N; ++s)
        for (int u = 0; u < n_u; ++u)
            for (int k = 0; k < m; ++k)
                MG[(size_t)(s * n_u + u) * m + k] = (k >= box || u == k % n_u) ? MGf[(size_t)s * m + k] : 0.f;
    for (int i = 0; i < m; ++i)
        for (int s = 0; s < N; ++s)
            for (int u = 0; u < n_u; ++u)
                GL[(size_t)i * n + s * n_u + u] = (i >= box || u == i % n_u) ? GLf[(size_t)i * N + s] : 0.f;
    return GPAD_OK;
}

int gpad_closed_loop(gpad_problem_t p, gpad_handle_t h, int B, const double* x0, const double* xref, int samples,
                     const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj, double* u_traj) {
    if (!p || !h || !x0 || !theta || !beta || B < 1 || samples < 1 || max_iter < 1) return GPAD_ERR_INVALID_ARG;
    if (warm_start < GPAD_WARM_COLD || warm_start > GPAD_WARM_SHIFTED) return GPAD_ERR_INVALID_ARG;
    if (p->n_par - p->nx > 0 && !xref) return GPAD_ERR_INVALID_ARG;
    // the handle must solve THIS problem: a mismatch would run gpad_solve past the ends of the loop's buffers
    int hn = 0, hN = 0, hm = 0, hmode = 0, hB = 0;
    if (gpad_handle_dims(h, &hn, &hN, &hm, &hmode, &hB, nullptr) != GPAD_OK) return GPAD_ERR_INVALID_ARG;
    if (hn != p->n_u || hN != p->N || hm != p->m || B > hB || (hmode == GPAD_MODE_LATENCY && B != 1)) return GPAD_ERR_INVALID_ARG;
    // the whole loop runs on the device (csrc/closed_loop.cu): instance build, solve, state advance
    return gpad::closed_loop_device(p, h, B, x0, xref, samples, theta, beta, max_iter, warm_start, x_traj, u_traj);
}

int gpad_schedule(float* theta, float* beta, int count, int variant) {
    if (!theta || !beta || count < 0) return GPAD_ERR_INVALID_ARG;
    double th_prev = 1.0, th = 1.0, lagged = 0.0;      // acceldualgrad.m:17,27
    for (int v = 0; v < count; ++v) {
        const double paper_beta = th * (1.0 / th_prev - 1.0);            // paper (8e)
        theta[v] = (float)th;
        beta[v] = (float)(variant == GPAD_SCHEDULE_MATLAB_LAG ? lagged : paper_beta);   // acceldualgrad.m:56 lag
        lagged = paper_beta;
        const double t2 = th * th;
        const double next = (std::sqrt(t2 * t2 + 4.0 * t2) - t2) / 2.0;  // acceldualgrad.m:55
        th_prev = th;
        th = next;
    }
    return GPAD_OK;
}

}  // extern "C"
