// closed_loop.cu -- device-resident receding-horizon simulation (gpad.m:79-95), SURVEY 8(f) rows 1 and 2.
//
// Per sample, for a whole batch of plants: (1) the per-instance affine maps  g_P = Kg p,  p_D = -(b0 + Bb p) / L  with
// p = [x; x_ref] are evaluated on the device (acceldualgrad.m:21,23; one thread per output entry, fp64 like the host
// path, unfused so both paths round identically), (2) gpad_solve runs on device buffers, warm-started from the
// previous duals if asked, (3) u = z[0:n_u] is applied and x <- A x + B u advanced on the device.  Nothing but the
// state / input trajectories crosses PCIe.  Called by gpad_closed_loop (host/problem.cpp), which owns the problem data.
#include <cuda_runtime.h>

#include <vector>

#include "gpad_internal.h"

namespace gpad {

namespace {

// out[b][i] = (float) sum_c M[i][c] * par[b][c]      (rows = n)
// out[b][i] = (float) (-(b0[i] + sum_c M[i][c] * par[b][c]) * invL)      (b0 != nullptr)
__global__ void affine_kernel(float* __restrict__ out, const double* __restrict__ M, const double* __restrict__ b0, double invL,
                              const double* __restrict__ x, int nx, const double* __restrict__ xref, int nref, int rows, int B) {
    const size_t total = (size_t)B * rows;
    const int np = nx + nref;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / rows), i = (int)(idx % rows);
        const double* Mi = M + (size_t)i * np;
        double s = b0 ? b0[i] : 0.0;
        for (int c = 0; c < nx; ++c) s = __dadd_rn(s, __dmul_rn(Mi[c], x[(size_t)b * nx + c]));
        for (int c = 0; c < nref; ++c) s = __dadd_rn(s, __dmul_rn(Mi[nx + c], xref[(size_t)b * nref + c]));
        out[idx] = b0 ? (float)(-s * invL) : (float)s;
    }
}

// x <- A x + B u with u = z[b][0:nu]; records u and the new state
__global__ void advance_kernel(double* __restrict__ x, const float* __restrict__ z, int n, const double* __restrict__ A,
                               const double* __restrict__ Bm, int nx, int nu, int B, double* __restrict__ u_out,
                               double* __restrict__ x_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xn[32];
    const double* xb = x + (size_t)b * nx;
    const float* u = z + (size_t)b * n;
    for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int j = 0; j < nx; ++j) s = __dadd_rn(s, __dmul_rn(A[i * nx + j], xb[j]));
        for (int j = 0; j < nu; ++j) s = __dadd_rn(s, __dmul_rn(Bm[i * nu + j], (double)u[j]));
        xn[i] = s;
    }
    for (int i = 0; i < nx; ++i) { x[(size_t)b * nx + i] = xn[i]; if (x_out) x_out[(size_t)b * nx + i] = xn[i]; }
    if (u_out) for (int j = 0; j < nu; ++j) u_out[(size_t)b * nu + j] = (double)u[j];
}

struct DevBufs {
    std::vector<void*> p;
    template <typename T> int alloc(T** out, size_t count) {
        void* q = nullptr;
        if (cudaMalloc(&q, (count ? count : 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); set_error("closed loop: cudaMalloc failed"); return GPAD_ERR_ALLOC; }
        p.push_back(q); *out = static_cast<T*>(q); return GPAD_OK;
    }
    ~DevBufs() { for (void* q : p) cudaFree(q); }
};

}  // namespace

int closed_loop_device(gpad_handle_t h, int B, int nx, int nu, int n, int m, int npar, double L, const double* Kg,
                       const double* Bb, const double* b0, const double* A, const double* Bm, const double* x0,
                       const double* xref, int samples, const float* theta, const float* beta, int max_iter, int warm_start,
                       double* x_traj, double* u_traj) {
    if (nx > 32) { set_error("closed loop: nx = %d > 32 is not supported", nx); return GPAD_ERR_UNSUPPORTED; }
    const int nref = npar - nx;
    DevBufs d;
    double *dKg, *dBb, *db0, *dA, *dB, *dx, *dxref = nullptr, *dut, *dxt;
    float *gP, *pD, *z, *ya[2], *yb[2];
#define TRYA(e) do { int rc_ = (e); if (rc_ != GPAD_OK) return rc_; } while (0)
    TRYA(d.alloc(&dKg, (size_t)n * npar)); TRYA(d.alloc(&dBb, (size_t)m * npar)); TRYA(d.alloc(&db0, m));
    TRYA(d.alloc(&dA, (size_t)nx * nx)); TRYA(d.alloc(&dB, (size_t)nx * nu)); TRYA(d.alloc(&dx, (size_t)B * nx));
    if (nref > 0) TRYA(d.alloc(&dxref, (size_t)B * nref));
    TRYA(d.alloc(&dut, u_traj ? (size_t)samples * B * nu : 0)); TRYA(d.alloc(&dxt, x_traj ? (size_t)(samples + 1) * B * nx : 0));
    TRYA(d.alloc(&gP, (size_t)B * n)); TRYA(d.alloc(&pD, (size_t)B * m)); TRYA(d.alloc(&z, (size_t)B * n));
    for (int k = 0; k < 2; ++k) { TRYA(d.alloc(&ya[k], (size_t)B * m)); TRYA(d.alloc(&yb[k], (size_t)B * m)); }
#undef TRYA
    cudaStream_t s = nullptr;
    GPAD_CUDA(cudaMemcpyAsync(dKg, Kg, sizeof(double) * n * npar, cudaMemcpyHostToDevice, s));
    GPAD_CUDA(cudaMemcpyAsync(dBb, Bb, sizeof(double) * m * npar, cudaMemcpyHostToDevice, s));
    GPAD_CUDA(cudaMemcpyAsync(db0, b0, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    GPAD_CUDA(cudaMemcpyAsync(dA, A, sizeof(double) * nx * nx, cudaMemcpyHostToDevice, s));
    GPAD_CUDA(cudaMemcpyAsync(dB, Bm, sizeof(double) * nx * nu, cudaMemcpyHostToDevice, s));
    GPAD_CUDA(cudaMemcpyAsync(dx, x0, sizeof(double) * B * nx, cudaMemcpyHostToDevice, s));
    if (nref > 0) GPAD_CUDA(cudaMemcpyAsync(dxref, xref, sizeof(double) * B * nref, cudaMemcpyHostToDevice, s));
    if (x_traj) GPAD_CUDA(cudaMemcpyAsync(dxt, dx, sizeof(double) * B * nx, cudaMemcpyDeviceToDevice, s));
    const int grid_n = (int)std::min<size_t>(((size_t)B * n + 255) / 256, 148 * 32);
    const int grid_m = (int)std::min<size_t>(((size_t)B * m + 255) / 256, 148 * 32);
    for (int k = 0; k < samples; ++k) {
        affine_kernel<<<grid_n, 256, 0, s>>>(gP, dKg, nullptr, 0.0, dx, nx, dxref, nref, n, B);                 // gpad.m:81
        affine_kernel<<<grid_m, 256, 0, s>>>(pD, dBb, db0, 1.0 / L, dx, nx, dxref, nref, m, B);                 // gpad.m:85
        GPAD_CUDA(cudaGetLastError());
        gpad_solve_args_t a{};
        a.batch = B; a.mem = GPAD_MEM_DEVICE; a.stream = s;
        a.g_P = gP; a.p_D = pD; a.theta = theta; a.beta = beta; a.max_iter = max_iter;
        const int cur = k & 1, prev = cur ^ 1;
        if (warm_start && k > 0) { a.y0 = ya[prev]; a.y_prev0 = yb[prev]; }
        a.z = z; a.y_next = ya[cur]; a.y = yb[cur];
        const int rc = gpad_solve(h, &a);                                                                       // gpad.m:90
        if (rc != GPAD_OK) return rc;
        advance_kernel<<<(B + 127) / 128, 128, 0, s>>>(dx, z, n, dA, dB, nx, nu, B, u_traj ? dut + (size_t)k * B * nu : nullptr,
                                                       x_traj ? dxt + (size_t)(k + 1) * B * nx : nullptr);      // gpad.m:91-93
        GPAD_CUDA(cudaGetLastError());
    }
    if (u_traj) GPAD_CUDA(cudaMemcpyAsync(u_traj, dut, sizeof(double) * (size_t)samples * B * nu, cudaMemcpyDeviceToHost, s));
    if (x_traj) GPAD_CUDA(cudaMemcpyAsync(x_traj, dxt, sizeof(double) * (size_t)(samples + 1) * B * nx, cudaMemcpyDeviceToHost, s));
    GPAD_CUDA(cudaStreamSynchronize(s));
    return GPAD_OK;
}

}  // namespace gpad
