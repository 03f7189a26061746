// closed_loop.cu -- on-device instance build and device-resident receding-horizon simulation (gpad.m:79-95),
// SURVEY 8(f) rows 1 and 2.
//
// Instance build: the per-instance affine maps  g_P = Kg p,  f = Ff p,  p_D = -(b0 + Bb p) / L  with p = [x; x_ref]
// (acceldualgrad.m:21,23; gpad.m:81,85), one thread per output entry, fp64 and unfused like the host path
// (host/problem.cpp:gpad_problem_instances) so both round identically.  The problem's matrices are uploaded once per
// device and cached in the problem.
// Closed loop, per sample and for a whole batch of plants: (1) instance build from the current states, (2) gpad_solve on
// device buffers, warm-started from the previous duals (as they are, or shifted one stage) if asked, (3) u = z[0:n_u]
// is applied and x <- A x + B u advanced on the device.  Nothing but the trajectories crosses PCIe.  The plants may be
// one shared problem (gpad_closed_loop) or one plant per instance (gpad_closed_loop_plants, BASELINE config 5).
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "../host/plants_internal.h"
#include "gpad_internal.h"

namespace gpad {

namespace {

// Instance maps, one launch per output vector:
//   out[b][i] = (float) sum_c M[i][c] * par[b][c]                               (b0 == nullptr)
//   out[b][i] = (float) (-(b0[i] + sum_c M[i][c] * par[b][c]) * invL)           (b0 != nullptr)
// with par[b] = [x[b]; xref[b]] (np = nx + nref doubles), evaluated in fp64 with unfused multiply and add in the order
// c = 0 .. np-1, exactly like host/problem.cpp:gpad_problem_instances, so both builds agree bit for bit.
// mat_stride / invL_b: one matrix (and one 1/L) per instance when non-zero / non-null (per-instance plants).
// A CTA owns 128 output rows i and kAffTB batch rows at a time: the matrix tile sits transposed in shared memory
// ([c][i]: conflict-free, read once per CTA instead of once per output), the parameter rows are broadcast from shared
// memory, and every store is a 512-byte coalesced row segment.  (The first version, one thread per output with the
// matrix row read through 32-way strided loads, took 10.8 ms for the 64K quadrotor batch; this one is bound by the
// fp64 pipe and the 734 MB it writes.)
constexpr int kAffRows = 128, kAffTB = 32, kAffMaxNp = 32;

__global__ void __launch_bounds__(kAffRows)
affine_kernel(float* __restrict__ out, int ld, const double* __restrict__ M, size_t mat_stride, const double* __restrict__ b0,
              double invL, const double* __restrict__ invL_b, const double* __restrict__ x, int nx, const double* __restrict__ xref,
              int nref, int rows, int B) {
    __shared__ double Ms[kAffMaxNp][kAffRows + 1];
    __shared__ double Xs[kAffTB][kAffMaxNp];
    const int np = nx + nref;
    const int i0 = blockIdx.x * kAffRows, i = i0 + threadIdx.x;
    const bool shared_mat = mat_stride == 0;
    if (shared_mat)
        for (int k = threadIdx.x; k < kAffRows * np; k += kAffRows) {
            const int r = k / np, c = k % np;
            Ms[c][r] = i0 + r < rows ? M[(size_t)(i0 + r) * np + c] : 0.0;
        }
    const double bias = (b0 && i < rows) ? b0[i] : 0.0;
    for (int bb = blockIdx.y * kAffTB; bb < B; bb += gridDim.y * kAffTB) {
        const int nb = min(kAffTB, B - bb);
        __syncthreads();
        for (int k = threadIdx.x; k < nb * np; k += kAffRows) {
            const int r = k / np, c = k % np;
            Xs[r][c] = c < nx ? x[(size_t)(bb + r) * nx + c] : xref[(size_t)(bb + r) * nref + (c - nx)];
        }
        __syncthreads();
        if (i >= rows) continue;
        for (int r = 0; r < nb; ++r) {
            const int b = bb + r;
            double s = bias;
            if (shared_mat) {
                for (int c = 0; c < np; ++c) s = __dadd_rn(s, __dmul_rn(Ms[c][threadIdx.x], Xs[r][c]));
            } else {
                const double* Mi = M + (size_t)b * mat_stride + (size_t)i * np;
                for (int c = 0; c < np; ++c) s = __dadd_rn(s, __dmul_rn(Mi[c], Xs[r][c]));
            }
            out[(size_t)b * ld + i] = b0 ? (float)(-s * (invL_b ? invL_b[b] : invL)) : (float)s;
        }
    }
}

// one matrix per instance (per-instance plants): every matrix entry is used once, so one thread per output reading its own
// row is already coalesced across the outputs of an instance
__global__ void affine_plants_kernel(float* __restrict__ out, int ld, const double* __restrict__ M, size_t mat_stride,
                                     const double* __restrict__ b0, const double* __restrict__ invL_b, const double* __restrict__ x,
                                     int nx, int rows, int B) {
    const size_t total = (size_t)B * rows;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / rows), i = (int)(idx % rows);
        const double* Mi = M + (size_t)b * mat_stride + (size_t)i * nx;
        double s = b0 ? b0[i] : 0.0;
        for (int c = 0; c < nx; ++c) s = __dadd_rn(s, __dmul_rn(Mi[c], x[(size_t)b * nx + c]));
        out[(size_t)b * ld + i] = b0 ? (float)(-s * invL_b[b]) : (float)s;
    }
}

inline dim3 affine_grid(int rows, int B) {
    const int gx = (rows + kAffRows - 1) / kAffRows;
    const int gy = std::max(1, std::min((B + kAffTB - 1) / kAffTB, (148 * 8 + gx - 1) / gx));
    return dim3(gx, gy);
}

// x <- A x + B u with u = z[b][0:nu]; records u and the new state; b_stride != 0: one input matrix per instance
__global__ void advance_kernel(double* __restrict__ x, const float* __restrict__ z, int n, const double* __restrict__ A,
                               const double* __restrict__ Bm, size_t b_stride, int nx, int nu, int B, double* __restrict__ u_out,
                               double* __restrict__ x_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xn[32];
    const double* xb = x + (size_t)b * nx;
    const double* Bb = Bm + (size_t)b * b_stride;
    const float* u = z + (size_t)b * n;
    for (int i = 0; i < nx; ++i) {
        double s = 0.0;
        for (int j = 0; j < nx; ++j) s = __dadd_rn(s, __dmul_rn(A[i * nx + j], xb[j]));
        for (int j = 0; j < nu; ++j) s = __dadd_rn(s, __dmul_rn(Bb[i * nu + j], (double)u[j]));
        xn[i] = s;
    }
    for (int i = 0; i < nx; ++i) { x[(size_t)b * nx + i] = xn[i]; if (x_out) x_out[(size_t)b * nx + i] = xn[i]; }
    if (u_out) for (int j = 0; j < nu; ++j) u_out[(size_t)b * nu + j] = (double)u[j];
}

// receding-horizon shift of a dual vector: inside every constraint block {offset, rows per stage} stage s takes the
// multipliers of stage s + 1, the last stage keeps its own
__global__ void shift_duals_kernel(float* __restrict__ dst, const float* __restrict__ src, int m, int B, const int* __restrict__ blocks,
                                   int nblocks, int N) {
    const size_t total = (size_t)B * m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / m), i = (int)(idx % m);
        int from = i;
        for (int k = 0; k < nblocks; ++k) {
            const int off = blocks[2 * k], r = blocks[2 * k + 1];
            if (i >= off && i < off + r * N) {
                const int s = (i - off) / r, j = (i - off) % r;
                from = off + min(s + 1, N - 1) * r + j;
                break;
            }
        }
        dst[idx] = src[(size_t)b * m + from];
    }
}

struct DevBufs {
    std::vector<void*> p;
    template <typename T> int alloc(T** out, size_t count) {
        void* q = nullptr;
        if (cudaMalloc(&q, (count ? count : 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); set_error("closed loop: cudaMalloc failed"); return GPAD_ERR_ALLOC; }
        p.push_back(q); *out = static_cast<T*>(q); return GPAD_OK;
    }
    template <typename T> int upload(T** out, const T* src, size_t count, cudaStream_t s) {
        int rc = alloc(out, count);
        if (rc != GPAD_OK) return rc;
        GPAD_CUDA(cudaMemcpyAsync(*out, src, sizeof(T) * count, cudaMemcpyHostToDevice, s));
        return GPAD_OK;
    }
    ~DevBufs() { for (void* q : p) cudaFree(q); }
};

// the loop's working buffers: one device allocation that persists with the problem / plants cache and only ever grows
// (cudaMalloc / cudaFree of ~100 MB per call cost tens to hundreds of milliseconds, more than the loop itself)
struct Workspace {
    unsigned char* base = nullptr;
    size_t cap = 0, used = 0;
    int reserve(size_t bytes) {
        used = 0;
        if (bytes <= cap) return GPAD_OK;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        if (cudaMalloc(reinterpret_cast<void**>(&base), bytes) != cudaSuccess) { cudaGetLastError(); set_error("closed loop: cudaMalloc of %zu bytes failed", bytes); return GPAD_ERR_ALLOC; }
        cap = bytes;
        return GPAD_OK;
    }
    template <typename T> T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used += (count * sizeof(T) + 255) / 256 * 256;
        return p;
    }
    static size_t need(size_t count, size_t elem) { return (count * elem + 255) / 256 * 256; }
    ~Workspace() { if (base) cudaFree(base); }
};

struct DeviceScope {             // the handle's device for the duration of the loop, the caller's afterwards
    int prev = -1;
    cudaError_t status = cudaSuccess;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev) status = cudaSetDevice(dev);
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

struct StreamScope {
    cudaStream_t s = nullptr;
    cudaError_t status;
    StreamScope() { status = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); }
    ~StreamScope() { if (s) cudaStreamDestroy(s); }
};

// device copies of a problem's instance maps and plant, one per device
struct ProblemDev {
    int device = -1;
    DevBufs mem;
    double *Kg = nullptr, *Ff = nullptr, *Bb = nullptr, *b0 = nullptr, *A = nullptr, *B = nullptr;
    int* blocks = nullptr;
    Workspace ws;
};

int problem_dev_get(gpad_problem_t p, int device, ProblemDev** out) {
    std::lock_guard<std::mutex> lock(p->dev_mutex);
    for (void* c : p->dev_cache)
        if (static_cast<ProblemDev*>(c)->device == device) { *out = static_cast<ProblemDev*>(c); return GPAD_OK; }
    ProblemDev* d = new ProblemDev;
    d->device = device;
    std::vector<int> blk;
    for (auto& b : p->blocks) { blk.push_back(b.first); blk.push_back(b.second); }
    int rc = GPAD_OK;
    do {
        if ((rc = d->mem.upload(&d->Kg, p->Kg.a.data(), p->Kg.a.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->Ff, p->Ff.a.data(), p->Ff.a.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->Bb, p->Bb.a.data(), p->Bb.a.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->b0, p->b0.data(), p->b0.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->A, p->A.a.data(), p->A.a.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->B, p->B.a.data(), p->B.a.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->blocks, blk.data(), blk.size(), nullptr)) != GPAD_OK) break;
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) { cudaGetLastError(); rc = GPAD_ERR_CUDA; }
    } while (0);
    if (rc != GPAD_OK) { delete d; return rc; }
    p->dev_cache.push_back(d);
    *out = d;
    return GPAD_OK;
}

inline int grid_for(size_t total) { return (int)std::min<size_t>((total + 255) / 256, 148 * 32); }

// the loop both entry points share; `pl` != nullptr: one plant per instance (plants [first, first + B))
int run_loop(gpad_handle_t h, int B, int nx, int nu, int n, int m, int npar, int N, const double* dKg, size_t kg_stride,
             const double* dBb, size_t bb_stride, const double* db0, double invL, const double* dinvL, const double* dA,
             const double* dBm, size_t bm_stride, const int* dblocks, int nblocks, const double* x0, const double* xref,
             int samples, const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj, double* u_traj,
             Workspace& ws, cudaStream_t s) {
    if (nx > 32 || npar > kAffMaxNp) { set_error("closed loop: nx = %d / n_par = %d > 32 is not supported", nx, npar); return GPAD_ERR_UNSUPPORTED; }
    const int nref = npar - nx;
    const size_t Bn = (size_t)B * n, Bm = (size_t)B * m;
    const size_t n_ut = u_traj ? (size_t)samples * B * nu : 0, n_xt = x_traj ? (size_t)(samples + 1) * B * nx : 0;
    const bool shifted = warm_start == GPAD_WARM_SHIFTED;
    size_t bytes = Workspace::need((size_t)B * nx, 8) + Workspace::need((size_t)B * std::max(nref, 1), 8) + Workspace::need(n_ut, 8) +
                   Workspace::need(n_xt, 8) + 2 * Workspace::need(Bn, 4) + (size_t)(5 + (shifted ? 2 : 0)) * Workspace::need(Bm, 4);
    GPAD_TRY(ws.reserve(bytes));
    double* dx = ws.take<double>((size_t)B * nx);
    double* dxref = ws.take<double>((size_t)B * std::max(nref, 1));
    double* dut = ws.take<double>(n_ut);
    double* dxt = ws.take<double>(n_xt);
    float* gP = ws.take<float>(Bn); float* z = ws.take<float>(Bn); float* pD = ws.take<float>(Bm);
    float *ya[2], *yb[2], *sa = nullptr, *sb = nullptr;
    for (int k = 0; k < 2; ++k) { ya[k] = ws.take<float>(Bm); yb[k] = ws.take<float>(Bm); }
    if (shifted) { sa = ws.take<float>(Bm); sb = ws.take<float>(Bm); }
    GPAD_CUDA(cudaMemcpyAsync(dx, x0, sizeof(double) * B * nx, cudaMemcpyHostToDevice, s));
    if (nref > 0) GPAD_CUDA(cudaMemcpyAsync(dxref, xref, sizeof(double) * B * nref, cudaMemcpyHostToDevice, s));
    else dxref = nullptr;
    if (x_traj) GPAD_CUDA(cudaMemcpyAsync(dxt, dx, sizeof(double) * B * nx, cudaMemcpyDeviceToDevice, s));
    for (int k = 0; k < samples; ++k) {
        if (kg_stride) {        // one plant per instance
            affine_plants_kernel<<<grid_for((size_t)B * n), 256, 0, s>>>(gP, n, dKg, kg_stride, nullptr, nullptr, dx, nx, n, B);             // gpad.m:81
            affine_plants_kernel<<<grid_for((size_t)B * m), 256, 0, s>>>(pD, m, dBb, bb_stride, db0, dinvL, dx, nx, m, B);                   // gpad.m:85
        } else {
            affine_kernel<<<affine_grid(n, B), kAffRows, 0, s>>>(gP, n, dKg, 0, nullptr, 0.0, nullptr, dx, nx, dxref, nref, n, B);           // gpad.m:81
            affine_kernel<<<affine_grid(m, B), kAffRows, 0, s>>>(pD, m, dBb, 0, db0, invL, nullptr, dx, nx, dxref, nref, m, B);              // gpad.m:85
        }
        GPAD_CUDA(cudaGetLastError());
        gpad_solve_args_t a{};
        a.batch = B; a.mem = GPAD_MEM_DEVICE; a.stream = s;
        a.g_P = gP; a.p_D = pD; a.theta = theta; a.beta = beta; a.max_iter = max_iter;
        const int cur = k & 1, prev = cur ^ 1;
        if (warm_start != GPAD_WARM_COLD && k > 0) {
            if (warm_start == GPAD_WARM_SHIFTED) {
                shift_duals_kernel<<<grid_for((size_t)B * m), 256, 0, s>>>(sa, ya[prev], m, B, dblocks, nblocks, N);
                shift_duals_kernel<<<grid_for((size_t)B * m), 256, 0, s>>>(sb, yb[prev], m, B, dblocks, nblocks, N);
                GPAD_CUDA(cudaGetLastError());
                a.y0 = sa; a.y_prev0 = sb;
            } else {
                a.y0 = ya[prev]; a.y_prev0 = yb[prev];
            }
        }
        a.z = z; a.y_next = ya[cur]; a.y = yb[cur];
        const int rc = gpad_solve(h, &a);                                                                       // gpad.m:90
        if (rc != GPAD_OK) return rc;
        advance_kernel<<<(B + 127) / 128, 128, 0, s>>>(dx, z, n, dA, dBm, bm_stride, nx, nu, B, u_traj ? dut + (size_t)k * B * nu : nullptr,
                                                       x_traj ? dxt + (size_t)(k + 1) * B * nx : nullptr);      // gpad.m:91-93
        GPAD_CUDA(cudaGetLastError());
    }
    if (u_traj) GPAD_CUDA(cudaMemcpyAsync(u_traj, dut, sizeof(double) * (size_t)samples * B * nu, cudaMemcpyDeviceToHost, s));
    if (x_traj) GPAD_CUDA(cudaMemcpyAsync(x_traj, dxt, sizeof(double) * (size_t)(samples + 1) * B * nx, cudaMemcpyDeviceToHost, s));
    GPAD_CUDA(cudaStreamSynchronize(s));
    return GPAD_OK;
}

}  // namespace

void problem_dev_free(void* cache) {
    ProblemDev* d = static_cast<ProblemDev*>(cache);
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaSetDevice(d->device);
    delete d;
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

int instances_device(gpad_problem_t p, int device, int B, const double* params_dev, float* g_P, int ld_g, float* p_D, int ld_p,
                     float* f, int ld_f, cudaStream_t s) {
    ProblemDev* d = nullptr;
    int rc = problem_dev_get(p, device, &d);
    if (rc != GPAD_OK) return rc;
    const int n = p->n, m = p->m, np = p->n_par;
    if (np > kAffMaxNp) { set_error("instance build: n_par = %d > %d is not supported", np, kAffMaxNp); return GPAD_ERR_UNSUPPORTED; }
    // the whole parameter row plays the role of x (nref = 0): the maps are [rows][n_par]
    if (g_P) affine_kernel<<<affine_grid(n, B), kAffRows, 0, s>>>(g_P, ld_g, d->Kg, 0, nullptr, 0.0, nullptr, params_dev, np, nullptr, 0, n, B);
    if (p_D) affine_kernel<<<affine_grid(m, B), kAffRows, 0, s>>>(p_D, ld_p, d->Bb, 0, d->b0, 1.0 / p->L, nullptr, params_dev, np, nullptr, 0, m, B);
    if (f) affine_kernel<<<affine_grid(n, B), kAffRows, 0, s>>>(f, ld_f, d->Ff, 0, nullptr, 0.0, nullptr, params_dev, np, nullptr, 0, n, B);
    GPAD_CUDA(cudaGetLastError());
    return GPAD_OK;
}

int closed_loop_device(gpad_problem_t p, gpad_handle_t h, int B, const double* x0, const double* xref, int samples,
                       const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj, double* u_traj) {
    int device = 0;
    GPAD_TRY_RC(gpad_handle_dims(h, nullptr, nullptr, nullptr, nullptr, nullptr, &device));
    DeviceScope dev(device);
    GPAD_CUDA(dev.status);
    StreamScope st;
    GPAD_CUDA(st.status);
    ProblemDev* d = nullptr;
    GPAD_TRY_RC(problem_dev_get(p, device, &d));
    return run_loop(h, B, p->nx, p->n_u, p->n, p->m, p->n_par, p->N, d->Kg, 0, d->Bb, 0, d->b0, 1.0 / p->L, nullptr, d->A, d->B, 0,
                    d->blocks, (int)p->blocks.size(), x0, xref, samples, theta, beta, max_iter, warm_start, x_traj, u_traj, d->ws, st.s);
}

// device copies of one shard of plants (uploaded once per (device, first, count): 1.6 KB of maps per plant)
struct PlantsDev {
    int device = -1, first = 0, count = 0;
    DevBufs mem;
    double *Kg = nullptr, *Bb = nullptr, *b0 = nullptr, *A = nullptr, *Bm = nullptr, *invL = nullptr;
    int* blocks = nullptr;
    Workspace ws;
};

void plants_dev_free(void* cache) {
    PlantsDev* d = static_cast<PlantsDev*>(cache);
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaSetDevice(d->device);
    delete d;
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
}

static int plants_dev_get(gpad_plants_t p, int device, int first, int count, PlantsDev** out) {
    std::lock_guard<std::mutex> lock(p->dev_mutex);
    for (void* c : p->dev_cache) {
        PlantsDev* d = static_cast<PlantsDev*>(c);
        if (d->device == device && d->first == first && d->count == count) { *out = d; return GPAD_OK; }
    }
    PlantsDev* d = new PlantsDev;
    d->device = device; d->first = first; d->count = count;
    const int n = p->n, m = p->m, np = p->n_par, nx = p->nx, nu = p->n_u;
    std::vector<double> invL(count);
    for (int b = 0; b < count; ++b) invL[b] = 1.0 / p->L[first + b];
    std::vector<int> blk;
    for (auto& b : p->blocks) { blk.push_back(b.first); blk.push_back(b.second); }
    int rc = GPAD_OK;
    do {
        if ((rc = d->mem.upload(&d->Kg, &p->Kg[(size_t)first * n * np], (size_t)count * n * np, nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->Bb, &p->Bb[(size_t)first * m * np], (size_t)count * m * np, nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->b0, p->b0.data(), p->b0.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->A, p->A.data(), p->A.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->Bm, &p->Bm[(size_t)first * nx * nu], (size_t)count * nx * nu, nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->invL, invL.data(), invL.size(), nullptr)) != GPAD_OK) break;
        if ((rc = d->mem.upload(&d->blocks, blk.data(), blk.size(), nullptr)) != GPAD_OK) break;
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) { cudaGetLastError(); rc = GPAD_ERR_CUDA; }     // invL / blk are locals
    } while (0);
    if (rc != GPAD_OK) { delete d; return rc; }
    p->dev_cache.push_back(d);
    *out = d;
    return GPAD_OK;
}

int closed_loop_plants_device(gpad_plants_t p, gpad_handle_t h, int first, int count, const double* x0, int samples,
                              const float* theta, const float* beta, int max_iter, int warm_start, double* x_traj,
                              double* u_traj) {
    int device = 0;
    GPAD_TRY_RC(gpad_handle_dims(h, nullptr, nullptr, nullptr, nullptr, nullptr, &device));
    DeviceScope dev(device);
    GPAD_CUDA(dev.status);
    StreamScope st;
    GPAD_CUDA(st.status);
    PlantsDev* d = nullptr;
    GPAD_TRY_RC(plants_dev_get(p, device, first, count, &d));
    const int n = p->n, m = p->m, np = p->n_par, nx = p->nx, nu = p->n_u;
    return run_loop(h, count, nx, nu, n, m, np, p->N, d->Kg, (size_t)n * np, d->Bb, (size_t)m * np, d->b0, 0.0, d->invL, d->A, d->Bm,
                    (size_t)nx * nu, d->blocks, (int)p->blocks.size(), x0, nullptr, samples, theta, beta, max_iter, warm_start, x_traj,
                    u_traj, d->ws, st.s);
}

}  // namespace gpad

extern "C" int gpad_instances_device(gpad_problem_t p, int B, const double* params, float* g_P, float* p_D, float* f, void* stream) {
    using namespace gpad;
    GPAD_REQUIRE(p && params && B >= 1 && (g_P || p_D || f), "gpad_instances_device: bad argument");
    int device = 0;
    GPAD_CUDA(cudaGetDevice(&device));
    return instances_device(p, device, B, params, g_P, p->n, p_D, p->m, f, p->n, static_cast<cudaStream_t>(stream));
}
