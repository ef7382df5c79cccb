// Frechet distance between two Gaussians on the GPU, fp64.
//
//   FAD = |mu1 - mu2|^2 + tr C1 + tr C2 - 2 tr sqrt(C1 C2)
//
// The reference (fadtk/fad.py:88-120) evaluates tr sqrt(C1 C2) through a non-symmetric
// eigen-decomposition V sqrt(D) V^-1 of C1 C2 (LAPACK dgeev + zgetri) - and scipy sqrtm for a
// warning.  Here the eigenvalues of C1 C2 are obtained from the similar SYMMETRIC PSD matrix
//   M = C1^(1/2) C2 C1^(1/2),      tr sqrt(C1 C2) = tr sqrt(M),
// and both square roots come from the coupled Newton-Schulz iteration
//   Y0 = A / |A|_F, Z0 = I;   W = 1.5 I - 0.5 Z Y;   Y <- Y W;   Z <- W Z;   Y -> (A/|A|_F)^(1/2)
// which is nothing but a chain of d x d x d GEMMs.  Zero eigenvalues (rank-deficient covariances,
// n < d) stay exactly zero in Y, so singular inputs need no eps-regularisation branch.
// C1^(1/2) of the baseline can be cached and reused by FAD-inf / per-song scoring.
//
// All scalars (norms, traces) stay on the device; the chain is launched without host syncs.
#pragma once
#include <stdint.h>

namespace fad {

// C = alpha * A * B + beta_diag * I   (row-major d x d, fp64), optional trace(C) accumulation.
// Up to two independent problems per launch (blockIdx.z): the Y <- Y W and Z <- W Z updates of
// one Newton-Schulz iteration run side by side.  64 x 32 tile / 256 threads (dgemm_tile below).
//
// Convergence control without host round trips: a launch whose `dev_in` (max |W - I| of the previous
// iteration) is below `tol` returns immediately, so a fixed-length launch sequence costs only
// launch latency once the iteration has converged.  `dev_out` receives max |C - I| (float bits,
// atomicMax), `dev_clear` is zeroed for a later iteration (three rotating slots, see host code).
struct DgemmProblem { const double* A; const double* B; double* C; double alpha, beta_diag; double* trace_out; };
struct DgemmBatch {
    DgemmProblem p[2];
    const float* dev_in; float* dev_out; float* dev_clear; float tol;
};

// One 64 x 32 tile of C = alpha A B + beta_diag I per 256-thread CTA; returns max |C - I| over this thread's
// outputs and its share of tr C (both zero for threads that hold no final outputs).
// Layout: 4 k-groups x 2 warps.  A warp owns 32 x 32 outputs (4 x 8 threads of 8 x 4 each: 32 DFMAs per
// 6 LDS.128, so the fp64 pipe and not shared memory is the limit); the four k-groups split every 16-wide
// K chunk between them (in-CTA split-K keeps 8 warps per CTA busy although a d = 768 problem has only 288
// tiles) and are summed through shared memory in a fixed order.  Global loads of chunk c+1 are issued
// before the FMAs of chunk c (register staging, two smem buffers, one barrier per chunk).
constexpr int kDgTileM = 64, kDgTileN = 32, kDgKC = 16;

__device__ __forceinline__ void dgemm_tile(const double* __restrict__ A, const double* __restrict__ B,
                                           double* __restrict__ C, int d, double alpha, double beta_diag,
                                           float& dev, double& tr)
{
    __shared__ __align__(16) double As[2][kDgKC][kDgTileM + 2];
    __shared__ __align__(16) double Bs[2][kDgKC][kDgTileN + 2];
    const int bi = blockIdx.y * kDgTileM, bj = blockIdx.x * kDgTileN;
    const int t = threadIdx.x, kg = t >> 6, lane = t & 31;
    const int r0 = ((t >> 5) & 1) * 32 + (lane >> 3) * 8, c0 = (lane & 7) * 4;
    // loader roles: A chunk = 64 rows x 16 k (4 consecutive k per thread), B chunk = 16 k x 32 cols (2 cols per thread)
    const int la_row = t >> 2, la_k = (t & 3) * 4, lb_k = t >> 4, lb_c = (t & 15) * 2;
    const bool a_row_ok = bi + la_row < d;
    const double* a_src = A + (size_t)(bi + la_row) * d + la_k;
    double ra[4], rb[2];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) ra[e] = (a_row_ok && k0 + la_k + e < d) ? a_src[k0 + e] : 0.0;
        const bool kok = k0 + lb_k < d;
        const double* b_src = B + (size_t)(k0 + lb_k) * d + bj + lb_c;
        rb[0] = (kok && bj + lb_c < d) ? b_src[0] : 0.0;
        rb[1] = (kok && bj + lb_c + 1 < d) ? b_src[1] : 0.0;
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int e = 0; e < 4; ++e) As[buf][la_k + e][la_row] = ra[e];
        *reinterpret_cast<double2*>(&Bs[buf][lb_k][lb_c]) = make_double2(rb[0], rb[1]);
    };
    double c[8][4] = {};
    const int chunks = (d + kDgKC - 1) / kDgKC;
    fetch(0);
    stage(0);
    __syncthreads();
    for (int ch = 0; ch < chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < chunks) fetch((ch + 1) * kDgKC);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = kg * 4 + kk;
            double a[8], b[4];
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const double2 v = *reinterpret_cast<const double2*>(&As[buf][k][r0 + u]);
                a[u] = v.x; a[u + 1] = v.y;
            }
#pragma unroll
            for (int v = 0; v < 4; v += 2) {
                const double2 w = *reinterpret_cast<const double2*>(&Bs[buf][k][c0 + v]);
                b[v] = w.x; b[v + 1] = w.y;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) c[u][v] = fma(a[u], b[v], c[u][v]);
        }
        if (ch + 1 < chunks) stage(buf ^ 1);
        __syncthreads();
    }
    // sum the k-groups: groups 1..3 hand their partial tiles to group 0 one after the other
    double* red = &As[0][0][0];                                 // 64 x 32 doubles <= sizeof(As)
    static_assert(sizeof(double) * kDgTileM * kDgTileN <= sizeof(As), "reduction scratch");
    const int slot = (t & 63) * 32;                             // 32 doubles per thread, thread-major
    for (int g = 1; g < 4; ++g) {
        if (kg == g) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) red[slot + ((u * 4 + v + (t & 31)) & 31)] = c[u][v];
        }
        __syncthreads();
        if (kg == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) c[u][v] += red[slot + ((u * 4 + v + (t & 31)) & 31)];
        }
        __syncthreads();
    }
    tr = 0.0;
    dev = 0.0f;
    if (kg != 0) return;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int gi = bi + r0 + u, gj = bj + c0 + v;
            if (gi < d && gj < d) {
                double val = alpha * c[u][v];
                if (gi == gj) { val += beta_diag; tr += val; }
                C[(size_t)gi * d + gj] = val;
                dev = fmaxf(dev, (float)fabs(val - (gi == gj ? 1.0 : 0.0)));
            }
        }
}

// grid (ceil(d / 32), ceil(d / 64), problems)
__global__ void __launch_bounds__(256, 2)
dgemm_kernel(const DgemmBatch batch, int d)
{
    if (batch.dev_in != nullptr && *batch.dev_in < batch.tol) return;      // already converged
    if (batch.dev_clear != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
        *batch.dev_clear = 0.0f;
    const DgemmProblem pr = blockIdx.z ? batch.p[1] : batch.p[0];   // no dynamically indexed copy of the parameter block
    float dev;
    double tr;
    dgemm_tile(pr.A, pr.B, pr.C, d, pr.alpha, pr.beta_diag, dev, tr);
    if (batch.dev_out != nullptr) {
        for (int o = 16; o > 0; o >>= 1) dev = fmaxf(dev, __shfl_xor_sync(0xffffffffu, dev, o));
        if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(batch.dev_out), __float_as_uint(dev));
    }
    if (pr.trace_out != nullptr && (blockIdx.x >> 1) == blockIdx.y) {
        // tiles that meet the diagonal only; reduce inside the block, one atomic per block
        __shared__ double red[256];
        red[threadIdx.x] = tr;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) atomicAdd(pr.trace_out, red[0]);
    }
}

// out[0] = tr A      (single block)
__global__ void trace_kernel(const double* __restrict__ A, int d, double* __restrict__ out)
{
    __shared__ double r[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < d; i += 256) t += A[(size_t)i * d + i];
    r[threadIdx.x] = t;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) r[threadIdx.x] += r[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = r[0];
}

// scal[0] = |A|_F, scal[1] = tr A     (single block of 1024 threads: fixed summation order)
__global__ void __launch_bounds__(1024) norm_trace_kernel(const double* __restrict__ A, int d, double* __restrict__ scal)
{
    __shared__ double r1[1024], r2[1024];
    double s0 = 0.0, s1 = 0.0, t = 0.0;
    const size_t total = (size_t)d * d;
    size_t e = threadIdx.x;
    for (; e + 1024 < total; e += 2048) {                      // two independent chains per thread
        const double v0 = A[e], v1 = A[e + 1024];
        s0 += v0 * v0; s1 += v1 * v1;
    }
    if (e < total) { const double v = A[e]; s0 += v * v; }
    for (int i = threadIdx.x; i < d; i += 1024) t += A[(size_t)i * d + i];
    r1[threadIdx.x] = s0 + s1; r2[threadIdx.x] = t;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if (threadIdx.x < k) { r1[threadIdx.x] += r1[threadIdx.x + k]; r2[threadIdx.x] += r2[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { scal[0] = sqrt(r1[0]); scal[1] = r2[0]; }
}

// Rank-deficient covariances (n < d) have eigenvalues that are zero up to roundoff, i.e. possibly
// slightly NEGATIVE, and Newton-Schulz diverges on a negative eigenvalue.  The iteration therefore
// runs on An + kNsDelta I (An = A/|A|_F), and the trace is corrected analytically:
//   tr sqrt(An) ~= tr Y - kNsDelta tr Z,      Y -> (An + dI)^(1/2),  Z -> (An + dI)^(-1/2)
// which is exact for null directions (sqrt(d) - d/sqrt(d) = 0) and O(d) ~ 1e-13 elsewhere.
constexpr double kNsDelta = 1e-13;

// Y = (A + A^T) / (2 |A|_F) + delta I, Z = I      (|A|_F read from scal[0])
__global__ void ns_init_kernel(const double* __restrict__ A, int d, const double* __restrict__ scal,
                               double* __restrict__ Y, double* __restrict__ Z)
{
    const double nrm = scal[0];
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)d * d;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        Y[e] = 0.5 * (A[e] + A[(size_t)j * d + i]) * inv + ((i == j) ? kNsDelta : 0.0);
        Z[e] = (i == j) ? 1.0 : 0.0;
    }
}

// S = sqrt(|A|_F) * Y        (un-normalise a converged square root)
__global__ void ns_unscale_kernel(const double* __restrict__ Y, int d, const double* __restrict__ scal,
                                  double* __restrict__ S)
{
    const double f = sqrt(scal[0]);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)d * d;
         e += (size_t)gridDim.x * blockDim.x) S[e] = f * Y[e];
}

// out[0] = fad, out[1] = tr sqrt(C1 C2), out[2] = relative residual |Y^2 - M/|M|_F|_F,
// out[3] = iterations, out[4] = |mu1-mu2|^2, out[5] = tr C1, out[6] = tr C2
__global__ void frechet_assemble_kernel(const double* __restrict__ mu1, const double* __restrict__ mu2, int d,
                                        const double* __restrict__ scalA /*|C1|_F, trC1*/,
                                        const double* __restrict__ trC2,
                                        const double* __restrict__ scalM /*|M|_F, trM*/,
                                        const double* __restrict__ trY, const double* __restrict__ trZ,
                                        const double* __restrict__ resid,
                                        int iters, double* __restrict__ out)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < d; i += 256) { const double df = mu1[i] - mu2[i]; s += df * df; }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double tr_sqrt = sqrt(scalM[0]) * (trY[0] - kNsDelta * trZ[0]);
        out[0] = red[0] + scalA[1] + trC2[1] - 2.0 * tr_sqrt;
        out[1] = tr_sqrt;
        out[2] = resid ? sqrt(resid[0]) : 0.0;
        out[3] = (double)iters;
        out[4] = red[0];
        out[5] = scalA[1];
        out[6] = trC2[1];
    }
}

// resid[0] = | Y Y - Mn |_F^2 where Mn = sym(M)/|M|_F : computed as a fused GEMM epilogue would
// be overkill here; one extra GEMM into T then this reduction.
__global__ void resid_kernel(const double* __restrict__ YY, const double* __restrict__ M, int d,
                             const double* __restrict__ scalM, double* __restrict__ resid)
{
    __shared__ double red[256];
    const double inv = scalM[0] > 0.0 ? 1.0 / scalM[0] : 0.0;
    double s = 0.0;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)d * d;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        const double m = 0.5 * (M[e] + M[(size_t)j * d + i]) * inv + ((i == j) ? kNsDelta : 0.0);
        const double r = YY[e] - m;
        s += r * r;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(resid, red[0]);
}

}  // namespace fad
