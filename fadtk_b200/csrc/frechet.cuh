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
// one Newton-Schulz iteration run side by side.  TM x TM tile / 256 threads, K step 16;
// TM = 32 for small d (more CTAs), 64 otherwise.
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

// one TM x TM tile of C = alpha A B + beta_diag I; returns max |C - I| over this thread's outputs and
// its share of tr C
template <int TM>
__device__ __forceinline__ void dgemm_tile(const double* __restrict__ A, const double* __restrict__ B,
                                           double* __restrict__ C, int d, double alpha, double beta_diag,
                                           float& dev, double& tr)
{
    constexpr int R = TM / 16;                                 // outputs per thread per dimension
    __shared__ double As[16][TM + 1], Bs[16][TM + 1];
    const int bi = blockIdx.y * TM, bj = blockIdx.x * TM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double c[R][R] = {};
    for (int k0 = 0; k0 < d; k0 += 16) {
        for (int i = threadIdx.x; i < TM * 16; i += 256) {
            const int r = i >> 4, k = i & 15;                  // A tile: TM rows x 16 k
            const int gi = bi + r, gk = k0 + k;
            As[k][r] = (gi < d && gk < d) ? A[(size_t)gi * d + gk] : 0.0;
        }
        for (int i = threadIdx.x; i < 16 * TM; i += 256) {
            const int k = i / TM, cc = i % TM;                 // B tile: 16 k x TM cols
            const int gk = k0 + k, gj = bj + cc;
            Bs[k][cc] = (gk < d && gj < d) ? B[(size_t)gk * d + gj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[R], b[R];
#pragma unroll
            for (int u = 0; u < R; ++u) { a[u] = As[k][ty * R + u]; b[u] = Bs[k][tx * R + u]; }
#pragma unroll
            for (int u = 0; u < R; ++u)
#pragma unroll
                for (int v = 0; v < R; ++v) c[u][v] = fma(a[u], b[v], c[u][v]);
        }
        __syncthreads();
    }
    tr = 0.0;
    dev = 0.0f;
#pragma unroll
    for (int u = 0; u < R; ++u)
#pragma unroll
        for (int v = 0; v < R; ++v) {
            const int gi = bi + ty * R + u, gj = bj + tx * R + v;
            if (gi < d && gj < d) {
                double val = alpha * c[u][v];
                if (gi == gj) { val += beta_diag; tr += val; }
                C[(size_t)gi * d + gj] = val;
                dev = fmaxf(dev, (float)fabs(val - (gi == gj ? 1.0 : 0.0)));
            }
        }
}

template <int TM>
__global__ void __launch_bounds__(256)
dgemm_kernel(const DgemmBatch batch, int d)
{
    if (batch.dev_in != nullptr && *batch.dev_in < batch.tol) return;      // already converged
    if (batch.dev_clear != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
        *batch.dev_clear = 0.0f;
    const DgemmProblem pr = batch.p[blockIdx.z];
    float dev;
    double tr;
    dgemm_tile<TM>(pr.A, pr.B, pr.C, d, pr.alpha, pr.beta_diag, dev, tr);
    if (batch.dev_out != nullptr) {
        for (int o = 16; o > 0; o >>= 1) dev = fmaxf(dev, __shfl_xor_sync(0xffffffffu, dev, o));
        if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(batch.dev_out), __float_as_uint(dev));
    }
    if (pr.trace_out != nullptr && blockIdx.x == blockIdx.y) {
        // diagonal blocks only; reduce inside the block, one atomic per block
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
