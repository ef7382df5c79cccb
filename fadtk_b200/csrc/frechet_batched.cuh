// Ragged-batched Frechet distance: many eval sets (songs) against ONE baseline, in lock-step.
//
// Replaces the per-file loop of FrechetAudioDistance.score_individual (fadtk/fad.py:353-395):
// for every eval file the reference computes calc_embd_statistics (fad.py:42-48: np.mean in the
// input dtype - fp16 -, np.cov in fp64) and calc_frechet_distance (fad.py:51-120) against the
// same baseline - 5 000 songs x ~0.6 s of LAPACK at BASELINE config 4.  Here item z owns rows
// [offsets[z], offsets[z+1]) of one fp16 [N, d] matrix and every stage is one launch over all
// items (blockIdx.z): statistics, M_z = S C_z S with S = C_base^(1/2) cached, the coupled
// Newton-Schulz iteration (frechet.cuh) with PER-ITEM convergence flags, traces, assembly.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "frechet.cuh"

namespace fad {

// mean (rounded to fp16 like np.mean of an fp16 array, returned as fp64) and covariance (ddof = 1,
// exact fp64 products of y = x - s, s = the item's first row) of item z.
// grid (ceil(d/32), ceil(d/32), items), 256 threads: thread (ty, tx) owns a 2 x 2 block of the tile.
// ok[z] = 0 for an item with fewer than 2 rows (the reference asserts, fad.py:46): its covariance is
// set to the identity so the lock-step chain stays finite, and the assembly writes NaN.
__global__ void __launch_bounds__(256)
song_stats_kernel(const __half* __restrict__ emb, const long long* __restrict__ offsets, int d,
                  double* __restrict__ mu /*[items][d]*/, double* __restrict__ cov /*[items][d][d]*/,
                  int* __restrict__ ok)
{
    __shared__ double Yi[32][33], Yj[32][33];
    __shared__ double si[32], sj[32];
    const int z = blockIdx.z;
    const long long r0 = offsets[z], r1 = offsets[z + 1];
    const long long n = r1 - r0;
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double* C = cov + (size_t)z * d * d;
    if (n < 2) {
        for (int e = threadIdx.x; e < 1024; e += 256) {
            const int gi = bi + (e >> 5), gj = bj + (e & 31);
            if (gi < d && gj < d) C[(size_t)gi * d + gj] = gi == gj ? 1.0 : 0.0;
        }
        if (bj == 0 && threadIdx.x < 32 && bi + threadIdx.x < d) mu[(size_t)z * d + bi + threadIdx.x] = 0.0;
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ok[z] = 0;
        return;
    }
    const __half* base = emb + (size_t)r0 * d;
    double c[2][2] = {};
    double colsum = 0.0;                                        // threads 0..31: column bi + t; 32..63: column bj + t
    for (long long rr = 0; rr < n; rr += 32) {
        for (int e = threadIdx.x; e < 2048; e += 256) {
            const int which = e >> 10, r = (e >> 5) & 31, cidx = e & 31;
            const int gc = (which ? bj : bi) + cidx;
            double v = 0.0;
            if (rr + r < n && gc < d)
                v = (double)__half2float(base[(size_t)(rr + r) * d + gc]) - (double)__half2float(base[gc]);
            (which ? Yj : Yi)[r][cidx] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const double a0 = Yi[r][ty * 2], a1 = Yi[r][ty * 2 + 1];
            const double b0 = Yj[r][tx * 2], b1 = Yj[r][tx * 2 + 1];
            c[0][0] = fma(a0, b0, c[0][0]); c[0][1] = fma(a0, b1, c[0][1]);
            c[1][0] = fma(a1, b0, c[1][0]); c[1][1] = fma(a1, b1, c[1][1]);
        }
        if (threadIdx.x < 64) {
            const int t = threadIdx.x & 31;
            double s = 0.0;
            if (threadIdx.x < 32) { for (int r = 0; r < 32; ++r) s += Yi[r][t]; }
            else                  { for (int r = 0; r < 32; ++r) s += Yj[r][t]; }
            colsum += s;
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) si[threadIdx.x] = colsum;
    else if (threadIdx.x < 64) sj[threadIdx.x - 32] = colsum;
    __syncthreads();
    const double inv_n = 1.0 / (double)n, inv_n1 = 1.0 / (double)(n - 1);
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int li = ty * 2 + u, lj = tx * 2 + v;
            const int gi = bi + li, gj = bj + lj;
            if (gi < d && gj < d) C[(size_t)gi * d + gj] = (c[u][v] - si[li] * sj[lj] * inv_n) * inv_n1;
        }
    if (bj == 0 && threadIdx.x < 32 && bi + threadIdx.x < d) {
        const double m = (double)__half2float(base[bi + threadIdx.x]) + si[threadIdx.x] * inv_n;
        mu[(size_t)z * d + bi + threadIdx.x] = (double)__half2float(__double2half(m));     // fp16 mean (fad.py:48)
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ok[z] = 1;
}

// The same per-item statistics on the FP64 tensor pipe (d a multiple of 64): one CTA = (item, 64 x 64 tile pair
// ti <= tj), the item's rows streamed 16 at a time through the DMMA tile of stats.cuh (exact products of
// y = x - first row, fp64 accumulation in a fixed order), covariance written to both triangles from the same
// accumulators.  5000 x [750, 128]: the CUDA-core kernel above was ~half of the whole per-song pass.
// grid (n_pairs, items), 256 threads.
__global__ void __launch_bounds__(256, 2)
song_stats_dmma_kernel(const __half* __restrict__ emb, const long long* __restrict__ offsets, int d,
                       double* __restrict__ mu, double* __restrict__ cov, int* __restrict__ ok)
{
    constexpr int T = 64, R = 16, P = T + 4;
    __shared__ __align__(16) double Ys[2][2][R][P];              // [panel i | j][buffer][row][col]
    __shared__ double s_sum[2][T];
    const int z = blockIdx.y;
    const long long r0 = offsets[z], r1 = offsets[z + 1];
    const long long n = r1 - r0;
    const int n_tiles = d / T;
    int ti = 0, rem = blockIdx.x;
    while (rem >= n_tiles - ti) { rem -= n_tiles - ti; ++ti; }
    const int tj = ti + rem;
    const bool diag = ti == tj;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
    const int fr = lane >> 2, fk = lane & 3;
    const int lrow = t >> 4, lcol = (t & 15) * 4;
    double* C = cov + (size_t)z * d * d;
    if (n < 2) {                                                 // the reference asserts (fad.py:46): identity keeps the chain finite
        for (int e = t; e < T * T; e += 256) {
            const int gi = ti * T + e / T, gj = tj * T + e % T;
            C[(size_t)gi * d + gj] = gi == gj ? 1.0 : 0.0;
            if (!diag) C[(size_t)gj * d + gi] = 0.0;
        }
        if (diag && t < T) mu[(size_t)z * d + ti * T + t] = 0.0;
        if (blockIdx.x == 0 && t == 0) ok[z] = 0;
        return;
    }
    const __half* base = emb + (size_t)r0 * d;
    double si[4], sj[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        si[e] = (double)__half2float(base[ti * T + lcol + e]);
        sj[e] = (double)__half2float(base[tj * T + lcol + e]);
    }
    const __half* src_i = base + ti * T + lcol;
    const __half* src_j = base + tj * T + lcol;
    uint2 ri = make_uint2(0, 0), rj = make_uint2(0, 0);
    bool rok = false;
    auto fetch = [&](long long rr) {
        const long long r = rr + lrow;
        rok = r < n;
        if (rok) {
            ri = __ldg(reinterpret_cast<const uint2*>(src_i + (size_t)r * d));
            if (!diag) rj = __ldg(reinterpret_cast<const uint2*>(src_j + (size_t)r * d));
        }
    };
    double cs_i[4] = {0.0, 0.0, 0.0, 0.0}, cs_j[4] = {0.0, 0.0, 0.0, 0.0};
    auto unpack = [](uint2 v, const double (&s)[4], bool okr, double (&y)[4]) {
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        y[0] = okr ? (double)f0.x - s[0] : 0.0;  y[1] = okr ? (double)f0.y - s[1] : 0.0;
        y[2] = okr ? (double)f1.x - s[2] : 0.0;  y[3] = okr ? (double)f1.y - s[3] : 0.0;
    };
    auto stage = [&](int buf) {
        double y[4];
        unpack(ri, si, rok, y);
        *reinterpret_cast<double2*>(&Ys[0][buf][lrow][lcol]) = make_double2(y[0], y[1]);
        *reinterpret_cast<double2*>(&Ys[0][buf][lrow][lcol + 2]) = make_double2(y[2], y[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) cs_i[e] += y[e];
        if (!diag) {
            unpack(rj, sj, rok, y);
            *reinterpret_cast<double2*>(&Ys[1][buf][lrow][lcol]) = make_double2(y[0], y[1]);
            *reinterpret_cast<double2*>(&Ys[1][buf][lrow][lcol + 2]) = make_double2(y[2], y[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) cs_j[e] += y[e];
        }
    };
    double c[4][2][2] = {};
    const int stages = (int)((n + R - 1) / R);
    fetch(0);
    stage(0);
    __syncthreads();
    const int bp = diag ? 0 : 1;
    for (int st = 0; st < stages; ++st) {
        const int buf = st & 1;
        if (st + 1 < stages) fetch((long long)(st + 1) * R);
#pragma unroll
        for (int kk = 0; kk < R; kk += 4) {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ys[0][buf][kk + fk][wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Ys[bp][buf][kk + fk][wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
        if (st + 1 < stages) stage(buf ^ 1);
        __syncthreads();
    }
    // column sums of y for both panels: 16 loader rows per column group, summed in a fixed order
    double* red = &Ys[0][0][0][0];                               // 2 x 16 x 64 doubles fit the first panel's two buffers
    static_assert(2 * R * T <= 2 * R * P, "reduction scratch");
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        red[lrow * T + lcol + e] = cs_i[e];
        red[R * T + lrow * T + lcol + e] = diag ? cs_i[e] : cs_j[e];
    }
    __syncthreads();
    if (t < 2 * T) {
        const int which = t >> 6, col = t & 63;
        double v = 0.0;
        for (int k = 0; k < R; ++k) v += red[which * R * T + k * T + col];
        s_sum[which][col] = v;
    }
    __syncthreads();
    const double inv_n = 1.0 / (double)n, inv_n1 = 1.0 / (double)(n - 1);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int li = wm + i * 8 + fr, lj = wn + j * 8 + 2 * fk + e;
                const double v = (c[i][j][e] - s_sum[0][li] * s_sum[1][lj] * inv_n) * inv_n1;
                const int gi = ti * T + li, gj = tj * T + lj;
                C[(size_t)gi * d + gj] = v;
                if (!diag) C[(size_t)gj * d + gi] = v;
            }
    if (diag && t < T) {
        const double m = (double)__half2float(base[ti * T + t]) + s_sum[0][t] * inv_n;
        mu[(size_t)z * d + ti * T + t] = (double)__half2float(__double2half(m));      // fp16 mean (fad.py:48)
    }
    if (blockIdx.x == 0 && t == 0) ok[z] = 1;
}

// C_z = alpha A_z B_z + beta_diag I for two strided families of problems in one launch:
// blockIdx.z = item + family * items.  A stride of 0 shares the operand between items (the cached
// baseline root).  flags: per-item float[3] rotating max|W - I| slots as in dgemm_kernel.
struct DgemmFamily { const double* A; const double* B; double* C; long long sA, sB, sC; double alpha, beta_diag; };
struct DgemmStrided {
    DgemmFamily f[2];
    int items;
    float* flags;          // [items][3] or null
    int in_slot, out_slot, clear_slot;     // -1 = unused
    float tol;
};

__global__ void __launch_bounds__(256, 2)
dgemm_strided_kernel(const DgemmStrided p, int d)
{
    const int fam = blockIdx.z >= p.items ? 1 : 0;
    const int z = blockIdx.z - fam * p.items;
    float* fl = p.flags ? p.flags + (size_t)z * 3 : nullptr;
    if (fl && p.in_slot >= 0 && fl[p.in_slot] < p.tol) return;             // this item has converged
    if (fl && p.clear_slot >= 0 && fam == 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) fl[p.clear_slot] = 0.0f;
    const DgemmFamily f = p.f[fam];
    float dev;
    double tr;
    dgemm_tile(f.A + (size_t)z * f.sA, f.B + (size_t)z * f.sB, f.C + (size_t)z * f.sC, d, f.alpha, f.beta_diag, dev, tr);
    if (fl && p.out_slot >= 0) {
        for (int o = 16; o > 0; o >>= 1) dev = fmaxf(dev, __shfl_xor_sync(0xffffffffu, dev, o));
        if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(fl + p.out_slot), __float_as_uint(dev));
    }
}

// scal[z] = {|A_z|_F, tr A_z}; one block per item
__global__ void __launch_bounds__(256)
norm_trace_batched_kernel(const double* __restrict__ A, int d, double* __restrict__ scal /*[items][2]*/)
{
    __shared__ double r1[256], r2[256];
    const double* Az = A + (size_t)blockIdx.x * d * d;
    double s = 0.0, t = 0.0;
    for (size_t e = threadIdx.x; e < (size_t)d * d; e += 256) { const double v = Az[e]; s += v * v; }
    for (int i = threadIdx.x; i < d; i += 256) t += Az[(size_t)i * d + i];
    r1[threadIdx.x] = s; r2[threadIdx.x] = t;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) { r1[threadIdx.x] += r1[threadIdx.x + k]; r2[threadIdx.x] += r2[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { scal[2 * blockIdx.x] = sqrt(r1[0]); scal[2 * blockIdx.x + 1] = r2[0]; }
}

// Y_z = sym(A_z)/|A_z|_F + delta I, Z_z = I, flags_z = {0, 0, 1e30}; grid (blocks, items)
__global__ void __launch_bounds__(256)
ns_init_batched_kernel(const double* __restrict__ A, int d, const double* __restrict__ scal,
                       double* __restrict__ Y, double* __restrict__ Z, float* __restrict__ flags)
{
    const int z = blockIdx.y;
    const double nrm = scal[2 * z];
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    const double* Az = A + (size_t)z * d * d;
    double* Yz = Y + (size_t)z * d * d;
    double* Zz = Z + (size_t)z * d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)d * d; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        Yz[e] = 0.5 * (Az[e] + Az[(size_t)j * d + i]) * inv + ((i == j) ? kNsDelta : 0.0);
        Zz[e] = (i == j) ? 1.0 : 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { flags[3 * z] = 0.0f; flags[3 * z + 1] = 0.0f; flags[3 * z + 2] = 1.0e30f; }
}

// out[z][0..7] in the layout of frechet_assemble_kernel; one block per item
__global__ void __launch_bounds__(256)
frechet_assemble_batched_kernel(const double* __restrict__ mu1, const double* __restrict__ mu2 /*[items][d]*/, int d,
                                const double* __restrict__ scal1 /*|C1|_F, tr C1*/,
                                const double* __restrict__ scalC /*[items][2]: |C_z|_F, tr C_z*/,
                                const double* __restrict__ scalM /*[items][2]*/,
                                const double* __restrict__ Y, const double* __restrict__ Z,
                                const int* __restrict__ ok, const long long* __restrict__ offsets,
                                int iters, double* __restrict__ out)
{
    __shared__ double r0[256], r1[256], r2[256];
    const int z = blockIdx.x;
    const double* Yz = Y + (size_t)z * d * d;
    const double* Zz = Z + (size_t)z * d * d;
    double s = 0.0, ty = 0.0, tz = 0.0;
    for (int i = threadIdx.x; i < d; i += 256) {
        const double df = mu1[i] - mu2[(size_t)z * d + i];
        s += df * df;
        ty += Yz[(size_t)i * d + i];
        tz += Zz[(size_t)i * d + i];
    }
    r0[threadIdx.x] = s; r1[threadIdx.x] = ty; r2[threadIdx.x] = tz;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) { r0[threadIdx.x] += r0[threadIdx.x + k]; r1[threadIdx.x] += r1[threadIdx.x + k]; r2[threadIdx.x] += r2[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double* o = out + (size_t)z * 8;
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        const double tr_sqrt = sqrt(scalM[2 * z]) * (r1[0] - kNsDelta * r2[0]);
        const bool good = ok[z] != 0;
        o[0] = good ? r0[0] + scal1[1] + scalC[2 * z + 1] - 2.0 * tr_sqrt : nan;
        o[1] = good ? tr_sqrt : nan;
        o[2] = 0.0;
        o[3] = (double)iters;
        o[4] = r0[0];
        o[5] = scal1[1];
        o[6] = scalC[2 * z + 1];
        o[7] = (double)(offsets[z + 1] - offsets[z]);
    }
}

}  // namespace fad
