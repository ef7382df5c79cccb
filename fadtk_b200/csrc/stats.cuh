// Sufficient statistics of an embedding matrix E[N, d] (fp16):   n,  sum(x - s),  sum(y y^T)   with y = x - s and
// s a shared fp16 shift vector.  Replaces np.mean / np.cov in fadtk/fad.py:42-48 and the per-file scatter + Chan
// merge in fadtk/utils.py:13-46 with one shifted E^T E contraction.
//
// Three kernels compute the same packed accumulator (fad_stats_accumulate's `mode`):
//
// stats_dmma_kernel<In>  (mode 0, the product default)   fp64 tensor pipe.  x and s are fp16, so y = x - s is exact
//   in fp64; products and sums are fp64 (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4): the result is the Gram matrix of
//   the data to ~1e-16, hence positive semi-definite.  Parity needs that: a covariance with cond ~1e9 (CLAP/MERT) or
//   a rank-deficient per-song covariance perturbed at the 1e-6 level of an fp32-accumulating path is indefinite -
//   Newton-Schulz diverges on it and tr sqrt(C1 C2) moves by percents.  One CTA = (64x64 output tile ti <= tj, row
//   range); Y tiles are staged in shared memory as doubles with pitch 68 (== 4 mod 16: conflict-free fragment loads)
//   and double-buffered; each job's tile goes to a workspace and stats_dmma_reduce_kernel sums the jobs in a fixed
//   order (deterministic).  In = double with no shift is the variant score() feeds with per-file fp16-rounded means
//   (fad_stats_accumulate_f64).  ncu: profiles/r2_ncu_fp64_dmma.md.
//
// stats_umma_kernel  (mode 1, opt-in)   tcgen05 fp16 hi/lo: yh = fp16(y), yl = fp16(y - yh),
//       sum y y^T ~= sum yh yh^T + yh yl^T + yl yh^T            (yl yl^T ~ 2^-22 is dropped)
//   Every fp16 x fp16 product is exact in the fp32 accumulator; the accumulation itself is cut every 256 rows and
//   drained into an fp64 tile, because tensor-core fp32 adds truncate.  One CTA = (128x128 output tile, row range).
//   E is row-major, so both operands of E^T E are "MN-major": a TMA box [32 rows x 64 cols] with 128-B swizzle IS
//   the canonical MN-major SWIZZLE_128B UMMA layout (K = row index).
//     warp 0       TMA producer: per 32-row stage, two 64-column boxes per panel
//     warps 4-11   transform: in smem, x -> (yh in place, yl into a second panel), zero rows past the end, exact
//                  column sums of x - s and of yh + yl in fp64 registers
//     warp 1       MMA issuer: per stage 2 k-steps x {hh, hl, lh} tcgen05.mma, fp32 in TMEM
//     warps 12-15  drain: every 256 rows the TMEM tile is added into an fp64 tile in shared memory
//   Good to ~1e-6 relative: fine for full-rank, well-conditioned sets only.
//
// stats_simt_kernel  (mode 2)   fp64 CUDA-core contraction of the exact y: the round-1 default, kept as the
//   independent cross-check of mode 0 (tests/test_gpu_kernels.py compares all three).
//
// Packed accumulator (fp64, caller-owned, all-reduced across GPUs as-is):
//   acc[0] = n,  acc[1 .. d] = sum(x - s) (exact),  acc[1+d .. 1+d+d*d) = sum(y y^T)
//   (d x d, full, row-major),  acc[1+d+d*d ..] = sum(yh + yl)  (centring term of the covariance; = sum y in modes 0, 2)
#pragma once
#include "sm100.cuh"

namespace fad {

constexpr int kStTile = 128;
constexpr int kStStageRows = 32;
constexpr int kStStagesPerChunk = 8;               // 256 rows per fp32 TMEM accumulation
constexpr int kStStages = 3;
constexpr uint32_t kStBlockBytes = kStStageRows * 128;            // one 64-col box: 4 KiB
constexpr uint32_t kStPanelBytes = 2 * kStBlockBytes;             // 128 cols x 32 rows: 8 KiB
constexpr uint32_t kStStageBytes = 4 * kStPanelBytes;             // Ah | Bh | Al | Bl = 32 KiB
constexpr int kStThreads = 512;
constexpr int kStTransformThreads = 256;
constexpr uint32_t kStSmemBytes = kStStages * kStStageBytes + kStTile * kStTile * 8 + 1024 + 256;

struct StatsJobParams {
    long long n_rows;          // valid rows in E
    int d;
    int n_tiles;               // d / 128
    int n_pairs;               // n_tiles (n_tiles + 1) / 2
    int n_splits;              // row splits per tile pair
    long long rows_per_split;  // multiple of 32
    const __half* shift;       // [d]
    double* ws_tiles;          // [n_pairs * n_splits][128 (col)][128 (row)]
    double* ws_sums;           // [n_tiles * n_splits][2][128]  (exact x-s sums | yh+yl sums)
};

__device__ __forceinline__ void pair_to_tiles(int pair, int n_tiles, int& ti, int& tj) {
    ti = 0;
    int rem = pair;
    while (rem >= n_tiles - ti) { rem -= n_tiles - ti; ++ti; }
    tj = ti + rem;
}

// y = x - s (exact in fp32) -> hi/lo fp16 pair
__device__ __forceinline__ void split_hi_lo(__half2 x, __half2 s, __half2& hi, __half2& lo, float2& y) {
    const float2 fx = __half22float2(x), fs = __half22float2(s);
    y = make_float2(fx.x - fs.x, fx.y - fs.y);
    hi = __floats2half2_rn(y.x, y.y);
    const float2 fh = __half22float2(hi);
    lo = __floats2half2_rn(y.x - fh.x, y.y - fh.y);
}

__global__ void __launch_bounds__(kStThreads, 1)
stats_umma_kernel(const __grid_constant__ CUtensorMap map_e, const StatsJobParams p)
{
    using namespace sm100;
    constexpr uint32_t kIdesc = make_idesc(FMT_F16, kStTile, kStTile, /*a MN-major*/1, /*b MN-major*/1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    double* acc64 = reinterpret_cast<double*>(smem + kStStages * kStStageBytes);     // [col][row]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStStages * kStStageBytes + kStTile * kStTile * 8);
    uint64_t* full = bars;                       // TMA landed
    uint64_t* ready = bars + kStStages;          // transform done
    uint64_t* empty = bars + 2 * kStStages;      // MMAs retired
    uint64_t* tmem_full = bars + 3 * kStStages;
    uint64_t* tmem_empty = bars + 3 * kStStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.x;
    const int pair = job / p.n_splits, split = job % p.n_splits;
    int ti, tj;
    pair_to_tiles(pair, p.n_tiles, ti, tj);
    const bool diag = (ti == tj);
    const long long row_begin = (long long)split * p.rows_per_split;
    long long row_end = row_begin + p.rows_per_split;
    if (row_end > p.n_rows) row_end = p.n_rows;
    const long long span = row_end > row_begin ? row_end - row_begin : 0;
    const int n_stages_total = (int)((span + kStStageRows - 1) / kStStageRows);
    const int n_chunks = (n_stages_total + kStStagesPerChunk - 1) / kStStagesPerChunk;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&map_e);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStStages; ++s) {
            mbar_init(&full[s], 1); mbar_init(&ready[s], kStTransformThreads / 32); mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<256>(tmem_slot);
    for (int i = threadIdx.x; i < kStTile * kStTile; i += kStThreads) acc64[i] = 0.0;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < n_stages_total; ++it) {
                const int r0 = (int)(row_begin + (long long)it * kStStageRows);
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], diag ? kStPanelBytes : 2 * kStPanelBytes);
                uint8_t* st = smem + s * kStStageBytes;
                tma_load_2d(st, &map_e, &full[s], ti * kStTile, r0);
                tma_load_2d(st + kStBlockBytes, &map_e, &full[s], ti * kStTile + 64, r0);
                if (!diag) {
                    tma_load_2d(st + kStPanelBytes, &map_e, &full[s], tj * kStTile, r0);
                    tma_load_2d(st + kStPanelBytes + kStBlockBytes, &map_e, &full[s], tj * kStTile + 64, r0);
                }
                if (++s == kStStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t acc_ph = 0;
            int it = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * kStTile;
                const int n_st = min(kStStagesPerChunk, n_stages_total - it);
                for (int q = 0; q < n_st; ++q, ++it) {
                    mbar_wait(&ready[s], ph);
                    tc_fence_after_sync();
                    const uint32_t base = smem_u32(smem + s * kStStageBytes);
                    const uint32_t ah = base, bh = diag ? base : base + kStPanelBytes;
                    const uint32_t al = base + 2 * kStPanelBytes, bl = diag ? al : al + kStPanelBytes;
                    // MN-major SW128: 64-col blocks kStBlockBytes apart (LBO), 8-row K groups 1024 B apart (SBO)
                    const uint64_t d_ah = mnmajor_sw128_desc(ah, kStBlockBytes, 1024);
                    const uint64_t d_bh = mnmajor_sw128_desc(bh, kStBlockBytes, 1024);
                    const uint64_t d_al = mnmajor_sw128_desc(al, kStBlockBytes, 1024);
                    const uint64_t d_bl = mnmajor_sw128_desc(bl, kStBlockBytes, 1024);
#pragma unroll
                    for (int k = 0; k < kStStageRows / 16; ++k) {
                        const uint32_t off = 128 * k;          // 16 K-rows = 2048 B = 128 x 16 B
                        umma_f16(d_tmem, d_ah + off, d_bh + off, kIdesc, (q | k) != 0);
                        umma_f16(d_tmem, d_ah + off, d_bl + off, kIdesc, 1);
                        umma_f16(d_tmem, d_al + off, d_bh + off, kIdesc, 1);
                    }
                    umma_commit(&empty[s]);
                    if (++s == kStStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ------------------------------------------------------- shift + hi/lo split
        const int t = threadIdx.x - 128;              // 0..255
        const int cg = t & 15;                        // 16-B column group inside the 128-col panel
        const int rl = (t >> 4) & 7;                  // row lane inside an 8-row swizzle group
        const int hf = t >> 7;                        // which half of the 32 rows
        const int cb = cg >> 3, lc = cg & 7;
        const uint32_t chunk_off = cb * kStBlockBytes + ((lc ^ rl) << 4);   // swizzled (row%8 == rl, lc)
        __half2 shA[4], shB[4];
        {
            const uint4 a = *reinterpret_cast<const uint4*>(p.shift + ti * kStTile + cg * 8);
            const uint4 b = *reinterpret_cast<const uint4*>(p.shift + tj * kStTile + cg * 8);
            shA[0] = *reinterpret_cast<const __half2*>(&a.x); shA[1] = *reinterpret_cast<const __half2*>(&a.y);
            shA[2] = *reinterpret_cast<const __half2*>(&a.z); shA[3] = *reinterpret_cast<const __half2*>(&a.w);
            shB[0] = *reinterpret_cast<const __half2*>(&b.x); shB[1] = *reinterpret_cast<const __half2*>(&b.y);
            shB[2] = *reinterpret_cast<const __half2*>(&b.z); shB[3] = *reinterpret_cast<const __half2*>(&b.w);
        }
        double sum_x[8], sum_y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sum_x[j] = 0.0; sum_y[j] = 0.0; }
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < n_stages_total; ++it) {
            const long long r0 = row_begin + (long long)it * kStStageRows;
            mbar_wait(&full[s], ph);
            uint8_t* st = smem + s * kStStageBytes;
#pragma unroll
            for (int panel = 0; panel < 2; ++panel) {
                if (panel == 1 && diag) break;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = rl + 8 * (hf * 2 + i);
                    uint4* ph_ptr = reinterpret_cast<uint4*>(st + panel * kStPanelBytes + chunk_off + r * 128);
                    uint4* pl_ptr = reinterpret_cast<uint4*>(st + (2 + panel) * kStPanelBytes + chunk_off + r * 128);
                    uint4 v = *ph_ptr, w = make_uint4(0, 0, 0, 0);
                    if (r0 + r < row_end) {
                        __half2* hx = reinterpret_cast<__half2*>(&v);
                        __half2* hl = reinterpret_cast<__half2*>(&w);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __half2 hi, lo; float2 y;
                            split_hi_lo(hx[j], panel ? shB[j] : shA[j], hi, lo, y);
                            hx[j] = hi; hl[j] = lo;
                            if (panel == 0 && diag) {
                                const float2 fh = __half22float2(hi), fl = __half22float2(lo);
                                sum_x[2 * j] += (double)y.x;            sum_x[2 * j + 1] += (double)y.y;
                                sum_y[2 * j] += (double)fh.x + (double)fl.x;
                                sum_y[2 * j + 1] += (double)fh.y + (double)fl.y;
                            }
                        }
                    } else {
                        v = make_uint4(0, 0, 0, 0);
                    }
                    *ph_ptr = v;
                    *pl_ptr = w;
                }
            }
            fence_proxy_async_smem();                 // generic-proxy stores -> visible to UMMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
            if (++s == kStStages) { s = 0; ph ^= 1; }
        }
        // column sums: reduce the 16 row lanes through stage 0 once every MMA has retired
        if (diag) {
            if (n_chunks > 0) {
                const int last = n_chunks - 1;
                mbar_wait(&tmem_full[last & 1], (uint32_t)((last >> 1) & 1));
            }
            asm volatile("bar.sync 1, 256;");
            double* red = reinterpret_cast<double*>(smem);       // [2][16 row lanes][128 cols] = 32 KiB
            const int lane16 = hf * 8 + rl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                red[lane16 * 128 + cg * 8 + j] = sum_x[j];
                red[2048 + lane16 * 128 + cg * 8 + j] = sum_y[j];
            }
            asm volatile("bar.sync 1, 256;");
            if (t < 128) {
                double tx = 0.0, ty = 0.0;
                for (int k = 0; k < 16; ++k) { tx += red[k * 128 + t]; ty += red[2048 + k * 128 + t]; }
                double* wsum = p.ws_sums + ((size_t)ti * p.n_splits + split) * 2 * kStTile;
                wsum[t] = tx;
                wsum[kStTile + t] = ty;
            }
        }
    } else if (warp >= 12) {
        // -------------------------------------------------- drain TMEM -> fp64 smem
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int acc = 0; uint32_t acc_ph = 0;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after_sync();
            const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * kStTile;
#pragma unroll 1
            for (int cc = 0; cc < kStTile / 32; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + cc * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    acc64[(cc * 32 + j) * kStTile + row] += (double)__uint_as_float(v[j]);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
        double* dst = p.ws_tiles + (size_t)job * kStTile * kStTile;
        for (int col = 0; col < kStTile; ++col) dst[col * kStTile + row] = acc64[col * kStTile + row];
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<256>(tmem_base);
}

// acc += sum over row splits of the job tiles (fixed order => deterministic).
// grid = (n_pairs), block = 256
__global__ void stats_reduce_kernel(StatsJobParams p, double* __restrict__ acc)
{
    const int pair = blockIdx.x;
    int ti, tj;
    pair_to_tiles(pair, p.n_tiles, ti, tj);
    const int d = p.d;
    double* outer = acc + 1 + d;
    for (int e = threadIdx.x; e < kStTile * kStTile; e += blockDim.x) {
        const int col = e / kStTile, row = e % kStTile;        // workspace layout is [col][row]
        double v = 0.0;
        for (int s = 0; s < p.n_splits; ++s)
            v += p.ws_tiles[((size_t)pair * p.n_splits + s) * kStTile * kStTile + e];
        const int I = ti * kStTile + row, J = tj * kStTile + col;
        outer[(size_t)I * d + J] += v;
        if (ti != tj) outer[(size_t)J * d + I] += v;
    }
    if (ti == tj) {
        for (int c = threadIdx.x; c < kStTile; c += blockDim.x) {
            double vx = 0.0, vy = 0.0;
            for (int s = 0; s < p.n_splits; ++s) {
                const double* wsum = p.ws_sums + ((size_t)ti * p.n_splits + s) * 2 * kStTile;
                vx += wsum[c]; vy += wsum[kStTile + c];
            }
            acc[1 + ti * kStTile + c] += vx;
            acc[1 + (size_t)d + (size_t)d * d + ti * kStTile + c] += vy;
        }
    }
    if (pair == 0 && threadIdx.x == 0) acc[0] += (double)p.n_rows;
}

// --------------------------------------------------------------------------------------
// stats_dmma_kernel: the product default.  EXACT Gram matrix on the FP64 tensor pipe.
//   y = x - s is exact in fp64 for any two fp16 values (<= 22 significant bits), every product
//   y_a y_b is exact (<= 44 bits), and mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) accumulates in fp64 in
//   a fixed order: the result is the Gram matrix of the data to ~1e-16 - positive semi-definite, which
//   is what Newton-Schulz on cond-1e9 / rank-deficient covariances needs (see the header) - at the
//   tensor-pipe rate instead of the CUDA-core DFMA rate, with no atomics (bit-reproducible).
// One CTA = one job = (64 x 64 output tile (ti <= tj), row range).  256 threads = 8 warps as 2 x 4,
// a warp owns 32 x 16 outputs (4 x 2 DMMA blocks).  E is row-major [row][col]: for E^T E both operand
// fragments read smem as Y[k = row][m or n = col] with pitch 68 doubles (= 4 mod 16: the m8n8k4
// fragment pattern k = lane%4, m = lane/4 touches 16 distinct 8-byte banks per half-warp).
// Loads: 8 bytes (4 fp16) per thread and panel, coalesced 128-B row segments, converted and shifted on
// the way into shared memory; the next 16-row stage is in flight while the current one is multiplied.
// Column sums of y come from the loader's own registers (no extra smem reads).
// Jobs write their fp64 tile to a workspace; stats_dmma_reduce_kernel sums row splits in a fixed order.
constexpr int kSdTile = 64, kSdRows = 16, kSdPitch = kSdTile + 4;

struct StatsDmmaParams {
    long long n_rows;
    int d, n_tiles, n_pairs, n_splits;
    long long rows_per_split;      // multiple of kSdRows
    const __half* shift;           // [d]
    double* ws_tiles;              // [n_pairs * n_splits][64 (row of the tile)][64 (col)]
    double* ws_sums;               // [n_tiles * n_splits][64]
};

__device__ __forceinline__ void stats_dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// In = __half (embeddings; shift vector applied) or double (per-file mean rows of the reference's online merge, no shift)
template <typename In>
__global__ void __launch_bounds__(256, 2)
stats_dmma_kernel(const In* __restrict__ E, const StatsDmmaParams p)
{
    constexpr bool kHalf = sizeof(In) == 2;
    __shared__ __align__(16) double Ys[2][2][kSdRows][kSdPitch];     // [panel i | j][buffer][row][col]
    const int job = blockIdx.x;
    const int pair = job / p.n_splits, split = job % p.n_splits;
    int ti, tj;
    pair_to_tiles(pair, p.n_tiles, ti, tj);
    const bool diag = ti == tj;
    const long long row_begin = (long long)split * p.rows_per_split;
    const long long row_end = min(p.n_rows, row_begin + p.rows_per_split);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
    const int fr = lane >> 2, fk = lane & 3;
    const int lrow = t >> 4, lcol = (t & 15) * 4;               // loader: row of the stage, first of 4 columns

    double si[4] = {0.0, 0.0, 0.0, 0.0}, sj[4] = {0.0, 0.0, 0.0, 0.0};
    if (kHalf && p.shift != nullptr) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            si[e] = (double)__half2float(p.shift[ti * kSdTile + lcol + e]);
            sj[e] = (double)__half2float(p.shift[tj * kSdTile + lcol + e]);
        }
    }
    const In* src_i = E + (size_t)ti * kSdTile + lcol;
    const In* src_j = E + (size_t)tj * kSdTile + lcol;
    double ri[4] = {0.0, 0.0, 0.0, 0.0}, rj[4] = {0.0, 0.0, 0.0, 0.0};        // the 4 values of this thread, already as fp64
    bool rok = false;
    auto load4 = [&](const In* ptr, double (&v)[4]) {
        if constexpr (kHalf) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(ptr));
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            v[0] = (double)f0.x; v[1] = (double)f0.y; v[2] = (double)f1.x; v[3] = (double)f1.y;
        } else {
            const double2 a = __ldg(reinterpret_cast<const double2*>(ptr)), b = __ldg(reinterpret_cast<const double2*>(ptr) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        }
    };
    auto fetch = [&](long long r0) {
        const long long r = r0 + lrow;
        rok = r < row_end;
        if (rok) {
            load4(src_i + (size_t)r * p.d, ri);
            if (!diag) load4(src_j + (size_t)r * p.d, rj);
        }
    };
    double colsum[4] = {0.0, 0.0, 0.0, 0.0};
    auto unpack = [](const double (&v)[4], const double (&s)[4], bool ok, double (&y)[4]) {
#pragma unroll
        for (int e = 0; e < 4; ++e) y[e] = ok ? v[e] - s[e] : 0.0;
    };
    auto stage = [&](int buf) {
        double y[4];
        unpack(ri, si, rok, y);
        *reinterpret_cast<double2*>(&Ys[0][buf][lrow][lcol]) = make_double2(y[0], y[1]);
        *reinterpret_cast<double2*>(&Ys[0][buf][lrow][lcol + 2]) = make_double2(y[2], y[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) colsum[e] += y[e];
        if (!diag) {
            unpack(rj, sj, rok, y);
            *reinterpret_cast<double2*>(&Ys[1][buf][lrow][lcol]) = make_double2(y[0], y[1]);
            *reinterpret_cast<double2*>(&Ys[1][buf][lrow][lcol + 2]) = make_double2(y[2], y[3]);
        }
    };

    double c[4][2][2] = {};
    const int stages = row_end > row_begin ? (int)((row_end - row_begin + kSdRows - 1) / kSdRows) : 0;
    if (stages > 0) {
        fetch(row_begin);
        stage(0);
    }
    __syncthreads();
    for (int st = 0; st < stages; ++st) {
        const int buf = st & 1;
        if (st + 1 < stages) fetch(row_begin + (long long)(st + 1) * kSdRows);
        const int bp = diag ? 0 : 1;
#pragma unroll
        for (int kk = 0; kk < kSdRows; kk += 4) {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ys[0][buf][kk + fk][wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Ys[bp][buf][kk + fk][wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) stats_dmma_884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
        if (st + 1 < stages) stage(buf ^ 1);
        __syncthreads();
    }
    double* dst = p.ws_tiles + (size_t)job * kSdTile * kSdTile;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
            *reinterpret_cast<double2*>(dst + (wm + i * 8 + fr) * kSdTile + wn + j * 8 + 2 * fk) =
                make_double2(c[i][j][0], c[i][j][1]);
    if (diag) {
        // column sums: 16 loader rows per column group -> one fixed-order sum per column
        double* red = &Ys[0][0][0][0];                          // 16 x 64 doubles, the stages are done with it
        red[lrow * kSdTile + lcol + 0] = colsum[0]; red[lrow * kSdTile + lcol + 1] = colsum[1];
        red[lrow * kSdTile + lcol + 2] = colsum[2]; red[lrow * kSdTile + lcol + 3] = colsum[3];
        __syncthreads();
        if (t < kSdTile) {
            double v = 0.0;
            for (int k = 0; k < 16; ++k) v += red[k * kSdTile + t];
            p.ws_sums[((size_t)ti * p.n_splits + split) * kSdTile + t] = v;
        }
    }
}

// acc += sum over row splits of the job tiles, fixed order.  grid = (n_pairs, 16): block (pair, y) owns 256 of the
// tile's 4096 entries (a 3-block grid at d = 128 took 0.33 ms for 12 MB of partial tiles - longer than the Gram itself)
__global__ void __launch_bounds__(256) stats_dmma_reduce_kernel(StatsDmmaParams p, double* __restrict__ acc)
{
    const int pair = blockIdx.x;
    int ti, tj;
    pair_to_tiles(pair, p.n_tiles, ti, tj);
    const int d = p.d;
    double* outer = acc + 1 + d;
    {
        const int e = blockIdx.y * 256 + threadIdx.x;
        const int row = e / kSdTile, col = e % kSdTile;
        const double* src = p.ws_tiles + (size_t)pair * p.n_splits * kSdTile * kSdTile + e;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int s = 0;
        for (; s + 4 <= p.n_splits; s += 4) {                   // four loads in flight; the order of the sum stays fixed
            const double a = src[(size_t)(s + 0) * kSdTile * kSdTile], b = src[(size_t)(s + 1) * kSdTile * kSdTile];
            const double c = src[(size_t)(s + 2) * kSdTile * kSdTile], e4 = src[(size_t)(s + 3) * kSdTile * kSdTile];
            v0 += a; v1 += b; v2 += c; v3 += e4;
        }
        for (; s < p.n_splits; ++s) v0 += src[(size_t)s * kSdTile * kSdTile];
        const double v = (v0 + v1) + (v2 + v3);
        const int I = ti * kSdTile + row, J = tj * kSdTile + col;
        outer[(size_t)I * d + J] += v;
        if (ti != tj) outer[(size_t)J * d + I] += v;
    }
    if (ti == tj && blockIdx.y == 0) {
        for (int cidx = threadIdx.x; cidx < kSdTile; cidx += blockDim.x) {
            double v = 0.0;
            for (int s = 0; s < p.n_splits; ++s) v += p.ws_sums[((size_t)ti * p.n_splits + s) * kSdTile + cidx];
            acc[1 + ti * kSdTile + cidx] += v;                                        // sum(x - s), exact
            acc[1 + (size_t)d + (size_t)d * d + ti * kSdTile + cidx] += v;            // centring term: same y
        }
    }
    if (pair == 0 && blockIdx.y == 0 && threadIdx.x == 0) acc[0] += (double)p.n_rows;
}

// --------------------------------------------------------------------------------------
// fp64 CUDA-core version (verification path only: FADTK_STATS=simt / mode 2).  grid = (row chunks, d/64, d/64) upper tiles only.
constexpr int kSimtRows = 1024;
__global__ void __launch_bounds__(256)
stats_simt_kernel(const __half* __restrict__ E, long long n_rows, int d,
                  const __half* __restrict__ shift, double* __restrict__ acc)
{
    const int ti = blockIdx.y, tj = blockIdx.z;
    if (tj < ti) return;
    __shared__ double yi[32][65], yj[32][65];
    const long long r_begin = (long long)blockIdx.x * kSimtRows;
    const long long r_end = min(n_rows, r_begin + kSimtRows);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 4x4 outputs per thread
    double c[4][4] = {};
    double csum = 0.0, csum_x = 0.0;
    for (long long r0 = r_begin; r0 < r_end; r0 += 32) {
        for (int i = threadIdx.x; i < 32 * 64; i += 256) {
            const int r = i >> 6, cc = i & 63;
            double a = 0.0, b = 0.0;
            if (r0 + r < r_end) {
                const __half* rowp = E + (size_t)(r0 + r) * d;
                // x - s in fp64 is exact for any two fp16 values; products of such differences carry
                // <= 2 x 40 bits, rounded once to fp64: relative error 1e-16 per term
                a = (double)__half2float(rowp[ti * 64 + cc]) - (double)__half2float(shift[ti * 64 + cc]);
                b = (double)__half2float(rowp[tj * 64 + cc]) - (double)__half2float(shift[tj * 64 + cc]);
            }
            yi[r][cc] = a; yj[r][cc] = b;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = yi[r][ty * 4 + u]; b[u] = yj[r][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) c[u][v] = fma(a[u], b[v], c[u][v]);
        }
        if (ti == tj && threadIdx.x < 64)
            for (int r = 0; r < 32; ++r) { csum += yi[r][threadIdx.x]; }
        csum_x = csum;                                      // y is exact here: both sums coincide
        __syncthreads();
    }
    double* outer = acc + 1 + d;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int I = ti * 64 + ty * 4 + u, J = tj * 64 + tx * 4 + v;
            atomicAdd(&outer[(size_t)I * d + J], c[u][v]);
            if (ti != tj) atomicAdd(&outer[(size_t)J * d + I], c[u][v]);
        }
    if (ti == tj && threadIdx.x < 64) {
        atomicAdd(&acc[1 + ti * 64 + threadIdx.x], csum_x);
        atomicAdd(&acc[1 + (size_t)d + (size_t)d * d + ti * 64 + threadIdx.x], csum);
    }
    if (blockIdx.x == 0 && ti == 0 && tj == 0 && threadIdx.x == 0) atomicAdd(&acc[0], (double)n_rows);
}

// Per-file means of equal-length files (file f = rows [f r, (f + 1) r) of emb): the exact mean in fp64 and the mean as
// the reference's _process_file returns it for an fp16 .npy (np.mean of an fp16 array: fp32 accumulation, result rounded
// to fp16 - fadtk/utils.py:14), both stored as fp64 rows for the Gram kernel.  grid-stride over files x d.
__global__ void file_means_kernel(const __half* __restrict__ emb, long long n_files, int r, int d,
                                  double* __restrict__ m64, double* __restrict__ m16)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_files * d; e += (long long)gridDim.x * blockDim.x) {
        const long long f = e / d;
        const int c = (int)(e - f * d);
        const __half* src = emb + ((size_t)f * r) * d + c;
        double s64 = 0.0;
        float s32 = 0.f;
        for (int k = 0; k < r; ++k) { const float v = __half2float(src[(size_t)k * d]); s64 += (double)v; s32 += v; }
        m64[e] = s64 / (double)r;
        m16[e] = (double)__half2float(__float2half_rn(s32 / (float)r));
    }
}

// mu, cov as the reference's online merge computes them for n_files files of r rows each (fadtk/utils.py:13-46):
//   mu_ref = sum_f r m16_f / n,   S_ref = S_exact - sum_f r (m64_f - mu)(m64_f - mu)^T + sum_f r (m16_f - mu_ref)(m16_f - mu_ref)^T
// from the three packed accumulators (rows; exact file means; fp16-rounded file means - the latter two unshifted).
// r == 1 reproduces the reference's all-NaN covariance (np.cov of a single row) unless keep_single.
__global__ void stats_finalize_mirrored_kernel(const double* __restrict__ acc, const double* __restrict__ acc64,
                                               const double* __restrict__ acc16, const __half* __restrict__ shift,
                                               int r, int d, int keep_single, double* __restrict__ mu_out, double* __restrict__ cov_out)
{
    const double n = acc[0];
    const double* sum_x = acc + 1;
    const double* outer = acc + 1 + d;
    const double* sum_y = acc + 1 + (size_t)d + (size_t)d * d;
    const double* s64f = acc64 + 1;            // sum_f m64_f
    const double* o64 = acc64 + 1 + d;         // sum_f m64_f m64_f^T
    const double* s16f = acc16 + 1;
    const double* o16 = acc16 + 1 + d;
    const double w = (double)r;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)d * d; e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        const double mu_i = (double)__half2float(shift[i]) + (n > 0.0 ? sum_x[i] / n : 0.0);
        const double mu_j = (double)__half2float(shift[j]) + (n > 0.0 ? sum_x[j] / n : 0.0);
        const double mr_i = n > 0.0 ? w * s16f[i] / n : 0.0, mr_j = n > 0.0 ? w * s16f[j] / n : 0.0;
        double c = 0.0;
        if (n >= 2.0) {
            const double s_exact = outer[e] - sum_y[i] * sum_y[j] / n;                               // (n - 1) cov_exact
            const double b64 = w * o64[e] - mu_i * (w * s64f[j]) - (w * s64f[i]) * mu_j + n * mu_i * mu_j;
            const double b16 = w * o16[e] - mr_i * (w * s16f[j]) - (w * s16f[i]) * mr_j + n * mr_i * mr_j;
            c = (s_exact - b64 + b16) / (n - 1.0);
            if (r == 1 && !keep_single) c = nan;
        }
        cov_out[e] = c;
        if (j == 0) mu_out[i] = mr_i;
    }
}

// mu = shift + sum(x-s)/n ; cov = (outer - sum(y) sum(y)^T / n) / (n - 1)   (cov = 0 when n < 2,
// fadtk/utils.py:42-43).  grid-stride over d*d.
__global__ void stats_finalize_kernel(const double* __restrict__ acc, const __half* __restrict__ shift,
                                      int d, double* __restrict__ mu, double* __restrict__ cov)
{
    const double n = acc[0];
    const double* sum_x = acc + 1;
    const double* outer = acc + 1 + d;
    const double* sum = acc + 1 + (size_t)d + (size_t)d * d;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)d * d;
         e += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(e / d), j = (int)(e % d);
        cov[e] = n < 2.0 ? 0.0 : (outer[e] - sum[i] * sum[j] / n) / (n - 1.0);
        if (j == 0) mu[i] = (double)__half2float(shift[i]) + (n > 0.0 ? sum_x[i] / n : 0.0);
    }
}

// out[i, :] = src[idx[i], :]  (FAD-inf bootstrap gather, fadtk/fad.py:333-334); 16-B vectors.
__global__ void gather_rows_kernel(const __half* __restrict__ src, const long long* __restrict__ idx,
                                   long long n_out, int d, __half* __restrict__ out)
{
    const int vec_per_row = d / 8;
    const long long total = n_out * vec_per_row;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / vec_per_row;
        const int v = (int)(e % vec_per_row);
        reinterpret_cast<uint4*>(out)[e] = reinterpret_cast<const uint4*>(src + (size_t)idx[i] * d)[v];
    }
}

}  // namespace fad
