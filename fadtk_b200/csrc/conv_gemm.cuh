// Implicit-GEMM 3x3 convolution / fully-connected layer on tcgen05 tensor cores.
//
//   D[128 pixels, N_TILE channels] = sum over (tap, 64-channel block) of
//         X_tap[128 pixels, 64 ch] (fp16, K-major, TMA im2col-by-coordinates)
//       x W[N_TILE, (tap, 64 ch)]^T (fp16, K-major)
//   epilogue: + bias, ReLU, optional 2x2 max-pool, fp32 -> fp16, NHWC store.
//
// Replaces the cuDNN / cuBLAS calls behind torchvggish's ``features`` and ``embeddings``
// (reference call site fadtk/model_loader.py:107-108; shapes in SURVEY.md appendix A, K3/K4).
//
// Data movement: activations are NHWC fp16 in HBM.  One output tile covers a
// BW x BH x BN box of pixels (BW*BH*BN = 128); for filter tap (kh, kw) the A operand is the
// same box shifted by (kh-1, kw-1), fetched with ONE 4-D TMA whose out-of-bounds elements are
// zero-filled by hardware - that is the conv padding, and there is no im2col buffer.  TMA
// writes the 128-byte swizzled K-major layout tcgen05.mma consumes directly.
//
// Accumulation.  The tensor core adds into its fp32 accumulator with truncation: measured on
// B200, a K-long reduction comes out scaled by (1 - 1.0e-9 K) (tests/diag_accum_bias.py: -1.25e-5
// at K = 12288) - over the eight layers that alone moves the FAD by ~6e-5 relative.  So the MMA
// warp accumulates at most kChunkSteps x 64 = 512 of K into one TMEM buffer, and the epilogue
// warps sum the chunks in registers with ordinary round-to-nearest fp32 adds.  The two TMEM
// buffers alternate per CHUNK, so draining chunk c overlaps the MMAs of chunk c+1.
//
// Warp roles (384 threads, persistent over tiles):
//   warp 0  TMA producer (one elected lane)        warp 2  TMEM allocator
//   warp 1  MMA issuer   (one elected lane)        warps 4-11 epilogue: TMEM lane quarter = warp%4,
//                                                  column half = (warp-4)/4
// Pipelines: smem ring full[]/empty[] (TMA <-> MMA), TMEM chunk buffers tmem_full[]/tmem_empty[].
#pragma once
#include "sm100.cuh"

namespace fad {

struct ConvGemmParams {
    int taps;        // 9 (conv3x3, pad 1) or 1 (fully connected / 1x1)
    int cblks;       // Cin / 64
    int box_w, box_h, box_n;   // pixel box of one tile, product == 128
    int tiles_w, tiles_h;      // tiles per image along W and H
    int img_groups;            // ceil(NB / box_n)
    int n_tiles;               // Cout / N_TILE
    int H, W, NB, Cout;
    float lo_scale;            // WMODE 2: 2^-s of the E4M3 low parts
    int lo8_group;             // WMODE 2: k-steps whose E4M3 MMAs are issued together (1 .. min(4, STAGES - 2))
    int ld_out, n_valid;       // un-pooled outputs: row stride and number of columns actually stored
                               // (Cout is padded to the tile width; columns >= n_valid are dropped)
    int relu, pool;            // relu: 0 = none, 1 = ReLU, 2 = GELU (erf form), 3 = ELU
    const float* bias;         // [Cout]
    __half* out;               // NHWC fp16 [NB, H(/2), W(/2), Cout]; may be null when out_f32 is set
    float* out_f32;            // optional fp32 copy of the un-pooled output (may be null)
    uint8_t* out8;             // optional E4M3 copy of `out` (same layout) for a WMODE 2 consumer (may be null)
    // optional fused residual update (transformer blocks): resid[token(row)][0:resid_C] += result,
    // where row -> token undoes the (shifted-)window ordering of the rows (resid_res = 0: identity)
    float* resid;
    int resid_C, resid_res, resid_shift;
};

// token index (b*res*res + y*res + x) of window-ordered row o (8x8 windows, cyclic shift)
__device__ __forceinline__ long long window_row_to_token(long long o, int res, int shift) {
    const int lg_nw = 28 - __clz(res);                 // res is a power of two >= 8: log2(res / 8)
    const int nw = res >> 3;
    const int in = (int)(o & 63);
    const long long wi = o >> 6;
    const int wx = (int)wi & (nw - 1);
    const int wy = (int)(wi >> lg_nw) & (nw - 1);
    const long long b = wi >> (2 * lg_nw);
    int y = wy * 8 + (in >> 3) + shift, xx = wx * 8 + (in & 7) + shift;
    if (y >= res) y -= res;
    if (xx >= res) xx -= res;
    return (b * res + y) * res + xx;
}

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                        // fp16 elements per 128-B swizzled row
constexpr int kConvGemmThreads = 384;
constexpr int kEpilogueWarps = 8;
constexpr int kChunkSteps = 8;                     // k-steps (of 64) per TMEM accumulation chunk
// The tensor core truncates when it adds into its fp32 accumulator: a sum over T accumulated elements (T = K of the
// chunk, x 2 when the hi and lo weight MMAs share the accumulator) comes out scaled by (1 - 1.0e-9 T) - measured on
// B200 for random and post-ReLU operands, shapes K = 384 ... 12288 (tests/diag_accum_bias.py; profiles/
// r2_gemm_bias_probe_before.json: -8.1e-7 at T = 768, -1.08e-6 at T = 1024, -4.1e-7 at T = 384).  Cutting K into chunks
// bounds it, but what is left is SYSTEMATIC: through the ~50 GEMMs of a transformer encoder it adds up to a per-
// dimension offset of the hidden states (wav2vec: 8e-5 of their rms at layer 12, 10x what independent errors would
// give, and a -2e-4 offset of the FAD, profiles/r2_w2v_layer_bias_before.json).  The epilogue therefore scales every
// chunk by the inverse of its expected shrink when it sums the chunks in registers: v + v * eps as one FMA (1 + eps
// itself is not representable finely enough in fp32: eps ~ 1e-6 is only 8 ulps of 1).
constexpr float kAccumShrinkPerElement = 1.06e-9f;
constexpr uint32_t kABytes = kTileM * kBlockK * 2; // 16 KiB per stage
constexpr uint32_t kStagingBytes = 32 * 128;       // per epilogue warp: 32 rows x 128 B output staging

// SPLIT_W: the weights are an fp16 hi/lo pair (W = Wh + Wl, 22 bits).  fp16 rounding of the
// weights is a fixed perturbation of the model that does not average out over samples: it alone
// moves the FAD by ~1.4e-4 relative (CPU experiment), more than the whole 1e-4 budget,
// whereas fp16 activations cost 2e-5.  The hi and lo rows of one N tile are stored back to back
// ([Wh: N_TILE rows | Wl: N_TILE rows] per tile), so ONE TMA box brings both and the MMA warp
// issues A x Wh and A x Wl into the same TMEM accumulator.
// WMODE 0: fp16 weights.  1: fp16 hi/lo pair, two kind::f16 MMAs per K step.  2: fp16 hi + E4M3 lo:
// the low part (|Wl| <= 2^-11 |W|, needed to ~4 bits) is applied by a kind::f8f6f4 MMA - twice the
// rate and half the operand bytes of a second fp16 MMA - against an E4M3 copy of the activation
// (written next to the fp16 one by the producing kernel, fetched by its own TMA box), into its own
// TMEM accumulator that the epilogue adds with the power-of-two scale of the E4M3 weights.
// PAIR: two CTAs (one cluster, the two SMs of a TPC) share every MMA with cta_group::2 - M = 256, each CTA holds its
// own 128-row A tile and HALF of the weight tile, so a stage carries half the weight bytes per CTA (the kernel is
// bound by what each SM pulls from L2 per k-step, not by the tensor pipe: ncu, profiles/r2_ncu_conv_gemm_pair.md).
template <int N_TILE, int WMODE, int PAIR = 0>
__host__ __device__ constexpr uint32_t conv_gemm_stage_bytes() {
    constexpr int kNLoc = PAIR ? N_TILE / 2 : N_TILE;
    return WMODE == 2 ? kABytes + kNLoc * kBlockK * 2 + kNLoc * kBlockK + kTileM * kBlockK
                      : kABytes + (WMODE == 1 ? 2 : 1) * kNLoc * kBlockK * 2;
}

template <int N_TILE, int STAGES, int WMODE, int PAIR = 0>
__host__ __device__ constexpr uint32_t conv_gemm_smem_bytes() {
    return STAGES * conv_gemm_stage_bytes<N_TILE, WMODE, PAIR>() + 1024 /*align slack*/ + 256 /*barriers*/
         + kEpilogueWarps * kStagingBytes;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// GELU(x) = x/2 (1 + erf(x / sqrt 2)), the exact-erf form torch.nn.GELU() / HTSAT use, with
//   erf(z) = 1 - 2^(z q(z)),  z = |x| / sqrt 2 clamped to 4.3   (erfc(4.3) = 1.2e-9)
// q = degree-6 weighted-minimax fit of log2(erfc(z)) / z (oracle/fit_gelu.py): |erf error| <= 1.4e-7
// in fp32, far below the fp16 rounding of the value this feeds.  One MUFU (ex2) + ~13 FMA-pipe
// instructions per element - erff costs ~30, and two MUFUs made the fc1 epilogue MUFU-bound.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fminf(fabsf(x) * 0.70710678118654752f, 4.3f);
    float q = 1.04899843e-04f;
    q = fmaf(q, z, -4.92790774e-04f);
    q = fmaf(q, z, -2.22368206e-03f);
    q = fmaf(q, z, 2.93586859e-02f);
    q = fmaf(q, z, -1.48908889e-01f);
    q = fmaf(q, z, -9.18342944e-01f);
    q = fmaf(q, z, -1.62791250e+00f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * z));
    return 0.5f * x * (1.0f + copysignf(1.0f - e, x));
}

// ELU (alpha = 1) in ~10 instructions (expm1f is ~25, in an epilogue that is issue-bound): the negative branch is
// 2^(x log2 e) - 1 with one MUFU (absolute error 2^-22) below -1/16, and the Taylor polynomial of degree 4 above it, where
// the subtraction would cancel (truncation < 8e-9, relative error ~1e-7 of a result that is then rounded to fp16).
__device__ __forceinline__ float elu_ex2(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 0.f) * 1.4426950408889634f));
    const float t = x * fmaf(x, fmaf(x, fmaf(x, 4.16666667e-2f, 1.66666667e-1f), 0.5f), 1.0f);
    return x > 0.f ? x : (x > -0.0625f ? t : e - 1.0f);
}

// 8 halves -> 8 E4M3 bytes (round to nearest, saturate)
__device__ __forceinline__ uint32_t f16x4_to_e4m3x4(uint32_t a, uint32_t b) {
    __half2_raw h0, h1;
    h0.x = (unsigned short)(a & 0xffffu); h0.y = (unsigned short)(a >> 16);
    h1.x = (unsigned short)(b & 0xffffu); h1.y = (unsigned short)(b >> 16);
    return (uint32_t)__nv_cvt_halfraw2_to_fp8x2(h0, __NV_SATFINITE, __NV_E4M3)
         | ((uint32_t)__nv_cvt_halfraw2_to_fp8x2(h1, __NV_SATFINITE, __NV_E4M3) << 16);
}
__device__ __forceinline__ uint2 f16x8_to_e4m3x8(uint4 v) {
    return make_uint2(f16x4_to_e4m3x4(v.x, v.y), f16x4_to_e4m3x4(v.z, v.w));
}

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

template <int N_TILE, int STAGES, int WMODE, int PAIR = 0, int STACK = 0>
__global__ void __launch_bounds__(kConvGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_x,
                 const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_wl8,       // WMODE 2: E4M3 low parts [Cout, K]
                 const __grid_constant__ CUtensorMap map_x8,        // WMODE 2: E4M3 copy of the activation
                 const ConvGemmParams p)
{
    using namespace sm100;
    constexpr bool SPLIT_W = WMODE == 1;
    constexpr bool LO8 = WMODE == 2;
    constexpr uint32_t kStageBytes = conv_gemm_stage_bytes<N_TILE, WMODE, PAIR>();
    constexpr int kBRows = (WMODE != 0 ? 2 : 1) * N_TILE;           // rows of the packed weight tensor per N tile
    constexpr int kNLoc = PAIR ? N_TILE / 2 : N_TILE;               // weight rows (output channels) THIS CTA stages
    // STACKED (SPLIT_W && STACK; measurement variant, off by default): hi and lo weight rows as ONE B operand of N = 2 N_TILE
    // rows ([Wh | Wl] is how a stage holds them); the product lands in two column halves of the accumulator (hi: [0, N_TILE),
    // lo: [N_TILE, 2 N_TILE)) that the epilogue adds.  One MMA per K slice instead of two, A fetched from shared memory
    // once: 16 % faster on the bare tensor pipe (profiles/r2_umma_issue_patterns.json: 2.24 vs 1.89 PFLOP/s issued) - but
    // not in this kernel, which runs against the power cap, and it costs a second TMEM read per epilogue group
    // (profiles/r2_ncu_vggish_pair.md); the default accumulates A Wh^T and A Wl^T into the same TMEM tile.
    constexpr bool STACKED = SPLIT_W && STACK != 0;
    constexpr uint32_t kBufCols = (STACKED ? 2 : 1) * N_TILE;       // TMEM columns of one chunk buffer
    constexpr uint32_t kTmemCols = LO8 ? 4 * N_TILE : 2 * kBufCols; // two chunk buffers (+ two low-part buffers)
    constexpr uint32_t kIdesc = make_idesc(FMT_F16, PAIR ? 2 * kTileM : kTileM, STACKED ? 2 * N_TILE : N_TILE);
    constexpr uint32_t kIdesc8 = make_idesc(FMT_E4M3, PAIR ? 2 * kTileM : kTileM, N_TILE);
    constexpr uint32_t kWhBytes = kNLoc * kBlockK * 2;
    constexpr uint32_t kOffWl8 = kABytes + kWhBytes;                // stage layout (LO8): A16 | Wh | Wl8 | A8
    constexpr uint32_t kOffA8 = kOffWl8 + kNLoc * kBlockK;
    static_assert(!PAIR || WMODE != 0, "pairs are built for the split-weight modes");
    constexpr int kColsPerWarp = N_TILE / 2;            // each lane quarter is shared by two warps
    constexpr int kGroups = kColsPerWarp / 32;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint64_t* corr_empty = bars + 2 * STAGES + 4;                    // LO8: low-part accumulator drained
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
    uint8_t* staging = smem + STAGES * kStageBytes + 256;            // kEpilogueWarps x kStagingBytes

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ksteps = p.taps * p.cblks;
    const int n_chunks = (ksteps + kChunkSteps - 1) / kChunkSteps;
    const int chunk_len = (ksteps + n_chunks - 1) / n_chunks;       // balanced chunks
    const int m_tiles = p.img_groups * p.tiles_h * p.tiles_w;
    // work unit = one N tile x (one M tile | PAIR: two consecutive M tiles, one per CTA of the pair; the second of an
    // odd count is past the batch: its TMA boxes are zero-filled and its rows masked in the epilogue)
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int n_workers = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
    const int worker = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
    const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;
    const int total_tiles = m_units * p.n_tiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        if (LO8) { tma_prefetch_desc(&map_wl8); tma_prefetch_desc(&map_x8); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        // PAIR: the leader's tmem_empty / corr_empty collect the epilogue warps of BOTH CTAs
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], kEpilogueWarps * (PAIR ? 2 : 1)); }
        for (int a = 0; a < 2; ++a) mbar_init(&corr_empty[a], kEpilogueWarps * (PAIR ? 2 : 1));
        mbar_fence_init();
    }
    if (warp == 2) { if (PAIR) tmem_alloc_pair<kTmemCols>(tmem_base_slot); else tmem_alloc<kTmemCols>(tmem_base_slot); }
    tc_fence_before_sync();
    if (PAIR) cluster_sync(); else __syncthreads();           // PAIR: the peer's barriers exist before anything signals them
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_slot;

    // warps 0-3 (one warpgroup) only issue TMA / MMA: hand their registers to the two epilogue
    // warpgroups, which hold N_TILE/2 fp32 partial sums per thread.  128*88 + 256*208 <= 64 K.
    if (warp < 4) {
      setmaxnreg_dec<88>();
      if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int tile = worker; tile < total_tiles; tile += n_workers) {
                const int nt = tile % p.n_tiles;
                const int m = PAIR ? 2 * (tile / p.n_tiles) + (int)rank : tile / p.n_tiles;
                const int w0 = (m % p.tiles_w) * p.box_w;
                const int h0 = ((m / p.tiles_w) % p.tiles_h) * p.box_h;
                const int n0 = (m / (p.tiles_w * p.tiles_h)) * p.box_n;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int tap = ks / p.cblks;
                    const int cb = ks - tap * p.cblks;
                    int dh = 0, dw = 0;
                    if (p.taps == 9) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * kStageBytes;
                    if (PAIR) {
                        // both CTAs' bytes land on the LEADER's barrier (the MMA issuer waits there); a peer's bytes may
                        // complete before the leader arms the phase - the transaction count is signed, that is fine
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * kStageBytes);
                        const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
                        tma_load_4d_pair(st, &map_x, bar, cb * kBlockK, w0 + dw, h0 + dh, n0);
                        // this CTA's half of the hi rows (and of the lo rows: a second box).  STACKED: B of the pair's N = 2 N_TILE
                        // MMA is [Wh | Wl]: rank 0 stages all hi rows, rank 1 all lo rows (one box of N_TILE rows each)
                        if (STACKED) tma_load_2d_pair(st + kABytes, &map_w, bar, ks * kBlockK, nt * kBRows + (int)rank * N_TILE);
                        else {
                            tma_load_2d_pair(st + kABytes, &map_w, bar, ks * kBlockK, nt * kBRows + (int)rank * kNLoc);
                            if (SPLIT_W)
                                tma_load_2d_pair(st + kABytes + kWhBytes, &map_w, bar, ks * kBlockK, nt * kBRows + N_TILE + (int)rank * kNLoc);
                        }
                        if (LO8) {
                            tma_load_2d_pair(st + kOffWl8, &map_wl8, bar, ks * kBlockK, nt * N_TILE + (int)rank * kNLoc);
                            tma_load_4d_pair(st + kOffA8, &map_x8, bar, cb * kBlockK, w0 + dw, h0 + dh, n0);
                        }
                    } else {
                    mbar_expect_tx(&full[s], kStageBytes);
                    tma_load_4d(st, &map_x, &full[s], cb * kBlockK, w0 + dw, h0 + dh, n0);
                    tma_load_2d(st + kABytes, &map_w, &full[s], ks * kBlockK, nt * kBRows);
                    if (LO8) {
                        tma_load_2d(st + kOffWl8, &map_wl8, &full[s], ks * kBlockK, nt * N_TILE);
                        tma_load_4d(st + kOffA8, &map_x8, &full[s], cb * kBlockK, w0 + dw, h0 + dh, n0);
                    }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
      } else if (warp == 1 && rank == 0) {
        // -------------------------------------------------------------- MMA issuer (PAIR: the leader CTA only)
        if (elect_one()) {
            int s = 0; uint32_t ph = 0;
            int buf = 0; uint32_t buf_ph = 0;
            int cpar = 0; uint32_t cpar_ph = 0;
            int pend_stage[4] = {0, 0, 0, 0}; bool pend_first[4] = {false, false, false, false}; int n_pend = 0;
            for (int tile = worker; tile < total_tiles; tile += n_workers) {
                const uint32_t d_corr = tmem_base + 2 * N_TILE + cpar * N_TILE;
                if (LO8) { mbar_wait(&corr_empty[cpar], cpar_ph ^ 1); tc_fence_after_sync(); }
                for (int ks0 = 0; ks0 < ksteps; ks0 += chunk_len) {
                    const int ks1 = min(ks0 + chunk_len, ksteps);
                    mbar_wait(&tmem_empty[buf], buf_ph ^ 1);
                    tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + buf * kBufCols;
                    for (int ks = ks0; ks < ks1; ++ks) {
                        mbar_wait(&full[s], ph);
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
                        const uint64_t a_desc = kmajor_sw128_desc(a_addr);
                        const uint64_t b_desc = kmajor_sw128_desc(a_addr + kABytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            // +32 B along K inside the 128-B swizzle atom == +2 in the 16-B address field
                            if (PAIR) umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, kIdesc, (ks > ks0) || (k > 0));
                            else      umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, kIdesc, (ks > ks0) || (k > 0));
                            if (SPLIT_W && !STACKED) {         // lo rows: kNLoc rows (x 128 B) further down the stage, same accumulator
                                if (PAIR) umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k + (kNLoc * 128 / 16), kIdesc, 1);
                                else      umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k + (kNLoc * 128 / 16), kIdesc, 1);
                            }
                        }
                        if (LO8) {
                            // The E4M3 low-part MMAs of the last `lo8_group` k-steps are issued together, after their
                            // fp16 MMAs: alternating kind::f16 / kind::f8f6f4 every k-step drains the tensor pipe at
                            // each switch (ncu: tensor pipe 58-71 % with per-k-step alternation, r2_ncu_wlo8).  A stage
                            // is released once BOTH its MMAs have been issued, so a group holds `lo8_group` stages.
                            pend_stage[n_pend] = s; pend_first[n_pend] = (ks == 0); ++n_pend;
                            if (n_pend == p.lo8_group || ks == ks1 - 1) {
                                for (int q = 0; q < n_pend; ++q) {
                                    const uint32_t q_addr = smem_u32(smem + pend_stage[q] * kStageBytes);
                                    const uint64_t a8_desc = kmajor_sw64_desc(q_addr + kOffA8);
                                    const uint64_t w8_desc = kmajor_sw64_desc(q_addr + kOffWl8);
#pragma unroll
                                    for (int k = 0; k < kBlockK / 32; ++k) { // K = 32 per kind::f8f6f4 MMA, +32 B inside the 64-B atom
                                        if (PAIR) umma_f8_pair(d_corr, a8_desc + 2 * k, w8_desc + 2 * k, kIdesc8, !pend_first[q] || (k > 0));
                                        else      umma_f8(d_corr, a8_desc + 2 * k, w8_desc + 2 * k, kIdesc8, !pend_first[q] || (k > 0));
                                    }
                                    if (PAIR) umma_commit_pair(&empty[pend_stage[q]]); else umma_commit(&empty[pend_stage[q]]);
                                }
                                n_pend = 0;
                            }
                        } else {
                            if (PAIR) umma_commit_pair(&empty[s]); else umma_commit(&empty[s]);   // smem slot(s) free once these MMAs retire
                        }
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                    if (PAIR) umma_commit_pair(&tmem_full[buf]); else umma_commit(&tmem_full[buf]);   // chunk complete -> epilogue warps
                    if (++buf == 2) { buf = 0; buf_ph ^= 1; }
                }
                if (++cpar == 2) { cpar = 0; cpar_ph ^= 1; }
            }
        }
      }
    } else {
        // ---------------------------------------------------------------- epilogue
        setmaxnreg_inc<208>();
        const int q = warp & 3;                           // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;                 // which half of the tile's columns
        const int r = q * 32 + lane;                      // row of the tile == TMEM lane
        const int bw = p.box_w, bh = p.box_h;
        const int pw = r % bw;
        const int phh = (r / bw) % bh;
        const int pn = r / (bw * bh);
        int buf = 0; uint32_t buf_ph = 0;
        int cpar = 0;
        // plain row-major GEMM (1x1 "image", 128 rows per tile): no per-tile divisions
        const bool plain = p.tiles_w == 1 && p.tiles_h == 1 && bw == 1 && bh == 1;
        // (nt, mu) of this worker's current unit, stepped without divisions; m = the M tile of THIS CTA
        int nt = worker % p.n_tiles, mu = worker / p.n_tiles;
        const int step_nt = n_workers % p.n_tiles, step_m = n_workers / p.n_tiles;
        int m = PAIR ? 2 * mu + (int)rank : mu;
        const uint32_t te_addr[2] = {PAIR ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u, PAIR ? mapa_u32(smem_u32(&tmem_empty[1]), 0) : 0u};
        const uint32_t ce_addr[2] = {PAIR ? mapa_u32(smem_u32(&corr_empty[0]), 0) : 0u, PAIR ? mapa_u32(smem_u32(&corr_empty[1]), 0) : 0u};
        // residual rows are read-modify-written in the epilogue: pull the NEXT tile's rows into L2
        // while this tile is computed, so the loads do not pay HBM latency on the critical path
        auto prefetch_resid = [&](int nt_, int m_) {
            if (p.resid == nullptr || !plain) return;
            const int n_ = m_ * kTileM + r;
            if (n_ >= p.NB) return;
            const long long tok = p.resid_res ? window_row_to_token((long long)n_, p.resid_res, p.resid_shift) : (long long)n_;
            const int c0 = nt_ * N_TILE + half * kColsPerWarp;
            for (int c = c0; c < c0 + kColsPerWarp && c < p.resid_C; c += 32)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(p.resid + tok * p.resid_C + c));
        };
        if (mu < m_units) prefetch_resid(nt, m);
        // The unit index itself is not carried: unit = mu * n_tiles + nt < total_tiles  <=>  mu < m_units.  (One loop scalar
        // less: the round-1 form kept the counter and tmem_base in LOCAL memory - ptxas spills what crosses the setmaxnreg
        // split - and reloaded both on the critical path of every tile: 9 % of the epilogue warps' stall samples on the
        // CLAP layers, profiles/r2_ncu_clap_gemm.md.)
        auto load_bias = [&](float4 (&b)[8], int col0) {
            const float4* src = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = __ldg(src + j);
        };
        while (mu < m_units) {
            int w, h, n;
            if (plain) { w = 0; h = 0; n = m * kTileM + r; }
            else {
                w = (m % p.tiles_w) * bw + pw;
                h = ((m / p.tiles_w) % p.tiles_h) * bh + phh;
                n = (m / (p.tiles_w * p.tiles_h)) * p.box_n + pn;
            }
            const bool valid = n < p.NB;
            int nt_next = nt + step_nt, mu_next = mu + step_m;
            if (nt_next >= p.n_tiles) { nt_next -= p.n_tiles; ++mu_next; }
            const int m_next = PAIR ? 2 * mu_next + (int)rank : mu_next;
            if (mu_next < m_units) prefetch_resid(nt_next, m_next);
            // bias of the first 32-column group: requested BEFORE the wait for the accumulator (the mbarrier / TMEM asm
            // statements are compiler barriers: a load written after them is issued after them, and its latency then
            // sits on the critical path of the tile - the single largest stall of the epilogue in the ncu source page)
            constexpr bool kBiasAhead = kColsPerWarp <= 64;  // N_TILE = 256 holds 128 partial sums per thread: no room for it
            float4 bnext[8];
            if (kBiasAhead) load_bias(bnext, nt * N_TILE + half * kColsPerWarp);
            uint32_t tmem_base_e;                            // re-read per tile (LDS) instead of a local-memory reload
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base_e) : "r"(smem_u32(tmem_base_slot)));

            // sum the K chunks in registers (round-to-nearest adds), undoing the expected truncation shrink of each
            float acc[kColsPerWarp];
            for (int c = 0; c < n_chunks; ++c) {
                const int len_c = min(chunk_len, ksteps - c * chunk_len);
                // products accumulated per TMEM element: hi and lo share one accumulator (x 2) unless they are stacked side by side
                const float unshrink = kAccumShrinkPerElement * (float)(len_c * kBlockK * (SPLIT_W && !STACKED ? 2 : 1));
                mbar_wait(&tmem_full[buf], buf_ph);
                tc_fence_after_sync();
                const uint32_t t_row = tmem_base_e + (uint32_t(q * 32) << 16) + buf * kBufCols + half * kColsPerWarp;
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + g * 32, v);
                    if (STACKED) {                             // A Wh^T + A Wl^T: the two column halves of the stacked product
                        uint32_t w[32];
                        tmem_ld_32x32(t_row + N_TILE + g * 32, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
                    } else {
                        tmem_ld_wait();
                    }
                    if (c == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[g * 32 + j] = fmaf(__uint_as_float(v[j]), unshrink, __uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[g * 32 + j] += fmaf(__uint_as_float(v[j]), unshrink, __uint_as_float(v[j]));
                    }
                    if (LO8 && c == n_chunks - 1) {           // + A8 * Wl8^T / 2^s: complete once the last chunk is
                        tmem_ld_32x32(tmem_base_e + (uint32_t(q * 32) << 16) + 2 * N_TILE + cpar * N_TILE + half * kColsPerWarp + g * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[g * 32 + j] = fmaf(__uint_as_float(v[j]), p.lo_scale, acc[g * 32 + j]);
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) {                               // the leader's MMA warp waits for both CTAs' drains
                        mbar_arrive_cluster(te_addr[buf]);
                        if (LO8 && c == n_chunks - 1) mbar_arrive_cluster(ce_addr[cpar]);
                    } else {
                        mbar_arrive(&tmem_empty[buf]);
                        if (LO8 && c == n_chunks - 1) mbar_arrive(&corr_empty[cpar]);
                    }
                }
                if (++buf == 2) { buf = 0; buf_ph ^= 1; }
            }
            cpar ^= 1;

            const int ch0 = nt * N_TILE + half * kColsPerWarp;
            // Destination of this lane's row (element offsets, -1 = row beyond the batch).  Stores go
            // through a per-warp staging tile (32 rows x 128 B, 16-B chunks XOR-swizzled by row) so a
            // warp writes whole 128-B lines of 4 rows per instruction instead of 16 B into 32
            // different rows: the direct pattern made the small-K GEMMs LSU-bound (ncu: 32 sectors
            // per store request, stall_lg/long_scoreboard on the bias loads queued behind them).
            long long out_off = -1, res_off = -1;
            if (valid && !p.pool) {
                out_off = (long long)((size_t(n) * p.H + h) * p.W + w) * p.ld_out;
                if (p.resid != nullptr)
                    res_off = (p.resid_res ? window_row_to_token((long long)n, p.resid_res, p.resid_shift) : (long long)n)
                              * p.resid_C;
            }
            const uint32_t stg = smem_u32(staging) + (warp - 4) * kStagingBytes;     // shared-space addresses (sts128 / lds128)
            const uint32_t stg_mine = stg + lane * 128;
            const int sw = lane & 7;
            const int cq = lane & 7, rq = lane >> 3;      // flush role: 16-B chunk cq of rows it*4 + rq
            const bool f32_path = p.out_f32 != nullptr || p.resid != nullptr;
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                float f[32];
                if (!kBiasAhead) load_bias(bnext, ch0 + g * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = bnext[j];
                    f[4 * j + 0] = acc[g * 32 + 4 * j + 0] + b.x;
                    f[4 * j + 1] = acc[g * 32 + 4 * j + 1] + b.y;
                    f[4 * j + 2] = acc[g * 32 + 4 * j + 2] + b.z;
                    f[4 * j + 3] = acc[g * 32 + 4 * j + 3] + b.w;
                }
                if (kBiasAhead && g + 1 < kGroups) load_bias(bnext, ch0 + (g + 1) * 32);   // the next group's, one group ahead
                if (p.relu == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
                } else if (p.relu == 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                } else if (p.relu == 3) {                           // ELU (alpha = 1): SEANet's activation
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = elu_ex2(f[j]);
                }
                uint32_t h2[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) h2[j] = pack_half2(f[2 * j], f[2 * j + 1]);

                if (!p.pool) {
                    if (f32_path) {
                        // fp32 outputs: this group's 32 columns = one 128-B row segment per row
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            sts128(stg_mine + ((j ^ sw) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                                   __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
                        __syncwarp();
                        float4 v[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rr = it * 4 + rq;
                            const uint4 u = lds128(stg + rr * 128 + ((cq ^ (rr & 7)) << 4));
                            v[it] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
                        }
                        const int col = ch0 + g * 32 + cq * 4;
                        if (p.out_f32 != nullptr) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const long long off = __shfl_sync(0xffffffffu, out_off, it * 4 + rq);
                                if (off >= 0 && col < p.n_valid) *reinterpret_cast<float4*>(p.out_f32 + off + col) = v[it];
                            }
                        }
                        if (p.resid != nullptr && ch0 + g * 32 < p.resid_C) {     // resid_C is a multiple of 32
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {               // 4 loads in flight, then 4 stores
                                long long offs[4];
                                float4 xv[4];
#pragma unroll
                                for (int it = 0; it < 4; ++it) {
                                    offs[it] = __shfl_sync(0xffffffffu, res_off, (hf * 4 + it) * 4 + rq);
                                    if (offs[it] >= 0) xv[it] = *reinterpret_cast<const float4*>(p.resid + offs[it] + col);
                                }
#pragma unroll
                                for (int it = 0; it < 4; ++it) {
                                    if (offs[it] >= 0) {
                                        const float4 a = v[hf * 4 + it];
                                        xv[it].x += a.x; xv[it].y += a.y; xv[it].z += a.z; xv[it].w += a.w;
                                        *reinterpret_cast<float4*>(p.resid + offs[it] + col) = xv[it];
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (p.out != nullptr && out_off >= 0 && ch0 + g * 32 + 32 <= p.n_valid) {   // rare (last VGGish layer): both precisions, direct
                            uint4* dst = reinterpret_cast<uint4*>(p.out + out_off + ch0 + g * 32);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                dst[j] = make_uint4(h2[4 * j], h2[4 * j + 1], h2[4 * j + 2], h2[4 * j + 3]);
                        }
                    } else if (p.out != nullptr) {
                        // fp16 output: two groups (64 columns) fill the 128-B row segment, then flush
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            sts128(stg_mine + ((((g & 1) * 4 + j) ^ sw) << 4), h2[4 * j], h2[4 * j + 1], h2[4 * j + 2], h2[4 * j + 3]);
                        if (g & 1) {
                            __syncwarp();
                            const int col = ch0 + (g - 1) * 32 + cq * 8;
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int rr = it * 4 + rq;
                                const uint4 v = lds128(stg + rr * 128 + ((cq ^ (rr & 7)) << 4));
                                const long long off = __shfl_sync(0xffffffffu, out_off, rr);
                                if (off >= 0 && col < p.n_valid) {
                                    *reinterpret_cast<uint4*>(p.out + off + col) = v;
                                    if (p.out8 != nullptr) *reinterpret_cast<uint2*>(p.out8 + off + col) = f16x8_to_e4m3x8(v);
                                }
                            }
                            __syncwarp();
                        }
                    }
                } else {
                    // 2x2 max-pool: partners are lane^1 (w) and lane^box_w (h), box_w in {8,16}
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t o = __shfl_xor_sync(0xffffffffu, h2[j], 1);
                        h2[j] = hmax2_u32(h2[j], o);
                        o = __shfl_xor_sync(0xffffffffu, h2[j], bw);
                        h2[j] = hmax2_u32(h2[j], o);
                    }
                    // the four lanes of a 2x2 group now hold the same 32 channels; each stores 8
                    const int sub = (pw & 1) | ((phh & 1) << 1);
                    uint4 o;
                    o.x = sub == 0 ? h2[0] : sub == 1 ? h2[4] : sub == 2 ? h2[8]  : h2[12];
                    o.y = sub == 0 ? h2[1] : sub == 1 ? h2[5] : sub == 2 ? h2[9]  : h2[13];
                    o.z = sub == 0 ? h2[2] : sub == 1 ? h2[6] : sub == 2 ? h2[10] : h2[14];
                    o.w = sub == 0 ? h2[3] : sub == 1 ? h2[7] : sub == 2 ? h2[11] : h2[15];
                    if (valid) {
                        const size_t pix = (size_t(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                        *reinterpret_cast<uint4*>(p.out + pix * p.Cout + ch0 + g * 32 + sub * 8) = o;
                        if (p.out8 != nullptr) *reinterpret_cast<uint2*>(p.out8 + pix * p.Cout + ch0 + g * 32 + sub * 8) = f16x8_to_e4m3x8(o);
                    }
                }
            }
            nt = nt_next; mu = mu_next; m = m_next;
        }
    }

    tc_fence_before_sync();
    if (PAIR) cluster_sync(); else __syncthreads();           // PAIR: neither CTA may exit while the other still signals it
    if (warp == 2) { if (PAIR) tmem_dealloc_pair<kTmemCols>(tmem_base); else tmem_dealloc<kTmemCols>(tmem_base); }
}

}  // namespace fad
