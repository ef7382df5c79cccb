// Host side of the C ABI declared in include/fadtk_b200.h: owns device weights and workspaces,
// encodes TMA descriptors, launches the sm_100a kernels.  No torch, no CUTLASS.
#include "../../include/fadtk_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "conv_gemm.cuh"
#include "frontend.cuh"
#include "stats.cuh"
#include "frechet.cuh"
#include "frechet_batched.cuh"
#include "clap.cuh"
#include "resample.cuh"
#include "whisper.cuh"
#include "encodec.cuh"
#include "wav2vec.cuh"
#include "umma_bench.cuh"
#include "attention_umma.cuh"

namespace {

thread_local std::string g_err;

int fail(const std::string& m) { g_err = m; return 1; }

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 tensor, innermost dimension first; 128-B swizzle, zero fill out of bounds.
int encode_f16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                    gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
                 (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
        return fail(buf);
    }
    return 0;
}

struct LayerGeom {
    int taps, Cin, Cout, H, W, relu, pool;
    int box_w, box_h, box_n, n_tile;
    int split_w;      // 0: fp16 weights; 1: fp16 hi/lo pair (interleaved per 128-row tile), two fp16 MMAs;
                      // 2: same packed tensor, low part applied as E4M3 (kind::f8f6f4) - see conv_gemm.cuh
    int pair;         // 1: CTA pairs (cta_group::2, M = 256 per MMA, each CTA stages half of the weight tile)
    int stack;        // 1 (mode 1 only, FADTK_STACK=1): hi | lo rows as one N = 2 n_tile MMA operand (measurement variant)
};

// how the low part of split weights is applied: FADTK_WLO=fp16 (two fp16 MMAs) | fp8 (E4M3 correction MMA)
int wlo_mode() {
    static const int mode = [] { const char* e = getenv("FADTK_WLO"); return (e && std::string(e) == "fp8") ? 2 : 1; }();
    return mode;
}
// CTA pairs (cta_group::2) per VGGish tensor-core layer, bit i = conv2, conv3_1, conv3_2, conv4_1, conv4_2, fc1, fc2, fc3.
// FADTK_PAIR = auto (default) | 0 | 1 (every layer) | 0x<mask>.  auto = the layers where the pair measured faster
// (profiles/r2_bench_pair_*.json): K >= 2304 convolutions and the FC layers; conv2 / conv3_1 (K = 576 / 1152, epilogue-
// heavy: one output tile per 9 / 18 k-steps) lose a little to the lock-step of two SMs and stay single-CTA.
unsigned pair_mask() {
    static const unsigned mask = [] {
        const char* e = getenv("FADTK_PAIR");
        if (!e || std::string(e) == "auto") return 0x7Cu;
        if (std::string(e) == "1") return 0xFFu;
        return (unsigned)strtoul(e, nullptr, 0) & 0xFFu;
    }();
    return mask;
}
int pair_all() {                    // stage-test entry (fad_umma_layer): pairs only when explicitly forced on
    const char* e = getenv("FADTK_PAIR");
    return (e && std::string(e) == "1") ? 1 : 0;
}

int make_geom(LayerGeom& g, int H, int W, int Cin, int Cout, int taps, int relu, int pool, int split_w) {
    g.taps = taps; g.Cin = Cin; g.Cout = Cout; g.H = H; g.W = W; g.relu = relu; g.pool = pool;
    g.split_w = split_w;                     // 0, 1 or 2 - the caller decides (wlo_mode() for the VGGish pipeline)
    g.pair = 0;                              // set by the caller after make_geom (split modes only)
    {
        static const int stk = [] { const char* e = getenv("FADTK_STACK"); return (e && e[0] == '1') ? 1 : 0; }();
        g.stack = (g.split_w == 1 && stk) ? 1 : 0;
    }
    if (Cin % 64 != 0) return fail("Cin must be a multiple of 64");
    if (taps != 1 && taps != 9) return fail("taps must be 1 or 9");
    if (H == 1 && W == 1) { g.box_w = 1; g.box_h = 1; g.box_n = 128; }
    else {
        g.box_w = W < 16 ? W : 16;
        if (g.box_w != 8 && g.box_w != 16) return fail("W must be 8 or a multiple of 16");
        if (W % g.box_w != 0) return fail("W must be a multiple of the 16-pixel box");
        int bh = 1;
        while (bh * 2 <= 128 / g.box_w && H % (bh * 2) == 0) bh *= 2;
        g.box_h = bh;
        g.box_n = 128 / (g.box_w * g.box_h);
    }
    if (pool && (g.box_h % 2 != 0 || g.box_w % 2 != 0)) return fail("pooling needs even tile boxes");
    if (Cout % 128 != 0) return fail("Cout must be a multiple of 128");
    g.n_tile = (!g.split_w && Cout % 256 == 0) ? 256 : 128;
    return 0;
}

}  // namespace

struct fad_handle {
    int device = 0;
    int num_sms = 0;
    int max_examples = 0;
    long long launches = 0;

    // front-end tables
    double *d_twiddle = nullptr, *d_hann = nullptr, *d_melw = nullptr;
    int *d_mel_start = nullptr, *d_mel_count = nullptr;

    // VGGish parameters (device)
    bool vgg_loaded = false;
    float *conv1_w = nullptr, *conv1_b = nullptr;
    __half* conv_w[5] = {};
    float* conv_b[5] = {};
    __half* fc_w[3] = {};
    float* fc_b[3] = {};

    // activations (device), sized for max_examples
    float* logmel = nullptr;
    __half* act[9] = {};       // act[0]=conv1 out ... act[5]=conv6 out (flattened), act[6..7]=fc1, fc2 out

    // per-layer cached descriptors for the fixed VGGish pipeline
    CUtensorMap map_x[8], map_w[8], map_x8[8];
    uint8_t* act8[8] = {};     // E4M3 copies of act[0..7] (inputs of the 8 tensor-core layers) when the low parts run in fp8
    uint8_t* x8_scratch = nullptr;  size_t x8_scratch_cap = 0;   // fad_umma_layer (stage test) only
    LayerGeom geom[8];

    // statistics workspace
    double *ws_tiles = nullptr, *ws_sums = nullptr;
    size_t ws_tiles_cap = 0, ws_sums_cap = 0;
    __half* gather_buf = nullptr;
    size_t gather_cap = 0;

    // Frechet workspace (fp64 d x d matrices)
    double* fr_buf = nullptr;
    size_t fr_cap = 0;
    unsigned char* frb_buf = nullptr;   // fad_frechet_batched workspace
    size_t frb_cap = 0;
    float* rs_bank = nullptr;  size_t rs_bank_cap = 0;  int rs_in = 0, rs_out = 0;     // resampler filter bank
    float* rs_mono = nullptr;  size_t rs_mono_cap = 0;
    struct Lo8 { uint8_t* w8; float inv_scale; };
    std::map<const void*, Lo8> lo8;          // E4M3 low parts per packed weight tensor (built on first use)
    double* fr_scal = nullptr;   // 32 doubles

    void* nccl_comm = nullptr;   // ncclComm_t created by fad_comm_init (NCCL is dlopen'ed, never linked)

    void* clap_state = nullptr;  // ClapState (clap_host.inc)
    void* whisper_state = nullptr;   // WhisperState (whisper_host.inc)
    void* encodec_state = nullptr;   // EncodecState (encodec_host.inc)
    void* w2v_state = nullptr;       // W2vState (wav2vec_host.inc)

    // optional per-category timing with CUDA events recorded on the launching stream
    bool prof_on = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int cat; size_t e0, e1; };
    std::vector<Span> spans;
    double prof_ms[FAD_PROF_CATEGORIES] = {};
    long long prof_count[FAD_PROF_CATEGORIES] = {};
};

namespace {

template <int N_TILE, int STAGES, int WMODE, int PAIR = 0, int STACK = 0>
int launch_conv_gemm(fad_handle* h, const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mw8,
                     const CUtensorMap& mx8, const fad::ConvGemmParams& p, cudaStream_t st) {
    static bool attr_set = false;
    constexpr uint32_t smem = fad::conv_gemm_smem_bytes<N_TILE, STAGES, WMODE, PAIR>();
    static_assert(smem <= 227 * 1024, "over the per-CTA shared-memory limit");
    auto kern = fad::conv_gemm_kernel<N_TILE, STAGES, WMODE, PAIR, STACK>;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int m_tiles = p.img_groups * p.tiles_h * p.tiles_w;
    const int total = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;           // work units (PAIR: two M tiles each)
    if (total == 0) return 0;
    if (PAIR) {
        // one cluster of two CTAs per unit in flight: the pair lands on the two SMs of a TPC
        const int pairs = total < h->num_sms / 2 ? total : h->num_sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(fad::kConvGemmThreads);
        cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, mx, mw, mw8, mx8, p));
    } else {
        const int grid = total < h->num_sms ? total : h->num_sms;
        kern<<<grid, fad::kConvGemmThreads, smem, st>>>(mx, mw, mw8, mx8, p);
    }
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

// a_cols: channels actually stored per pixel/row of the activation (<= g.Cin); the TMA box reads
// zeros beyond it, so a K that is not a multiple of 64 needs no padding columns in HBM.
int encode_layer_maps(const LayerGeom& g, const void* x, long long nb_dim, const void* w,
                      CUtensorMap* mx, CUtensorMap* mw, int a_cols = 0, long long a_row_stride = 0) {
    const uint64_t ac = a_cols > 0 ? (uint64_t)a_cols : (uint64_t)g.Cin;
    const uint64_t xd[4] = {ac, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)nb_dim};
    uint64_t xs[3] = {ac * 2, (uint64_t)g.W * ac * 2, (uint64_t)g.H * g.W * ac * 2};
    // 1x1 geometry only: rows a_row_stride elements apart (< a_cols = overlapping rows, e.g. the sliding
    // windows of a strided 1-D convolution read straight from the [T][C] activation, no im2col copy)
    if (a_row_stride > 0) xs[0] = xs[1] = xs[2] = (uint64_t)a_row_stride * 2;
    const uint32_t xb[4] = {64, (uint32_t)g.box_w, (uint32_t)g.box_h, (uint32_t)g.box_n};
    if (encode_f16_map(mx, x, 4, xd, xs, xb)) return 1;
    const uint64_t K = (uint64_t)g.taps * g.Cin;
    const uint64_t rows_mul = g.split_w ? 2 : 1;
    const uint64_t wd[2] = {K, (uint64_t)g.Cout * rows_mul};
    const uint64_t ws[1] = {K * 2};
    // rows per weight box: mode 1 fetches hi + lo of a tile at once, mode 2 the hi rows only; a CTA of a pair fetches
    // its half of the hi rows and its half of the lo rows as two boxes
    // (mode 1 pair: one box of n_tile rows - rank 0 the hi rows, rank 1 the lo rows of the stacked N = 2 n_tile operand)
    const uint32_t wb[2] = {64, (uint32_t)(g.pair ? (g.stack ? g.n_tile : g.n_tile / 2) : g.n_tile * (g.split_w == 1 ? 2 : 1))};
    return encode_f16_map(mw, w, 2, wd, ws, wb);
}

int lo8_for(fad_handle* h, const LayerGeom& g, const void* w, CUtensorMap* mw8, float* inv_scale, cudaStream_t st);

// Encoder self-attention on tcgen05 (attention_umma.cuh).  qkv: fp16 [clips * S][3 d] (q | k | v, head h at column h * 64),
// out: fp16 [clips * S][d].  FADTK_ATTN=legacy selects the mma.sync flash kernel (whisper.cuh) instead.
bool attention_umma_enabled() {
    static const bool on = [] { const char* e = getenv("FADTK_ATTN"); return !(e && std::string(e) == "legacy"); }();
    return on;
}
int launch_attention_umma(fad_handle* h, const __half* qkv, long long n_clips, int S, int d, int heads, __half* out, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(fad::attention_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fad::kAtSmem));
        CK(cudaFuncSetAttribute(fad::attention_umma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));   // two CTAs of 112 KiB per SM
        attr_set = true;
    }
    if (heads * 64 != d) return fail("attention_umma: head dimension must be 64");
    CUtensorMap map;
    const uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)S, (uint64_t)n_clips};
    const uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)S * 3 * d * 2};
    const uint32_t box[3] = {64, 128, 1};
    if (encode_f16_map(&map, qkv, 3, dims, strides, box)) return 1;
    fad::AttnParams p;
    p.S = S; p.d = d; p.heads = heads; p.out = out;
    fad::attention_umma_kernel<<<dim3((S + 127) / 128, heads, (unsigned)n_clips), fad::kAtThreads, fad::kAtSmem, st>>>(map, p);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

// the E4M3 low parts are cached per weight POINTER: drop the entry whenever that memory is rewritten
void lo8_forget(fad_handle* h, const void* w) {
    auto it = h->lo8.find(w);
    if (it != h->lo8.end()) { cudaFree(it->second.w8); h->lo8.erase(it); }
}

// E4M3 copy of an NHWC activation: same box geometry as the fp16 map, 64-B rows, SWIZZLE_64B
int encode_x8_map(const LayerGeom& g, const void* x8, long long nb_dim, CUtensorMap* mx8) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)nb_dim};
    cuuint64_t gstr[3] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W * g.Cin, (cuuint64_t)g.H * g.W * g.Cin};
    cuuint32_t bdim[4] = {64, (cuuint32_t)g.box_w, (cuuint32_t)g.box_h, (cuuint32_t)g.box_n}, estr[4] = {1, 1, 1, 1};
    CUresult r = fn(mx8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(x8), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (E4M3 activation) failed");
    return 0;
}

int run_layer(fad_handle* h, const LayerGeom& g, const CUtensorMap& mx, const CUtensorMap& mw, const void* w,
              int NB, const float* bias, void* out, float* out_f32, cudaStream_t st,
              float* resid = nullptr, int resid_C = 0, int resid_res = 0, int resid_shift = 0, int n_valid = 0,
              const CUtensorMap* mx8 = nullptr, uint8_t* out8 = nullptr) {
    fad::ConvGemmParams p;
    p.taps = g.taps; p.cblks = g.Cin / 64;
    p.box_w = g.box_w; p.box_h = g.box_h; p.box_n = g.box_n;
    p.tiles_w = g.W / g.box_w; p.tiles_h = g.H / g.box_h;
    p.img_groups = (NB + g.box_n - 1) / g.box_n;
    p.n_tiles = g.Cout / g.n_tile;
    p.H = g.H; p.W = g.W; p.NB = NB; p.Cout = g.Cout;
    p.n_valid = n_valid > 0 ? n_valid : g.Cout;
    p.ld_out = p.n_valid;
    p.relu = g.relu; p.pool = g.pool;
    p.bias = bias; p.out = reinterpret_cast<__half*>(out); p.out_f32 = out_f32; p.out8 = out8;
    p.resid = resid; p.resid_C = resid_C; p.resid_res = resid_res; p.resid_shift = resid_shift;
    p.lo_scale = 0.0f;
    p.lo8_group = 1;
    {
        static const int grp = [] { const char* e = getenv("FADTK_LO8_GROUP"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 3 ? 3 : v); }();
        p.lo8_group = grp;                                 // <= STAGES - 2 so the producer always has a stage to fill
    }
    if (g.split_w == 2) {
        if (mx8 == nullptr) return fail("fp8 low-part mode needs the E4M3 copy of the activation");
        CUtensorMap mw8;
        if (lo8_for(h, g, w, &mw8, &p.lo_scale, st)) return 1;
        if (g.pair) return launch_conv_gemm<128, 5, 2, 1>(h, mx, mw, mw8, *mx8, p, st);
        return launch_conv_gemm<128, 4, 2>(h, mx, mw, mw8, *mx8, p, st);
    }
    if (g.split_w && g.pair && g.stack) return launch_conv_gemm<128, 6, 1, 1, 1>(h, mx, mw, mw, mx, p, st);
    if (g.split_w && g.pair) return launch_conv_gemm<128, 6, 1, 1>(h, mx, mw, mw, mx, p, st);
    if (g.split_w && g.stack) return launch_conv_gemm<128, 4, 1, 0, 1>(h, mx, mw, mw, mx, p, st);
    if (g.split_w) return launch_conv_gemm<128, 4, 1>(h, mx, mw, mw, mx, p, st);
    if (g.n_tile == 256) return launch_conv_gemm<256, 4, 0>(h, mx, mw, mw, mx, p, st);
    return launch_conv_gemm<128, 6, 0>(h, mx, mw, mw, mx, p, st);
}

// E4M3 copy of the low parts of a packed hi/lo weight tensor, scaled by a power of two so the largest
// |Wl| lands near 224 (E4M3 max 448), plus its tensor map (64-B rows, SWIZZLE_64B).  Built once per tensor.
__global__ void wlo_absmax_kernel(const __half* __restrict__ w, long long n_tiles, long long K, unsigned int* __restrict__ out) {
    float m = 0.f;
    const long long total = n_tiles * 128 * K;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / K, k = e - row * K;
        const long long t = row >> 7, j = row & 127;
        m = fmaxf(m, fabsf(__half2float(w[((t * 256 + 128 + j) * K) + k])));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
__global__ void wlo_to_e4m3_kernel(const __half* __restrict__ w, long long n_tiles, long long K, float scale, uint8_t* __restrict__ out) {
    const long long total = n_tiles * 128 * K;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / K, k = e - row * K;
        const long long t = row >> 7, j = row & 127;
        const float v = __half2float(w[((t * 256 + 128 + j) * K) + k]) * scale;
        out[e] = (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
    }
}

__global__ void f16_to_e4m3_kernel(const __half* __restrict__ x, size_t count, uint8_t* __restrict__ out) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x)
        out[e] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(x[e]), __NV_SATFINITE, __NV_E4M3);
}

int lo8_for(fad_handle* h, const LayerGeom& g, const void* w, CUtensorMap* mw8, float* inv_scale, cudaStream_t st) {
    const long long K = (long long)g.taps * g.Cin, n_tiles = g.Cout / 128;
    auto it = h->lo8.find(w);
    if (it == h->lo8.end()) {
        unsigned int* d_max = nullptr;
        uint8_t* w8 = nullptr;
        CK(cudaMalloc(&d_max, 4));
        CK(cudaMalloc(&w8, (size_t)(n_tiles * 128 * K)));
        CK(cudaMemsetAsync(d_max, 0, 4, st));
        const unsigned blocks = (unsigned)std::min<long long>((n_tiles * 128 * K + 255) / 256, (long long)h->num_sms * 16);
        wlo_absmax_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(w), n_tiles, K, d_max);
        float mx = 0.f;
        CK(cudaMemcpyAsync(&mx, d_max, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int e = 0;
        if (mx > 0.f) e = (int)std::floor(std::log2(224.0f / mx));
        if (e > 100) e = 100;
        const float scale = std::ldexp(1.0f, e);
        wlo_to_e4m3_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(w), n_tiles, K, scale, w8);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        cudaFree(d_max);
        it = h->lo8.emplace(w, fad_handle::Lo8{w8, 1.0f / scale}).first;
    }
    *inv_scale = it->second.inv_scale;
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)(n_tiles * 128)};
    cuuint64_t gstr[1] = {(cuuint64_t)K};
    cuuint32_t bdim[2] = {64, (cuuint32_t)(g.pair ? 64 : 128)}, estr[2] = {1, 1};
    CUresult r = fn(mw8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, it->second.w8, gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (E4M3 low parts) failed");
    return 0;
}

// VGGish layer table: H, W are the conv's spatial size (input == un-pooled output)
struct VggLayer { int H, W, Cin, Cout, taps, relu, pool; };
const VggLayer kVgg[8] = {
    {48, 32, 64, 128, 9, 1, 1},     // conv2  -> [24,16,128]
    {24, 16, 128, 256, 9, 1, 0},    // conv3_1
    {24, 16, 256, 256, 9, 1, 1},    // conv3_2 -> [12,8,256]
    {12, 8, 256, 512, 9, 1, 0},     // conv4_1
    {12, 8, 512, 512, 9, 1, 1},     // conv4_2 -> [6,4,512] == [12288]
    {1, 1, 12288, 4096, 1, 1, 0},   // fc1
    {1, 1, 4096, 4096, 1, 1, 0},    // fc2
    {1, 1, 4096, 128, 1, 0, 0},     // fc3 (no ReLU: fadtk/model_loader.py:102-103)
};
// bytes of fp16 activation per example produced by: conv1, conv2, conv3_1, conv3_2, conv4_1, conv4_2, fc1, fc2
const size_t kActElems[8] = {48 * 32 * 64, 24 * 16 * 128, 24 * 16 * 256, 12 * 8 * 256,
                             12 * 8 * 512, 6 * 4 * 512, 4096, 4096};

void build_frontend_tables(std::vector<double>& tw, std::vector<double>& hann,
                           std::vector<double>& melw, std::vector<int>& mstart, std::vector<int>& mcount) {
    const double PI = 3.14159265358979323846;
    tw.resize(512); hann.resize(fad::kWin);
    for (int k = 0; k < 256; ++k) {
        tw[2 * k] = std::cos(-2.0 * PI * k / 512.0);
        tw[2 * k + 1] = std::sin(-2.0 * PI * k / 512.0);
    }
    for (int n = 0; n < fad::kWin; ++n) hann[n] = 0.5 - 0.5 * std::cos(2.0 * PI / fad::kWin * n);
    // HTK mel filterbank, 64 bands over 125..7500 Hz on 257 bins of 0..8000 Hz; DC bin zeroed
    auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
    std::vector<double> bins(fad::kBins), edges(fad::kMel + 2);
    for (int i = 0; i < fad::kBins; ++i) bins[i] = mel(8000.0 * i / (fad::kBins - 1));
    const double lo = mel(125.0), hi = mel(7500.0);
    for (int i = 0; i < fad::kMel + 2; ++i) edges[i] = lo + (hi - lo) * i / (fad::kMel + 1);
    melw.assign(fad::kMel * fad::kMelMaxTaps, 0.0);
    mstart.assign(fad::kMel, 0); mcount.assign(fad::kMel, 0);
    for (int b = 0; b < fad::kMel; ++b) {
        const double l = edges[b], c = edges[b + 1], u = edges[b + 2];
        int first = -1, cnt = 0;
        for (int i = 1; i < fad::kBins; ++i) {            // bin 0 (DC) has zero weight
            const double w = std::fmax(0.0, std::fmin((bins[i] - l) / (c - l), (u - bins[i]) / (u - c)));
            if (w > 0.0) {
                if (first < 0) first = i;
                const int off = i - first;
                if (off < fad::kMelMaxTaps) { melw[b * fad::kMelMaxTaps + off] = w; cnt = off + 1; }
            }
        }
        mstart[b] = first < 0 ? 0 : first;
        mcount[b] = cnt;
    }
}

int ensure(void** ptr, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    CK(cudaMalloc(ptr, bytes));
    *cap = bytes;
    return 0;
}

size_t prof_begin(fad_handle* h, cudaStream_t st) {
    if (!h->prof_on) return 0;
    if (h->ev_used == h->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
    cudaEventRecord(h->ev_pool[h->ev_used], st);
    return h->ev_used++;
}
void prof_end(fad_handle* h, int cat, size_t e0, cudaStream_t st) {
    if (!h->prof_on) return;
    const size_t e1 = prof_begin(h, st);
    h->spans.push_back({cat, e0, e1});
}

}  // namespace

extern "C" {

int fad_version(void) { return 1; }

int fad_profile_enable(fad_handle* h, int on) {
    if (!h) return fail("null handle");
    h->prof_on = on != 0;
    return 0;
}

int fad_profile_collect(fad_handle* h, double* ms_out, long long* count_out, int reset) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    for (const auto& sp : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_pool[sp.e0], h->ev_pool[sp.e1]) == cudaSuccess) {
            h->prof_ms[sp.cat] += ms; h->prof_count[sp.cat]++;
        }
    }
    h->spans.clear(); h->ev_used = 0;
    for (int i = 0; i < FAD_PROF_CATEGORIES; ++i) {
        if (ms_out) ms_out[i] = h->prof_ms[i];
        if (count_out) count_out[i] = h->prof_count[i];
        if (reset) { h->prof_ms[i] = 0.0; h->prof_count[i] = 0; }
    }
    return 0;
}
const char* fad_last_error(void) { return g_err.c_str(); }

int fad_create(int device, int max_examples, fad_handle** out) {
    if (!out) return fail("null out pointer");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail("no CUDA device: fadtk_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail("bad device index");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail("fadtk_b200 kernels are built for sm_100a (Blackwell B200) only");
    fad_handle* h = new fad_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->max_examples = max_examples > 0 ? max_examples : 2048;

    std::vector<double> tw, hann, melw; std::vector<int> ms, mc;
    build_frontend_tables(tw, hann, melw, ms, mc);
    CK(cudaMalloc(&h->d_twiddle, tw.size() * 8)); CK(cudaMemcpy(h->d_twiddle, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_hann, hann.size() * 8)); CK(cudaMemcpy(h->d_hann, hann.data(), hann.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_melw, melw.size() * 8)); CK(cudaMemcpy(h->d_melw, melw.data(), melw.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_mel_start, ms.size() * 4)); CK(cudaMemcpy(h->d_mel_start, ms.data(), ms.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_mel_count, mc.size() * 4)); CK(cudaMemcpy(h->d_mel_count, mc.data(), mc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->fr_scal, 32 * sizeof(double)));
    CK(cudaFuncSetAttribute(fad::logmel_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fad::logmel_smem_bytes<double>()));
    CK(cudaFuncSetAttribute(fad::logmel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fad::logmel_smem_bytes<float>()));
    CK(cudaFuncSetAttribute(fad::stats_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)fad::kStSmemBytes));
    *out = h;
    return 0;
}

static void clap_free_state(void* p);
static void whisper_free_state(void* p);
static void encodec_free_state(void* p);
static void w2v_free_state(void* p);

int fad_destroy(fad_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    fad_comm_destroy(h);
    clap_free_state(h->clap_state);
    whisper_free_state(h->whisper_state);
    encodec_free_state(h->encodec_state);
    w2v_free_state(h->w2v_state);
    void* ptrs[] = {h->d_twiddle, h->d_hann, h->d_melw, h->d_mel_start, h->d_mel_count, h->conv1_w, h->conv1_b,
                    h->logmel, h->ws_tiles, h->ws_sums, h->gather_buf, h->fr_buf, h->fr_scal, h->frb_buf, h->rs_bank, h->rs_mono};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& kv : h->lo8) cudaFree(kv.second.w8);
    for (int i = 0; i < 5; ++i) { if (h->conv_w[i]) cudaFree(h->conv_w[i]); if (h->conv_b[i]) cudaFree(h->conv_b[i]); }
    for (int i = 0; i < 3; ++i) { if (h->fc_w[i]) cudaFree(h->fc_w[i]); if (h->fc_b[i]) cudaFree(h->fc_b[i]); }
    for (int i = 0; i < 9; ++i) if (h->act[i]) cudaFree(h->act[i]);
    for (int i = 0; i < 8; ++i) if (h->act8[i]) cudaFree(h->act8[i]);
    if (h->x8_scratch) cudaFree(h->x8_scratch);
    delete h;
    return 0;
}

long long fad_launch_count(fad_handle* h) { return h ? h->launches : 0; }

// ------------------------------------------------------------------------------ VGGish
int fad_vggish_load(fad_handle* h, const fad_vggish_weights* w) {
    if (!h || !w) return fail("null argument");
    CK(cudaSetDevice(h->device));
    auto up = [&](void** dst, const void* src, size_t bytes) -> int {
        if (!src) return fail("missing weight pointer");
        if (*dst) { lo8_forget(h, *dst); cudaFree(*dst); *dst = nullptr; }          // sizes depend on split_mask
        CK(cudaMalloc(dst, bytes));
        lo8_forget(h, *dst);
        CK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return 0;
    };
    if (up((void**)&h->conv1_w, w->conv1_w_host, 64 * 9 * 4)) return 1;
    if (up((void**)&h->conv1_b, w->conv1_b_host, 64 * 4)) return 1;
    for (int i = 0; i < 5; ++i) {
        const VggLayer& L = kVgg[i];
        const size_t mul = ((w->split_mask >> i) & 1) ? 2 : 1;
        if (up((void**)&h->conv_w[i], w->conv_w_host[i], mul * L.Cout * 9 * L.Cin * 2)) return 1;
        if (up((void**)&h->conv_b[i], w->conv_b_host[i], (size_t)L.Cout * 4)) return 1;
    }
    for (int i = 0; i < 3; ++i) {
        const VggLayer& L = kVgg[5 + i];
        const size_t mul = ((w->split_mask >> (5 + i)) & 1) ? 2 : 1;
        if (up((void**)&h->fc_w[i], w->fc_w_host[i], mul * L.Cout * L.Cin * 2)) return 1;
        if (up((void**)&h->fc_b[i], w->fc_b_host[i], (size_t)L.Cout * 4)) return 1;
    }
    const size_t B = (size_t)h->max_examples;
    if (!h->logmel) CK(cudaMalloc(&h->logmel, B * 96 * 64 * 4));
    for (int i = 0; i < 8; ++i)
        if (!h->act[i]) CK(cudaMalloc(&h->act[i], B * kActElems[i] * 2));
    // descriptors of the fixed pipeline (batch dimension = max_examples; tiles past the live
    // batch are never scheduled and rows past it are masked in the epilogue)
    for (int i = 0; i < 8; ++i) {
        const VggLayer& L = kVgg[i];
        if (make_geom(h->geom[i], L.H, L.W, L.Cin, L.Cout, L.taps, L.relu, L.pool, ((w->split_mask >> i) & 1) ? wlo_mode() : 0)) return 1;
        h->geom[i].pair = (h->geom[i].split_w && ((pair_mask() >> i) & 1)) ? 1 : 0;
        const void* wptr = i < 5 ? (const void*)h->conv_w[i] : (const void*)h->fc_w[i - 5];
        if (encode_layer_maps(h->geom[i], h->act[i], (long long)B, wptr, &h->map_x[i], &h->map_w[i])) return 1;
        if (h->geom[i].split_w == 2) {
            if (!h->act8[i]) CK(cudaMalloc(&h->act8[i], B * kActElems[i]));
            if (encode_x8_map(h->geom[i], h->act8[i], (long long)B, &h->map_x8[i])) return 1;
        }
    }
    h->vgg_loaded = true;
    return 0;
}

long long fad_vggish_num_examples(long long n_samples) {
    if (n_samples < fad::kWin) return 0;
    const long long t = 1 + (n_samples - fad::kWin) / fad::kHop;
    if (t < fad::kExFrames) return 0;
    return 1 + (t - fad::kExFrames) / fad::kExFrames;
}

long long fad_vggish_plan(const long long* clip_offsets_host, long long n_clips,
                          long long* ex_start_host, long long capacity, long long* rows_per_clip_host) {
    long long n = 0;
    for (long long c = 0; c < n_clips; ++c) {
        const long long len = clip_offsets_host[c + 1] - clip_offsets_host[c];
        const long long k = fad_vggish_num_examples(len);
        if (rows_per_clip_host) rows_per_clip_host[c] = k;
        for (long long e = 0; e < k; ++e, ++n)
            if (ex_start_host && n < capacity)
                ex_start_host[n] = clip_offsets_host[c] + e * (long long)(fad::kExFrames * fad::kHop);
    }
    return n;
}

static int launch_logmel(fad_handle* h, const int16_t* pcm, const long long* ex_start, long long n,
                         float* out, int use_double, cudaStream_t st) {
    fad::FrontendTables tab{h->d_twiddle, h->d_hann, h->d_melw, h->d_mel_start, h->d_mel_count};
    const long long frames = n * fad::kExFrames;
    long long blocks = (frames + fad::kFeWarps - 1) / fad::kFeWarps;
    const long long cap = (long long)h->num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return 0;
    if (use_double)
        fad::logmel_kernel<double><<<(int)blocks, fad::kFeWarps * 32, fad::logmel_smem_bytes<double>(), st>>>(
            pcm, ex_start, (int)n, tab, out);
    else
        fad::logmel_kernel<float><<<(int)blocks, fad::kFeWarps * 32, fad::logmel_smem_bytes<float>(), st>>>(
            pcm, ex_start, (int)n, tab, out);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

int fad_vggish_logmel(fad_handle* h, const int16_t* pcm, const long long* ex_start,
                      long long n_examples, float* logmel_out, int use_double, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    return launch_logmel(h, pcm, ex_start, n_examples, logmel_out, use_double, (cudaStream_t)stream);
}

int fad_vggish_forward(fad_handle* h, const int16_t* pcm, const long long* ex_start,
                       long long n_examples, void* emb_out_f16, void* stream) {
    if (!h) return fail("null handle");
    if (!h->vgg_loaded) return fail("fad_vggish_load has not been called");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    static const int fe_double = []() { const char* e = getenv("FADTK_FRONTEND_FP64"); return (e && e[0] == '1') ? 1 : 0; }();
    for (long long base = 0; base < n_examples; base += h->max_examples) {
        const int nb = (int)((n_examples - base) < h->max_examples ? (n_examples - base) : h->max_examples);
        size_t ev = prof_begin(h, st);
        if (launch_logmel(h, pcm, ex_start + base, nb, h->logmel, fe_double, st)) return 1;
        prof_end(h, FAD_PROF_LOGMEL, ev, st);
        ev = prof_begin(h, st);
        {
            static const bool simt = []() { const char* e = getenv("FADTK_CONV1"); return !(e && std::string(e) == "mma"); }();   // default: CUDA-core stencil
            if (simt) fad::conv1_kernel<<<dim3(6, nb), 256, 0, st>>>(h->logmel, h->conv1_w, h->conv1_b, h->act[0], h->act8[0]);
            else      fad::conv1_mma_kernel<<<nb, 256, 0, st>>>(h->logmel, h->conv1_w, h->conv1_b, h->act[0], h->act8[0]);
        }
        CK(cudaGetLastError());
        h->launches++;
        prof_end(h, FAD_PROF_CONV1, ev, st);
        for (int i = 0; i < 8; ++i) {
            const float* bias = i < 5 ? h->conv_b[i] : h->fc_b[i - 5];
            void* out = (i == 7) ? (void*)((__half*)emb_out_f16 + (size_t)base * 128) : (void*)h->act[i + 1];
            ev = prof_begin(h, st);
            if (run_layer(h, h->geom[i], h->map_x[i], h->map_w[i], (i < 5 ? (const void*)h->conv_w[i] : (const void*)h->fc_w[i - 5]), nb, bias, out, nullptr, st,
                          nullptr, 0, 0, 0, 0, h->geom[i].split_w == 2 ? &h->map_x8[i] : nullptr, i < 7 ? h->act8[i + 1] : nullptr)) return 1;
            prof_end(h, FAD_PROF_LAYER0 + i, ev, st);
        }
    }
    return 0;
}

// Stage entry (parity test): conv1 (3x3, 1 -> 64, pad 1) + bias + ReLU + 2x2 max-pool on fp32 log-mel examples.
int fad_vggish_conv1(fad_handle* h, const float* logmel, long long n_examples, void* out_f16, void* stream) {
    if (!h) return fail("null handle");
    if (!h->vgg_loaded) return fail("fad_vggish_load has not been called");
    if (n_examples <= 0) return 0;
    CK(cudaSetDevice(h->device));
    fad::conv1_kernel<<<dim3(6, (unsigned)n_examples), 256, 0, (cudaStream_t)stream>>>(
        logmel, h->conv1_w, h->conv1_b, reinterpret_cast<__half*>(out_f16), nullptr);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

int fad_umma_layer(fad_handle* h, const void* x_f16, int NB, int H, int W, int Cin,
                   const void* w_f16, const float* bias, int Cout, int taps, int relu, int pool,
                   int split_w, void* out_f16, float* out_f32_or_null, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    LayerGeom g;
    if (make_geom(g, H, W, Cin, Cout, taps, relu, pool, split_w)) return 1;
    g.pair = (g.split_w && pair_all()) ? 1 : 0;
    if (pool && out_f32_or_null) return fail("fp32 copy is only available for un-pooled layers");
    CUtensorMap mx, mw, mx8;
    if (encode_layer_maps(g, x_f16, NB, w_f16, &mx, &mw)) return 1;
    if (g.split_w == 2) {                    // stage test entry: make the E4M3 copy a producer would have written
        lo8_forget(h, w_f16);                // caller-owned weights: never trust a cached conversion
        const size_t count = (size_t)NB * H * W * Cin;
        if (ensure((void**)&h->x8_scratch, &h->x8_scratch_cap, count)) return 1;
        f16_to_e4m3_kernel<<<(unsigned)std::min<size_t>((count / 8 + 255) / 256 + 1, (size_t)h->num_sms * 16), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const __half*>(x_f16), count, h->x8_scratch);
        CK(cudaGetLastError());
        if (encode_x8_map(g, h->x8_scratch, NB, &mx8)) return 1;
    }
    return run_layer(h, g, mx, mw, w_f16, NB, bias, out_f16, out_f32_or_null, (cudaStream_t)stream,
                     nullptr, 0, 0, 0, 0, g.split_w == 2 ? &mx8 : nullptr, nullptr);
}

// -------------------------------------------------------------------------- statistics
size_t fad_stats_acc_len(int d) { return 1 + 2 * (size_t)d + (size_t)d * d; }

int fad_stats_accumulate(fad_handle* h, const void* emb_f16, long long n_rows, int d,
                         const void* shift_f16, double* acc, int tensor_core, void* stream) {
    if (!h) return fail("null handle");
    if (n_rows <= 0) return 0;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const __half* E = reinterpret_cast<const __half*>(emb_f16);
    const __half* shift = reinterpret_cast<const __half*>(shift_f16);
    // mode 0 (default): exact fp64 Gram on the FP64 tensor pipe (DMMA); 1: tcgen05 fp16 hi/lo (fp32 accumulation,
    // full-rank well-conditioned sets only); 2: fp64 CUDA-core kernel (verification).  FADTK_STATS overrides.
    static const int forced = [] {
        const char* e = getenv("FADTK_STATS");
        if (!e) return -1;
        const std::string v(e);
        return v == "dmma" ? 0 : v == "umma" ? 1 : v == "simt" ? 2 : -1;
    }();
    if (forced >= 0) tensor_core = forced;
    if (tensor_core == 2) {
        if (d % 64 != 0) return fail("d must be a multiple of 64");
        dim3 grid((unsigned)((n_rows + fad::kSimtRows - 1) / fad::kSimtRows), d / 64, d / 64);
        fad::stats_simt_kernel<<<grid, 256, 0, st>>>(E, n_rows, d, shift, acc);
        CK(cudaGetLastError());
        h->launches++;
        return 0;
    }
    if (tensor_core == 0) {
        if (d % 64 != 0) return fail("d must be a multiple of 64");
        fad::StatsDmmaParams p;
        p.n_rows = n_rows; p.d = d; p.n_tiles = d / fad::kSdTile;
        p.n_pairs = p.n_tiles * (p.n_tiles + 1) / 2;
        const long long stages = (n_rows + fad::kSdRows - 1) / fad::kSdRows;
        long long want = (4LL * h->num_sms + p.n_pairs - 1) / p.n_pairs;       // ~2 waves at 2 CTAs per SM
        if (want < 1) want = 1;
        long long per = (stages + want - 1) / want;                             // 16-row stages per split
        if (per < 4) per = 4;
        p.n_splits = (int)((stages + per - 1) / per);
        p.rows_per_split = per * fad::kSdRows;
        p.shift = shift;
        const size_t jobs = (size_t)p.n_pairs * p.n_splits;
        if (ensure((void**)&h->ws_tiles, &h->ws_tiles_cap, jobs * fad::kSdTile * fad::kSdTile * 8)) return 1;
        if (ensure((void**)&h->ws_sums, &h->ws_sums_cap, (size_t)p.n_tiles * p.n_splits * fad::kSdTile * 8)) return 1;
        p.ws_tiles = h->ws_tiles; p.ws_sums = h->ws_sums;
        size_t ev = prof_begin(h, st);
        fad::stats_dmma_kernel<__half><<<(unsigned)jobs, 256, 0, st>>>(E, p);
        CK(cudaGetLastError());
        prof_end(h, FAD_PROF_STATS, ev, st);
        ev = prof_begin(h, st);
        fad::stats_dmma_reduce_kernel<<<dim3(p.n_pairs, fad::kSdTile * fad::kSdTile / 256), 256, 0, st>>>(p, acc);
        CK(cudaGetLastError());
        prof_end(h, FAD_PROF_STATS_REDUCE, ev, st);
        h->launches += 2;
        return 0;
    }
    if (d % 128 != 0) return fail("d must be a multiple of 128 for the tensor-core statistics kernel");
    fad::StatsJobParams p;
    p.n_rows = n_rows; p.d = d; p.n_tiles = d / 128;
    p.n_pairs = p.n_tiles * (p.n_tiles + 1) / 2;
    const long long stages = (n_rows + fad::kStStageRows - 1) / fad::kStStageRows;
    long long want = (2LL * h->num_sms) / p.n_pairs;          // ~2 waves of jobs
    if (p.n_pairs == 1) want = h->num_sms;
    if (want < 1) want = 1;
    long long per = (stages + want - 1) / want;                // 32-row stages per split
    if (per < fad::kStStagesPerChunk) per = fad::kStStagesPerChunk;   // at least one 256-row chunk
    p.n_splits = (int)((stages + per - 1) / per);
    p.rows_per_split = per * fad::kStStageRows;
    p.shift = shift;
    const size_t jobs = (size_t)p.n_pairs * p.n_splits;
    if (ensure((void**)&h->ws_tiles, &h->ws_tiles_cap, jobs * 128 * 128 * 8)) return 1;
    if (ensure((void**)&h->ws_sums, &h->ws_sums_cap, (size_t)p.n_tiles * p.n_splits * 2 * 128 * 8)) return 1;
    p.ws_tiles = h->ws_tiles; p.ws_sums = h->ws_sums;
    CUtensorMap me;
    const uint64_t ed[2] = {(uint64_t)d, (uint64_t)n_rows};
    const uint64_t es[1] = {(uint64_t)d * 2};
    const uint32_t eb[2] = {64, (uint32_t)fad::kStStageRows};
    if (encode_f16_map(&me, E, 2, ed, es, eb)) return 1;
    size_t ev = prof_begin(h, st);
    fad::stats_umma_kernel<<<(unsigned)jobs, fad::kStThreads, fad::kStSmemBytes, st>>>(me, p);
    CK(cudaGetLastError());
    prof_end(h, FAD_PROF_STATS, ev, st);
    ev = prof_begin(h, st);
    fad::stats_reduce_kernel<<<p.n_pairs, 256, 0, st>>>(p, acc);
    CK(cudaGetLastError());
    prof_end(h, FAD_PROF_STATS_REDUCE, ev, st);
    h->launches += 2;
    return 0;
}

// ---- the reference's per-file statistics semantics for equal-length files, on the device (fadtk/utils.py:13-46) ----
int fad_file_means(fad_handle* h, const void* emb_f16, long long n_files, int rows_per_file, int d,
                   double* m64_out, double* m16_out, void* stream) {
    if (!h) return fail("null handle");
    if (n_files <= 0) return 0;
    if (rows_per_file <= 0 || d <= 0) return fail("bad shape");
    CK(cudaSetDevice(h->device));
    long long blocks = (n_files * d + 255) / 256;
    if (blocks > (long long)h->num_sms * 16) blocks = (long long)h->num_sms * 16;
    fad::file_means_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(emb_f16), n_files, rows_per_file, d, m64_out, m16_out);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

// exact Gram statistics of fp64 rows (no shift): acc[0] += n, acc[1..d] += sum x, outer += sum x x^T (DMMA, fixed order)
int fad_stats_accumulate_f64(fad_handle* h, const double* rows, long long n_rows, int d, double* acc, void* stream) {
    if (!h) return fail("null handle");
    if (n_rows <= 0) return 0;
    if (d % 64 != 0) return fail("d must be a multiple of 64");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    fad::StatsDmmaParams p;
    p.n_rows = n_rows; p.d = d; p.n_tiles = d / fad::kSdTile;
    p.n_pairs = p.n_tiles * (p.n_tiles + 1) / 2;
    const long long stages = (n_rows + fad::kSdRows - 1) / fad::kSdRows;
    long long want = (4LL * h->num_sms + p.n_pairs - 1) / p.n_pairs;
    if (want < 1) want = 1;
    long long per = (stages + want - 1) / want;
    if (per < 4) per = 4;
    p.n_splits = (int)((stages + per - 1) / per);
    p.rows_per_split = per * fad::kSdRows;
    p.shift = nullptr;
    const size_t jobs = (size_t)p.n_pairs * p.n_splits;
    if (ensure((void**)&h->ws_tiles, &h->ws_tiles_cap, jobs * fad::kSdTile * fad::kSdTile * 8)) return 1;
    if (ensure((void**)&h->ws_sums, &h->ws_sums_cap, (size_t)p.n_tiles * p.n_splits * fad::kSdTile * 8)) return 1;
    p.ws_tiles = h->ws_tiles; p.ws_sums = h->ws_sums;
    fad::stats_dmma_kernel<double><<<(unsigned)jobs, 256, 0, st>>>(rows, p);
    fad::stats_dmma_reduce_kernel<<<dim3(p.n_pairs, fad::kSdTile * fad::kSdTile / 256), 256, 0, st>>>(p, acc);
    CK(cudaGetLastError());
    h->launches += 2;
    return 0;
}

int fad_stats_finalize_mirrored(fad_handle* h, const double* acc, const double* acc_means64, const double* acc_means16,
                                const void* shift_f16, int rows_per_file, int d, double* mu_out, double* cov_out, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    static const int keep = [] { const char* e = getenv("FADTK_SINGLE_FRAME_FILES"); return (e && std::string(e) == "keep") ? 1 : 0; }();
    const size_t total = (size_t)d * d;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > (unsigned)h->num_sms * 8) blocks = h->num_sms * 8;
    fad::stats_finalize_mirrored_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        acc, acc_means64, acc_means16, reinterpret_cast<const __half*>(shift_f16), rows_per_file, d, keep, mu_out, cov_out);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

int fad_stats_accumulate_gather(fad_handle* h, const void* emb_f16, long long n_src_rows,
                                const long long* idx, long long n_idx, int d,
                                const void* shift_f16, double* acc, void* stream) {
    if (!h) return fail("null handle");
    (void)n_src_rows;
    if (n_idx <= 0) return 0;
    if (d % 8 != 0) return fail("d must be a multiple of 8");
    CK(cudaSetDevice(h->device));
    if (ensure((void**)&h->gather_buf, &h->gather_cap, (size_t)n_idx * d * 2)) return 1;
    const long long vecs = n_idx * (d / 8);
    long long blocks = (vecs + 255) / 256;
    if (blocks > (long long)h->num_sms * 16) blocks = (long long)h->num_sms * 16;
    fad::gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(emb_f16), idx, n_idx, d, h->gather_buf);
    CK(cudaGetLastError());
    h->launches++;
    return fad_stats_accumulate(h, h->gather_buf, n_idx, d, shift_f16, acc, 0, stream);   // exact fp64 path
}

int fad_stats_finalize(fad_handle* h, const double* acc, const void* shift_f16, int d,
                       double* mu_out, double* cov_out, void* stream) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    const size_t total = (size_t)d * d;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > (unsigned)h->num_sms * 8) blocks = h->num_sms * 8;
    fad::stats_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        acc, reinterpret_cast<const __half*>(shift_f16), d, mu_out, cov_out);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

// ------------------------------------------------------------------- fp64 tensor-pipe peak
// Roofline denominator of the DMMA kernels (exact Gram, Newton-Schulz): MEASURED_PEAKS.json only
// carries the bf16 GEMM and HBM copy rates, so the fp64 tensor-pipe rate is measured here - every warp
// of a full grid issues independent m8n8k4 DMMAs from registers, nothing else.
namespace {
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = -1.0 * threadIdx.x; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) fad::dmma_884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) sink[0] = s;                       // keeps the chain alive, never true
}
}  // namespace

// Stage entry (parity test): encoder self-attention of n_clips sequences of S positions, heads of 64 dims.
// qkv fp16 [n_clips * S][3 d] (q | k | v), out fp16 [n_clips * S][d]; legacy != 0 runs the mma.sync kernel instead.
extern "C" int fad_attention(fad_handle* h, const void* qkv_f16, long long n_clips, int S, int d, void* out_f16, int legacy, void* stream) {
    if (!h || !qkv_f16 || !out_f16) return fail("null argument");
    if (d % 64 != 0 || S <= 0 || n_clips <= 0) return fail("bad shape");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!legacy) return launch_attention_umma(h, reinterpret_cast<const __half*>(qkv_f16), n_clips, S, d, d / 64, reinterpret_cast<__half*>(out_f16), st);
    fad::whisper_flash_attention_kernel<<<dim3((S + 63) / 64, d / 64, (unsigned)n_clips), 128, 0, st>>>(
        reinterpret_cast<const __half*>(qkv_f16), S, d, reinterpret_cast<__half*>(out_f16));
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}

// time (ms) of `ksteps` hi/lo-split K steps per SM under issue pattern `mode` (umma_bench.cuh), every SM busy
extern "C" int fad_bench_umma_mode(fad_handle* h, int mode, int ksteps, double* ms_out_host) {
    if (!h || !ms_out_host) return fail("null argument");
    if (mode < 0 || mode > 5 || ksteps < 16) return fail("bad mode / ksteps");
    CK(cudaSetDevice(h->device));
    CK(cudaFuncSetAttribute(fad::umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fad::kUbSmem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fad::umma_bench_kernel<<<h->num_sms, fad::kUbThreads, fad::kUbSmem>>>(mode, ksteps / 8);     // warm-up
    CK(cudaEventRecord(e0));
    fad::umma_bench_kernel<<<h->num_sms, fad::kUbThreads, fad::kUbSmem>>>(mode, ksteps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_out_host = ms;
    h->launches += 2;
    return 0;
}

extern "C" int fad_bench_dmma_peak(fad_handle* h, int iters, double* tflops_out_host) {
    if (!h || !tflops_out_host) return fail("null argument");
    CK(cudaSetDevice(h->device));
    if (iters <= 0) iters = 20000;
    double* sink = h->fr_scal + 30;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = h->num_sms * 4;
    dmma_peak_kernel<<<blocks, 256>>>(iters / 10, sink);            // warm-up
    CK(cudaEventRecord(e0));
    dmma_peak_kernel<<<blocks, 256>>>(iters, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double flop = (double)blocks * 8 /*warps*/ * (double)iters * 8 /*DMMAs*/ * 512.0;
    *tflops_out_host = flop / (ms * 1e-3) / 1e12;
    h->launches += 2;
    return 0;
}

// ------------------------------------------------------------------- cross-GPU statistics merge
// The one exchange step of the path: all-reduce(sum) of the packed fp64 accumulator over
// NVLink.  NCCL is resolved at run time with dlopen (the process usually has torch's libnccl.so.2 mapped already;
// $FADTK_NCCL_LIB overrides) so the library keeps loading on hosts without NCCL and links nothing but the C++ runtime.
#include <dlfcn.h>
namespace {
struct NcclId { char bytes[128]; };
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char* env = getenv("FADTK_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        void* lib = nullptr;
        for (const char* n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { a.why = "libnccl.so.2 not found (set FADTK_NCCL_LIB)"; return a; }
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(lib, "ncclAllReduce"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
        if (!a.ok) a.why = "NCCL library lacks an expected symbol";
        return a;
    }();
    return api;
}
int nccl_fail(const char* what, int rc) {
    return fail(std::string(what) + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(rc) : "NCCL error"));
}
}  // namespace

extern "C" {
int fad_comm_unique_id(void* id_out_host) {
    if (!id_out_host) return fail("null argument");
    NcclApi& n = nccl_api();
    if (!n.ok) return fail(n.why);
    NcclId id;
    const int rc = n.GetUniqueId(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id_out_host, id.bytes, sizeof id.bytes);
    return 0;
}

int fad_comm_init(fad_handle* h, const void* id_host, int rank, int world) {
    if (!h || !id_host) return fail("null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank / world size");
    NcclApi& n = nccl_api();
    if (!n.ok) return fail(n.why);
    CK(cudaSetDevice(h->device));
    if (h->nccl_comm) { n.CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
    NcclId id;
    memcpy(id.bytes, id_host, sizeof id.bytes);
    const int rc = n.CommInitRank(&h->nccl_comm, world, id, rank);
    if (rc != 0) { h->nccl_comm = nullptr; return nccl_fail("ncclCommInitRank", rc); }
    return 0;
}

int fad_comm_destroy(fad_handle* h) {
    if (h && h->nccl_comm) { nccl_api().CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
    return 0;
}

int fad_allreduce_sum_f64(fad_handle* h, void* nccl_comm_or_null, double* buf, long long n_values, void* stream) {
    if (!h || !buf) return fail("null argument");
    void* comm = nccl_comm_or_null ? nccl_comm_or_null : h->nccl_comm;
    if (!comm) return fail("no communicator: call fad_comm_init or pass an ncclComm_t");
    NcclApi& n = nccl_api();
    if (!n.ok) return fail(n.why);
    CK(cudaSetDevice(h->device));
    const int rc = n.AllReduce(buf, buf, (size_t)n_values, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, (cudaStream_t)stream);
    if (rc != 0) return nccl_fail("ncclAllReduce", rc);
    return 0;
}

int fad_stats_allreduce(fad_handle* h, void* nccl_comm_or_null, double* acc, int d, void* stream) {
    if (d <= 0) return fail("bad dimension");
    return fad_allreduce_sum_f64(h, nccl_comm_or_null, acc, (long long)fad_stats_acc_len(d), stream);
}
}  // extern "C"

// ----------------------------------------------------------------------------- Frechet
namespace {
int launch_dgemm2(fad_handle* h, const fad::DgemmBatch& batch, int nprob, int d, cudaStream_t st) {
    dim3 grid((d + fad::kDgTileN - 1) / fad::kDgTileN, (d + fad::kDgTileM - 1) / fad::kDgTileM, nprob);
    fad::dgemm_kernel<<<grid, 256, 0, st>>>(batch, d);
    CK(cudaGetLastError());
    h->launches++;
    return 0;
}
int launch_dgemm(fad_handle* h, const double* A, const double* B, double* C, int d, double alpha,
                 double beta_diag, double* trace, cudaStream_t st) {
    fad::DgemmBatch batch = {};
    batch.p[0] = {A, B, C, alpha, beta_diag, trace};
    batch.p[1] = batch.p[0];
    return launch_dgemm2(h, batch, 1, d, st);
}

// Coupled Newton-Schulz: on return Y ~ sqrt(sym(A)/|A|_F + delta I), Z its inverse;
// scal[0..1] = |A|_F, tr A; trY / trZ = traces of the converged iterates.
int newton_schulz(fad_handle* h, const double* A, int d, int iters, double* Y, double* Z, double* W,
                  double* T, double* scal, double* trY, double* trZ, float* dev /*3 floats*/, cudaStream_t st) {
    const size_t total = (size_t)d * d;
    unsigned eb = (unsigned)((total + 255) / 256);
    if (eb > (unsigned)h->num_sms * 8) eb = h->num_sms * 8;
    static const float dev_init[3] = {0.0f, 0.0f, 1.0e30f};            // slot (k-1)%3 for k = 0 is slot 2
    CK(cudaMemcpyAsync(dev, dev_init, sizeof dev_init, cudaMemcpyHostToDevice, st));
    fad::norm_trace_kernel<<<1, 1024, 0, st>>>(A, d, scal);
    fad::ns_init_kernel<<<eb, 256, 0, st>>>(A, d, scal, Y, Z);
    CK(cudaGetLastError());
    h->launches += 2;
    // (Y, Z) <-> (T, T+total) ping-pong on a fixed host schedule; an even iteration count lands the
    // last scheduled update in (Y, Z).  If the device stops early (dev < tol) the two pairs differ by
    // one factor W with |W - I| < tol, i.e. by < 1e-12 relative - either is the converged iterate.
    if (iters & 1) ++iters;
    double* Yc = Y; double* Zc = Z; double* Yn = T; double* Zn = T + total;
    const float tol = 1e-12f;
    for (int it = 0; it < iters; ++it) {
        float* d_prev = dev + (it + 2) % 3;       // max |W - I| of iteration it-1
        float* d_cur = dev + it % 3;
        float* d_next = dev + (it + 1) % 3;
        fad::DgemmBatch wb = {};
        wb.p[0] = {Zc, Yc, W, -0.5, 1.5, nullptr};                               // W = 1.5 I - 0.5 Z Y
        wb.p[1] = wb.p[0];
        wb.dev_in = d_prev; wb.dev_out = d_cur; wb.dev_clear = nullptr; wb.tol = tol;
        if (launch_dgemm2(h, wb, 1, d, st)) return 1;
        fad::DgemmBatch yz = {};
        yz.p[0] = {Yc, W, Yn, 1.0, 0.0, nullptr};                                // Y <- Y W
        yz.p[1] = {W, Zc, Zn, 1.0, 0.0, nullptr};                                // Z <- W Z
        yz.dev_in = d_prev; yz.dev_out = nullptr; yz.dev_clear = d_next; yz.tol = tol;
        if (launch_dgemm2(h, yz, 2, d, st)) return 1;
        double* t = Yc; Yc = Yn; Yn = t;
        t = Zc; Zc = Zn; Zn = t;
    }
    fad::trace_kernel<<<1, 256, 0, st>>>(Y, d, trY);
    fad::trace_kernel<<<1, 256, 0, st>>>(Z, d, trZ);
    CK(cudaGetLastError());
    h->launches += 2;
    return 0;
}
}  // namespace

// S = C^(1/2) and tr C of a baseline covariance, computed once and reused for every eval set
// (FAD-inf steps, per-song scores).  sqrt_out: d*d doubles, scal_out: 2 doubles (|C|_F, tr C), device.
int fad_sqrt_psd(fad_handle* h, const double* cov, int d, int iters, double* sqrt_out, double* scal_out,
                 void* stream) {
    if (!h) return fail("null handle");
    if (d <= 0) return fail("bad dimension");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (iters <= 0) iters = 60;
    const size_t total = (size_t)d * d;
    if (ensure((void**)&h->fr_buf, &h->fr_cap, 8 * total * 8)) return 1;
    double* Y = h->fr_buf;  double* Z = Y + total;  double* W = Z + total;  double* T = W + total;
    double* scalA = h->fr_scal;  double* trS = scalA + 8;  double* trZs = scalA + 10;
    float* devf = reinterpret_cast<float*>(scalA + 16);
    unsigned eb = (unsigned)((total + 255) / 256);
    if (eb > (unsigned)h->num_sms * 8) eb = h->num_sms * 8;
    const size_t ev = prof_begin(h, st);
    if (newton_schulz(h, cov, d, iters, Y, Z, W, T, scalA, trS, trZs, devf, st)) return 1;
    fad::ns_unscale_kernel<<<eb, 256, 0, st>>>(Y, d, scalA, sqrt_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(scal_out, scalA, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    h->launches++;
    prof_end(h, FAD_PROF_FRECHET, ev, st);
    return 0;
}

// FAD against a baseline given by (mu1, S1 = C1^(1/2), scal1 = {|C1|_F, tr C1}).
int fad_frechet_presqrt(fad_handle* h, const double* mu1, const double* sqrt1, const double* scal1,
                        const double* mu2, const double* cov2, int d, int iters, double* out, void* stream) {
    if (!h) return fail("null handle");
    if (d <= 0) return fail("bad dimension");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (iters <= 0) iters = 60;
    const size_t total = (size_t)d * d;
    if (ensure((void**)&h->fr_buf, &h->fr_cap, 8 * total * 8)) return 1;
    double* Y = h->fr_buf;  double* Z = Y + total;  double* W = Z + total;  double* T = W + total;
    double* M = T + 2 * total;  double* P = h->fr_buf + 7 * total;   // slots: Y Z W T T M S(fad_frechet) P
    double* scalB = h->fr_scal + 2;  double* scalM = h->fr_scal + 4;
    double* trY = h->fr_scal + 6;    double* resid = h->fr_scal + 7;  double* trZ = h->fr_scal + 9;
    float* devf = reinterpret_cast<float*>(h->fr_scal + 16);
    unsigned eb = (unsigned)((total + 255) / 256);
    if (eb > (unsigned)h->num_sms * 8) eb = h->num_sms * 8;
    const size_t ev_fr = prof_begin(h, st);
    fad::norm_trace_kernel<<<1, 1024, 0, st>>>(cov2, d, scalB);
    CK(cudaGetLastError());
    h->launches++;
    if (launch_dgemm(h, sqrt1, cov2, P, d, 1.0, 0.0, nullptr, st)) return 1;     // M = S C2 S
    if (launch_dgemm(h, P, sqrt1, M, d, 1.0, 0.0, nullptr, st)) return 1;
    if (newton_schulz(h, M, d, iters, Y, Z, W, T, scalM, trY, trZ, devf, st)) return 1;
    CK(cudaMemsetAsync(resid, 0, sizeof(double), st));
    if (launch_dgemm(h, Y, Y, P, d, 1.0, 0.0, nullptr, st)) return 1;
    fad::resid_kernel<<<eb, 256, 0, st>>>(P, M, d, scalM, resid);
    fad::frechet_assemble_kernel<<<1, 256, 0, st>>>(mu1, mu2, d, scal1, scalB, scalM, trY, trZ, resid, iters, out);
    CK(cudaGetLastError());
    h->launches += 2;
    prof_end(h, FAD_PROF_FRECHET, ev_fr, st);
    return 0;
}

int fad_frechet(fad_handle* h, const double* mu1, const double* cov1, const double* mu2,
                const double* cov2, int d, int iters, double* out, void* stream) {
    if (!h) return fail("null handle");
    if (d <= 0) return fail("bad dimension");
    CK(cudaSetDevice(h->device));
    const size_t total = (size_t)d * d;
    if (ensure((void**)&h->fr_buf, &h->fr_cap, 8 * total * 8)) return 1;
    double* S = h->fr_buf + 6 * total;            // scratch slot not used by the chains
    double* scal1 = h->fr_scal + 12;
    if (fad_sqrt_psd(h, cov1, d, iters, S, scal1, stream)) return 1;
    return fad_frechet_presqrt(h, mu1, S, scal1, mu2, cov2, d, iters, out, stream);
}

// Ragged-batched FAD of n_items eval sets against one cached baseline (see frechet_batched.cuh).
int fad_frechet_batched(fad_handle* h, const double* mu1, const double* sqrt1, const double* scal1,
                        const void* emb_f16, const long long* offsets, long long n_items, int d, int iters,
                        double* out, void* stream) {
    if (!h) return fail("null handle");
    if (d <= 0 || n_items < 0) return fail("bad dimension");
    if (n_items == 0) return 0;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (iters <= 0) iters = 60;
    if (iters & 1) ++iters;
    const size_t total = (size_t)d * d;
    long long G = (long long)((size_t(1) << 31) / (64 * total));          // 8 matrices of d*d doubles per item in 2 GiB
    if (G < 1) G = 1;
    if (G > 32767) G = 32767;                                              // two families share gridDim.z
    if (G > n_items) G = n_items;
    const size_t per_item = 8 * total * 8 + (size_t)d * 8 + 4 * 8 + 3 * 4 + 4;
    if (ensure((void**)&h->frb_buf, &h->frb_cap, per_item * (size_t)G + 256)) return 1;
    double* cov = reinterpret_cast<double*>(h->frb_buf);
    double* P = cov + G * total;   double* M = P + G * total;
    double* Y = M + G * total;     double* Z = Y + G * total;   double* W = Z + G * total;
    double* Yn = W + G * total;    double* Zn = Yn + G * total;
    double* mu = Zn + G * total;
    double* scalC = mu + G * d;    double* scalM = scalC + 2 * G;
    float* flags = reinterpret_cast<float*>(scalM + 2 * G);
    int* ok = reinterpret_cast<int*>(flags + 3 * G);
    const int dt = (d + 31) / 32;
    auto gemm = [&](const fad::DgemmStrided& p, int families) -> int {
        dim3 grid((d + fad::kDgTileN - 1) / fad::kDgTileN, (d + fad::kDgTileM - 1) / fad::kDgTileM, (unsigned)(p.items * families));
        fad::dgemm_strided_kernel<<<grid, 256, 0, st>>>(p, d);
        CK(cudaGetLastError());
        h->launches++;
        return 0;
    };
    const float tol = 1e-12f;
    const size_t ev = prof_begin(h, st);
    for (long long g0 = 0; g0 < n_items; g0 += G) {
        const int g = (int)((n_items - g0) < G ? (n_items - g0) : G);
        if (d % 64 == 0) {                                   // fp64 tensor pipe (DMMA), upper tile triangle per item
            const int nt = d / 64;
            fad::song_stats_dmma_kernel<<<dim3(nt * (nt + 1) / 2, g), 256, 0, st>>>(reinterpret_cast<const __half*>(emb_f16), offsets + g0, d, mu, cov, ok);
        } else {
            fad::song_stats_kernel<<<dim3(dt, dt, g), 256, 0, st>>>(reinterpret_cast<const __half*>(emb_f16), offsets + g0, d, mu, cov, ok);
        }
        fad::norm_trace_batched_kernel<<<g, 256, 0, st>>>(cov, d, scalC);
        CK(cudaGetLastError());
        fad::DgemmStrided p = {};
        p.items = g; p.flags = nullptr; p.in_slot = p.out_slot = p.clear_slot = -1; p.tol = tol;
        p.f[0] = {sqrt1, cov, P, 0, (long long)total, (long long)total, 1.0, 0.0};          // P = S C_z
        if (gemm(p, 1)) return 1;
        p.f[0] = {P, sqrt1, M, (long long)total, 0, (long long)total, 1.0, 0.0};            // M = P S
        if (gemm(p, 1)) return 1;
        fad::norm_trace_batched_kernel<<<g, 256, 0, st>>>(M, d, scalM);
        unsigned eb = (unsigned)((total + 255) / 256);
        if (eb > 64) eb = 64;
        fad::ns_init_batched_kernel<<<dim3(eb, g), 256, 0, st>>>(M, d, scalM, Y, Z, flags);
        CK(cudaGetLastError());
        h->launches += 4;
        double* Yc = Y; double* Zc = Z; double* Yx = Yn; double* Zx = Zn;
        for (int it = 0; it < iters; ++it) {
            fad::DgemmStrided w = {};
            w.items = g; w.flags = flags; w.tol = tol;
            w.in_slot = (it + 2) % 3; w.out_slot = it % 3; w.clear_slot = -1;
            w.f[0] = {Zc, Yc, W, (long long)total, (long long)total, (long long)total, -0.5, 1.5};   // W = 1.5 I - 0.5 Z Y
            if (gemm(w, 1)) return 1;
            fad::DgemmStrided yz = {};
            yz.items = g; yz.flags = flags; yz.tol = tol;
            yz.in_slot = (it + 2) % 3; yz.out_slot = -1; yz.clear_slot = (it + 1) % 3;
            yz.f[0] = {Yc, W, Yx, (long long)total, (long long)total, (long long)total, 1.0, 0.0};   // Y <- Y W
            yz.f[1] = {W, Zc, Zx, (long long)total, (long long)total, (long long)total, 1.0, 0.0};   // Z <- W Z
            if (gemm(yz, 2)) return 1;
            double* t = Yc; Yc = Yx; Yx = t;
            t = Zc; Zc = Zx; Zx = t;
        }
        fad::frechet_assemble_batched_kernel<<<g, 256, 0, st>>>(mu1, mu, d, scal1, scalC, scalM, Y, Z, ok, offsets + g0,
                                                                iters, out + g0 * 8);
        CK(cudaGetLastError());
        h->launches++;
    }
    prof_end(h, FAD_PROF_FRECHET, ev, st);
    return 0;
}

}  // extern "C"

#include "resample_host.inc"
#include "clap_host.inc"

static void clap_free_state(void* p) { clap_free(reinterpret_cast<ClapState*>(p)); }

#include "whisper_host.inc"
#include "encodec_host.inc"
#include "wav2vec_host.inc"
