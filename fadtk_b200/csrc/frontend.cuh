// VGGish front-end on the GPU: PCM16 -> log-mel examples, and the first convolution.
//
//  logmel_kernel   int16 PCM -> /32768 -> 25 ms periodic-Hann frames (hop 10 ms) -> |rFFT_512|
//                  -> 64 HTK-mel bands (125..7500 Hz) -> log(x + 0.01) -> fp32 [B, 96, 64].
//                  Replaces torchvggish's numpy front-end reached from
//                  fadtk/model_loader.py:107-108 (+ load_wav :63-70); SURVEY.md appendix A K1/K2.
//                  One warp per STFT frame: 512-point real FFT as a 256-point complex radix-2
//                  FFT in shared memory.  This stage is <3 % of the model FLOPs and its output
//                  feeds a log(), so it runs in full precision on the CUDA cores (template T =
//                  double matches the reference's float64 numpy to ~1e-13; float is ~4x cheaper).
//  conv1_kernel    3x3 conv 1->64 + bias + ReLU + 2x2 max-pool on fp32 input (K = 9 is too thin
//                  for the tensor pipe); writes NHWC fp16 [B, 48, 32, 64] for the tcgen05 layers.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace fad {

constexpr int kWin = 400, kHop = 160, kFft = 512, kBins = 257, kMel = 64, kExFrames = 96;
constexpr int kMelMaxTaps = 24;          // widest triangular filter spans < 24 FFT bins
constexpr int kFeWarps = 8;

// Host-built tables (double precision, converted on upload).
struct FrontendTables {
    const double* twiddle;   // [256][2]  exp(-2 pi i k / 512)
    const double* hann;      // [400]
    const double* mel_w;     // [64][kMelMaxTaps]
    const int* mel_start;    // [64] first FFT bin with non-zero weight
    const int* mel_count;    // [64]
};

template <typename T> struct Cx { T re, im; };

template <typename T>
__host__ __device__ constexpr size_t logmel_smem_bytes() {
    return sizeof(T) * (size_t)(kFeWarps * (2 * 256 + 260) + 2 * 256 + kWin + kMel * kMelMaxTaps)
         + sizeof(int) * 2 * kMel;
}

template <typename T>
__global__ void __launch_bounds__(kFeWarps * 32)
logmel_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ ex_start,
              int n_examples, FrontendTables tab, float* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char fe_smem[];
    T* sm = reinterpret_cast<T*>(fe_smem);
    Cx<T>* tw = reinterpret_cast<Cx<T>*>(sm);                 // 256 complex
    T* hann = sm + 512;                                        // 400
    T* melw = hann + kWin;                                     // 64*24
    T* wbuf = melw + kMel * kMelMaxTaps;                       // per-warp: 256 complex + 260 mags
    int* mstart = reinterpret_cast<int*>(wbuf + kFeWarps * (512 + 260));
    int* mcount = mstart + kMel;

    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        tw[i].re = (T)tab.twiddle[2 * i]; tw[i].im = (T)tab.twiddle[2 * i + 1];
    }
    for (int i = threadIdx.x; i < kWin; i += blockDim.x) hann[i] = (T)tab.hann[i];
    for (int i = threadIdx.x; i < kMel * kMelMaxTaps; i += blockDim.x) melw[i] = (T)tab.mel_w[i];
    for (int i = threadIdx.x; i < kMel; i += blockDim.x) { mstart[i] = tab.mel_start[i]; mcount[i] = tab.mel_count[i]; }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Cx<T>* z = reinterpret_cast<Cx<T>*>(wbuf + warp * (512 + 260));
    T* mag = wbuf + warp * (512 + 260) + 512;

    const long long total = (long long)n_examples * kExFrames;
    for (long long g = (long long)blockIdx.x * kFeWarps + warp; g < total;
         g += (long long)gridDim.x * kFeWarps) {
        const int e = (int)(g / kExFrames), f = (int)(g % kExFrames);
        const int16_t* src = pcm + ex_start[e] + (long long)f * kHop;

        // windowed samples packed as z[j] = x[2j] + i x[2j+1], stored bit-reversed (8 bits)
        for (int j = lane; j < 256; j += 32) {
            T re = 0, im = 0;
            if (j < kWin / 2) {
                re = (T)src[2 * j] * (T)(1.0 / 32768.0) * hann[2 * j];
                im = (T)src[2 * j + 1] * (T)(1.0 / 32768.0) * hann[2 * j + 1];
            }
            const int r = __brev((unsigned)j) >> 24;
            z[r].re = re; z[r].im = im;
        }
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
            const int half = 1 << s;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = lane + 32 * i;
                const int k = j & (half - 1);
                const int i0 = ((j >> s) << (s + 1)) + k, i1 = i0 + half;
                const Cx<T> w = tw[k * (256 >> s)];
                const Cx<T> u = z[i0], v = z[i1];
                const T vr = v.re * w.re - v.im * w.im, vi = v.re * w.im + v.im * w.re;
                z[i0].re = u.re + vr; z[i0].im = u.im + vi;
                z[i1].re = u.re - vr; z[i1].im = u.im - vi;
            }
            __syncwarp();
        }
        // split the packed transform into the 257 bins of the real FFT; keep magnitudes
        for (int k = lane; k <= 128; k += 32) {
            const Cx<T> a = z[k], b = z[(256 - k) & 255];
            const T er = (T)0.5 * (a.re + b.re), ei = (T)0.5 * (a.im - b.im);      // even part
            const T orr = (T)0.5 * (a.im + b.im), oi = (T)-0.5 * (a.re - b.re);    // odd part
            const Cx<T> w = tw[k];
            const T pr = orr * w.re - oi * w.im, pi = orr * w.im + oi * w.re;
            const T xr = er + pr, xi = ei + pi, yr = er - pr, yi = ei - pi;
            mag[k] = sqrt(xr * xr + xi * xi);
            mag[256 - k] = sqrt(yr * yr + yi * yi);
        }
        __syncwarp();
        float* dst = out + g * kMel;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = lane + 32 * h;
            const int st = mstart[b], cnt = mcount[b];
            T acc = 0;
            for (int i = 0; i < cnt; ++i) acc += mag[st + i] * melw[b * kMelMaxTaps + i];
            dst[b] = (float)log(acc + (T)0.01);
        }
        __syncwarp();
    }
}

// conv1: grid = (6 strips, B); block = 256 threads.  Lane = output-channel pair, warp loops
// over pooled pixels of an 8-row strip; the fp32 input strip (with halo) sits in smem.
__global__ void __launch_bounds__(256)
conv1_kernel(const float* __restrict__ logmel /*[B,96,64]*/, const float* __restrict__ w /*[64,9]*/,
             const float* __restrict__ bias, __half* __restrict__ out /*[B,48,32,64]*/)
{
    __shared__ float tile[18][68];
    const int e = blockIdx.y, strip = blockIdx.x;          // strip: pooled rows [8*strip, 8*strip+8)
    const int row0 = strip * 16 - 1;                       // first input row held (halo)
    const float* src = logmel + (size_t)e * 96 * 64;
    for (int i = threadIdx.x; i < 18 * 66; i += 256) {
        const int r = i / 66, c = i % 66;
        const int gr = row0 + r, gc = c - 1;
        tile[r][c] = (gr >= 0 && gr < 96 && gc >= 0 && gc < 64) ? src[gr * 64 + gc] : 0.0f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float w0[9], w1[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { w0[i] = w[(2 * lane) * 9 + i]; w1[i] = w[(2 * lane + 1) * 9 + i]; }
    const float b0 = bias[2 * lane], b1 = bias[2 * lane + 1];
    __syncthreads();

    const int py = warp;                                   // pooled row inside the strip
#pragma unroll 2
    for (int px = 0; px < 32; ++px) {
        float patch[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) patch[r][c] = tile[2 * py + r][2 * px + c];
        float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const float x = patch[dy + kh][dx + kw];
                        a0 = fmaf(w0[kh * 3 + kw], x, a0);
                        a1 = fmaf(w1[kh * 3 + kw], x, a1);
                    }
                m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1);
            }
        m0 = fmaxf(m0 + b0, 0.f); m1 = fmaxf(m1 + b1, 0.f);
        const size_t pix = ((size_t)e * 48 + strip * 8 + py) * 32 + px;
        *reinterpret_cast<__half2*>(out + pix * 64 + 2 * lane) = __floats2half2_rn(m0, m1);
    }
}

}  // namespace fad
