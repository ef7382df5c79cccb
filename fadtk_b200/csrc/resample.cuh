// Mono mix-down + polyphase windowed-sinc resampling + PCM16 quantisation on the GPU.
//
// Replaces the conversion step of FrechetAudioDistance.load_audio (fadtk/fad.py:139-160):
//   x = mean over channels (:150);  torchaudio.transforms.Resample(fs, model_sr,
//   lowpass_filter_width=64, rolloff=0.9475937167399596, resampling_method="sinc_interp_kaiser",
//   beta=14.769656459379492) (:151-158);  save as PCM_S 16 (:160).
// torchaudio's algorithm (functional._get_sinc_resample_kernel / _apply_sinc_resample_kernel):
// with orig = fs/gcd, new = sr/gcd, the output sample j = m*new + p is a dot product of the p-th
// filter (2*width + orig taps, Kaiser-windowed sinc evaluated in float64, stored as float32)
// with the zero-padded input starting at m*orig - width.  The filter bank is built on the host
// (resample_host.inc); here one thread owns one output sample (four fp32 partial sums combined in
// fp64; the library accumulates in fp32 in an unspecified order, so results agree to ~1e-6, i.e. the
// same PCM16 sample except near rounding ties).
#pragma once
#include <stdint.h>

namespace fad {

// in_i16: interleaved [length][channels] PCM16 (scaled by 1/32768 like load_wav), or
// in_f32: planar [channels][length]; exactly one is non-null.  mono: fp32 [length].
__global__ void __launch_bounds__(256)
mono_mix_kernel(const int16_t* __restrict__ in_i16, const float* __restrict__ in_f32, int channels,
                long long length, float* __restrict__ mono)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < length; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        if (in_i16 != nullptr) {
            for (int c = 0; c < channels; ++c) s += (float)in_i16[i * channels + c] * (1.0f / 32768.0f);
        } else {
            for (int c = 0; c < channels; ++c) s += in_f32[(long long)c * length + i];
        }
        mono[i] = s / (float)channels;                     // torch.mean(x, 0)
    }
}

// bank: [new_][taps] float32, taps = 2*width + orig.  out: PCM16 [target_len] =
// clamp(round(y * 32768), -32768, 32767) (round half to even, like torch.round).
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ mono, long long length, const float* __restrict__ bank,
                int orig, int new_, int width, long long target_len, int16_t* __restrict__ out,
                float* __restrict__ out_f32 /* optional un-quantised copy */)
{
    const int taps = 2 * width + orig;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < target_len; j += (long long)gridDim.x * blockDim.x) {
        const long long m = j / new_;
        const int p = (int)(j - m * new_);
        const float* f = bank + (size_t)p * taps;
        const long long x0 = m * orig - width;              // index of tap 0 in the un-padded signal
        int k_lo = 0, k_hi = taps;
        if (x0 < 0) k_lo = (int)(-x0);
        if (x0 + taps > length) k_hi = (int)(length - x0);
        // four independent fp32 chains (no fp64 converts in the loop), combined in fp64
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* xs = mono + x0;
        int k = k_lo;
        for (; k + 4 <= k_hi; k += 4) {
            a0 = fmaf(__ldg(f + k), __ldg(xs + k), a0);
            a1 = fmaf(__ldg(f + k + 1), __ldg(xs + k + 1), a1);
            a2 = fmaf(__ldg(f + k + 2), __ldg(xs + k + 2), a2);
            a3 = fmaf(__ldg(f + k + 3), __ldg(xs + k + 3), a3);
        }
        for (; k < k_hi; ++k) a0 = fmaf(__ldg(f + k), __ldg(xs + k), a0);
        const float y = (float)(((double)a0 + (double)a1) + ((double)a2 + (double)a3));
        if (out_f32 != nullptr) out_f32[j] = y;
        float q = rintf(y * 32768.0f);
        q = fminf(fmaxf(q, -32768.0f), 32767.0f);
        out[j] = (int16_t)q;
    }
}

// same-rate case (torchaudio returns the waveform unchanged): mono -> PCM16 only
__global__ void __launch_bounds__(256)
quantize_pcm16_kernel(const float* __restrict__ mono, long long length, int16_t* __restrict__ out, float* __restrict__ out_f32)
{
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < length; j += (long long)gridDim.x * blockDim.x) {
        const float y = mono[j];
        if (out_f32 != nullptr) out_f32[j] = y;
        out[j] = (int16_t)fminf(fmaxf(rintf(y * 32768.0f), -32768.0f), 32767.0f);
    }
}

}  // namespace fad
