// Tensor-pipe microbenchmark: which tcgen05.mma issue pattern does a hi/lo-split K step cost the least in?
// One CTA per SM, operands resident in shared memory (no TMA, no epilogue): the time is what the tensor pipe and its
// shared-memory operand fetch need for one "K step" (128 rows x 128 channels x 64 K, weights as a hi/lo pair) issued as
//   mode 0  8 x kind::f16 M128 N128 K16            (hi and lo as two MMAs per K slice: the shipped WMODE 1)
//   mode 1  4 x kind::f16 M128 N256 K16            (hi | lo stacked along N: A is fetched once per K slice)
//   mode 2  4 x kind::f16 N128 K16 + 2 x kind::f8f6f4 N128 K32, alternating every K step   (WMODE 2)
//   mode 3  the same MMAs, kinds grouped over 4 K steps (16 x f16, then 8 x f8)
//   mode 4  4 x kind::f8f6f4 N128 K32              (pure E4M3 rate with 64-byte operand rows)
//   mode 5  4 x kind::f16 N128 K16                 (unsplit fp16 weights: the lower bound)
// Used by benchmarks/umma_modes.py; results in profiles/.
#pragma once
#include "sm100.cuh"

namespace fad {

constexpr int kUbThreads = 128;
constexpr uint32_t kUbSmem = 16384 /*A16*/ + 32768 /*B16: hi | lo*/ + 8192 /*A8*/ + 8192 /*B8*/ + 1024 /*align*/ + 64;

__global__ void __launch_bounds__(kUbThreads, 1)
umma_bench_kernel(int mode, int ksteps)
{
    using namespace sm100;
    extern __shared__ uint8_t ub_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ub_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    // operands: any finite pattern (small values, so the accumulators stay finite)
    for (int i = threadIdx.x; i < 65536 / 4; i += kUbThreads)
        reinterpret_cast<uint32_t*>(smem)[i] = (i < 49152 / 4) ? 0x1c001c00u /* two fp16 2^-8 */ : 0x20202020u /* four E4M3 2^-3 */;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc<512>(slot);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *slot;
    if (warp == 0 && elect_one()) {
        const uint32_t a16 = smem_u32(smem), b16 = a16 + 16384, a8 = a16 + 49152, b8 = a8 + 8192;
        const uint64_t da = kmajor_sw128_desc(a16), db = kmajor_sw128_desc(b16);
        const uint64_t da8 = kmajor_sw64_desc(a8), db8 = kmajor_sw64_desc(b8);
        constexpr uint32_t i128 = make_idesc(FMT_F16, 128, 128), i256 = make_idesc(FMT_F16, 128, 256);
        constexpr uint32_t i8 = make_idesc(FMT_E4M3, 128, 128);
        auto f16_hi = [&](uint32_t d) { for (int k = 0; k < 4; ++k) umma_f16(d, da + 2 * k, db + 2 * k, i128, 1); };
        auto f16_256 = [&](uint32_t d) { for (int k = 0; k < 4; ++k) umma_f16(d, da + 2 * k, db + 2 * k, i256, 1); };
        auto f8_lo = [&](uint32_t d) { for (int k = 0; k < 2; ++k) umma_f8(d, da8 + 2 * k, db8 + 2 * k, i8, 1); };
        // a commit every 8 K steps, waiting for the one before it: at most 16 K steps in flight
        uint32_t ph[2] = {0, 0};
        int pending[2] = {0, 0};
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t d = tmem + ((ks >> 3) & 1) * 256;
            if (mode == 0) { for (int k = 0; k < 4; ++k) { umma_f16(d, da + 2 * k, db + 2 * k, i128, 1); umma_f16(d, da + 2 * k, db + 2 * k + 1024, i128, 1); } }
            else if (mode == 1) f16_256(d);
            else if (mode == 2) { f16_hi(d); f8_lo(d + 128); }
            else if (mode == 3) { f16_hi(d); if ((ks & 3) == 3) for (int q = 0; q < 4; ++q) f8_lo(d + 128); }
            else if (mode == 4) { f8_lo(d); f8_lo(d + 128); }
            else f16_hi(d);
            if ((ks & 7) == 7) {
                const int b = (ks >> 3) & 1;
                umma_commit(&bar[b]);
                pending[b] = 1;
                const int o = b ^ 1;
                if (pending[o]) { mbar_wait(&bar[o], ph[o]); ph[o] ^= 1; pending[o] = 0; }
            }
        }
        umma_commit(&bar[0]);
        if (pending[0]) { mbar_wait(&bar[0], ph[0]); ph[0] ^= 1; }
        mbar_wait(&bar[0], ph[0]);
        if (pending[1]) mbar_wait(&bar[1], ph[1]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace fad
