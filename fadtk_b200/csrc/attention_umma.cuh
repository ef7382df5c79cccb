// Encoder self-attention (head dim 64, no mask) on tcgen05: S = Q K^T and O = P V as UMMA tiles with the scores in TMEM.
//
// Replaces the mma.sync flash kernel (whisper.cuh) on the Whisper / wav2vec 2.0 / HuBERT / MERT encoder paths - the
// reference's `WhisperModel(...)` / `AutoModel(...)` forwards (fadtk/model_loader.py:656-672, 254-288, 525-596) spend
// their attention time in torch SDPA; here one CTA owns a 128-query block of one (clip, head):
//
//   warp 0      TMA producer: Q once, then K_j / V_j tiles (128 keys x 64 dims, 128-B swizzle) into a 2-stage ring.
//               The tensor map is 3-D [clips][S][3 d]: rows past S are zero-filled by the hardware.
//   warp 1      MMA issuer: S_j = Q K_j^T  (M 128, N 128, K 64; both operands K-major) into the TMEM score buffer,
//               and - one block behind - O_j = P_j V_j (M 128, N 64, K 128; P K-major from shared memory, V as an
//               MN-major B operand: the TMA tile [keys][dims] IS that layout) into one of two TMEM output buffers.
//   warps 2-9   softmax, TWO threads per query row (64 keys and 32 output dims each): two passes over the score row
//               straight out of TMEM (row maximum - exchanged between the two threads through shared memory -, then
//               exp2 / row sum), P written as fp16 into the 128-B-swizzled K-major tile the PV MMA reads; the partial
//               output of the previous block is folded into fp32 registers with the usual rescaling
//               O <- (O + O_{j-1}) * 2^(m_{j-1} - m_j), so the accumulator never has to be rescaled inside TMEM.
//
// Two CTAs per SM (7 tiles of 16 KiB of shared memory and 256 TMEM columns each): while one CTA's softmax warps work on
// S_j the other CTA's MMAs and loads run - the softmax side (exp2: 16 MUFU results per clock and SM, i.e. 1024 clocks per
// 128 x 128 block against 512 of MMA, plus the TMEM round trips) is the bound, so it is what must stay busy.
#pragma once
#include <cuda_bf16.h>
#include <type_traits>
#include "sm100.cuh"

namespace fad {

// s0 += lo half, s1 += hi half of a packed fp16 pair, in fp32 (add.rn.f32.f16 -> SASS FHADD: one instruction per value)
__device__ __forceinline__ void add_f16x2(float& s0, float& s1, uint32_t h2) {
    asm("{.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tadd.rn.f32.f16 %0, lo, %0;\n\tadd.rn.f32.f16 %1, hi, %1;}"
        : "+f"(s0), "+f"(s1) : "r"(h2));
}

constexpr int kAtThreads = 320;                                // TMA warp, MMA warp, 8 softmax warps
constexpr uint32_t kAtTile = 128 * 128;                       // bytes of one 128-row x 64-col fp16 tile: 16 KiB
constexpr uint32_t kAtSmem = kAtTile /*Q*/ + 2 * 2 * kAtTile /*K,V x 2 stages*/ + 2 * kAtTile /*P*/ + 256 /*barriers*/
                           + 128 * 4 /*row-maximum exchange*/;               // 115 456 B: x 2 CTAs + 2 x 1 KiB reserved <= 228 KiB

struct AttnParams {
    int S, d, heads;
    __half* out;             // [clips * S][d]
};

__global__ void __launch_bounds__(kAtThreads, 2)
attention_umma_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttnParams p)
{
    using namespace sm100;
    // no static shared memory in this kernel: the dynamic window starts at the CTA's 1 KiB-aligned base, which the
    // 128-B swizzle needs (checked below - the usual align-up slack would not leave room for two CTAs per SM)
    extern __shared__ __align__(1024) uint8_t at_raw[];
    uint8_t* smem = at_raw;
    if ((smem_u32(smem) & 1023u) != 0) asm volatile("trap;");
    uint8_t* q_s = smem;
    uint8_t* kv_s = smem + kAtTile;                            // stage st: K at kv_s + st * 2 tiles, V one tile further
    uint8_t* p_s = smem + 5 * kAtTile;                         // two 64-key blocks of 16 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kAtTile);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // 2
    uint64_t* kv_empty = bars + 3;      // 2
    uint64_t* s_full = bars + 5;        // 1
    uint64_t* s_empty = bars + 6;       // 1
    uint64_t* p_full = bars + 7;        // 1
    uint64_t* p_empty = bars + 8;       // 1
    uint64_t* o_full = bars + 9;        // 2
    uint64_t* o_empty = bars + 11;      // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* mx_s = reinterpret_cast<float*>(smem + 7 * kAtTile + 256);      // [row]: one slot per query row

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qb = blockIdx.x, h = blockIdx.y, clip = blockIdx.z;
    const int n_blocks = (p.S + 127) / 128;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&map_qkv);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(s_full, 1);  mbar_init(s_empty, 8);
        mbar_init(p_full, 8);  mbar_init(p_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
            mbar_init(&o_full[i], 1);  mbar_init(&o_empty[i], 8);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<256>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_s = tmem, tm_o0 = tmem + 128;      // O buffer b at columns 128 + 64 b (arithmetic, not a local array: a runtime index put it in local memory)

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(q_full, kAtTile);
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(q_s)), "l"(reinterpret_cast<uint64_t>(&map_qkv)), "r"(smem_u32(q_full)),
                           "r"(h * 64), "r"(qb * 128), "r"(clip) : "memory");
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&kv_full[st], 2 * kAtTile);
                uint8_t* kd = kv_s + st * 2 * kAtTile;
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(kd)), "l"(reinterpret_cast<uint64_t>(&map_qkv)), "r"(smem_u32(&kv_full[st])),
                               "r"(p.d + h * 64), "r"(j * 128), "r"(clip) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(kd + kAtTile)), "l"(reinterpret_cast<uint64_t>(&map_qkv)), "r"(smem_u32(&kv_full[st])),
                               "r"(2 * p.d + h * 64), "r"(j * 128), "r"(clip) : "memory");
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t kIdS = make_idesc(FMT_F16, 128, 128);
            constexpr uint32_t kIdO = make_idesc(FMT_F16, 128, 64, /*a MN-major*/ 0, /*b MN-major*/ 1);
            const uint64_t dq = kmajor_sw128_desc(smem_u32(q_s));
            auto issue_pv = [&](int j) {                       // O_j = P_j V_j
                const int b = j & 1, st = j & 1;
                mbar_wait(p_full, j & 1);
                mbar_wait(&o_empty[b], ((j >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                const uint32_t pa = smem_u32(p_s);
                const uint64_t dv = mnmajor_sw128_desc(smem_u32(kv_s + st * 2 * kAtTile + kAtTile), kAtTile, 1024);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {               // 16 keys per MMA: P advances 32 B inside its 64-key block, V two 8-key groups
                    const uint64_t dp = kmajor_sw128_desc(pa + (kk >> 2) * kAtTile) + 2 * (kk & 3);
                    umma_f16(tm_o0 + 64u * b, dp, dv + 128 * kk, kIdO, kk > 0);
                }
                umma_commit(&o_full[b]);
                umma_commit(p_empty);
                umma_commit(&kv_empty[st]);
            };
            mbar_wait(q_full, 0);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j & 1;
                mbar_wait(&kv_full[st], (j >> 1) & 1);
                mbar_wait(s_empty, (j & 1) ^ 1);
                tc_fence_after_sync();
                const uint64_t dk = kmajor_sw128_desc(smem_u32(kv_s + st * 2 * kAtTile));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tm_s, dq + 2 * k, dk + 2 * k, kIdS, k > 0);
                umma_commit(s_full);
                if (j > 0) issue_pv(j - 1);
            }
            issue_pv(n_blocks - 1);
        }
    } else {
        // ------------------------------------------------------------------ softmax + output accumulation
        // TWO threads per query row (warps w and w + 4 share a TMEM lane quarter): thread `half` owns keys
        // [64 half, 64 half + 64) of every block - i.e. one of the two 64-key blocks of the P tile - and output dims
        // [32 half, 32 half + 32).  The row maximum is exchanged through shared memory once per block (one named barrier
        // over the 256 softmax threads, slots double-buffered by block parity); the row sums stay per thread and are added
        // at the very end (both threads apply the same rescaling factors).
        const int quarter = warp & 3;                          // TMEM lanes this warp may read
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;                   // query row of the tile
        const uint32_t lane_base = uint32_t(quarter * 32) << 16;
        const float sc = 0.125f * 1.4426950408889634f;         // head_dim^-0.5 and log2(e): softmax in the exp2 domain
        float o[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = 0.f;
        float m = -3.0e38f, l = 0.f;
        const uint32_t sw = row & 7;
        uint8_t* p_blk = p_s + half * kAtTile + row * 128;      // this thread's row of its 64-key block of P
        const uint32_t s_cols = tm_s + lane_base + half * 64;
        auto ex2 = [](float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };   // one MUFU, x <= 0
        for (int j = 0; j < n_blocks; ++j) {
            const int valid = min(128, p.S - j * 128) - half * 64;   // keys of this thread's 64 that exist (may be <= 0)
            const bool full = valid >= 64;                      // every block but the last: no per-key masking
            mbar_wait(s_full, j & 1);
            tc_fence_after_sync();
            float raw = -3.0e38f;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t v[32];
                tmem_ld_32x32(s_cols + g * 32, v);
                tmem_ld_wait();
                if (full) {
                    float r0 = __uint_as_float(v[0]), r1 = __uint_as_float(v[1]), r2 = __uint_as_float(v[2]), r3 = __uint_as_float(v[3]);
#pragma unroll
                    for (int c = 4; c < 32; c += 4) {          // four independent chains
                        r0 = fmaxf(r0, __uint_as_float(v[c]));     r1 = fmaxf(r1, __uint_as_float(v[c + 1]));
                        r2 = fmaxf(r2, __uint_as_float(v[c + 2])); r3 = fmaxf(r3, __uint_as_float(v[c + 3]));
                    }
                    raw = fmaxf(raw, fmaxf(fmaxf(r0, r1), fmaxf(r2, r3)));
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (g * 32 + c < valid) raw = fmaxf(raw, __uint_as_float(v[c]));
                }
            }
            // Both threads of a row need the row maximum.  One fp32 slot per row (there is no room for more next to two
            // CTAs' tiles), two steps: the thread of the first 64 keys posts its maximum, the other folds its own in and
            // posts the result back.
            if (half == 0) mx_s[row] = raw;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 1) { raw = fmaxf(raw, mx_s[row]); mx_s[row] = raw; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0) raw = mx_s[row];
            const float mx = fmaxf(m, raw * sc);               // sc > 0: the maximum commutes with the scaling
            const float alpha = ex2(m - mx);
            m = mx;
            mbar_wait(p_empty, (j & 1) ^ 1);                   // PV_{j-1} has finished reading the P tile
            float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
            // FULL blocks (all but the last) carry no per-key masking: a separate instantiation, not a predicate per key
            auto exp_and_store = [&](auto full_c) {
                constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t v[32];
                    tmem_ld_32x32(s_cols + g * 32, v);
                    tmem_ld_wait();
                    uint32_t h2[16];
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        float e0 = ex2(fmaf(__uint_as_float(v[c]), sc, -m)), e1 = ex2(fmaf(__uint_as_float(v[c + 1]), sc, -m));
                        if (!FULL) {
                            if (g * 32 + c >= valid) e0 = 0.f;
                            if (g * 32 + c + 1 >= valid) e1 = 0.f;
                        }
                        const __half2 hh = __floats2half2_rn(e0, e1);
                        h2[c >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                        // the row sum is taken over the fp16 values the P V MMA reads (the weights of a row then sum to
                        // exactly 1): fp32 += fp16 is one instruction (FHADD) per value
                        if (c & 2) add_f16x2(rs2, rs3, h2[c >> 1]); else add_f16x2(rs0, rs1, h2[c >> 1]);
                    }
                    // keys g*32 .. g*32+31 of this thread's block = its 16-B chunks g * 4 .. +3 (XOR-swizzled by the row)
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint32_t chunk = uint32_t(g * 4 + q4) ^ sw;
                        *reinterpret_cast<uint4*>(p_blk + (chunk << 4)) = make_uint4(h2[4 * q4], h2[4 * q4 + 1], h2[4 * q4 + 2], h2[4 * q4 + 3]);
                    }
                }
            };
            if (full) exp_and_store(std::true_type{}); else exp_and_store(std::false_type{});
            l = l * alpha + ((rs0 + rs1) + (rs2 + rs3));
            tc_fence_before_sync();
            fence_proxy_async_smem();                          // P stores -> visible to the UMMA (async proxy)
            __syncwarp();
            if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(p_full); }
            if (j > 0) {                                       // fold in the previous block's P V (relative to the old maximum)
                const int bp = (j - 1) & 1;
                mbar_wait(&o_full[bp], ((j - 1) >> 1) & 1);
                tc_fence_after_sync();
                uint32_t v[32];
                tmem_ld_32x32(tm_o0 + 64u * bp + lane_base + half * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = (o[c] + __uint_as_float(v[c])) * alpha;
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[bp]);
            }
        }
        {
            const int bp = (n_blocks - 1) & 1;
            mbar_wait(&o_full[bp], ((n_blocks - 1) >> 1) & 1);  // the last P V has completed: the P tile is free
            tc_fence_after_sync();
            // total row sum = the two threads' partial sums, exchanged in fp32 through this row's (now idle) P rows
            *reinterpret_cast<float*>(p_blk) = l;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            l += *reinterpret_cast<const float*>(p_s + (half ^ 1) * kAtTile + row * 128);
            const float inv = 1.0f / l;
            const int q = qb * 128 + row;
            __half* dst = p.out + ((size_t)clip * p.S + q) * p.d + h * 64 + half * 32;
            uint32_t v[32];
            tmem_ld_32x32(tm_o0 + 64u * bp + lane_base + half * 32, v);
            tmem_ld_wait();
            if (q < p.S) {
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 hh = __floats2half2_rn((o[c + 2 * e] + __uint_as_float(v[c + 2 * e])) * inv,
                                                             (o[c + 2 * e + 1] + __uint_as_float(v[c + 2 * e + 1])) * inv);
                        w[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(dst + c) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc<256>(tmem);
}

}  // namespace fad
