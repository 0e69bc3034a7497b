// Shared device helpers for the monoloco_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/monoloco_b200.h"

namespace mlb {

constexpr int MP = 32;        // detection rows per CTA tile as laid out in shared memory (2 groups x 16)
constexpr int KC = 8;         // k-steps per weight chunk (one TMA bulk copy = KC * L floats = 32 KB at L = 1024)
constexpr int NSTAGE = 3;     // weight-ring depth
constexpr int KIN_MAX = 72;   // largest network input (monstereo, 68) rounded up to the chunk depth
constexpr int OUT_LD = 16;    // raw-output staging row stride (output_size <= 16)
constexpr int MAX_THREADS = 384;   // 2 consumer warpgroups + 1 producer warpgroup (setmaxnreg 240 / 24)

// error codes written to the device error flag
enum { ERR_NONE = 0, ERR_MBAR_TIMEOUT = 1 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (PTX ISA: mbarrier.*) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps with an error flag instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            if (err_flag) *reinterpret_cast<volatile int*>(err_flag) = ERR_MBAR_TIMEOUT;  // plain store: the flag may live in mapped host memory
            __threadfence_system();
            __trap();
        }
    }
}

// Producer-side wait: sleeps between polls (the ring is normally full, so this warp mostly waits).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, int* err_flag) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(128);
        if (++spins > (1u << 22)) {
            if (err_flag) *reinterpret_cast<volatile int*>(err_flag) = ERR_MBAR_TIMEOUT;  // plain store: the flag may live in mapped host memory
            __threadfence_system();
            __trap();
        }
    }
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ----
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- Tensor Memory as a per-thread stash (tcgen05.alloc / st / ld; SASS: STTM / LDTM) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- packed fp32x2 math (sm_100: fma.rn.f32x2 -> SASS FFMA2), IEEE fma per half: bit-identical to two fmaf
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ---- counter-based RNG for the dropout keep masks (net.py:135-161 MC dropout; nn.Dropout in training):
// murmur3 fmix32 of (seed, site, row, col) -- regenerated identically wherever the mask is needed (forward, backward)
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t mix32(uint64_t x) {  // 64-bit counter -> 32 random bits (Laplace sampler)
    return fmix32((uint32_t)x ^ fmix32((uint32_t)(x >> 32) + 0x9E3779B9u));
}
// Dropout keep decision for element (site, row, col): r = fmix32(S ^ row*K ^ fmix32(col, site)), keep iff the top 24 bits
// as a fraction are >= p.  Split so that callers hoist the per-column / per-row / per-launch parts out of their loops.
__device__ __forceinline__ uint32_t drop_seed_mix(uint64_t seed) { return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B1u); }
__device__ __forceinline__ uint32_t drop_col_hash(uint32_t col, uint32_t site) {
    return fmix32(col * 0x85EBCA77u + site * 0xC2B2AE3Du + 0x27D4EB2Fu);
}
__device__ __forceinline__ uint32_t drop_row_mix(uint32_t seed_mix, uint32_t row) { return seed_mix ^ (row * 0x9E3779B1u); }
// (float)(r >> 8) * 2^-24 >= p  <=>  (r >> 8) >= ceil(p * 2^24)   (both sides exact)
__device__ __forceinline__ uint32_t drop_threshold(float p) { return (uint32_t)ceilf(p * 16777216.0f); }
__device__ __forceinline__ bool drop_keep(uint32_t row_mix, uint32_t col_hash, uint32_t thr) {
    return (fmix32(row_mix ^ col_hash) >> 8) >= thr;
}
// 4 mask bytes (little endian) -> 4 keep bits
__device__ __forceinline__ uint32_t bytes_to_bits(uint32_t a) {
    return ((a & 0xFFu) ? 1u : 0u) | ((a & 0xFF00u) ? 2u : 0u) | ((a & 0xFF0000u) ? 4u : 0u) | ((a & 0xFF000000u) ? 8u : 0u);
}
__device__ __forceinline__ bool keep_draw(uint64_t seed, uint32_t site, uint32_t row, uint32_t col, float p) {
    return drop_keep(drop_row_mix(drop_seed_mix(seed), row), drop_col_hash(col, site), drop_threshold(p));
}

// One output column of a narrow head, one row slot per lane:  bias + sum_k act[k * ld + slot] * w[k]   (K % 128 == 0,
// K <= 1024, w 16-byte aligned).  The K weights are fetched ONCE, coalesced (up to 8 float4 per lane, a single L2 round
// trip) and broadcast by shuffle; a uniform __ldg per 4 k paid one L2 latency every 32 k (measured 24 us per column at
// K = 1024 in the tile forward kernel).  Accumulation order: 4 interleaved chains over increasing k.
__device__ __forceinline__ float head_column(const float* __restrict__ w, float bias, int K, const float* act, int lane,
                                             int slot, int ld) {
    float4 wr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        wr[i] = (i * 128 < K) ? __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i * 128 < K) {
            const float* a = act + (size_t)(i * 128) * ld + slot;
#pragma unroll 4
            for (int s = 0; s < 32; ++s) {
                const float wx = __shfl_sync(0xffffffffu, wr[i].x, s), wy = __shfl_sync(0xffffffffu, wr[i].y, s);
                const float wz = __shfl_sync(0xffffffffu, wr[i].z, s), ww = __shfl_sync(0xffffffffu, wr[i].w, s);
                a0 = fmaf(a[(s * 4 + 0) * ld], wx, a0);
                a1 = fmaf(a[(s * 4 + 1) * ld], wy, a1);
                a2 = fmaf(a[(s * 4 + 2) * ld], wz, a2);
                a3 = fmaf(a[(s * 4 + 3) * ld], ww, a3);
            }
        }
    }
    return ((a0 + a1) + (a2 + a3)) + bias;
}

// All N <= 16 output columns of a narrow head, K split over the warps of the CTA (warp w: k in [128w, 128w + 128), K == 128 *
// nwarps), one row slot per lane.  Lane l fetches W[o][128w + 4l .. +3] for every column (N coalesced float4 loads, one
// L2 round trip) and the warp broadcasts them by shuffle.  Partial sums go to part[(w * 16 + o) * 32 + lane]; after the
// caller's barrier, head_ksplit_sum() adds the nwarps partials in a fixed order.
__device__ __forceinline__ void head_ksplit_partial(const float* __restrict__ W, int N, int K, const float* act, int warp, int lane,
                                                    int slot, int ld, float* part) {
    float4 wr[16];
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        acc[o] = 0.f;
        wr[o] = o < N ? __ldg(reinterpret_cast<const float4*>(W + (size_t)o * K + warp * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* a = act + (size_t)(warp * 128) * ld + slot;
#pragma unroll 2
    for (int s = 0; s < 32; ++s) {
        const float a0 = a[(s * 4 + 0) * ld], a1 = a[(s * 4 + 1) * ld], a2 = a[(s * 4 + 2) * ld], a3 = a[(s * 4 + 3) * ld];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            if (o < N) {
                float t = acc[o];
                t = fmaf(a0, __shfl_sync(0xffffffffu, wr[o].x, s), t);
                t = fmaf(a1, __shfl_sync(0xffffffffu, wr[o].y, s), t);
                t = fmaf(a2, __shfl_sync(0xffffffffu, wr[o].z, s), t);
                t = fmaf(a3, __shfl_sync(0xffffffffu, wr[o].w, s), t);
                acc[o] = t;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 16; ++o)
        if (o < N) part[(warp * 16 + o) * 32 + lane] = acc[o];
}
__device__ __forceinline__ float head_ksplit_sum(const float* part, int nwarps, int o, int lane, float bias) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += part[(w * 16 + o) * 32 + lane];
    return v + bias;
}

}  // namespace mlb
