// Register-tiled fp32 "rows x 1024-wide" GEMM micro-kernel shared by the training phases (same scheme as the
// inference kernel in forward.cu):  C[2*TM rows, L cols] += A[k][row] * B[k][col]  with
//   A  k-major in shared memory (stride MP floats per k),  2 row groups x 16 slots,
//   B  streamed by the producer warp through the TMA ring as chunks [KC][L] (k-major),
//   accumulators packed as f32x2 pairs over consecutive rows (fma.rn.f32x2 / FFMA2).
#pragma once
#include "common.cuh"

namespace mlb {

struct RingState {
    unsigned q;       // chunks consumed / produced so far in this CTA
    unsigned stage;   // ring slot of chunk q
    unsigned parity;  // fill parity of chunk q
};

__device__ __forceinline__ void ring_advance(RingState& rs) {
    rs.q++;
    if (++rs.stage == NSTAGE) rs.stage = 0, rs.parity ^= 1;
}

// Consume `nchunks` chunks of the stream.  The first chunk is waited for on entry; the mbarrier of chunk c+1 is probed
// (non-blocking) before the FFMAs of chunk c so its round trip hides under the math.  a_of_chunk(ch, stage) returns the
// k-major A pointer for k-step 0 of chunk `ch` (the resident activation tile, or a per-stage A buffer).
template <int TM, typename AOfChunk>
__device__ __forceinline__ void tile_gemm(unsigned long long (&acc2)[TM / 2][8], int nchunks, AOfChunk a_of_chunk,
                                          const float* ring, uint64_t* full, uint64_t* empty, RingState& rs, int n0, int g,
                                          int lane, int L, int* err_flag) {
    if (nchunks > 0 && !mbar_test_wait(&full[rs.stage], rs.parity)) mbar_wait(&full[rs.stage], rs.parity, err_flag);
    // several chunks per loop trip: less loop-carried register shuffling (forward.cu measured 1.357 -> 1.245 ms with it)
#pragma unroll(TM <= 14 ? 4 : 2)
    for (int ch = 0; ch < nchunks; ++ch) {
        unsigned nstage = rs.stage + 1, nparity = rs.parity;
        if (nstage == NSTAGE) nstage = 0, nparity ^= 1;
        const bool has_next = ch + 1 < nchunks;
        const bool next_ready = has_next ? mbar_test_wait(&full[nstage], nparity) : true;
        const float* b_ptr = ring + (size_t)rs.stage * KC * L + n0;
        const float* a_ptr = a_of_chunk(ch, rs.stage) + g * 16;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            unsigned long long a2[(TM + 1) / 2];
            const float* ap = a_ptr + kk * MP;
#pragma unroll
            for (int v = 0; v < TM / 4; ++v) {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(ap + v * 4);
                a2[v * 2 + 0] = t.x, a2[v * 2 + 1] = t.y;
            }
            if (TM % 4) a2[(TM / 4) * 2] = *reinterpret_cast<const unsigned long long*>(ap + (TM / 4) * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(b_ptr + kk * L);
            const float4 b1 = *reinterpret_cast<const float4*>(b_ptr + kk * L + 64);
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned long long bd = pack2(b[j], b[j]);
#pragma unroll
                for (int i = 0; i < TM / 2; ++i) acc2[i][j] = ffma2(a2[i], bd, acc2[i][j]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[rs.stage]);
        if (!next_ready) mbar_wait(&full[nstage], nparity, err_flag);
        rs.q++;
        rs.stage = nstage, rs.parity = nparity;
    }
}

template <int TM>
__device__ __forceinline__ void acc_zero(unsigned long long (&acc2)[TM / 2][8]) {
#pragma unroll
    for (int i = 0; i < TM / 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc2[i][j] = 0ull;
}

template <int TM>
__device__ __forceinline__ void acc_unpack(const unsigned long long (&acc2)[TM / 2][8], float (&acc)[TM][8]) {
#pragma unroll
    for (int i = 0; i < TM / 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) unpack2(acc2[i][j], acc[2 * i][j], acc[2 * i + 1][j]);
}

// column index of accumulator slot j for a thread whose first column is n0
__device__ __forceinline__ int col_of(int n0, int j) { return n0 + (j & 3) + (j >> 2) * 64; }

// local row r of a tile -> shared-memory row slot (2 groups of 16 slots, TM used per group)
__device__ __forceinline__ int slot_of_row(int r, int tm) { return (r / tm) * 16 + (r % tm); }

__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace mlb
