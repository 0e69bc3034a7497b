// monoloco_b200 -- latency kernel for one image's worth of detections (<= 32 rows): the WHOLE grid works on one row tile.
//
// What the reference does per image (net.py:92-124): preprocess_monoloco -> model(inputs) -> extract_outputs on the 1-30
// people of one frame.  At that size the forward pass is a stream of 33.8 MB of weights against a [<=32 x 1024] activation
// tile; a single CTA (forward.cu) or an 8-CTA cluster (forward_small.cu) can only pull the weights through 1 / 8 SMs'
// worth of L2 bandwidth.  Here every layer is split by OUTPUT COLUMNS over L/8 CTAs (128 at L = 1024):
//
//   CTA c owns columns [8c, 8c+8) of every layer.  Its weights are one contiguous slab  Wt[k][8]  per layer (32 KB at
//   K = 1024, re-packed at mlb_create), streamed by TMA through a 2-stage ring that runs ahead of the layer loop.
//   Per layer:  256 threads = (k-subset, row) compute partial sums of the [R x 8] block  ->  shared-memory reduction ->
//   folded-BN / ReLU / dropout / residual epilogue (one output per thread, residual kept in that thread's register) ->
//   the block goes to a global k-major exchange buffer xg[parity][L][R] -> grid barrier -> every CTA pulls the complete
//   next-layer input tile (L x R floats, L2-resident) back into shared memory with one TMA bulk copy.
//   Narrow heads, decode and the stores run on CTA 0 (the other CTAs exit after the last exchange).
//
// Cooperative launch (co-residency for the hand-rolled grid barrier); the barrier counter is monotonic across launches
// (the host passes the base value), so no memset precedes the kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "fwd_common.cuh"

namespace mlb {

constexpr int WC = 8;        // output columns per CTA
constexpr int WNT = 256;     // threads per CTA
constexpr int WNST = 2;      // weight-slab ring stages

struct WideExtra {
    const float* wslab;                // per GEMM op: [L/8 CTAs][Kpad][8]
    long long wslab_off[MLB_MAX_OPS];  // float offset of each op's slab block
    float* xg;                         // [2][L][32] exchange buffer (k-major tiles, double-buffered by layer parity)
    unsigned* bar;                     // monotonic grid-barrier counter
    unsigned bar_base;                 // counter value when this launch starts
};

__device__ __forceinline__ unsigned wide_ld_acquire(const unsigned* ptr) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}

template <int R>  // row slots of the tile: 16 or 32
__global__ void __launch_bounds__(WNT, 1) loco_forward_wide_kernel(const __grid_constant__ FwdParams p,
                                                                   const __grid_constant__ WideExtra ex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int S = WNT / R;  // k-subsets: thread (s, r) accumulates k = s, s + S, s + 2S, ...
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, L = p.L;

    float* act = reinterpret_cast<float*>(smem_raw);   // [L][R]  k-major input tile of the current layer
    float* ring = act + (size_t)L * R;                  // [WNST][L][WC] weight slabs
    float* part = ring + (size_t)WNST * L * WC;         // [S][WC][R] partial sums
    float* outs = part + WNT * WC;                      // [R][OUT_LD]
    float* cen = outs + R * OUT_LD;                     // [R][4]
    uint64_t* wfull = reinterpret_cast<uint64_t*>(cen + R * 4);  // [WNST]
    uint64_t* gfull = wfull + WNST;                     // exchange-tile arrival

    if (tid == 0) {
        for (int s = 0; s < WNST; ++s) mbar_init(&wfull[s], 1);
        mbar_init(gfull, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < R * OUT_LD; i += WNT) outs[i] = 0.f;
    __syncthreads();

    // ---- weight stream: GEMM op #g -> ring stage g % WNST (issued by thread 0, two layers ahead)
    int gemm_ops[MLB_MAX_OPS];
    int n_gemm = 0;
    for (int oi = 0; oi < p.n_ops; ++oi)
        if (p.ops[oi].type == MLB_OP_GEMM) gemm_ops[n_gemm++] = oi;
    auto issue_slab = [&](int g) {
        const mlb_op& op = p.ops[gemm_ops[g]];
        const uint32_t bytes = (uint32_t)(op.Kpad * WC * sizeof(float));
        const int st = g % WNST;
        mbar_expect_tx(&wfull[st], bytes);
        tma_bulk_g2s(ring + (size_t)st * L * WC, ex.wslab + ex.wslab_off[gemm_ops[g]] + (size_t)cta * op.Kpad * WC, bytes, &wfull[st]);
    };
    if (tid == 0)
        for (int g = 0; g < WNST && g < n_gemm; ++g) issue_slab(g);

    const int rows_here = p.n_rows;  // <= R, a single tile
    stage_input_tile(p, 0, rows_here, R, R, act, cen, tid, WNT, [] { __syncthreads(); });
    __syncthreads();
    if (cta == 0 && p.out_x != nullptr && p.input_kind != MLB_IN_X) {
        for (int idx = tid; idx < rows_here * p.in_size; idx += WNT) {
            const int r = idx / p.in_size, k = idx % p.in_size;
            p.out_x[(size_t)r * p.in_size + k] = act[k * R + r];
        }
    }

    // GEMM mapping: k-subset s, row r.  R = 32: s = warp, r = lane;  R = 16: s = 2 * warp + (lane >> 4), r = lane & 15
    const int gs = tid / R, gr = tid % R;
    // epilogue mapping (threads < WC * R): column ec of the CTA's 8, row er
    const int ec = tid / R, er = tid % R;
    const bool epi = tid < WC * R;
    const int gcol = cta * WC + ec;
    float res = 0.f;
    int site = 0, g = 0, par = 0;
    unsigned bar_target = ex.bar_base;

    for (int oi = 0; oi < p.n_ops; ++oi) {
        const mlb_op& op = p.ops[oi];
        if (op.type == MLB_OP_GEMM) {
            const int st = g % WNST;
            mbar_wait(&wfull[st], (g / WNST) & 1, p.err_flag);
            const float* w = ring + (size_t)st * L * WC;
            float acc[WC] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int k = gs; k < op.Kpad; k += S) {
                const float a = act[k * R + gr];
                const float4 w0 = *reinterpret_cast<const float4*>(w + k * WC);
                const float4 w1 = *reinterpret_cast<const float4*>(w + k * WC + 4);
                acc[0] = fmaf(a, w0.x, acc[0]), acc[1] = fmaf(a, w0.y, acc[1]);
                acc[2] = fmaf(a, w0.z, acc[2]), acc[3] = fmaf(a, w0.w, acc[3]);
                acc[4] = fmaf(a, w1.x, acc[4]), acc[5] = fmaf(a, w1.y, acc[5]);
                acc[6] = fmaf(a, w1.z, acc[6]), acc[7] = fmaf(a, w1.w, acc[7]);
            }
#pragma unroll
            for (int c = 0; c < WC; ++c) part[(gs * WC + c) * R + gr] = acc[c];
            __syncthreads();  // partials complete; nobody reads `act` / this ring stage any more
            if (tid == 0 && g + WNST < n_gemm) issue_slab(g + WNST);
            const bool last_gemm = g + 1 == n_gemm;
            if (epi) {
                float v = 0.f;
#pragma unroll
                for (int s = 0; s < S; ++s) v += part[(s * WC + ec) * R + er];
                v = fmaf(v, __ldg(p.blob + op.scale_off + gcol), __ldg(p.blob + op.shift_off + gcol));
                if (op.flags & MLB_F_RELU) v = fmaxf(v, 0.f);
                if ((op.flags & MLB_F_DROPOUT) && (p.flags & MLB_FWD_DROPOUT)) {
                    bool keep;
                    if (p.drop_mask != nullptr)
                        keep = er < p.n_rows ? p.drop_mask[((size_t)site * p.n_rows + er) * L + gcol] != 0 : true;
                    else
                        keep = keep_draw(p.drop_seed, site, er, gcol, p.p_drop);
                    v = keep ? v * (1.0f / (1.0f - p.p_drop)) : 0.f;
                }
                if (op.flags & MLB_F_ADD_RES) v += res;
                if (op.flags & MLB_F_SAVE_RES) res = v;
                ex.xg[((size_t)par * L + gcol) * R + er] = er < rows_here ? v : 0.f;
            }
            if (op.flags & MLB_F_DROPOUT) site++;
            // ---- grid barrier, then pull the complete tile back (TMA bulk copy, L2 -> shared)
            __syncthreads();
            bar_target += gridDim.x;
            if (tid == 0) {
                __threadfence();
                atomicAdd(ex.bar, 1u);
            }
            if (last_gemm && cta != 0) return;  // heads / decode / stores run on CTA 0 only
            if (tid == 0) {
                unsigned spins = 0;
                while ((int)(wide_ld_acquire(ex.bar) - bar_target) < 0) {
                    if (++spins > (1u << 24)) {
                        if (p.err_flag != nullptr) atomicExch(p.err_flag, 3);
                        __threadfence_system();
                        __trap();
                    }
                }
                asm volatile("fence.proxy.async;" ::: "memory");  // peers' generic-proxy stores -> this async-proxy read
                const uint32_t bytes = (uint32_t)((size_t)L * R * sizeof(float));
                mbar_expect_tx(gfull, bytes);
                for (uint32_t off = 0; off < bytes; off += 32768u)
                    tma_bulk_g2s(reinterpret_cast<unsigned char*>(act) + off,
                                 reinterpret_cast<const unsigned char*>(ex.xg + (size_t)par * L * R) + off,
                                 min(32768u, bytes - off), gfull);
            }
            mbar_wait(gfull, g & 1, p.err_flag);
            par ^= 1;
            g++;
        } else if (cta == 0) {
            // ---- narrow head on CTA 0: one warp per output column, lane = row slot
            for (int o = 7 - warp; o < op.N; o += 8) {
                const float y = head_column(p.blob + op.w_off + (size_t)o * op.K, __ldg(p.blob + op.shift_off + o), op.K, act, lane,
                                            lane & (R - 1), R);
                if (lane < R) outs[lane * OUT_LD + op.out_col + o] = y;
            }
        }
    }
    // ---- decode + store (CTA 0, one thread per row)
    __syncthreads();
    if (tid < rows_here) store_row(p, (size_t)tid, outs + tid * OUT_LD, cen + tid * 4);
}

// W^T [Kpad][L] -> per-CTA slabs [L/8][Kpad][8]
__global__ void wide_pack_kernel(const float* __restrict__ wt, float* __restrict__ slab, int kpad, int L) {
    const int n = kpad * L;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i / L, col = i % L;
        slab[((size_t)(col / WC) * kpad + k) * WC + col % WC] = wt[i];
    }
}

template <int R>
static size_t wide_smem(int L) {
    return ((size_t)L * R + (size_t)WNST * L * WC + (size_t)WNT * WC + (size_t)R * OUT_LD + (size_t)R * 4) * sizeof(float) +
           (WNST + 1) * sizeof(uint64_t);
}

}  // namespace mlb

using namespace mlb;

// total floats of the per-CTA slab copy and each op's offset in it
size_t mlb_wide_slab_floats(const mlb_op* ops, int n_ops, int L, long long* slab_off) {
    size_t off = 0;
    for (int i = 0; i < n_ops; ++i) {
        slab_off[i] = -1;
        if (ops[i].type != MLB_OP_GEMM) continue;
        slab_off[i] = (long long)off;
        off += (size_t)ops[i].Kpad * L;
    }
    return off;
}

cudaError_t mlb_wide_pack(const float* blob, const mlb_op* ops, int n_ops, int L, float* slab, const long long* slab_off,
                          cudaStream_t st) {
    for (int i = 0; i < n_ops; ++i) {
        if (ops[i].type != MLB_OP_GEMM) continue;
        wide_pack_kernel<<<128, 256, 0, st>>>(blob + ops[i].w_off, slab + slab_off[i], ops[i].Kpad, L);
    }
    return cudaGetLastError();
}

// can the whole grid (L/8 CTAs) be co-resident?  (cooperative launch requirement)
bool mlb_wide_supported(int L, int n_sms) {
    if (L % 128 != 0 || L / WC > n_sms) return false;
    int occ = 0;
    if (cudaFuncSetAttribute(loco_forward_wide_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem<32>(L)) != cudaSuccess)
        return false;
    if (cudaFuncSetAttribute(loco_forward_wide_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem<16>(L)) != cudaSuccess)
        return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, loco_forward_wide_kernel<32>, WNT, wide_smem<32>(L)) != cudaSuccess || occ < 1)
        return false;
    return true;
}

// number of grid barriers one launch performs (the host advances its copy of the counter by n * grid)
int mlb_wide_barriers(const mlb_op* ops, int n_ops) {
    int n = 0;
    for (int i = 0; i < n_ops; ++i) n += ops[i].type == MLB_OP_GEMM;
    return n;
}

cudaError_t mlb_wide_launch(const FwdParams& p, const float* wslab, const long long* wslab_off, float* xg, unsigned* bar,
                            unsigned bar_base, cudaStream_t st) {
    WideExtra ex;
    ex.wslab = wslab;
    for (int i = 0; i < MLB_MAX_OPS; ++i) ex.wslab_off[i] = i < p.n_ops ? wslab_off[i] : 0;
    ex.xg = xg, ex.bar = bar, ex.bar_base = bar_base;
    void* args[] = {(void*)&p, (void*)&ex};
    const int grid = p.L / WC;
    if (p.n_rows <= 16)
        return cudaLaunchCooperativeKernel((void*)loco_forward_wide_kernel<16>, dim3(grid), dim3(WNT), args, wide_smem<16>(p.L), st);
    return cudaLaunchCooperativeKernel((void*)loco_forward_wide_kernel<32>, dim3(grid), dim3(WNT), args, wide_smem<32>(p.L), st);
}
