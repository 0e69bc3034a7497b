// monoloco_b200 -- EXPERIMENTAL tensor-core forward (round-2 candidate; compile-checked only, never selected by mlb_forward).
//
// The fp32 network on the 5th-gen tensor cores without leaving the 1e-5 parity rule (DESIGN.md "What comes next",
// tools/tf32x3_study.py): every fp32 operand is split into two TF32 terms, a = a_hi + a_lo, and each layer product runs as
// three tcgen05.mma kind::tf32 (a_hi.w_hi into a main TMEM accumulator, a_lo.w_hi + a_hi.w_lo into a second one).
//
//   weights      re-packed once per model: per GEMM op  [4 column tiles][K/16 k blocks][hi | lo][256 x 16]  (canonical K-major
//                no-swizzle UMMA layout: core matrix 8 rows x 16 B, SBO 128 B, LBO rows x 16 B)
//   activations  between layers in the SAME layout,  [B/128 row tiles][1024/16][hi | lo][128 x 16], ping-pong in global / L2,
//                so a pipeline stage is two 1-D TMA bulk copies (16 KB of X planes + 32 KB of W planes), no tensor maps
//   kernel       cluster of 4 CTAs = one 128-row tile, CTA n owns output columns [256n, 256n + 256).  Per layer: warp 1 lane 0
//                streams the stages through a 4-slot ring, warp 0 lane 0 issues 2 k-steps x 3 MMAs (M 128, N 256, K 8) per
//                stage and releases it with tcgen05.commit; all 128 threads (thread = row) read main + cross back
//                (tcgen05.ld), apply folded BN / ReLU / residual, and write the result straight into the next layer's hi / lo
//                planes (+ an fp32 copy where a residual or a head needs it); barrier.cluster separates the layers.
//   heads        a CUDA-core kernel (warp per 4 rows) on the fp32 copies; decode through the existing mlb_decode.
// MC-dropout is not implemented on this path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "common.cuh"

namespace mlb {

constexpr int TCM = 128, TCN = 256, TCKB = 16, TCNST = 4;
constexpr uint32_t TC_A_PLANE = TCM * TCKB * 4, TC_W_PLANE = TCN * TCKB * 4;  // bytes
constexpr uint32_t TC_STAGE = 2 * TC_A_PLANE + 2 * TC_W_PLANE;                 // 48 KB
constexpr uint32_t TC_LBO_A = TCM * 16, TC_LBO_W = TCN * 16, TC_SBO = 128;
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);
constexpr int TC_MAX_LAYERS = 16;

struct TcLayer {
    const float* wplanes;  // [L/256][n_kb][hi|lo][256 x 16]
    const float* scale;    // folded BatchNorm scale [L]
    const float* shift;    // folded BatchNorm shift (+ bias) [L]
    int n_kb;              // K / 16
    int flags;             // MLB_F_RELU | MLB_F_SAVE_RES | MLB_F_ADD_RES
    int head_buf;          // index of the fp32 buffer that keeps this layer's output for a head, or -1
};

struct TcParams {
    TcLayer layer[TC_MAX_LAYERS];
    int n_layers, L, rows_pad;
    float* xplanes[2];   // [rows_pad/128][L/16][hi|lo][128 x 16]; [0] also holds the network input (n_kb of layer 0)
    float* res_f32;      // [rows_pad][L] stage input kept for the residual add
    float* head_f32[2];  // [rows_pad][L] outputs that feed a narrow head
    int* err_flag;
};

__device__ __forceinline__ float tc_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((TC_SBO >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// float offset of element (row r, k) inside one [tile_rows x 16] plane
__device__ __forceinline__ size_t tc_plane_off(int r, int k_in_block, int tile_rows) {
    return (size_t)(k_in_block >> 2) * tile_rows * 4 + (size_t)(r >> 3) * 32 + (size_t)(r & 7) * 4 + (k_in_block & 3);
}

// network input [B][in_size] fp32 -> hi / lo planes with K padded to n_kb * 16, rows padded to rows_pad (zeros)
__global__ void tc_pack_input_kernel(const float* __restrict__ x, float* __restrict__ planes, int B, int in_size, int n_kb, int rows_pad) {
    const int K = n_kb * TCKB;
    const size_t plane = (size_t)TCM * TCKB;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows_pad * K; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / K), k = (int)(i % K);
        const float v = (row < B && k < in_size) ? x[(size_t)row * in_size + k] : 0.f;
        const float h = tc_tf32(v), l = tc_tf32(v - h);
        float* blk = planes + ((size_t)(row / TCM) * n_kb + k / TCKB) * 2 * plane;
        const size_t off = tc_plane_off(row % TCM, k % TCKB, TCM);
        blk[off] = h;
        blk[plane + off] = l;
    }
}

// W^T [Kpad][L] (the packed blob's layout) -> W planes with K padded to n_kb * 16
__global__ void tc_pack_weights_kernel(const float* __restrict__ wt, float* __restrict__ planes, int Kpad, int L, int n_kb) {
    const int K = n_kb * TCKB;
    const size_t plane = (size_t)TCN * TCKB;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)L * K; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / L), n = (int)(i % L);  // consecutive threads -> consecutive n: coalesced reads of W^T
        const float v = k < Kpad ? wt[(size_t)k * L + n] : 0.f;
        const float h = tc_tf32(v), l = tc_tf32(v - h);
        float* blk = planes + ((size_t)(n / TCN) * n_kb + k / TCKB) * 2 * plane;
        const size_t off = tc_plane_off(n % TCN, k % TCKB, TCN);
        blk[off] = h;
        blk[plane + off] = l;
    }
}

__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(128, 1) tc_forward_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[TCNST], empty[TCNST], done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rb = blockIdx.x, nt = blockIdx.y, L = p.L;

    if (tid == 0) {
        for (int s = 0; s < TCNST; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
        mbar_init(&done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);  // [0,256) main accumulator, [256,512) cross terms
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const size_t grow = (size_t)rb * TCM + tid;  // this thread's row in the epilogue

    unsigned it_p = 0, it_m = 0;  // stages issued / consumed so far (producer lane, MMA lane)
    int par = 0;                  // activation plane buffer the current layer reads
    for (int g = 0; g < p.n_layers; ++g) {
        const TcLayer& ly = p.layer[g];
        if (warp == 1 && lane == 0) {
            // ---- producer: this row tile's X planes and this column tile's W planes, 48 KB per stage
            asm volatile("fence.proxy.async;" ::: "memory");  // the cluster peers' epilogue stores -> this thread's TMA reads
            const unsigned char* xsrc = reinterpret_cast<const unsigned char*>(p.xplanes[par]) + (size_t)rb * ly.n_kb * 2 * TC_A_PLANE;
            const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(ly.wplanes) + (size_t)nt * ly.n_kb * 2 * TC_W_PLANE;
            for (int kb = 0; kb < ly.n_kb; ++kb, ++it_p) {
                const unsigned s = it_p % TCNST;
                if (it_p >= TCNST) mbar_wait(&empty[s], ((it_p / TCNST) - 1) & 1, p.err_flag);
                unsigned char* st = smem_raw + (size_t)s * TC_STAGE;
                mbar_expect_tx(&full[s], TC_STAGE);
                tma_bulk_g2s(st, xsrc + (size_t)kb * 2 * TC_A_PLANE, 2 * TC_A_PLANE, &full[s]);
                tma_bulk_g2s(st + 2 * TC_A_PLANE, wsrc + (size_t)kb * 2 * TC_W_PLANE, 2 * TC_W_PLANE, &full[s]);
            }
        } else if (warp == 0 && lane == 0) {
            // ---- MMA issuer
            tmem_fence_after();
            uint32_t main_acc = 0, cross_acc = 0;
            for (int kb = 0; kb < ly.n_kb; ++kb, ++it_m) {
                const unsigned s = it_m % TCNST;
                mbar_wait(&full[s], (it_m / TCNST) & 1, p.err_flag);
                tmem_fence_after();
                const uint32_t a_hi = smem_u32(smem_raw + (size_t)s * TC_STAGE), a_lo = a_hi + TC_A_PLANE;
                const uint32_t w_hi = a_hi + 2 * TC_A_PLANE, w_lo = w_hi + TC_W_PLANE;
#pragma unroll
                for (int j = 0; j < TCKB / 8; ++j) {
                    const uint64_t ah = tc_desc(a_hi + 2 * j * TC_LBO_A, TC_LBO_A), al = tc_desc(a_lo + 2 * j * TC_LBO_A, TC_LBO_A);
                    const uint64_t wh = tc_desc(w_hi + 2 * j * TC_LBO_W, TC_LBO_W), wl = tc_desc(w_lo + 2 * j * TC_LBO_W, TC_LBO_W);
                    tc_mma(tmem + TCN, al, wh, cross_acc), cross_acc = 1;
                    tc_mma(tmem + TCN, ah, wl, 1u);
                    tc_mma(tmem, ah, wh, main_acc), main_acc = 1;
                }
                tc_commit(&empty[s]);
            }
            tc_commit(&done);
        }
        __syncwarp();
        mbar_wait_backoff(&done, (uint32_t)(g & 1), p.err_flag);
        tmem_fence_after();

        // ---- epilogue: thread = row; columns [256 nt, 256 nt + 256) in steps of 8 (= two 16-byte chunks of a k block)
        float* nxt = p.xplanes[par ^ 1];
        const bool relu = (ly.flags & MLB_F_RELU) != 0, add_res = (ly.flags & MLB_F_ADD_RES) != 0, save_res = (ly.flags & MLB_F_SAVE_RES) != 0;
        const size_t plane = (size_t)TCM * TCKB;
        for (int c0 = 0; c0 < TCN; c0 += 8) {
            float m[8], c[8];
            tmem_ld8(lane_base + (uint32_t)c0, m);
            tmem_ld8(lane_base + (uint32_t)(TCN + c0), c);
            const int col = nt * TCN + c0;
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(ly.scale + col)), s1 = __ldg(reinterpret_cast<const float4*>(ly.scale + col + 4));
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(ly.shift + col)), t1 = __ldg(reinterpret_cast<const float4*>(ly.shift + col + 4));
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = fmaf(m[j] + c[j], sc[j], sh[j]);
                if (relu) v[j] = fmaxf(v[j], 0.f);
            }
            if (add_res) {
                const float4 r0 = *reinterpret_cast<const float4*>(p.res_f32 + grow * L + col);
                const float4 r1 = *reinterpret_cast<const float4*>(p.res_f32 + grow * L + col + 4);
                v[0] += r0.x, v[1] += r0.y, v[2] += r0.z, v[3] += r0.w, v[4] += r1.x, v[5] += r1.y, v[6] += r1.z, v[7] += r1.w;
            }
            if (save_res) {
                *reinterpret_cast<float4*>(p.res_f32 + grow * L + col) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(p.res_f32 + grow * L + col + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            if (ly.head_buf >= 0) {
                float* hb = p.head_f32[ly.head_buf] + grow * L + col;
                *reinterpret_cast<float4*>(hb) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(hb + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            // next layer's A operand: output column `col` is its k index
            float* blk = nxt + ((size_t)rb * (L / TCKB) + col / TCKB) * 2 * plane;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 h = make_float4(tc_tf32(v[4 * q]), tc_tf32(v[4 * q + 1]), tc_tf32(v[4 * q + 2]), tc_tf32(v[4 * q + 3]));
                const float4 l = make_float4(tc_tf32(v[4 * q] - h.x), tc_tf32(v[4 * q + 1] - h.y), tc_tf32(v[4 * q + 2] - h.z),
                                             tc_tf32(v[4 * q + 3] - h.w));
                const size_t off = tc_plane_off(tid, (col % TCKB) + 4 * q, TCM);
                *reinterpret_cast<float4*>(blk + off) = h;
                *reinterpret_cast<float4*>(blk + plane + off) = l;
            }
        }
        tmem_fence_before();
        tc_cluster_sync();  // all four column tiles of this row tile are written; TMEM reads are complete
        tmem_fence_after();
        par ^= 1;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// narrow head on an fp32 activation buffer: warp per 4 rows, lanes split K, shuffle reduction
__global__ void tc_heads_kernel(const float* __restrict__ act, const float* __restrict__ W, const float* __restrict__ bias, int N, int K,
                                int B, float* __restrict__ out_raw, int out_size, int out_col) {
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int row0 = warp * 4;
    if (row0 >= B) return;
    for (int o = 0; o < N; ++o) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane * 4; k < K; k += 128) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)o * K + k));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (row0 + r < B) {
                    const float4 a = *reinterpret_cast<const float4*>(act + (size_t)(row0 + r) * K + k);
                    acc[r] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[r]))));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float s = acc[r];
            for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
            if (lane == 0 && row0 + r < B) out_raw[(size_t)(row0 + r) * out_size + out_col + o] = s + __ldg(bias + o);
        }
    }
}

}  // namespace mlb

// ================================================================================================ host side
using namespace mlb;

struct mlb_tc {
    int device, L, in_size, out_size, n_ops, max_rows_pad;
    mlb_op ops[MLB_MAX_OPS];
    float* blob_dev;
    float* wplanes[MLB_MAX_OPS];
    int n_kb[MLB_MAX_OPS];
    float *xplanes[2], *res_f32, *head_f32[2];
    int* err_flag;
};

extern thread_local std::string g_mlb_err;
void mlb_count_launch();
static int tc_fail(const std::string& m) {
    g_mlb_err = m;
    return -1;
}
#define TCC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return tc_fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" int mlb_tc_create(const mlb_model_desc* desc, const mlb_op* ops, const float* packed_host, size_t n_floats, int device,
                             int max_rows, mlb_tc_handle* out) {
    if (!desc || !ops || !packed_host || !out || max_rows < 1) return tc_fail("mlb_tc_create: bad argument");
    if (desc->linear_size != 1024) return tc_fail("mlb_tc_create: the tensor-core path is written for linear_size == 1024");
    if (desc->n_ops < 1 || desc->n_ops > MLB_MAX_OPS) return tc_fail("mlb_tc_create: n_ops out of range");
    TCC(cudaSetDevice(device));
    mlb_tc* t = new mlb_tc();
    t->device = device, t->L = desc->linear_size, t->in_size = desc->input_size, t->out_size = desc->output_size, t->n_ops = desc->n_ops;
    memcpy(t->ops, ops, sizeof(mlb_op) * desc->n_ops);
    t->max_rows_pad = ((max_rows + TCM - 1) / TCM) * TCM;
    TCC(cudaMalloc(&t->blob_dev, n_floats * sizeof(float)));
    TCC(cudaMemcpy(t->blob_dev, packed_host, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    int n_gemm = 0;
    for (int i = 0; i < desc->n_ops; ++i) {
        if (ops[i].type != MLB_OP_GEMM) continue;
        if (++n_gemm > TC_MAX_LAYERS) return tc_fail("mlb_tc_create: too many layers");
        t->n_kb[i] = (ops[i].Kpad + TCKB - 1) / TCKB;
        const size_t fl = (size_t)2 * t->n_kb[i] * TCKB * t->L;
        TCC(cudaMalloc(&t->wplanes[i], fl * sizeof(float)));
        tc_pack_weights_kernel<<<296, 256>>>(t->blob_dev + ops[i].w_off, t->wplanes[i], ops[i].Kpad, t->L, t->n_kb[i]);
    }
    const size_t act = (size_t)t->max_rows_pad * t->L;
    for (int b = 0; b < 2; ++b) {
        TCC(cudaMalloc(&t->xplanes[b], 2 * act * sizeof(float)));
        TCC(cudaMemset(t->xplanes[b], 0, 2 * act * sizeof(float)));
        TCC(cudaMalloc(&t->head_f32[b], act * sizeof(float)));
    }
    TCC(cudaMalloc(&t->res_f32, act * sizeof(float)));
    TCC(cudaMalloc(&t->err_flag, sizeof(int)));
    TCC(cudaMemset(t->err_flag, 0, sizeof(int)));
    TCC(cudaGetLastError());
    TCC(cudaDeviceSynchronize());
    *out = t;
    return 0;
}

extern "C" void mlb_tc_destroy(mlb_tc_handle t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaFree(t->blob_dev);
    for (int i = 0; i < t->n_ops; ++i) cudaFree(t->wplanes[i]);
    cudaFree(t->xplanes[0]), cudaFree(t->xplanes[1]), cudaFree(t->head_f32[0]), cudaFree(t->head_f32[1]);
    cudaFree(t->res_f32), cudaFree(t->err_flag);
    delete t;
}

// x_dev: pre-processed inputs [B, input_size] (MLB_IN_X); out_raw_dev: [B, output_size] raw network outputs
extern "C" int mlb_tc_forward(mlb_tc_handle t, const float* x_dev, int B, float* out_raw_dev, void* stream) {
    if (!t || !x_dev || !out_raw_dev || B < 1) return tc_fail("mlb_tc_forward: bad argument");
    const int rows_pad = ((B + TCM - 1) / TCM) * TCM;
    if (rows_pad > t->max_rows_pad) return tc_fail("mlb_tc_forward: more rows than mlb_tc_create reserved");
    TCC(cudaSetDevice(t->device));
    cudaStream_t st = (cudaStream_t)stream;
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.L = t->L, p.rows_pad = rows_pad, p.err_flag = t->err_flag;
    p.xplanes[0] = t->xplanes[0], p.xplanes[1] = t->xplanes[1], p.res_f32 = t->res_f32;
    p.head_f32[0] = t->head_f32[0], p.head_f32[1] = t->head_f32[1];
    int n_heads = 0, first = -1;
    int head_of_layer[MLB_MAX_OPS];  // GEMM op index whose output each head op reads
    for (int i = 0, last_gemm = -1; i < t->n_ops; ++i) {
        head_of_layer[i] = -1;
        if (t->ops[i].type == MLB_OP_GEMM) {
            last_gemm = i;
            if (first < 0) first = i;
        } else {
            if (last_gemm < 0) return tc_fail("mlb_tc_forward: a head before any layer");
            head_of_layer[i] = last_gemm;
        }
    }
    int buf_of_gemm[MLB_MAX_OPS];
    for (int i = 0; i < t->n_ops; ++i) buf_of_gemm[i] = -1;
    for (int i = 0; i < t->n_ops; ++i)
        if (head_of_layer[i] >= 0 && buf_of_gemm[head_of_layer[i]] < 0) {
            if (n_heads >= 2) return tc_fail("mlb_tc_forward: more than two head inputs");
            buf_of_gemm[head_of_layer[i]] = n_heads++;
        }
    for (int i = 0; i < t->n_ops; ++i) {
        if (t->ops[i].type != MLB_OP_GEMM) continue;
        TcLayer& ly = p.layer[p.n_layers++];
        ly.wplanes = t->wplanes[i], ly.n_kb = t->n_kb[i], ly.flags = t->ops[i].flags, ly.head_buf = buf_of_gemm[i];
        ly.scale = t->blob_dev + t->ops[i].scale_off, ly.shift = t->blob_dev + t->ops[i].shift_off;
    }
    tc_pack_input_kernel<<<296, 256, 0, st>>>(x_dev, t->xplanes[0], B, t->in_size, t->n_kb[first], rows_pad);
    mlb_count_launch();
    const size_t smem = (size_t)TCNST * TC_STAGE;
    TCC(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_forward_kernel<<<dim3(rows_pad / TCM, 4), 128, smem, st>>>(p);
    mlb_count_launch();
    for (int i = 0; i < t->n_ops; ++i) {
        if (t->ops[i].type != MLB_OP_HEAD) continue;
        const mlb_op& op = t->ops[i];
        const int warps = (B + 3) / 4;
        tc_heads_kernel<<<(warps + 7) / 8, 256, 0, st>>>(t->head_f32[buf_of_gemm[head_of_layer[i]]], t->blob_dev + op.w_off,
                                                        t->blob_dev + op.shift_off, op.N, op.K, B, out_raw_dev, t->out_size, op.out_col);
        mlb_count_launch();
    }
    TCC(cudaGetLastError());
    return 0;
}
