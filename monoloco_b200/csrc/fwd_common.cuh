// Shared by the two inference kernels (forward.cu: one CTA per 2*TM detections; forward_small.cu: an 8-CTA cluster per
// 16 detections): kernel parameters and the per-row decode.
#pragma once
#include "common.cuh"

namespace mlb {

struct FwdParams {
    const float* blob;
    mlb_op ops[MLB_MAX_OPS];
    int n_ops, in_size, out_size, L, decode_kind;
    int input_kind, flags, n_rows, n_right, n_tiles, kpad0;
    float kinv[9];
    float z_met;
    const float* x;
    const float* xr;
    float* out_raw;
    float* out_dec;
    float* out_xyzc;
    float* out_x;
    const uint8_t* drop_mask;
    unsigned long long drop_seed;
    float p_drop;
    float* res_scratch;
    int* err_flag;
    float* gather[MLB_MAX_PEERS];
    int n_gather;
    long long gather_row0;
};

// Laplace / spherical / orientation decode of one raw output row (process.py:231-278, 330-360; net.py:95-100).
// Explicit __f*_rn intrinsics pin the reference's operation order (no FMA contraction).
__device__ __forceinline__ void decode_row(int kind, int out_size, const float* o, float& x, float& y, float& z, float& d,
                                           float& bi, float& yaw_p, float& yaw_o, float& aux) {
    x = y = z = d = bi = yaw_p = yaw_o = aux = 0.f;
    if (kind == MLB_DECODE_LOCO) {
        const float th = o[0], ps = o[1];
        d = o[2];
        bi = __fmul_rn(expf(o[3]), d);                    // process.py:132
        x = __fmul_rn(__fmul_rn(d, sinf(ps)), cosf(th));  // camera.py:232
        y = __fmul_rn(d, cosf(ps));                       // camera.py:236
        z = sqrtf(__fsub_rn(__fsub_rn(__fmul_rn(d, d), __fmul_rn(x, x)), __fmul_rn(y, y)));  // process.py:265
        yaw_p = atan2f(o[7], o[8]);                       // process.py:272
        if (out_size == 10) aux = 1.0f / (1.0f + expf(-o[9]));  // process.py:277
    } else if (kind == MLB_DECODE_MONO) {
        x = o[0], y = o[1], z = o[2];
        d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));  // process.py:350
        bi = __fmul_rn(expf(o[3]), o[2]);
        yaw_p = atan2f(o[7], o[8]);
    } else if (kind == MLB_DECODE_DB) {
        d = o[0];
        bi = __fmul_rn(expf(o[1]), o[0]);  // net.py:98
    }
    if (kind == MLB_DECODE_LOCO || kind == MLB_DECODE_MONO) {
        yaw_o = __fadd_rn(yaw_p, atan2f(x, z));  // camera.py:203-204
        if (yaw_o > 3.14159265358979323846f) yaw_o = __fsub_rn(yaw_o, 6.28318530717958647692f);
        if (yaw_o < -3.14159265358979323846f) yaw_o = __fadd_rn(yaw_o, 6.28318530717958647692f);
    }
}

}  // namespace mlb
