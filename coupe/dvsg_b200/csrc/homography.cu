// homography.cu -- gradient of the grid stage of ProjectiveTransformer / AffineTransformer w.r.t. theta
// (spatial_transformer.py:423-452, 73-91).
//
// The transformers' forward is fused with the sampler (dvsg_homography_warp_fwd); their backward w.r.t. the input and the
// sampling coordinates is the sampler's (dvsg_bilinear_bwd, B3).  What is left is the chain from (grad x_s, grad y_s) to
// theta -- what tf.gradients assembles for  T_g = theta @ [x_t; y_t; 1] (:437)  and  div_no_nan(T_g[0|1], T_g[2]) (:446-447):
//     d num  = div_no_nan(g, z)                         DivNoNan, first argument
//     d z    = g * div_no_nan(div_no_nan(-num, z), z)   DivNoNan, second argument: zero where z == 0
//     d theta[i][j] = sum_pix d T_g[i][pix] * grid[j][pix]
// No reference call site differentiates the transformers (model.py:156-167 feeds random constants), so this is built for
// completeness, not speed: one CTA per (entry of theta, frame) walks the frame with fp64 partial sums -- deterministic.
#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

constexpr int HG_THREADS = 512;

__global__ void __launch_bounds__(HG_THREADS) homography_grid_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ gx,
                                                                         const float* __restrict__ gy, float* __restrict__ grad_theta, int projective,
                                                                         int oh, int ow, float step_x, float step_y) {
    __shared__ double s_part[HG_THREADS / 32];
    __shared__ float s_h[9];
    const int q = blockIdx.x, b = blockIdx.y, nt = projective ? 8 : 6;
    const int i = q / 3, j = q % 3;                    // entry (i, j) of the 3x3 (2x3) matrix
    if (threadIdx.x < 9) s_h[threadIdx.x] = threadIdx.x < nt ? __ldg(theta + (size_t)b * nt + threadIdx.x) : (threadIdx.x == 8 ? 1.0f : 0.0f);
    __syncthreads();
    const long long n = (long long)oh * ow;
    const float* gxb = gx + (size_t)b * n;
    const float* gyb = gy + (size_t)b * n;
    double acc = 0.0;
    for (long long pix = threadIdx.x; pix < n; pix += HG_THREADS) {
        const int row = (int)(pix / ow), col = (int)(pix % ow);
        const float xt = lin_coord(col, step_x), yt = lin_coord(row, step_y);
        const float gj = j == 0 ? xt : (j == 1 ? yt : 1.0f);
        float g;                                        // d T_g[i][pix]
        if (!projective) {
            g = i == 0 ? __ldg(gxb + pix) : __ldg(gyb + pix);
        } else {
            // the forward's own numerators and denominator (warp_fwd.cu, MODE_HOMOG)
            const float zn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_h[6], xt), DVSG_MUL(s_h[7], yt)), s_h[8]);
            if (zn == 0.0f) {
                g = 0.0f;
            } else if (i < 2) {
                g = DVSG_DIV(i == 0 ? __ldg(gxb + pix) : __ldg(gyb + pix), zn);
            } else {
                const float xn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_h[0], xt), DVSG_MUL(s_h[1], yt)), s_h[2]);
                const float yn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_h[3], xt), DVSG_MUL(s_h[4], yt)), s_h[5]);
                const float gxz = DVSG_DIV(__ldg(gxb + pix), zn), gyz = DVSG_DIV(__ldg(gyb + pix), zn);
                g = -(gxz * DVSG_DIV(xn, zn) + gyz * DVSG_DIV(yn, zn));
            }
        }
        acc += (double)g * (double)gj;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < HG_THREADS / 32; ++w) t += s_part[w];
        grad_theta[(size_t)b * nt + q] = (float)t;
    }
}

static float hg_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_homography_grid_bwd(const float* theta, const float* grad_x, const float* grad_y, int projective, float* grad_theta, int B,
                                        int oh, int ow, void* stream) {
    DVSG_REQUIRE(B >= 0 && oh > 0 && ow > 0, "homography_grid_bwd: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(theta && grad_x && grad_y && grad_theta, "homography_grid_bwd: null pointer");
    DVSG_REQUIRE(B <= 65535 && (long long)oh * ow < (1LL << 31), "homography_grid_bwd: batch / frame too large");
    homography_grid_bwd_kernel<<<dim3(projective ? 8u : 6u, (unsigned)B), HG_THREADS, 0, (cudaStream_t)stream>>>(theta, grad_x, grad_y, grad_theta,
                                                                                                                projective ? 1 : 0, oh, ow, hg_step(ow),
                                                                                                                hg_step(oh));
    count_launch();
    return check_launch("homography_grid_bwd_kernel");
}
