// elastic.cu -- the dense sampling grid of ElasticTransformer (spatial_transformer.py:93-362) and its backward.
//
// ElasticTransformer is the reference's second thin-plate-spline formulation: radial basis
// U(r2) = r2 * log(r2) with log(0) -> 0 and NO epsilon (:300-310), coefficients ordered (x, y, 1, w_1..w_pn) (:288,343-360),
// one regular control mesh fixed at construction, the zero-padded sampler bilinear_interp.  No reference call site uses it
// (model.py imports ProjectiveTransformer only), so it is built for completeness, not for speed: one thread per output
// pixel evaluates  x_s, y_s = coefficients @ [x_t; y_t; 1; U_1..U_pn]  (:286-296) -- the [pn+3, h*w] right_mat of the
// reference is never materialised -- and dvsg_bilinear_fwd / _bwd do the sampling.
#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

constexpr int EL_THREADS = 256;
constexpr int EL_CHUNK = 256;      // control points staged in shared memory per round

__device__ __forceinline__ float elastic_u(float d2) { return d2 > 0.0f ? DVSG_MUL(d2, logf(d2)) : 0.0f; }

// coef [B,2,pn+3] rows (x, y, 1, w_k); sp [pn,2] control points; x_out, y_out flat [B*oh*ow]
__global__ void __launch_bounds__(EL_THREADS) elastic_grid_kernel(const float* __restrict__ sp, const float* __restrict__ coef, float* __restrict__ x_out,
                                                                  float* __restrict__ y_out, int oh, int ow, int pn, float step_x, float step_y) {
    __shared__ float4 s_cp[EL_CHUNK];      // (px, py, wx, wy)
    const int b = blockIdx.y, N = pn + 3;
    const long long pix = (long long)blockIdx.x * EL_THREADS + threadIdx.x;
    const bool ok = pix < (long long)oh * ow;
    const int row = ok ? (int)(pix / ow) : 0, col = ok ? (int)(pix % ow) : 0;
    const float xt = lin_coord(col, step_x), yt = lin_coord(row, step_y);      // get_meshgrid(out_w, out_h), :313-322
    const float* cx = coef + (size_t)b * 2 * N;
    const float* cy = cx + N;
    // rows of right_mat in order: x, y, 1, then the radial terms (:288); sequential fp32 sum with separately rounded products
    float X = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(cx), xt), DVSG_MUL(__ldg(cx + 1), yt)), __ldg(cx + 2));
    float Y = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(cy), xt), DVSG_MUL(__ldg(cy + 1), yt)), __ldg(cy + 2));
    for (int k0 = 0; k0 < pn; k0 += EL_CHUNK) {
        const int kc = min(EL_CHUNK, pn - k0);
        __syncthreads();
        for (int k = threadIdx.x; k < kc; k += EL_THREADS)
            s_cp[k] = make_float4(__ldg(sp + 2 * (k0 + k)), __ldg(sp + 2 * (k0 + k) + 1), __ldg(cx + 3 + k0 + k), __ldg(cy + 3 + k0 + k));
        __syncthreads();
        for (int k = 0; k < kc; ++k) {
            const float4 c = s_cp[k];
            const float u = elastic_u(tps_d2(xt, yt, c.x, c.y));
            X = DVSG_ADD(X, DVSG_MUL(c.z, u));
            Y = DVSG_ADD(Y, DVSG_MUL(c.w, u));
        }
    }
    if (ok) { x_out[(size_t)b * oh * ow + pix] = X; y_out[(size_t)b * oh * ow + pix] = Y; }
}

// grad_coef[b][c][q] = sum_pix grad_(x|y)[pix] * right_mat[q][pix]; one CTA per (row q of right_mat, frame), fp64 partial sums
__global__ void __launch_bounds__(EL_THREADS) elastic_grid_bwd_kernel(const float* __restrict__ sp, const float* __restrict__ gx, const float* __restrict__ gy,
                                                                      float* __restrict__ grad_coef, int oh, int ow, int pn, float step_x, float step_y) {
    __shared__ double s_x[EL_THREADS / 32], s_y[EL_THREADS / 32];
    const int q = blockIdx.x, b = blockIdx.y, N = pn + 3;
    const long long n = (long long)oh * ow;
    const float px = q >= 3 ? __ldg(sp + 2 * (q - 3)) : 0.0f, py = q >= 3 ? __ldg(sp + 2 * (q - 3) + 1) : 0.0f;
    double ax = 0.0, ay = 0.0;
    for (long long pix = threadIdx.x; pix < n; pix += EL_THREADS) {
        const int row = (int)(pix / ow), col = (int)(pix % ow);
        const float xt = lin_coord(col, step_x), yt = lin_coord(row, step_y);
        const float v = q == 0 ? xt : (q == 1 ? yt : (q == 2 ? 1.0f : elastic_u(tps_d2(xt, yt, px, py))));
        ax += (double)__ldg(gx + (size_t)b * n + pix) * (double)v;
        ay += (double)__ldg(gy + (size_t)b * n + pix) * (double)v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); }
    if ((threadIdx.x & 31) == 0) { s_x[threadIdx.x >> 5] = ax; s_y[threadIdx.x >> 5] = ay; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tx = 0.0, ty = 0.0;
        for (int w = 0; w < EL_THREADS / 32; ++w) { tx += s_x[w]; ty += s_y[w]; }
        grad_coef[((size_t)b * 2 + 0) * N + q] = (float)tx;
        grad_coef[((size_t)b * 2 + 1) * N + q] = (float)ty;
    }
}

static float el_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_elastic_grid(const float* source_points, const float* coef, float* x_out, float* y_out, int B, int oh, int ow, int pn,
                                 void* stream) {
    DVSG_REQUIRE(B >= 0 && oh >= 0 && ow >= 0 && pn >= 1, "elastic_grid: bad shape");
    if (B == 0 || oh == 0 || ow == 0) return DVSG_OK;
    DVSG_REQUIRE(source_points && coef && x_out && y_out, "elastic_grid: null pointer");
    DVSG_REQUIRE(B <= 65535 && (long long)oh * ow < (1LL << 31), "elastic_grid: batch / frame too large");
    const long long n = (long long)oh * ow;
    elastic_grid_kernel<<<dim3((unsigned)((n + EL_THREADS - 1) / EL_THREADS), (unsigned)B), EL_THREADS, 0, (cudaStream_t)stream>>>(
        source_points, coef, x_out, y_out, oh, ow, pn, el_step(ow), el_step(oh));
    count_launch();
    return check_launch("elastic_grid_kernel");
}

extern "C" int dvsg_elastic_grid_bwd(const float* source_points, const float* grad_x, const float* grad_y, float* grad_coef, int B, int oh, int ow,
                                     int pn, void* stream) {
    DVSG_REQUIRE(B >= 0 && oh > 0 && ow > 0 && pn >= 1, "elastic_grid_bwd: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(source_points && grad_x && grad_y && grad_coef, "elastic_grid_bwd: null pointer");
    DVSG_REQUIRE(B <= 65535 && (long long)oh * ow < (1LL << 31), "elastic_grid_bwd: batch / frame too large");
    elastic_grid_bwd_kernel<<<dim3((unsigned)(pn + 3), (unsigned)B), EL_THREADS, 0, (cudaStream_t)stream>>>(source_points, grad_x, grad_y, grad_coef, oh,
                                                                                                               ow, pn, el_step(ow), el_step(oh));
    count_launch();
    return check_launch("elastic_grid_bwd_kernel");
}
