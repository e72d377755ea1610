// losses.cu -- the consumers of the warp path's outputs inside the training graph:
//
//   sparse TPS evaluation   trainer.py:363-386 (get_surf_loss) reads the dense sampling grid x, y of
//                           ThinPlateSpline at a few hundred feature points per frame (tf.batch_gather at
//                           idx = x + y*w, with one extra entry -1 at idx = h*w for padded features).  Evaluating
//                           the spline AT those points gives the same numbers -- the arithmetic below is the dense
//                           kernels' arithmetic term for term, so the result equals x_flat[b*h*w + idx] bit for
//                           bit -- without materialising 8 B per pixel of grid in training.
//   masked MSE              trainer.py:232-243 (masked_MSE): sum((pred*mask - gt*mask)^2) / sum(mask) per frame
//                           (tf.div_no_nan), mean over the batch; used by the pixel and temporal losses (:245-250).
//                           Two deterministic reduction stages (fixed chunking, fixed order): the same inputs give
//                           the same bits on every run.
#include "dvsg_common.cuh"

namespace dvsg {

constexpr float LN2F = 0.6931471805599453f;

// ---- sparse TPS evaluation ---------------------------------------------------------------------------
// one warp per (frame, 32 points); lane = point.  Same operations in the same order as tile_tps_basis /
// tps_accumulate: affine part by two fmas, then k = 0..pn-1: d2 = (x_t - px)^2 + max((y_t - py)^2, tiny),
// r = d2 * lg2(d2), X = fma(c_k ln2, r, X).
__global__ void __launch_bounds__(128) tps_eval_points_kernel(const float* __restrict__ coord, long long cstride, const float* __restrict__ T,
                                                              const int* __restrict__ idx, float* __restrict__ x_out, float* __restrict__ y_out,
                                                              int oh, int ow, int pn, int P, float step_x, float step_y) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, p = (blockIdx.x * 4 + warp) * 32 + lane;
    if ((blockIdx.x * 4 + warp) * 32 >= P) return;           // whole warp out of range (warp-uniform)
    const int N = pn + 3;
    const float* Tb = T + (size_t)b * 2 * N;
    const float* cb = coord + (size_t)b * cstride;
    const float l0 = tps_affine0(Tb, pn, lane), l3 = tps_affine0(Tb + N, pn, lane);
    if (p >= P) return;
    const int i = __ldg(idx + (size_t)b * P + p);
    float X = -1.0f, Y = -1.0f;                              // idx == h*w: the appended -1 entry (trainer.py:364-365)
    if (i >= 0 && i < oh * ow) {
        const float xt = lin_coord(i % ow, step_x), yt = lin_coord(i / ow, step_y);
        X = fmaf(__ldg(Tb + 2), yt, fmaf(__ldg(Tb + 1), xt, l0));
        Y = fmaf(__ldg(Tb + N + 2), yt, fmaf(__ldg(Tb + N + 1), xt, l3));
        for (int k = 0; k < pn; ++k) {
            const float dx = DVSG_ADD(xt, -__ldg(cb + 2 * k));
            const float d2 = DVSG_ADD(DVSG_MUL(dx, dx), tps_dy2(yt, __ldg(cb + 2 * k + 1)));
            const float r = DVSG_MUL(d2, lg2_approx(d2));
            X = fmaf(__ldg(Tb + 3 + k) * LN2F, r, X);
            Y = fmaf(__ldg(Tb + N + 3 + k) * LN2F, r, Y);
        }
    }
    x_out[(size_t)b * P + p] = X;
    y_out[(size_t)b * P + p] = Y;
}

// grad_T[b, o, :] = sum_p g_o(p) * d(X or Y)(p) / dT: (1, x_t, y_t, r_k ln2 + 1e-6).  One warp per frame, lanes stride
// over the points, one shuffle tree per coefficient: fixed order, no atomics, grad_T is overwritten.
__global__ void __launch_bounds__(32) tps_eval_points_bwd_kernel(const float* __restrict__ coord, long long cstride, const int* __restrict__ idx,
                                                                 const float* __restrict__ gx, const float* __restrict__ gy, float* __restrict__ grad_T,
                                                                 int oh, int ow, int pn, int P, float step_x, float step_y) {
    const int lane = threadIdx.x, b = blockIdx.x, N = pn + 3;
    const float* cb = coord + (size_t)b * cstride;
    float* gT = grad_T + (size_t)b * 2 * N;
    for (int k = -3; k < pn; ++k) {
        const float px = k >= 0 ? __ldg(cb + 2 * k) : 0.0f, py = k >= 0 ? __ldg(cb + 2 * k + 1) : 0.0f;
        float ax = 0.0f, ay = 0.0f;
        for (int p = lane; p < P; p += 32) {
            const int i = __ldg(idx + (size_t)b * P + p);
            if (i < 0 || i >= oh * ow) continue;
            const float xt = lin_coord(i % ow, step_x), yt = lin_coord(i / ow, step_y);
            float w;
            if (k == -3) w = 1.0f;
            else if (k == -2) w = xt;
            else if (k == -1) w = yt;
            else {
                const float dx = DVSG_ADD(xt, -px);
                const float d2 = DVSG_ADD(DVSG_MUL(dx, dx), tps_dy2(yt, py));
                w = fmaf(DVSG_MUL(d2, lg2_approx(d2)), LN2F, TPS_EPS);
            }
            ax = fmaf(__ldg(gx + (size_t)b * P + p), w, ax);
            ay = fmaf(__ldg(gy + (size_t)b * P + p), w, ay);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); }
        if (lane == 0) { gT[k + 3] = ax; gT[N + k + 3] = ay; }
    }
}

// ---- masked MSE ------------------------------------------------------------------------------------------
constexpr int MSE_NT = 256;
constexpr int MSE_CHUNK = 16384;      // elements per CTA: 64 per thread

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < MSE_NT / 32; ++w) t += s_red[w];
    }
    __syncthreads();
    return t;                          // valid on thread 0
}

// stage 1: partial[b][chunk] = (sum (pred*mask - gt*mask)^2, sum mask) over one chunk of frame b
__global__ void __launch_bounds__(MSE_NT) masked_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                                    const float* __restrict__ mask, int mask_c, int C, long long n,
                                                                    float2* __restrict__ partial) {
    __shared__ float s_red[MSE_NT / 32];
    const int b = blockIdx.y;
    const long long e0 = (long long)blockIdx.x * MSE_CHUNK, e1 = e0 + MSE_CHUNK < n ? e0 + MSE_CHUNK : n;
    const float* pb = pred + (size_t)b * n;
    const float* gb = gt + (size_t)b * n;
    const float* mb = mask + (size_t)b * (mask_c == C ? n : n / C);
    float sq = 0.0f, ms = 0.0f;
    for (long long e = e0 + threadIdx.x; e < e1; e += MSE_NT) {
        const float m = __ldg(mb + (mask_c == C ? e : e / C));
        const float d = DVSG_SUB(DVSG_MUL(__ldg(pb + e), m), DVSG_MUL(__ldg(gb + e), m));   // squared_difference(pred*mask, gt*mask)
        sq = fmaf(d, d, sq);
        // reduce_sum(mask) runs over the mask's own elements: a 1-channel mask is counted once per pixel
        if (mask_c == C || e % C == 0) ms += m;
    }
    const float tsq = block_sum(sq, s_red), tms = block_sum(ms, s_red);
    if (threadIdx.x == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = make_float2(tsq, tms);
}

// stage 2: per-frame totals in chunk order, div_no_nan, mean over the batch
__global__ void __launch_bounds__(32) masked_mse_final_kernel(const float2* __restrict__ partial, int chunks, int B, float* __restrict__ sq_out,
                                                              float* __restrict__ msum_out, float* __restrict__ loss_out) {
    const int lane = threadIdx.x;
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) {
        float sq = 0.0f, ms = 0.0f;
        for (int c = lane; c < chunks; c += 32) { const float2 v = partial[(size_t)b * chunks + c]; sq += v.x; ms += v.y; }
#pragma unroll
        for (int o = 16; o; o >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o); ms += __shfl_xor_sync(0xffffffffu, ms, o); }
        if (lane == 0) { sq_out[b] = sq; msum_out[b] = ms; }
        acc += ms != 0.0f ? sq / ms : 0.0f;              // tf.div_no_nan (trainer.py:242)
    }
    if (lane == 0) loss_out[0] = acc / (float)B;         // tf.reduce_mean (:243)
}

// backward of loss = mean_b sq_b / msum_b: d/dpred = 2 (pred - gt) m^2 / msum_b / B, d/dgt = -d/dpred,
// d/dmask = (2 (pred - gt)^2 m / msum_b - sq_b / msum_b^2) / B (summed over channels for a 1-channel mask)
__global__ void __launch_bounds__(MSE_NT) masked_mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask,
                                                                int mask_c, int C, long long n, const float* __restrict__ sq, const float* __restrict__ msum,
                                                                float gscale, float* __restrict__ g_pred, float* __restrict__ g_gt, float* __restrict__ g_mask) {
    const int b = blockIdx.y;
    const float ms = __ldg(msum + b), s = __ldg(sq + b);
    const float inv = ms != 0.0f ? gscale / ms : 0.0f, tail = ms != 0.0f ? gscale * s / (ms * ms) : 0.0f;
    const long long npx = n / C;
    const float* pb = pred + (size_t)b * n;
    const float* gb = gt + (size_t)b * n;
    const float* mb = mask + (size_t)b * (mask_c == C ? n : npx);
    for (long long px = (long long)blockIdx.x * MSE_NT + threadIdx.x; px < npx; px += (long long)gridDim.x * MSE_NT) {
        float gm1 = 0.0f;
        for (int ch = 0; ch < C; ++ch) {
            const long long e = px * C + ch;
            const float m = __ldg(mb + (mask_c == C ? e : px));
            const float d = __ldg(pb + e) - __ldg(gb + e);
            const float gp = 2.0f * d * m * m * inv;
            if (g_pred) g_pred[(size_t)b * n + e] = gp;
            if (g_gt) g_gt[(size_t)b * n + e] = -gp;
            const float gm = 2.0f * d * d * m * inv;
            if (mask_c == C) { if (g_mask) g_mask[(size_t)b * n + e] = gm - tail; }
            else gm1 += gm;
        }
        if (mask_c != C && g_mask) g_mask[(size_t)b * npx + px] = gm1 - tail;
    }
}

static float lin_step_l(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_tps_eval_points(const float* coord, long long coord_batch_stride, const float* T, const int* idx, float* x_out,
                                    float* y_out, int B, int oh, int ow, int pn, int P, void* stream) {
    DVSG_REQUIRE(B >= 0 && oh > 0 && ow > 0 && pn > 0 && P >= 0, "tps_eval_points: bad shape");
    if (B == 0 || P == 0) return DVSG_OK;
    DVSG_REQUIRE(coord && T && idx && x_out && y_out, "tps_eval_points: null pointer");
    DVSG_REQUIRE(coord_batch_stride == 0 || coord_batch_stride >= 2LL * pn, "tps_eval_points: coord stride %lld < 2*pn", coord_batch_stride);
    DVSG_REQUIRE(B <= 65535 && (long long)oh * ow < (1LL << 31) - 1, "tps_eval_points: batch or grid too large");
    const dim3 grid((P + 127) / 128, B);
    tps_eval_points_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(coord, coord_batch_stride, T, idx, x_out, y_out, oh, ow, pn, P,
                                                                   lin_step_l(ow), lin_step_l(oh));
    count_launch();
    return check_launch("tps_eval_points_kernel");
}

extern "C" int dvsg_tps_eval_points_bwd(const float* coord, long long coord_batch_stride, const int* idx, const float* grad_x,
                                        const float* grad_y, float* grad_T, int B, int oh, int ow, int pn, int P, void* stream) {
    DVSG_REQUIRE(B >= 0 && oh > 0 && ow > 0 && pn > 0 && P >= 0, "tps_eval_points_bwd: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(coord && grad_T && (P == 0 || (idx && grad_x && grad_y)), "tps_eval_points_bwd: null pointer");
    DVSG_REQUIRE(coord_batch_stride == 0 || coord_batch_stride >= 2LL * pn, "tps_eval_points_bwd: coord stride %lld < 2*pn", coord_batch_stride);
    tps_eval_points_bwd_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(coord, coord_batch_stride, idx, grad_x, grad_y, grad_T, oh, ow, pn, P,
                                                                   lin_step_l(ow), lin_step_l(oh));
    count_launch();
    return check_launch("tps_eval_points_bwd_kernel");
}

extern "C" size_t dvsg_masked_mse_workspace_bytes(int B, long long n_per_frame) {
    if (B <= 0 || n_per_frame <= 0) return 0;
    return (size_t)B * (size_t)((n_per_frame + MSE_CHUNK - 1) / MSE_CHUNK) * sizeof(float2);
}

extern "C" int dvsg_masked_mse_fwd(const float* pred, const float* gt, const float* mask, int mask_channels, float* sq_out, float* msum_out,
                                   float* loss_out, void* workspace, size_t workspace_bytes, int B, long long n_pixels, int C, void* stream) {
    DVSG_REQUIRE(B > 0 && n_pixels > 0 && C > 0, "masked_mse_fwd: bad shape");
    DVSG_REQUIRE(mask_channels == C || mask_channels == 1, "masked_mse_fwd: mask has %d channels, expected %d or 1", mask_channels, C);
    DVSG_REQUIRE(pred && gt && mask && sq_out && msum_out && loss_out, "masked_mse_fwd: null pointer");
    DVSG_REQUIRE(B <= 65535, "masked_mse_fwd: batch %d exceeds the grid y limit", B);
    const long long n = n_pixels * C;
    const long long chunks = (n + MSE_CHUNK - 1) / MSE_CHUNK;
    DVSG_REQUIRE(chunks < (1LL << 31), "masked_mse_fwd: frame too large");
    if (!workspace || workspace_bytes < dvsg_masked_mse_workspace_bytes(B, n)) {
        set_error("masked_mse_fwd: workspace of %zu bytes needed, %zu given", dvsg_masked_mse_workspace_bytes(B, n), workspace_bytes);
        return DVSG_ERR_WORKSPACE;
    }
    masked_mse_partial_kernel<<<dim3((unsigned)chunks, B), MSE_NT, 0, (cudaStream_t)stream>>>(pred, gt, mask, mask_channels, C, n,
                                                                                               reinterpret_cast<float2*>(workspace));
    masked_mse_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(workspace), (int)chunks, B, sq_out, msum_out, loss_out);
    count_launch(2);
    return check_launch("masked_mse kernels");
}

extern "C" int dvsg_masked_mse_bwd(const float* pred, const float* gt, const float* mask, int mask_channels, const float* sq,
                                   const float* msum, float grad_loss, float* grad_pred, float* grad_gt, float* grad_mask, int B,
                                   long long n_pixels, int C, void* stream) {
    DVSG_REQUIRE(B > 0 && n_pixels > 0 && C > 0, "masked_mse_bwd: bad shape");
    DVSG_REQUIRE(mask_channels == C || mask_channels == 1, "masked_mse_bwd: mask has %d channels, expected %d or 1", mask_channels, C);
    DVSG_REQUIRE(pred && gt && mask && sq && msum, "masked_mse_bwd: null pointer");
    DVSG_REQUIRE(B <= 65535, "masked_mse_bwd: batch %d exceeds the grid y limit", B);
    const long long blocks = (n_pixels + MSE_NT - 1) / MSE_NT;
    const unsigned gx = (unsigned)(blocks < 148 * 8 ? blocks : 148 * 8);
    masked_mse_bwd_kernel<<<dim3(gx, B), MSE_NT, 0, (cudaStream_t)stream>>>(pred, gt, mask, mask_channels, C, n_pixels * C, sq, msum,
                                                                             grad_loss / (float)B, grad_pred, grad_gt, grad_mask);
    count_launch();
    return check_launch("masked_mse_bwd_kernel");
}
