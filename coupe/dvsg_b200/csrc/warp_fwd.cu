// warp_fwd.cu -- forward kernels of the DVSG warp path for sm_100a.
//
// One kernel template covers the four coordinate sources of the path
//   MODE_TPS    fused TPS grid generation + A4 sampler   (ThinPlateSpline.py:92-141, 30-90)
//   MODE_GIVEN  bilinear_interp on caller-given x, y      (spatial_transformer.py:496-563)
//   MODE_FLOW   tf_warp                                   (warp_with_optical_flow.py:96-176)
//   MODE_HOMOG  Projective/AffineTransformer + bilinear   (spatial_transformer.py:73-91, 423-452)
// and two data paths
//   STAGED  (C == 3, 16-B aligned rows): each CTA owns a 64x16 output tile; it computes the
//           sampling coordinates, reduces the exact bounding box of the source pixels the
//           tile touches, pulls that footprint into shared memory with 1-D bulk async
//           copies (cp.async.bulk, completion on an mbarrier), gathers the four corners
//           from shared memory, blends, stages the output tile in shared memory and writes
//           it back with bulk async stores.  The sampling grid never exists in HBM unless
//           the caller asks for x, y.
//   DIRECT  (any C, any alignment, or a footprint larger than the staging buffer): same
//           arithmetic, corners gathered straight from global memory.
//
// HBM traffic (algorithmic): 4C B read + 4C B written per output pixel (+8 B if x, y are
// materialised, +8 B flow read for MODE_FLOW, +8 B x,y read for MODE_GIVEN).
// The TPS basis costs pn MUFU.LG2 per pixel: for pn = 16 the kernel is co-limited by the
// MUFU pipe (16 lanes/clk/SM) and HBM; for pn = 256 it is MUFU-bound.
#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

enum { MODE_TPS = 0, MODE_GIVEN = 1, MODE_FLOW = 2, MODE_HOMOG = 3 };

constexpr int TW = 64;        // tile width in output pixels (one thread per column)
constexpr int PR = 4;         // consecutive rows per thread
constexpr int RG = 4;         // row groups per CTA
constexpr int TH = PR * RG;   // tile height = 16
constexpr int NT = TW * RG;   // 256 threads
constexpr int KC = 256;       // control points held in shared memory at a time
constexpr float LN2 = 0.6931471805599453f;

struct FwdParams {
    const float* src;   // [B,H,W,C]
    float* out;         // [B,oh,ow,C]
    float* x_out;       // optional [B*oh*ow]
    float* y_out;
    float* mask_out;    // optional [B,oh,ow] (MODE_TPS)
    int B, H, W, C, oh, ow;
    // MODE_TPS
    const float* coord;
    long long coord_stride;
    const float* T;
    int pn, kc_cap;
    float step_x, step_y;   // 2/(ow-1), 2/(oh-1) in fp32 (tf.linspace step)
    // MODE_GIVEN
    const float* x_in;
    const float* y_in;
    // MODE_FLOW
    const float* flow;
    // MODE_HOMOG
    const float* theta;
    int projective;
    // STAGED
    int src_smem_bytes;
};

// Accumulate sum_k c_k * d2_k * ln(d2_k + 1e-6) for PR consecutive rows of one column.
// pt[k] = (px, py, cx*ln2, cy*ln2); dy2[k*TH + r] = (y_t(row0+r) - py_k)^2.
template <bool PACK>
__device__ __forceinline__ void tps_accumulate(const float4* __restrict__ pt, const float* __restrict__ dy2, int kc,
                                               float xt, int rbase, float (&xs)[PR], float (&ys)[PR]) {
    if (PACK) {
        // packed fp32x2 (FADD2/FMUL2/FFMA2): two rows per instruction, halves the issue slots
        float2 xa = make_float2(xs[0], xs[1]), xb = make_float2(xs[2], xs[3]);
        float2 ya = make_float2(ys[0], ys[1]), yb = make_float2(ys[2], ys[3]);
        const float2 eps = make_float2(1e-6f, 1e-6f);
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
            const float4 q = pt[k];
            const float dx = DVSG_SUB(xt, q.x);
            const float dx2 = DVSG_MUL(dx, dx);
            const float4 d = *reinterpret_cast<const float4*>(dy2 + k * TH + rbase);
            const float2 dxx = make_float2(dx2, dx2);
            const float2 d2a = __fadd2_rn(dxx, make_float2(d.x, d.y));
            const float2 d2b = __fadd2_rn(dxx, make_float2(d.z, d.w));
            const float2 ta = __fadd2_rn(d2a, eps);
            const float2 tb = __fadd2_rn(d2b, eps);
            const float2 ra = __fmul2_rn(d2a, make_float2(lg2_approx(ta.x), lg2_approx(ta.y)));
            const float2 rb = __fmul2_rn(d2b, make_float2(lg2_approx(tb.x), lg2_approx(tb.y)));
            const float2 cx = make_float2(q.z, q.z), cy = make_float2(q.w, q.w);
            xa = __ffma2_rn(cx, ra, xa);
            xb = __ffma2_rn(cx, rb, xb);
            ya = __ffma2_rn(cy, ra, ya);
            yb = __ffma2_rn(cy, rb, yb);
        }
        xs[0] = xa.x; xs[1] = xa.y; xs[2] = xb.x; xs[3] = xb.y;
        ys[0] = ya.x; ys[1] = ya.y; ys[2] = yb.x; ys[3] = yb.y;
    } else {
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
            const float4 q = pt[k];
            const float dx = DVSG_SUB(xt, q.x);
            const float dx2 = DVSG_MUL(dx, dx);
            const float4 d = *reinterpret_cast<const float4*>(dy2 + k * TH + rbase);
            const float dv[PR] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int p = 0; p < PR; ++p) {
                const float d2 = DVSG_ADD(dx2, dv[p]);
                const float r = DVSG_MUL(d2, lg2_approx(DVSG_ADD(d2, 1e-6f)));
                xs[p] = fmaf(q.z, r, xs[p]);
                ys[p] = fmaf(q.w, r, ys[p]);
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ Corners corners_of(float xs, float ys, int W, int H) {
    if (MODE == MODE_TPS) return a4_corners(xs, ys, W, H);
    return zp_corners(xs, ys, W, H);   // xs, ys already pixel-space for the ZP modes
}

template <int MODE, bool STAGED, bool PACK>
__global__ void __launch_bounds__(NT, STAGED ? 4 : 3) warp_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ int s_bbox[4];      // xmin, xmax, ymin, ymax of the source pixels touched
    __shared__ float s_lin[12];    // affine TPS coefficients or the 3x3 homography

    float* s_out = reinterpret_cast<float*>(smem);
    float* s_src = s_out + (STAGED ? TH * TW * 3 : 0);
    float4* s_pt = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s_src) + (STAGED ? p.src_smem_bytes : 0));
    float* s_dy2 = reinterpret_cast<float*>(s_pt + p.kc_cap);

    const int tid = threadIdx.x;
    const int tx = tid & (TW - 1);
    const int rg = tid >> 6;
    const int col0 = blockIdx.x * TW, row0 = blockIdx.y * TH, b = blockIdx.z;
    const int col = col0 + tx;
    const int rbase = rg * PR;
    const int H = p.H, W = p.W, C = p.C, oh = p.oh, ow = p.ow;
    const bool col_ok = col < ow;

    if (STAGED && tid == 0) {
        mbar_init(smem_u32(&s_mbar), 1);
        fence_mbar_init();
        s_bbox[0] = 0x7fffffff; s_bbox[1] = -1; s_bbox[2] = 0x7fffffff; s_bbox[3] = -1;
    }

    // ---- phase 1: sampling coordinates for PR rows of this column --------------------
    // MODE_TPS: xs, ys normalised; other modes: pixel-space (before the clip).
    float xs[PR], ys[PR];
    if (MODE == MODE_TPS) {
        const int N = p.pn + 3;
        const float* Tb = p.T + (size_t)b * 2 * N;
        const float* cb = p.coord + (size_t)b * p.coord_stride;
        if (tid < 6) s_lin[tid] = __ldg(Tb + (tid < 3 ? tid : N + tid - 3));
        for (int k0 = 0; k0 < p.pn; k0 += p.kc_cap) {
            const int kc = min(p.kc_cap, p.pn - k0);
            if (k0 > 0) __syncthreads();
            for (int k = tid; k < kc; k += NT)
                s_pt[k] = make_float4(__ldg(cb + 2 * (k0 + k)), __ldg(cb + 2 * (k0 + k) + 1),
                                      __ldg(Tb + 3 + k0 + k) * LN2, __ldg(Tb + N + 3 + k0 + k) * LN2);
            for (int i = tid; i < kc * TH; i += NT) {
                const int k = i / TH, r = i % TH;
                const float dy = DVSG_SUB(lin_coord(row0 + r, p.step_y), __ldg(cb + 2 * (k0 + k) + 1));
                s_dy2[i] = DVSG_MUL(dy, dy);
            }
            __syncthreads();
            const float xt = lin_coord(col, p.step_x);
            if (k0 == 0) {
#pragma unroll
                for (int q = 0; q < PR; ++q) {
                    const float yt = lin_coord(row0 + rbase + q, p.step_y);
                    xs[q] = fmaf(s_lin[2], yt, fmaf(s_lin[1], xt, s_lin[0]));
                    ys[q] = fmaf(s_lin[5], yt, fmaf(s_lin[4], xt, s_lin[3]));
                }
            }
            tps_accumulate<PACK>(s_pt, s_dy2, kc, xt, rbase, xs, ys);
        }
    } else if (MODE == MODE_GIVEN) {
#pragma unroll
        for (int q = 0; q < PR; ++q) {
            const int row = row0 + rbase + q;
            xs[q] = ys[q] = 0.0f;
            if (col_ok && row < oh) {
                const size_t i = ((size_t)b * oh + row) * ow + col;
                xs[q] = zp_pix_from_norm(__ldg(p.x_in + i), W);
                ys[q] = zp_pix_from_norm(__ldg(p.y_in + i), H);
            }
        }
    } else if (MODE == MODE_FLOW) {
#pragma unroll
        for (int q = 0; q < PR; ++q) {
            const int row = row0 + rbase + q;
            xs[q] = ys[q] = 0.0f;
            if (col_ok && row < oh) {
                const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + ((size_t)b * oh + row) * ow + col);
                xs[q] = DVSG_ADD((float)col, f.x);   // warp_with_optical_flow.py:107-120
                ys[q] = DVSG_ADD((float)row, f.y);
            }
        }
    } else {  // MODE_HOMOG
        const int nt = p.projective ? 8 : 6;
        if (tid < 9) s_lin[tid] = tid < nt ? __ldg(p.theta + (size_t)b * nt + tid) : (tid == 8 ? 1.0f : 0.0f);
        __syncthreads();
        const float xt = lin_coord(col, p.step_x);
#pragma unroll
        for (int q = 0; q < PR; ++q) {
            const float yt = lin_coord(row0 + rbase + q, p.step_y);
            // rows of theta @ [x_t; y_t; 1], accumulated k = 0,1,2 (spatial_transformer.py:437)
            float xn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[0], xt), DVSG_MUL(s_lin[1], yt)), s_lin[2]);
            float yn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[3], xt), DVSG_MUL(s_lin[4], yt)), s_lin[5]);
            if (p.projective) {
                const float zn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[6], xt), DVSG_MUL(s_lin[7], yt)), s_lin[8]);
                xn = zn != 0.0f ? DVSG_DIV(xn, zn) : 0.0f;   // tf.div_no_nan, :446-447
                yn = zn != 0.0f ? DVSG_DIV(yn, zn) : 0.0f;
            }
            const int row = row0 + rbase + q;
            if (p.x_out && col_ok && row < oh) {
                const size_t i = ((size_t)b * oh + row) * ow + col;
                p.x_out[i] = xn;
                p.y_out[i] = yn;
            }
            xs[q] = zp_pix_from_norm(xn, W);
            ys[q] = zp_pix_from_norm(yn, H);
        }
    }
    if (MODE == MODE_TPS && p.x_out && col_ok) {
#pragma unroll
        for (int q = 0; q < PR; ++q) {
            const int row = row0 + rbase + q;
            if (row < oh) {
                const size_t i = ((size_t)b * oh + row) * ow + col;
                p.x_out[i] = xs[q];
                p.y_out[i] = ys[q];
            }
        }
    }

    const float* srcb = p.src + (size_t)b * H * W * C;

    // ---- phase 2 (STAGED): exact source footprint of the tile -> shared memory ---------
    bool fits = false;
    int fx0 = 0, fy0 = 0, frow = 0;   // footprint origin (px) and row pitch (floats)
    if (STAGED) {
        int xmin = 0x7fffffff, xmax = -1, ymin = 0x7fffffff, ymax = -1;
#pragma unroll
        for (int q = 0; q < PR; ++q) {
            if (col_ok && row0 + rbase + q < oh) {
                const Corners c = corners_of<MODE>(xs[q], ys[q], W, H);
                if (MODE == MODE_TPS) {
                    xmin = min(xmin, c.x0); xmax = max(xmax, c.x1);
                    ymin = min(ymin, c.y0); ymax = max(ymax, c.y1);
                } else {
                    // padded coords -> real pixels; only corners that land inside the frame
                    const int ax0 = max(c.x0, 1) - 1, ax1 = min(c.x1, W) - 1;
                    const int ay0 = max(c.y0, 1) - 1, ay1 = min(c.y1, H) - 1;
                    if (ax0 <= ax1 && ay0 <= ay1) {
                        xmin = min(xmin, ax0); xmax = max(xmax, ax1);
                        ymin = min(ymin, ay0); ymax = max(ymax, ay1);
                    }
                }
            }
        }
        __syncthreads();   // s_bbox / mbarrier initialised
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        ymin = __reduce_min_sync(0xffffffffu, ymin);
        ymax = __reduce_max_sync(0xffffffffu, ymax);
        if ((tid & 31) == 0) {
            atomicMin(&s_bbox[0], xmin); atomicMax(&s_bbox[1], xmax);
            atomicMin(&s_bbox[2], ymin); atomicMax(&s_bbox[3], ymax);
        }
        __syncthreads();
        xmin = s_bbox[0]; xmax = s_bbox[1]; ymin = s_bbox[2]; ymax = s_bbox[3];
        if (xmax >= xmin && ymax >= ymin) {
            fx0 = xmin & ~3;                                    // 4 px = 48 B keeps rows 16-B aligned
            const int wpx = min(((xmax - fx0 + 1) + 3) & ~3, W - fx0);
            const int nrows = ymax - ymin + 1;
            const unsigned row_bytes = (unsigned)wpx * 12u;
            fy0 = ymin;
            frow = wpx * 3;
            fits = (size_t)row_bytes * nrows <= (size_t)p.src_smem_bytes;
            if (fits && tid < 32) {
                const uint32_t mbar = smem_u32(&s_mbar);
                if (tid == 0) mbar_arrive_expect_tx(mbar, row_bytes * (unsigned)nrows);
                __syncwarp();
                for (int r = tid; r < nrows; r += 32)
                    bulk_g2s(smem_u32(s_src + (size_t)r * frow), srcb + ((size_t)(fy0 + r) * W + fx0) * 3, row_bytes, mbar);
            }
            if (fits) mbar_wait(smem_u32(&s_mbar), 0);
        }
    }

    // ---- phase 3: gather, blend, store ---------------------------------------------------
#pragma unroll
    for (int q = 0; q < PR; ++q) {
        const int row = row0 + rbase + q;
        if (!(col_ok && row < oh)) continue;
        const Corners c = corners_of<MODE>(xs[q], ys[q], W, H);
        if (MODE == MODE_TPS && p.mask_out) {
            const float wa = DVSG_MUL(c.ax1, c.ay1), wb = DVSG_MUL(c.ax1, c.ay0);
            const float wc = DVSG_MUL(c.ax0, c.ay1), wd = DVSG_MUL(c.ax0, c.ay0);
            p.mask_out[((size_t)b * oh + row) * ow + col] = DVSG_ADD(DVSG_ADD(DVSG_ADD(wa, wb), wc), wd);
        }
        // element offsets of the four corners; v* = corner lies inside the frame
        bool v00 = true, v01 = true, v10 = true, v11 = true;
        int x0 = c.x0, x1 = c.x1, y0 = c.y0, y1 = c.y1;
        if (MODE != MODE_TPS) {
            const bool vx0 = zp_valid(x0, W), vx1 = zp_valid(x1, W), vy0 = zp_valid(y0, H), vy1 = zp_valid(y1, H);
            v00 = vx0 && vy0; v01 = vx1 && vy0; v10 = vx0 && vy1; v11 = vx1 && vy1;
            x0 = max(x0, 1) - 1; x1 = min(x1, W) - 1; y0 = max(y0, 1) - 1; y1 = min(y1, H) - 1;   // keep addresses legal
            x0 = min(x0, W - 1); x1 = max(x1, 0); y0 = min(y0, H - 1); y1 = max(y1, 0);
        }
        if (STAGED) {
            float* o = s_out + ((rbase + q) * TW + tx) * 3;
            if (fits) {
                const float* r0 = s_src + (y0 - fy0) * frow - fx0 * 3;
                const float* r1 = s_src + (y1 - fy0) * frow - fx0 * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if (MODE == MODE_TPS) {
                        o[ch] = a4_blend(c, r0[x0 * 3 + ch], r1[x0 * 3 + ch], r0[x1 * 3 + ch], r1[x1 * 3 + ch]);
                    } else {
                        const float i00 = v00 ? r0[x0 * 3 + ch] : 0.0f, i01 = v01 ? r0[x1 * 3 + ch] : 0.0f;
                        const float i10 = v10 ? r1[x0 * 3 + ch] : 0.0f, i11 = v11 ? r1[x1 * 3 + ch] : 0.0f;
                        o[ch] = zp_blend(c, i00, i01, i10, i11);
                    }
                }
            } else {
                const float* r0 = srcb + (size_t)y0 * W * 3;
                const float* r1 = srcb + (size_t)y1 * W * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    if (MODE == MODE_TPS) {
                        o[ch] = a4_blend(c, __ldg(r0 + x0 * 3 + ch), __ldg(r1 + x0 * 3 + ch), __ldg(r0 + x1 * 3 + ch),
                                         __ldg(r1 + x1 * 3 + ch));
                    } else {
                        const float i00 = v00 ? __ldg(r0 + x0 * 3 + ch) : 0.0f, i01 = v01 ? __ldg(r0 + x1 * 3 + ch) : 0.0f;
                        const float i10 = v10 ? __ldg(r1 + x0 * 3 + ch) : 0.0f, i11 = v11 ? __ldg(r1 + x1 * 3 + ch) : 0.0f;
                        o[ch] = zp_blend(c, i00, i01, i10, i11);
                    }
                }
            }
        } else {
            float* o = p.out + (((size_t)b * oh + row) * ow + col) * C;
            const float* p00 = srcb + ((size_t)y0 * W + x0) * C;
            const float* p01 = srcb + ((size_t)y0 * W + x1) * C;
            const float* p10 = srcb + ((size_t)y1 * W + x0) * C;
            const float* p11 = srcb + ((size_t)y1 * W + x1) * C;
            for (int ch = 0; ch < C; ++ch) {
                if (MODE == MODE_TPS) {
                    // A4 corner order: a=(x0,y0) b=(x0,y1) c=(x1,y0) d=(x1,y1)
                    o[ch] = a4_blend(c, __ldg(p00 + ch), __ldg(p10 + ch), __ldg(p01 + ch), __ldg(p11 + ch));
                } else {
                    const float i00 = v00 ? __ldg(p00 + ch) : 0.0f, i01 = v01 ? __ldg(p01 + ch) : 0.0f;
                    const float i10 = v10 ? __ldg(p10 + ch) : 0.0f, i11 = v11 ? __ldg(p11 + ch) : 0.0f;
                    o[ch] = zp_blend(c, i00, i01, i10, i11);
                }
            }
        }
    }

    // ---- phase 4 (STAGED): output tile -> global with bulk async stores --------------------
    if (STAGED) {
        fence_proxy_async_smem();
        __syncthreads();
        const int vcols = min(TW, ow - col0);
        if (tid < TH && row0 + tid < oh) {
            bulk_s2g(p.out + (((size_t)b * oh + row0 + tid) * ow + col0) * 3, smem_u32(s_out + tid * TW * 3), (unsigned)vcols * 12u);
            bulk_commit();
            bulk_wait_read0();
        }
    }
}

// ---- spatial_transformer._meshgrid (spatial_transformer.py:460-482) --------------------------
__global__ void st_meshgrid_kernel(float* __restrict__ grid, int oh, int ow, float step_x, float step_y) {
    const int n = oh * ow;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        grid[i] = lin_coord(i % ow, step_x);
        grid[n + i] = lin_coord(i / ow, step_y);
        grid[2 * n + i] = 1.0f;
    }
}

// ---- host launchers ----------------------------------------------------------------------------
static float lin_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

static int g_src_smem_bytes = 28 * 1024;   // staging buffer per CTA; tunable for experiments
static int g_pack = 1;

// fast path (warp_fwd_strip.cu)
bool strip_path_ok(const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn_or_0);
int strip_tps(const float* U, const float* coord, long long cstride, const float* T, float* out, float* x_out, float* y_out,
              float* mask_out, int B, int H, int W, int oh, int ow, int pn, cudaStream_t st);
int strip_given(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int oh, int ow, cudaStream_t st);
int strip_flow(const float* im, const float* flow, float* out, int B, int H, int W, cudaStream_t st);
int strip_homog(const float* im, const float* theta, int projective, float* out, float* x_out, float* y_out, int B, int H, int W,
                int oh, int ow, cudaStream_t st);
void strip_set_tuning(int smem_bytes, int pack, int target_ctas, int pipe);

// fast path (warp_fwd_tile.cu): warp-autonomous 32x8 tiles
bool tile_path_ok(const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn_or_0);
int tile_tps(const float* U, const float* coord, long long cstride, const float* T, float* out, float* x_out, float* y_out,
             float* mask_out, int B, int H, int W, int oh, int ow, int pn, cudaStream_t st);
int tile_given(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int oh, int ow, cudaStream_t st);
int tile_flow(const float* im, const float* flow, float* out, int B, int H, int W, cudaStream_t st);
int tile_homog(const float* im, const float* theta, int projective, float* out, float* x_out, float* y_out, int B, int H, int W,
               int oh, int ow, cudaStream_t st);
void tile_set_tuning(int stage_bytes, int target_ctas, int minb);

constexpr int FLAG_LEGACY_STAGED = 2;   // experiments only: the non-pipelined one-tile-per-CTA staged kernel
constexpr int FLAG_STRIP = 4;           // experiments only: the CTA-synchronous strip kernel (warp_fwd_strip.cu)

static bool use_tile(int flags, const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn) {
    return !(flags & (DVSG_FLAG_FORCE_DIRECT | FLAG_LEGACY_STAGED | FLAG_STRIP)) && tile_path_ok(src, out, H, W, C, oh, ow, pn);
}

static bool use_strip(int flags, const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn) {
    return (flags & FLAG_STRIP) && strip_path_ok(src, out, H, W, C, oh, ow, pn);
}

template <int MODE>
static int launch_fwd(FwdParams p, int flags, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    DVSG_REQUIRE(p.B <= 65535, "batch %d exceeds the grid z limit 65535: split the call", p.B);
    const bool staged = (flags & FLAG_LEGACY_STAGED) && p.C == 3 && p.W % 4 == 0 && p.ow % 4 == 0 &&
                        aligned16(p.src) && aligned16(p.out);
    p.kc_cap = MODE == MODE_TPS ? (p.pn < KC ? p.pn : KC) : 0;
    p.src_smem_bytes = staged ? g_src_smem_bytes : 0;
    const size_t tps_bytes = (size_t)p.kc_cap * (sizeof(float4) + TH * sizeof(float));
    const size_t smem = (staged ? (size_t)TH * TW * 3 * sizeof(float) + p.src_smem_bytes : 0) + tps_bytes;
    dim3 grid((p.ow + TW - 1) / TW, (p.oh + TH - 1) / TH, p.B);
    DVSG_REQUIRE(grid.y <= 65535, "output height %d too large", p.oh);
    const bool pack = g_pack && MODE == MODE_TPS;
#define DVSG_LAUNCH(ST, PK)                                                                                      \
    do {                                                                                                         \
        auto k = warp_fwd_kernel<MODE, ST, PK>;                                                                  \
        if (smem > 40 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        k<<<grid, NT, smem, st>>>(p);                                                                            \
    } while (0)
    if (staged) { if (pack) DVSG_LAUNCH(true, true); else DVSG_LAUNCH(true, false); }
    else        { if (pack) DVSG_LAUNCH(false, true); else DVSG_LAUNCH(false, false); }
#undef DVSG_LAUNCH
    count_launch();
    return check_launch("warp_fwd_kernel");
}

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_set_tuning(int src_smem_bytes, int pack) {
    if (src_smem_bytes >= 0) g_src_smem_bytes = src_smem_bytes & ~15;
    if (pack >= 0) g_pack = pack;
    strip_set_tuning(src_smem_bytes, pack, -1, -1);
    return DVSG_OK;
}

extern "C" int dvsg_set_tile_tuning(int stage_bytes, int target_ctas, int min_ctas_per_sm) {
    tile_set_tuning(stage_bytes, target_ctas, min_ctas_per_sm);
    return DVSG_OK;
}

extern "C" int dvsg_set_strip_tuning(int target_ctas, int pipe) {
    strip_set_tuning(-1, -1, target_ctas, pipe);
    return DVSG_OK;
}

extern "C" int dvsg_tps_warp_fwd(const float* U, const float* coord, long long coord_batch_stride, const float* T,
                                 float* out, float* x_out, float* y_out, float* mask_out, int B, int H, int W, int C,
                                 int oh, int ow, int pn, int flags, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0 && pn > 0, "tps_warp_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (U && coord && T && out), "tps_warp_fwd: null pointer");
    DVSG_REQUIRE((x_out == nullptr) == (y_out == nullptr), "tps_warp_fwd: x_out and y_out must be given together");
    DVSG_REQUIRE(coord_batch_stride == 0 || coord_batch_stride >= 2LL * pn, "tps_warp_fwd: coord stride %lld < 2*pn", coord_batch_stride);
    DVSG_REQUIRE((long long)H * W < (1LL << 31) / C && (long long)oh * ow < (1LL << 31) / C, "tps_warp_fwd: frame too large for int32 indexing");
    if (use_tile(flags, U, out, H, W, C, oh, ow, pn))
        return B == 0 ? DVSG_OK : tile_tps(U, coord, coord_batch_stride, T, out, x_out, y_out, mask_out, B, H, W, oh, ow, pn, (cudaStream_t)stream);
    if (use_strip(flags, U, out, H, W, C, oh, ow, pn))
        return B == 0 ? DVSG_OK : strip_tps(U, coord, coord_batch_stride, T, out, x_out, y_out, mask_out, B, H, W, oh, ow, pn, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = U; p.out = out; p.x_out = x_out; p.y_out = y_out; p.mask_out = mask_out;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = coord_batch_stride; p.T = T; p.pn = pn;
    p.step_x = lin_step(ow); p.step_y = lin_step(oh);
    return launch_fwd<MODE_TPS>(p, flags, (cudaStream_t)stream);
}

extern "C" int dvsg_bilinear_fwd(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int C,
                                 int oh, int ow, int flags, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0, "bilinear_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && x && y && out), "bilinear_fwd: null pointer");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "bilinear_fwd: frame too large for int32 indexing");
    if (use_tile(flags, im, out, H, W, C, oh, ow, 0))
        return B == 0 ? DVSG_OK : tile_given(im, x, y, out, B, H, W, oh, ow, (cudaStream_t)stream);
    if (use_strip(flags, im, out, H, W, C, oh, ow, 0))
        return B == 0 ? DVSG_OK : strip_given(im, x, y, out, B, H, W, oh, ow, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = im; p.out = out; p.x_in = x; p.y_in = y;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    return launch_fwd<MODE_GIVEN>(p, flags, (cudaStream_t)stream);
}

extern "C" int dvsg_flow_warp_fwd(const float* im, const float* flow, float* out, int B, int H, int W, int C, int flags,
                                  void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, "flow_warp_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && flow && out), "flow_warp_fwd: null pointer");
    DVSG_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7u) == 0, "flow_warp_fwd: flow must be 8-byte aligned");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "flow_warp_fwd: frame too large for int32 indexing");
    if (use_tile(flags, im, out, H, W, C, H, W, 0))
        return B == 0 ? DVSG_OK : tile_flow(im, flow, out, B, H, W, (cudaStream_t)stream);
    if (use_strip(flags, im, out, H, W, C, H, W, 0))
        return B == 0 ? DVSG_OK : strip_flow(im, flow, out, B, H, W, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = im; p.out = out; p.flow = flow;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = H; p.ow = W;
    return launch_fwd<MODE_FLOW>(p, flags, (cudaStream_t)stream);
}

extern "C" int dvsg_homography_warp_fwd(const float* im, const float* theta, int projective, float* out, float* x_out,
                                        float* y_out, int B, int H, int W, int C, int oh, int ow, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0, "homography_warp_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && theta && out), "homography_warp_fwd: null pointer");
    DVSG_REQUIRE((x_out == nullptr) == (y_out == nullptr), "homography_warp_fwd: x_out and y_out must be given together");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "homography_warp_fwd: frame too large for int32 indexing");
    if (use_tile(0, im, out, H, W, C, oh, ow, 0))
        return B == 0 ? DVSG_OK : tile_homog(im, theta, projective, out, x_out, y_out, B, H, W, oh, ow, (cudaStream_t)stream);
    if (false)
        return B == 0 ? DVSG_OK : strip_homog(im, theta, projective, out, x_out, y_out, B, H, W, oh, ow, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = im; p.out = out; p.x_out = x_out; p.y_out = y_out; p.theta = theta; p.projective = projective;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    p.step_x = lin_step(ow); p.step_y = lin_step(oh);
    return launch_fwd<MODE_HOMOG>(p, 0, (cudaStream_t)stream);
}

extern "C" int dvsg_st_meshgrid(float* grid, int oh, int ow, void* stream) {
    DVSG_REQUIRE(oh > 0 && ow > 0 && grid, "st_meshgrid: bad argument");
    const int n = oh * ow;
    st_meshgrid_kernel<<<(n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184, 256, 0, (cudaStream_t)stream>>>(grid, oh, ow, lin_step(ow), lin_step(oh));
    count_launch();
    return check_launch("st_meshgrid_kernel");
}
