// warp_fwd.cu -- C-ABI entry points of the forward warp and the GENERIC forward kernel.
//
// One kernel template covers the four coordinate sources of the path
//   MODE_TPS    fused TPS grid generation + A4 sampler   (ThinPlateSpline.py:92-141, 30-90)
//   MODE_GIVEN  bilinear_interp on caller-given x, y      (spatial_transformer.py:496-563)
//   MODE_FLOW   tf_warp                                   (warp_with_optical_flow.py:96-176)
//   MODE_HOMOG  Projective/AffineTransformer + bilinear   (spatial_transformer.py:73-91, 423-452)
// for ANY channel count, alignment and frame size: each CTA owns a 64x16 output tile (one
// thread per column, four rows per thread), computes the sampling coordinates and gathers
// the four corners straight from global memory.  Frames with C == 3 and 16-byte aligned rows
// -- every configuration the path is measured on -- take the warp-autonomous TMA-staged
// kernel of warp_fwd_tile.cu instead; both perform the same arithmetic in the same order, so
// their results are identical bit for bit (tests/test_gpu_forward.py).
//
// HBM traffic (algorithmic): 4C B read + 4C B written per output pixel (+8 B if x, y are
// materialised, +8 B flow read for MODE_FLOW, +8 B x,y read for MODE_GIVEN).
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>

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
};

// Accumulate sum_k c_k * d2_k * ln(d2_k) (the radial term as restated in dvsg_common.cuh) for PR consecutive rows of one column in packed fp32x2
// (two rows per instruction).  pt[k] = (px, py, cx*ln2, cy*ln2); dy2[k*TH + r] = (y_t(row0+r) - py_k)^2.
// Same operations in the same order as the tile kernel, so both produce identical coordinates.
__device__ __forceinline__ void tps_accumulate(const float4* __restrict__ pt, const float* __restrict__ dy2, int kc,
                                               float xt, int rbase, float (&xs)[PR], float (&ys)[PR]) {
    float2 xa = make_float2(xs[0], xs[1]), xb = make_float2(xs[2], xs[3]);
    float2 ya = make_float2(ys[0], ys[1]), yb = make_float2(ys[2], ys[3]);
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
        const float4 q = pt[k];
        const float dx = DVSG_SUB(xt, q.x);
        const float dx2 = DVSG_MUL(dx, dx);
        const float4 d = *reinterpret_cast<const float4*>(dy2 + k * TH + rbase);
        const float2 dxx = make_float2(dx2, dx2);
        const float2 d2a = __fadd2_rn(dxx, make_float2(d.x, d.y));
        const float2 d2b = __fadd2_rn(dxx, make_float2(d.z, d.w));
        const float2 ra = __fmul2_rn(d2a, make_float2(lg2_approx(d2a.x), lg2_approx(d2a.y)));
        const float2 rb = __fmul2_rn(d2b, make_float2(lg2_approx(d2b.x), lg2_approx(d2b.y)));
        const float2 cx = make_float2(q.z, q.z), cy = make_float2(q.w, q.w);
        xa = __ffma2_rn(cx, ra, xa);
        xb = __ffma2_rn(cx, rb, xb);
        ya = __ffma2_rn(cy, ra, ya);
        yb = __ffma2_rn(cy, rb, yb);
    }
    xs[0] = xa.x; xs[1] = xa.y; xs[2] = xb.x; xs[3] = xb.y;
    ys[0] = ya.x; ys[1] = ya.y; ys[2] = yb.x; ys[3] = yb.y;
}

template <int MODE>
__device__ __forceinline__ Corners corners_of(float xs, float ys, int W, int H) {
    if (MODE == MODE_TPS) return a4_corners(xs, ys, W, H);
    return zp_corners(xs, ys, W, H);   // xs, ys already pixel-space for the ZP modes
}

template <int MODE>
__global__ void __launch_bounds__(NT, 3) warp_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float s_lin[12];    // affine TPS coefficients or the 3x3 homography

    float4* s_pt = reinterpret_cast<float4*>(smem);
    float* s_dy2 = reinterpret_cast<float*>(s_pt + p.kc_cap);

    const int tid = threadIdx.x;
    const int tx = tid & (TW - 1);
    const int rg = tid >> 6;
    const int col0 = blockIdx.x * TW, row0 = blockIdx.y * TH, b = blockIdx.z;
    const int col = col0 + tx;
    const int rbase = rg * PR;
    const int H = p.H, W = p.W, C = p.C, oh = p.oh, ow = p.ow;
    const bool col_ok = col < ow;

    // ---- phase 1: sampling coordinates for PR rows of this column --------------------
    // MODE_TPS: xs, ys normalised; other modes: pixel-space (before the clip).
    float xs[PR], ys[PR];
    if (MODE == MODE_TPS) {
        const int N = p.pn + 3;
        const float* Tb = p.T + (size_t)b * 2 * N;
        const float* cb = p.coord + (size_t)b * p.coord_stride;
        if (tid < 64) {      // warps 0, 1: affine rows of x_s, y_s (constant with the folded epsilon term, x, y)
            const int w = tid >> 5, l = tid & 31;
            const float c0 = tps_affine0(Tb + w * N, p.pn, l);
            if (l == 0) s_lin[3 * w] = c0;
            else if (l < 3) s_lin[3 * w + l] = __ldg(Tb + w * N + l);
        }
        for (int k0 = 0; k0 < p.pn; k0 += p.kc_cap) {
            const int kc = min(p.kc_cap, p.pn - k0);
            if (k0 > 0) __syncthreads();
            for (int k = tid; k < kc; k += NT)
                s_pt[k] = make_float4(__ldg(cb + 2 * (k0 + k)), __ldg(cb + 2 * (k0 + k) + 1),
                                      __ldg(Tb + 3 + k0 + k) * LN2, __ldg(Tb + N + 3 + k0 + k) * LN2);
            for (int i = tid; i < kc * TH; i += NT) {
                const int k = i / TH, r = i % TH;
                s_dy2[i] = tps_dy2(lin_coord(row0 + r, p.step_y), __ldg(cb + 2 * (k0 + k) + 1));
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
            tps_accumulate(s_pt, s_dy2, kc, xt, rbase, xs, ys);
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

    // ---- phase 2: gather, blend, store ---------------------------------------------------
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

// ---- wide pixels (C even, C != 3): lanes over (pixel, channel pair) --------------------------------------------------
// bilinear_interp's only live call sites carry C = 18 (model.py:156-167): a pixel is 72 contiguous bytes.  The
// per-column mapping of warp_fwd_kernel reads 4-byte words 72 bytes apart (25 % of the HBM roofline at C = 18), and a
// thread per (pixel, channel pair) that recomputes the pixel's coordinates is instruction-bound (243 instructions per
// 8 output bytes with the projective grid's two IEEE divisions: 29 %).  So the CTA works in two phases on a tile of
// WIDE_PX x WIDE_ROWS output pixels: (1) one thread per pixel computes the sampling coordinate, the four corner offsets
// (-1 = outside the frame, i.e. the zero padding) and the four weights once -- same operations in the same order as
// the other kernels -- into a 32-byte shared-memory record; (2) all threads run over (pixel, channel pair): two
// LDS.128, four coalesced LDG.64, the blend in the reference's add_n order, one coalesced STG.64.  Identical bits;
// 59-62 % of the HBM roofline at C = 18 (288x512x32 and 1080p x 8), 2.4x the per-column kernel.
constexpr int WIDE_PX = 64, WIDE_ROWS = 4, WIDE_NT = 256;
struct __align__(16) WideRec { int o00, o01, o10, o11; float w00, w01, w10, w11; };

template <int MODE>
__global__ void __launch_bounds__(WIDE_NT) warp_fwd_wide_kernel(const FwdParams p) {
    __shared__ WideRec s_rec[WIDE_ROWS * WIDE_PX];
    const int H = p.H, W = p.W, C = p.C, oh = p.oh, ow = p.ow, cv = C >> 1;
    const int tid = threadIdx.x, col0 = blockIdx.x * WIDE_PX, row0 = blockIdx.y * WIDE_ROWS, b = blockIdx.z;
    // ---- phase 1: one record per pixel ---------------------------------------------------------------
    {
        const int r = tid / WIDE_PX, col = col0 + tid % WIDE_PX, row = row0 + r;     // WIDE_NT == WIDE_ROWS * WIDE_PX
        WideRec rec = {-1, -1, -1, -1, 0.0f, 0.0f, 0.0f, 0.0f};
        if (col < ow && row < oh) {
            const size_t pix = ((size_t)b * oh + row) * ow + col;
            float xq, yq;      // pixel-space coordinate before the clip
            if (MODE == MODE_GIVEN) {
                xq = zp_pix_from_norm(__ldg(p.x_in + pix), W);
                yq = zp_pix_from_norm(__ldg(p.y_in + pix), H);
            } else if (MODE == MODE_FLOW) {
                const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + pix);
                xq = DVSG_ADD((float)col, f.x);   // warp_with_optical_flow.py:107-120
                yq = DVSG_ADD((float)row, f.y);
            } else {
                const int nt = p.projective ? 8 : 6;
                const float* th = p.theta + (size_t)b * nt;
                const float xt = lin_coord(col, p.step_x), yt = lin_coord(row, p.step_y);
                // rows of theta @ [x_t; y_t; 1], accumulated k = 0,1,2 (spatial_transformer.py:437)
                float xn = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(th + 0), xt), DVSG_MUL(__ldg(th + 1), yt)), __ldg(th + 2));
                float yn = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(th + 3), xt), DVSG_MUL(__ldg(th + 4), yt)), __ldg(th + 5));
                if (p.projective) {
                    const float zn = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(th + 6), xt), DVSG_MUL(__ldg(th + 7), yt)), 1.0f);
                    xn = zn != 0.0f ? DVSG_DIV(xn, zn) : 0.0f;   // tf.div_no_nan, :446-447
                    yn = zn != 0.0f ? DVSG_DIV(yn, zn) : 0.0f;
                }
                if (p.x_out) { p.x_out[pix] = xn; p.y_out[pix] = yn; }
                xq = zp_pix_from_norm(xn, W);
                yq = zp_pix_from_norm(yn, H);
            }
            const Corners c = zp_corners(xq, yq, W, H);
            const bool vx0 = zp_valid(c.x0, W), vx1 = zp_valid(c.x1, W), vy0 = zp_valid(c.y0, H), vy1 = zp_valid(c.y1, H);
            // offsets of the corners inside the frame in channel pairs (H*W*C < 2^31); padded index - 1 = frame index
            rec.o00 = (vx0 && vy0) ? ((c.y0 - 1) * W + (c.x0 - 1)) * cv : -1;
            rec.o01 = (vx1 && vy0) ? ((c.y0 - 1) * W + (c.x1 - 1)) * cv : -1;
            rec.o10 = (vx0 && vy1) ? ((c.y1 - 1) * W + (c.x0 - 1)) * cv : -1;
            rec.o11 = (vx1 && vy1) ? ((c.y1 - 1) * W + (c.x1 - 1)) * cv : -1;
            rec.w00 = DVSG_MUL(c.ax1, c.ay1); rec.w01 = DVSG_MUL(c.ax0, c.ay1);       // as zp_blend forms them
            rec.w10 = DVSG_MUL(c.ax1, c.ay0); rec.w11 = DVSG_MUL(c.ax0, c.ay0);
        }
        s_rec[tid] = rec;
    }
    __syncthreads();
    // ---- phase 2: (pixel, channel pair) -----------------------------------------------------------------
    const float2* srcb = reinterpret_cast<const float2*>(p.src + (size_t)b * H * W * C);
    const int npx = min(WIDE_PX, ow - col0), per_row = npx * cv;
    const float2 z = make_float2(0.0f, 0.0f);
    const int px_first = tid / cv, c2_first = tid - px_first * cv, dpx = WIDE_NT / cv, dc2 = WIDE_NT - dpx * cv;   // e -> (px, c2) without a division per element
    for (int r = 0; r < WIDE_ROWS && row0 + r < oh; ++r) {
        float2* orow = reinterpret_cast<float2*>(p.out) + (((size_t)b * oh + row0 + r) * ow + col0) * cv;
        int px = px_first, c2 = c2_first;
#pragma unroll 3
        for (int e = tid; e < per_row; e += WIDE_NT, px += dpx, c2 += dc2) {
            if (c2 >= cv) { c2 -= cv; ++px; }
            const WideRec rec = s_rec[r * WIDE_PX + px];
            const float2 i00 = rec.o00 >= 0 ? __ldg(srcb + (rec.o00 + c2)) : z;
            const float2 i01 = rec.o01 >= 0 ? __ldg(srcb + (rec.o01 + c2)) : z;
            const float2 i10 = rec.o10 >= 0 ? __ldg(srcb + (rec.o10 + c2)) : z;
            const float2 i11 = rec.o11 >= 0 ? __ldg(srcb + (rec.o11 + c2)) : z;
            // add_n([w00*I00, w01*I01, w10*I10, w11*I11]) left to right (spatial_transformer.py:557-562)
            float2 o;
            o.x = DVSG_ADD(DVSG_ADD(DVSG_ADD(DVSG_MUL(rec.w00, i00.x), DVSG_MUL(rec.w01, i01.x)), DVSG_MUL(rec.w10, i10.x)), DVSG_MUL(rec.w11, i11.x));
            o.y = DVSG_ADD(DVSG_ADD(DVSG_ADD(DVSG_MUL(rec.w00, i00.y), DVSG_MUL(rec.w01, i01.y)), DVSG_MUL(rec.w10, i10.y)), DVSG_MUL(rec.w11, i11.y));
            orow[e] = o;
        }
    }
}

// ---- wide pixels, TMA-staged (round 2) ----------------------------------------------------------------------------
// The kernel above gathers its four corners straight from global memory and spends 78 instructions per (pixel, channel
// pair): it is ISSUE-bound (80 % issue-slot utilisation) at 0.60 of the HBM roofline.  Here the CTA (128 threads, a
// 16 x 8-pixel output tile) stages the tile's source FOOTPRINT in shared memory with two side-by-side 3-D TMA tensor
// copies ([B][H][W*C] floats; a box is at most 256 elements wide = 14 pixels of 18 channels), zero-filled outside the
// frame -- the zero padding of bilinear_interp / tf_warp -- and every thread then owns ONE pixel:
//   1  sampling coordinate, padded corners and weights (same operations, same order as everywhere else);
//   2  CTA bounding box of the corners (REDUX per warp, 4 x 4 words of shared memory), origin aligned to 16 bytes;
//   3  thread 0 issues the two TMA copies; every thread turns its corners into shared-memory offsets (registers, no records);
//   4  per channel pair: four LDS.64 (lanes 72 bytes apart: conflict-free per half-warp at C = 18), the blend in add_n order
//      in packed fp32x2, one STS.64 into the output tile; 15 instructions per pair;
//   5  the output tile leaves with two TMA tensor stores (8 pixels x 8 rows each), clipped at the frame edge by the hardware.
// A tile whose footprint does not fit 2 boxes x 12 rows gathers from global memory instead (same arithmetic).
// Tried on the way (C = 18, 32 x 288 x 512): (pixel, channel pair) work items with per-pixel records in shared memory and
// direct STG.64 -- 0.58 of HBM as first written, 0.68 with 32-bit indexing and a packed blend; three pairs per item 0.64.
constexpr int WS_PX = 16, WS_ROWS = 8, WS_NT = WS_PX * WS_ROWS, WS_BOX_ROWS = 12, WS_OBOX_PX = 8;

__device__ __forceinline__ void ws_tma_load_3d(uint32_t dst_smem, const void* tmap, int x, int y, int z, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
}
__device__ __forceinline__ void ws_tma_store_3d(const void* tmap, int x, int y, int z, uint32_t src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(z), "r"(src_smem) : "memory");
}
struct alignas(64) WideMaps { CUtensorMap src, out; };

// The CTA walks `seg_len` tiles of a strip as a software pipeline (the chain coordinates -> footprint -> TMA copy -> gather ->
// TMA store of ONE tile is ~4 us of mostly latency: one tile per CTA reached 0.66 of HBM): while tile t is gathered, the
// copy of tile t+1's footprint is in flight into the other staging buffer and the coordinates of tile t+2 are being loaded.
template <int MODE>
__global__ void __launch_bounds__(WS_NT) warp_fwd_wide_staged_kernel(const FwdParams p, const __grid_constant__ WideMaps maps, const int pxb,
                                                                     const int pal_shift, const int stage_bytes, const int seg_len, const int box_rows, const float one) {
    extern __shared__ __align__(128) unsigned char ws_smem[];      // 2 x [2 boxes][WS_BOX_ROWS][pxb * C] floats, then the output tile
    __shared__ int s_bb[2][4][4];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    const int H = p.H, W = p.W, C = p.C, oh = p.oh, ow = p.ow, cv = C >> 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.y * WS_ROWS, b = blockIdx.z;
    const int n_tx = (ow + WS_PX - 1) / WS_PX, t_begin = blockIdx.x * seg_len, t_end = min(t_begin + seg_len, n_tx);
    if (tid == 0) { mbar_init(smem_u32(&s_mbar[0]), 1); mbar_init(smem_u32(&s_mbar[1]), 1); fence_mbar_init(); }
    const int pr = tid / WS_PX, pc = tid % WS_PX;
    const int row = min(row0 + pr, oh - 1);
    const bool row_live = row0 + pr < oh;
    const int box_f2 = box_rows * pxb * cv;         // float2 elements per box
    const int box1_f2 = (box_f2 + 15) & ~15;        // the second box starts at the next multiple of 128 bytes (TMA destination alignment)
    const float2 one2 = make_float2(one, one);      // 1.0f from the kernel arguments: a + b as fma(a, 1, b), rounded once, not contracted
    // this thread's pixel in the output tile: two boxes of 8 pixels x 8 rows, each dense [row][8 * C floats]
    float2* const otile = reinterpret_cast<float2*>(ws_smem + 2 * stage_bytes);
    float2* const o = otile + (pc / WS_OBOX_PX) * (WS_ROWS * WS_OBOX_PX * cv) + (pr * WS_OBOX_PX + pc % WS_OBOX_PX) * cv;
    const float2* __restrict__ srcb = reinterpret_cast<const float2*>(p.src + (size_t)b * H * W * C);

    // raw per-pixel inputs of a tile (GIVEN: x, y; FLOW: dx, dy): issued two tiles ahead of their use
    auto load_raw = [&](const int t, float& rx, float& ry) {
        if (MODE == MODE_GIVEN || MODE == MODE_FLOW) {
            const int col = min(t * WS_PX + pc, ow - 1);
            const size_t pix = ((size_t)b * oh + row) * ow + col;
            if (MODE == MODE_GIVEN) { rx = __ldg(p.x_in + pix); ry = __ldg(p.y_in + pix); }
            else { const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + pix); rx = f.x; ry = f.y; }
        }
    };
    // 1: coordinate, padded corners and weights of this thread's pixel of tile t (pixels past the frame edge repeat the edge pixel)
    auto corners_of_tile = [&](const int t, const float rx, const float ry) -> Corners {
        const int col = min(t * WS_PX + pc, ow - 1);
        float xq, yq;      // pixel-space coordinate before the clip
        if (MODE == MODE_GIVEN) {
            xq = zp_pix_from_norm(rx, W);
            yq = zp_pix_from_norm(ry, H);
        } else if (MODE == MODE_FLOW) {
            xq = DVSG_ADD((float)col, rx);   // warp_with_optical_flow.py:107-120
            yq = DVSG_ADD((float)row, ry);
        } else {
            const int nt = p.projective ? 8 : 6;
            const float* th = p.theta + (size_t)b * nt;
            const float xt = lin_coord(col, p.step_x), yt = lin_coord(row, p.step_y);
            float xn = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(th + 0), xt), DVSG_MUL(__ldg(th + 1), yt)), __ldg(th + 2));
            float yn = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(th + 3), xt), DVSG_MUL(__ldg(th + 4), yt)), __ldg(th + 5));
            if (p.projective) {
                const float zn = DVSG_ADD(DVSG_ADD(DVSG_MUL(__ldg(th + 6), xt), DVSG_MUL(__ldg(th + 7), yt)), 1.0f);
                xn = zn != 0.0f ? DVSG_DIV(xn, zn) : 0.0f;   // tf.div_no_nan, :446-447
                yn = zn != 0.0f ? DVSG_DIV(yn, zn) : 0.0f;
            }
            if (p.x_out && row_live && t * WS_PX + pc < ow) {
                const size_t pix = ((size_t)b * oh + row) * ow + col;
                p.x_out[pix] = xn; p.y_out[pix] = yn;
            }
            xq = zp_pix_from_norm(xn, W);
            yq = zp_pix_from_norm(yn, H);
        }
        return zp_corners(xq, yq, W, H);      // padded corners in [0, W+1] x [0, H+1]; frame index = padded index - 1
    };
    // 2a: per-warp bounding box of the tile's corners (frame coordinates; the padding ring is part of it) -> s_bb[sel]
    auto bbox_post = [&](const Corners& c, const int sel) {
        const int xl = __reduce_min_sync(0xffffffffu, c.x0 - 1), xh = __reduce_max_sync(0xffffffffu, c.x1 - 1);
        const int yl = __reduce_min_sync(0xffffffffu, c.y0 - 1), yh = __reduce_max_sync(0xffffffffu, c.y1 - 1);
        if (lane == 0) { s_bb[sel][warp][0] = xl; s_bb[sel][warp][1] = xh; s_bb[sel][warp][2] = yl; s_bb[sel][warp][3] = yh; }
    };
    // 2b + 3 (after a CTA barrier): footprint origin, does it fit, thread 0 issues the two TMA copies into staging buffer `sel`
    auto footprint_issue = [&](const int sel, int& fx0, int& y_lo) -> bool {
        const int x_lo = min(min(s_bb[sel][0][0], s_bb[sel][1][0]), min(s_bb[sel][2][0], s_bb[sel][3][0]));
        const int x_hi = max(max(s_bb[sel][0][1], s_bb[sel][1][1]), max(s_bb[sel][2][1], s_bb[sel][3][1]));
        y_lo = min(min(s_bb[sel][0][2], s_bb[sel][1][2]), min(s_bb[sel][2][2], s_bb[sel][3][2]));
        const int y_hi = max(max(s_bb[sel][0][3], s_bb[sel][1][3]), max(s_bb[sel][2][3], s_bb[sel][3][3]));
        // origin aligned down to 2^pal_shift pixels (that many pixels of C floats are a multiple of 16 bytes: TMA box origins must be 16-byte aligned)
        fx0 = (x_lo >> pal_shift) << pal_shift;      // arithmetic shift: rounds toward -infinity
        const bool staged = x_hi - fx0 + 1 <= 2 * pxb && y_hi - y_lo + 1 <= box_rows;
        if (staged && tid == 0) {
            const uint32_t mb = smem_u32(&s_mbar[sel]), dst = smem_u32(ws_smem) + (unsigned)(sel * stage_bytes);
            mbar_arrive_expect_tx(mb, (unsigned)(2 * box_f2 * 8));
            ws_tma_load_3d(dst, &maps.src, fx0 * C, y_lo, b, mb);
            ws_tma_load_3d(dst + (unsigned)(box1_f2 * 8), &maps.src, (fx0 + pxb) * C, y_lo, b, mb);
        }
        return staged;
    };

    if (t_begin >= t_end) return;
    float rx1 = 0.0f, ry1 = 0.0f, rx2 = 0.0f, ry2 = 0.0f;      // raw inputs of tiles t+1, t+2
    Corners cur, nxt = {};
    bool cur_staged, nxt_staged = false;
    int cur_fx0, cur_ylo, nxt_fx0 = 0, nxt_ylo = 0;
    unsigned ph0 = 0u, ph1 = 0u;      // mbarrier parities of the two staging buffers
    {   // prologue: tile t_begin's footprint on its way, tile t_begin + 1's raw inputs loading
        float rx0 = 0.0f, ry0 = 0.0f;
        load_raw(t_begin, rx0, ry0);
        if (t_begin + 1 < t_end) load_raw(t_begin + 1, rx1, ry1);
        cur = corners_of_tile(t_begin, rx0, ry0);
        bbox_post(cur, t_begin & 1);
        __syncthreads();
        cur_staged = footprint_issue(t_begin & 1, cur_fx0, cur_ylo);
    }
    for (int t = t_begin; t < t_end; ++t) {
        const int sel = t & 1;
        const bool has_next = t + 1 < t_end;
        if (t + 2 < t_end) load_raw(t + 2, rx2, ry2);
        if (has_next) {
            nxt = corners_of_tile(t + 1, rx1, ry1);
            bbox_post(nxt, sel ^ 1);
        }
        if (tid == 0 && t > t_begin) bulk_wait_read0();      // the previous tile's stores have read the output tile
        __syncthreads();                                     // s_bb of tile t+1 visible; output tile and the other staging buffer free
        if (has_next) nxt_staged = footprint_issue(sel ^ 1, nxt_fx0, nxt_ylo);
        // ---- 4: per channel pair: gather, blend, stage --------------------------------------------------------------
        const float2 w00 = make_float2(DVSG_MUL(cur.ax1, cur.ay1), DVSG_MUL(cur.ax1, cur.ay1)), w01 = make_float2(DVSG_MUL(cur.ax0, cur.ay1), DVSG_MUL(cur.ax0, cur.ay1));
        const float2 w10 = make_float2(DVSG_MUL(cur.ax1, cur.ay0), DVSG_MUL(cur.ax1, cur.ay0)), w11 = make_float2(DVSG_MUL(cur.ax0, cur.ay0), DVSG_MUL(cur.ax0, cur.ay0));
        if (cur_staged) {
            // offsets in float2 units inside the staging buffer; corners in the padding ring read the TMA's zero fill
            const int xa = cur.x0 - 1 - cur_fx0, xb = cur.x1 - 1 - cur_fx0, ya = cur.y0 - 1 - cur_ylo, yb = cur.y1 - 1 - cur_ylo;
            const int oa = xa >= pxb ? box1_f2 + (xa - pxb) * cv : xa * cv, ob = xb >= pxb ? box1_f2 + (xb - pxb) * cv : xb * cv;
            const float2* __restrict__ stage = reinterpret_cast<const float2*>(ws_smem + sel * stage_bytes);
            const float2 *__restrict__ p00 = stage + ya * pxb * cv + oa, *__restrict__ p01 = stage + ya * pxb * cv + ob;
            const float2 *__restrict__ p10 = stage + yb * pxb * cv + oa, *__restrict__ p11 = stage + yb * pxb * cv + ob;
            mbar_wait(smem_u32(&s_mbar[sel]), sel ? ph1 : ph0);
            if (sel) ph1 ^= 1u; else ph0 ^= 1u;
#pragma unroll 3
            for (int k = 0; k < cv; ++k) {
                // add_n([w00*I00, w01*I01, w10*I10, w11*I11]) left to right (spatial_transformer.py:557-562), both channels of the pair at once
                const float2 t00 = __fmul2_rn(w00, p00[k]), t01 = __fmul2_rn(w01, p01[k]), t10 = __fmul2_rn(w10, p10[k]), t11 = __fmul2_rn(w11, p11[k]);
                o[k] = __ffma2_rn(__ffma2_rn(__ffma2_rn(t00, one2, t01), one2, t10), one2, t11);
            }
        } else {
            const bool vx0 = zp_valid(cur.x0, W), vx1 = zp_valid(cur.x1, W), vy0 = zp_valid(cur.y0, H), vy1 = zp_valid(cur.y1, H);
            const float2 z = make_float2(0.0f, 0.0f);
            const bool v00 = vx0 && vy0, v01 = vx1 && vy0, v10 = vx0 && vy1, v11 = vx1 && vy1;
            const float2 *p00 = srcb + (v00 ? ((cur.y0 - 1) * W + (cur.x0 - 1)) * cv : 0), *p01 = srcb + (v01 ? ((cur.y0 - 1) * W + (cur.x1 - 1)) * cv : 0);
            const float2 *p10 = srcb + (v10 ? ((cur.y1 - 1) * W + (cur.x0 - 1)) * cv : 0), *p11 = srcb + (v11 ? ((cur.y1 - 1) * W + (cur.x1 - 1)) * cv : 0);
            for (int k = 0; k < cv; ++k) {
                const float2 t00 = __fmul2_rn(w00, v00 ? __ldg(p00 + k) : z), t01 = __fmul2_rn(w01, v01 ? __ldg(p01 + k) : z);
                const float2 t10 = __fmul2_rn(w10, v10 ? __ldg(p10 + k) : z), t11 = __fmul2_rn(w11, v11 ? __ldg(p11 + k) : z);
                o[k] = __ffma2_rn(__ffma2_rn(__ffma2_rn(t00, one2, t01), one2, t10), one2, t11);
            }
        }
        // ---- 5: output tile -> global with two TMA tensor stores (clipped at the frame edge) ------------------------------
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const uint32_t os = smem_u32(otile);
            const int col0 = t * WS_PX;
            ws_tma_store_3d(&maps.out, col0 * C, row0, b, os);
            if (col0 + WS_OBOX_PX < ow) ws_tma_store_3d(&maps.out, (col0 + WS_OBOX_PX) * C, row0, b, os + (unsigned)(WS_ROWS * WS_OBOX_PX * C * 4));
            bulk_commit();
        }
        cur = nxt; cur_staged = nxt_staged; cur_fx0 = nxt_fx0; cur_ylo = nxt_ylo;
        rx1 = rx2; ry1 = ry2;
    }
    if (tid == 0) bulk_wait_read0();      // shared memory must outlive the stores' reads
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

// fast path (warp_fwd_tile.cu): warp-autonomous 32x8 tiles, TMA staging
bool tile_path_ok(const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn_or_0);
int tile_tps(const float* U, const float* coord, long long cstride, const float* T, float* out, float* x_out, float* y_out,
             float* mask_out, int B, int H, int W, int oh, int ow, int pn, int flags, cudaStream_t st);
bool tps_nodes_ok(int H, int W, int C, int oh, int ow, int pn, int flags);
int tile_tps_fused(const float* U, const float* coord, const float* vector, const double* winv, float* T_out, float* out, float* x_out,
                   float* y_out, float* mask_out, int B, int H, int W, int oh, int ow, int pn, int flags, cudaStream_t st);
constexpr int FUSE_N = 32;      // == TFUSE_N of tile_common.cuh
int tile_given(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int oh, int ow, cudaStream_t st);
int tile_flow(const float* im, const float* flow, float* out, int B, int H, int W, cudaStream_t st);
int tile_homog(const float* im, const float* theta, int projective, float* out, float* x_out, float* y_out, int B, int H, int W,
               int oh, int ow, cudaStream_t st);
void tile_set_tuning(int debug_mask, int target_ctas, int minb);

static bool use_tile(int flags, const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn) {
    return !(flags & DVSG_FLAG_FORCE_DIRECT) && tile_path_ok(src, out, H, W, C, oh, ow, pn);
}

template <int MODE>
static int launch_fwd(FwdParams p, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    DVSG_REQUIRE(p.B <= 65535, "batch %d exceeds the grid z limit 65535: split the call", p.B);
    p.kc_cap = MODE == MODE_TPS ? (p.pn < KC ? p.pn : KC) : 0;
    const size_t smem = (size_t)p.kc_cap * (sizeof(float4) + TH * sizeof(float));
    dim3 grid((p.ow + TW - 1) / TW, (p.oh + TH - 1) / TH, p.B);
    DVSG_REQUIRE(grid.y <= 65535, "output height %d too large", p.oh);
    warp_fwd_kernel<MODE><<<grid, NT, smem, st>>>(p);
    count_launch();
    return check_launch("warp_fwd_kernel");
}

static bool wide_path_ok(int flags, const void* src, const void* out, int C, int oh, int ow) {
    return !(flags & DVSG_FLAG_FORCE_DIRECT) && C >= 4 && C % 2 == 0 && (reinterpret_cast<uintptr_t>(src) & 7u) == 0 &&
           (reinterpret_cast<uintptr_t>(out) & 7u) == 0 && (oh + WIDE_ROWS - 1) / WIDE_ROWS <= 65535;
}

// Staging box: 10 pixels x 12 rows, two side by side = 20 x 12 source pixels for a 16 x 8 output tile.  The box size is the
// first-order knob of this kernel: every staged byte crosses L2 -> shared memory through the TMA unit, and with the widest
// box that 256 elements allow (14 px at C = 18, 28 x 12 pixels staged per tile = 2.6x the output) the kernel sat at 0.62-0.68
// of HBM whatever else was done to it; 10 x 12: 0.74 at 288 x 512, 0.81 at 1080p (10 x 10 is no faster and sends sheared
// grids to the global-memory path).  0 when ten pixels exceed a box (C > 25).
static int ws_box_px(int C) { return (10 * C <= 256 && (10 * C) % 4 == 0) ? 10 : 0; }

template <int MODE>
static int launch_fwd_wide(FwdParams p, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    DVSG_REQUIRE(p.B <= 65535, "batch %d exceeds the grid z limit 65535: split the call", p.B);
    // TMA-staged variant: rows of W*C and ow*C floats at 16-byte multiples, 16-byte aligned bases, boxes wide enough for a tile
    static const bool no_staged = getenv("DVSG_WIDE_DIRECT") != nullptr;      // A/B experiments
    int pxb = ws_box_px(p.C), box_rows = WS_BOX_ROWS;
    if (const char* e = getenv("DVSG_WIDE_BOX")) {      // experiments only: "pixels per box,rows"
        int a = 0, r = 0;
        if (sscanf(e, "%d,%d", &a, &r) == 2 && a >= 9 && a * p.C <= 256 && (a * p.C) % 4 == 0 && r >= 9 && r <= 16) { pxb = a; box_rows = r; }
    }
    if (!no_staged && pxb > 0 && WS_OBOX_PX * p.C <= 256 && ((long long)p.W * p.C) % 4 == 0 && ((long long)p.ow * p.C) % 4 == 0 && aligned16(p.src) &&
        aligned16(p.out) && (p.oh + WS_ROWS - 1) / WS_ROWS <= 65535 && (long long)(p.W + 2 * pxb + 2) * p.C < (1LL << 31) &&
        (long long)(p.ow + WS_PX) * p.C < (1LL << 31)) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn enc = []() -> EncodeFn {
            void* f = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
            return reinterpret_cast<EncodeFn>(f);
        }();
        if (enc) {
            WideMaps maps;
            const cuuint32_t estr[3] = {1, 1, 1};
            const cuuint64_t sdims[3] = {(cuuint64_t)p.W * p.C, (cuuint64_t)p.H, (cuuint64_t)p.B};
            const cuuint64_t sstr[2] = {(cuuint64_t)p.W * p.C * 4, (cuuint64_t)p.H * p.W * p.C * 4};
            const cuuint32_t sbox[3] = {(cuuint32_t)(pxb * p.C), (cuuint32_t)box_rows, 1};
            const cuuint64_t odims[3] = {(cuuint64_t)p.ow * p.C, (cuuint64_t)p.oh, (cuuint64_t)p.B};
            const cuuint64_t ostr[2] = {(cuuint64_t)p.ow * p.C * 4, (cuuint64_t)p.oh * p.ow * p.C * 4};
            const cuuint32_t obox[3] = {(cuuint32_t)(WS_OBOX_PX * p.C), (cuuint32_t)WS_ROWS, 1};
            if (enc(&maps.src, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.src), sdims, sstr, sbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                enc(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.out, odims, ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                const int pal_shift = (p.C % 4 == 0) ? 0 : 1;      // 2^pal_shift pixels of C floats are a multiple of 16 bytes
                const size_t box_bytes = (size_t)box_rows * pxb * p.C * sizeof(float);
                const size_t stage_bytes = (((box_bytes + 127) & ~(size_t)127) + box_bytes + 127) & ~(size_t)127;
                const size_t smem = 2 * stage_bytes + (size_t)WS_NT * p.C * sizeof(float);
                // tiles per CTA: long enough to amortise the pipeline's fill (8), short enough for >= 4 waves of CTAs
                const int n_tx = (p.ow + WS_PX - 1) / WS_PX;
                const long long strips = (long long)p.B * ((p.oh + WS_ROWS - 1) / WS_ROWS);
                int seg_len = 8;
                if (const char* e = getenv("DVSG_WIDE_SEGLEN")) seg_len = max(atoi(e), 1);      // experiments only
                while (seg_len > 2 && strips * ((n_tx + seg_len - 1) / seg_len) < 4LL * 148 * 3) seg_len /= 2;
                seg_len = min(seg_len, n_tx);
                const dim3 grid((unsigned)((n_tx + seg_len - 1) / seg_len), (unsigned)((p.oh + WS_ROWS - 1) / WS_ROWS), (unsigned)p.B);
                auto k = warp_fwd_wide_staged_kernel<MODE>;
                ensure_dynamic_smem(reinterpret_cast<const void*>(k), (int)smem);
                k<<<grid, WS_NT, smem, st>>>(p, maps, pxb, pal_shift, (int)stage_bytes, seg_len, box_rows, 1.0f);
                count_launch();
                return check_launch("warp_fwd_wide_staged_kernel");
            }
        }
    }
    const dim3 grid((unsigned)((p.ow + WIDE_PX - 1) / WIDE_PX), (unsigned)((p.oh + WIDE_ROWS - 1) / WIDE_ROWS), (unsigned)p.B);
    warp_fwd_wide_kernel<MODE><<<grid, WIDE_NT, 0, st>>>(p);
    count_launch();
    return check_launch("warp_fwd_wide_kernel");
}

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_set_tile_tuning(int debug_mask, int target_ctas, int min_ctas_per_sm) {
    tile_set_tuning(debug_mask, target_ctas, min_ctas_per_sm);
    return DVSG_OK;
}

extern "C" int dvsg_tps_coords_mode(int H, int W, int C, int oh, int ow, int pn, int flags) {
    return tps_nodes_ok(H, W, C, oh, ow, pn, flags) ? 1 : 0;
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
        return B == 0 ? DVSG_OK : tile_tps(U, coord, coord_batch_stride, T, out, x_out, y_out, mask_out, B, H, W, oh, ow, pn, flags, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = U; p.out = out; p.x_out = x_out; p.y_out = y_out; p.mask_out = mask_out;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = coord_batch_stride; p.T = T; p.pn = pn;
    p.step_x = lin_step(ow); p.step_y = lin_step(oh);
    return launch_fwd<MODE_TPS>(p, (cudaStream_t)stream);
}

extern "C" int dvsg_bilinear_fwd(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int C,
                                 int oh, int ow, int flags, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0, "bilinear_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && x && y && out), "bilinear_fwd: null pointer");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "bilinear_fwd: frame too large for int32 indexing");
    if (use_tile(flags, im, out, H, W, C, oh, ow, 0))
        return B == 0 ? DVSG_OK : tile_given(im, x, y, out, B, H, W, oh, ow, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = im; p.out = out; p.x_in = x; p.y_in = y;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    if (wide_path_ok(flags, im, out, C, oh, ow)) return launch_fwd_wide<MODE_GIVEN>(p, (cudaStream_t)stream);
    return launch_fwd<MODE_GIVEN>(p, (cudaStream_t)stream);
}

extern "C" int dvsg_flow_warp_fwd(const float* im, const float* flow, float* out, int B, int H, int W, int C, int flags,
                                  void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, "flow_warp_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && flow && out), "flow_warp_fwd: null pointer");
    DVSG_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7u) == 0, "flow_warp_fwd: flow must be 8-byte aligned");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "flow_warp_fwd: frame too large for int32 indexing");
    if (use_tile(flags, im, out, H, W, C, H, W, 0))
        return B == 0 ? DVSG_OK : tile_flow(im, flow, out, B, H, W, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = im; p.out = out; p.flow = flow;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = H; p.ow = W;
    if (wide_path_ok(flags, im, out, C, H, W)) return launch_fwd_wide<MODE_FLOW>(p, (cudaStream_t)stream);
    return launch_fwd<MODE_FLOW>(p, (cudaStream_t)stream);
}

extern "C" int dvsg_homography_warp_fwd(const float* im, const float* theta, int projective, float* out, float* x_out,
                                        float* y_out, int B, int H, int W, int C, int oh, int ow, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0, "homography_warp_fwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && theta && out), "homography_warp_fwd: null pointer");
    DVSG_REQUIRE((x_out == nullptr) == (y_out == nullptr), "homography_warp_fwd: x_out and y_out must be given together");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "homography_warp_fwd: frame too large for int32 indexing");
    if (use_tile(0, im, out, H, W, C, oh, ow, 0))
        return B == 0 ? DVSG_OK : tile_homog(im, theta, projective, out, x_out, y_out, B, H, W, oh, ow, (cudaStream_t)stream);
    FwdParams p = {};
    p.src = im; p.out = out; p.x_out = x_out; p.y_out = y_out; p.theta = theta; p.projective = projective;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    p.step_x = lin_step(ow); p.step_y = lin_step(oh);
    if (wide_path_ok(0, im, out, C, oh, ow)) return launch_fwd_wide<MODE_HOMOG>(p, (cudaStream_t)stream);
    return launch_fwd<MODE_HOMOG>(p, (cudaStream_t)stream);
}

extern "C" int dvsg_st_meshgrid(float* grid, int oh, int ow, void* stream) {
    DVSG_REQUIRE(oh > 0 && ow > 0 && grid, "st_meshgrid: bad argument");
    const int n = oh * ow;
    st_meshgrid_kernel<<<(n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184, 256, 0, (cudaStream_t)stream>>>(grid, oh, ow, lin_step(ow), lin_step(oh));
    count_launch();
    return check_launch("st_meshgrid_kernel");
}

// One call per batch of frames for the online loop (eval.py:106-110 runs the graph once per frame): coefficients from the
// prepared inverse of the clip's constant mesh, then the fused warp -- two launches, one trip through the ABI.  `target`
// is coord + vector (ThinPlateSpline.py:161); T [B,2,pn+3] is scratch the caller owns (and may read afterwards).
extern "C" int dvsg_tps_warp_frames(const float* U, const float* coord, const float* target, void* prepared, size_t prepared_bytes,
                                    float* T, float* out, float* x_out, float* y_out, float* mask_out, int B, int H, int W, int C, int oh,
                                    int ow, int pn, void* stream) {
    const int rc = dvsg_tps_solve_prepared(coord, 0, target, T, B, pn, prepared, prepared_bytes, stream);
    if (rc) return rc;
    return dvsg_tps_warp_fwd(U, coord, 0, T, out, x_out, y_out, mask_out, B, H, W, C, oh, ow, pn, 0, stream);
}

// Same with the regressed offsets `vector` [B,pn,2] as the argument (coord + vector is formed inside the prepared solve)
extern "C" int dvsg_tps_warp_frames_offsets(const float* U, const float* coord, const float* vector, void* prepared, size_t prepared_bytes,
                                            float* T, float* out, float* x_out, float* y_out, float* mask_out, int B, int H, int W, int C,
                                            int oh, int ow, int pn, void* stream) {
    if (pn + 3 <= FUSE_N && B > 0 && H > 0 && W > 0 && oh > 0 && ow > 0 && U && coord && vector && T && out && prepared &&
        prepared_bytes >= dvsg_tps_prepare_workspace_bytes(B, pn, 0) && (x_out == nullptr) == (y_out == nullptr) &&
        use_tile(0, U, out, H, W, C, oh, ow, pn))
        // ONE launch: the prepared solve runs in the warp kernel's prologue (same arithmetic, same bits)
        return tile_tps_fused(U, coord, vector, reinterpret_cast<const double*>(prepared), T, out, x_out, y_out, mask_out, B, H, W, oh, ow, pn, 0,
                              (cudaStream_t)stream);
    const int rc = dvsg_tps_solve_offsets_prepared(coord, 0, vector, T, B, pn, prepared, prepared_bytes, stream);
    if (rc) return rc;
    return dvsg_tps_warp_fwd(U, coord, 0, T, out, x_out, y_out, mask_out, B, H, W, C, oh, ow, pn, 0, stream);
}
