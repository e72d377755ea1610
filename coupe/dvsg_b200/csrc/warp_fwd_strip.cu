// warp_fwd_strip.cu -- the sm_100a fast path of the forward warp (C == 3, 16-B aligned rows).
//
// One persistent CTA walks a horizontal strip of 64x16 output tiles of one frame and runs a
// two-stage software pipeline over them:
//
//   iteration t:   coordinates(t+1)  ->  exact source footprint of tile t+1 (warp REDUX +
//                  shared atomics)   ->  cp.async.bulk rows of that footprint into staging
//                  buffer (t+1)&1, completion on mbarrier (t+1)&1
//                  wait mbarrier t&1 (issued one iteration ago: normally already complete)
//                  gather the four corners of tile t from staging buffer t&1, blend, stage the
//                  output tile in shared memory, cp.async.bulk it to global
//
// so the HBM/L2 latency of the source rows hides behind a whole tile of arithmetic, the TPS
// tables ((px,py,cx,cy) per control point and (y_t - py)^2 per tile row) are built once per
// strip, and nothing but the frames themselves touches HBM (the [B, pn+3, h*w] basis and the
// sampling grid of the reference never exist).
//
// Instruction diet (the kernel is issue/XU-bound for TPS): corners are
// computed once per pixel and carried across the pipeline as two packed 16-bit pairs plus the
// pixel-space coordinate; int->float conversions use the 2^23 magic constant on the FMA/ALU
// pipes so that the XU pipe (16 lanes/clk/SM, shared with MUFU.LG2) only sees the pn
// logarithms and two F2I.FLOOR per pixel; the TPS basis runs in packed fp32x2 (FADD2/FMUL2/
// FFMA2); the blend stays in the reference's op order with explicitly rounded operations.
#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

enum { SMODE_TPS = 0, SMODE_GIVEN = 1, SMODE_FLOW = 2, SMODE_HOMOG = 3 };

constexpr int STW = 64;          // tile width (one thread per column)
constexpr int SPR = 4;           // rows per thread
constexpr int SRG = 4;           // row groups
constexpr int STH = SPR * SRG;   // tile height 16
constexpr int SNT = STW * SRG;   // 256 threads
constexpr int SKC = 256;         // max control points (tables resident for the whole strip)
constexpr float SLN2 = 0.6931471805599453f;

struct StripParams {
    const float* src;
    float* out;
    float* x_out;
    float* y_out;
    float* mask_out;
    int B, H, W, oh, ow;
    const float* coord;
    long long coord_stride;
    const float* T;
    int pn;
    float step_x, step_y;
    const float* x_in;
    const float* y_in;
    const float* flow;
    const float* theta;
    int projective;
    int src_smem_bytes;   // per staging buffer
    int n_tx, n_ty;       // tiles per row / tile rows
    int segs, seg_len;    // segments per strip, tiles per segment
};

// pixel state carried from the coordinate stage to the gather stage (4 registers)
struct Pix {
    float xp, yp;        // A4: pixel-space coordinate; ZP: clipped+1 coordinate in the padded frame
    unsigned ix, iy;     // x0 | x1 << 16, y0 | y1 << 16 (A4: clamped; ZP: padded-frame indices)
};

__device__ __forceinline__ float u2f(unsigned v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }   // v < 2^23, exact

__device__ __forceinline__ int floor_i32(float f) {
    // floor + the reference's CPU cast semantics (out-of-range / NaN -> INT_MIN); one F2I.FLOOR
    const int v = __float2int_rd(f);
    return fabsf(f) < 2147483648.0f ? v : (int)0x80000000;
}

// A4 corner stage: identical arithmetic to sampler_math.cuh::a4_corners (bit-exact indices)
__device__ __forceinline__ Pix a4_pix(float x, float y, int W, int H, int& x0, int& x1, int& y0, int& y1) {
    Pix s;
    s.xp = DVSG_MUL(DVSG_MUL(DVSG_ADD(x, 1.0f), (float)W), 0.5f);
    s.yp = DVSG_MUL(DVSG_MUL(DVSG_ADD(y, 1.0f), (float)H), 0.5f);
    const int fx = floor_i32(s.xp), fy = floor_i32(s.yp);
    x0 = min(max(fx, 0), W - 1);
    x1 = min(max((int)((unsigned)fx + 1u), 0), W - 1);
    y0 = min(max(fy, 0), H - 1);
    y1 = min(max((int)((unsigned)fy + 1u), 0), H - 1);
    s.ix = (unsigned)x0 | ((unsigned)x1 << 16);
    s.iy = (unsigned)y0 | ((unsigned)y1 << 16);
    return s;
}

// ZP corner stage (sampler_math.cuh::zp_corners); corners are indices into the padded frame
__device__ __forceinline__ Pix zp_pix(float xpix, float ypix, int W, int H, int& x0, int& x1, int& y0, int& y1) {
    Pix s;
    const float wf = (float)W, hf = (float)H;
    s.xp = DVSG_ADD(fminf(fmaxf(xpix, -1.0f), wf), 1.0f);
    s.yp = DVSG_ADD(fminf(fmaxf(ypix, -1.0f), hf), 1.0f);
    x0 = __float2int_rd(s.xp);                 // in [0, W+1]
    y0 = __float2int_rd(s.yp);
    x1 = min(x0 + 1, W + 1);
    y1 = min(y0 + 1, H + 1);
    s.ix = (unsigned)x0 | ((unsigned)x1 << 16);
    s.iy = (unsigned)y0 | ((unsigned)y1 << 16);
    return s;
}

__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// ptxas (CUDA 12.9) contracts mul.rn.f32x2 feeding add.rn.f32x2 into FFMA2 although both carry .rn (and
// regardless of -fmad), which would fuse roundings the reference keeps apart.  Products that feed a sum
// are therefore multiplied packed (FMUL2) but summed with scalar, explicitly rounded adds.
__device__ __forceinline__ float2 fmul2_exact(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fadd2_scalar(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }   // a - b, IEEE rn per lane
__device__ __forceinline__ float2 u2f2(unsigned a, unsigned b) {
    return __fadd2_rn(make_float2(__uint_as_float(0x4B000000u | a), __uint_as_float(0x4B000000u | b)), make_float2(-8388608.0f, -8388608.0f));
}

// PIPE = true : two-stage software pipeline inside the CTA (two staging buffers, 3 CTAs/SM)
// PIPE = false: one staging buffer, the load latency is hidden by the other resident CTAs (4/SM)
template <int MODE, bool PACK, bool PIPE>
__global__ void __launch_bounds__(SNT, PIPE ? 3 : 4) warp_fwd_strip_kernel(const StripParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    __shared__ int s_bbox[2][4];
    __shared__ float s_lin[12];

    float* s_out = reinterpret_cast<float*>(smem);                                    // [STH][STW*3]
    unsigned char* s_src0 = smem + STH * STW * 3 * sizeof(float);
    float4* s_pt = reinterpret_cast<float4*>(s_src0 + (PIPE ? 2 : 1) * (size_t)p.src_smem_bytes);  // [pn]
    float* s_dy2 = reinterpret_cast<float*>(s_pt + (MODE == SMODE_TPS ? p.pn : 0));   // [pn][STH]

    const int tid = threadIdx.x;
    const int tx = tid & (STW - 1);
    const int rbase = (tid >> 6) * SPR;
    const int H = p.H, W = p.W, oh = p.oh, ow = p.ow;

    // strip segment of this CTA
    int bid = blockIdx.x;
    const int seg = bid % p.segs; bid /= p.segs;
    const int ty = bid % p.n_ty;
    const int b = bid / p.n_ty;
    const int t_begin = seg * p.seg_len;
    const int t_end = min(t_begin + p.seg_len, p.n_tx);
    const int row0 = ty * STH;
    const float* srcb = p.src + (size_t)b * H * W * 3;

    if (tid == 0) {
        mbar_init(smem_u32(&s_mbar[0]), 1);
        mbar_init(smem_u32(&s_mbar[1]), 1);
        fence_mbar_init();
        s_bbox[0][0] = s_bbox[1][0] = 0x7fffffff; s_bbox[0][1] = s_bbox[1][1] = -1;
        s_bbox[0][2] = s_bbox[1][2] = 0x7fffffff; s_bbox[0][3] = s_bbox[1][3] = -1;
    }
    // ---- per-strip tables ------------------------------------------------------------------
    if (MODE == SMODE_TPS) {
        const int N = p.pn + 3;
        const float* Tb = p.T + (size_t)b * 2 * N;
        const float* cb = p.coord + (size_t)b * p.coord_stride;
        if (tid < 6) s_lin[tid] = __ldg(Tb + (tid < 3 ? tid : N + tid - 3));
        for (int k = tid; k < p.pn; k += SNT)
            s_pt[k] = make_float4(__ldg(cb + 2 * k), __ldg(cb + 2 * k + 1), __ldg(Tb + 3 + k) * SLN2, __ldg(Tb + N + 3 + k) * SLN2);
        for (int i = tid; i < p.pn * STH; i += SNT) {
            const int k = i >> 4, r = i & (STH - 1);
            const float dy = DVSG_SUB(lin_coord(min(row0 + r, oh - 1), p.step_y), __ldg(cb + 2 * k + 1));
            s_dy2[i] = DVSG_MUL(dy, dy);
        }
    } else if (MODE == SMODE_HOMOG) {
        const int nt = p.projective ? 8 : 6;
        if (tid < 9) s_lin[tid] = tid < nt ? __ldg(p.theta + (size_t)b * nt + tid) : (tid == 8 ? 1.0f : 0.0f);
    }
    __syncthreads();

    // rows / columns past the frame edge are computed as duplicates of the edge pixel (their
    // results land in the unused part of the output tile): no divergence in the hot loops
    float yt[SPR];
#pragma unroll
    for (int q = 0; q < SPR; ++q) yt[q] = lin_coord(min(row0 + rbase + q, oh - 1), p.step_y);

    // raw inputs of the coordinate stage (GIVEN / FLOW), prefetched one tile ahead
    auto load_raw = [&](int t, float (&rx)[SPR], float (&ry)[SPR]) {
        const int col = min(t * STW + tx, ow - 1);
#pragma unroll
        for (int q = 0; q < SPR; ++q) {
            rx[q] = ry[q] = 0.0f;
            if ((MODE == SMODE_GIVEN || MODE == SMODE_FLOW) && t < t_end) {
                const size_t i = ((size_t)b * oh + min(row0 + rbase + q, oh - 1)) * ow + col;
                if (MODE == SMODE_GIVEN) { rx[q] = __ldg(p.x_in + i); ry[q] = __ldg(p.y_in + i); }
                if (MODE == SMODE_FLOW) { const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + i); rx[q] = f.x; ry[q] = f.y; }
            }
        }
    };

    // footprint descriptor of a staged tile: origin (px), row pitch (bytes); pitch 0 = not staged
    struct Foot { int fx0, fy0, pitch; };
    unsigned uses[2] = {0u, 0u};   // completed uses of each mbarrier (phase parity)

    // ---- coordinate stage for tile t: Pix state, optional x/y/mask outputs, footprint, loads --
    auto stage_coords = [&](int t, Pix (&px)[SPR], Foot& foot, const float (&rawx)[SPR], const float (&rawy)[SPR]) {
        const bool col_ok = t * STW + tx < ow;
        const int col = min(t * STW + tx, ow - 1);
        float xs[SPR], ys[SPR];
        if (MODE == SMODE_TPS) {
            const float xt = lin_coord(col, p.step_x);
#pragma unroll
            for (int q = 0; q < SPR; ++q) {
                xs[q] = fmaf(s_lin[2], yt[q], fmaf(s_lin[1], xt, s_lin[0]));
                ys[q] = fmaf(s_lin[5], yt[q], fmaf(s_lin[4], xt, s_lin[3]));
            }
            if (PACK) {
                float2 xa = make_float2(xs[0], xs[1]), xb = make_float2(xs[2], xs[3]);
                float2 ya = make_float2(ys[0], ys[1]), yb = make_float2(ys[2], ys[3]);
                const float2 eps = make_float2(1e-6f, 1e-6f);
#pragma unroll 4
                for (int k = 0; k < p.pn; ++k) {
                    const float4 c = s_pt[k];
                    const float dx = DVSG_SUB(xt, c.x);
                    const float dx2 = DVSG_MUL(dx, dx);
                    const float4 d = *reinterpret_cast<const float4*>(s_dy2 + k * STH + rbase);
                    const float2 dxx = make_float2(dx2, dx2);
                    const float2 d2a = __fadd2_rn(dxx, make_float2(d.x, d.y));
                    const float2 d2b = __fadd2_rn(dxx, make_float2(d.z, d.w));
                    const float2 ta = __fadd2_rn(d2a, eps);
                    const float2 tb = __fadd2_rn(d2b, eps);
                    const float2 ra = __fmul2_rn(d2a, make_float2(lg2_approx(ta.x), lg2_approx(ta.y)));
                    const float2 rb = __fmul2_rn(d2b, make_float2(lg2_approx(tb.x), lg2_approx(tb.y)));
                    const float2 cx = make_float2(c.z, c.z), cy = make_float2(c.w, c.w);
                    xa = __ffma2_rn(cx, ra, xa);
                    xb = __ffma2_rn(cx, rb, xb);
                    ya = __ffma2_rn(cy, ra, ya);
                    yb = __ffma2_rn(cy, rb, yb);
                }
                xs[0] = xa.x; xs[1] = xa.y; xs[2] = xb.x; xs[3] = xb.y;
                ys[0] = ya.x; ys[1] = ya.y; ys[2] = yb.x; ys[3] = yb.y;
            } else {
#pragma unroll 4
                for (int k = 0; k < p.pn; ++k) {
                    const float4 c = s_pt[k];
                    const float dx = DVSG_SUB(xt, c.x);
                    const float dx2 = DVSG_MUL(dx, dx);
                    const float4 d = *reinterpret_cast<const float4*>(s_dy2 + k * STH + rbase);
                    const float dv[SPR] = {d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int q = 0; q < SPR; ++q) {
                        const float d2 = DVSG_ADD(dx2, dv[q]);
                        const float r = DVSG_MUL(d2, lg2_approx(DVSG_ADD(d2, 1e-6f)));
                        xs[q] = fmaf(c.z, r, xs[q]);
                        ys[q] = fmaf(c.w, r, ys[q]);
                    }
                }
            }
        } else if (MODE == SMODE_GIVEN) {
#pragma unroll
            for (int q = 0; q < SPR; ++q) { xs[q] = zp_pix_from_norm(rawx[q], W); ys[q] = zp_pix_from_norm(rawy[q], H); }
        } else if (MODE == SMODE_FLOW) {
#pragma unroll
            for (int q = 0; q < SPR; ++q) { xs[q] = DVSG_ADD((float)col, rawx[q]); ys[q] = DVSG_ADD((float)min(row0 + rbase + q, oh - 1), rawy[q]); }
        } else {
            const float xt = lin_coord(col, p.step_x);
#pragma unroll
            for (int q = 0; q < SPR; ++q) {
                float xn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[0], xt), DVSG_MUL(s_lin[1], yt[q])), s_lin[2]);
                float yn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[3], xt), DVSG_MUL(s_lin[4], yt[q])), s_lin[5]);
                if (p.projective) {
                    const float zn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[6], xt), DVSG_MUL(s_lin[7], yt[q])), s_lin[8]);
                    xn = zn != 0.0f ? DVSG_DIV(xn, zn) : 0.0f;
                    yn = zn != 0.0f ? DVSG_DIV(yn, zn) : 0.0f;
                }
                xs[q] = xn; ys[q] = yn;
            }
        }
        if ((MODE == SMODE_TPS || MODE == SMODE_HOMOG) && p.x_out && col_ok) {
#pragma unroll
            for (int q = 0; q < SPR; ++q) {
                const int row = row0 + rbase + q;
                if (row < oh) {
                    const size_t i = ((size_t)b * oh + row) * ow + col;
                    p.x_out[i] = xs[q];
                    p.y_out[i] = ys[q];
                }
            }
        }
        if (MODE == SMODE_HOMOG) {
#pragma unroll
            for (int q = 0; q < SPR; ++q) { xs[q] = zp_pix_from_norm(xs[q], W); ys[q] = zp_pix_from_norm(ys[q], H); }
        }
        // corners (once per pixel) + thread-local footprint
        int xmin = 0x7fffffff, xmax = -1, ymin = 0x7fffffff, ymax = -1;
#pragma unroll
        for (int q = 0; q < SPR; ++q) {
            int x0, x1, y0, y1;
            if (MODE == SMODE_TPS) {
                px[q] = a4_pix(xs[q], ys[q], W, H, x0, x1, y0, y1);
            } else {
                px[q] = zp_pix(xs[q], ys[q], W, H, x0, x1, y0, y1);
                x0 = max(x0, 1) - 1; x1 = min(x1, W) - 1;   // real pixels covered by valid corners
                y0 = max(y0, 1) - 1; y1 = min(y1, H) - 1;
            }
            if (x0 <= x1 && y0 <= y1) {
                xmin = min(xmin, x0); xmax = max(xmax, x1);
                ymin = min(ymin, y0); ymax = max(ymax, y1);
            }
        }
        const int slot = t & 1;                  // bbox slot
        const int bslot = PIPE ? slot : 0;       // staging buffer / mbarrier slot
        xmin = __reduce_min_sync(0xffffffffu, xmin);
        xmax = __reduce_max_sync(0xffffffffu, xmax);
        ymin = __reduce_min_sync(0xffffffffu, ymin);
        ymax = __reduce_max_sync(0xffffffffu, ymax);
        if ((tid & 31) == 0) {
            atomicMin(&s_bbox[slot][0], xmin); atomicMax(&s_bbox[slot][1], xmax);
            atomicMin(&s_bbox[slot][2], ymin); atomicMax(&s_bbox[slot][3], ymax);
        }
        if (tid < STH) bulk_wait_read0();      // previous tile's output rows have left s_out
        if (tid == 0) {
            // re-arm the other slot for tile t+1: its last readers (tile t-1) finished before the
            // S2 of the gather that preceded this call, its next writers come after this S1
            s_bbox[slot ^ 1][0] = 0x7fffffff; s_bbox[slot ^ 1][1] = -1; s_bbox[slot ^ 1][2] = 0x7fffffff; s_bbox[slot ^ 1][3] = -1;
        }
        __syncthreads();                       // S1
        xmin = s_bbox[slot][0]; xmax = s_bbox[slot][1]; ymin = s_bbox[slot][2]; ymax = s_bbox[slot][3];
        foot.fx0 = 0; foot.fy0 = 0; foot.pitch = 0;
        if (xmax >= xmin && ymax >= ymin) {
            const int fx0 = xmin & ~3;                                  // 4 px = 48 B keeps rows 16-B aligned
            const int wpx = min(((xmax - fx0 + 1) + 3) & ~3, W - fx0);
            const int nrows = ymax - ymin + 1;
            const unsigned row_bytes = (unsigned)wpx * 12u;
            if ((size_t)row_bytes * nrows <= (size_t)p.src_smem_bytes) {
                foot.fx0 = fx0; foot.fy0 = ymin; foot.pitch = (int)row_bytes;
                if (tid < 32) {
                    const uint32_t mbar = smem_u32(&s_mbar[bslot]);
                    if (tid == 0) mbar_arrive_expect_tx(mbar, row_bytes * (unsigned)nrows);
                    __syncwarp();
                    const uint32_t dst = smem_u32(s_src0 + (size_t)bslot * p.src_smem_bytes);
                    for (int r = tid; r < nrows; r += 32)
                        bulk_g2s(dst + (unsigned)r * row_bytes, srcb + ((size_t)(ymin + r) * W + fx0) * 3, row_bytes, mbar);
                }
            }
        }
    };

    // ---- gather stage for tile t -------------------------------------------------------------
    // Weight factors, weights and the blend run in packed fp32x2 (per-lane IEEE rn, so still the
    // reference's exactly rounded op sequence): pixel pairs (q, q+1) share an instruction for the
    // factors / weights / channel 2, channels (0, 1) of one pixel share an instruction.
    auto stage_gather = [&](int t, const Pix (&px)[SPR], const Foot& foot) {
        const int col = t * STW + tx;
        const int bslot = PIPE ? (t & 1) : 0;
        const bool staged = foot.pitch != 0;
        if (staged) {
            mbar_wait(smem_u32(&s_mbar[bslot]), uses[bslot] & 1u);
            uses[bslot]++;
        }
        // byte address (shared window) of source pixel (0,0) of the frame inside the staging buffer
        const uint32_t sbase = smem_u32(s_src0 + (size_t)bslot * p.src_smem_bytes) - (uint32_t)(foot.fy0 * foot.pitch + foot.fx0 * 12);
        const uint32_t obase = smem_u32(s_out) + (uint32_t)((rbase * STW + tx) * 12);
#pragma unroll
        for (int qq = 0; qq < SPR; qq += 2) {
            float2 w00, w01, w10, w11;              // lanes = pixels qq, qq+1
            unsigned cx0[2], cx1[2], cy0[2], cy1[2];
            bool v00[2], v01[2], v10[2], v11[2];
            {
                const unsigned ax_ = px[qq].ix, bx_ = px[qq + 1].ix, ay_ = px[qq].iy, by_ = px[qq + 1].iy;
                const unsigned x0a = ax_ & 0xffffu, x1a = ax_ >> 16, x0b = bx_ & 0xffffu, x1b = bx_ >> 16;
                const unsigned y0a = ay_ & 0xffffu, y1a = ay_ >> 16, y0b = by_ & 0xffffu, y1b = by_ >> 16;
                const float2 xp2 = make_float2(px[qq].xp, px[qq + 1].xp), yp2 = make_float2(px[qq].yp, px[qq + 1].yp);
                float2 ax0, ax1, ay0, ay1;
                if (MODE == SMODE_TPS) {
                    ax1 = f2sub(u2f2(x1a, x1b), xp2); ax0 = f2sub(xp2, u2f2(x0a, x0b));
                    ay1 = f2sub(u2f2(y1a, y1b), yp2); ay0 = f2sub(yp2, u2f2(y0a, y0b));
                    cx0[0] = x0a; cx1[0] = x1a; cy0[0] = y0a; cy1[0] = y1a;
                    cx0[1] = x0b; cx1[1] = x1b; cy0[1] = y0b; cy1[1] = y1b;
                    v00[0] = v01[0] = v10[0] = v11[0] = v00[1] = v01[1] = v10[1] = v11[1] = true;
                } else {
                    const float2 x0f = u2f2(x0a, x0b), y0f = u2f2(y0a, y0b), one = make_float2(1.0f, 1.0f);
                    ax1 = f2sub(__fadd2_rn(x0f, one), xp2); ax0 = f2sub(xp2, x0f);
                    ay1 = f2sub(__fadd2_rn(y0f, one), yp2); ay0 = f2sub(yp2, y0f);
                    const unsigned xs_[2][2] = {{x0a, x1a}, {x0b, x1b}}, ys_[2][2] = {{y0a, y1a}, {y0b, y1b}};
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const bool vx0 = zp_valid((int)xs_[j][0], W), vx1 = zp_valid((int)xs_[j][1], W);
                        const bool vy0 = zp_valid((int)ys_[j][0], H), vy1 = zp_valid((int)ys_[j][1], H);
                        v00[j] = vx0 && vy0; v01[j] = vx1 && vy0; v10[j] = vx0 && vy1; v11[j] = vx1 && vy1;
                        cx0[j] = (unsigned)min(max((int)xs_[j][0], 1) - 1, W - 1); cx1[j] = (unsigned)max(min((int)xs_[j][1], W) - 1, 0);
                        cy0[j] = (unsigned)min(max((int)ys_[j][0], 1) - 1, H - 1); cy1[j] = (unsigned)max(min((int)ys_[j][1], H) - 1, 0);
                    }
                }
                // reference pairing: 00 = (x0,y0), 01 = (x1,y0), 10 = (x0,y1), 11 = (x1,y1)
                w00 = fmul2_exact(ax1, ay1); w01 = fmul2_exact(ax0, ay1); w10 = fmul2_exact(ax1, ay0); w11 = fmul2_exact(ax0, ay0);
            }
            if (MODE == SMODE_TPS && p.mask_out) {   // A4 add_n order: a=(x0,y0), b=(x0,y1), c=(x1,y0), d=(x1,y1)
                const float2 m = fadd2_scalar(fadd2_scalar(fadd2_scalar(w00, w10), w01), w11);
                if (col < ow && row0 + rbase + qq < oh) p.mask_out[((size_t)b * oh + row0 + rbase + qq) * ow + col] = m.x;
                if (col < ow && row0 + rbase + qq + 1 < oh) p.mask_out[((size_t)b * oh + row0 + rbase + qq + 1) * ow + col] = m.y;
            }
            float i00[2][3], i01[2][3], i10[2][3], i11[2][3];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (staged) {
                    const uint32_t r0 = sbase + cy0[j] * (unsigned)foot.pitch, r1 = sbase + cy1[j] * (unsigned)foot.pitch;
                    const uint32_t a00 = r0 + cx0[j] * 12u, a01 = r0 + cx1[j] * 12u, a10 = r1 + cx0[j] * 12u, a11 = r1 + cx1[j] * 12u;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        i00[j][ch] = v00[j] ? lds_f32(a00 + 4 * ch) : 0.0f; i01[j][ch] = v01[j] ? lds_f32(a01 + 4 * ch) : 0.0f;
                        i10[j][ch] = v10[j] ? lds_f32(a10 + 4 * ch) : 0.0f; i11[j][ch] = v11[j] ? lds_f32(a11 + 4 * ch) : 0.0f;
                    }
                } else {
                    const float* a00 = srcb + ((size_t)cy0[j] * W + cx0[j]) * 3;
                    const float* a01 = srcb + ((size_t)cy0[j] * W + cx1[j]) * 3;
                    const float* a10 = srcb + ((size_t)cy1[j] * W + cx0[j]) * 3;
                    const float* a11 = srcb + ((size_t)cy1[j] * W + cx1[j]) * 3;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        i00[j][ch] = v00[j] ? __ldg(a00 + ch) : 0.0f; i01[j][ch] = v01[j] ? __ldg(a01 + ch) : 0.0f;
                        i10[j][ch] = v10[j] ? __ldg(a10 + ch) : 0.0f; i11[j][ch] = v11[j] ? __ldg(a11 + ch) : 0.0f;
                    }
                }
            }
            // blend; second / third term swap between the two samplers' add_n orders
            float2 o01[2], o2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float wa = j ? w00.y : w00.x, wb = j ? w01.y : w01.x, wc = j ? w10.y : w10.x, wd = j ? w11.y : w11.x;
                const float2 t00 = fmul2_exact(make_float2(wa, wa), make_float2(i00[j][0], i00[j][1]));
                const float2 t01 = fmul2_exact(make_float2(wb, wb), make_float2(i01[j][0], i01[j][1]));
                const float2 t10 = fmul2_exact(make_float2(wc, wc), make_float2(i10[j][0], i10[j][1]));
                const float2 t11 = fmul2_exact(make_float2(wd, wd), make_float2(i11[j][0], i11[j][1]));
                if (MODE == SMODE_TPS) o01[j] = fadd2_scalar(fadd2_scalar(fadd2_scalar(t00, t10), t01), t11);   // ThinPlateSpline.py:89
                else o01[j] = fadd2_scalar(fadd2_scalar(fadd2_scalar(t00, t01), t10), t11);                       // spatial_transformer.py:562
            }
            {
                const float2 t00 = fmul2_exact(w00, make_float2(i00[0][2], i00[1][2])), t01 = fmul2_exact(w01, make_float2(i01[0][2], i01[1][2]));
                const float2 t10 = fmul2_exact(w10, make_float2(i10[0][2], i10[1][2])), t11 = fmul2_exact(w11, make_float2(i11[0][2], i11[1][2]));
                if (MODE == SMODE_TPS) o2 = fadd2_scalar(fadd2_scalar(fadd2_scalar(t00, t10), t01), t11);
                else o2 = fadd2_scalar(fadd2_scalar(fadd2_scalar(t00, t01), t10), t11);
            }
            const uint32_t oa = obase + (uint32_t)(qq * STW * 12);
            sts_f32(oa, o01[0].x); sts_f32(oa + 4, o01[0].y); sts_f32(oa + 8, o2.x);
            sts_f32(oa + STW * 12, o01[1].x); sts_f32(oa + STW * 12 + 4, o01[1].y); sts_f32(oa + STW * 12 + 8, o2.y);
        }
        fence_proxy_async_smem();
        __syncthreads();                       // S2
        const int col0 = t * STW;
        const int vcols = min(STW, ow - col0);
        if (tid < STH && row0 + tid < oh) {
            bulk_s2g(p.out + (((size_t)b * oh + row0 + tid) * ow + col0) * 3, smem_u32(s_out + tid * STW * 3), (unsigned)vcols * 12u);
            bulk_commit();
        }
    };

    // ---- the pipeline -------------------------------------------------------------------------
    Pix pa[SPR], pb[SPR];
    Foot fa, fb;
    float ax[SPR], ay[SPR], bx[SPR], by[SPR];   // raw inputs: (ax, ay) for even steps, (bx, by) for odd steps
    load_raw(t_begin, ax, ay);
    if (PIPE) {
        load_raw(t_begin + 1, bx, by);
        stage_coords(t_begin, pa, fa, ax, ay);
        for (int t = t_begin; t < t_end; t += 2) {
            if (t + 1 < t_end) {
                load_raw(t + 2, ax, ay);
                stage_coords(t + 1, pb, fb, bx, by);
            } else {
                if (tid < STH) bulk_wait_read0();
                __syncthreads();
            }
            stage_gather(t, pa, fa);
            if (t + 1 >= t_end) break;
            if (t + 2 < t_end) {
                load_raw(t + 3, bx, by);
                stage_coords(t + 2, pa, fa, ax, ay);
            } else {
                if (tid < STH) bulk_wait_read0();
                __syncthreads();
            }
            stage_gather(t + 1, pb, fb);
        }
    } else {
        for (int t = t_begin; t < t_end; ++t) {
            load_raw(t + 1, bx, by);
            stage_coords(t, pa, fa, ax, ay);
            stage_gather(t, pa, fa);
#pragma unroll
            for (int q = 0; q < SPR; ++q) { ax[q] = bx[q]; ay[q] = by[q]; }
        }
    }
    if (tid < STH) bulk_wait_read0();   // shared memory must outlive the last bulk stores' reads
}

static float strip_lin_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }
static int g_strip_smem = 28 * 1024;
static int g_strip_pack = 1;
static int g_strip_pipe = 0;
static int g_strip_target_ctas = 148 * 3 * 4;

bool strip_path_ok(const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn_or_0) {
    return C == 3 && W % 4 == 0 && ow % 4 == 0 && aligned16(src) && aligned16(out) && W < 32760 && H < 32760 && pn_or_0 <= SKC;
}

template <int MODE>
static int launch_strip(StripParams p, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    p.src_smem_bytes = g_strip_smem;
    p.n_tx = (p.ow + STW - 1) / STW;
    p.n_ty = (p.oh + STH - 1) / STH;
    const long long strips = (long long)p.B * p.n_ty;
    int segs = 1;
    if (strips < g_strip_target_ctas) segs = (int)min((long long)p.n_tx, (g_strip_target_ctas + strips - 1) / strips);
    p.seg_len = (p.n_tx + segs - 1) / segs;
    p.segs = (p.n_tx + p.seg_len - 1) / p.seg_len;
    const long long ctas = strips * p.segs;
    DVSG_REQUIRE(ctas < (1LL << 31), "strip kernel: %lld CTAs exceed the grid limit: split the batch", ctas);
    const size_t tps_bytes = MODE == SMODE_TPS ? (size_t)p.pn * (sizeof(float4) + STH * sizeof(float)) : 0;
    const bool pipe = g_strip_pipe != 0;
    const size_t smem = (size_t)STH * STW * 3 * sizeof(float) + (pipe ? 2 : 1) * (size_t)p.src_smem_bytes + tps_bytes;
    const bool pack = g_strip_pack && MODE == SMODE_TPS;
#define DVSG_LAUNCH_STRIP(PK, PP)                                                                              \
    do {                                                                                                       \
        auto k = warp_fwd_strip_kernel<MODE, PK, PP>;                                                          \
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
        k<<<(unsigned)ctas, SNT, smem, st>>>(p);                                                               \
    } while (0)
    if (pack) { if (pipe) DVSG_LAUNCH_STRIP(true, true); else DVSG_LAUNCH_STRIP(true, false); }
    else      { if (pipe) DVSG_LAUNCH_STRIP(false, true); else DVSG_LAUNCH_STRIP(false, false); }
#undef DVSG_LAUNCH_STRIP
    count_launch();
    return check_launch("warp_fwd_strip_kernel");
}

int strip_tps(const float* U, const float* coord, long long cstride, const float* T, float* out, float* x_out, float* y_out,
              float* mask_out, int B, int H, int W, int oh, int ow, int pn, cudaStream_t st) {
    StripParams p = {};
    p.src = U; p.out = out; p.x_out = x_out; p.y_out = y_out; p.mask_out = mask_out;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = cstride; p.T = T; p.pn = pn;
    p.step_x = strip_lin_step(ow); p.step_y = strip_lin_step(oh);
    return launch_strip<SMODE_TPS>(p, st);
}

int strip_given(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int oh, int ow, cudaStream_t st) {
    StripParams p = {};
    p.src = im; p.out = out; p.x_in = x; p.y_in = y;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    return launch_strip<SMODE_GIVEN>(p, st);
}

int strip_flow(const float* im, const float* flow, float* out, int B, int H, int W, cudaStream_t st) {
    StripParams p = {};
    p.src = im; p.out = out; p.flow = flow;
    p.B = B; p.H = H; p.W = W; p.oh = H; p.ow = W;
    return launch_strip<SMODE_FLOW>(p, st);
}

int strip_homog(const float* im, const float* theta, int projective, float* out, float* x_out, float* y_out, int B, int H, int W,
                int oh, int ow, cudaStream_t st) {
    StripParams p = {};
    p.src = im; p.out = out; p.x_out = x_out; p.y_out = y_out; p.theta = theta; p.projective = projective;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    p.step_x = strip_lin_step(ow); p.step_y = strip_lin_step(oh);
    return launch_strip<SMODE_HOMOG>(p, st);
}

void strip_set_tuning(int smem_bytes, int pack, int target_ctas, int pipe) {
    if (smem_bytes >= 0) g_strip_smem = smem_bytes & ~15;
    if (pack >= 0) g_strip_pack = pack;
    if (target_ctas > 0) g_strip_target_ctas = target_ctas;
    if (pipe >= 0) g_strip_pipe = pipe;
}

}  // namespace dvsg
