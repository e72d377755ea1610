// warp_fwd_tile.cu -- the sm_100a fast path of the forward warp (C == 3, 16-B aligned rows).
//
// Every WARP is an autonomous pipeline over 32x8-pixel output tiles (lane = column, 8 rows per
// thread); the four warps of a CTA share only the per-strip TPS tables and never meet at a
// CTA barrier after the prologue, so warps drift apart and the SM always has some warps in
// the arithmetic phase while others wait for their source rows:
//
//   A  coordinates   TPS basis in packed fp32x2 (FADD2 / FMUL2 / FFMA2), one MUFU.LG2 per
//                    (pixel, control point), 64-B table records per control point read as
//                    shared-memory broadcasts; or flow / given x,y / homography
//   F  footprint     NaN-propagating FMNMX3 bounding box of the tile's sampling coordinates,
//                    4 warp REDUX, no shared atomics, no CTA barrier
//   L  staging       the footprint -> the warp's private staging buffer with ONE 3-D TMA tensor
//                    copy (cp.async.bulk.tensor, SASS UTMALDG) of a 128- or 160-float wide box,
//                    completion on the warp's mbarrier; out-of-frame parts arrive as zeros,
//                    which IS the zero padding of bilinear_interp / tf_warp
//   G  gather+blend  four corners from shared memory; weights and add_n in the reference's
//                    op order with separately rounded products and sums (bit-exact sampler)
//   S  store         output tile staged in shared memory, then one TMA tensor store (clipped at
//                    the frame edge by the hardware)
//
// Gather variants, chosen per tile (warp-uniform):
//   interior  every corner of every pixel lies strictly inside the frame: no clamps, corners
//             are a, a+12, a+pitch, a+pitch+12; floor (FADD2.RM with the 2^23 constant) and
//             the address arithmetic run packed on the FMA/ALU pipes, so the XU pipe only
//             sees the logarithms
//   clamped   TPS tiles touching the frame border: same scheme plus the integer clamps of
//             ThinPlateSpline.py:57-60 done on exact fp32 integers
//   general   per-pixel scalar code with the full reference semantics, corners straight from
//             global memory: footprints larger than the biggest box, non-finite coordinates
// Nothing but the frames themselves touches HBM: the [B, pn+3, h*w] basis and the sampling
// grid of the reference never exist (x, y are written only when the caller asks).
#include <stdlib.h>

#include "tile_common.cuh"

namespace dvsg {

// ---- per-pixel general gather: full reference semantics --------------------------------------------
// xp, yp: TPS -> pixel-space coordinate of the A4 sampler; other modes -> clipped+1 coordinate in the
// zero-padded frame.  Corners come straight from global memory.
// Wide corner loads: the two corners of one source row are 24 contiguous bytes at a 4-byte alignment (12 x0 mod 16 is 0, 12,
// 8 or 4).  Twelve scalar loads per pixel made this path bound by L1 tag lookups (each LDG.32 of a scattered warp is 32
// lookups); here a row is fetched with two 16-byte aligned LDG.128 (plus one LDG.32 for the word that spills over when
// x0 = 1 mod 4) and the six floats are rotated into place with selects: 4-5 lookups per pixel instead of 12.  The tile
// path guarantees W % 4 == 0 and a 16-byte aligned frame, so every row starts on a 16-byte boundary and no load that is
// issued reaches past the frame (each is predicated on containing a needed word).
__device__ __forceinline__ void row_corners(const float* __restrict__ srcb, int y, int x0, int x1, int W, float (&c0)[3], float (&c1)[3]) {
    const size_t e = ((size_t)y * W + x0) * 3;
    const int k = (int)(e & 3);
    const float* g = srcb + (e - k);
    const bool two = x1 != x0;
    const int last = k + (two ? 5 : 2);               // last needed word, relative to g
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(g));
    float4 q1 = make_float4(0.f, 0.f, 0.f, 0.f);
    float q2 = 0.f;
    if (last >= 4) q1 = __ldg(reinterpret_cast<const float4*>(g) + 1);
    if (last >= 8) q2 = __ldg(g + 8);
    const float v[9] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2};
    float t[7], f[6];
#pragma unroll
    for (int j = 0; j < 7; ++j) t[j] = (k & 2) ? v[j + 2] : v[j];
#pragma unroll
    for (int j = 0; j < 6; ++j) f[j] = (k & 1) ? t[j + 1] : t[j];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { c0[ch] = f[ch]; c1[ch] = two ? f[3 + ch] : f[ch]; }
}

template <int MODE, bool WIDE_LD = true>
__device__ __noinline__ void general_pixel(float xp, float yp, int W, int H, const float* __restrict__ srcb, float* __restrict__ optr, float* mask_ptr) {
    int x0, x1, y0, y1;
    float ax0, ax1, ay0, ay1;
    bool v00 = true, v01 = true, v10 = true, v11 = true;   // 00 = (x0,y0), 01 = (x1,y0), 10 = (x0,y1), 11 = (x1,y1)
    if (MODE == TMODE_TPS) {
        const int fx = t_floor_i32(xp), fy = t_floor_i32(yp);
        x0 = min(max(fx, 0), W - 1);
        x1 = min(max((int)((unsigned)fx + 1u), 0), W - 1);
        y0 = min(max(fy, 0), H - 1);
        y1 = min(max((int)((unsigned)fy + 1u), 0), H - 1);
        ax1 = DVSG_SUB(t_u2f((unsigned)x1), xp); ax0 = DVSG_SUB(xp, t_u2f((unsigned)x0));
        ay1 = DVSG_SUB(t_u2f((unsigned)y1), yp); ay0 = DVSG_SUB(yp, t_u2f((unsigned)y0));
    } else {
        const int qx0 = __float2int_rd(xp), qy0 = __float2int_rd(yp);     // in [0, W+1] / [0, H+1]
        const int qx1 = min(qx0 + 1, W + 1), qy1 = min(qy0 + 1, H + 1);
        const float x0f = t_u2f((unsigned)qx0), y0f = t_u2f((unsigned)qy0);
        ax1 = DVSG_SUB(DVSG_ADD(x0f, 1.0f), xp); ax0 = DVSG_SUB(xp, x0f);
        ay1 = DVSG_SUB(DVSG_ADD(y0f, 1.0f), yp); ay0 = DVSG_SUB(yp, y0f);
        const bool vx0 = zp_valid(qx0, W), vx1 = zp_valid(qx1, W), vy0 = zp_valid(qy0, H), vy1 = zp_valid(qy1, H);
        v00 = vx0 && vy0; v01 = vx1 && vy0; v10 = vx0 && vy1; v11 = vx1 && vy1;
        x0 = min(max(qx0, 1) - 1, W - 1); x1 = max(min(qx1, W) - 1, 0);   // keep addresses legal
        y0 = min(max(qy0, 1) - 1, H - 1); y1 = max(min(qy1, H) - 1, 0);
    }
    const float w00 = DVSG_MUL(ax1, ay1), w01 = DVSG_MUL(ax0, ay1), w10 = DVSG_MUL(ax1, ay0), w11 = DVSG_MUL(ax0, ay0);
    if (MODE == TMODE_TPS && mask_ptr) *mask_ptr = DVSG_ADD(DVSG_ADD(DVSG_ADD(w00, w10), w01), w11);   // A4 add_n order
    float i00[3], i01[3], i10[3], i11[3];
    if (WIDE_LD && (unsigned)(x1 - x0) <= 1u) {          // always, short of the int32 wrap of a coordinate beyond 2^31
        row_corners(srcb, y0, x0, x1, W, i00, i01);
        row_corners(srcb, y1, x0, x1, W, i10, i11);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            if (!v00) i00[ch] = 0.0f;
            if (!v01) i01[ch] = 0.0f;
            if (!v10) i10[ch] = 0.0f;
            if (!v11) i11[ch] = 0.0f;
        }
    } else {
        const float* a00 = srcb + ((size_t)y0 * W + x0) * 3;
        const float* a01 = srcb + ((size_t)y0 * W + x1) * 3;
        const float* a10 = srcb + ((size_t)y1 * W + x0) * 3;
        const float* a11 = srcb + ((size_t)y1 * W + x1) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            i00[ch] = v00 ? __ldg(a00 + ch) : 0.0f; i01[ch] = v01 ? __ldg(a01 + ch) : 0.0f;
            i10[ch] = v10 ? __ldg(a10 + ch) : 0.0f; i11[ch] = v11 ? __ldg(a11 + ch) : 0.0f;
        }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float t00 = DVSG_MUL(w00, i00[ch]), t01 = DVSG_MUL(w01, i01[ch]), t10 = DVSG_MUL(w10, i10[ch]), t11 = DVSG_MUL(w11, i11[ch]);
        float o;
        if (MODE == TMODE_TPS) o = DVSG_ADD(DVSG_ADD(DVSG_ADD(t00, t10), t01), t11);   // ThinPlateSpline.py:89
        else o = DVSG_ADD(DVSG_ADD(DVSG_ADD(t00, t01), t10), t11);                       // spatial_transformer.py:562
        optr[ch] = o;
    }
}

// ---- packed gather of one pixel pair (rows 2J, 2J+1 of the thread's column) ------------------------
// sb: staging-buffer address of frame pixel (0,0) (padded-frame modes: of padded pixel (0,0)); ot: this
// lane's column of the output tile.  All index quantities are exact fp32 integers below 2^22; floor() is
// FADD2.RM against 2^23 and the byte offset y*pitch + x*12 rides on the same constant, so the XU pipe
// (F2I / I2F) is not used at all.  Plain loads / stores (not volatile asm) and __restrict__ so that the
// compiler may overlap the pairs.  Sums of products are add2x (FFMA2 by an opaque 1.0): separately rounded.
// CLAMP = false: no corner of the tile touches the frame border (corners a, a+12, a+pitch, a+pitch+12).
// CLAMP = true : TPS sampler at the frame border, corners clamped first and weights taken FROM the clamped
//                corners (ThinPlateSpline.py:57-60, 81-88).
template <int MODE, bool CLAMP, bool MASK, int J>
__device__ __forceinline__ void gather_pair(const float2 xp, const float2 yp, const int pitch, const unsigned char* __restrict__ sb,
                                            float* __restrict__ ot, float2& msum, const float wm1, const float hm1, const float2 one) {
    const float2 one2 = f2dup(1.0f), m23 = f2dup(MAGIC23), pitchf = f2dup((float)pitch), twelve = f2dup(12.0f);
    float2 x0f, y0f, x1f, y1f;
    if (!CLAMP) {
        x0f = floor2_pos(xp); y0f = floor2_pos(yp);
        x1f = __fadd2_rn(x0f, one2); y1f = __fadd2_rn(y0f, one2);
    } else {
        const float2 fx = floor2_any(xp), fy = floor2_any(yp);
        const float2 gx = __fadd2_rn(fx, one2), gy = __fadd2_rn(fy, one2);
        x0f = f2(fminf(fmaxf(fx.x, 0.0f), wm1), fminf(fmaxf(fx.y, 0.0f), wm1));
        x1f = f2(fminf(fmaxf(gx.x, 0.0f), wm1), fminf(fmaxf(gx.y, 0.0f), wm1));
        y0f = f2(fminf(fmaxf(fy.x, 0.0f), hm1), fminf(fmaxf(fy.y, 0.0f), hm1));
        y1f = f2(fminf(fmaxf(gy.x, 0.0f), hm1), fminf(fmaxf(gy.y, 0.0f), hm1));
    }
    const float2 ax1 = sub2(x1f, xp), ax0 = sub2(xp, x0f), ay1 = sub2(y1f, yp), ay0 = sub2(yp, y0f);
    // 00 = (x0,y0), 01 = (x1,y0), 10 = (x0,y1), 11 = (x1,y1)
    const float2 w00 = __fmul2_rn(ax1, ay1), w01 = __fmul2_rn(ax0, ay1), w10 = __fmul2_rn(ax1, ay0), w11 = __fmul2_rn(ax0, ay0);
    if (MASK) msum = add2x(add2x(add2x(w00, w10, one), w01, one), w11, one);   // A4 add_n order (mask = warp of ones)
    // byte offsets as exact fp32 integers riding on 2^23: the float's bit pattern is 0x4B000000 + offset, and sb already has
    // 0x4B000000 subtracted, so pattern + sb is the corner's address (one LDS [R + UR + imm] per channel, no integer ops)
    const float2 tx0 = __ffma2_rn(x0f, twelve, m23);
    const float2 o00 = __ffma2_rn(y0f, pitchf, tx0);
    const float* __restrict__ p00a = reinterpret_cast<const float*>(sb + __float_as_uint(o00.x));
    const float* __restrict__ p00b = reinterpret_cast<const float*>(sb + __float_as_uint(o00.y));
    const float *__restrict__ p01a, *__restrict__ p01b, *__restrict__ p10a, *__restrict__ p10b, *__restrict__ p11a, *__restrict__ p11b;
    if (!CLAMP) {
        p01a = p00a + 3; p01b = p00b + 3;
        p10a = reinterpret_cast<const float*>(sb + pitch + __float_as_uint(o00.x));
        p10b = reinterpret_cast<const float*>(sb + pitch + __float_as_uint(o00.y));
        p11a = p10a + 3; p11b = p10b + 3;
    } else {
        const float2 tx1 = __ffma2_rn(x1f, twelve, m23);
        const float2 o01 = __ffma2_rn(y0f, pitchf, tx1), o10 = __ffma2_rn(y1f, pitchf, tx0), o11 = __ffma2_rn(y1f, pitchf, tx1);
        p01a = reinterpret_cast<const float*>(sb + __float_as_uint(o01.x));
        p01b = reinterpret_cast<const float*>(sb + __float_as_uint(o01.y));
        p10a = reinterpret_cast<const float*>(sb + __float_as_uint(o10.x));
        p10b = reinterpret_cast<const float*>(sb + __float_as_uint(o10.y));
        p11a = reinterpret_cast<const float*>(sb + __float_as_uint(o11.x));
        p11b = reinterpret_cast<const float*>(sb + __float_as_uint(o11.y));
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float2 i00 = f2(p00a[ch], p00b[ch]), i01 = f2(p01a[ch], p01b[ch]);
        const float2 i10 = f2(p10a[ch], p10b[ch]), i11 = f2(p11a[ch], p11b[ch]);
        const float2 t00 = __fmul2_rn(w00, i00), t01 = __fmul2_rn(w01, i01), t10 = __fmul2_rn(w10, i10), t11 = __fmul2_rn(w11, i11);
        float2 o;
        if (MODE == TMODE_TPS) o = add2x(add2x(add2x(t00, t10, one), t01, one), t11, one);   // ThinPlateSpline.py:89
        else o = add2x(add2x(add2x(t00, t01, one), t10, one), t11, one);                       // spatial_transformer.py:562
        ot[(2 * J) * TC * 3 + ch] = o.x;
        ot[(2 * J + 1) * TC * 3 + ch] = o.y;
    }
}

// ---- TPS tiles whose footprint fits no staging box: one pixel PAIR per call, straight from global memory ---------------
// The arithmetic of gather_pair<CLAMP = true> (corners clamped first, weights from the clamped corners, add_n order; packed
// fp32x2) with global corner addresses: half the instructions per pixel of general_pixel, identical bits.  Coordinates must
// be "sane" (|x| < 2^22, the footprint code checks it); anything else goes through general_pixel, which also reproduces the
// int32 wrap of the reference's cast.
template <bool MASK>
__device__ __noinline__ void general_pair_tps(const float2 xp, const float2 yp, const int W, const float wm1, const float hm1,
                                              const float* __restrict__ srcb, float* __restrict__ opair, float* __restrict__ mask0,
                                              float* __restrict__ mask1, const float2 one) {
    const float2 one2 = f2dup(1.0f);
    const float2 fx = floor2_any(xp), fy = floor2_any(yp);
    const float2 gx = __fadd2_rn(fx, one2), gy = __fadd2_rn(fy, one2);
    const float2 x0f = f2(fminf(fmaxf(fx.x, 0.0f), wm1), fminf(fmaxf(fx.y, 0.0f), wm1));
    const float2 x1f = f2(fminf(fmaxf(gx.x, 0.0f), wm1), fminf(fmaxf(gx.y, 0.0f), wm1));
    const float2 y0f = f2(fminf(fmaxf(fy.x, 0.0f), hm1), fminf(fmaxf(fy.y, 0.0f), hm1));
    const float2 y1f = f2(fminf(fmaxf(gy.x, 0.0f), hm1), fminf(fmaxf(gy.y, 0.0f), hm1));
    const float2 ax1 = sub2(x1f, xp), ax0 = sub2(xp, x0f), ay1 = sub2(y1f, yp), ay0 = sub2(yp, y0f);
    const float2 w00 = __fmul2_rn(ax1, ay1), w01 = __fmul2_rn(ax0, ay1), w10 = __fmul2_rn(ax1, ay0), w11 = __fmul2_rn(ax0, ay0);
    if (MASK) {
        const float2 ms = add2x(add2x(add2x(w00, w10, one), w01, one), w11, one);   // A4 add_n order
        if (mask0) *mask0 = ms.x;
        if (mask1) *mask1 = ms.y;
    }
    // exact small integers: the float -> int conversion is a bit trick (no F2I on the XU pipe)
    const float2 m23 = f2dup(MAGIC23);
    const float2 bx0 = __fadd2_rn(x0f, m23), bx1 = __fadd2_rn(x1f, m23), by0 = __fadd2_rn(y0f, m23), by1 = __fadd2_rn(y1f, m23);
    const int xa0 = __float_as_int(bx0.x) & 0x7fffff, xa1 = __float_as_int(bx1.x) & 0x7fffff, ya0 = __float_as_int(by0.x) & 0x7fffff,
              ya1 = __float_as_int(by1.x) & 0x7fffff;
    const int xb0 = __float_as_int(bx0.y) & 0x7fffff, xb1 = __float_as_int(bx1.y) & 0x7fffff, yb0 = __float_as_int(by0.y) & 0x7fffff,
              yb1 = __float_as_int(by1.y) & 0x7fffff;
    const float* __restrict__ ra0 = srcb + (size_t)ya0 * W * 3;
    const float* __restrict__ ra1 = srcb + (size_t)ya1 * W * 3;
    const float* __restrict__ rb0 = srcb + (size_t)yb0 * W * 3;
    const float* __restrict__ rb1 = srcb + (size_t)yb1 * W * 3;
    const float *p00a = ra0 + xa0 * 3, *p01a = ra0 + xa1 * 3, *p10a = ra1 + xa0 * 3, *p11a = ra1 + xa1 * 3;
    const float *p00b = rb0 + xb0 * 3, *p01b = rb0 + xb1 * 3, *p10b = rb1 + xb0 * 3, *p11b = rb1 + xb1 * 3;
    float2 i00[3], i01[3], i10[3], i11[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        i00[ch] = f2(__ldg(p00a + ch), __ldg(p00b + ch)); i01[ch] = f2(__ldg(p01a + ch), __ldg(p01b + ch));
        i10[ch] = f2(__ldg(p10a + ch), __ldg(p10b + ch)); i11[ch] = f2(__ldg(p11a + ch), __ldg(p11b + ch));
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float2 t00 = __fmul2_rn(w00, i00[ch]), t01 = __fmul2_rn(w01, i01[ch]), t10 = __fmul2_rn(w10, i10[ch]), t11 = __fmul2_rn(w11, i11[ch]);
        const float2 o = add2x(add2x(add2x(t00, t10, one), t01, one), t11, one);   // ThinPlateSpline.py:89
        opair[ch] = o.x;
        opair[TC * 3 + ch] = o.y;
    }
}

// gather + blend of the thread's 8 pixels from the staged footprint (one variant per tile, warp-uniform)
template <int MODE, bool CLAMP, bool MASK>
__device__ __forceinline__ void gather_tile(const float2 (&XC)[TR / 2], const float2 (&YC)[TR / 2], const int pitch,
                                            const unsigned char* __restrict__ sb, float* __restrict__ ot, const float wm1, const float hm1,
                                            const float2 one, float* __restrict__ mask_col, const int ow, const int rows_ok) {
    float2 ms[TR / 2];
    gather_pair<MODE, CLAMP, MASK, 0>(XC[0], YC[0], pitch, sb, ot, ms[0], wm1, hm1, one);
    gather_pair<MODE, CLAMP, MASK, 1>(XC[1], YC[1], pitch, sb, ot, ms[1], wm1, hm1, one);
    gather_pair<MODE, CLAMP, MASK, 2>(XC[2], YC[2], pitch, sb, ot, ms[2], wm1, hm1, one);
    gather_pair<MODE, CLAMP, MASK, 3>(XC[3], YC[3], pitch, sb, ot, ms[3], wm1, hm1, one);
    if (MASK && mask_col) {      // mask_col == nullptr: column past the frame edge
#pragma unroll
        for (int j = 0; j < TR / 2; ++j) {
            if (2 * j < rows_ok) mask_col[(2 * j) * ow] = ms[j].x;
            if (2 * j + 1 < rows_ok) mask_col[(2 * j + 1) * ow] = ms[j].y;
        }
    }
}

// G > 0: kernel specialised for G x G meshes; the compact separable-mesh tables are used when the frame's mesh is
// separable (checked in the prologue), the generic records otherwise.  G == 0: any mesh, generic records.  Both evaluate
// every radial term per pixel (DVSG_FLAG_TPS_EXACT).  G < 0: tile-node evaluation (tile_common.cuh), the default: -1 any
// mesh, -4 / -5 / -16 with the separable node pass for G x G meshes (checked per frame in the prologue).
template <int MODE, bool MASK, int G>
__global__ void __launch_bounds__(TNT, 5) warp_fwd_tile_kernel(const TileParams p, const __grid_constant__ TileMaps maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_mbar[TNW];
    __shared__ float s_lin[12];
    __shared__ __align__(16) float s_yt[TR];
    __shared__ float s_T[MODE == TMODE_TPS ? 2 * TFUSE_N : 1];      // coefficients of this frame when the solve is fused (online loop)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler: tile bookkeeping lives in uniform registers
    const int H = p.H, W = p.W, oh = p.oh, ow = p.ow;
    const int seg = blockIdx.x, b = blockIdx.z;      // grid = (CTAs per strip, strips per frame, frames)
    const int row0 = blockIdx.y * TR;
    const int rows_ok = min(TR, oh - row0);
    const int t_begin = seg * p.seg_len, t_end = min(t_begin + p.seg_len, p.n_tx);

    unsigned char* w_out = smem + (size_t)warp * (TOUT_BYTES + p.stage_bytes);
    unsigned char* w_stage = w_out + TOUT_BYTES;
    const unsigned char* recs = smem + (size_t)TNW * (TOUT_BYTES + p.stage_bytes);
    const uint32_t out_s = smem_u32(w_out), stage_s = smem_u32(w_stage), mbar = smem_u32(&s_mbar[warp]);
    const int pn8 = (p.pn + 7) & ~7;                 // table padded with zero-weight records to a multiple of 8
    constexpr bool NODES = MODE == TMODE_TPS && G < 0;
    constexpr int NG = G < -1 ? -G : 0;              // side of the separable mesh the node pass is specialised for
    // node mode: [node tables][per-warp exchange buffers of 32 float2] follow the per-warp staging
    const NodeTables nt = node_tables_at(smem + (size_t)TNW * (TOUT_BYTES + p.stage_bytes), p.pn);
    float2* const w_nodes = reinterpret_cast<float2*>(smem + (size_t)TNW * (TOUT_BYTES + p.stage_bytes) + node_tables_bytes(p.pn)) + warp * 32;

    // ---- prologue: mbarriers, per-strip tables (the only CTA barrier of the kernel) -----------------
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    if (tid < TR) s_yt[tid] = lin_coord(min(row0 + tid, oh - 1), p.step_y);
    bool sep = false;
    if (MODE == TMODE_TPS) {
        unsigned char* tab = smem + (size_t)TNW * (TOUT_BYTES + p.stage_bytes);
        const float* cb = p.coord + (size_t)b * p.coord_stride;
        const float* Tb = p.T + (size_t)b * 2 * (p.pn + 3);
        if (p.winv != nullptr) {      // fused prepared solve: one launch per call of the online loop
            tile_fused_solve(p.winv, cb, p.vec + (size_t)b * p.pn * 2, p.pn, tid, s_T);
            __syncthreads();
            if (blockIdx.x == 0 && blockIdx.y == 0 && tid < 2 * (p.pn + 3)) p.T_out[(size_t)b * 2 * (p.pn + 3) + tid] = s_T[tid];
            Tb = s_T;
        }
        if (NODES) tile_node_tables<NG>(Tb, cb, p.pn, row0, oh, p.step_x, p.step_y, tid, TNT, s_lin, nt, t_begin * TC, p.seg_len == SNODE_PER_CTA * SNODE_TILES);
        else if (G > 0) sep = tile_tps_tables_sep<(G > 0 ? G : 1)>(Tb, cb, pn8, row0, oh, p.step_y, tid, TNT, s_lin, tab);
        else tile_tps_tables(Tb, cb, p.pn, pn8, row0, oh, p.step_y, tid, TNT, s_lin, reinterpret_cast<TpsRec*>(tab));
    } else if (MODE == TMODE_HOMOG) {
        const int nt = p.projective ? 8 : 6;
        if (tid < 9) s_lin[tid] = tid < nt ? __ldg(p.theta + (size_t)b * nt + tid) : (tid == 8 ? 1.0f : 0.0f);
    }
    __syncthreads();

    // node mode: this lane's node (lanes past the last node repeat it; their values are never read)
    const int node_l = min(lane, NNX * NNY - 1);
    const float node_xoff = NODES ? NODE_XOFF[node_l % NNX] : 0.0f;
    const float node_yn = NODES ? fmaf(p.step_y, (float)row0 + NODE_YOFF[node_l / NNX], -1.0f) : 0.0f;
    float node_lx[NNX];                              // Lagrange weights of this lane's column; reloaded for a ragged last tile
    if (NODES) node_load_lx(lane, node_lx);

    const float* srcb = p.src + (size_t)b * H * W * 3;
    const float2 one2 = f2dup(1.0f), onex = f2dup(p.one);
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    float* const ot = reinterpret_cast<float*>(w_out) + lane * 3;      // this lane's column of the output tile
    unsigned phase = 0;
    bool out_pending = false;

    struct Tile {               // warp-uniform description of a tile whose coordinates are known
        int col0, col;
        bool col_ok, staged, interior, sane;
        int pitch;
        const unsigned char* sb;
    };

    // ---- pieces of the per-tile work -----------------------------------------------------------------
    // coordinates, part 1.  TPS: affine part + basis.  GIVEN / FLOW: the raw loads only (consumed after the
    // current tile's gather, so their latency hides behind it).
    auto coords_begin = [&](const int tt, Tile& T, float& xt, float2 (&X)[TR / 2], float2 (&Y)[TR / 2], float (&rx)[TR], float (&ry)[TR]) {
        T.col0 = tt * TC;
        T.col = min(T.col0 + lane, ow - 1);          // columns / rows past the edge are duplicates of the edge pixel
        T.col_ok = T.col0 + lane < ow;
        xt = lin_coord(T.col, p.step_x);
        if (NODES) {
            if (T.col0 + TC > ow) node_load_lx(min(lane, ow - 1 - T.col0), node_lx);      // the strip's last tile = this warp's last tile
            tile_node_coords<NG>(nt, p.pn, s_lin, s_yt, T.col0, node_lx, xt, node_xoff, node_yn, p.step_x, rows_ok, w_nodes, lane, X, Y);
        } else if (MODE == TMODE_TPS) {
            const float bx = fmaf(s_lin[1], xt, s_lin[0]), by = fmaf(s_lin[4], xt, s_lin[3]);
            const float2 l2 = f2dup(s_lin[2]), l5 = f2dup(s_lin[5]);
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                const float2 ytp = *reinterpret_cast<const float2*>(s_yt + 2 * j);
                X[j] = __ffma2_rn(l2, ytp, f2dup(bx));
                Y[j] = __ffma2_rn(l5, ytp, f2dup(by));
            }
            if (G > 0 && sep) tile_tps_basis_sep<(G > 0 ? G : 1)>(recs, xt, X, Y);
            else tile_tps_basis(recs, pn8, xt, X, Y);
        } else if (MODE == TMODE_GIVEN || MODE == TMODE_FLOW) {
#pragma unroll
            for (int q = 0; q < TR; ++q) {
                const int row = min(row0 + q, oh - 1);
                const size_t i = ((size_t)b * oh + row) * ow + T.col;
                if (MODE == TMODE_GIVEN) { rx[q] = __ldg(p.x_in + i); ry[q] = __ldg(p.y_in + i); }
                else { const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + i); rx[q] = f.x; ry[q] = f.y; }
            }
        }
    };
    // coordinates, last part: x / y outputs and the sampler's own coordinate convention
    //   TPS: A4 pixel-space coordinate; others: clipped+1 coordinate in the zero-padded frame
    // (XO, YO may alias X, Y: the software pipeline passes the current tile's arrays, free after its gather, so that the next
    //  tile's coordinates need no register copy at the end of the iteration)
    auto coords_end = [&](const Tile& T, const float xt, const float2 (&X)[TR / 2], const float2 (&Y)[TR / 2], float2 (&XO)[TR / 2], float2 (&YO)[TR / 2],
                          float (&rx)[TR], float (&ry)[TR]) {
        if (MODE == TMODE_TPS) {
            if (p.x_out && T.col_ok) {
#pragma unroll
                for (int j = 0; j < TR / 2; ++j) {
                    const size_t i = ((size_t)b * oh + row0 + 2 * j) * ow + T.col;
                    if (2 * j < rows_ok) { p.x_out[i] = X[j].x; p.y_out[i] = Y[j].x; }
                    if (2 * j + 1 < rows_ok) { p.x_out[i + ow] = X[j].y; p.y_out[i + ow] = Y[j].y; }
                }
            }
            // x_pix = ((x + 1) * W) / 2   (ThinPlateSpline.py:48-49), separately rounded
            const float2 wf = f2dup((float)W), hf = f2dup((float)H), half2 = f2dup(0.5f);
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                XO[j] = __fmul2_rn(__fmul2_rn(__fadd2_rn(X[j], one2), wf), half2);
                YO[j] = __fmul2_rn(__fmul2_rn(__fadd2_rn(Y[j], one2), hf), half2);
            }
        } else {
            if (MODE == TMODE_GIVEN || MODE == TMODE_FLOW) {
#pragma unroll
                for (int q = 0; q < TR; ++q) {
                    if (MODE == TMODE_GIVEN) { rx[q] = zp_pix_from_norm(rx[q], W); ry[q] = zp_pix_from_norm(ry[q], H); }
                    else { rx[q] = DVSG_ADD((float)T.col, rx[q]); ry[q] = DVSG_ADD((float)min(row0 + q, oh - 1), ry[q]); }   // warp_with_optical_flow.py:107-120
                }
            } else {
#pragma unroll
                for (int q = 0; q < TR; ++q) {
                    const float yt = s_yt[q];
                    // rows of theta @ [x_t; y_t; 1], accumulated k = 0,1,2 (spatial_transformer.py:437)
                    float xn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[0], xt), DVSG_MUL(s_lin[1], yt)), s_lin[2]);
                    float yn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[3], xt), DVSG_MUL(s_lin[4], yt)), s_lin[5]);
                    if (p.projective) {
                        const float zn = DVSG_ADD(DVSG_ADD(DVSG_MUL(s_lin[6], xt), DVSG_MUL(s_lin[7], yt)), s_lin[8]);
                        xn = zn != 0.0f ? DVSG_DIV(xn, zn) : 0.0f;   // tf.div_no_nan, :446-447
                        yn = zn != 0.0f ? DVSG_DIV(yn, zn) : 0.0f;
                    }
                    if (p.x_out && T.col_ok && q < rows_ok) {
                        const size_t i = ((size_t)b * oh + row0 + q) * ow + T.col;
                        p.x_out[i] = xn; p.y_out[i] = yn;
                    }
                    rx[q] = zp_pix_from_norm(xn, W); ry[q] = zp_pix_from_norm(yn, H);
                }
            }
            // clip to [-1, W] and shift into the zero-padded frame (spatial_transformer.py:517-521)
            const float wf = (float)W, hf = (float)H;
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                XO[j] = f2(DVSG_ADD(fminf(fmaxf(rx[2 * j], -1.0f), wf), 1.0f), DVSG_ADD(fminf(fmaxf(rx[2 * j + 1], -1.0f), wf), 1.0f));
                YO[j] = f2(DVSG_ADD(fminf(fmaxf(ry[2 * j], -1.0f), hf), 1.0f), DVSG_ADD(fminf(fmaxf(ry[2 * j + 1], -1.0f), hf), 1.0f));
            }
        }
    };
    // F + L: footprint of the tile (warp-uniform after the REDUX) and ONE TMA tensor copy of it into the staging buffer
    auto footprint_load = [&](const float2 (&XP)[TR / 2], const float2 (&YP)[TR / 2], Tile& T) {
        const float xmn = min8n(XP), xmx = max8n(XP), ymn = min8n(YP), ymx = max8n(YP);
        // finite and small enough for the 2^23 arithmetic; NaN fails the comparison
        const bool sane = (fabsf(xmn) + fabsf(xmx)) + (fabsf(ymn) + fabsf(ymx)) < 4.0e6f;
        int x_lo = sane ? floor_small(xmn) : -(1 << 30), x_hi = sane ? floor_small(xmx) + 1 : (1 << 30);
        int y_lo = sane ? floor_small(ymn) : -(1 << 30), y_hi = sane ? floor_small(ymx) + 1 : (1 << 30);
        x_lo = __reduce_min_sync(0xffffffffu, x_lo); x_hi = __reduce_max_sync(0xffffffffu, x_hi);
        y_lo = __reduce_min_sync(0xffffffffu, y_lo); y_hi = __reduce_max_sync(0xffffffffu, y_hi);
        const bool all_sane = x_lo > -(1 << 30);
        int fx_lo, fx_hi, fy_lo, fy_hi;    // source pixels to stage (inclusive; may reach outside the frame for the padded modes)
        if (MODE == TMODE_TPS) {
            T.interior = x_lo >= 0 && x_hi <= W - 1 && y_lo >= 0 && y_hi <= H - 1;
            fx_lo = min(max(x_lo, 0), W - 1); fx_hi = min(max(x_hi, 0), W - 1);
            fy_lo = min(max(y_lo, 0), H - 1); fy_hi = min(max(y_hi, 0), H - 1);
        } else {
            // padded-frame corners lie in [x_lo, x_hi]; real pixel = idx - 1.  Whatever falls outside the frame is
            // zero-filled by the TMA copy -- exactly the zero padding of the reference -- so every tile is "interior"
            T.interior = true;
            fx_lo = x_lo - 1; fx_hi = x_hi - 1; fy_lo = y_lo - 1; fy_hi = y_hi - 1;
        }
        // the box starts at a 16-byte aligned float (TMA faults on unaligned box origins)
        const int fx0 = (fx_lo * 3) & ~3;
        const int fw = (fx_hi + 1) * 3 - fx0, nrows = fy_hi - fy_lo + 1;
        // box 0 / 1: bw[0] wide, bh[0] / bh[1] rows; box 2: bw[2] x bh[2] (floats x rows; a box disabled by the host has height 0)
        const bool wide = fw > p.bw[0];
        const int box = wide ? 2 : (nrows <= p.bh[0] ? 0 : 1);
        const int box_rows = wide ? p.bh[2] : (nrows <= p.bh[0] ? p.bh[0] : p.bh[1]);
        T.pitch = (wide ? p.bw[2] : p.bw[0]) * 4;
        // p.bh[] are 0 for frames whose byte offsets would not fit the packed gather's exact fp32 integers (< 2^22)
        T.staged = all_sane && fw <= p.bw[2] && nrows <= box_rows;
        T.sane = all_sane;
        // sb = staging address of frame pixel (0,0) minus the bit pattern of 2^23 (see gather_pair); padded-frame modes
        // index pixel idx-1: fold the -1 into it
        T.sb = w_stage - (fy_lo * T.pitch + fx0 * 4) - (MODE == TMODE_TPS ? 0 : T.pitch + 12) - 0x4B000000ll;
        if (T.staged && lane == 0) {
            mbar_arrive_expect_tx(mbar, (unsigned)(T.pitch * box_rows));
            tma_load_3d(stage_s, &maps.src[box], fx0, fy_lo, b, mbar);
        }
        if ((MODE == TMODE_TPS || MODE == TMODE_HOMOG) && !T.staged && all_sane && !(p.dbg & 4)) {
            // per-pixel tile of a smooth map (rare; rough given-grid / flow fields would only double their L2 requests):
            // its gathers are issued a whole coordinate phase from now -- pull the two source rows of every pixel into
            // L2 meanwhile (the x1 corner shares the line of x0 in all but 1 of 10 cases)
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float xq = h ? XP[j].y : XP[j].x, yq = h ? YP[j].y : YP[j].x;
                    const int off = MODE == TMODE_TPS ? 0 : 1;
                    const int xi = min(max(floor_small(xq) - off, 0), W - 1), yi = min(max(floor_small(yq) - off, 0), H - 1);
                    const float* a0 = srcb + ((size_t)yi * W + xi) * 3;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a0));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + (yi + 1 < H ? 3 * W : 0)));
                }
            }
        }
    };

    // ---- the software pipeline: the coordinates of tile t+1 are computed BEFORE tile t is gathered, so the TMA copy of
    // a footprint has a whole coordinate phase to land before the warp waits for it
    int t = t_begin + warp;
    if (t < t_end) {
        Tile cur, nxt;
        float2 XC[TR / 2], YC[TR / 2], XN[TR / 2], YN[TR / 2];
        float rx[TR], ry[TR], xt;
        coords_begin(t, cur, xt, XC, YC, rx, ry);
        coords_end(cur, xt, XC, YC, XC, YC, rx, ry);
        footprint_load(XC, YC, cur);
        while (true) {
            const int tn = t + TNW;
            const bool has_next = tn < t_end;
            if (has_next) coords_begin(tn, nxt, xt, XN, YN, rx, ry);
            if (out_pending) {             // the previous tile's tensor store must have read the output tile
                if (lane == 0) bulk_wait_read0();
                out_pending = false;
            }
            __syncwarp();
            float* mask_col = (MASK && cur.col_ok) ? p.mask_out + ((size_t)b * oh + row0) * ow + cur.col : nullptr;
            if (cur.staged) {
                mbar_wait(mbar, phase); phase ^= 1u;
                if (MODE != TMODE_TPS || cur.interior)
                    gather_tile<MODE, false, MASK>(XC, YC, cur.pitch, cur.sb, ot, wm1, hm1, onex, mask_col, ow, rows_ok);
                else
                    gather_tile<MODE, true, MASK>(XC, YC, cur.pitch, cur.sb, ot, wm1, hm1, onex, mask_col, ow, rows_ok);
            } else if (MODE == TMODE_TPS && cur.sane && !(p.dbg & 8)) {
                // TPS tile that fits no box (a stretched or sheared neighbourhood): pixel pairs from global memory
#pragma unroll
                for (int j = 0; j < TR / 2; ++j)
                    general_pair_tps<MASK>(XC[j], YC[j], W, wm1, hm1, srcb, ot + 2 * j * TC * 3, (MASK && mask_col && 2 * j < rows_ok) ? mask_col + 2 * j * ow : nullptr,
                                           (MASK && mask_col && 2 * j + 1 < rows_ok) ? mask_col + (2 * j + 1) * ow : nullptr, onex);
            } else {
                // wide corner loads pay for a rough flow field (white-noise +-8 px flow: 60 -> 74 Gpix/s); the neighbouring
                // lanes of a smooth map already share sectors, where the selects only cost (+-0.3 TPS: 137 -> 128 Gpix/s)
#pragma unroll
                for (int q = 0; q < TR; ++q) {
                    const float xq = (q & 1) ? XC[q >> 1].y : XC[q >> 1].x, yq = (q & 1) ? YC[q >> 1].y : YC[q >> 1].x;
                    general_pixel<MODE, MODE == TMODE_FLOW>(xq, yq, W, H, srcb, ot + q * TC * 3, (MASK && mask_col && q < rows_ok) ? mask_col + q * ow : nullptr);
                }
            }

            // S: output tile -> global with one TMA tensor store (clipped at the frame edge)
            fence_proxy_async_smem();      // this lane's generic-proxy writes -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&maps.out, cur.col0 * 3, row0, b, out_s);
                bulk_commit();
            }
            out_pending = true;
            if (!has_next) break;

            coords_end(nxt, xt, XN, YN, XC, YC, rx, ry);      // straight into the current tile's arrays (free after its gather)
            footprint_load(XC, YC, nxt);   // the staging buffer is free: every lane passed the __syncwarp above
            cur = nxt;
            t = tn;
        }
    }
    if (out_pending && lane == 0) bulk_wait_read0();   // shared memory must outlive the store's reads
}

// ---- host side ---------------------------------------------------------------------------------
static float tile_lin_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }
static int g_tile_dbg = 0;
// staging boxes of the forward kernel (floats x rows): box 0 = bw0 x bh0, box 1 = bw0 x bh1, box 2 = bw2 x bh2.
// TPS / homography (smooth maps): 128 x 10 (5 KB) covers 60 % of the tiles of a +-0.1 TPS warp at 720p; 128 x 15 and
// 160 x 12 (7.5 KB, the most that still leaves 5 CTAs per SM) cover all but ~0.05 % -- a tile that fits no box takes
// the per-pixel path, which costs about six staged tiles, so the tall boxes are worth their extra TMA bytes
// (measured: 63.1 % -> 67.5 % of HBM).  Given-grid and flow samplers take arbitrary fields: their boxes stay at
// 128 x 13 / 160 x 10, because the per-pixel path of a rough field lives on the L1 cache that a larger shared-memory
// carve-out takes away (white-noise +-8 px flow: 0.57 ms with the small boxes, 1.17 ms with the large ones).
struct BoxSet { int bw0, bh0, bh1, bw2, bh2; };
static BoxSet g_box_smooth = {128, 10, 15, 160, 12}, g_box_field = {BOX_W0, BOX_H0, BOX_H1, BOX_W2, BOX_H2};
static void tile_env_once() {
    static bool done = false;
    if (done) return;
    done = true;
    if (const char* e = getenv("DVSG_TILE_BOXES")) {       // experiments only: "bw0,bh0,bh1,bw2,bh2"
        int v[5];
        if (sscanf(e, "%d,%d,%d,%d,%d", v, v + 1, v + 2, v + 3, v + 4) == 5 && v[0] % 32 == 0 && v[3] % 4 == 0)
            g_box_smooth = g_box_field = BoxSet{v[0], v[1], v[2], v[3], v[4]};
    }
}

bool tile_path_ok(const void* src, const void* out, int H, int W, int C, int oh, int ow, int pn_or_0) {
    return C == 3 && W % 4 == 0 && ow % 4 == 0 && ow >= TC && oh >= TR && aligned16(src) && aligned16(out) && W < (1 << 20) && H < (1 << 20) &&
           (long long)H * W < (1LL << 28) && (long long)oh * ow < (1LL << 28) && pn_or_0 <= TKC;
}

template <int MODE>
static int launch_tile(TileParams p, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    tile_env_once();
    const BoxSet bx = (MODE == TMODE_TPS || MODE == TMODE_HOMOG) ? g_box_smooth : g_box_field;
    p.stage_bytes = 4 * max(bx.bw0 * max(bx.bh0, bx.bh1), bx.bw2 * bx.bh2);
    p.dbg = g_tile_dbg;
    p.one = 1.0f;
    p.n_tx = (p.ow + TC - 1) / TC;
    p.n_ty = (p.oh + TR - 1) / TR;
    const long long strips = (long long)p.B * p.n_ty;
    // tiles per CTA (tile_pick_seg_len): the TPS prologue builds the per-strip tables, the field samplers have none
    p.seg_len = tile_pick_seg_len(strips, p.n_tx, 148 * 5, MODE == TMODE_TPS ? 0.75 : 0.2, "DVSG_FWD_SEGLEN");
    // 16 x 16 meshes in node mode: SNODE_PER_CTA super-tiles per CTA (two-level evaluation, tile_common.cuh); DVSG_TPS_ONE_LEVEL: A/B
    if (MODE == TMODE_TPS && p.nodes && p.pn == TKS * TKS && !getenv("DVSG_TPS_ONE_LEVEL")) p.seg_len = SNODE_PER_CTA * SNODE_TILES;
    p.segs = (p.n_tx + p.seg_len - 1) / p.seg_len;
    DVSG_REQUIRE(p.B <= 65535 && p.n_ty <= 65535, "tile kernel: batch %d / %d strips exceed the grid limits: split the call", p.B, p.n_ty);
    TileMaps maps;
    // the packed gather forms byte offsets y*pitch + x*12 as exact fp32 integers below 2^22 (corners reach one pixel
    // past the frame in the padded modes); larger frames take the per-pixel path
    const bool offsets_fit = (long long)(p.H + 3) * (max(bx.bw0, bx.bw2) * 4) + (long long)(p.W + 3) * 12 < (1LL << 22);
    for (int i = 0; i < NBOX; ++i) {
        p.bw[i] = min(i == 2 ? bx.bw2 : bx.bw0, 3 * p.W);      // a box may not exceed the tensor (tiny frames)
        p.bh[i] = min(i == 0 ? bx.bh0 : (i == 1 ? bx.bh1 : bx.bh2), p.H);
        const int rc = encode_frames(&maps.src[i], p.src, p.B, p.H, p.W, p.bw[i], p.bh[i]);
        if (rc) return rc;
        if (!offsets_fit || (i > 0 && (p.dbg & i))) p.bh[i] = 0;     // box disabled (debug mask: bit 0 = box 1, bit 1 = box 2)
    }
    const int rc = encode_frames(&maps.out, p.out, p.B, p.oh, p.ow, TC * 3, TR);
    if (rc) return rc;
    const bool nodes = MODE == TMODE_TPS && p.nodes;
    const size_t smem = (size_t)TNW * (TOUT_BYTES + p.stage_bytes) +
                        (nodes ? node_tables_bytes(p.pn) + (size_t)TNW * 32 * sizeof(float2) : (MODE == TMODE_TPS ? (size_t)((p.pn + 7) & ~7) * sizeof(TpsRec) : 0));
    const dim3 grid((unsigned)p.segs, (unsigned)p.n_ty, (unsigned)p.B);
    auto go = [&](auto k) {
        ensure_dynamic_smem(reinterpret_cast<const void*>(k), (int)smem);
        k<<<grid, TNT, smem, st>>>(p, maps);
    };
    if constexpr (MODE == TMODE_TPS) {
        const bool m = p.mask_out != nullptr;
        if (nodes && p.pn == 16) { if (m) go(warp_fwd_tile_kernel<MODE, true, -4>); else go(warp_fwd_tile_kernel<MODE, false, -4>); }
        else if (nodes && p.pn == 25) { if (m) go(warp_fwd_tile_kernel<MODE, true, -5>); else go(warp_fwd_tile_kernel<MODE, false, -5>); }
        else if (nodes && p.pn == 256) { if (m) go(warp_fwd_tile_kernel<MODE, true, -16>); else go(warp_fwd_tile_kernel<MODE, false, -16>); }
        else if (nodes) { if (m) go(warp_fwd_tile_kernel<MODE, true, -1>); else go(warp_fwd_tile_kernel<MODE, false, -1>); }
        else if (p.pn == 16) { if (m) go(warp_fwd_tile_kernel<MODE, true, 4>); else go(warp_fwd_tile_kernel<MODE, false, 4>); }
        else if (p.pn == 25) { if (m) go(warp_fwd_tile_kernel<MODE, true, 5>); else go(warp_fwd_tile_kernel<MODE, false, 5>); }
        else { if (m) go(warp_fwd_tile_kernel<MODE, true, 0>); else go(warp_fwd_tile_kernel<MODE, false, 0>); }
    } else {
        go(warp_fwd_tile_kernel<MODE, false, 0>);
    }
    count_launch();
    return check_launch("warp_fwd_tile_kernel");
}

// Shapes for which the tile kernels evaluate the TPS map on tile nodes (tile_common.cuh) unless DVSG_FLAG_TPS_EXACT asks
// for the per-pixel evaluation: both tile kernels must be applicable (so that forward and backward see the same
// coordinates), the mesh large enough for the node pass to pay (9 instructions per control point and tile + ~130 for the
// interpolation, against 28 per control point), and a tile small against the frame (the error study covers >= 200 x 400).
// Measured (sustained, B200): 4x4 mesh 64 x 1080p +11 %, 64 x 720p +9 %, 5x5 at 720p +25 %, 16x16 at 4K 3.8x; at the
// 288 x 512 training shape (two waves of short CTAs, latency-bound) the 4 x 4 mesh LOSES 5 % (forward 43.3 vs 41.2 us,
// backward 110 vs 104) -- hence the pixel floor -- while the model's 5 x 5 mesh gains there (forward 45.3 vs 51.3 us, backward
// 142.5 vs 152.7, once the backward kernel's node mode kept its fourth CTA per SM): the floor applies below 25 points.
bool tps_nodes_ok(int H, int W, int C, int oh, int ow, int pn, int flags) {
    static const bool env_exact = getenv("DVSG_TPS_EXACT") != nullptr;      // A/B experiments
    static const bool env_force = getenv("DVSG_TPS_NODES_FORCE") != nullptr;      // experiments: no pixel floor
    (void)H;
    return !(flags & DVSG_FLAG_TPS_EXACT) && !env_exact && C == 3 && W % 4 == 0 && ow % 4 == 0 && pn >= 8 && pn <= TKC && ow >= 400 && oh >= 200 &&
           (env_force || pn >= 25 || (long long)oh * ow >= 500000);
}

int tile_tps(const float* U, const float* coord, long long cstride, const float* T, float* out, float* x_out, float* y_out,
             float* mask_out, int B, int H, int W, int oh, int ow, int pn, int flags, cudaStream_t st) {
    TileParams p = {};
    p.nodes = tps_nodes_ok(H, W, 3, oh, ow, pn, flags) ? 1 : 0;
    p.src = U; p.out = out; p.x_out = x_out; p.y_out = y_out; p.mask_out = mask_out;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = cstride; p.T = T; p.pn = pn;
    p.step_x = tile_lin_step(ow); p.step_y = tile_lin_step(oh);
    return launch_tile<TMODE_TPS>(p, st);
}

// ThinPlateSpline of the online loop in ONE launch: the prepared solve (coord + vector, W^-1 from dvsg_tps_prepare) runs in the
// prologue of every CTA -- 2N threads x pn fp64 multiply-adds, the arithmetic of tps_apply_kernel, so T and the frames are
// bit-identical to dvsg_tps_solve_offsets_prepared followed by dvsg_tps_warp_fwd
int tile_tps_fused(const float* U, const float* coord, const float* vector, const double* winv, float* T_out, float* out, float* x_out,
                   float* y_out, float* mask_out, int B, int H, int W, int oh, int ow, int pn, int flags, cudaStream_t st) {
    TileParams p = {};
    p.nodes = tps_nodes_ok(H, W, 3, oh, ow, pn, flags) ? 1 : 0;
    p.src = U; p.out = out; p.x_out = x_out; p.y_out = y_out; p.mask_out = mask_out;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = 0; p.T = T_out; p.pn = pn;
    p.winv = winv; p.vec = vector; p.T_out = T_out;
    p.step_x = tile_lin_step(ow); p.step_y = tile_lin_step(oh);
    return launch_tile<TMODE_TPS>(p, st);
}

int tile_given(const float* im, const float* x, const float* y, float* out, int B, int H, int W, int oh, int ow, cudaStream_t st) {
    TileParams p = {};
    p.src = im; p.out = out; p.x_in = x; p.y_in = y;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    return launch_tile<TMODE_GIVEN>(p, st);
}

int tile_flow(const float* im, const float* flow, float* out, int B, int H, int W, cudaStream_t st) {
    TileParams p = {};
    p.src = im; p.out = out; p.flow = flow;
    p.B = B; p.H = H; p.W = W; p.oh = H; p.ow = W;
    return launch_tile<TMODE_FLOW>(p, st);
}

int tile_homog(const float* im, const float* theta, int projective, float* out, float* x_out, float* y_out, int B, int H, int W,
               int oh, int ow, cudaStream_t st) {
    TileParams p = {};
    p.src = im; p.out = out; p.x_out = x_out; p.y_out = y_out; p.theta = theta; p.projective = projective;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    p.step_x = tile_lin_step(ow); p.step_y = tile_lin_step(oh);
    return launch_tile<TMODE_HOMOG>(p, st);
}

void tile_set_tuning(int stage_bytes, int target_ctas, int minb) {
    if (stage_bytes >= 0 && stage_bytes < 16) g_tile_dbg = stage_bytes;      // debug mask (the staging buffer is sized by the largest TMA box)
    (void)target_ctas;      // superseded by tile_pick_seg_len
    (void)minb;
}

}  // namespace dvsg

// tuning ABI (exported, not in the public header): the tiles-per-CTA cut of a launch, so that the host logic is testable
// without a device
extern "C" int dvsg_debug_seg_len(long long strips, int n_tx, int slots, double cta_cost) {
    return dvsg::tile_pick_seg_len(strips, n_tx, slots, cta_cost, "DVSG_DEBUG_SEGLEN");
}
