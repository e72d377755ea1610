// sampler_math.cuh -- the per-pixel arithmetic of the two bilinear samplers on the DVSG
// warp path, written once for all device kernels (it also compiles as plain host C++).
//
// Everything that feeds an integer sample index is spelled with explicitly rounded
// single operations (no FMA contraction) in exactly the reference's op order, so that the
// int32 corners are bit-exact given identical fp32 coordinates:
//   A4  TPS sampler        ThinPlateSpline.py:48-89     W/2 scaling, clamp-then-weight
//   ZP  zero-padded sampler spatial_transformer.py:515-562 == warp_with_optical_flow.py:128-173
#pragma once

#if defined(__CUDA_ARCH__)
#define DVSG_HD __host__ __device__ __forceinline__
#define DVSG_ADD(a, b) __fadd_rn((a), (b))
#define DVSG_SUB(a, b) __fsub_rn((a), (b))
#define DVSG_MUL(a, b) __fmul_rn((a), (b))
#define DVSG_DIV(a, b) __fdiv_rn((a), (b))
#else
#include <cmath>
#if defined(__CUDACC__)
#define DVSG_HD __host__ __device__ inline
#else
#define DVSG_HD inline
#endif
// host build: compiled with -ffp-contract=off so that these stay single rounded ops
#define DVSG_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define DVSG_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define DVSG_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define DVSG_DIV(a, b) ((float)((float)(a) / (float)(b)))
#endif

namespace dvsg {

// float -> int32 the way the reference's CPU cast behaves (x86 cvttss2si): values outside
// int32 (and NaN) become INT_MIN instead of saturating.  Input is already floor()ed.
DVSG_HD int cast_i32(float f) {
    return (fabsf(f) < 2147483648.0f) ? (int)f : (int)0x80000000;
}

DVSG_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// tf.linspace(-1, 1, n)[i] = -1 + step*i with step = 2/(n-1) computed by the caller in
// fp32 (ThinPlateSpline.py:94,96; TF 1.x LinSpace kernel, separate mul and add).
DVSG_HD float lin_coord(int i, float step) { return DVSG_ADD(-1.0f, DVSG_MUL(step, (float)i)); }

struct Corners {
    int x0, x1, y0, y1;       // integer corners (A4: clamped to the image; ZP: in padded coords)
    float ax0, ax1, ay0, ay1; // ax1 = x1f - xp, ax0 = xp - x0f, likewise in y
};

// ---- A4: ThinPlateSpline._interpolate ------------------------------------------------
// x, y normalised.  Returns clamped corners and the weight factors taken FROM the clamped
// corners (:57-60 then :81-88), which is what makes out-of-frame samples cancel to ~0.
DVSG_HD Corners a4_corners(float x, float y, int W, int H) {
    Corners c;
    const float xp = DVSG_MUL(DVSG_MUL(DVSG_ADD(x, 1.0f), (float)W), 0.5f);   // ((x+1)*W)/2, :48
    const float yp = DVSG_MUL(DVSG_MUL(DVSG_ADD(y, 1.0f), (float)H), 0.5f);   // :49
    const int x0 = cast_i32(floorf(xp));                                      // :52
    const int y0 = cast_i32(floorf(yp));
    const int x1 = (int)((unsigned)x0 + 1u);                                  // :53 (wraps like int32 add)
    const int y1 = (int)((unsigned)y0 + 1u);
    c.x0 = clampi(x0, 0, W - 1);                                              // :57-60
    c.x1 = clampi(x1, 0, W - 1);
    c.y0 = clampi(y0, 0, H - 1);
    c.y1 = clampi(y1, 0, H - 1);
    c.ax1 = DVSG_SUB((float)c.x1, xp);                                        // :81-88
    c.ax0 = DVSG_SUB(xp, (float)c.x0);
    c.ay1 = DVSG_SUB((float)c.y1, yp);
    c.ay0 = DVSG_SUB(yp, (float)c.y0);
    return c;
}
// weights: wa=(x0,y0) ax1*ay1, wb=(x0,y1) ax1*ay0, wc=(x1,y0) ax0*ay1, wd=(x1,y1) ax0*ay0
// blend  : add_n([wa*Ia, wb*Ib, wc*Ic, wd*Id]) left to right (:89)
DVSG_HD float a4_blend(const Corners& c, float ia, float ib, float ic, float id) {
    const float wa = DVSG_MUL(c.ax1, c.ay1), wb = DVSG_MUL(c.ax1, c.ay0);
    const float wc = DVSG_MUL(c.ax0, c.ay1), wd = DVSG_MUL(c.ax0, c.ay0);
    return DVSG_ADD(DVSG_ADD(DVSG_ADD(DVSG_MUL(wa, ia), DVSG_MUL(wb, ib)), DVSG_MUL(wc, ic)), DVSG_MUL(wd, id));
}

// ---- ZP: bilinear_interp / tf_warp ---------------------------------------------------
// pixel-space coordinate for bilinear_interp: ((x+1)/2)*(W-1), divide first (:515-516)
DVSG_HD float zp_pix_from_norm(float x, int W) {
    return DVSG_MUL(DVSG_MUL(DVSG_ADD(x, 1.0f), 0.5f), DVSG_SUB((float)W, 1.0f));
}
// From the pixel-space coordinate before the clip.  Corners are indices into the image
// zero-padded by one pixel: column x in [0, W+1], real pixel = x-1, valid iff 1 <= x <= W.
DVSG_HD Corners zp_corners(float xpix, float ypix, int W, int H) {
    Corners c;
    const float wf = (float)W, hf = (float)H;
    float xq = fminf(fmaxf(xpix, -1.0f), wf);                                 // clip to [-1, W], :517
    float yq = fminf(fmaxf(ypix, -1.0f), hf);
    xq = DVSG_ADD(xq, 1.0f);                                                  // :520
    yq = DVSG_ADD(yq, 1.0f);
    const float x0f = floorf(xq), y0f = floorf(yq);                           // :524-527
    const float x1f = DVSG_ADD(x0f, 1.0f), y1f = DVSG_ADD(y0f, 1.0f);
    c.x0 = (int)x0f;                                                          // :529-532
    c.y0 = (int)y0f;
    c.x1 = (int)fminf(x1f, DVSG_ADD(wf, 1.0f));
    c.y1 = (int)fminf(y1f, DVSG_ADD(hf, 1.0f));
    c.ax1 = DVSG_SUB(x1f, xq);                                                // :557-560, unclamped
    c.ax0 = DVSG_SUB(xq, x0f);
    c.ay1 = DVSG_SUB(y1f, yq);
    c.ay0 = DVSG_SUB(yq, y0f);
    return c;
}
// weights: w00=(y0,x0) ax1*ay1, w01=(y0,x1) ax0*ay1, w10=(y1,x0) ax1*ay0, w11=(y1,x1) ax0*ay0
DVSG_HD float zp_blend(const Corners& c, float i00, float i01, float i10, float i11) {
    const float w00 = DVSG_MUL(c.ax1, c.ay1), w01 = DVSG_MUL(c.ax0, c.ay1);
    const float w10 = DVSG_MUL(c.ax1, c.ay0), w11 = DVSG_MUL(c.ax0, c.ay0);
    return DVSG_ADD(DVSG_ADD(DVSG_ADD(DVSG_MUL(w00, i00), DVSG_MUL(w01, i01)), DVSG_MUL(w10, i10)), DVSG_MUL(w11, i11));
}
DVSG_HD bool zp_valid(int v, int n) { return v >= 1 && v <= n; }

// ---- TPS radial basis ----------------------------------------------------------------
// d2 = (x_t-px)^2 + (y_t-py)^2, r = d2*log(d2 + 1e-6)  (ThinPlateSpline.py:104-105,152-153)
DVSG_HD float tps_d2(float xt, float yt, float px, float py) {
    const float dx = DVSG_SUB(xt, px), dy = DVSG_SUB(yt, py);
    return DVSG_ADD(DVSG_MUL(dx, dx), DVSG_MUL(dy, dy));
}

}  // namespace dvsg
