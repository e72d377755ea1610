// warp_bwd_tile.cu -- the sm_100a fast path of the backward warp (C == 3, 16-B aligned rows).
//
// Same skeleton as warp_fwd_tile.cu: every WARP is an autonomous pipeline over 32x8-pixel output
// tiles (lane = column, 8 rows per thread) and recomputes the sampling coordinates (the TPS basis
// was never stored).  What TF autodiff derives for the reference graphs:
//
//   grad x,y   d out / d x_pix from the four corners (read from the TMA-staged source footprint),
//              chained to the caller's coordinates (A4: *W/2; ZP: *(W-1)/2 or 1, and the float
//              clip_by_value mask -1 <= x_pix <= W)
//   grad_im    scatter-add of w_k * grad_out over the four corners.  Float atomics on shared memory
//              are CAS loops on this architecture, but INTEGER ones are native (ATOMS.ADD): the warp
//              accumulates into a PRIVATE shared buffer shaped like its staged footprint in 32-bit
//              fixed point, scaled per tile by a power of two taken from max |grad_out| (22 significant
//              bits; weights of an interior tile lie in [0, 1]) -- exact, order-independent sums with no
//              ordering between lanes, rows or corner classes.  The buffer is converted to fp32 in
//              place and goes to global memory with ONE TMA tensor reduce-add
//              (cp.reduce.async.bulk.tensor .add.f32); box parts outside the frame are dropped by the
//              hardware.
//   grad_T     sum over pixels of grad(x_s, y_s) * basis: a second pass over the basis multiplies
//              it with the per-pixel gradients in packed fp32x2, 8 control points x (x, y) = 16
//              partial sums per lane are reduced across the warp with a 16-shuffle transpose
//              reduction, accumulated per warp in shared memory, and issued as red.global once per
//              warp at the end of its strip.
// TPS tiles at the frame border take the same path: a sample outside the frame has both clamped corners of an axis on
// one pixel with exactly opposite weights, so its contributions cancel exactly and are not scattered at all.  The padded
// samplers' border tiles read their corners from the staged footprint too and scatter with red.global.add.f32 (their
// reduce-add box would start outside the tensor); tiles whose footprint fits no box, and tiles with non-finite
// gradients, take a per-pixel path with global gathers and red.global.  grad_im is reproducible to the rounding of the
// fp32 adds between tiles only.
#include "tile_common.cuh"

namespace dvsg {

bool tps_nodes_ok(int H, int W, int C, int oh, int ow, int pn, int flags);      // warp_fwd_tile.cu

// Staging boxes of the backward kernel (floats x rows) and the CTAs per SM it is compiled for.  Every warp owns TWO buffers
// of the largest box (source footprint + fixed-point accumulator), so the box size sets the occupancy:
//   default          128 x 10, 128 x 13, 160 x 10 -> 6656 B, 4 CTAs (16 warps) per SM
//   DVSG_BWD_SMALL   128 x 10, 128 x 10, 160 x 8  -> 5120 B, 5 CTAs (20 warps) per SM, registers capped at 102
#ifdef DVSG_BWD_SMALL
constexpr int BB_W0 = 128, BB_W2 = 160, BB_H0 = 10, BB_H1 = 10, BB_H2 = 8, BSTAGE = 5120, BWD_MINB = 5;
#else
constexpr int BB_W0 = BOX_W0, BB_W2 = BOX_W2, BB_H0 = BOX_H0, BB_H1 = BOX_H1, BB_H2 = BOX_H2, BSTAGE = TSTAGE_BYTES, BWD_MINB = 4;
#endif

struct BwdTileParams {
    const float* src;        // [B,H,W,3]
    const float* grad_out;   // [B,oh,ow,3]
    float* grad_src;         // [B,H,W,3] accumulated (may be null)
    float* grad_x;           // flat [B*oh*ow] w.r.t. the caller's coordinates (may be null)
    float* grad_y;
    float* grad_flow;        // [B,H,W,2] (FLOW, may be null)
    int B, H, W, oh, ow;
    const float* coord;
    long long coord_stride;
    const float* T;
    const float* grad_x_in;  // optional upstream gradient on the returned x, y (TPS)
    const float* grad_y_in;
    float* grad_T;           // [B,2,pn+3], pre-zeroed by the launcher (may be null)
    int pn;
    float step_x, step_y;
    const float* x_in;
    const float* y_in;
    const float* flow;
    int n_tx, n_ty, segs, seg_len;
    int bw[3], bh[3];
    int nodes;               // TPS: coordinates from the tile-node evaluation (must match the forward call)
};

struct alignas(64) BwdTileMaps {
    CUtensorMap src[NBOX];    // source frames, one map per box shape
    CUtensorMap gsrc[NBOX];   // grad_src frames, same shapes (targets of the reduce-add)
};

__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, int x, int y, int z, uint32_t src_smem) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(z), "r"(src_smem) : "memory");
}

// 16 partial sums per lane -> lane l ends with the warp total of v[(l >> 1) & 15] (16 shuffles)
__device__ __forceinline__ float reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int h = 8, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- per-pixel general path: global gathers, red.global scatter (full reference semantics) ---------
template <int MODE>
__device__ __noinline__ void bwd_general_pixel(float xp, float yp, int W, int H, const float* __restrict__ srcb, float* __restrict__ gsrcb,
                                               float g0, float g1, float g2, float& dxp, float& dyp) {
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
    const int a00 = (y0 * W + x0) * 3, a01 = (y0 * W + x1) * 3, a10 = (y1 * W + x0) * 3, a11 = (y1 * W + x1) * 3;
    const float g[3] = {g0, g1, g2};
    dxp = 0.0f; dyp = 0.0f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float i00 = v00 ? __ldg(srcb + a00 + ch) : 0.0f, i01 = v01 ? __ldg(srcb + a01 + ch) : 0.0f;
        const float i10 = v10 ? __ldg(srcb + a10 + ch) : 0.0f, i11 = v11 ? __ldg(srcb + a11 + ch) : 0.0f;
        dxp += g[ch] * (ay1 * (i01 - i00) + ay0 * (i11 - i10));
        dyp += g[ch] * (ax1 * (i10 - i00) + ax0 * (i11 - i01));
        if (gsrcb) {
            if (v00) atomicAdd(gsrcb + a00 + ch, w00 * g[ch]);
            if (v10) atomicAdd(gsrcb + a10 + ch, w10 * g[ch]);
            if (v01) atomicAdd(gsrcb + a01 + ch, w01 * g[ch]);
            if (v11) atomicAdd(gsrcb + a11 + ch, w11 * g[ch]);
        }
    }
}

// ---- one pixel pair (rows 2j, 2j+1 of the lane's column) of a staged tile --------------------------------------------
// Returns (d out/d x_pix of the two pixels, d out/d y_pix of the two pixels) and scatters w_k * grad_out.
// The scatter accumulates in the warp's private buffer (= staging buffer + BSTAGE, same layout) as 32-bit FIXED
// POINT with the native integer shared-memory atomic (ATOMS.ADD; float atomics on shared memory are CAS loops on sm_100
// and lose to everything else that was tried): every scattered weight lies in [0, 1] (see CLAMP below for the frame
// border), so |contribution| <= max |grad_out| of the tile.  Contributions are scaled by a power of two chosen per
// tile from that maximum so that they carry 22 significant bits (rounded to nearest by the 1.5 * 2^23 trick: absolute
// error <= 2^-23 of the tile's largest gradient, fp32's own resolution there); integer sums are exact, independent of
// the order, and cannot overflow below 512 contributions per source pixel.  No ordering between lanes, rows or corner
// classes is needed -- the previous scheme (plain read-modify-write rounds, one per row and corner class, duplicates
// merged by shuffles) cost ~170 instructions per pair, 32 warp barriers per tile and a monotonicity test.
// CLAMP (TPS tiles at the frame border): corners clamped first, weights FROM the clamped corners (ThinPlateSpline.py:57-60,
// 81-88).  A sample outside the frame in x has x0 == x1 after the clamp, so its weights are exact negatives of each other
// (x1f - x == -(x - x0f) in floating point) and its two contributions to that one pixel cancel EXACTLY -- the reference's
// scatter leaves their rounding noise instead, in whatever order its segment sum runs.  Such samples are therefore not
// scattered at all; the others have unclamped corners and weights in [0, 1] like those of an interior tile.
template <bool CLAMP>
__device__ __forceinline__ float4 bwd_pair_fixed(const float2 xp, const float2 yp, const float ga0, const float ga1, const float ga2, const float gb0,
                                                 const float gb1, const float gb2, const unsigned char* __restrict__ sb, const int pitch,
                                                 const bool scatter, const float scale, const float wm1, const float hm1) {
    const float2 one2 = f2dup(1.0f), m23 = f2dup(MAGIC23), pitchf = f2dup((float)pitch), twelve = f2dup(12.0f);
    float2 x0f, y0f, x1f, y1f;
    bool inside[2] = {true, true};
    if (!CLAMP) {
        x0f = floor2_pos(xp); y0f = floor2_pos(yp);
        x1f = __fadd2_rn(x0f, one2); y1f = __fadd2_rn(y0f, one2);
    } else {
        const float2 fx = floor2_any(xp), fy = floor2_any(yp);
        const float2 hx = __fadd2_rn(fx, one2), hy = __fadd2_rn(fy, one2);
        x0f = f2(fminf(fmaxf(fx.x, 0.0f), wm1), fminf(fmaxf(fx.y, 0.0f), wm1));
        x1f = f2(fminf(fmaxf(hx.x, 0.0f), wm1), fminf(fmaxf(hx.y, 0.0f), wm1));
        y0f = f2(fminf(fmaxf(fy.x, 0.0f), hm1), fminf(fmaxf(fy.y, 0.0f), hm1));
        y1f = f2(fminf(fmaxf(hy.x, 0.0f), hm1), fminf(fmaxf(hy.y, 0.0f), hm1));
        inside[0] = x0f.x == fx.x && x1f.x == hx.x && y0f.x == fy.x && y1f.x == hy.x;
        inside[1] = x0f.y == fx.y && x1f.y == hx.y && y0f.y == fy.y && y1f.y == hy.y;
    }
    const float2 ax1 = sub2(x1f, xp), ax0 = sub2(xp, x0f), ay1 = sub2(y1f, yp), ay0 = sub2(yp, y0f);
    const float2 w00 = __fmul2_rn(ax1, ay1), w01 = __fmul2_rn(ax0, ay1), w10 = __fmul2_rn(ax1, ay0), w11 = __fmul2_rn(ax0, ay0);
    const float2 tx0 = __ffma2_rn(x0f, twelve, m23);
    const float2 o00 = __ffma2_rn(y0f, pitchf, tx0);
    const float* p0[2] = {reinterpret_cast<const float*>(sb + (__float_as_int(o00.x) & 0x7fffff)), reinterpret_cast<const float*>(sb + (__float_as_int(o00.y) & 0x7fffff))};
    const float* p1[2] = {reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[0]) + pitch),
                          reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[1]) + pitch)};
    // corner (x1, y0) / (x1, y1): 12 bytes after (x0, .) unless the clamp folded them onto it; likewise the row
    int dxo[2] = {12, 12}, dyo[2] = {pitch, pitch};
    if (CLAMP) {
        dxo[0] = x1f.x == x0f.x ? 0 : 12; dxo[1] = x1f.y == x0f.y ? 0 : 12;
        dyo[0] = y1f.x == y0f.x ? 0 : pitch; dyo[1] = y1f.y == y0f.y ? 0 : pitch;
        p1[0] = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[0]) + dyo[0]);
        p1[1] = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[1]) + dyo[1]);
    }
    const float* p0x[2] = {reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[0]) + dxo[0]),
                           reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[1]) + dxo[1])};
    const float* p1x[2] = {reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p1[0]) + dxo[0]),
                           reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p1[1]) + dxo[1])};
    const float gq[2][3] = {{ga0, ga1, ga2}, {gb0, gb1, gb2}};
    const float2 sc2 = f2dup(scale), m15 = f2dup(MAGIC15);
    float2 dx = f2dup(0.0f), dy = f2dup(0.0f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float2 g = f2(gq[0][ch], gq[1][ch]);
        const float2 i00 = f2(p0[0][ch], p0[1][ch]), i01 = f2(p0x[0][ch], p0x[1][ch]);
        const float2 i10 = f2(p1[0][ch], p1[1][ch]), i11 = f2(p1x[0][ch], p1x[1][ch]);
        const float2 ux = __ffma2_rn(ay0, sub2(i11, i10), __fmul2_rn(ay1, sub2(i01, i00)));
        const float2 uy = __ffma2_rn(ax0, sub2(i11, i01), __fmul2_rn(ax1, sub2(i10, i00)));
        dx = __ffma2_rn(g, ux, dx);
        dy = __ffma2_rn(g, uy, dy);
        if (scatter) {
            const float2 gs = __fmul2_rn(g, sc2);      // exact: scale is a power of two
            const float2 q00 = __ffma2_rn(w00, gs, m15), q01 = __ffma2_rn(w01, gs, m15), q10 = __ffma2_rn(w10, gs, m15), q11 = __ffma2_rn(w11, gs, m15);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (CLAMP && !inside[h]) continue;      // clamped sample: its paired contributions cancel exactly (see above)
                int* a0 = const_cast<int*>(reinterpret_cast<const int*>(p0[h])) + BSTAGE / 4 + ch;
                int* a1 = const_cast<int*>(reinterpret_cast<const int*>(p1[h])) + BSTAGE / 4 + ch;
                atomicAdd(a0, __float_as_int(h ? q00.y : q00.x) - 0x4B400000);
                atomicAdd(a1, __float_as_int(h ? q10.y : q10.x) - 0x4B400000);
                atomicAdd(a0 + 3, __float_as_int(h ? q01.y : q01.x) - 0x4B400000);
                atomicAdd(a1 + 3, __float_as_int(h ? q11.y : q11.x) - 0x4B400000);
            }
        }
    }
    return make_float4(dx.x, dx.y, dy.x, dy.y);
}

// Zero-padded samplers at the frame border (bilinear_interp / tf_warp): corners a, a+12, a+pitch, a+pitch+12 of the
// staged footprint, whose out-of-frame parts the TMA copy filled with zeros -- the reference's zero padding -- and
// red.global.add.f32 for the corners that lie inside the frame (the padding receives no gradient).  xp, yp are
// clipped+1 coordinates in the padded frame; sb has the -1 pixel shift folded in.
__device__ __forceinline__ float4 bwd_pair_padded(const float2 xp, const float2 yp, const float ga0, const float ga1, const float ga2, const float gb0,
                                                  const float gb1, const float gb2, const unsigned char* __restrict__ sb, const int pitch, const int W,
                                                  const int H, float* __restrict__ gsrcb, const bool ok_a, const bool ok_b) {
    const float2 one2 = f2dup(1.0f), m23 = f2dup(MAGIC23), pitchf = f2dup((float)pitch), twelve = f2dup(12.0f);
    const float2 x0f = floor2_pos(xp), y0f = floor2_pos(yp);
    const float2 ax1 = sub2(__fadd2_rn(x0f, one2), xp), ax0 = sub2(xp, x0f);
    const float2 ay1 = sub2(__fadd2_rn(y0f, one2), yp), ay0 = sub2(yp, y0f);
    const float2 w00 = __fmul2_rn(ax1, ay1), w01 = __fmul2_rn(ax0, ay1), w10 = __fmul2_rn(ax1, ay0), w11 = __fmul2_rn(ax0, ay0);
    const float2 o00 = __ffma2_rn(y0f, pitchf, __ffma2_rn(x0f, twelve, m23));
    const float* p0[2] = {reinterpret_cast<const float*>(sb + (__float_as_int(o00.x) & 0x7fffff)), reinterpret_cast<const float*>(sb + (__float_as_int(o00.y) & 0x7fffff))};
    const float* p1[2] = {reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[0]) + pitch),
                          reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(p0[1]) + pitch)};
    // frame coordinates of the corners: padded index - 1
    const int xa[2] = {(__float_as_int(x0f.x + MAGIC23) & 0x7fffff) - 1, (__float_as_int(x0f.y + MAGIC23) & 0x7fffff) - 1};
    const int ya[2] = {(__float_as_int(y0f.x + MAGIC23) & 0x7fffff) - 1, (__float_as_int(y0f.y + MAGIC23) & 0x7fffff) - 1};
    const bool ok[2] = {ok_a, ok_b};
    const float gq[2][3] = {{ga0, ga1, ga2}, {gb0, gb1, gb2}};
    float2 dx = f2dup(0.0f), dy = f2dup(0.0f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float2 g = f2(gq[0][ch], gq[1][ch]);
        const float2 i00 = f2(p0[0][ch], p0[1][ch]), i01 = f2(p0[0][3 + ch], p0[1][3 + ch]);
        const float2 i10 = f2(p1[0][ch], p1[1][ch]), i11 = f2(p1[0][3 + ch], p1[1][3 + ch]);
        const float2 ux = __ffma2_rn(ay0, sub2(i11, i10), __fmul2_rn(ay1, sub2(i01, i00)));
        const float2 uy = __ffma2_rn(ax0, sub2(i11, i01), __fmul2_rn(ax1, sub2(i10, i00)));
        dx = __ffma2_rn(g, ux, dx);
        dy = __ffma2_rn(g, uy, dy);
        if (gsrcb) {
            const float2 c00 = __fmul2_rn(w00, g), c01 = __fmul2_rn(w01, g), c10 = __fmul2_rn(w10, g), c11 = __fmul2_rn(w11, g);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!ok[h]) continue;
                const int x0 = xa[h], y0 = ya[h];
                const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
                float* q = gsrcb + ((long long)y0 * W + x0) * 3 + ch;
                if (vx0 && vy0) atomicAdd(q, h ? c00.y : c00.x);
                if (vx1 && vy0) atomicAdd(q + 3, h ? c01.y : c01.x);
                if (vx0 && vy1) atomicAdd(q + (size_t)W * 3, h ? c10.y : c10.x);
                if (vx1 && vy1) atomicAdd(q + (size_t)W * 3 + 3, h ? c11.y : c11.x);
            }
        }
    }
    return make_float4(dx.x, dx.y, dy.x, dy.y);
}

// NM: 0 = every radial term per pixel; 1 = tile-node evaluation, any mesh; 4 / 5 / 16 = with the separable node pass
template <int MODE, int NM>
__global__ void __launch_bounds__(TNT, BWD_MINB) warp_bwd_tile_kernel(const BwdTileParams p, const __grid_constant__ BwdTileMaps maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_mbar[TNW];
    __shared__ float s_lin[12];
    __shared__ __align__(16) float s_yt[TR];

    constexpr bool NODES = NM > 0;
    constexpr int NG = NM > 1 ? NM : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, oh = p.oh, ow = p.ow;
    const int seg = blockIdx.x, b = blockIdx.z;
    const int row0 = blockIdx.y * TR;
    const int t_begin = seg * p.seg_len, t_end = min(t_begin + p.seg_len, p.n_tx);
    const int rows_ok = min(TR, oh - row0);
    const int pn4 = (p.pn + 3) & ~3, pn8 = (p.pn + 7) & ~7;
    const bool want_gT = MODE == TMODE_TPS && p.grad_T != nullptr;

    unsigned char* w_stage = smem + (size_t)warp * (2 * BSTAGE);
    unsigned char* w_acc = w_stage + BSTAGE;
    // after the per-warp buffers: [per-pixel tables (exact mode only)][per-warp grad_T accumulators]
    //                             [node mode: node tables][per-warp node exchange buffers][Lagrange weights NODE_LX]
    const size_t rec_bytes = NODES ? 0 : (size_t)pn8 * sizeof(TpsRec);
    const unsigned char* recs = smem + (size_t)TNW * (2 * BSTAGE);
    float* w_gt = reinterpret_cast<float*>(smem + (size_t)TNW * (2 * BSTAGE) + rec_bytes) + warp * (2 * pn8 + 16);
    const uint32_t stage_s = smem_u32(w_stage), acc_s = smem_u32(w_acc), mbar = smem_u32(&s_mbar[warp]);
    unsigned char* const node_base = smem + (size_t)TNW * (2 * BSTAGE) + rec_bytes + (size_t)TNW * (2 * pn8 + 16) * sizeof(float);
    const NodeTables nt = node_tables_at(node_base, p.pn);
    float2* const w_nodes = reinterpret_cast<float2*>(node_base + node_tables_bytes(p.pn)) + warp * 32;

    // ---- prologue -----------------------------------------------------------------------------------
    if (lane == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    if (tid < TR) s_yt[tid] = lin_coord(min(row0 + tid, oh - 1), p.step_y);
    if (MODE == TMODE_TPS) {
        if (NODES) {
            tile_node_tables<NG>(p.T + (size_t)b * 2 * (p.pn + 3), p.coord + (size_t)b * p.coord_stride, p.pn, row0, oh, p.step_x, p.step_y, tid, TNT,
                                 s_lin, nt, t_begin * TC, p.seg_len == SNODE_PER_CTA * SNODE_TILES);
        } else {
            tile_tps_tables(p.T + (size_t)b * 2 * (p.pn + 3), p.coord + (size_t)b * p.coord_stride, p.pn, pn8, row0, oh, p.step_y, tid, TNT,
                            s_lin, reinterpret_cast<TpsRec*>(smem + (size_t)TNW * (2 * BSTAGE)));
        }
        if (want_gT) for (int i = lane; i < 2 * pn8 + 16; i += 32) w_gt[i] = 0.0f;
    }
    __syncthreads();
    const int node_l = min(lane, NNX * NNY - 1);
    const float node_xoff = NODES ? NODE_XOFF[node_l % NNX] : 0.0f;
    const float node_yn = NODES ? fmaf(p.step_y, (float)row0 + NODE_YOFF[node_l / NNX], -1.0f) : 0.0f;
    float node_lx[NNX];
    if (NODES) node_load_lx(lane, node_lx);

    const float* srcb = p.src + (size_t)b * H * W * 3;
    float* gsrcb = p.grad_src ? p.grad_src + (size_t)b * H * W * 3 : nullptr;
    const float2 one2 = f2dup(1.0f);
    unsigned phase = 0;
    bool red_pending = false;

    for (int t = t_begin + warp; t < t_end; t += TNW) {
        const int col0 = t * TC;
        const int col = min(col0 + lane, ow - 1);
        const bool col_ok = col0 + lane < ow;
        const float xt = lin_coord(col, p.step_x);

        // grad_out of the thread's 8 pixels: issued first, the global-memory latency hides behind the basis
        float gq[TR][3];
        const size_t opix0 = ((size_t)b * oh + row0) * ow + col;      // flat index of the lane's pixel in row 0 of the tile
        {
            const float* gp = p.grad_out + opix0 * 3;
            const int rstride = ow * 3;
#pragma unroll
            for (int q = 0; q < TR; ++q) {
                const bool ok = col_ok && q < rows_ok;
                const float* gr = gp + min(q, rows_ok - 1) * rstride;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) gq[q][ch] = ok ? __ldg(gr + ch) : 0.0f;
            }
        }

        // ================= A: coordinates (identical arithmetic to the forward kernel) =================
        float2 XP[TR / 2], YP[TR / 2];
        unsigned clipmask = 0;      // padded modes: bit q = x inside the clip range, bit 8+q = y inside
        if (MODE == TMODE_TPS) {
            if (NODES) {
                if (col0 + TC > ow) node_load_lx(min(lane, ow - 1 - col0), node_lx);
                tile_node_coords<NG>(nt, p.pn, s_lin, s_yt, col0, node_lx, xt, node_xoff, node_yn, p.step_x, rows_ok, w_nodes, lane, XP, YP);
            } else {
                const float bx = fmaf(s_lin[1], xt, s_lin[0]), by = fmaf(s_lin[4], xt, s_lin[3]);
                const float2 l2 = f2dup(s_lin[2]), l5 = f2dup(s_lin[5]);
#pragma unroll
                for (int j = 0; j < TR / 2; ++j) {
                    const float2 ytp = *reinterpret_cast<const float2*>(s_yt + 2 * j);
                    XP[j] = __ffma2_rn(l2, ytp, f2dup(bx));
                    YP[j] = __ffma2_rn(l5, ytp, f2dup(by));
                }
                tile_tps_basis(recs, pn4, xt, XP, YP);
            }
            const float2 wf = f2dup((float)W), hf = f2dup((float)H), half2 = f2dup(0.5f);
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                XP[j] = __fmul2_rn(__fmul2_rn(__fadd2_rn(XP[j], one2), wf), half2);
                YP[j] = __fmul2_rn(__fmul2_rn(__fadd2_rn(YP[j], one2), hf), half2);
            }
        } else {
            float xs[TR], ys[TR];
#pragma unroll
            for (int q = 0; q < TR; ++q) {
                const int row = min(row0 + q, oh - 1);
                const size_t i = ((size_t)b * oh + row) * ow + col;
                if (MODE == TMODE_GIVEN) { xs[q] = __ldg(p.x_in + i); ys[q] = __ldg(p.y_in + i); }
                else { const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + i); xs[q] = f.x; ys[q] = f.y; }
            }
            const float wf = (float)W, hf = (float)H;
#pragma unroll
            for (int q = 0; q < TR; ++q) {
                if (MODE == TMODE_GIVEN) { xs[q] = zp_pix_from_norm(xs[q], W); ys[q] = zp_pix_from_norm(ys[q], H); }
                else { xs[q] = DVSG_ADD((float)col, xs[q]); ys[q] = DVSG_ADD((float)min(row0 + q, oh - 1), ys[q]); }
                // clip_by_value passes gradient on -1 <= x_pix <= W (inclusive)
                if (xs[q] >= -1.0f && xs[q] <= wf) clipmask |= 1u << q;
                if (ys[q] >= -1.0f && ys[q] <= hf) clipmask |= 256u << q;
            }
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                XP[j] = f2(DVSG_ADD(fminf(fmaxf(xs[2 * j], -1.0f), wf), 1.0f), DVSG_ADD(fminf(fmaxf(xs[2 * j + 1], -1.0f), wf), 1.0f));
                YP[j] = f2(DVSG_ADD(fminf(fmaxf(ys[2 * j], -1.0f), hf), 1.0f), DVSG_ADD(fminf(fmaxf(ys[2 * j + 1], -1.0f), hf), 1.0f));
            }
        }

        // ================= F: footprint =================
        const float xmn = min8n(XP), xmx = max8n(XP), ymn = min8n(YP), ymx = max8n(YP);
        const bool sane = (fabsf(xmn) + fabsf(xmx)) + (fabsf(ymn) + fabsf(ymx)) < 4.0e6f;
        int x_lo = sane ? floor_small(xmn) : -(1 << 30), x_hi = sane ? floor_small(xmx) + 1 : (1 << 30);
        int y_lo = sane ? floor_small(ymn) : -(1 << 30), y_hi = sane ? floor_small(ymx) + 1 : (1 << 30);
        x_lo = __reduce_min_sync(0xffffffffu, x_lo); x_hi = __reduce_max_sync(0xffffffffu, x_hi);
        y_lo = __reduce_min_sync(0xffffffffu, y_lo); y_hi = __reduce_max_sync(0xffffffffu, y_hi);
        const bool all_sane = x_lo > -(1 << 30);
        bool interior;
        int fx_lo, fx_hi, fy_lo, fy_hi;
        if (MODE == TMODE_TPS) {
            interior = x_lo >= 0 && x_hi <= W - 1 && y_lo >= 0 && y_hi <= H - 1;
            // border tiles: the A4 sampler clamps its corners into the frame (ThinPlateSpline.py:57-60), so does the footprint
            fx_lo = min(max(x_lo, 0), W - 1); fx_hi = min(max(x_hi, 0), W - 1);
            fy_lo = min(max(y_lo, 0), H - 1); fy_hi = min(max(y_hi, 0), H - 1);
        } else {
            fx_lo = x_lo - 1; fx_hi = x_hi - 1; fy_lo = y_lo - 1; fy_hi = y_hi - 1;
            // frame-border tiles (footprint reaching into the zero padding) cannot use the accumulation buffer -- the TMA
            // reduce-add faults on boxes that start outside the tensor -- they scatter with red.global (bwd_pair_padded)
            interior = fx_lo >= 0 && fy_lo >= 0 && fx_hi <= W - 1 && fy_hi <= H - 1;
        }
        const int fx0 = (fx_lo * 3) & ~3;
        const int fw = (fx_hi + 1) * 3 - fx0, nrows = fy_hi - fy_lo + 1;
        int box = -1;
        if (fw <= p.bw[0]) box = nrows <= p.bh[0] ? 0 : (nrows <= p.bh[1] ? 1 : -1);
        else if (fw <= p.bw[2] && nrows <= p.bh[2]) box = 2;
        const int pitch = (box == 2 ? p.bw[2] : p.bw[0]) * 4;
        const int box_rows = box == 0 ? p.bh[0] : (box == 1 ? p.bh[1] : p.bh[2]);
        bool staged = box >= 0 && all_sane && (long long)(fy_hi + 3) * pitch + (long long)(fx_hi + 3) * 12 < (1LL << 22);
        // Staged tiles accumulate grad_im in the warp's shared buffer in fixed point ("fast"; TPS border tiles with the
        // clamped-corner variant), except the padded samplers' border tiles, which scatter with red.global.
        // per-tile fixed-point scale of the scatter: 2^(21 - floor(log2 max|grad_out|)), from the exponent bits of the maximum
        unsigned gmax_bits = 0;
#pragma unroll
        for (int q = 0; q < TR; ++q)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) gmax_bits = max(gmax_bits, __float_as_uint(fabsf(gq[q][ch])));
        gmax_bits = __reduce_max_sync(0xffffffffu, gmax_bits);
        const int gexp = min(max((int)(gmax_bits >> 23), 40), 254);
        const bool g_finite = (gmax_bits >> 23) < 255, g_zero = (gmax_bits >> 23) < 40;      // |g| < 2^-87 everywhere: nothing to scatter
        // the padded samplers scatter their frame-border tiles with red.global: the reduce-add box would start outside the tensor
        const bool fast = staged && g_finite && (interior || MODE == TMODE_TPS);
        const float scale = __int_as_float((275 - gexp) << 23), inv_scale = __int_as_float((gexp - 21) << 23);
        if (!g_finite) staged = false;              // inf / NaN gradients: per-pixel path (float red.global keeps their semantics)

        // ================= L: stage the source footprint, clear the accumulation buffer =================
        if (red_pending) {                 // the previous tile's reduce-add must have read the accumulation buffer
            if (lane == 0) bulk_wait_read0();
            red_pending = false;
        }
        __syncwarp();
        if (staged) {
            if (lane == 0) {
                mbar_arrive_expect_tx(mbar, (unsigned)(pitch * box_rows));
                tma_load_3d(stage_s, &maps.src[box], fx0, fy_lo, b, mbar);
            }
            if (gsrcb && fast && !g_zero) {
                float4* z = reinterpret_cast<float4*>(w_acc);
                const int n16 = pitch * box_rows / 16;
                for (int i = lane; i < n16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();
            mbar_wait(mbar, phase); phase ^= 1u;
        }

        // ================= G: d out / d coordinates, scatter of grad_im =================
        float2 GX[TR / 2], GY[TR / 2];     // gradient w.r.t. the sampler's pixel-space coordinate, then chained
        if (fast) {
            const unsigned char* sb = w_stage - (fy_lo * pitch + fx0 * 4) - (MODE == TMODE_TPS ? 0 : pitch + 12);
            const bool sc = gsrcb != nullptr && !g_zero;
            // two rounds of two pairs (the pair bodies are ~200 instructions each: four inlined copies per variant put the
            // kernel's hot path past the instruction cache); the register arrays rotate between the rounds
            const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
            GX[2] = GX[3] = GY[2] = GY[3] = f2dup(0.0f);
#pragma unroll 1
            for (int it = 0; it < 2; ++it) {
                float4 d0, d1;
                if (MODE != TMODE_TPS || interior) {
                    d0 = bwd_pair_fixed<false>(XP[0], YP[0], gq[0][0], gq[0][1], gq[0][2], gq[1][0], gq[1][1], gq[1][2], sb, pitch, sc, scale, 0.0f, 0.0f);
                    d1 = bwd_pair_fixed<false>(XP[1], YP[1], gq[2][0], gq[2][1], gq[2][2], gq[3][0], gq[3][1], gq[3][2], sb, pitch, sc, scale, 0.0f, 0.0f);
                } else {
                    d0 = bwd_pair_fixed<true>(XP[0], YP[0], gq[0][0], gq[0][1], gq[0][2], gq[1][0], gq[1][1], gq[1][2], sb, pitch, sc, scale, wm1, hm1);
                    d1 = bwd_pair_fixed<true>(XP[1], YP[1], gq[2][0], gq[2][1], gq[2][2], gq[3][0], gq[3][1], gq[3][2], sb, pitch, sc, scale, wm1, hm1);
                }
                GX[0] = GX[2]; GX[1] = GX[3]; GY[0] = GY[2]; GY[1] = GY[3];
                GX[2] = f2(d0.x, d0.y); GY[2] = f2(d0.z, d0.w); GX[3] = f2(d1.x, d1.y); GY[3] = f2(d1.z, d1.w);
                XP[0] = XP[2]; XP[1] = XP[3]; YP[0] = YP[2]; YP[1] = YP[3];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) gq[q][ch] = gq[q + 4][ch];
            }
        } else if (MODE != TMODE_TPS && staged) {
            // frame-border tiles of the padded samplers: corners still come from the staged footprint, the scatter goes to global memory
            const unsigned char* sb = w_stage - (fy_lo * pitch + fx0 * 4) - (pitch + 12);
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                const float4 d = bwd_pair_padded(XP[j], YP[j], gq[2 * j][0], gq[2 * j][1], gq[2 * j][2], gq[2 * j + 1][0], gq[2 * j + 1][1], gq[2 * j + 1][2],
                                                 sb, pitch, W, H, gsrcb, col_ok && row0 + 2 * j < oh, col_ok && row0 + 2 * j + 1 < oh);
                GX[j] = f2(d.x, d.y); GY[j] = f2(d.z, d.w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < TR; ++q) {
                const float xp = (q & 1) ? XP[q >> 1].y : XP[q >> 1].x, yp = (q & 1) ? YP[q >> 1].y : YP[q >> 1].x;
                const int row = row0 + q;
                const bool ok = col_ok && row < oh;
                float dxp = 0.0f, dyp = 0.0f;
                if (ok) bwd_general_pixel<MODE>(xp, yp, W, H, srcb, gsrcb, gq[q][0], gq[q][1], gq[q][2], dxp, dyp);
                if (q & 1) { GX[q >> 1].y = dxp; GY[q >> 1].y = dyp; } else { GX[q >> 1].x = dxp; GY[q >> 1].x = dyp; }
            }
        }

        // ================= chain to the caller's coordinates, write grad x / y / flow =================
        // (per-tile base pointers + 32-bit row offsets: the flat 64-bit index per pixel and array cost 39 instructions per 32 px)
        {
            const int r_last = rows_ok - 1;
            if (MODE == TMODE_TPS) {
                // block-uniform bases + 32-bit per-lane offsets; rows / columns past the edge alias the edge pixel, so the
                // loads need no guard (their values are zeroed below)
                const size_t ubase = ((size_t)b * oh + row0) * ow;
                const float* gxi = p.grad_x_in ? p.grad_x_in + ubase : nullptr;
                const float* gyi = p.grad_x_in ? p.grad_y_in + ubase : nullptr;
                float* gxo = p.grad_x ? p.grad_x + ubase : nullptr;
                float* gyo = p.grad_x ? p.grad_y + ubase : nullptr;
                const float2 wf = f2dup((float)W), hf = f2dup((float)H), half2 = f2dup(0.5f);
#pragma unroll
                for (int j = 0; j < TR / 2; ++j) {
                    const unsigned oa = (unsigned)(min(2 * j, r_last) * ow + col), ob = (unsigned)(min(2 * j + 1, r_last) * ow + col);
                    const bool oka = col_ok && 2 * j < rows_ok, okb = col_ok && 2 * j + 1 < rows_ok;
                    float2 gx = __fmul2_rn(__fmul2_rn(GX[j], wf), half2);      // x_pix = (x+1)*W/2
                    float2 gy = __fmul2_rn(__fmul2_rn(GY[j], hf), half2);
                    if (gxi) {
                        gx = __fadd2_rn(gx, f2(__ldg(gxi + oa), __ldg(gxi + ob)));
                        gy = __fadd2_rn(gy, f2(__ldg(gyi + oa), __ldg(gyi + ob)));
                    }
                    if (!oka) { gx.x = 0.0f; gy.x = 0.0f; }
                    if (!okb) { gx.y = 0.0f; gy.y = 0.0f; }
                    if (gxo) {
                        if (oka) { gxo[oa] = gx.x; gyo[oa] = gy.x; }
                        if (okb) { gxo[ob] = gx.y; gyo[ob] = gy.y; }
                    }
                    GX[j] = gx; GY[j] = gy;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TR / 2; ++j) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = 2 * j + h;
                        const bool ok = col_ok && q < rows_ok;
                        const size_t opix = opix0 + (size_t)(min(q, rows_ok - 1) * ow);
                        float gx = h ? GX[j].y : GX[j].x, gy = h ? GY[j].y : GY[j].x;
                        gx = (clipmask >> q) & 1u ? gx : 0.0f;
                        gy = (clipmask >> (8 + q)) & 1u ? gy : 0.0f;
                        if (ok) {
                            if (MODE == TMODE_GIVEN) {
                                if (p.grad_x) {
                                    p.grad_x[opix] = gx * ((float)W - 1.0f) * 0.5f;   // x_pix = (x+1)/2*(W-1)
                                    p.grad_y[opix] = gy * ((float)H - 1.0f) * 0.5f;
                                }
                            } else if (p.grad_flow) {
                                reinterpret_cast<float2*>(p.grad_flow)[opix] = make_float2(gx, gy);
                            }
                        }
                    }
                }
            }
        }

        // ================= S: accumulation buffer -> grad_im with one TMA tensor reduce-add =================
        if (fast && gsrcb && !g_zero) {
            __syncwarp();
            {   // fixed point -> fp32 in place (exact scaling by a power of two), then the reduce-add sees ordinary floats
                int4* zi = reinterpret_cast<int4*>(w_acc);
                const int n16 = pitch * box_rows / 16;
                for (int i = lane; i < n16; i += 32) {
                    const int4 v = zi[i];
                    reinterpret_cast<float4*>(zi)[i] = make_float4(__int2float_rn(v.x) * inv_scale, __int2float_rn(v.y) * inv_scale,
                                                                   __int2float_rn(v.z) * inv_scale, __int2float_rn(v.w) * inv_scale);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_reduce_add_3d(&maps.gsrc[box], fx0, fy_lo, b, acc_s);
                bulk_commit();
            }
            red_pending = true;
        }

        // ================= P2: grad_T = sum over pixels of grad(x_s, y_s) * (1, x_t, y_t, r_1..r_pn) =================
        if (want_gT && NODES) {
            // ---- node mode: the adjoint of what tile_node_coords computes ------------------------------------------------
            // coordinates = affine + sum_n L_n(pixel) F(n) + near field, F(n) = sum_{far k} c_k phi_k(node n):
            //   A(n)      = sum_pixels L_n(pixel) g(pixel)              adjoint interpolation: rows first, then columns
            //   grad c_k  = sum_n A(n) phi_k(n)                         far control points of this tile (lane = control point)
            //   grad c_k  = sum_pixels g(pixel) phi_k(pixel)            its near control points, per pixel
            // Scratch = the warp's staging buffer, idle between the gather of this tile and the TMA copy of the next.
            __syncwarp();                                                          // every lane has finished reading the staged footprint
            float* const hbuf = reinterpret_cast<float*>(w_stage);                 // [32 lanes][12]: (hx_b, hy_b), b = 0..4
            float2* const abuf = reinterpret_cast<float2*>(w_stage + 32 * 48);     // [32 nodes]
            const int n_row_near = nt.near_cnt[0];
            unsigned near = 0;
            if (n_row_near > 0) {
                const float lo = (float)col0 - NODE_NEAR_X, hi = (float)(col0 + TC - 1) + NODE_NEAR_X;
                for (int i = 0; i < n_row_near; ++i) {
                    const float c = nt.near_col[i];
                    if (c > lo && c < hi) near |= 1u << i;
                }
            }
            if (n_row_near >= 0) {
                // rows: h_b = sum_r My[b][r] g(r)
                float hrow[2 * NNY];
#pragma unroll
                for (int bb = 0; bb < NNY; ++bb) {
                    float2 ax = f2dup(0.0f), ay = f2dup(0.0f);
#pragma unroll
                    for (int j = 0; j < TR / 2; ++j) {
                        const float2 m = f2(NODE_MY[bb][2 * j], NODE_MY[bb][2 * j + 1]);
                        ax = __ffma2_rn(m, GX[j], ax);
                        ay = __ffma2_rn(m, GY[j], ay);
                    }
                    hrow[2 * bb] = ax.x + ax.y; hrow[2 * bb + 1] = ay.x + ay.y;
                }
                float4* hb4 = reinterpret_cast<float4*>(hbuf + lane * 12);
                hb4[0] = make_float4(hrow[0], hrow[1], hrow[2], hrow[3]);
                hb4[1] = make_float4(hrow[4], hrow[5], hrow[6], hrow[7]);
                hb4[2] = make_float4(hrow[8], hrow[9], 0.0f, 0.0f);
                __syncwarp();
                // columns: A(a, b) = sum_c Lx[a](c) h_b(c); lane = node (a, b)
                {
                    const int na = node_l % NNX, nb = node_l / NNX;
                    const float* __restrict__ hp = hbuf + 2 * nb;
                    // Lagrange weights straight from the (L1-resident, 1 KB) table: a copy in shared memory cost the 5 x 5 mesh
                    // its fourth CTA per SM
                    const float* __restrict__ lp = &NODE_LX[0][0] + na;
                    float ax = 0.0f, ay = 0.0f;
#pragma unroll 8
                    for (int c = 0; c < TC; ++c) {
                        const float2 h = *reinterpret_cast<const float2*>(hp + c * 12);
                        const float w = lp[c * 8];
                        ax = fmaf(w, h.x, ax);
                        ay = fmaf(w, h.y, ay);
                    }
                    abuf[lane] = lane < NNX * NNY ? make_float2(ax, ay) : make_float2(0.0f, 0.0f);
                }
                __syncwarp();
                // far control points: lane = control point, loop over the nodes
                {
                    float xnode[NNX], ynode[NNY];
#pragma unroll
                    for (int a = 0; a < NNX; ++a) xnode[a] = fmaf(p.step_x, (float)col0 + NODE_XOFF[a], -1.0f);
#pragma unroll
                    for (int bb = 0; bb < NNY; ++bb) ynode[bb] = fmaf(p.step_y, (float)row0 + NODE_YOFF[bb], -1.0f);
                    for (int kk = lane; kk < p.pn; kk += 32) {
                        bool is_near = false;
                        for (unsigned m = near; m; m &= m - 1) is_near = is_near || nt.near_idx[__ffs(m) - 1] == kk;
                        const float4 cp = nt.cp[kk];
                        float dx2[NNX], dy2[NNY];
#pragma unroll
                        for (int a = 0; a < NNX; ++a) { const float d = xnode[a] - cp.x; dx2[a] = d * d; }
#pragma unroll
                        for (int bb = 0; bb < NNY; ++bb) { const float d = ynode[bb] - cp.y; dy2[bb] = fmaf(d, d, TPS_TINY); }
                        float sx = 0.0f, sy = 0.0f;
#pragma unroll
                        for (int n = 0; n < NNX * NNY; ++n) {
                            const float2 a = abuf[n];
                            const float d2 = dx2[n % NNX] + dy2[n / NNX];
                            const float r = d2 * lg2_approx(d2);
                            sx = fmaf(a.x, r, sx);
                            sy = fmaf(a.y, r, sy);
                        }
                        if (!is_near) { w_gt[2 * kk] += sx; w_gt[2 * kk + 1] += sy; }
                    }
                }
                __syncwarp();
            }
            // near control points (all of them in a degenerate strip): per pixel, the forward's own radial term
            {
                const int n_loop = n_row_near < 0 ? p.pn : n_row_near;
                for (int i = 0; i < n_loop; ++i) {
                    if (n_row_near >= 0 && !((near >> i) & 1u)) continue;
                    const int kk = n_row_near < 0 ? i : nt.near_idx[i];
                    const float4 cp = nt.cp[kk];
                    const float dx = xt - cp.x;
                    const float2 dxx = f2dup(dx * dx), npy = f2dup(-cp.y);
                    float2 sx = f2dup(0.0f), sy = f2dup(0.0f);
#pragma unroll
                    for (int j = 0; j < TR / 2; ++j) {
                        const float2 dy = __fadd2_rn(*reinterpret_cast<const float2*>(s_yt + 2 * j), npy);
                        const float2 d2 = __ffma2_rn(dy, dy, dxx);
                        const float2 t = __fadd2_rn(d2, f2dup(TPS_EPS));
                        const float2 r = __ffma2_rn(d2, f2(lg2_approx(t.x), lg2_approx(t.y)), f2dup(-TPS_EPS_LG2));
                        sx = __ffma2_rn(GX[j], r, sx);
                        sy = __ffma2_rn(GY[j], r, sy);
                    }
                    float vx = sx.x + sx.y, vy = sy.x + sy.y;
#pragma unroll
                    for (int o = 16; o; o >>= 1) { vx += __shfl_xor_sync(0xffffffffu, vx, o); vy += __shfl_xor_sync(0xffffffffu, vy, o); }
                    if (lane == 0) { w_gt[2 * kk] += vx; w_gt[2 * kk + 1] += vy; }
                }
            }
            {   // affine part: (1, x_t, y_t)
                float v[16];
                float2 sx = f2dup(0.0f), sy = f2dup(0.0f), sxy = f2dup(0.0f), syy = f2dup(0.0f);
#pragma unroll
                for (int j = 0; j < TR / 2; ++j) {
                    const float2 ytp = *reinterpret_cast<const float2*>(s_yt + 2 * j);
                    sx = __fadd2_rn(sx, GX[j]); sy = __fadd2_rn(sy, GY[j]);
                    sxy = __ffma2_rn(GX[j], ytp, sxy); syy = __ffma2_rn(GY[j], ytp, syy);
                }
                const float ax = sx.x + sx.y, ay = sy.x + sy.y;
                v[0] = ax; v[1] = ax * xt; v[2] = sxy.x + sxy.y; v[3] = ay; v[4] = ay * xt; v[5] = syy.x + syy.y;
#pragma unroll
                for (int i = 6; i < 16; ++i) v[i] = 0.0f;
                const float tot = reduce16(v, lane);
                if ((lane & 1) == 0) w_gt[2 * pn8 + (lane >> 1)] += tot;
            }
            __syncwarp();
        } else if (want_gT) {
            const unsigned char* rp = recs;
            for (int k = 0; k < pn8; k += 8) {
                float v[16];
#pragma unroll
                for (int u = 0; u < 8; ++u, rp += sizeof(TpsRec)) {
                    const float4 pc = *reinterpret_cast<const float4*>(rp);
                    const float4 da = *reinterpret_cast<const float4*>(rp + 16);
                    const float4 db = *reinterpret_cast<const float4*>(rp + 32);
                    const float dx = DVSG_ADD(xt, pc.x);
                    const float2 dxx = f2dup(DVSG_MUL(dx, dx));
                    const float2 dy[TR / 2] = {f2(da.x, da.y), f2(da.z, da.w), f2(db.x, db.y), f2(db.z, db.w)};
                    float2 sx = f2dup(0.0f), sy = f2dup(0.0f);
#pragma unroll
                    for (int j = 0; j < TR / 2; ++j) {
                        const float2 d2 = __fadd2_rn(dxx, dy[j]);
                        const float2 r = __fmul2_rn(d2, f2(lg2_approx(d2.x), lg2_approx(d2.y)));   // radial term as in tile_tps_basis
                        sx = __ffma2_rn(GX[j], r, sx);
                        sy = __ffma2_rn(GY[j], r, sy);
                    }
                    v[2 * u] = sx.x + sx.y;
                    v[2 * u + 1] = sy.x + sy.y;
                }
                const float tot = reduce16(v, lane);                  // lane l: entry (l >> 1) of this chunk
                if ((lane & 1) == 0) w_gt[2 * k + (lane >> 1)] += tot;
            }
            {   // affine part: (1, x_t, y_t)
                float v[16];
                float2 sx = f2dup(0.0f), sy = f2dup(0.0f), sxy = f2dup(0.0f), syy = f2dup(0.0f);
#pragma unroll
                for (int j = 0; j < TR / 2; ++j) {
                    const float2 ytp = *reinterpret_cast<const float2*>(s_yt + 2 * j);
                    sx = __fadd2_rn(sx, GX[j]); sy = __fadd2_rn(sy, GY[j]);
                    sxy = __ffma2_rn(GX[j], ytp, sxy); syy = __ffma2_rn(GY[j], ytp, syy);
                }
                const float ax = sx.x + sx.y, ay = sy.x + sy.y;
                v[0] = ax; v[1] = ax * xt; v[2] = sxy.x + sxy.y; v[3] = ay; v[4] = ay * xt; v[5] = syy.x + syy.y;
#pragma unroll
                for (int i = 6; i < 16; ++i) v[i] = 0.0f;
                const float tot = reduce16(v, lane);
                if ((lane & 1) == 0) w_gt[2 * pn8 + (lane >> 1)] += tot;
            }
            __syncwarp();
        }
    }
    if (red_pending && lane == 0) bulk_wait_read0();   // shared memory must outlive the reduce's reads

    // ---- grad_T of this warp's tiles -> global ---------------------------------------------------------
    if (want_gT) {
        __syncwarp();
        const int N = p.pn + 3;
        float* gTb = p.grad_T + (size_t)b * 2 * N;
        for (int i = lane; i < 2 * pn8; i += 32) {
            const int k = i >> 1, xy = i & 1;               // entry 2*k + xy
            // + 1e-6 * sum_pixels g: every c_k also enters the affine constant through the folded epsilon (tps_affine0)
            if (k < p.pn) atomicAdd(gTb + xy * N + 3 + k, fmaf(w_gt[i], TLN2, TPS_EPS * w_gt[2 * pn8 + 3 * xy]));
        }
        if (lane < 6) atomicAdd(gTb + (lane / 3) * N + lane % 3, w_gt[2 * pn8 + lane]);
    }
}

// ---- host side -----------------------------------------------------------------------------------------
static float btile_lin_step(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }

bool bwd_tile_path_ok(const void* src, const void* grad_src, int H, int W, int C, int oh, int ow, int pn_or_0) {
    return C == 3 && W % 4 == 0 && ow >= TC && oh >= TR && aligned16(src) && (grad_src == nullptr || aligned16(grad_src)) &&
           W < (1 << 20) && H < (1 << 20) && (long long)H * W < (1LL << 28) && (long long)oh * ow < (1LL << 28) && pn_or_0 <= TKC;
}

template <int MODE>
static int launch_bwd_tile(BwdTileParams p, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    p.n_tx = (p.ow + TC - 1) / TC;
    p.n_ty = (p.oh + TR - 1) / TR;
    const long long strips = (long long)p.B * p.n_ty;
    // the TPS prologue builds the per-strip tables; the field samplers have none and prefer short CTAs (tf_warp backward,
    // 32 x 288 x 512: 16 / 8 / 4 tiles per CTA 156 / 144 / 138 us; 16 x 720p: 40 / 20 / 4 tiles 220 / 199 / 189 us)
    p.seg_len = tile_pick_seg_len(strips, p.n_tx, 148 * BWD_MINB, MODE == TMODE_TPS ? 0.5 : 0.05, "DVSG_BWD_SEGLEN");
    // 16 x 16 meshes in node mode: SNODE_PER_CTA super-tiles per CTA, as in the forward kernel -- identical coordinates in both
    if (MODE == TMODE_TPS && p.nodes && p.pn == TKS * TKS && !getenv("DVSG_TPS_ONE_LEVEL")) p.seg_len = SNODE_PER_CTA * SNODE_TILES;
    p.segs = (p.n_tx + p.seg_len - 1) / p.seg_len;
    DVSG_REQUIRE(p.B <= 65535 && p.n_ty <= 65535, "bwd tile kernel: batch %d / %d strips exceed the grid limits: split the call", p.B, p.n_ty);
    BwdTileMaps maps;
    for (int i = 0; i < NBOX; ++i) {
        p.bw[i] = min(i == 2 ? BB_W2 : BB_W0, 3 * p.W);
        p.bh[i] = min(i == 0 ? BB_H0 : (i == 1 ? BB_H1 : BB_H2), p.H);
        int rc = encode_frames(&maps.src[i], p.src, p.B, p.H, p.W, p.bw[i], p.bh[i]);
        if (rc) return rc;
        rc = encode_frames(&maps.gsrc[i], p.grad_src ? p.grad_src : p.src, p.B, p.H, p.W, p.bw[i], p.bh[i]);
        if (rc) return rc;
    }
    const int pn8 = MODE == TMODE_TPS ? (p.pn + 7) & ~7 : 0;
    const bool nodes = MODE == TMODE_TPS && p.nodes;
    const size_t smem = (size_t)TNW * 2 * BSTAGE + (nodes ? 0 : (size_t)pn8 * sizeof(TpsRec)) + (size_t)TNW * (2 * pn8 + 16) * sizeof(float) +
                        (nodes ? node_tables_bytes(p.pn) + (size_t)TNW * 32 * sizeof(float2) : 0);
    auto go = [&](auto k) {
        ensure_dynamic_smem(reinterpret_cast<const void*>(k), (int)smem);
        k<<<dim3((unsigned)p.segs, (unsigned)p.n_ty, (unsigned)p.B), TNT, smem, st>>>(p, maps);
    };
    if constexpr (MODE == TMODE_TPS) {
        if (!nodes) go(warp_bwd_tile_kernel<MODE, 0>);
        else if (p.pn == 16) go(warp_bwd_tile_kernel<MODE, 4>);
        else if (p.pn == 25) go(warp_bwd_tile_kernel<MODE, 5>);
        else if (p.pn == 256) go(warp_bwd_tile_kernel<MODE, 16>);
        else go(warp_bwd_tile_kernel<MODE, 1>);
    } else {
        go(warp_bwd_tile_kernel<MODE, 0>);
    }
    count_launch();
    return check_launch("warp_bwd_tile_kernel");
}

int bwd_tile_tps(const float* U, const float* coord, long long cstride, const float* T, const float* grad_out, const float* grad_x_in,
                 const float* grad_y_in, float* grad_U, float* grad_T, float* grad_xs, float* grad_ys, int B, int H, int W, int oh, int ow,
                 int pn, int flags, cudaStream_t st) {
    BwdTileParams p = {};
    p.nodes = tps_nodes_ok(H, W, 3, oh, ow, pn, flags) ? 1 : 0;
    p.src = U; p.grad_out = grad_out; p.grad_src = grad_U; p.grad_x = grad_xs; p.grad_y = grad_ys;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = cstride; p.T = T; p.pn = pn;
    p.grad_x_in = grad_x_in; p.grad_y_in = grad_y_in; p.grad_T = grad_T;
    p.step_x = btile_lin_step(ow); p.step_y = btile_lin_step(oh);
    return launch_bwd_tile<TMODE_TPS>(p, st);
}

int bwd_tile_given(const float* im, const float* x, const float* y, const float* grad_out, float* grad_im, float* grad_x, float* grad_y,
                   int B, int H, int W, int oh, int ow, cudaStream_t st) {
    BwdTileParams p = {};
    p.src = im; p.grad_out = grad_out; p.grad_src = grad_im; p.grad_x = grad_x; p.grad_y = grad_y;
    p.B = B; p.H = H; p.W = W; p.oh = oh; p.ow = ow; p.x_in = x; p.y_in = y;
    return launch_bwd_tile<TMODE_GIVEN>(p, st);
}

int bwd_tile_flow(const float* im, const float* flow, const float* grad_out, float* grad_im, float* grad_flow, int B, int H, int W,
                  cudaStream_t st) {
    BwdTileParams p = {};
    p.src = im; p.grad_out = grad_out; p.grad_src = grad_im; p.flow = flow; p.grad_flow = grad_flow;
    p.B = B; p.H = H; p.W = W; p.oh = H; p.ow = W;
    return launch_bwd_tile<TMODE_FLOW>(p, st);
}

}  // namespace dvsg
