// warp_bwd.cu -- backward kernels of the DVSG warp path for sm_100a (K4, K5-bwd, K6-bwd).
//
// What TF autodiff derives for the reference graphs:
//   grad_im   scatter-add of w_k * grad_out over the four corners (gather gradient);
//   grad x,y  d out / d x_pix from the corner values, chained to the caller's coordinates
//             (A4: *W/2, integer clamp passes no gradient; ZP: *(W-1)/2 and the float
//             clip_by_value mask -1 <= x_pix <= W);
//   grad_T    sum over pixels of grad(x_s, y_s) * basis, with the TPS basis recomputed
//             (it was never stored).
//
// grad_im is reduced before it touches memory: neighbouring output pixels hit the same
// source pixels (lane i's right corners are lane i+1's left corners; a row's bottom
// corners are the next row's top corners), so contributions are first merged across
// lanes with warp shuffles and down the thread's rows in registers, and only the
// remainder is issued as red.global.add.f32.  Atomic ordering makes grad_im reproducible
// only to rounding (tolerance, not bit-exactness).
#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

enum { BMODE_TPS = 0, BMODE_GIVEN = 1, BMODE_FLOW = 2 };

constexpr int BTW = 64, BPR = 4, BRG = 4, BTH = BPR * BRG, BNT = BTW * BRG, BKC = 256;
constexpr float BLN2 = 0.6931471805599453f;

struct BwdParams {
    const float* src;        // [B,H,W,C]
    const float* grad_out;   // [B,oh,ow,C]
    float* grad_src;         // [B,H,W,C] accumulated (may be null)
    float* grad_x;           // flat [B*oh*ow] w.r.t. the caller's coordinates (may be null)
    float* grad_y;
    int B, H, W, C, oh, ow;
    // TPS
    const float* coord;
    long long coord_stride;
    const float* T;
    const float* grad_x_in;  // optional upstream gradient on the returned x, y
    const float* grad_y_in;
    float* grad_T;           // [B,2,pn+3], pre-zeroed by the launcher (may be null)
    int pn, kc_cap;
    float step_x, step_y;
    // GIVEN
    const float* x_in;
    const float* y_in;
    // FLOW
    const float* flow;
    float* grad_flow;        // [B,H,W,2] (may be null)
    int merge;               // shuffle / register pre-reduction of grad_src on (C == 3 only)
};

__device__ __forceinline__ void red_add(float* addr, float v) { atomicAdd(addr, v); }

template <int MODE>
__global__ void __launch_bounds__(BNT, 3) warp_bwd_kernel(const BwdParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float s_lin[6];
    __shared__ float s_gaff[6];          // grad of the affine TPS coefficients
    float4* s_pt = reinterpret_cast<float4*>(smem);
    float* s_dy2 = reinterpret_cast<float*>(s_pt + p.kc_cap);
    float* s_gt = s_dy2 + p.kc_cap * BTH;   // [2][kc_cap] partial grad of the rbf weights

    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid & (BTW - 1), rg = tid >> 6;
    const int col0 = blockIdx.x * BTW, row0 = blockIdx.y * BTH, b = blockIdx.z;
    const int col = col0 + tx, rbase = rg * BPR;
    const int H = p.H, W = p.W, C = p.C, oh = p.oh, ow = p.ow;
    const bool col_ok = col < ow;
    const int N = p.pn + 3;
    const float* Tb = MODE == BMODE_TPS ? p.T + (size_t)b * 2 * N : nullptr;
    const float* cb = MODE == BMODE_TPS ? p.coord + (size_t)b * p.coord_stride : nullptr;

    // ---- coordinates (same evaluation as the forward kernel) ------------------------------
    float xs[BPR], ys[BPR];
    const float xt = lin_coord(col, p.step_x);
    if (MODE == BMODE_TPS) {
        if (tid < 6) s_gaff[tid] = 0.0f;
        if (tid < 64) {      // warps 0, 1: affine rows of x_s, y_s (constant with the folded epsilon term, x, y)
            const int w = tid >> 5, l = tid & 31;
            const float c0 = tps_affine0(Tb + w * N, p.pn, l);
            if (l == 0) s_lin[3 * w] = c0;
            else if (l < 3) s_lin[3 * w + l] = __ldg(Tb + w * N + l);
        }
        for (int k0 = 0; k0 < p.pn; k0 += p.kc_cap) {
            const int kc = min(p.kc_cap, p.pn - k0);
            if (k0 > 0) __syncthreads();
            for (int k = tid; k < kc; k += BNT)
                s_pt[k] = make_float4(__ldg(cb + 2 * (k0 + k)), __ldg(cb + 2 * (k0 + k) + 1), __ldg(Tb + 3 + k0 + k) * BLN2,
                                      __ldg(Tb + N + 3 + k0 + k) * BLN2);
            for (int i = tid; i < kc * BTH; i += BNT) {
                const int k = i / BTH, r = i % BTH;
                s_dy2[i] = tps_dy2(lin_coord(row0 + r, p.step_y), __ldg(cb + 2 * (k0 + k) + 1));
            }
            __syncthreads();
            if (k0 == 0) {
#pragma unroll
                for (int q = 0; q < BPR; ++q) {
                    const float yt = lin_coord(row0 + rbase + q, p.step_y);
                    xs[q] = fmaf(s_lin[2], yt, fmaf(s_lin[1], xt, s_lin[0]));
                    ys[q] = fmaf(s_lin[5], yt, fmaf(s_lin[4], xt, s_lin[3]));
                }
            }
#pragma unroll 4
            for (int k = 0; k < kc; ++k) {
                const float4 pk = s_pt[k];
                const float dx = DVSG_SUB(xt, pk.x);
                const float dx2 = DVSG_MUL(dx, dx);
                const float4 d = *reinterpret_cast<const float4*>(s_dy2 + k * BTH + rbase);
                const float dv[BPR] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int q = 0; q < BPR; ++q) {
                    const float d2 = DVSG_ADD(dx2, dv[q]);
                    const float r = DVSG_MUL(d2, lg2_approx(d2));
                    xs[q] = fmaf(pk.z, r, xs[q]);
                    ys[q] = fmaf(pk.w, r, ys[q]);
                }
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < BPR; ++q) {
            const int row = row0 + rbase + q;
            xs[q] = ys[q] = 0.0f;
            if (col_ok && row < oh) {
                const size_t i = ((size_t)b * oh + row) * ow + col;
                if (MODE == BMODE_GIVEN) {
                    xs[q] = zp_pix_from_norm(__ldg(p.x_in + i), W);
                    ys[q] = zp_pix_from_norm(__ldg(p.y_in + i), H);
                } else {
                    const float2 f = __ldg(reinterpret_cast<const float2*>(p.flow) + i);
                    xs[q] = DVSG_ADD((float)col, f.x);
                    ys[q] = DVSG_ADD((float)row, f.y);
                }
            }
        }
    }

    const float* srcb = p.src + (size_t)b * H * W * C;
    float* gsrcb = p.grad_src ? p.grad_src + (size_t)b * H * W * C : nullptr;
    const bool merge = p.merge && C == 3 && gsrcb;

    // vertical carry of the (already lane-merged) bottom-left contribution
    int carry_addr = -1;
    float carry[3] = {0.f, 0.f, 0.f};
    float gxs[BPR], gys[BPR];   // gradient w.r.t. x_s, y_s (TPS) kept for the grad_T pass

#pragma unroll
    for (int q = 0; q < BPR; ++q) {
        const int row = row0 + rbase + q;
        const bool ok = col_ok && row < oh;
        gxs[q] = gys[q] = 0.0f;
        Corners c;
        if (MODE == BMODE_TPS) c = a4_corners(xs[q], ys[q], W, H);
        else c = zp_corners(xs[q], ys[q], W, H);
        // weights and legal element indices of the four corners: tl=(x0,y0) tr=(x1,y0) bl=(x0,y1) br=(x1,y1)
        const float wtl = DVSG_MUL(c.ax1, c.ay1), wtr = DVSG_MUL(c.ax0, c.ay1);
        const float wbl = DVSG_MUL(c.ax1, c.ay0), wbr = DVSG_MUL(c.ax0, c.ay0);
        bool vtl = ok, vtr = ok, vbl = ok, vbr = ok;
        int x0 = c.x0, x1 = c.x1, y0 = c.y0, y1 = c.y1;
        if (MODE != BMODE_TPS) {
            const bool vx0 = zp_valid(x0, W), vx1 = zp_valid(x1, W), vy0 = zp_valid(y0, H), vy1 = zp_valid(y1, H);
            vtl = ok && vx0 && vy0; vtr = ok && vx1 && vy0; vbl = ok && vx0 && vy1; vbr = ok && vx1 && vy1;
            x0 = min(max(x0, 1) - 1, W - 1); x1 = max(min(x1, W) - 1, 0);
            y0 = min(max(y0, 1) - 1, H - 1); y1 = max(min(y1, H) - 1, 0);
        }
        const int atl = (y0 * W + x0) * C, atr = (y0 * W + x1) * C, abl = (y1 * W + x0) * C, abr = (y1 * W + x1) * C;
        const size_t opix = ((size_t)b * oh + (ok ? row : 0)) * ow + (ok ? col : 0);
        const float* go = p.grad_out + opix * C;

        float dxp = 0.0f, dyp = 0.0f;
        if (merge) {
            float g[3], ctl[3], ctr[3], cbl[3], cbr[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                g[ch] = ok ? __ldg(go + ch) : 0.0f;
                const float itl = vtl ? __ldg(srcb + atl + ch) : 0.0f, itr = vtr ? __ldg(srcb + atr + ch) : 0.0f;
                const float ibl = vbl ? __ldg(srcb + abl + ch) : 0.0f, ibr = vbr ? __ldg(srcb + abr + ch) : 0.0f;
                dxp += g[ch] * (c.ay1 * (itr - itl) + c.ay0 * (ibr - ibl));
                dyp += g[ch] * (c.ax1 * (ibl - itl) + c.ax0 * (ibr - itr));
                ctl[ch] = wtl * g[ch]; ctr[ch] = wtr * g[ch]; cbl[ch] = wbl * g[ch]; cbr[ch] = wbr * g[ch];
            }
            // horizontal merge: my right corners are usually my right neighbour's left corners
            int ktl = vtl ? atl : -1, ktr = vtr ? atr : -2, kbl = vbl ? abl : -3, kbr = vbr ? abr : -4;
            const int ptr_ = __shfl_up_sync(0xffffffffu, ktr, 1), pbr_ = __shfl_up_sync(0xffffffffu, kbr, 1);
            const int ntl_ = __shfl_down_sync(0xffffffffu, ktl, 1), nbl_ = __shfl_down_sync(0xffffffffu, kbl, 1);
            const bool in_t = lane > 0 && ptr_ == ktl, in_b = lane > 0 && pbr_ == kbl;
            const bool out_t = lane < 31 && ntl_ == ktr, out_b = lane < 31 && nbl_ == kbr;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float ft = __shfl_up_sync(0xffffffffu, ctr[ch], 1), fb = __shfl_up_sync(0xffffffffu, cbr[ch], 1);
                if (in_t) ctl[ch] += ft;
                if (in_b) cbl[ch] += fb;
            }
            // vertical merge: the previous row's bottom-left usually is this row's top-left
            if (carry_addr >= 0) {
                if (carry_addr == ktl) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) ctl[ch] += carry[ch];
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) red_add(gsrcb + carry_addr + ch, carry[ch]);
                }
            }
            if (ktl >= 0) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) red_add(gsrcb + ktl + ch, ctl[ch]);
            }
            if (ktr >= 0 && !out_t) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) red_add(gsrcb + ktr + ch, ctr[ch]);
            }
            if (kbr >= 0 && !out_b) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) red_add(gsrcb + kbr + ch, cbr[ch]);
            }
            carry_addr = kbl >= 0 ? kbl : -1;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) carry[ch] = cbl[ch];
        } else if (ok) {
            for (int ch = 0; ch < C; ++ch) {
                const float g = __ldg(go + ch);
                const float itl = vtl ? __ldg(srcb + atl + ch) : 0.0f, itr = vtr ? __ldg(srcb + atr + ch) : 0.0f;
                const float ibl = vbl ? __ldg(srcb + abl + ch) : 0.0f, ibr = vbr ? __ldg(srcb + abr + ch) : 0.0f;
                dxp += g * (c.ay1 * (itr - itl) + c.ay0 * (ibr - ibl));
                dyp += g * (c.ax1 * (ibl - itl) + c.ax0 * (ibr - itr));
                if (gsrcb) {
                    if (vtl) red_add(gsrcb + atl + ch, wtl * g);
                    if (vbl) red_add(gsrcb + abl + ch, wbl * g);
                    if (vtr) red_add(gsrcb + atr + ch, wtr * g);
                    if (vbr) red_add(gsrcb + abr + ch, wbr * g);
                }
            }
        }

        // ---- chain d/d x_pix to the caller's coordinates ---------------------------------------
        if (ok) {
            float gx, gy;
            if (MODE == BMODE_TPS) {
                gx = dxp * (float)W * 0.5f;                      // x_pix = (x+1)*W/2
                gy = dyp * (float)H * 0.5f;
                if (p.grad_x_in) { gx += __ldg(p.grad_x_in + opix); gy += __ldg(p.grad_y_in + opix); }
                gxs[q] = gx; gys[q] = gy;
                if (p.grad_x) { p.grad_x[opix] = gx; p.grad_y[opix] = gy; }
            } else {
                // clip_by_value passes gradient on -1 <= x_pix <= W (inclusive)
                const bool mx = xs[q] >= -1.0f && xs[q] <= (float)W, my = ys[q] >= -1.0f && ys[q] <= (float)H;
                gx = mx ? dxp : 0.0f;
                gy = my ? dyp : 0.0f;
                if (MODE == BMODE_GIVEN) {
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
    if (merge && carry_addr >= 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) red_add(gsrcb + carry_addr + ch, carry[ch]);
    }

    // ---- grad_T: sum over the tile of grad(x_s, y_s) * (1, x_t, y_t, r_1..r_pn) ----------------
    if (MODE == BMODE_TPS && p.grad_T) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < BPR; ++q) {
            const float yt = lin_coord(row0 + rbase + q, p.step_y);
            a[0] += gxs[q]; a[1] += gxs[q] * xt; a[2] += gxs[q] * yt;
            a[3] += gys[q]; a[4] += gys[q] * xt; a[5] += gys[q] * yt;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) atomicAdd(&s_gaff[i], a[i]);
        }
        float* gTb = p.grad_T + (size_t)b * 2 * N;
        for (int k0 = 0; k0 < p.pn; k0 += p.kc_cap) {
            const int kc = min(p.kc_cap, p.pn - k0);
            __syncthreads();
            if (p.pn > p.kc_cap) {   // tables of this chunk (single-chunk case: still resident)
                for (int k = tid; k < kc; k += BNT)
                    s_pt[k] = make_float4(__ldg(cb + 2 * (k0 + k)), __ldg(cb + 2 * (k0 + k) + 1), 0.f, 0.f);
                for (int i = tid; i < kc * BTH; i += BNT) {
                    const int k = i / BTH, r = i % BTH;
                    s_dy2[i] = tps_dy2(lin_coord(row0 + r, p.step_y), __ldg(cb + 2 * (k0 + k) + 1));
                }
            }
            for (int i = tid; i < 2 * kc; i += BNT) s_gt[i] = 0.0f;
            __syncthreads();
            for (int k = 0; k < kc; ++k) {
                const float4 pk = s_pt[k];
                const float dx = DVSG_SUB(xt, pk.x);
                const float dx2 = DVSG_MUL(dx, dx);
                const float4 d = *reinterpret_cast<const float4*>(s_dy2 + k * BTH + rbase);
                const float dv[BPR] = {d.x, d.y, d.z, d.w};
                float ax = 0.0f, ay = 0.0f;
#pragma unroll
                for (int q = 0; q < BPR; ++q) {
                    const float d2 = DVSG_ADD(dx2, dv[q]);
                    const float r = DVSG_MUL(d2, lg2_approx(d2));
                    ax = fmaf(gxs[q], r, ax);
                    ay = fmaf(gys[q], r, ay);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    ax += __shfl_xor_sync(0xffffffffu, ax, off);
                    ay += __shfl_xor_sync(0xffffffffu, ay, off);
                }
                if (lane == 0) { atomicAdd(&s_gt[k], ax); atomicAdd(&s_gt[kc + k], ay); }
            }
            __syncthreads();
            for (int i = tid; i < 2 * kc; i += BNT) {
                const int k = i % kc, rowsel = i / kc;
                atomicAdd(gTb + rowsel * N + 3 + k0 + k, s_gt[i] * BLN2);
            }
        }
        if (tid < 6) atomicAdd(gTb + (tid < 3 ? tid : N + tid - 3), s_gaff[tid]);
    }
}

static float lin_step_b(int n) { return n > 1 ? 2.0f / (float)(n - 1) : 0.0f; }
static int g_bwd_merge = 1;
static int g_bwd_notile = 0;   // experiments: force this generic kernel

// fast path (warp_bwd_tile.cu): warp-autonomous 32x8 tiles, TMA staging + TMA reduce-add
bool bwd_tile_path_ok(const void* src, const void* grad_src, int H, int W, int C, int oh, int ow, int pn_or_0);
int bwd_tile_tps(const float* U, const float* coord, long long cstride, const float* T, const float* grad_out, const float* grad_x_in,
                 const float* grad_y_in, float* grad_U, float* grad_T, float* grad_xs, float* grad_ys, int B, int H, int W, int oh, int ow,
                 int pn, int flags, cudaStream_t st);
int bwd_tile_given(const float* im, const float* x, const float* y, const float* grad_out, float* grad_im, float* grad_x, float* grad_y,
                   int B, int H, int W, int oh, int ow, cudaStream_t st);
int bwd_tile_flow(const float* im, const float* flow, const float* grad_out, float* grad_im, float* grad_flow, int B, int H, int W,
                  cudaStream_t st);

template <int MODE>
static int launch_bwd(BwdParams p, cudaStream_t st) {
    if (p.B == 0 || p.oh == 0 || p.ow == 0) return DVSG_OK;
    DVSG_REQUIRE(p.B <= 65535, "batch %d exceeds the grid z limit 65535: split the call", p.B);
    p.kc_cap = MODE == BMODE_TPS ? (p.pn < BKC ? p.pn : BKC) : 0;
    p.merge = g_bwd_merge;
    const size_t smem = (size_t)p.kc_cap * (sizeof(float4) + BTH * sizeof(float) + 2 * sizeof(float));
    dim3 grid((p.ow + BTW - 1) / BTW, (p.oh + BTH - 1) / BTH, p.B);
    warp_bwd_kernel<MODE><<<grid, BNT, smem, st>>>(p);
    count_launch();
    return check_launch("warp_bwd_kernel");
}

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_set_bwd_tuning(int merge) {
    if (merge >= 0) { g_bwd_merge = merge & 1; g_bwd_notile = (merge >> 1) & 1; }   // bit 0: shuffle merge of the generic kernel, bit 1: force the generic kernel
    return DVSG_OK;
}

extern "C" int dvsg_tps_warp_bwd_ex(const float* U, const float* coord, long long coord_batch_stride, const float* T,
                                    const float* grad_out, const float* grad_x_in, const float* grad_y_in, float* grad_U,
                                    float* grad_T, float* grad_xs, float* grad_ys, int B, int H, int W, int C, int oh, int ow,
                                    int pn, int flags, void* stream);
extern "C" int dvsg_tps_warp_bwd(const float* U, const float* coord, long long coord_batch_stride, const float* T,
                                 const float* grad_out, const float* grad_x_in, const float* grad_y_in, float* grad_U,
                                 float* grad_T, float* grad_xs, float* grad_ys, int B, int H, int W, int C, int oh, int ow,
                                 int pn, void* stream) {
    return dvsg_tps_warp_bwd_ex(U, coord, coord_batch_stride, T, grad_out, grad_x_in, grad_y_in, grad_U, grad_T, grad_xs, grad_ys, B, H, W, C, oh,
                                ow, pn, 0, stream);
}

extern "C" int dvsg_tps_warp_bwd_ex(const float* U, const float* coord, long long coord_batch_stride, const float* T,
                                    const float* grad_out, const float* grad_x_in, const float* grad_y_in, float* grad_U,
                                    float* grad_T, float* grad_xs, float* grad_ys, int B, int H, int W, int C, int oh, int ow,
                                    int pn, int flags, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0 && pn > 0, "tps_warp_bwd: bad shape");
    DVSG_REQUIRE(B == 0 || (U && coord && T && grad_out), "tps_warp_bwd: null pointer");
    DVSG_REQUIRE((grad_x_in == nullptr) == (grad_y_in == nullptr), "tps_warp_bwd: grad_x_in and grad_y_in go together");
    DVSG_REQUIRE((grad_xs == nullptr) == (grad_ys == nullptr), "tps_warp_bwd: grad_xs and grad_ys go together");
    DVSG_REQUIRE(coord_batch_stride == 0 || coord_batch_stride >= 2LL * pn, "tps_warp_bwd: coord stride %lld < 2*pn", coord_batch_stride);
    DVSG_REQUIRE((long long)H * W < (1LL << 31) / C && (long long)oh * ow < (1LL << 31) / C, "tps_warp_bwd: frame too large for int32 indexing");
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_T && B > 0) {
        if (cudaMemsetAsync(grad_T, 0, (size_t)B * 2 * (pn + 3) * sizeof(float), st) != cudaSuccess) {
            set_error("tps_warp_bwd: cudaMemsetAsync(grad_T) failed");
            return DVSG_ERR_CUDA;
        }
    }
    if (!g_bwd_notile && bwd_tile_path_ok(U, grad_U, H, W, C, oh, ow, pn))
        return B == 0 ? DVSG_OK : bwd_tile_tps(U, coord, coord_batch_stride, T, grad_out, grad_x_in, grad_y_in, grad_U, grad_T, grad_xs, grad_ys,
                                               B, H, W, oh, ow, pn, flags, st);
    BwdParams p = {};
    p.src = U; p.grad_out = grad_out; p.grad_src = grad_U; p.grad_x = grad_xs; p.grad_y = grad_ys;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow;
    p.coord = coord; p.coord_stride = coord_batch_stride; p.T = T; p.pn = pn;
    p.grad_x_in = grad_x_in; p.grad_y_in = grad_y_in; p.grad_T = grad_T;
    p.step_x = lin_step_b(ow); p.step_y = lin_step_b(oh);
    return launch_bwd<BMODE_TPS>(p, st);
}

extern "C" int dvsg_bilinear_bwd(const float* im, const float* x, const float* y, const float* grad_out, float* grad_im,
                                 float* grad_x, float* grad_y, int B, int H, int W, int C, int oh, int ow, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oh >= 0 && ow >= 0, "bilinear_bwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && x && y && grad_out), "bilinear_bwd: null pointer");
    DVSG_REQUIRE((grad_x == nullptr) == (grad_y == nullptr), "bilinear_bwd: grad_x and grad_y go together");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "bilinear_bwd: frame too large for int32 indexing");
    if (!g_bwd_notile && bwd_tile_path_ok(im, grad_im, H, W, C, oh, ow, 0))
        return B == 0 ? DVSG_OK : bwd_tile_given(im, x, y, grad_out, grad_im, grad_x, grad_y, B, H, W, oh, ow, (cudaStream_t)stream);
    BwdParams p = {};
    p.src = im; p.grad_out = grad_out; p.grad_src = grad_im; p.grad_x = grad_x; p.grad_y = grad_y;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = oh; p.ow = ow; p.x_in = x; p.y_in = y;
    return launch_bwd<BMODE_GIVEN>(p, (cudaStream_t)stream);
}

extern "C" int dvsg_flow_warp_bwd(const float* im, const float* flow, const float* grad_out, float* grad_im,
                                  float* grad_flow, int B, int H, int W, int C, void* stream) {
    DVSG_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, "flow_warp_bwd: bad shape");
    DVSG_REQUIRE(B == 0 || (im && flow && grad_out), "flow_warp_bwd: null pointer");
    DVSG_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7u) == 0 && (reinterpret_cast<uintptr_t>(grad_flow) & 7u) == 0,
                 "flow_warp_bwd: flow / grad_flow must be 8-byte aligned");
    DVSG_REQUIRE((long long)(H + 2) * (W + 2) < (1LL << 31) / C, "flow_warp_bwd: frame too large for int32 indexing");
    if (!g_bwd_notile && bwd_tile_path_ok(im, grad_im, H, W, C, H, W, 0))
        return B == 0 ? DVSG_OK : bwd_tile_flow(im, flow, grad_out, grad_im, grad_flow, B, H, W, (cudaStream_t)stream);
    BwdParams p = {};
    p.src = im; p.grad_out = grad_out; p.grad_src = grad_im; p.flow = flow; p.grad_flow = grad_flow;
    p.B = B; p.H = H; p.W = W; p.C = C; p.oh = H; p.ow = W;
    return launch_bwd<BMODE_FLOW>(p, (cudaStream_t)stream);
}
