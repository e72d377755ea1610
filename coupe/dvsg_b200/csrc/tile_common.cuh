// tile_common.cuh -- shared by the warp-autonomous tile kernels (warp_fwd_tile.cu, warp_bwd_tile.cu):
// tile geometry, TMA box shapes and tensor maps, the TPS table record, packed-fp32x2 helpers.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

enum { TMODE_TPS = 0, TMODE_GIVEN = 1, TMODE_FLOW = 2, TMODE_HOMOG = 3 };

constexpr int TR = 8;                         // rows per tile = pixels per thread
constexpr int TC = 32;                        // columns per tile = lanes
constexpr int TNW = 4;                        // warps per CTA
constexpr int TNT = TNW * 32;
constexpr int TKC = 256;                      // control points resident in shared memory
constexpr int TOUT_BYTES = TR * TC * 12;      // output tile, 3072 B
constexpr float TLN2 = 0.6931471805599453f;
constexpr float MAGIC23 = 8388608.0f;         // 2^23
constexpr float MAGIC15 = 12582912.0f;        // 1.5 * 2^23

struct TileParams {
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
    int stage_bytes;      // per-warp staging buffer
    int dbg;
    int bw[3], bh[3];     // staging boxes (floats x rows) of the three source tensor maps, clamped to the frame
    int n_tx, n_ty;       // tiles per strip / strips per frame
    int segs, seg_len;    // CTAs per strip, tiles per CTA
    float one;            // 1.0f, opaque to ptxas: add2x() below
};

// staging boxes (floats wide x rows): pitch = 512 or 640 B keeps the row pitch a multiple of 128 B, so the
// bank of a corner depends on its column only (3*x mod 32), whatever row each lane reads
constexpr int NBOX = 3;
constexpr int BOX_W0 = 128, BOX_W2 = 160;     // box 0: 128 x 10, box 1: 128 x 13, box 2: 160 x 10
constexpr int BOX_H0 = 10, BOX_H1 = 13, BOX_H2 = 10;
constexpr int TSTAGE_BYTES = 6656;            // largest box
__host__ __device__ constexpr int box_w(int i) { return i == 2 ? BOX_W2 : BOX_W0; }
__host__ __device__ constexpr int box_h(int i) { return i == 0 ? BOX_H0 : (i == 1 ? BOX_H1 : BOX_H2); }

struct alignas(64) TileMaps {
    CUtensorMap src[NBOX];   // source [B][H][3W] floats, one map per box shape
    CUtensorMap out;         // output [B][oh][3ow] floats, box = one 96 x 8 tile
};

struct __align__(16) TpsRec {   // one control point, 48 B, read with three broadcast LDS.128
    float4 pc;    // (-px, cx ln2, cy ln2, 0)
    float4 dya;   // max((y_t(row0 + r) - py)^2, TPS_TINY), r = 0..3
    float4 dyb;   // r = 4..7
};

// ---- small helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float t_lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void t_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float min3n(float a, float b, float c) { float r; asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float max3n(float a, float b, float c) { float r; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min2n(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float max2n(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2dup(float a) { return make_float2(a, a); }
__device__ __forceinline__ float min8n(const float2 (&v)[4]) { return min2n(min3n(v[0].x, v[0].y, v[1].x), min3n(v[1].y, v[2].x, min3n(v[2].y, v[3].x, v[3].y))); }
__device__ __forceinline__ float max8n(const float2 (&v)[4]) { return max2n(max3n(v[0].x, v[0].y, v[1].x), max3n(v[1].y, v[2].x, max3n(v[2].y, v[3].x, v[3].y))); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }   // per-lane IEEE a - b
// ptxas (CUDA 12.9) contracts mul.rn.f32x2 feeding add/fma.rn.f32x2 into one FFMA2 although both carry
// .rn, which would fuse roundings the reference keeps apart: sums of products use scalar rounded adds.
__device__ __forceinline__ float2 add2s(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
// the packed alternative: a + b as fma(a, one, b) with `one` = 1.0f loaded from the kernel parameters, which ptxas
// cannot fold; fma(a, 1, b) rounds a + b once, i.e. it IS add.rn, and a product feeding it stays a separate FMUL2
__device__ __forceinline__ float2 add2x(float2 a, float2 b, float2 one) { return __ffma2_rn(a, one, b); }
// per-lane add rounded toward -infinity (FADD2.RM): x +rm 2^23 drops the fraction downwards = floor
__device__ __forceinline__ float2 add2_rm(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rm.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 floor2_pos(float2 x) { return __fadd2_rn(add2_rm(x, f2dup(MAGIC23)), f2dup(-MAGIC23)); }   // 0 <= x < 2^22
__device__ __forceinline__ float2 floor2_any(float2 x) { return __fadd2_rn(add2_rm(x, f2dup(MAGIC15)), f2dup(-MAGIC15)); }   // |x| < 2^22
__device__ __forceinline__ int floor_small(float x) { return __float_as_int(__fadd_rd(x, MAGIC15)) - 0x4B400000; }             // |x| < 2^22
__device__ __forceinline__ float t_u2f(unsigned v) { return __uint_as_float(0x4B000000u | v) - MAGIC23; }   // v < 2^23, exact
__device__ __forceinline__ int t_floor_i32(float f) {
    // floor + the reference's CPU cast semantics (out-of-range / NaN -> INT_MIN)
    const int v = __float2int_rd(f);
    return fabsf(f) < 2147483648.0f ? v : (int)0x80000000;
}

// ---- tiles per CTA (host) ------------------------------------------------------------------------------------------
// A strip of n_tx tiles is cut into CTAs of seg_len tiles (TNW warps; tile t of a CTA goes to warp t % TNW).  A CTA lasts
// its prologue (tables, barrier, launch: cta_cost, in units of one tile time) plus ceil(seg_len / TNW) tile rounds; the
// launch lasts (CTAs / slots) of those (whole waves when there are fewer than four) plus a tail that grows with the
// CTA's duration (0.4 of it fits the measurements).
// The cut with the smallest product wins.  Measured: backward kernel at the 32 x 288 x 512 training shape, 3 CTAs of
// 6 / 6 / 4 tiles per strip (two half-empty rounds each) 127 us, one CTA of 16 tiles 107 us; forward kernel there 45.4 ->
// 40.4 us; flow warp 16 x 1080p, 60 / 30 / 20 tiles per CTA: 178.7 / 172.4 / 168.2 us.
static inline int tile_pick_seg_len(long long strips, int n_tx, int slots, double cta_cost, const char* env_override) {
    if (const char* e = getenv(env_override)) return max(atoi(e), 1);      // experiments only
    const int max_segs = max(n_tx / TNW, 1);
    int best_len = n_tx;
    double best = 1e300;
    for (int s = 1; s <= max_segs; ++s) {
        const int len = (n_tx + s - 1) / s, segs = (n_tx + len - 1) / len;
        double waves = (double)(strips * segs) / slots;
        if (waves < 4.0) waves = ceil(waves);       // a launch of a few waves is quantised; longer ones even out
        const double cost = (waves + 0.4) * (cta_cost + (double)((len + TNW - 1) / TNW));
        if (cost < best * (1.0 - 1e-9)) { best = cost; best_len = len; }
    }
    return best_len;
}

// ---- TPS tables shared by the forward and backward tile kernels (identical coordinates in both) -----------------------
// Per-strip tables of a tile kernel: s_lin[0..5] = affine rows (constant, x, y) of x_s and y_s, records of the pn8
// (padded) control points for the TR rows starting at row0.  Called by the whole CTA (>= 2 warps) before its barrier.
__device__ __forceinline__ void tile_tps_tables(const float* __restrict__ Tb, const float* __restrict__ cb, int pn, int pn8,
                                                int row0, int oh, float step_y, int tid, int nthreads, float* s_lin, TpsRec* wr) {
    const int N = pn + 3, lane = tid & 31, warp = tid >> 5;
    if (warp < 2) {
        const float c0 = tps_affine0(Tb + warp * N, pn, lane);
        if (lane == 0) s_lin[3 * warp] = c0;
        else if (lane < 3) s_lin[3 * warp + lane] = __ldg(Tb + warp * N + lane);
    }
    for (int k = tid; k < pn8; k += nthreads) {
        const bool real = k < pn;
        const float px = real ? __ldg(cb + 2 * k) : 0.0f, py = real ? __ldg(cb + 2 * k + 1) : 0.0f;
        const float cx = real ? __ldg(Tb + 3 + k) * TLN2 : 0.0f, cy = real ? __ldg(Tb + N + 3 + k) * TLN2 : 0.0f;
        float d[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r)
            d[r] = real ? tps_dy2(lin_coord(min(row0 + r, oh - 1), step_y), py) : 1.0f;   // padding: d2 >= 1, weight 0 -> adds exactly 0
        TpsRec rec;
        rec.pc = make_float4(-px, cx, cy, 0.0f);
        rec.dya = make_float4(d[0], d[1], d[2], d[3]); rec.dyb = make_float4(d[4], d[5], d[6], d[7]);
        wr[k] = rec;
    }
}
// TPS basis of all (padded) control points accumulated into X, Y (rows (2j, 2j+1) of the lane's column in pair j):
// packed fp32x2, one MUFU.LG2 per (pixel, control point), ln 2 folded into the coefficients.  The scalar, separately
// rounded (x_t - px)^2 enters the packed add as a broadcast operand (a packed mul feeding it would be contracted).
__device__ __forceinline__ void tile_tps_basis(const unsigned char* __restrict__ recs, const int pn8, const float xt,
                                               float2 (&X)[TR / 2], float2 (&Y)[TR / 2]) {
    const unsigned char* rp = recs;
    const unsigned char* const rend = recs + (size_t)pn8 * sizeof(TpsRec);
    do {      // pn8 >= 4: no guard in front of the loop
#pragma unroll
        for (int u = 0; u < 4; ++u, rp += sizeof(TpsRec)) {
            const float4 pc = *reinterpret_cast<const float4*>(rp);
            const float4 da = *reinterpret_cast<const float4*>(rp + 16);
            const float4 db = *reinterpret_cast<const float4*>(rp + 32);
            const float dx = DVSG_ADD(xt, pc.x);
            const float2 dxx = f2dup(DVSG_MUL(dx, dx));
            const float2 cfx = f2dup(pc.y), cfy = f2dup(pc.z);
            const float2 dy[TR / 2] = {f2(da.x, da.y), f2(da.z, da.w), f2(db.x, db.y), f2(db.z, db.w)};
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                const float2 d2 = __fadd2_rn(dxx, dy[j]);
                const float2 r = __fmul2_rn(d2, f2(lg2_approx(d2.x), lg2_approx(d2.y)));
                X[j] = __ffma2_rn(cfx, r, X[j]);
                Y[j] = __ffma2_rn(cfy, r, Y[j]);
            }
        }
    } while (rp != rend);
}

// ---- separable meshes (control point k = (gx[k % G], gy[k / G]), the regular mesh of every reference call site) --------
// Same arithmetic in the same order as tile_tps_basis -- identical coordinates -- but the tables shrink from 48 B per
// control point to 8 B (cx ln2, cy ln2) + 32 B per mesh ROW ((y_t - gy)^2 of the 8 tile rows) + 4 B per mesh column
// (-gx): 34 instead of 96 shared-memory wavefronts per tile at G = 4 (a broadcast LDS.128 costs two wavefronts), and
// (x_t - gx)^2 is formed once per mesh column instead of once per control point.
template <int G>
struct SepLayout {
    static constexpr int DY = 0;                                   // float [G][TR]
    static constexpr int CF = G * TR * 4;                          // float2 [G*G]
    static constexpr int NPX = (CF + G * G * 8 + 15) & ~15;        // float [G]
    static constexpr int BYTES = NPX + ((G * 4 + 15) & ~15);
};
// Tables of one strip; returns true (CTA-uniform) when the frame's mesh is separable and the compact tables were built,
// false after building the generic records instead.  Contains one CTA barrier; the caller still needs its own.
template <int G>
__device__ __forceinline__ bool tile_tps_tables_sep(const float* __restrict__ Tb, const float* __restrict__ cb, int pn8, int row0, int oh,
                                                    float step_y, int tid, int nthreads, float* s_lin, unsigned char* recs) {
    constexpr int PN = G * G;
    bool ok = true;
    for (int k = tid; k < PN; k += nthreads)
        ok = ok && __ldg(cb + 2 * k) == __ldg(cb + 2 * (k % G)) && __ldg(cb + 2 * k + 1) == __ldg(cb + 2 * (k / G * G) + 1);
    if (!__syncthreads_and(ok)) {
        tile_tps_tables(Tb, cb, PN, pn8, row0, oh, step_y, tid, nthreads, s_lin, reinterpret_cast<TpsRec*>(recs));
        return false;
    }
    const int N = PN + 3, lane = tid & 31, warp = tid >> 5;
    if (warp < 2) {
        const float c0 = tps_affine0(Tb + warp * N, PN, lane);
        if (lane == 0) s_lin[3 * warp] = c0;
        else if (lane < 3) s_lin[3 * warp + lane] = __ldg(Tb + warp * N + lane);
    }
    float* dyt = reinterpret_cast<float*>(recs + SepLayout<G>::DY);
    float2* cf = reinterpret_cast<float2*>(recs + SepLayout<G>::CF);
    float* npx = reinterpret_cast<float*>(recs + SepLayout<G>::NPX);
    for (int i = tid; i < G * TR; i += nthreads)
        dyt[i] = tps_dy2(lin_coord(min(row0 + i % TR, oh - 1), step_y), __ldg(cb + 2 * (i / TR * G) + 1));
    for (int k = tid; k < PN; k += nthreads) cf[k] = make_float2(__ldg(Tb + 3 + k) * TLN2, __ldg(Tb + N + 3 + k) * TLN2);
    if (tid < G) npx[tid] = -__ldg(cb + 2 * tid);
    return true;
}
template <int G>
__device__ __forceinline__ void tile_tps_basis_sep(const unsigned char* __restrict__ recs, const float xt, float2 (&X)[TR / 2], float2 (&Y)[TR / 2]) {
    const float* __restrict__ npx = reinterpret_cast<const float*>(recs + SepLayout<G>::NPX);
    float dxx[G];
#pragma unroll
    for (int gx = 0; gx < G; ++gx) {
        const float dx = DVSG_ADD(xt, npx[gx]);
        dxx[gx] = DVSG_MUL(dx, dx);
    }
    const unsigned char* dp = recs + SepLayout<G>::DY;
    const unsigned char* cp = recs + SepLayout<G>::CF;
#pragma unroll 1
    for (int gy = 0; gy < G; ++gy, dp += TR * 4, cp += G * 8) {
        const float4 da = *reinterpret_cast<const float4*>(dp);
        const float4 db = *reinterpret_cast<const float4*>(dp + 16);
        const float2 dy[TR / 2] = {f2(da.x, da.y), f2(da.z, da.w), f2(db.x, db.y), f2(db.z, db.w)};
#pragma unroll
        for (int gx = 0; gx < G; ++gx) {
            const float2 c = *reinterpret_cast<const float2*>(cp + gx * 8);
            const float2 cfx = f2dup(c.x), cfy = f2dup(c.y), dx2 = f2dup(dxx[gx]);
#pragma unroll
            for (int j = 0; j < TR / 2; ++j) {
                const float2 d2 = __fadd2_rn(dx2, dy[j]);
                const float2 r = __fmul2_rn(d2, f2(lg2_approx(d2.x), lg2_approx(d2.y)));
                X[j] = __ffma2_rn(cfx, r, X[j]);
                Y[j] = __ffma2_rn(cfy, r, Y[j]);
            }
        }
    }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, int x, int y, int z, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, int x, int y, int z, uint32_t src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(z), "r"(src_smem) : "memory");
}


// ---- host side: tensor maps ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// [B][rows][3*cols] fp32 tensor, box = bw floats x bh rows x 1 frame, zero fill outside the tensor
inline int encode_frames(CUtensorMap* m, const float* base, int B, int rows, int cols, int bw, int bh) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("tile kernel: cuTensorMapEncodeTiled is not available from this driver"); return DVSG_ERR_CUDA; }
    const cuuint64_t dims[3] = {(cuuint64_t)cols * 3, (cuuint64_t)rows, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 12, (cuuint64_t)rows * cols * 12};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tile kernel: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DVSG_ERR_CUDA; }
    return DVSG_OK;
}


}  // namespace dvsg
