// tps_solve.cu -- K1: batched TPS coefficient solve (and its backward w.r.t. the targets).
//
// Replaces _solve_system (ThinPlateSpline.py:143-166): W = [[P, R], [0, P^T]] with
// R_ij = d2_ij * log(d2_ij + 1e-6), T = (W^-1 @ pad(target))^T.
//
// The matrix ENTRIES are formed in fp32 exactly as the reference forms them (so the system
// solved is the reference's system); the elimination runs in fp64 with partial pivoting and
// the coefficients are rounded to fp32 once at the end.  The reference's own fp32
// inverse-then-multiply carries ~cond(W)*eps of noise (2e-6 for a 4x4 mesh, 1.6e-3 for
// 16x16 -- SURVEY.md H4); the fp64 elimination removes this kernel's share of that noise
// at no measurable cost (the kernel is latency-bound).
//
//   N = pn+3 <= 32 : mesh shared by the batch (the only way the reference is ever called, model.py:68):
//                    one CTA eliminates [W | rhs of up to 32 frames] at once -- W is factorised once per
//                    CTA instead of once per frame, all 256 threads work on every elimination step;
//                    per-frame meshes: one warp per frame, lane i owns row i, pivot search with shuffles.
//   N  > 32        : one CTA per DISTINCT system (1 when the mesh is shared by the batch --
//                    SURVEY.md H6) runs Gauss-Jordan on [W | I] in a global-memory workspace,
//                    then a second kernel applies W^-1 (or its transpose) to every frame.
#include <float.h>

#include "dvsg_common.cuh"
#include "sampler_math.cuh"

namespace dvsg {

constexpr int SOLVE_WARPS = 4;
constexpr int SMALL_N = 32;
constexpr int SMALL_LD = SMALL_N + 2 + 1;   // odd row pitch (doubles): lane-per-row access is conflict-free

// W_ij in fp32, reference op order (ThinPlateSpline.py:147-158).  i, j in [0, N).
__device__ __forceinline__ float tps_w_entry(const float* __restrict__ c, int pn, int i, int j) {
    if (i < pn) {
        if (j == 0) return 1.0f;
        if (j == 1) return c[2 * i];
        if (j == 2) return c[2 * i + 1];
        const int k = j - 3;
        // reduce_sum(square(p_i - p_k)) over (1, x, y): the leading term is exactly 0
        const float d2 = tps_d2(c[2 * i], c[2 * i + 1], c[2 * k], c[2 * k + 1]);
        return DVSG_MUL(d2, logf(DVSG_ADD(d2, 1e-6f)));
    }
    if (j < 3) return 0.0f;
    const int r = i - pn, k = j - 3;
    return r == 0 ? 1.0f : c[2 * k + (r - 1)];
}

// L_ij of ElasticTransformer._initialize_tps (spatial_transformer.py:324-349), fp32, as written there:
//   rows 0, 1 : [0 0 0 | x_1..x_pn], [0 0 0 | y_1..y_pn]      (L_top)
//   row  2    : [0 0 1 | 1 .. 1]                               (L_mid: ones([1, pn+1]) starts at the constant column)
//   rows 3+i  : [x_i y_i 1 | U(|p_i - p_k|^2)],  U(r2) = r2 * log(r2) with log(0) -> 0   (U_func :300-310)
// Unknowns are ordered (a_x, a_y, a_1, w_1..w_pn) -- the rows of right_mat (:288).  c = control points [pn,2] (x, y).
__device__ __forceinline__ float elastic_l_entry(const float* __restrict__ c, int pn, int i, int j) {
    if (i < 2) return j < 3 ? 0.0f : c[2 * (j - 3) + i];
    if (i == 2) return j < 2 ? 0.0f : 1.0f;
    const int r = i - 3;
    if (j < 2) return c[2 * r + j];
    if (j == 2) return 1.0f;
    const int k = j - 3;
    const float d2 = tps_d2(c[2 * r], c[2 * r + 1], c[2 * k], c[2 * k + 1]);
    return d2 > 0.0f ? DVSG_MUL(d2, logf(d2)) : 0.0f;
}

// rhs layout: TRANSPOSED == false : forward,  A = W,   rhs = pad(target)   -> T[b][c][i]
//             TRANSPOSED == true  : backward, A = W^T, rhs = grad_T^T      -> grad_target[b][i][c], i < pn
template <bool TRANSPOSED>
__global__ void __launch_bounds__(SOLVE_WARPS * 32) tps_solve_warp_kernel(const float* __restrict__ coord, long long coord_stride,
                                                                          const float* __restrict__ rhs_in, float* __restrict__ out,
                                                                          int B, int pn) {
    __shared__ double s_a[SOLVE_WARPS][SMALL_N * SMALL_LD];
    __shared__ float s_c[SOLVE_WARPS][2 * SMALL_N];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * SOLVE_WARPS + w;
    if (b >= B) return;
    const int N = pn + 3, NC = N + 2;
    double* a = s_a[w];
    float* c = s_c[w];
    const float* cb = coord + (size_t)b * coord_stride;
    for (int i = lane; i < 2 * pn; i += 32) c[i] = cb[i];
    __syncwarp();
    // build [A | rhs]; lane = column here (coalesced smem writes), rows sequential
    for (int i = 0; i < N; ++i) {
        for (int j = lane; j < NC; j += 32) {
            double v;
            if (j < N) {
                v = TRANSPOSED ? (double)tps_w_entry(c, pn, j, i) : (double)tps_w_entry(c, pn, i, j);
            } else {
                const int cc = j - N;
                if (TRANSPOSED) v = (double)rhs_in[((size_t)b * 2 + cc) * N + i];                   // grad_T[b][cc][i]
                else v = i < pn ? (double)rhs_in[((size_t)b * pn + i) * 2 + cc] : 0.0;              // pad(target)
            }
            a[i * SMALL_LD + j] = v;
        }
    }
    __syncwarp();
    // Gauss-Jordan with partial pivoting; lane i owns row i
    for (int k = 0; k < N; ++k) {
        double mag = (lane >= k && lane < N) ? fabs(a[lane * SMALL_LD + k]) : -1.0;
        int piv = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double om = __shfl_xor_sync(0xffffffffu, mag, off);
            const int op = __shfl_xor_sync(0xffffffffu, piv, off);
            if (om > mag || (om == mag && op < piv)) { mag = om; piv = op; }
        }
        if (piv != k) {   // swap rows k and piv, lanes over columns
            for (int j = lane; j < NC; j += 32) {
                const double t = a[k * SMALL_LD + j];
                a[k * SMALL_LD + j] = a[piv * SMALL_LD + j];
                a[piv * SMALL_LD + j] = t;
            }
        }
        __syncwarp();
        const double inv = 1.0 / a[k * SMALL_LD + k];
        const double f = (lane < N && lane != k) ? a[lane * SMALL_LD + k] : 0.0;
        __syncwarp();
        if (lane < N) {
            if (lane == k) {
                for (int j = k; j < NC; ++j) a[k * SMALL_LD + j] *= inv;
            }
        }
        __syncwarp();
        if (lane < N && lane != k) {
            for (int j = k; j < NC; ++j) a[lane * SMALL_LD + j] -= f * a[k * SMALL_LD + j];
        }
        __syncwarp();
    }
    // solution X[i][cc] = a[i][N+cc]
    if (TRANSPOSED) {
        for (int i = lane; i < 2 * pn; i += 32) out[(size_t)b * pn * 2 + i] = (float)a[(i >> 1) * SMALL_LD + N + (i & 1)];
    } else {
        for (int i = lane; i < 2 * N; i += 32) out[(size_t)b * 2 * N + i] = (float)a[(i % N) * SMALL_LD + N + (i / N)];
    }
}

// ---- shared mesh, N <= 32: one CTA eliminates [A | rhs columns of SH_FRAMES frames] -----------------
// The kernel is latency-bound (a few thousand dependent instructions per warp), so the elimination update
// is spread over SH_ROWG x 128 threads: thread (g, j) owns column j for the rows i = g, g+SH_ROWG, ...;
// warp 0 does the pivot search of the next column in parallel over rows.  Three CTA barriers per step.
constexpr int SH_FRAMES = 32;                 // frames per CTA -> 64 right-hand-side columns
constexpr int SH_COLS = 128;                  // >= SMALL_N + 2*SH_FRAMES columns
constexpr int SH_ROWG = 8;                    // row groups
constexpr int SH_THREADS = SH_COLS * SH_ROWG;
constexpr int SH_LD = SMALL_N + 2 * SH_FRAMES + 1;   // odd pitch (doubles)

template <bool TRANSPOSED>
__global__ void __launch_bounds__(SH_THREADS) tps_solve_shared_kernel(const float* __restrict__ coord, const float* __restrict__ rhs_in,
                                                                      float* __restrict__ out, int B, int pn) {
    __shared__ double s_a[SMALL_N * SH_LD];
    __shared__ double s_col[SMALL_N];
    __shared__ double s_inv;
    __shared__ float s_c[2 * SMALL_N];
    __shared__ int s_piv;
    const int tid = threadIdx.x;
    const int N = pn + 3;
    const int b0 = blockIdx.x * SH_FRAMES;
    const int nf = min(SH_FRAMES, B - b0);
    const int NC = N + 2 * nf;                  // columns: A, then (frame, component) pairs
    for (int i = tid; i < 2 * pn; i += SH_THREADS) s_c[i] = coord[i];
    // right-hand sides: one coalesced pass over the contiguous block of this CTA's frames (independent loads
    // in flight together, instead of N dependent global loads per column)
    if (TRANSPOSED) {
        const float* g = rhs_in + (size_t)b0 * 2 * N;                     // grad_T[b][cc][i]
        for (int e = tid; e < nf * 2 * N; e += SH_THREADS) {
            const int f = e / (2 * N), r = e - f * 2 * N, cc = r / N, i = r - cc * N;
            s_a[i * SH_LD + N + 2 * f + cc] = (double)g[e];
        }
    } else {
        const float* g = rhs_in + (size_t)b0 * pn * 2;                    // target[b][i][cc], rows pn..N-1 are the zero pad
        for (int e = tid; e < nf * 2 * N; e += SH_THREADS) {
            const int f = e / (2 * N), r = e - f * 2 * N, i = r >> 1, cc = r & 1;
            s_a[i * SH_LD + N + 2 * f + cc] = i < pn ? (double)g[(size_t)f * pn * 2 + r] : 0.0;
        }
    }
    __syncthreads();
    for (int e = tid; e < N * N; e += SH_THREADS) {                       // the system matrix, all threads
        const int i = e / N, j = e - i * N;
        s_a[i * SH_LD + j] = TRANSPOSED ? (double)tps_w_entry(s_c, pn, j, i) : (double)tps_w_entry(s_c, pn, i, j);
    }
    // Gauss-Jordan with partial pivoting (same pivot rule and operation order as the per-frame kernel)
    for (int k = 0; k < N; ++k) {
        __syncthreads();                        // column k is final
        if (tid < 32) {
            const double ck = tid < N ? s_a[tid * SH_LD + k] : 0.0;
            // pivot = largest |.| of column k (lowest row on ties); magnitudes compared in fp32: candidates that
            // tie to fp32 precision are equally good pivots
            float mag = (tid >= k && tid < N) ? fabsf((float)ck) : -1.0f;
            int piv = tid;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, mag, off);
                const int op = __shfl_xor_sync(0xffffffffu, piv, off);
                if (om > mag || (om == mag && op < piv)) { mag = om; piv = op; }
            }
            // factors of the rows AFTER the swap k <-> piv: row k's old value moves to slot piv
            const double cpiv = __shfl_sync(0xffffffffu, ck, piv), ckk = __shfl_sync(0xffffffffu, ck, k);
            if (tid < N) s_col[tid] = tid == k ? 0.0 : (tid == piv ? ckk : ck);
            if (tid == 0) {
                // 1/pivot: fp32 reciprocal + two Newton steps in fp64 (full double precision without the
                // ~100-instruction IEEE division on the critical path of every elimination step)
                double r = (double)(1.0f / (float)cpiv);
                r = r * (2.0 - cpiv * r);
                r = r * (2.0 - cpiv * r);
                r = r * (2.0 - cpiv * r);
                s_piv = piv; s_inv = cpiv != 0.0 ? r : 1.0 / cpiv;
            }
        }
        __syncthreads();
        const int j = tid & (SH_COLS - 1), g = tid / SH_COLS;
        const int piv = s_piv;
        const bool act = j >= k && j < NC;
        double pr = 0.0, rowk_old = 0.0;
        if (act) {
            rowk_old = s_a[k * SH_LD + j];
            pr = s_a[piv * SH_LD + j] * s_inv;                 // scaled pivot row entry
        }
        __syncthreads();                                       // every row group has read rows k and piv
        if (act) {
            if (g == 0) s_a[k * SH_LD + j] = pr;
            // the swap k <-> piv needs no store of its own: slot piv receives row k's old value minus its elimination
            // term from the ONE thread that owns (piv, j) below (a separate `s_a[piv] = rowk_old` by row group 0 raced
            // with that store whenever piv % SH_ROWG != 0)
            for (int i = g; i < N; i += SH_ROWG) {
                if (i == k) continue;
                const double base = (i == piv && piv != k) ? rowk_old : s_a[i * SH_LD + j];
                s_a[i * SH_LD + j] = base - s_col[i] * pr;
            }
        }
    }
    __syncthreads();
    // solution X[i][f][cc] = a[i][N + 2f + cc]
    if (TRANSPOSED) {
        for (int e = tid; e < nf * 2 * pn; e += SH_THREADS) {
            const int f = e / (2 * pn), r = e - f * 2 * pn;          // r = i*2 + cc
            out[(size_t)(b0 + f) * pn * 2 + r] = (float)s_a[(r >> 1) * SH_LD + N + 2 * f + (r & 1)];
        }
    } else {
        for (int e = tid; e < nf * 2 * N; e += SH_THREADS) {
            const int f = e / (2 * N), r = e - f * 2 * N;            // r = cc*N + i
            out[(size_t)(b0 + f) * 2 * N + r] = (float)s_a[(r % N) * SH_LD + N + 2 * f + (r / N)];
        }
    }
}

// ---- large systems: W^-1 into a global workspace, one CTA per distinct system ---------------
constexpr int INV_THREADS = 1024;

__global__ void __launch_bounds__(INV_THREADS) tps_inverse_kernel(const float* __restrict__ coord, long long coord_stride, int pn,
                                                                  double* __restrict__ work /* [nsys][N][2N] */, int* __restrict__ status,
                                                                  int elastic = 0) {
    extern __shared__ double s_buf[];   // pivot row (2N) + factor column (N)
    __shared__ double s_red[32];
    __shared__ int s_redi[32];
    __shared__ int s_piv;
    const int N = pn + 3, M = 2 * N;
    const int tid = threadIdx.x;
    const float* c = coord + (size_t)blockIdx.x * coord_stride;
    double* a = work + (size_t)blockIdx.x * N * M;
    double* s_row = s_buf;
    double* s_col = s_buf + M;
    for (int e = tid; e < N * M; e += INV_THREADS) {
        const int i = e / M, j = e % M;
        a[e] = j < N ? (double)(elastic ? elastic_l_entry(c, pn, i, j) : tps_w_entry(c, pn, i, j)) : (j - N == i ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int k = 0; k < N; ++k) {
        // pivot search in column k, rows k..N-1
        double mag = -1.0;
        int piv = 0x7fffffff;
        for (int i = k + tid; i < N; i += INV_THREADS) {
            const double v = fabs(a[(size_t)i * M + k]);
            if (v > mag) { mag = v; piv = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double om = __shfl_xor_sync(0xffffffffu, mag, off);
            const int op = __shfl_xor_sync(0xffffffffu, piv, off);
            if (om > mag || (om == mag && op < piv)) { mag = om; piv = op; }
        }
        if ((tid & 31) == 0) { s_red[tid >> 5] = mag; s_redi[tid >> 5] = piv; }
        __syncthreads();
        if (tid < 32) {
            mag = s_red[tid]; piv = s_redi[tid];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double om = __shfl_xor_sync(0xffffffffu, mag, off);
                const int op = __shfl_xor_sync(0xffffffffu, piv, off);
                if (om > mag || (om == mag && op < piv)) { mag = om; piv = op; }
            }
            if (tid == 0) {
                s_piv = piv;
                if (!(mag > 0.0) && status) atomicExch(status, 1);   // singular (tf.matrix_inverse would raise)
            }
        }
        __syncthreads();
        piv = s_piv;
        // stage the (scaled) pivot row and the factor column; swap rows k <-> piv on the fly
        const double inv = 1.0 / a[(size_t)piv * M + k];
        for (int j = tid; j < M; j += INV_THREADS) s_row[j] = a[(size_t)piv * M + j] * inv;
        for (int i = tid; i < N; i += INV_THREADS) {
            const int src = i == k ? piv : (i == piv ? k : i);     // row that will live at i after the swap
            s_col[i] = i == k ? 0.0 : a[(size_t)src * M + k];
        }
        __syncthreads();
        if (piv != k) {
            for (int j = tid; j < M; j += INV_THREADS) a[(size_t)piv * M + j] = a[(size_t)k * M + j];
        }
        __syncthreads();
        for (int e = tid; e < N * M; e += INV_THREADS) {
            const int i = e / M, j = e % M;
            a[e] = i == k ? s_row[j] : a[e] - s_col[i] * s_row[j];
        }
        __syncthreads();
    }
}

// forward : T[b][c][i]           = sum_{j<pn} Winv[i][j]     * target[b][j][c]
// backward: grad_target[b][j][c] = sum_{i<N}  Winv[i][j]     * grad_T[b][c][i]     (j < pn)
// OFFSETS (forward only): rhs holds the regressed offsets `vector` and the right-hand side is coord + vector
// (ThinPlateSpline.py:161), added here in fp32 exactly as the separate elementwise add would
template <bool TRANSPOSED, bool OFFSETS = false>
__global__ void tps_apply_kernel(const double* __restrict__ work, int shared_sys, const float* __restrict__ rhs,
                                 float* __restrict__ out, int B, int pn, const float* __restrict__ coord = nullptr,
                                 long long coord_stride = 0) {
    const int N = pn + 3, M = 2 * N;
    const int n_out = TRANSPOSED ? pn : N;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * n_out) return;
    const int b = (int)(gid / n_out), o = (int)(gid % n_out);
    const double* winv = work + (size_t)(shared_sys ? 0 : b) * N * M + N;   // right half of [I | W^-1]
    double a0 = 0.0, a1 = 0.0;
    if (TRANSPOSED) {
        for (int i = 0; i < N; ++i) {
            const double w = winv[(size_t)i * M + o];
            a0 += w * (double)rhs[((size_t)b * 2 + 0) * N + i];
            a1 += w * (double)rhs[((size_t)b * 2 + 1) * N + i];
        }
        out[((size_t)b * pn + o) * 2 + 0] = (float)a0;
        out[((size_t)b * pn + o) * 2 + 1] = (float)a1;
    } else {
        const float* cb = OFFSETS ? coord + (size_t)b * coord_stride : nullptr;
        for (int j = 0; j < pn; ++j) {
            const double w = winv[(size_t)o * M + j];
            float r0 = rhs[((size_t)b * pn + j) * 2 + 0], r1 = rhs[((size_t)b * pn + j) * 2 + 1];
            if (OFFSETS) { r0 = __fadd_rn(cb[2 * j], r0); r1 = __fadd_rn(cb[2 * j + 1], r1); }
            a0 += w * (double)r0;
            a1 += w * (double)r1;
        }
        out[((size_t)b * 2 + 0) * N + o] = (float)a0;
        out[((size_t)b * 2 + 1) * N + o] = (float)a1;
    }
}

// ElasticTransformer._transform (spatial_transformer.py:283-285): coefficients[b][c][o] = sum_i theta[b][c][i] * L_inv[o][3 + i]
// with theta [B,2,pn] the absolute target coordinates (all x, then all y: :161); TRANSPOSED: the gradient w.r.t. theta,
// grad_theta[b][c][i] = sum_o grad_coef[b][c][o] * L_inv[o][3 + i].  fp64 accumulation, rounded once.
template <bool TRANSPOSED>
__global__ void elastic_apply_kernel(const double* __restrict__ work, const float* __restrict__ in, float* __restrict__ out, int B, int pn) {
    const int N = pn + 3, M = 2 * N;
    const int n_out = TRANSPOSED ? pn : N;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * 2 * n_out) return;
    const int o = (int)(gid % n_out);
    const long long bc = gid / n_out;
    const double* linv = work + N;      // right half of [I | L^-1]
    double a = 0.0;
    if (TRANSPOSED) { for (int q = 0; q < N; ++q) a += linv[(size_t)q * M + 3 + o] * (double)in[bc * N + q]; }
    else { for (int i = 0; i < pn; ++i) a += linv[(size_t)o * M + 3 + i] * (double)in[bc * pn + i]; }
    out[bc * n_out + o] = (float)a;
}

static size_t big_workspace_bytes(int B, int pn, long long stride) {
    const size_t N = (size_t)pn + 3;
    const size_t nsys = stride == 0 ? 1 : (size_t)B;
    return nsys * N * 2 * N * sizeof(double) + 256;   // + status word, kept 256-B apart
}

// prepared: 0 = invert into the workspace, then apply; 1 = the workspace already holds the inverse(s) of this mesh
// (dvsg_tps_prepare), apply only; 2 = invert only
template <bool TRANSPOSED>
static int solve_impl(const float* coord, long long stride, const float* rhs, float* out, int B, int pn, void* ws,
                      size_t ws_bytes, cudaStream_t st, const char* what, int prepared = 0, bool offsets = false) {
    DVSG_REQUIRE(B >= 0 && pn >= 3, "%s: need B >= 0 and at least 3 control points (got B=%d pn=%d)", what, B, pn);
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(coord && (prepared == 2 || (rhs && out)), "%s: null pointer", what);
    DVSG_REQUIRE(stride == 0 || stride >= 2LL * pn, "%s: coord stride %lld < 2*pn", what, stride);
    const int N = pn + 3;
    if (N <= SMALL_N && prepared == 0 && stride == 0) {
        tps_solve_shared_kernel<TRANSPOSED><<<(B + SH_FRAMES - 1) / SH_FRAMES, SH_THREADS, 0, st>>>(coord, rhs, out, B, pn);
        count_launch();
        return check_launch("tps_solve_shared_kernel");
    }
    if (N <= SMALL_N && prepared == 0) {
        tps_solve_warp_kernel<TRANSPOSED><<<(B + SOLVE_WARPS - 1) / SOLVE_WARPS, SOLVE_WARPS * 32, 0, st>>>(coord, stride, rhs, out, B, pn);
        count_launch();
        return check_launch("tps_solve_warp_kernel");
    }
    DVSG_REQUIRE(N <= 2048, "%s: %d control points exceed the supported maximum 2045", what, pn);
    const size_t need = big_workspace_bytes(B, pn, stride);
    if (!ws || ws_bytes < need) {
        set_error("%s: workspace of %zu bytes required, %zu given", what, need, ws_bytes);
        return DVSG_ERR_WORKSPACE;
    }
    DVSG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7u) == 0, "%s: workspace must be 8-byte aligned", what);
    const int nsys = stride == 0 ? 1 : B;
    double* work = reinterpret_cast<double*>(ws);
    if (prepared != 1) {
        // singular systems (duplicate / collinear control points; tf.matrix_inverse raises there) set the status word kept
        // in the last 256 bytes of the workspace: read it back with dvsg_tps_prepare_status
        int* status = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws) + need - 256);
        if (cudaMemsetAsync(status, 0, sizeof(int), st) != cudaSuccess) return check_launch("tps_inverse status reset");
        const size_t inv_smem = (size_t)3 * N * sizeof(double);      // 48 KB at the maximum N = 2048: above the default limit with the static arrays
        static bool attr_done = false;
        if (!attr_done) { cudaFuncSetAttribute(tps_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 2048 * (int)sizeof(double)); attr_done = true; }
        tps_inverse_kernel<<<nsys, INV_THREADS, inv_smem, st>>>(coord, stride, pn, work, status);
        count_launch();
        const int rc = check_launch("tps_inverse_kernel");
        if (rc || prepared == 2) return rc;
    }
    const long long n_out = (long long)B * (TRANSPOSED ? pn : N);
    if (offsets && !TRANSPOSED)
        tps_apply_kernel<false, true><<<(unsigned)((n_out + 127) / 128), 128, 0, st>>>(work, stride == 0, rhs, out, B, pn, coord, stride);
    else
        tps_apply_kernel<TRANSPOSED><<<(unsigned)((n_out + 127) / 128), 128, 0, st>>>(work, stride == 0, rhs, out, B, pn);
    count_launch();
    return check_launch("tps_apply_kernel");
}

// ---- gradient w.r.t. the control-point positions `coord` ------------------------------------------------------------
// No reference call site differentiates the mesh (model.py:62-68 builds it as a constant); provided for completeness.
// coord enters ThinPlateSpline.py in three places:
//   (1) the radial terms of the dense grid, r_k(pix) = f(|pix - c_k|^2), f(d2) = d2 log(d2 + 1e-6)   (:100-105, :129)
//         d c_k += sum_pix (gx Tx_k + gy Ty_k) f'(d2) * 2 (c_k - pix),   f'(d2) = log(d2 + 1e-6) + d2 / (d2 + 1e-6)
//   (2) the system matrix W of the solve (:147-159), through  Y = W^-1 tp:   dW = -G Y^T,  G = W^-T dY
//         d c_k += dW[k][1|2] + dW[pn+1|pn+2][3+k] + sum_j (dW[k][3+j] + dW[j][3+k]) f'(d2_kj) * 2 (c_k - c_j)
//   (3) the right-hand side coord + vector of ThinPlateSpline (:161) -- that term equals the gradient w.r.t. `vector`
//       and is added by the caller.
// One CTA per (control point, frame); fp64 throughout; deterministic.
constexpr int CB_THREADS = 256;

__device__ __forceinline__ double tps_fprime(double d2) { return log(d2 + 1e-6) + d2 / (d2 + 1e-6); }

__global__ void __launch_bounds__(CB_THREADS) tps_coord_bwd_kernel(const double* __restrict__ work, int shared_sys, const float* __restrict__ coord,
                                                                  long long coord_stride, const float* __restrict__ T,
                                                                  const float* __restrict__ grad_T, const float* __restrict__ gx,
                                                                  const float* __restrict__ gy, float* __restrict__ grad_coord, int oh, int ow, int pn,
                                                                  float step_x, float step_y) {
    extern __shared__ double s_g[];                       // G [N][2]
    __shared__ double s_rx[CB_THREADS / 32], s_ry[CB_THREADS / 32];
    const int k = blockIdx.x, b = blockIdx.y, N = pn + 3, M = 2 * N, tid = threadIdx.x;
    const float* cb = coord + (size_t)b * coord_stride;
    const float* Tb = T + (size_t)b * 2 * N;
    const float* gTb = grad_T + (size_t)b * 2 * N;
    const double* winv = work + (size_t)(shared_sys ? 0 : b) * N * M + N;
    const double ckx = (double)cb[2 * k], cky = (double)cb[2 * k + 1];
    double ax = 0.0, ay = 0.0;
    // (2) G[i][c] = sum_m Winv[m][i] * grad_T[c][m]
    for (int i = tid; i < N; i += CB_THREADS) {
        double g0 = 0.0, g1 = 0.0;
        for (int m = 0; m < N; ++m) {
            const double w = winv[(size_t)m * M + i];
            g0 += w * (double)gTb[m];
            g1 += w * (double)gTb[N + m];
        }
        s_g[2 * i] = g0; s_g[2 * i + 1] = g1;
    }
    __syncthreads();
    // dW[i][j] = -(G[i][0] Y[j][0] + G[i][1] Y[j][1]),  Y[j][c] = T[c][j]
    auto dW = [&](int i, int j) { return -(s_g[2 * i] * (double)Tb[j] + s_g[2 * i + 1] * (double)Tb[N + j]); };
    if (tid == 0) {
        ax += dW(k, 1) + dW(pn + 1, 3 + k);
        ay += dW(k, 2) + dW(pn + 2, 3 + k);
    }
    for (int j = tid; j < pn; j += CB_THREADS) {
        if (j == k) continue;                             // 2 (c_k - c_k) = 0
        const double dx = ckx - (double)cb[2 * j], dy = cky - (double)cb[2 * j + 1];
        // the reference forms d2 in fp32 (:152); the gradient is taken at that value
        const double d2 = (double)tps_d2(cb[2 * k], cb[2 * k + 1], cb[2 * j], cb[2 * j + 1]);
        const double f = (dW(k, 3 + j) + dW(j, 3 + k)) * tps_fprime(d2) * 2.0;
        ax += f * dx; ay += f * dy;
    }
    // (1) the dense grid
    if (gx != nullptr) {
        const double tx = (double)Tb[3 + k], ty = (double)Tb[N + 3 + k];
        const long long n = (long long)oh * ow;
        const float* gxb = gx + (size_t)b * n;
        const float* gyb = gy + (size_t)b * n;
        for (long long pix = tid; pix < n; pix += CB_THREADS) {
            const int row = (int)(pix / ow), col = (int)(pix % ow);
            const float xt = lin_coord(col, step_x), yt = lin_coord(row, step_y);
            const double w = (double)__ldg(gxb + pix) * tx + (double)__ldg(gyb + pix) * ty;
            const double d2 = (double)tps_d2(xt, yt, cb[2 * k], cb[2 * k + 1]);
            const double f = w * tps_fprime(d2) * 2.0;
            ax += f * (ckx - (double)xt); ay += f * (cky - (double)yt);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o); }
    if ((tid & 31) == 0) { s_rx[tid >> 5] = ax; s_ry[tid >> 5] = ay; }
    __syncthreads();
    if (tid == 0) {
        double sx = 0.0, sy = 0.0;
        for (int w = 0; w < CB_THREADS / 32; ++w) { sx += s_rx[w]; sy += s_ry[w]; }
        grad_coord[((size_t)b * pn + k) * 2] = (float)sx;
        grad_coord[((size_t)b * pn + k) * 2 + 1] = (float)sy;
    }
}

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_tps_coord_bwd(const float* coord, long long coord_batch_stride, const float* T, const float* grad_T, const float* grad_x,
                                  const float* grad_y, float* grad_coord, int B, int oh, int ow, int pn, const void* workspace,
                                  size_t workspace_bytes, void* stream) {
    DVSG_REQUIRE(B >= 0 && pn >= 3 && oh >= 0 && ow >= 0, "tps_coord_bwd: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(coord && T && grad_T && grad_coord && workspace, "tps_coord_bwd: null pointer");
    DVSG_REQUIRE((grad_x == nullptr) == (grad_y == nullptr), "tps_coord_bwd: grad_x and grad_y must be given together");
    DVSG_REQUIRE(coord_batch_stride == 0 || coord_batch_stride >= 2LL * pn, "tps_coord_bwd: coord stride %lld < 2*pn", coord_batch_stride);
    DVSG_REQUIRE(B <= 65535 && pn + 3 <= 2048 && (long long)oh * ow < (1LL << 31), "tps_coord_bwd: batch / mesh / frame too large");
    const size_t need = big_workspace_bytes(B, pn, coord_batch_stride);
    DVSG_REQUIRE(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0,
                 "tps_coord_bwd: the workspace dvsg_tps_prepare filled for this mesh is required (%zu bytes, 8-byte aligned; %zu given)", need,
                 workspace_bytes);
    const bool grid = grad_x != nullptr && oh > 0 && ow > 0;
    tps_coord_bwd_kernel<<<dim3((unsigned)pn, (unsigned)B), CB_THREADS, (size_t)2 * (pn + 3) * sizeof(double), (cudaStream_t)stream>>>(
        reinterpret_cast<const double*>(workspace), coord_batch_stride == 0, coord, coord_batch_stride, T, grad_T, grid ? grad_x : nullptr,
        grid ? grad_y : nullptr, grad_coord, oh, ow, pn, ow > 1 ? 2.0f / (float)(ow - 1) : 0.0f, oh > 1 ? 2.0f / (float)(oh - 1) : 0.0f);
    count_launch();
    return check_launch("tps_coord_bwd_kernel");
}

extern "C" size_t dvsg_tps_solve_workspace_bytes(int B, int pn, long long coord_batch_stride) {
    if (pn + 3 <= SMALL_N || B <= 0) return 0;
    return big_workspace_bytes(B, pn, coord_batch_stride);
}

extern "C" int dvsg_tps_solve(const float* coord, long long coord_batch_stride, const float* target, float* T, int B, int pn,
                              void* workspace, size_t workspace_bytes, void* stream) {
    return solve_impl<false>(coord, coord_batch_stride, target, T, B, pn, workspace, workspace_bytes, (cudaStream_t)stream, "tps_solve");
}

// The mesh of a clip is a constant (model.py:62-68): dvsg_tps_prepare inverts its system(s) into the
// workspace once, dvsg_tps_solve_prepared / _bwd_prepared then only apply W^-1 (a [B,N] x [N,N] product) per call.
// Works for every mesh size: with a 4x4 mesh the per-call work drops from a 23 us factorisation kernel to a ~3 us
// [B,16] x [16,19] product (size the workspace with dvsg_tps_prepare_workspace_bytes).
extern "C" size_t dvsg_tps_prepare_workspace_bytes(int B, int pn, long long coord_batch_stride) {
    if (B <= 0 || pn < 3) return 0;
    return big_workspace_bytes(B, pn, coord_batch_stride);
}

extern "C" int dvsg_tps_prepare(const float* coord, long long coord_batch_stride, int B, int pn, void* workspace, size_t workspace_bytes,
                                void* stream) {
    return solve_impl<false>(coord, coord_batch_stride, nullptr, nullptr, B, pn, workspace, workspace_bytes, (cudaStream_t)stream, "tps_prepare", 2);
}

// 0 = every system prepared in this workspace was invertible, 1 = at least one pivot was exactly zero (the reference's
// tf.matrix_inverse raises InvalidArgument there, ThinPlateSpline.py:159).  Synchronises `stream`.
extern "C" int dvsg_tps_prepare_status(const void* workspace, size_t workspace_bytes, int B, int pn, long long coord_batch_stride,
                                       void* stream, int* singular_out) {
    DVSG_REQUIRE(workspace && singular_out && B > 0 && pn >= 3, "tps_prepare_status: bad argument");
    const size_t need = big_workspace_bytes(B, pn, coord_batch_stride);
    DVSG_REQUIRE(workspace_bytes >= need, "tps_prepare_status: workspace of %zu bytes required, %zu given", need, workspace_bytes);
    const int* status = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(workspace) + need - 256);
    if (cudaMemcpyAsync(singular_out, status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
        cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        return check_launch("tps_prepare_status");
    return DVSG_OK;
}

extern "C" int dvsg_tps_solve_prepared(const float* coord, long long coord_batch_stride, const float* target, float* T, int B, int pn,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    return solve_impl<false>(coord, coord_batch_stride, target, T, B, pn, workspace, workspace_bytes, (cudaStream_t)stream, "tps_solve_prepared", 1);
}

// same with the regressed offsets as the argument: the right-hand side coord + vector is formed inside the apply kernel
extern "C" int dvsg_tps_solve_offsets_prepared(const float* coord, long long coord_batch_stride, const float* vector, float* T, int B, int pn,
                                               void* workspace, size_t workspace_bytes, void* stream) {
    return solve_impl<false>(coord, coord_batch_stride, vector, T, B, pn, workspace, workspace_bytes, (cudaStream_t)stream,
                             "tps_solve_offsets_prepared", 1, true);
}

extern "C" int dvsg_tps_solve_bwd_prepared(const float* coord, long long coord_batch_stride, const float* grad_T, float* grad_target, int B,
                                           int pn, void* workspace, size_t workspace_bytes, void* stream) {
    return solve_impl<true>(coord, coord_batch_stride, grad_T, grad_target, B, pn, workspace, workspace_bytes, (cudaStream_t)stream,
                            "tps_solve_bwd_prepared", 1);
}

extern "C" int dvsg_tps_solve_bwd(const float* coord, long long coord_batch_stride, const float* grad_T, float* grad_target,
                                  int B, int pn, void* workspace, size_t workspace_bytes, void* stream) {
    return solve_impl<true>(coord, coord_batch_stride, grad_T, grad_target, B, pn, workspace, workspace_bytes, (cudaStream_t)stream,
                            "tps_solve_bwd");
}

// ---- ElasticTransformer (spatial_transformer.py:93-362): the second TPS formulation of the reference ------------------
// One system for the layer's constant regular mesh, inverted once (dvsg_elastic_prepare = _initialize_tps), applied per call.
extern "C" size_t dvsg_elastic_workspace_bytes(int pn) { return pn >= 1 ? big_workspace_bytes(1, pn, 0) : 0; }

extern "C" int dvsg_elastic_prepare(const float* source_points, int pn, void* workspace, size_t workspace_bytes, void* stream) {
    DVSG_REQUIRE(source_points && workspace && pn >= 3 && pn + 3 <= 2048, "elastic_prepare: bad argument");
    const size_t need = big_workspace_bytes(1, pn, 0);
    DVSG_REQUIRE(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, "elastic_prepare: workspace of %zu bytes (8-byte aligned) required", need);
    const int N = pn + 3;
    cudaStream_t st = (cudaStream_t)stream;
    int* status = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(workspace) + need - 256);
    if (cudaMemsetAsync(status, 0, sizeof(int), st) != cudaSuccess) return check_launch("elastic_prepare status reset");
    cudaFuncSetAttribute(tps_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 2048 * (int)sizeof(double));
    tps_inverse_kernel<<<1, INV_THREADS, (size_t)3 * N * sizeof(double), st>>>(source_points, 0, pn, reinterpret_cast<double*>(workspace), status, 1);
    count_launch();
    return check_launch("tps_inverse_kernel (elastic)");
}

extern "C" int dvsg_elastic_solve(const float* theta_abs, float* coef, int B, int pn, const void* workspace, size_t workspace_bytes, void* stream) {
    DVSG_REQUIRE(B >= 0 && pn >= 3, "elastic_solve: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(theta_abs && coef && workspace && workspace_bytes >= big_workspace_bytes(1, pn, 0), "elastic_solve: bad argument");
    const long long n = (long long)B * 2 * (pn + 3);
    elastic_apply_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(workspace), theta_abs, coef, B, pn);
    count_launch();
    return check_launch("elastic_apply_kernel");
}

extern "C" int dvsg_elastic_solve_bwd(const float* grad_coef, float* grad_theta, int B, int pn, const void* workspace, size_t workspace_bytes,
                                      void* stream) {
    DVSG_REQUIRE(B >= 0 && pn >= 3, "elastic_solve_bwd: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(grad_coef && grad_theta && workspace && workspace_bytes >= big_workspace_bytes(1, pn, 0), "elastic_solve_bwd: bad argument");
    const long long n = (long long)B * 2 * pn;
    elastic_apply_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(workspace), grad_coef, grad_theta, B, pn);
    count_launch();
    return check_launch("elastic_apply_kernel (bwd)");
}
