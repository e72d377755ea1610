// dvsg_common.cuh -- error plumbing, launch bookkeeping and the sm_100a PTX wrappers
// (mbarrier + 1-D bulk async copies, i.e. the non-tensor TMA path: SASS UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/dvsg_warp.h"

namespace dvsg {

// ---- host-side error state (thread-local) ---------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);   // cudaGetLastError -> DVSG_ERR_CUDA

#define DVSG_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::dvsg::set_error(__VA_ARGS__);     \
            return DVSG_ERR_INVALID;            \
        }                                       \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Raise a kernel's dynamic shared-memory limit only when it has to grow (cudaFuncSetAttribute costs ~1 us per call, which the
// online loop pays per frame otherwise).  Keyed by the kernel's ADDRESS: template instantiations share one function-pointer
// type, so a `static` inside a generic lambda would be shared by all of them.
inline void ensure_dynamic_smem(const void* kernel, int bytes) {
    struct Entry { const void* k; int dev, bytes; };      // the attribute belongs to (kernel, device)
    static thread_local Entry seen[96];
    static thread_local int n_seen = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < n_seen; ++i)
        if (seen[i].k == kernel && seen[i].dev == dev) {
            if (bytes > seen[i].bytes) { cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); seen[i].bytes = bytes; }
            return;
        }
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (n_seen < 96) { seen[n_seen].k = kernel; seen[n_seen].dev = dev; seen[n_seen].bytes = bytes; ++n_seen; }
}

#if defined(__CUDACC__)
// ---- device-side PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy (16-B aligned addresses, size % 16 == 0), completion on mbar
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
// shared -> global bulk copy, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// single MUFU.LG2 (t is never denormal on this path: t >= 1e-30)
__device__ __forceinline__ float lg2_approx(float t) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}
#endif  // __CUDACC__

}  // namespace dvsg

#if defined(__CUDACC__)
#include "sampler_math.cuh"
namespace dvsg {
// ---- TPS radial term, shared by all four warp kernels (identical coordinates in every one of them) ---------
// Radial term: the reference's r = d2 * log(d2 + 1e-6) (ThinPlateSpline.py:105) is evaluated as d2 * log(d2) + 1e-6:
// d2 log(1 + 1e-6/d2) = 1e-6 - 1e-12/(2 d2) + ..., so the two differ by less than 1e-8 once d2 > 5e-5 (a pixel farther
// than 0.007 normalised units from the control point) and by at most 1e-6 next to it.  The constant is not added per
// term: sum_k c_k * 1e-6 is folded into the affine constant once per frame (it vanishes for coefficients that come out
// of the TPS system, whose last three rows force sum_k c_k = 0).  That removes one packed add per (pixel pair, control
// point).  d2 = 0 (a pixel exactly on a control point) would give 0 * -inf: the (y_t - p_y)^2 table entries are kept
// >= TPS_TINY, so d2 >= 1e-30 and the term is -1e-28 ~ -0 as in the reference.
constexpr float TPS_TINY = 1e-30f;
constexpr float TPS_EPS = 1e-6f;
// warp-collective: sum of pn coefficients in a fixed order (lane-strided partial sums, xor tree: same value on every lane)
__device__ __forceinline__ float tps_coef_sum(const float* __restrict__ c, int pn, int lane) {
    float s = 0.0f;
    for (int k = lane; k < pn; k += 32) s = __fadd_rn(s, c[k]);      // plain loads: the coefficients may live in shared memory
#pragma unroll
    for (int o = 16; o; o >>= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    return s;
}
// affine constant of one output row of T with the folded epsilon term (called by one whole warp)
__device__ __forceinline__ float tps_affine0(const float* __restrict__ Trow, int pn, int lane) {
    return __fmaf_rn(TPS_EPS, tps_coef_sum(Trow + 3, pn, lane), Trow[0]);
}
// (y_t - p_y)^2 as stored in the tables
__device__ __forceinline__ float tps_dy2(float yt, float py) {
    const float dy = DVSG_SUB(yt, py);
    return fmaxf(DVSG_MUL(dy, dy), TPS_TINY);
}
}  // namespace dvsg
#endif  // __CUDACC__
