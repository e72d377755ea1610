// frames_u8.cu -- frame ingest / egress on the device: the two conversions that sit
// immediately before and after the warp in the reference's inference loop,
//
//   ingest   eval.py:79-80    frame = cv2.cvtColor(frame, BGR2RGB); frame / 255.   (uint8 -> [0,1], fed as fp32)
//   egress   eval.py:112-113  np.uint8(frame * 255.); cv2.cvtColor(.., RGB2BGR)
//
// so that frames cross PCIe as 3 bytes per pixel instead of 12.  Both are exact restatements:
//   * u / 255 : numpy divides in fp64 and the feed casts to fp32; for the 256 possible inputs that equals the
//     correctly rounded fp32 quotient, which q = u*c, r = fma(-255, q, u), q' = fma(r, c, q) (c = fl(1/255))
//     reproduces bit for bit (checked for all 256 values by tests/test_cpu_library.py against numpy);
//   * uint8(x * 255.) : the fp64 product of an fp32 value and 255 is exact and the cast truncates toward zero;
//     mul.rz never crosses an integer, so trunc(mul.rz(x, 255)) is the same integer.  Out-of-range products
//     follow the x86 cast numpy performs (cvttsd2si, then the low byte); NaN -> 0.
// Both kernels are streaming (15 B per pixel): one thread per four output elements, coalesced vector stores,
// loads served by L1; the channel swap is index arithmetic (element i <-> i + 2 - 2*(i % 3)).
#include "dvsg_common.cuh"

namespace dvsg {

__device__ __forceinline__ float u8_to_unit(unsigned v) {
    const float c = 0.00392156885936856269836425781250f;       // fl32(1/255)
    const float u = (float)v;
    const float q = __fmul_rn(u, c);
    const float r = __fmaf_rn(-255.0f, q, u);
    return __fmaf_rn(r, c, q);
}

__device__ __forceinline__ unsigned unit_to_u8(float x) {
    const float t = __fmul_rz(x, 255.0f);
    // cvttsd2si semantics: |t| >= 2^31 and NaN give INT_MIN, whose low byte is 0
    const int v = (fabsf(t) < 2147483648.0f) ? __float2int_rz(t) : (int)0x80000000;
    return (unsigned)v & 0xffu;
}

// element i of a [n_px, 3] frame <-> element i + 2 - 2*(i % 3) of the channel-swapped frame
__device__ __forceinline__ long long swapped(long long i, int swap_rb) { return swap_rb ? i + 2 - 2 * (i % 3) : i; }
// the same for elements 4q .. 4q+3 with one 64-bit modulo: (4q + k) % 3 == (q + k) % 3
__device__ __forceinline__ void swapped4(long long q, int swap_rb, long long (&idx)[4]) {
    const int m = swap_rb ? (int)(q % 3) : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = 4 * q + k + (swap_rb ? 2 - 2 * ((m + k) % 3) : 0);
}

// One thread per 4 consecutive OUTPUT elements: stores are fully coalesced 128-bit (u8 -> f32) or 32-bit
// (f32 -> u8) accesses, the loads of a warp fall into one or two 128-byte lines and are served by L1.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, long long n, int swap_rb) {
    const long long n4 = n >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        long long j[4];
        swapped4(q, swap_rb, j);
        float4 v;
        v.x = u8_to_unit(__ldg(src + j[0]));
        v.y = u8_to_unit(__ldg(src + j[1]));
        v.z = u8_to_unit(__ldg(src + j[2]));
        v.w = u8_to_unit(__ldg(src + j[3]));
        *reinterpret_cast<float4*>(dst + 4 * q) = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // n % 4 trailing elements
        const long long i = 4 * n4 + threadIdx.x;
        dst[i] = u8_to_unit(__ldg(src + swapped(i, swap_rb)));
    }
}

__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ src, unsigned char* __restrict__ dst, long long n, int swap_rb) {
    const long long n4 = n >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
        long long j[4];
        swapped4(q, swap_rb, j);
        const unsigned w = unit_to_u8(__ldg(src + j[0])) | (unit_to_u8(__ldg(src + j[1])) << 8) | (unit_to_u8(__ldg(src + j[2])) << 16) |
                           (unit_to_u8(__ldg(src + j[3])) << 24);
        *reinterpret_cast<unsigned*>(dst + 4 * q) = w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = 4 * n4 + threadIdx.x;
        dst[i] = (unsigned char)unit_to_u8(__ldg(src + swapped(i, swap_rb)));
    }
}

// buffers whose OUTPUT is not aligned for the vector store: one element per thread
__global__ void u8_to_f32_elem_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, long long n, int swap_rb) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = u8_to_unit(__ldg(src + swapped(i, swap_rb)));
}
__global__ void f32_to_u8_elem_kernel(const float* __restrict__ src, unsigned char* __restrict__ dst, long long n, int swap_rb) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (unsigned char)unit_to_u8(__ldg(src + swapped(i, swap_rb)));
}

static int grid_for(long long n) {
    const long long blocks = (n + 255) / 256;
    return (int)(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);     // grid-stride: 8 CTAs of 256 threads per SM, two rounds
}

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_frames_u8_to_f32(const unsigned char* src, float* dst, long long n_pixels, int swap_rb, void* stream) {
    DVSG_REQUIRE(n_pixels >= 0, "frames_u8_to_f32: negative pixel count");
    if (n_pixels == 0) return DVSG_OK;
    DVSG_REQUIRE(src && dst, "frames_u8_to_f32: null pointer");
    const long long n = 3 * n_pixels;
    if (aligned16(dst)) u8_to_f32_kernel<<<grid_for(n / 4 + 4), 256, 0, (cudaStream_t)stream>>>(src, dst, n, swap_rb);
    else u8_to_f32_elem_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n, swap_rb);
    count_launch();
    return check_launch("u8_to_f32_kernel");
}

extern "C" int dvsg_frames_f32_to_u8(const float* src, unsigned char* dst, long long n_pixels, int swap_rb, void* stream) {
    DVSG_REQUIRE(n_pixels >= 0, "frames_f32_to_u8: negative pixel count");
    if (n_pixels == 0) return DVSG_OK;
    DVSG_REQUIRE(src && dst, "frames_f32_to_u8: null pointer");
    const long long n = 3 * n_pixels;
    if ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0) f32_to_u8_kernel<<<grid_for(n / 4 + 4), 256, 0, (cudaStream_t)stream>>>(src, dst, n, swap_rb);
    else f32_to_u8_elem_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n, swap_rb);
    count_launch();
    return check_launch("f32_to_u8_kernel");
}
