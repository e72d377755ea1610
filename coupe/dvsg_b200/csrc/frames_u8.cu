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

// ---- ingest with the resize of eval.py:80: cv2.resize(frame / 255., (out_w, out_h)), INTER_LINEAR on float64 ---------
// OpenCV's linear resize of a CV_64F image (third-party, version not pinned by the reference; restated from its
// observable behaviour -- impulse responses and 9 size pairs compared in tests/test_oracle_kat.py against cv2 4.13):
// pixel centres aligned, fx = (dx + 0.5) * scale - 0.5 with scale = 1 / (dst / src) in double, sx = floor(fx),
// fx -= sx; sx < 0 -> (0, fx = 0), sx >= src - 1 -> (src - 1, fx = 0); rows likewise but clamped without touching the
// weight; horizontal pass first, all in double, cast to fp32 by the feed.  The same formulas in double here.
__global__ void __launch_bounds__(256) u8_resize_to_f32_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int Hs, int Ws,
                                                               int h, int w, double scale_x, double scale_y, int swap_rb) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, b = blockIdx.z;
    if (dx >= w) return;
    double fx = __dadd_rn(__dmul_rn((double)dx + 0.5, scale_x), -0.5);
    int sx = (int)floor(fx);
    fx = __dadd_rn(fx, -(double)sx);
    if (sx < 0) { sx = 0; fx = 0.0; }
    if (sx >= Ws - 1) { sx = Ws - 1; fx = 0.0; }
    const int sx1 = min(sx + 1, Ws - 1);
    double fy = __dadd_rn(__dmul_rn((double)dy + 0.5, scale_y), -0.5);
    const int sy = (int)floor(fy);
    fy = __dadd_rn(fy, -(double)sy);
    const int sy0 = min(max(sy, 0), Hs - 1), sy1 = min(max(sy + 1, 0), Hs - 1);
    const double a0 = __dadd_rn(1.0, -fx), a1 = fx, b0 = __dadd_rn(1.0, -fy), b1 = fy;
    const unsigned char* f = src + (size_t)b * Hs * Ws * 3;
    const unsigned char* p00 = f + ((size_t)sy0 * Ws + sx) * 3;
    const unsigned char* p01 = f + ((size_t)sy0 * Ws + sx1) * 3;
    const unsigned char* p10 = f + ((size_t)sy1 * Ws + sx) * 3;
    const unsigned char* p11 = f + ((size_t)sy1 * Ws + sx1) * 3;
    float* o = dst + (((size_t)b * h + dy) * w + dx) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int cs = swap_rb ? 2 - ch : ch;      // cvtColor(BGR2RGB) comes first (eval.py:79)
        const double s00 = __ddiv_rn((double)__ldg(p00 + cs), 255.0), s01 = __ddiv_rn((double)__ldg(p01 + cs), 255.0);
        const double s10 = __ddiv_rn((double)__ldg(p10 + cs), 255.0), s11 = __ddiv_rn((double)__ldg(p11 + cs), 255.0);
        const double r0 = __dadd_rn(__dmul_rn(s00, a0), __dmul_rn(s01, a1));
        const double r1 = __dadd_rn(__dmul_rn(s10, a0), __dmul_rn(s11, a1));
        o[ch] = (float)__dadd_rn(__dmul_rn(r0, b0), __dmul_rn(r1, b1));
    }
}

// Flow ingest (data_loader.py:239): cv2.resize(np.load(flow), (w, h)) * [w, h] for a float32 [Hs, Ws, 2] field.  cv2's
// INTER_LINEAR on float32 input, restated from its observable behaviour: coordinates (dx + 0.5) * scale - 0.5 in double,
// cast to float; taps and sums in float with separately rounded products (no FMA) -- bit-identical to cv2 4.13 on every case
// tested (tests compare with cv2 itself); then the float64 product with (w, h), rounded once by the float32 feed, which is
// the correctly rounded float product.
__global__ void __launch_bounds__(256) flow_resize_scale_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int Hs, int Ws, int h,
                                                                int w, double scale_x, double scale_y) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, b = blockIdx.z;
    if (dx >= w) return;
    float fx = (float)__dadd_rn(__dmul_rn((double)dx + 0.5, scale_x), -0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { sx = 0; fx = 0.0f; }
    if (sx >= Ws - 1) { sx = Ws - 1; fx = 0.0f; }
    const int sx1 = min(sx + 1, Ws - 1);
    float fy = (float)__dadd_rn(__dmul_rn((double)dy + 0.5, scale_y), -0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int sy0 = min(max(sy, 0), Hs - 1), sy1 = min(max(sy + 1, 0), Hs - 1);
    const float a0 = __fsub_rn(1.0f, fx), a1 = fx, b0 = __fsub_rn(1.0f, fy), b1 = fy;
    const float2* f = src + (size_t)b * Hs * Ws;
    const float2 s00 = __ldg(f + (size_t)sy0 * Ws + sx), s01 = __ldg(f + (size_t)sy0 * Ws + sx1);
    const float2 s10 = __ldg(f + (size_t)sy1 * Ws + sx), s11 = __ldg(f + (size_t)sy1 * Ws + sx1);
    const float r0x = __fadd_rn(__fmul_rn(s00.x, a0), __fmul_rn(s01.x, a1)), r1x = __fadd_rn(__fmul_rn(s10.x, a0), __fmul_rn(s11.x, a1));
    const float r0y = __fadd_rn(__fmul_rn(s00.y, a0), __fmul_rn(s01.y, a1)), r1y = __fadd_rn(__fmul_rn(s10.y, a0), __fmul_rn(s11.y, a1));
    const float vx = __fadd_rn(__fmul_rn(r0x, b0), __fmul_rn(r1x, b1)), vy = __fadd_rn(__fmul_rn(r0y, b0), __fmul_rn(r1y, b1));
    dst[((size_t)b * h + dy) * w + dx] = make_float2(__fmul_rn(vx, (float)w), __fmul_rn(vy, (float)h));
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

extern "C" int dvsg_frames_u8_resize_to_f32(const unsigned char* src, float* dst, int B, int Hs, int Ws, int h, int w, int swap_rb,
                                            void* stream) {
    DVSG_REQUIRE(B >= 0 && Hs >= 2 && Ws >= 2 && h > 0 && w > 0, "frames_u8_resize_to_f32: bad shape (source frames need at least 2 x 2 pixels)");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(src && dst, "frames_u8_resize_to_f32: null pointer");
    DVSG_REQUIRE(B <= 65535 && h <= 65535, "frames_u8_resize_to_f32: batch %d / height %d exceed the grid limits", B, h);
    // cv::resize: inv_scale = dsize / ssize, scale = 1 / inv_scale (both double)
    const double scale_x = 1.0 / ((double)w / (double)Ws), scale_y = 1.0 / ((double)h / (double)Hs);
    u8_resize_to_f32_kernel<<<dim3((unsigned)((w + 255) / 256), (unsigned)h, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        src, dst, Hs, Ws, h, w, scale_x, scale_y, swap_rb);
    count_launch();
    return check_launch("u8_resize_to_f32_kernel");
}

extern "C" int dvsg_flow_resize_scale(const float* src, float* dst, int B, int Hs, int Ws, int h, int w, void* stream) {
    DVSG_REQUIRE(B >= 0 && Hs >= 1 && Ws >= 1 && h > 0 && w > 0, "flow_resize_scale: bad shape");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(src && dst, "flow_resize_scale: null pointer");
    DVSG_REQUIRE((reinterpret_cast<uintptr_t>(src) & 7u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0, "flow_resize_scale: buffers must be 8-byte aligned");
    DVSG_REQUIRE(B <= 65535 && h <= 65535 && w < (1 << 24) && h < (1 << 24), "flow_resize_scale: batch %d / size %d x %d exceed the limits", B, h, w);
    const double scale_x = 1.0 / ((double)w / (double)Ws), scale_y = 1.0 / ((double)h / (double)Hs);
    flow_resize_scale_kernel<<<dim3((unsigned)((w + 255) / 256), (unsigned)h, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(src), reinterpret_cast<float2*>(dst), Hs, Ws, h, w, scale_x, scale_y);
    count_launch();
    return check_launch("flow_resize_scale_kernel");
}
