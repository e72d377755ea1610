// api.cu -- library-level pieces of the C ABI: version, thread-local error string, launch
// bookkeeping, and the host-buffer pipeline (H2D -> solve -> fused warp -> D2H).
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "dvsg_common.cuh"

namespace dvsg {

static thread_local char t_err[512] = "";
static thread_local long long t_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { t_launches += n; }

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DVSG_ERR_CUDA;
    }
    return DVSG_OK;
}

}  // namespace dvsg

using namespace dvsg;

extern "C" int dvsg_version(void) { return 100; }   // 0.1.0
extern "C" const char* dvsg_last_error(void) { return t_err; }
extern "C" long long dvsg_launch_count(void) { return t_launches; }

// ---- host-buffer pipeline ------------------------------------------------------------------
// eval.py:106-110 feeds every frame from host memory and reads the warped frame back.  This
// object is that call for a batch of frames: chunks of frames rotate through n_slots device
// staging slots, each slot on its own stream, so chunk i's kernels overlap chunk i+1's H2D
// copy and chunk i-1's D2H copy (two copy engines + SMs busy at once).
struct dvsg_host_pipeline {
    int device, H, W, C, pn, fpc, n_slots;
    size_t frame_elems;
    float* d_coord;                 // [pn, 2] shared mesh
    void* d_winv;                   // inverse of the mesh's TPS system (dvsg_tps_prepare), refreshed by every call
    size_t winv_bytes;
    std::vector<cudaStream_t> streams;
    std::vector<float*> d_in, d_out, d_vec, d_T;
    std::vector<unsigned char*> d_in8, d_out8;     // uint8 staging of the frames (allocated by the first u8 call)
    std::vector<float*> d_flow;                    // flow staging (allocated by the first tf_warp call)
    std::vector<float> h_coord;                    // mesh whose inverse d_winv holds (re-prepared only when the caller's mesh changes)
    int async;                                     // 1: calls return once their work is enqueued (dvsg_host_pipeline_sync waits)
    int next_chunk;                                // slots rotate across calls, so consecutive calls overlap in async mode
};

#define DVSG_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) {                                          \
            set_error("%s: %s", #call, cudaGetErrorString(e_));           \
            return DVSG_ERR_CUDA;                                         \
        }                                                                 \
    } while (0)

extern "C" int dvsg_host_pipeline_create(dvsg_host_pipeline** out, int device, int H, int W, int C, int pn,
                                         int frames_per_chunk, int n_slots) {
    DVSG_REQUIRE(out && H > 0 && W > 0 && C > 0 && pn >= 3 && frames_per_chunk > 0 && n_slots > 0 && n_slots <= 16,
                 "host_pipeline_create: bad argument");
    DVSG_CUDA(cudaSetDevice(device));
    dvsg_host_pipeline* p = new dvsg_host_pipeline();
    p->device = device; p->H = H; p->W = W; p->C = C; p->pn = pn; p->fpc = frames_per_chunk; p->n_slots = n_slots;
    p->frame_elems = (size_t)H * W * C;
    p->d_coord = nullptr;
    p->d_winv = nullptr;
    p->async = 0;
    p->next_chunk = 0;
    p->winv_bytes = dvsg_tps_prepare_workspace_bytes(1, pn, 0);
    *out = p;
    DVSG_CUDA(cudaMalloc(&p->d_coord, (size_t)pn * 2 * sizeof(float)));
    DVSG_CUDA(cudaMalloc(&p->d_winv, p->winv_bytes));
    for (int s = 0; s < n_slots; ++s) {
        cudaStream_t st;
        float *a = nullptr, *b = nullptr, *v = nullptr, *t = nullptr;
        DVSG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        p->streams.push_back(st);
        DVSG_CUDA(cudaMalloc(&a, p->frame_elems * frames_per_chunk * sizeof(float)));
        p->d_in.push_back(a);
        DVSG_CUDA(cudaMalloc(&b, p->frame_elems * frames_per_chunk * sizeof(float)));
        p->d_out.push_back(b);
        DVSG_CUDA(cudaMalloc(&v, (size_t)frames_per_chunk * pn * 2 * sizeof(float)));
        p->d_vec.push_back(v);
        DVSG_CUDA(cudaMalloc(&t, (size_t)frames_per_chunk * 2 * (pn + 3) * sizeof(float)));
        p->d_T.push_back(t);
    }
    return DVSG_OK;
}

extern "C" void dvsg_host_pipeline_destroy(dvsg_host_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (auto s : p->streams) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (auto q : p->d_in) cudaFree(q);
    for (auto q : p->d_out) cudaFree(q);
    for (auto q : p->d_vec) cudaFree(q);
    for (auto q : p->d_T) cudaFree(q);
    for (auto q : p->d_in8) cudaFree(q);
    for (auto q : p->d_out8) cudaFree(q);
    for (auto q : p->d_flow) cudaFree(q);
    cudaFree(p->d_coord);
    cudaFree(p->d_winv);
    delete p;
}

// Upload the mesh and invert its system -- only when it differs from the one the pipeline already holds (a clip's mesh is a
// constant, model.py:62-68): later calls skip the upload, the inversion and the synchronisation they need.
static int host_prepare_mesh(dvsg_host_pipeline* p, const float* coord_host) {
    const size_t n = (size_t)p->pn * 2;
    if (p->h_coord.size() == n && memcmp(p->h_coord.data(), coord_host, n * sizeof(float)) == 0) return DVSG_OK;
    for (auto st : p->streams) DVSG_CUDA(cudaStreamSynchronize(st));      // chunks in flight still read the old inverse
    DVSG_CUDA(cudaMemcpyAsync(p->d_coord, coord_host, n * sizeof(float), cudaMemcpyHostToDevice, p->streams[0]));
    const int rc0 = dvsg_tps_prepare(p->d_coord, 0, 1, p->pn, p->d_winv, p->winv_bytes, p->streams[0]);
    if (rc0) return rc0;
    DVSG_CUDA(cudaStreamSynchronize(p->streams[0]));
    p->h_coord.assign(coord_host, coord_host + n);
    return DVSG_OK;
}

static int host_finish(dvsg_host_pipeline* p) {
    if (p->async) return DVSG_OK;
    for (auto st : p->streams) DVSG_CUDA(cudaStreamSynchronize(st));
    return DVSG_OK;
}

extern "C" int dvsg_host_pipeline_set_async(dvsg_host_pipeline* p, int async) {
    DVSG_REQUIRE(p, "host_pipeline_set_async: null pipeline");
    p->async = async ? 1 : 0;
    return DVSG_OK;
}

extern "C" int dvsg_host_pipeline_sync(dvsg_host_pipeline* p) {
    DVSG_REQUIRE(p, "host_pipeline_sync: null pipeline");
    DVSG_CUDA(cudaSetDevice(p->device));
    for (auto st : p->streams) DVSG_CUDA(cudaStreamSynchronize(st));
    return DVSG_OK;
}

extern "C" int dvsg_host_tps_warp(dvsg_host_pipeline* p, const float* U_host, const float* coord_host,
                                  const float* vector_host, float* out_host, int B) {
    DVSG_REQUIRE(p && B >= 0, "host_tps_warp: bad argument");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(U_host && coord_host && vector_host && out_host, "host_tps_warp: null pointer");
    DVSG_CUDA(cudaSetDevice(p->device));
    const size_t fe = p->frame_elems;
    const int pn = p->pn;
    {   // the inverse of the mesh's system is computed when the mesh changes; every chunk only applies it
        const int rc0 = host_prepare_mesh(p, coord_host);
        if (rc0) return rc0;
    }
    int& chunk = p->next_chunk;
    for (int f0 = 0; f0 < B; f0 += p->fpc, ++chunk) {
        const int s = chunk % p->n_slots;
        const int nf = B - f0 < p->fpc ? B - f0 : p->fpc;
        cudaStream_t st = p->streams[s];   // stream order serialises reuse of slot s
        DVSG_CUDA(cudaMemcpyAsync(p->d_in[s], U_host + (size_t)f0 * fe, fe * nf * sizeof(float), cudaMemcpyHostToDevice, st));
        DVSG_CUDA(cudaMemcpyAsync(p->d_vec[s], vector_host + (size_t)f0 * pn * 2, (size_t)nf * pn * 2 * sizeof(float),
                                  cudaMemcpyHostToDevice, st));
        // coord + vector, the prepared solve and the fused warp: one launch for meshes up to 29 points, two otherwise
        const int rc = dvsg_tps_warp_frames_offsets(p->d_in[s], p->d_coord, p->d_vec[s], p->d_winv, p->winv_bytes, p->d_T[s], p->d_out[s], nullptr,
                                                    nullptr, nullptr, nf, p->H, p->W, p->C, p->H, p->W, pn, st);
        if (rc) return rc;
        DVSG_CUDA(cudaMemcpyAsync(out_host + (size_t)f0 * fe, p->d_out[s], fe * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    return host_finish(p);
}

// Same pipeline with uint8 frames on the host side: the ingest u/255 (+ optional BGR->RGB) and the
// egress uint8(x*255) (+ RGB->BGR) of eval.py:79-80,112-113 run on the device, so a frame crosses PCIe as 3 bytes per
// pixel each way instead of 12.  The warp itself is the unchanged fp32 path.
extern "C" int dvsg_host_tps_warp_u8(dvsg_host_pipeline* p, const unsigned char* U_host, const float* coord_host,
                                     const float* vector_host, unsigned char* out_host, int B, int swap_rb) {
    DVSG_REQUIRE(p && B >= 0, "host_tps_warp_u8: bad argument");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(U_host && coord_host && vector_host && out_host, "host_tps_warp_u8: null pointer");
    DVSG_REQUIRE(p->C == 3, "host_tps_warp_u8: uint8 frames are 3-channel (the pipeline was created with C = %d)", p->C);
    DVSG_CUDA(cudaSetDevice(p->device));
    const size_t fe = p->frame_elems;
    const int pn = p->pn;
    while ((int)p->d_in8.size() < p->n_slots) {
        unsigned char *a = nullptr, *b = nullptr;
        DVSG_CUDA(cudaMalloc(&a, fe * p->fpc));
        p->d_in8.push_back(a);
        DVSG_CUDA(cudaMalloc(&b, fe * p->fpc));
        p->d_out8.push_back(b);
    }
    {
        const int rc0 = host_prepare_mesh(p, coord_host);
        if (rc0) return rc0;
    }
    int& chunk = p->next_chunk;
    for (int f0 = 0; f0 < B; f0 += p->fpc, ++chunk) {
        const int s = chunk % p->n_slots;
        const int nf = B - f0 < p->fpc ? B - f0 : p->fpc;
        cudaStream_t st = p->streams[s];
        DVSG_CUDA(cudaMemcpyAsync(p->d_in8[s], U_host + (size_t)f0 * fe, fe * nf, cudaMemcpyHostToDevice, st));
        DVSG_CUDA(cudaMemcpyAsync(p->d_vec[s], vector_host + (size_t)f0 * pn * 2, (size_t)nf * pn * 2 * sizeof(float),
                                  cudaMemcpyHostToDevice, st));
        int rc = dvsg_frames_u8_to_f32(p->d_in8[s], p->d_in[s], (long long)nf * p->H * p->W, swap_rb, st);
        if (rc) return rc;
        rc = dvsg_tps_warp_frames_offsets(p->d_in[s], p->d_coord, p->d_vec[s], p->d_winv, p->winv_bytes, p->d_T[s], p->d_out[s], nullptr, nullptr,
                                          nullptr, nf, p->H, p->W, p->C, p->H, p->W, pn, st);
        if (rc) return rc;
        rc = dvsg_frames_f32_to_u8(p->d_out[s], p->d_out8[s], (long long)nf * p->H * p->W, swap_rb, st);
        if (rc) return rc;
        DVSG_CUDA(cudaMemcpyAsync(out_host + (size_t)f0 * fe, p->d_out8[s], fe * nf, cudaMemcpyDeviceToHost, st));
    }
    return host_finish(p);
}

// tf_warp(im, flow, H, W) on HOST buffers (warp_with_optical_flow.py:96-176 as trainer.py:246-247 / model.py:87-88 call it):
// frames and flow fields go up (12 + 8 bytes per pixel), warped frames come back (12).  Same chunked, multi-stream pipeline.
extern "C" int dvsg_host_flow_warp(dvsg_host_pipeline* p, const float* im_host, const float* flow_host, float* out_host, int B) {
    DVSG_REQUIRE(p && B >= 0, "host_flow_warp: bad argument");
    if (B == 0) return DVSG_OK;
    DVSG_REQUIRE(im_host && flow_host && out_host, "host_flow_warp: null pointer");
    DVSG_CUDA(cudaSetDevice(p->device));
    const size_t fe = p->frame_elems, ff = (size_t)p->H * p->W * 2;
    while ((int)p->d_flow.size() < p->n_slots) {
        float* a = nullptr;
        DVSG_CUDA(cudaMalloc(&a, ff * p->fpc * sizeof(float)));
        p->d_flow.push_back(a);
    }
    int& chunk = p->next_chunk;
    for (int f0 = 0; f0 < B; f0 += p->fpc, ++chunk) {
        const int s = chunk % p->n_slots;
        const int nf = B - f0 < p->fpc ? B - f0 : p->fpc;
        cudaStream_t st = p->streams[s];
        DVSG_CUDA(cudaMemcpyAsync(p->d_in[s], im_host + (size_t)f0 * fe, fe * nf * sizeof(float), cudaMemcpyHostToDevice, st));
        DVSG_CUDA(cudaMemcpyAsync(p->d_flow[s], flow_host + (size_t)f0 * ff, ff * nf * sizeof(float), cudaMemcpyHostToDevice, st));
        const int rc = dvsg_flow_warp_fwd(p->d_in[s], p->d_flow[s], p->d_out[s], nf, p->H, p->W, p->C, 0, st);
        if (rc) return rc;
        DVSG_CUDA(cudaMemcpyAsync(out_host + (size_t)f0 * fe, p->d_out[s], fe * nf * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    return host_finish(p);
}
