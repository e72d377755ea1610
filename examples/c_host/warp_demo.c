/* A C host driving the frame-warping path through the C ABI alone (no Python, no torch): what a maintainer of a compiled
 * caller would write.  Solves the TPS coefficients of a 4x4 mesh for B frames, warps the frames (ThinPlateSpline.py:4-170
 * behind dvsg_tps_solve + dvsg_tps_warp_fwd), runs the backward, and prints checksums that tests/test_gpu_c_host.py compares
 * with the Python drop-in on the same inputs.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_host/warp_demo.c -o warp_demo \
 *       -L coupe/dvsg_b200 -ldvsg_warp -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/coupe/dvsg_b200
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "dvsg_warp.h"

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
    } while (0)
#define DV(call)                                                                         \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != DVSG_OK) { fprintf(stderr, "%s: %d %s\n", #call, rc_, dvsg_last_error()); return 3; } \
    } while (0)

/* the same deterministic inputs as the Python side of the test: a 32-bit LCG mapped to [0, 1) */
static unsigned lcg_state = 12345u;
static float lcg(void) { lcg_state = lcg_state * 1664525u + 1013904223u; return (float)(lcg_state >> 8) * (1.0f / 16777216.0f); }

static double checksum(const float* v, size_t n) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += (double)v[i] * (double)(1 + (i % 7));
    return s;
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 2, H = 288, W = 512, C = 3, m = 4, pn = m * m, N = pn + 3;
    const size_t n_px = (size_t)B * H * W, n_im = n_px * C;
    float* h_im = (float*)malloc(n_im * sizeof(float));
    float* h_mesh = (float*)malloc((size_t)pn * 2 * sizeof(float));
    float* h_tgt = (float*)malloc((size_t)B * pn * 2 * sizeof(float));
    float* h_out = (float*)malloc(n_im * sizeof(float));
    float* h_gT = (float*)malloc((size_t)B * 2 * N * sizeof(float));
    for (size_t i = 0; i < n_im; ++i) h_im[i] = lcg();
    for (int k = 0; k < pn; ++k) {        /* regular mesh on [-1, 1]^2, x fastest (model.py:62-68) */
        h_mesh[2 * k] = -1.0f + 2.0f * (float)(k % m) / (float)(m - 1);
        h_mesh[2 * k + 1] = -1.0f + 2.0f * (float)(k / m) / (float)(m - 1);
    }
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < 2 * pn; ++k) h_tgt[(size_t)b * 2 * pn + k] = h_mesh[k] + (lcg() - 0.5f) * 0.2f;

    float *d_im, *d_mesh, *d_tgt, *d_T, *d_out, *d_x, *d_y, *d_gout, *d_gU, *d_gT;
    CK(cudaMalloc((void**)&d_im, n_im * sizeof(float)));
    CK(cudaMalloc((void**)&d_out, n_im * sizeof(float)));
    CK(cudaMalloc((void**)&d_gout, n_im * sizeof(float)));
    CK(cudaMalloc((void**)&d_gU, n_im * sizeof(float)));
    CK(cudaMalloc((void**)&d_x, n_px * sizeof(float)));
    CK(cudaMalloc((void**)&d_y, n_px * sizeof(float)));
    CK(cudaMalloc((void**)&d_mesh, (size_t)pn * 2 * sizeof(float)));
    CK(cudaMalloc((void**)&d_tgt, (size_t)B * pn * 2 * sizeof(float)));
    CK(cudaMalloc((void**)&d_T, (size_t)B * 2 * N * sizeof(float)));
    CK(cudaMalloc((void**)&d_gT, (size_t)B * 2 * N * sizeof(float)));
    CK(cudaMemcpy(d_im, h_im, n_im * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_mesh, h_mesh, (size_t)pn * 2 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tgt, h_tgt, (size_t)B * pn * 2 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gout, h_im, n_im * sizeof(float), cudaMemcpyHostToDevice));      /* upstream gradient = the frames themselves */
    CK(cudaMemset(d_gU, 0, n_im * sizeof(float)));

    /* one mesh shared by the batch: coord_batch_stride = 0 */
    size_t ws_bytes = dvsg_tps_solve_workspace_bytes(B, pn, 0);
    void* d_ws = NULL;
    if (ws_bytes) CK(cudaMalloc(&d_ws, ws_bytes));
    DV(dvsg_tps_solve(d_mesh, 0, d_tgt, d_T, B, pn, d_ws, ws_bytes, NULL));
    DV(dvsg_tps_warp_fwd(d_im, d_mesh, 0, d_T, d_out, d_x, d_y, NULL, B, H, W, C, H, W, pn, 0, NULL));
    DV(dvsg_tps_warp_bwd(d_im, d_mesh, 0, d_T, d_gout, NULL, NULL, d_gU, d_gT, NULL, NULL, B, H, W, C, H, W, pn, NULL));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, n_im * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_gT, d_gT, (size_t)B * 2 * N * sizeof(float), cudaMemcpyDeviceToHost));
    printf("out_checksum %.9e\n", checksum(h_out, n_im));
    printf("gradT_checksum %.9e\n", checksum(h_gT, (size_t)B * 2 * N));
    printf("launches %lld\n", dvsg_launch_count());
    return 0;
}
