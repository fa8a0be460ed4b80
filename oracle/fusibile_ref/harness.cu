// Harness around the REFERENCE's own fusion kernel (gipuma/fusibile/fusibile.cu, compiled from where it lies under
// /root/reference by oracle/build.py into oracle/_ref/libfusibile_ref.so; nothing of it is copied into the repo).
//
// TEST INFRASTRUCTURE ONLY: tests/test_gpu_fusion.py loads the library to pin tmvs_fusibile_fwd
// against the reference's kernel `fusibile` (:89-173), its host compaction copy_pc_to_host (:175-210) and its
// per-camera launch loop fusibile_cu (:216-285) -- all executed unmodified through run_cuda().  What the harness itself
// restates is only the set-up main.cpp does with OpenCV (which is not available to build against):
//   * the texture objects of main.cpp:30-66 add_images_to_texture (float4 cudaArray, linear filter, element read mode,
//     unnormalised coordinates, wrap address mode) -- same calls, same descriptors;
//   * the Camera_cu fields that cameraGeometryUtils.h:104-156 fills from each 3x4 projection matrix: they arrive
//     precomputed from the caller (transmvsnet_b200.fusion.camera_record), [V][28] = P(12) RK_inv(9) C(3) P[:,3](3) K00.
#include <stdio.h>
#include <string.h>

#include "fusibile.h"          // /root/reference/gipuma/fusibile (include path)

extern "C" int fusibile_ref_run(const float *images, const float *cams, int V, int H, int W, float depth_threshold,
                                int consistent_threshold, float *points, long long capacity, long long *n_points)
{
    if (!images || !cams || !points || !n_points || V <= 0 || V > MAX_IMAGES) return -1;
    GlobalState *gs = new GlobalState;
    gs->algorithm = new AlgorithmParameters;
    gs->algorithm->depth_threshold = depth_threshold;
    gs->algorithm->consistent_threshold = consistent_threshold;
    gs->cameras = new CameraParameters_cu;
    gs->cameras->n_cameras = V;
    gs->cameras->cols = W;
    gs->cameras->rows = H;
    for (int v = 0; v < V; ++v) {
        const float *c = cams + (size_t)v * 28;
        Camera_cu &cam = gs->cameras->cameras[v];
        memset(cam.P, 0, sizeof(float) * 16);
        memset(cam.K, 0, sizeof(float) * 16);
        memset(cam.R, 0, sizeof(float) * 16);
        memset(cam.RK_inv, 0, sizeof(float) * 16);
        memcpy(cam.P, c, sizeof(float) * 12);
        memcpy(cam.RK_inv, c + 12, sizeof(float) * 9);
        cam.C4 = make_float4(c[21], c[22], c[23], 0.f);
        cam.P_col34 = make_float4(c[24], c[25], c[26], 0.f);
        cam.K[0] = c[27];
    }
    gs->pc = new PointCloud;
    gs->pc->resize(H * W);
    PointCloudList pc_list;
    pc_list.resize(H * W);
    pc_list.size = 0;
    cudaArray_t *arrays = new cudaArray_t[V];
    for (int v = 0; v < V; ++v) {            // main.cpp:30-66
        cudaChannelFormatDesc channelDesc = cudaCreateChannelDesc<float4>();
        if (cudaMallocArray(&arrays[v], &channelDesc, W, H) != cudaSuccess) return -2;
        if (cudaMemcpy2DToArray(arrays[v], 0, 0, images + (size_t)v * H * W * 4, (size_t)W * 16, (size_t)W * 16, H,
                                cudaMemcpyHostToDevice) != cudaSuccess) return -3;
        struct cudaResourceDesc resDesc;
        memset(&resDesc, 0, sizeof(resDesc));
        resDesc.resType = cudaResourceTypeArray;
        resDesc.res.array.array = arrays[v];
        struct cudaTextureDesc texDesc;
        memset(&texDesc, 0, sizeof(texDesc));
        texDesc.addressMode[0] = cudaAddressModeWrap;
        texDesc.addressMode[1] = cudaAddressModeWrap;
        texDesc.filterMode = cudaFilterModeLinear;
        texDesc.readMode = cudaReadModeElementType;
        texDesc.normalizedCoords = 0;
        if (cudaCreateTextureObject(&gs->color_images_textures[v], &resDesc, &texDesc, NULL) != cudaSuccess) return -4;
    }
    run_cuda(*gs, pc_list, V);               // the reference's loop: launch, synchronise, host scan, per camera
    cudaError_t err = cudaDeviceSynchronize();
    *n_points = (long long)pc_list.size;
    const long long n = pc_list.size < capacity ? pc_list.size : capacity;
    for (long long i = 0; i < n; ++i) {
        const Point_li &p = pc_list.points[i];
        float *o = points + i * 8;
        o[0] = p.coord.x; o[1] = p.coord.y; o[2] = p.coord.z; o[3] = p.coord.w;
        o[4] = p.texture4[0]; o[5] = p.texture4[1]; o[6] = p.texture4[2]; o[7] = p.texture4[3];
    }
    for (int v = 0; v < V; ++v) {
        cudaDestroyTextureObject(gs->color_images_textures[v]);
        cudaFreeArray(arrays[v]);
    }
    delete[] arrays;
    delete gs->pc;
    delete gs->cameras;
    delete gs->algorithm;
    delete gs;
    return err == cudaSuccess ? 0 : (int)err;
}
