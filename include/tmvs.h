/*
 * tmvs.h -- C ABI of libtmvs_sm100a.so: B200 (sm_100a) kernels for the TransMVSNet
 * cost-volume hot path.
 *
 * The reference has NO plugin / operator / FFI boundary for this path: the "interface" is two
 * plain Python functions bound by a star import (models/TransMVSNet.py:4) and the statements
 * of DepthNet.forward between them.  Each entry point below names the reference lines it
 * replaces; INTEGRATION.md shows the ctypes stub that binds it from the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless it says "host"; the library never
 *     allocates, frees or retains memory (PyTorch owns all buffers) and never synchronises;
 *   - launches go to `stream` (a cudaStream_t / CUstream passed as void*), current device;
 *   - returns 0 on success, a negative TMVS_E_* code for argument errors, or a positive
 *     cudaError_t if a launch failed; tmvs_error_string() describes either;
 *   - tensors are contiguous in the layouts written beside them; feature inputs additionally
 *     take element strides so NCHW and torch.channels_last tensors are both accepted;
 *   - "packed" source features use the kernel-native blocked channel-last layout
 *       [Nsrc][B][H][Wb][C4][8 px][4 ch],  C4 = ceil(C/4) (zero padded), Wb = ceil(W/8),
 *     produced by tmvs_pack_sources(): one bilinear tap of 4 channels is one 128-bit load, 8
 *     x-adjacent pixels share a 128-byte line and the C4 groups of a pixel are 128 bytes apart;
 *   - rot_trans is an array [Nsrc][B][12]: 3x3 `rot` (row major) then `trans` of
 *       proj = src_proj @ inverse(ref_proj)            (models/module.py:295-297),
 *     computed by the caller with the same torch ops as the reference; on the HOST by default (passed to the
 *     kernels by value) or, with TMVS_F_RT_DEVICE, on the device (read in place: no copy, no synchronisation);
 *   - depth hypotheses are [B][D] (per_pixel = 0) or [B][D][H][W] (per_pixel = 1), the two
 *     shapes models/module.py:288,306 accepts.
 */
#ifndef TMVS_H_
#define TMVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMVS_VERSION 200          /* 0.2.0: per-call flags, no process-wide state */
#define TMVS_MAX_SRC_VIEWS 16     /* source views per launch */
#define TMVS_MAX_DEPTH 256        /* depth planes per pixel */

enum {
    TMVS_OK = 0,
    TMVS_E_NULL = -1,        /* a required pointer is NULL */
    TMVS_E_SHAPE = -2,       /* a dimension is <= 0 or above a TMVS_MAX_* limit */
    TMVS_E_ALIGN = -3,       /* a packed / vector pointer is not 16-byte aligned */
    TMVS_E_UNSUPPORTED = -4  /* combination not implemented */
};

typedef void *tmvs_stream_t;

/*
 * Per-call options (`flags` argument).  The library keeps NO process-wide state and reads NO environment variable:
 * every choice below travels with the call, so concurrent callers with different options cannot interfere.
 *
 * Arithmetic.  The reference's own CPU and CUDA executions of models/module.py:311-313 differ in the last bit: ATen's
 * CPU kernel divides by the python scalar (W-1)/2, ATen's CUDA kernel multiplies by its reciprocal.  At 1152x1600 that
 * moves sample positions by ~1e-4 px and cost volumes by ~2e-4 (max-norm).  Without TMVS_F_ARITH_ATEN_CUDA the kernels
 * follow the CPU arithmetic (what the CPU-generated golden vectors pin); with it they follow the CUDA one -- the
 * arithmetic of the device the reference would have run on, and the default of the Python drop-in for CUDA tensors.
 */
#define TMVS_ARITH_IEEE 0
#define TMVS_ARITH_ATEN_CUDA 1
#define TMVS_F_ARITH_ATEN_CUDA 0x1u   /* follow ATen's CUDA arithmetic (reciprocal multiply) instead of the CPU one */
#define TMVS_F_RT_DEVICE       0x2u   /* rot_trans is a DEVICE array (same [Nsrc][B][12] layout), read by the kernels in
                                         place: nothing is copied to the host and nothing synchronises */
#define TMVS_F_FWD_TMA         0x4u   /* tmvs_costvol_fwd: TMA-staged shared-memory gather (tmvs_costvol_tma.cu) where
                                         it applies; same results up to fp32 re-association; measured slower */
#define TMVS_F_BWD_SCAN        0x8u   /* tmvs_costvol_bwd: grad_src through the tile-scan kernels only (no cell tables) */
#define TMVS_F_PACK_LDG        0x10u  /* tmvs_pack_sources: register-transpose kernel instead of the TMA engine */
#define TMVS_F_FWD_SPLIT       0x20u  /* tmvs_costvol_fwd, C = 32: two channel passes of 16 (64 registers, 4 CTAs per SM) instead of
                                         one pass over all 32 (128 registers, 2 CTAs per SM); measured slower */
#define TMVS_F_FWD_SWEEP       0x40u  /* tmvs_costvol_fwd (aggregated output, C = 8 or 16): epipolar-sweep kernel
                                         (tmvs_costvol_sweep.cu) -- bit-identical results, fewer loads where the
                                         hypotheses of a pixel are less than a pixel apart in the source image */
#define TMVS_F_RAY_UNFUSED     0x80u  /* rot @ (x, y, 1) of models/module.py:305 evaluated as ((r0*x) + (r1*y)) + r2 instead of
                                         fma(r2, 1, fma(r1, y, r0*x)): what cuBLAS does for the small stage-1 maps (probed) */
#define TMVS_F_TABLE_MB(mb)    ((unsigned)(mb) << 16)   /* tmvs_costvol_bwd(+_workspace_bytes): cap of the cell-table
                                         workspace in MiB (0 = default 3072), e.g. to exercise the multi-pass path */

int tmvs_version(void);
/*
 * Peer-mapped gather buffer: the multi-GPU exchange of SURVEY.md 8e (every view's depth + confidence map ends up on
 * one rank) without a collective.  The gathering rank creates the buffer and ships the 64-byte CUDA IPC handle to the
 * other ranks of the box (any host channel); each opens it with its own GPU current, which maps the buffer into that
 * GPU's address space over NVLink / NVSwitch.  The pointer is then an ordinary OUTPUT pointer of the entry points
 * below -- e.g. depth / conf of tmvs_softmax_wta_fwd -- so the kernel's stores cross the link and nothing else runs.
 * The only calls of this library that allocate or free device memory (cudaMalloc / cudaFree / IPC open / close).
 */
int tmvs_peer_buffer_create(size_t bytes, void **ptr, unsigned char *handle64);   /* zero-filled, current device */
int tmvs_peer_buffer_open(const unsigned char *handle64, void **ptr);             /* in another process of the box */
int tmvs_peer_buffer_release(void *ptr, int owner);                               /* owner: cudaFree, else IPC close */
/* Asynchronous copy (DMA engines, not SMs) from local device memory into a buffer opened with tmvs_peer_buffer_open:
 * the transport for the maps when the producing kernel should not wait on the NVLink port. */
int tmvs_peer_copy_async(void *dst, const void *src, size_t bytes, tmvs_stream_t stream);
const char *tmvs_error_string(int code);

/* Bytes of the packed source workspace for the given shape. */
size_t tmvs_packed_bytes(int n_src, int B, int C, int H, int W);

/*
 * Layout pre-pass: the N source feature maps  ->  packed [Nsrc][B][H][Wb][C4][8][4].
 * src[i] points at view i's [B,C,H,W] tensor with element strides (sB,sC,sH,sW), so the
 * NCHW output of the reference's FeatureNet/FMT (models/module.py:399-422, models/FMT.py:212-230)
 * and channels_last tensors are both read in place.   src is a HOST array of device pointers.
 */
int tmvs_pack_sources(const float *const *src, int n_src, int64_t sB, int64_t sC, int64_t sH, int64_t sW,
                      float *packed, int B, int C, int H, int W, unsigned flags, tmvs_stream_t stream);

/*
 * Drop-in homo_warping (models/module.py:284-322) for ONE source view: materialises the
 * warped volume out[B][C][D][H][W].  packed_view = this view's slice [B][H][Wb][C4][8][4];
 * rot_trans = host [B][12].
 */
int tmvs_homo_warp_fwd(const float *packed_view, const float *rot_trans, const float *depth, int per_pixel,
                       float *out, int B, int C, int D, int H, int W, unsigned flags, tmvs_stream_t stream);

/*
 * Backward of the drop-in homo_warping wrt the source features (autograd of F.grid_sample, models/module.py:318-320,
 * for an arbitrary upstream gradient): grad_out [B][C][D][H][W] -> grad_src [B][C][H][W] (contiguous NCHW, overwritten).
 * Deterministic and free of floating-point atomics: every source pixel gathers its contributions through the
 * cell table of tmvs_costvol_bwd.  workspace >= tmvs_homo_warp_bwd_workspace_bytes().
 */
int tmvs_homo_warp_bwd(const float *rot_trans, const float *depth, int per_pixel, const float *grad_out,
                       float *grad_src, void *workspace, size_t workspace_bytes, int B, int C, int D, int H, int W,
                       unsigned flags, tmvs_stream_t stream);
size_t tmvs_homo_warp_bwd_workspace_bytes(int B, int C, int D, int H, int W);

/*
 * Fused warp + bilinear sampling + correlation (+ view-weighted aggregation):
 * replaces the whole view loop models/TransMVSNet.py:71-93 (homo_warping :79,
 * (warped*ref).mean(1) :80, similarity_sum/weight_sum :88-93) without ever writing the
 * B x C x D x H x W warped volume.
 *   ref            reference features [B,C,H,W] with element strides (rB,rC,rH,rW)
 *   packed         packed sources [Nsrc][B][H][Wb][C4][8][4]
 *   view_weights   [B][Nsrc][H][W] or NULL
 *   sim_views      [Nsrc][B][D][H][W] per-view similarity (stage 1, feeds PixelwiseNet) or NULL
 *   agg            [B][D][H][W] = sum_i sim_i*w_i / (1e-5 + sum_i w_i)  or NULL (needs view_weights)
 */
int tmvs_costvol_fwd(const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW,
                     const float *packed, const float *rot_trans, const float *depth, int per_pixel,
                     const float *view_weights, float *sim_views, float *agg,
                     int B, int C, int D, int H, int W, int n_src, unsigned flags, tmvs_stream_t stream);

/*
 * The same kernel fed from per-view packed maps that live anywhere on the device -- what a scan-level cache of
 * packed feature pyramids hands over (each view of a scan is a source view of several reference views; it is
 * packed once, datasets/general_eval.py:25-57 pairing) -- and with the view weights read at a coarser stage's
 * resolution:
 *   packed_views   HOST array [n_src] of device pointers, each one view's packed map [B][H][Wb][C4][8][4]
 *   view_weights   [B][Nsrc][vw_h][vw_w]; the kernel reads w[y >> vw_shift][x >> vw_shift], i.e. the nearest x2
 *                  upsampling of models/TransMVSNet.py:193-194 applied vw_shift times, without materialising it
 *                  (vw_shift = 0: full resolution, vw_h = H, vw_w = W).
 */
int tmvs_costvol_fwd_cached(const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW,
                            const float *const *packed_views, const float *rot_trans, const float *depth,
                            int per_pixel, const float *view_weights, int vw_shift, int vw_h, int vw_w,
                            float *sim_views, float *agg, int B, int C, int D, int H, int W, int n_src,
                            unsigned flags, tmvs_stream_t stream);

/*
 * Aggregation alone (stage 1 after PixelwiseNet, models/TransMVSNet.py:71-72,88-93):
 * agg[B][D][H][W] = sum_i sim_views[i]*w[:,i] / (1e-5 + sum_i w[:,i]).
 */
int tmvs_aggregate_fwd(const float *sim_views, const float *view_weights, float *agg,
                       int B, int D, int H, int W, int n_src, tmvs_stream_t stream);

/*
 * SURVEY.md 8(f) N1 -- depth hypotheses of a cascade stage in one kernel, at the stage resolution:
 * replaces models/TransMVSNet.py:174-190 (bilinear upsample of the previous depth to the image size +
 * get_depth_samples, models/module.py:606-634) and :202-204 (trilinear resample to [D, h, w]) without
 * materialising the [B, D, Himg, Wimg] volume.
 *   prev_depth   [B][hp][wp] previous-stage depth (prev_planes = 0), or [B][n_planes] depth_values (prev_planes = n)
 *   out          [B][D][h][w], h = Himg / scale, w = Wimg / scale, scale in {1, 2, 4}
 *   interval     depth_inteval_pixel of the stage (ratio * depth_interval), used when prev_planes = 0
 */
int tmvs_depth_hypotheses_fwd(const float *prev_depth, int prev_planes, int hp, int wp, float interval, float *out,
                              int B, int D, int h, int w, int scale, tmvs_stream_t stream);

/*
 * SURVEY.md 8(f) N3 -- read-out -> wire format on the device (test.py:119-158, utils.py:11-21): the stage-1/2
 * confidence maps are bilinearly resized to the stage-3 size (cv2.resize INTER_LINEAR semantics), multiplied into
 * the final confidence, depth is zeroed where confidence < threshold, and the 8-bit alpha channel the PNG carries
 * (depth_normal: clamp to [depth_min, depth_max], scale to 0..255, truncate) is produced -- so only two float maps
 * and one byte map cross PCIe instead of every output including prob_volume (utils.py:66-73 tensor2numpy).
 *   depth, conf3 [B][H][W]; conf1 [B][h1][w1]; conf2 [B][h2][w2]
 *   depth_out, conf_out [B][H][W] fp32; alpha_out [B][H][W] uint8   (any output may be NULL)
 */
int tmvs_finalize_maps_fwd(const float *depth, const float *conf3, const float *conf1, int h1, int w1,
                           const float *conf2, int h2, int w2, float conf_threshold, float depth_min,
                           float depth_max, float *depth_out, float *conf_out, uint8_t *alpha_out, int B, int H,
                           int W, tmvs_stream_t stream);

/*
 * SURVEY.md 8(f) N2 -- eval-mode PixelwiseNet folded into the aggregation (models/TransMVSNet.py:10-30, 82-93):
 * for every source view   w_i = max_d sigmoid(MLP(sim_i[d])),  MLP = 1->16->8->1 per-voxel (1x1x1 Conv3d with the
 * BatchNorm3d running statistics folded in, ReLU),  then  agg = sum_i sim_i*w_i / (1e-5 + sum_i w_i).
 * mlp = HOST array of TMVS_PWN_PARAMS floats: w0[16], b0[16], w1[8][16], b1[8], w2[8], b2.
 * Outputs: view_weights [B][Nsrc][H][W] and agg [B][D][H][W].  Inference only (training needs batch statistics).
 */
#define TMVS_PWN_PARAMS 177
int tmvs_pixelwise_aggregate_fwd(const float *sim_views, const float *mlp, float *view_weights, float *agg,
                                 int B, int D, int H, int W, int n_src, tmvs_stream_t stream);

/*
 * SURVEY.md 8(f) N4 -- fusibile depth-map fusion (gipuma/fusibile/fusibile.cu:89-173 kernel `fusibile`, :175-210
 * copy_pc_to_host, :216-285 the per-camera launch / synchronise / host-scan loop; main.cpp:128-147 image set-up).
 *   images  [V][H][W][4] fp32: b, g, r in [0,1] and w = depth (425 + 512 * alpha/255, main.cpp:137); sampled through
 *           the texture unit with the reference's settings (float4 texels, bilinear, unnormalised + 0.5: main.cpp:30-66).
 *           16-byte aligned.  If the buffer is 512-byte aligned, W even and H*W a multiple of 32, the textures are
 *           built over it in place (pitch-linear resources); otherwise, or with TMVS_FUSE_ARRAY_TEXTURES, each view is
 *           copied into a cudaArray this call allocates and frees, as the reference does.  Identical samples either way.
 *   cams    HOST [V][TMVS_FUSE_CAM_FLOATS]: P (3x4 row major), RK_inv = inverse(P[:, :3]) (3x3), camera centre C (3),
 *           P[:, 3] (3), focal length K[0] of the decomposed P (cameraGeometryUtils.h:104-156).
 *   depth_threshold 0.25, consistent_threshold 3 (algorithmparameters.h:11-12).
 *   carry_over bit 0 set reproduces the reference's output exactly: its per-pixel point buffer is never cleared between
 *           cameras, so every later camera re-emits a pixel's latest fused point (fusibile.cu:165-166,188);
 *           clear: each camera's own points only.  Bit 1 (TMVS_FUSE_ARRAY_TEXTURES): force the cudaArray textures.
 *           Bit 2 (TMVS_FUSE_IEEE): IEEE division and square root.
 *           The default arithmetic is the reference's AS ITS OWN BUILD COMPILES IT (CMakeLists.txt:10 --use_fast_math:
 *           reciprocal-multiply divisions, MUFU.SQRT, the compiler's FMA contraction order), read off the SASS of
 *           gipuma/fusibile/fusibile.cu; TMVS_FUSE_IEEE is the same source without --use_fast_math.  Both are pinned
 *           bit for bit against the reference's compiled kernel (oracle/_ref, tests/test_gpu_fusion.py).
 *   points  [capacity][8] fp32 out: x, y, z, 0, b, g, r, 0 (point_cloud.h:7-11; the reference's float4 operator+ drops w),
 *           in the reference's order (camera, then y, then x).  *n_points (device, int64) = number of points found,
 *           which may exceed capacity (only the first `capacity` are written).
 * Unlike every other entry point this one synchronises `stream` before returning (it owns texture objects).
 */
#define TMVS_FUSE_CAM_FLOATS 28
#define TMVS_FUSE_MAX_VIEWS 1024       /* config.h:2 MAX_IMAGES */
#define TMVS_FUSE_ARRAY_TEXTURES 2
#define TMVS_FUSE_IEEE 4
int tmvs_fusibile_fwd(const float *images, const float *cams, int V, int H, int W, float depth_threshold,
                      int consistent_threshold, int carry_over, float *points, long long capacity,
                      long long *n_points, void *workspace, size_t workspace_bytes, tmvs_stream_t stream);
size_t tmvs_fusibile_workspace_bytes(int V, int H, int W);
/* Diagnostic: out[j] = tex2D<float4>(image, uv[j]) through the texture set-up tmvs_fusibile_fwd uses (image [H][W][4],
 * uv [n][2] unnormalised texture coordinates, out [n][4], all device).  The tests use it to measure the CPU oracle's
 * emulation of the hardware's 9-bit-weight bilinear filter.  Synchronises `stream`. */
int tmvs_fusibile_tex_probe(const float *image, int H, int W, const float *uv, float *out, int n, int mode,
                            tmvs_stream_t stream);      /* mode: 0 = pitch-linear texture, TMVS_FUSE_ARRAY_TEXTURES */

/*
 * Backward of the cost volume wrt the features (autograd of models/module.py:318-320 and
 * models/TransMVSNet.py:80).  grad_views = dL/d sim_i [Nsrc][B][D][H][W].
 *   grad_ref  [B][C][H][W]            (contiguous NCHW, overwritten)
 *   grad_src  [Nsrc][B][C][H][W]      (contiguous NCHW, overwritten)
 * Deterministic and free of floating-point atomics.  Either output may be NULL.
 */
int tmvs_costvol_bwd(const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW,
                     const float *packed, const float *rot_trans, const float *depth, int per_pixel,
                     const float *grad_views, float *grad_ref, float *grad_src, void *workspace,
                     size_t workspace_bytes, int B, int C, int D, int H, int W, int n_src,
                     unsigned flags, tmvs_stream_t stream);
size_t tmvs_costvol_bwd_workspace_bytes(int B, int C, int D, int H, int W, int n_src, unsigned flags);

/*
 * Read-out (models/TransMVSNet.py:99-103 + models/module.py:474-482) in one pass:
 *   prob = exp(log_softmax(logits, dim=1)); index = argmax_d prob (first maximal, int64);
 *   depth = depth_values[index]; conf = max_d prob.      prob may be NULL (not materialised).
 */
int tmvs_softmax_wta_fwd(const float *logits, const float *depth_values, float *prob, int64_t *index,
                         float *depth, float *conf, int B, int D, int H, int W, tmvs_stream_t stream);

/* depth_wta(p, depth_values) (models/module.py:474-482): index (int64, may be NULL) + depth. */
int tmvs_depth_wta(const float *p, const float *depth_values, int64_t *index, float *depth,
                   int B, int D, int H, int W, tmvs_stream_t stream);

/*
 * depth_regression(p, depth_values) = sum_d p*depth_values.  ABSENT from this fork of the
 * reference (north_star signature, upstream MVSNet definition).
 * bwd: grad_p[B][D][H][W] = grad_depth[B][H][W] * depth_values.
 */
int tmvs_depth_regression_fwd(const float *p, const float *depth_values, int per_pixel, float *depth,
                              int B, int D, int H, int W, tmvs_stream_t stream);
int tmvs_depth_regression_bwd(const float *grad_depth, const float *depth_values, int per_pixel,
                              float *grad_p, int B, int D, int H, int W, tmvs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TMVS_H_ */
