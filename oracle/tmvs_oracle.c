/*
 * tmvs_oracle.c -- CPU ORACLE for the TransMVSNet cost-volume hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (transmvsnet_b200/) never calls it and has
 * no CPU fallback.
 *
 * It is a plain-C, scalar (OpenMP over rows) restatement of the reference's
 * algorithm.  Nothing here is copied from the reference; each function cites the
 * reference lines whose arithmetic it restates (paths relative to the reference
 * tree) and the ATen semantics the reference relies on.
 *
 * Parity status: PINNED against golden vectors produced by importing the real
 * reference functions (tests/golden/make_golden.py, run in the build container
 * where /root/reference exists).  The one exception is depth_regression, which
 * does not exist in this fork of the reference: it
 * restates the upstream 3-line definition and is "parity unpinned" by the
 * reference itself (it is pinned only against a torch expression of that
 * definition).
 *
 * Build:  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC   (see oracle/build.py)
 * -ffp-contract=off matters: the reference runs separate mul/add ATen kernels,
 * so every intermediate is rounded to fp32; only the 3-term dot product of the
 * rot @ xyz matmul uses fmaf, in the order sgemm does (see oracle_source_index).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define TMVS_ORACLE_VERSION 100

int tmvs_oracle_version(void) { return TMVS_ORACLE_VERSION; }

/* ------------------------------------------------------------------------- */
/* Sampling coordinates: models/module.py:295-315 + ATen grid_sampler         */
/* un-normalisation (ATen/native/GridSampler.h:27-31, align_corners=True)     */
/* and the int-range guard of ATen/native/cuda/GridSampler.cuh:140-147.       */
/* rt = 12 floats: rot row-major (9) then trans (3), where                    */
/*   proj = src_proj @ inverse(ref_proj); rot = proj[:3,:3]; trans=proj[:3,3] */
/* ------------------------------------------------------------------------- */
static inline void oracle_source_index(const float *rt, float x, float y, float depth,
                                       int H, int W, float *ix, float *iy)
{
    /* module.py:305  rot_xyz = rot @ (x, y, 1).  sgemm (MKL and cuBLAS alike, probed bit for bit by
       scripts/probe_matmul.py) accumulates in k order with FMAs: r0*x, fma(r1, y, .), fma(r2, 1, .) */
    float rx = fmaf(rt[1], y, rt[0] * x) + rt[2];
    float ry = fmaf(rt[4], y, rt[3] * x) + rt[5];
    float rz = fmaf(rt[7], y, rt[6] * x) + rt[8];
    /* module.py:306-308  rot_depth_xyz = rot_xyz * depth ; proj_xyz = . + trans */
    float px = rx * depth; px = px + rt[9];
    float py = ry * depth; py = py + rt[10];
    float pz = rz * depth; pz = pz + rt[11];
    /* module.py:309  invalid = z < 1e-6 */
    int invalid = pz < 1e-6f;
    /* module.py:310  xy / z */
    float qx = px / pz;
    float qy = py / pz;
    /* module.py:311-314 normalise to [-1,1]; invalid -> -99 */
    float nx = qx / ((float)(W - 1) / 2.0f) - 1.0f;
    float ny = qy / ((float)(H - 1) / 2.0f) - 1.0f;
    if (invalid) { nx = -99.0f; ny = -99.0f; }
    /* ATen grid_sampler_unnormalize, align_corners=True */
    float fx = ((nx + 1.0f) / 2.0f) * (float)(W - 1);
    float fy = ((ny + 1.0f) / 2.0f) * (float)(H - 1);
    /* ATen safe_downgrade_to_int_range: non-finite / out of int range -> -100 */
    /* (guard tightened to |coord| < 2^31-128 so the int cast below is defined; such
       coordinates are far out of bounds either way, the sample is 0) */
    if (!(fx < 2147483520.0f && fx > -2147483520.0f)) fx = -100.0f;
    if (!(fy < 2147483520.0f && fy > -2147483520.0f)) fy = -100.0f;
    *ix = fx;
    *iy = fy;
}

/* Bilinear footprint per ATen grid_sampler_2d (cuda/GridSampler.cuh bilinear   */
/* branch): corners nw/ne/sw/se, weights from the (corner - coord) differences, */
/* zero padding applied PER TAP (within_bounds_2d).                             */
typedef struct {
    int x0, y0;          /* north-west corner */
    float w[4];          /* nw, ne, sw, se */
    int ok[4];           /* tap in bounds? */
} oracle_taps;

static inline void oracle_footprint(float ix, float iy, int H, int W, oracle_taps *t)
{
    float fx0 = floorf(ix), fy0 = floorf(iy);
    int x0 = (int)fx0, y0 = (int)fy0;
    int x1 = x0 + 1, y1 = y0 + 1;
    t->x0 = x0; t->y0 = y0;
    t->w[0] = ((float)x1 - ix) * ((float)y1 - iy);
    t->w[1] = (ix - (float)x0) * ((float)y1 - iy);
    t->w[2] = ((float)x1 - ix) * (iy - (float)y0);
    t->w[3] = (ix - (float)x0) * (iy - (float)y0);
    t->ok[0] = (y0 >= 0 && y0 < H && x0 >= 0 && x0 < W);
    t->ok[1] = (y0 >= 0 && y0 < H && x1 >= 0 && x1 < W);
    t->ok[2] = (y1 >= 0 && y1 < H && x0 >= 0 && x0 < W);
    t->ok[3] = (y1 >= 0 && y1 < H && x1 >= 0 && x1 < W);
}

static inline float oracle_depth_at(const float *depth, int per_pixel, int b, int d, int y, int x,
                                    int D, int H, int W)
{
    /* module.py:306  depth_values.view(B,1,D,-1): [B,D] broadcasts over pixels */
    return per_pixel ? depth[(((size_t)b * D + d) * H + y) * W + x] : depth[(size_t)b * D + d];
}

/* ------------------------------------------------------------------------- */
/* homo_warping: models/module.py:284-322.                                    */
/* src [B,C,H,W], rt [B,12], depth [B,D] or [B,D,H,W] -> out [B,C,D,H,W]       */
/* ------------------------------------------------------------------------- */
void tmvs_oracle_homo_warp(const float *src, const float *rt, const float *depth, int per_pixel,
                           float *out, int B, int C, int D, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(3) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float ix, iy; oracle_taps t;
                    float dep = oracle_depth_at(depth, per_pixel, b, d, y, x, D, H, W);
                    oracle_source_index(rt + (size_t)b * 12, (float)x, (float)y, dep, H, W, &ix, &iy);
                    oracle_footprint(ix, iy, H, W, &t);
                    for (int c = 0; c < C; ++c) {
                        const float *img = src + ((size_t)b * C + c) * HW;
                        float acc = 0.0f;
                        if (t.ok[0]) acc += img[(size_t)t.y0 * W + t.x0] * t.w[0];
                        if (t.ok[1]) acc += img[(size_t)t.y0 * W + t.x0 + 1] * t.w[1];
                        if (t.ok[2]) acc += img[(size_t)(t.y0 + 1) * W + t.x0] * t.w[2];
                        if (t.ok[3]) acc += img[(size_t)(t.y0 + 1) * W + t.x0 + 1] * t.w[3];
                        out[((((size_t)b * C + c) * D + d) * H + y) * W + x] = acc;
                    }
                }
}

/* ------------------------------------------------------------------------- */
/* Cost volume: the view loop of DepthNet.forward, models/TransMVSNet.py:71-93 */
/*   similarity_i = (warped_i * ref[:, :, None]).mean(1)            (:80)      */
/*   similarity_sum += similarity_i * w_i ; weight_sum(1e-5) += w_i (:71-72,88-89)
 *   similarity = similarity_sum / weight_sum                       (:93)      */
/* ref [B,C,H,W]; src [Nsrc,B,C,H,W]; rt [Nsrc,B,12]; depth as above;          */
/* weights [B,Nsrc,H,W] (may be NULL -> only per-view output);                 */
/* sim_views [Nsrc,B,D,H,W] (may be NULL); agg [B,D,H,W] (may be NULL).        */
/* ------------------------------------------------------------------------- */
void tmvs_oracle_costvol_fwd(const float *ref, const float *src, const float *rt, const float *depth,
                             int per_pixel, const float *weights, float *sim_views, float *agg,
                             int B, int C, int D, int H, int W, int Nsrc)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(3) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float dep = oracle_depth_at(depth, per_pixel, b, d, y, x, D, H, W);
                    float sim_sum = 0.0f;
                    float w_sum = 1e-5f;
                    for (int i = 0; i < Nsrc; ++i) {
                        float ix, iy; oracle_taps t;
                        oracle_source_index(rt + ((size_t)i * B + b) * 12, (float)x, (float)y, dep, H, W, &ix, &iy);
                        oracle_footprint(ix, iy, H, W, &t);
                        const float *simg = src + (((size_t)i * B + b) * C) * HW;
                        float s = 0.0f;
                        for (int c = 0; c < C; ++c) {
                            const float *img = simg + (size_t)c * HW;
                            float acc = 0.0f;
                            if (t.ok[0]) acc += img[(size_t)t.y0 * W + t.x0] * t.w[0];
                            if (t.ok[1]) acc += img[(size_t)t.y0 * W + t.x0 + 1] * t.w[1];
                            if (t.ok[2]) acc += img[(size_t)(t.y0 + 1) * W + t.x0] * t.w[2];
                            if (t.ok[3]) acc += img[(size_t)(t.y0 + 1) * W + t.x0 + 1] * t.w[3];
                            float prod = acc * ref[((size_t)b * C + c) * HW + (size_t)y * W + x];
                            s += prod;
                        }
                        s = s / (float)C; /* .mean(1) */
                        if (sim_views) sim_views[((((size_t)i * B + b) * D + d) * H + y) * W + x] = s;
                        if (weights) {
                            float w = weights[(((size_t)b * Nsrc + i) * H + y) * W + x];
                            float sw = s * w;
                            sim_sum += sw;
                            w_sum += w;
                        }
                    }
                    if (agg && weights) agg[(((size_t)b * D + d) * H + y) * W + x] = sim_sum / w_sum;
                }
}

/* Aggregation alone (stage 1, after PixelwiseNet produced the weights):        */
/* models/TransMVSNet.py:71-72,88-93.  sim_views [Nsrc,B,D,H,W], w [B,Nsrc,H,W] */
void tmvs_oracle_aggregate_fwd(const float *sim_views, const float *weights, float *agg,
                               int B, int D, int H, int W, int Nsrc)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (size_t p = 0; p < (size_t)H * W; ++p) {
                float sim_sum = 0.0f, w_sum = 1e-5f;
                for (int i = 0; i < Nsrc; ++i) {
                    float w = weights[((size_t)b * Nsrc + i) * H * W + p];
                    float s = sim_views[(((size_t)i * B + b) * D + d) * H * W + p];
                    float sw = s * w;
                    sim_sum += sw;
                    w_sum += w;
                }
                agg[((size_t)b * D + d) * H * W + p] = sim_sum / w_sum;
            }
}

/* ------------------------------------------------------------------------- */
/* Backward of the cost volume wrt features (autograd of module.py:318-320 and */
/* TransMVSNet.py:80): given G_i = dL/d similarity_i  [Nsrc,B,D,H,W]            */
/*   grad_ref[b,c,p]   = sum_i sum_d G_i[d,p] * warped_i[c,d,p] / C             */
/*   grad_src_i[b,c,q] = sum_{p,d,tap->q} G_i[d,p] * ref[c,p] / C * w_tap       */
/* The scatter follows ATen grid_sampler_2d_backward (per-tap bounds check,     */
/* cuda/GridSampler.cuh:250-260 safe_add_2d), done sequentially here.           */
/* No gradient flows to cameras / depth hypotheses (module.py:294 no_grad).     */
/* ------------------------------------------------------------------------- */
void tmvs_oracle_costvol_bwd(const float *ref, const float *src, const float *rt, const float *depth,
                             int per_pixel, const float *G, float *grad_ref, float *grad_src,
                             int B, int C, int D, int H, int W, int Nsrc)
{
    const size_t HW = (size_t)H * W;
    const float invC = 1.0f / (float)C;
    memset(grad_ref, 0, sizeof(float) * (size_t)B * C * HW);
    memset(grad_src, 0, sizeof(float) * (size_t)Nsrc * B * C * HW);
    /* parallel over (view, batch): each owns a disjoint grad_src slice; grad_ref
       is accumulated per view into a private buffer and summed in view order. */
    float *gref_views = (float *)calloc((size_t)Nsrc * B * C * HW, sizeof(float));
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < Nsrc; ++i)
        for (int b = 0; b < B; ++b) {
            const float *simg = src + (((size_t)i * B + b) * C) * HW;
            float *gsrc = grad_src + (((size_t)i * B + b) * C) * HW;
            float *gref = gref_views + (((size_t)i * B + b) * C) * HW;
            for (int d = 0; d < D; ++d)
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < W; ++x) {
                        float g = G[((((size_t)i * B + b) * D + d) * H + y) * W + x];
                        float dep = oracle_depth_at(depth, per_pixel, b, d, y, x, D, H, W);
                        float ix, iy; oracle_taps t;
                        oracle_source_index(rt + ((size_t)i * B + b) * 12, (float)x, (float)y, dep, H, W, &ix, &iy);
                        oracle_footprint(ix, iy, H, W, &t);
                        size_t o00 = (size_t)t.y0 * W + t.x0;
                        for (int c = 0; c < C; ++c) {
                            const float *img = simg + (size_t)c * HW;
                            float r = ref[((size_t)b * C + c) * HW + (size_t)y * W + x];
                            float gw = g * invC;           /* mean backward */
                            float gwarp = gw * r;          /* d/d warped */
                            float acc = 0.0f;
                            if (t.ok[0]) { acc += img[o00] * t.w[0];         gsrc[(size_t)c * HW + o00] += t.w[0] * gwarp; }
                            if (t.ok[1]) { acc += img[o00 + 1] * t.w[1];     gsrc[(size_t)c * HW + o00 + 1] += t.w[1] * gwarp; }
                            if (t.ok[2]) { acc += img[o00 + W] * t.w[2];     gsrc[(size_t)c * HW + o00 + W] += t.w[2] * gwarp; }
                            if (t.ok[3]) { acc += img[o00 + W + 1] * t.w[3]; gsrc[(size_t)c * HW + o00 + W + 1] += t.w[3] * gwarp; }
                            gref[(size_t)c * HW + (size_t)y * W + x] += gw * acc;
                        }
                    }
        }
    for (int i = 0; i < Nsrc; ++i)
        for (size_t k = 0; k < (size_t)B * C * HW; ++k)
            grad_ref[k] += gref_views[(size_t)i * B * C * HW + k];
    free(gref_views);
}

/* ------------------------------------------------------------------------- */
/* Read-out: models/TransMVSNet.py:99-103 + models/module.py:474-482           */
/*   prob = exp(log_softmax(x, dim=1)); idx = argmax_d prob (first maximal);    */
/*   depth = gather(depth_values, idx); conf = max_d prob.                      */
/* log_softmax follows ATen: x - max - log(sum exp(x - max)).                   */
/* logits, depth_values [B,D,H,W]; prob [B,D,H,W] (may be NULL); idx int64.     */
/* ------------------------------------------------------------------------- */
static inline int oracle_gt_nan_aware(float v, float best)
{   /* torch.argmax/max: NaN compares as the maximum; first occurrence wins */
    if (isnan(best)) return 0;
    if (isnan(v)) return 1;
    return v > best;
}

void tmvs_oracle_softmax_wta(const float *logits, const float *depth_values, float *prob,
                             int64_t *idx, float *depth, float *conf, int B, int D, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < HW; ++p) {
            const float *x = logits + (size_t)b * D * HW + p;
            float m = x[0];
            for (int d = 1; d < D; ++d) if (oracle_gt_nan_aware(x[(size_t)d * HW], m)) m = x[(size_t)d * HW];
            float s = 0.0f;
            for (int d = 0; d < D; ++d) s += expf(x[(size_t)d * HW] - m);
            float ls = logf(s);
            float best = 0.0f; int bi = 0;
            for (int d = 0; d < D; ++d) {
                float pv = expf((x[(size_t)d * HW] - m) - ls);
                if (prob) prob[((size_t)b * D + d) * HW + p] = pv;
                if (d == 0 || oracle_gt_nan_aware(pv, best)) { best = pv; bi = d; }
            }
            idx[(size_t)b * HW + p] = bi;
            depth[(size_t)b * HW + p] = depth_values[((size_t)b * D + bi) * HW + p];
            conf[(size_t)b * HW + p] = best;
        }
}

/* depth_wta: models/module.py:474-482 -- argmax (first maximal) + gather.      */
void tmvs_oracle_depth_wta(const float *p, const float *depth_values, int64_t *idx, float *depth,
                           int B, int D, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t q = 0; q < HW; ++q) {
            const float *x = p + (size_t)b * D * HW + q;
            float best = x[0]; int bi = 0;
            for (int d = 1; d < D; ++d)
                if (oracle_gt_nan_aware(x[(size_t)d * HW], best)) { best = x[(size_t)d * HW]; bi = d; }
            idx[(size_t)b * HW + q] = bi;
            depth[(size_t)b * HW + q] = depth_values[((size_t)b * D + bi) * HW + q];
        }
}

/* depth_regression(p, depth_values): ABSENT from this fork;    */
/* upstream definition: sum(p * depth_values, dim=1), depth_values [B,D] is     */
/* viewed as [B,D,1,1].  Parity unpinned by the reference.                      */
void tmvs_oracle_depth_regression(const float *p, const float *depth_values, int per_pixel,
                                  float *depth, int B, int D, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t q = 0; q < HW; ++q) {
            float acc = 0.0f;
            for (int d = 0; d < D; ++d) {
                float dv = per_pixel ? depth_values[((size_t)b * D + d) * HW + q] : depth_values[(size_t)b * D + d];
                float t = p[((size_t)b * D + d) * HW + q] * dv;
                acc += t;
            }
            depth[(size_t)b * HW + q] = acc;
        }
}

/* ------------------------------------------------------------------------- */
/* PixelwiseNet in eval mode: models/TransMVSNet.py:10-30 with ConvBnReLU3D    */
/* (models/module.py:214-221): 1x1x1 Conv3d (no bias) -> BatchNorm3d (running   */
/* statistics) -> ReLU, twice (1->16->8), then Conv3d 8->1 with bias, Sigmoid,  */
/* max over D.  NOT folded: conv and batch-norm are applied one after the other */
/* as the reference does.  sim [B,D,H,W] of ONE view -> weight [B,H,W].         */
/* bn arrays are [4][n]: gamma, beta, running_mean, running_var.                */
/* ------------------------------------------------------------------------- */
void tmvs_oracle_pixelwise_weight(const float *sim, const float *w0, const float *bn0, const float *w1,
                                  const float *bn1, const float *w2, float b2, float eps, float *weight,
                                  int B, int D, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < HW; ++p) {
            float best = -INFINITY;
            for (int d = 0; d < D; ++d) {
                float x = sim[((size_t)b * D + d) * HW + p];
                float h0[16], h1[8];
                for (int c = 0; c < 16; ++c) {
                    float y = w0[c] * x;
                    y = (y - bn0[2 * 16 + c]) / sqrtf(bn0[3 * 16 + c] + eps) * bn0[c] + bn0[16 + c];
                    h0[c] = y > 0.0f ? y : 0.0f;
                }
                for (int j = 0; j < 8; ++j) {
                    float y = 0.0f;
                    for (int c = 0; c < 16; ++c) y += w1[j * 16 + c] * h0[c];
                    y = (y - bn1[2 * 8 + j]) / sqrtf(bn1[3 * 8 + j] + eps) * bn1[j] + bn1[8 + j];
                    h1[j] = y > 0.0f ? y : 0.0f;
                }
                float o = b2;
                for (int j = 0; j < 8; ++j) o += w2[j] * h1[j];
                float sg = 1.0f / (1.0f + expf(-o));
                if (sg > best) best = sg;
            }
            weight[(size_t)b * HW + p] = best;
        }
}

/* ------------------------------------------------------------------------- */
/* fusibile depth-map fusion, restated on the CPU.  The    */
/* PIN of the GPU path is no longer this restatement but the reference's own */
/* kernel file compiled for sm_100a (oracle/build.py build_fusibile_ref,     */
/* tests/test_gpu_fusion.py: bit-identical); this C version remains the      */
/* CPU-side checker (IEEE arithmetic, modelled texture filter).              */
/*   fusibile.cu:89-173  kernel `fusibile` (one reference camera)            */
/*   fusibile.cu:54-69   get_3dpoint_cu, :71-85 project_on_camera,           */
/*   fusibile.cu:44-52   depth_convert_cu, :21-33 float4 operators (w := 0)  */
/*   fusibile.cu:175-210 copy_pc_to_host (buffer never cleared: carry-over)  */
/* nvcc contracts m0*x + m1*y + m2*z to mul(m1,y), fma(m0,x,.), fma(m2,z,.)   */
/* (read off the SASS of the reference's kernel); the same explicit sequence */
/* is used here and in csrc/tmvs_fusion.cu (its TMVS_FUSE_IEEE arithmetic).  */
/* The texture fetch (main.cpp:46-66: float4, cudaFilterModeLinear,          */
/* unnormalised coordinates) is EMULATED from the CUDA programming guide's   */
/* description of linear filtering (xB = x - 0.5, i = floor(xB), fractions in */
/* 9-bit fixed point with 8 fractional bits, clamp to edge) refined by a     */
/* probe of the B200 texture unit (see fuse_tex_linear).  The GPU test       */
/* measures how far this emulation is from the texture unit (<= 1 ulp).      */
/* cams: [V][28] = P(12) RK_inv(9) C(3) P34(3) K00(1).                        */
/* ------------------------------------------------------------------------- */
static void fuse_backproject(const float *cam, int px, int py, float depth, float *X)
{
    const float *m = cam + 12, *p34 = cam + 24;
    const float x = fmaf(depth, (float)px, -p34[0]);
    const float y = fmaf(depth, (float)py, -p34[1]);
    const float z = depth - p34[2];
    X[0] = fmaf(m[2], z, fmaf(m[0], x, m[1] * y));
    X[1] = fmaf(m[5], z, fmaf(m[3], x, m[4] * y));
    X[2] = fmaf(m[8], z, fmaf(m[6], x, m[7] * y));
}

static int fuse_clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

/* tex2D<float4>(tex, u, v), linear filter, unnormalised, clamp */
static void fuse_tex_linear(const float *img, int H, int W, float u, float v, float *out)
{
    const float xb = u - 0.5f, yb = v - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    float a = floorf((xb - fx) * 256.0f + 0.5f) / 256.0f;
    float b = floorf((yb - fy) * 256.0f + 0.5f) / 256.0f;
    const int i0 = fuse_clampi((int)fx, W - 1), i1 = fuse_clampi((int)fx + 1, W - 1);
    const int j0 = fuse_clampi((int)fy, H - 1), j1 = fuse_clampi((int)fy + 1, H - 1);
    const float *t00 = img + ((size_t)j0 * W + i0) * 4, *t01 = img + ((size_t)j0 * W + i1) * 4;
    const float *t10 = img + ((size_t)j1 * W + i0) * 4, *t11 = img + ((size_t)j1 * W + i1) * 4;
    /* measured on B200 (scripts/probe_tex_filter.py, 20000 random positions): the unit does not blend with the separable
       products of the two 8-bit fractions -- it rounds ONE product, w11 = round(a*b*256)/256, and derives the other three
       weights by subtraction (w01 = a - w11, w10 = b - w11, w00 = 1 - a - b + w11), so the four always sum to 1 and ramps
       are reproduced exactly.  With these weights a double-precision blend rounded to float matches the unit bit for
       bit on 99.7 % of the samples and to 1 ulp on the rest. */
    const float w11 = floorf(a * b * 256.0f + 0.5f) / 256.0f;
    const float w01 = a - w11, w10 = b - w11, w00 = 1.0f - a - b + w11;
    for (int k = 0; k < 4; ++k) {
        const double r = (double)w00 * t00[k] + (double)w01 * t01[k] + (double)w10 * t10[k] + (double)w11 * t11[k];
        out[k] = (float)r;
    }
}

/* returns the number of points found (may exceed capacity; only the first `capacity` are written).
   points: [capacity][8] = x y z 0 b g r 0.  scratch: [H*W][8] floats, the reference's per-pixel point buffer. */
long long tmvs_oracle_fusibile(const float *images, const float *cams, int V, int H, int W, float depth_threshold,
                               int consistent_threshold, int carry_over, float *points, long long capacity,
                               float *scratch)
{
    const size_t HW = (size_t)H * W;
    long long n = 0;
    memset(scratch, 0, HW * 8 * sizeof(float));                     /* point_cloud.h:20 */
    for (int c = 0; c < V; ++c) {
        const float *ref = cams + (size_t)c * 28;
        const float *img_c = images + (size_t)c * HW * 4;
        if (!carry_over) memset(scratch, 0, HW * 8 * sizeof(float));
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float sum_T[4];
                fuse_tex_linear(img_c, H, W, x + 0.5f, y + 0.5f, sum_T);
                float depth = sum_T[3];
                if ((double)depth <= 425.001) continue;
                float X[3], sum_X[3];
                fuse_backproject(ref, x, y, depth, X);
                sum_X[0] = X[0]; sum_X[1] = X[1]; sum_X[2] = X[2];
                int count = 0;
                for (int i = 0; i < V && count < 2 * consistent_threshold; ++i) {
                    if (i == c) continue;
                    const float *cam = cams + (size_t)i * 28;
                    const float tx = fmaf(cam[2], X[2], fmaf(cam[0], X[0], cam[1] * X[1])) + cam[3];
                    const float ty = fmaf(cam[6], X[2], fmaf(cam[4], X[0], cam[5] * X[1])) + cam[7];
                    const float tz = fmaf(cam[10], X[2], fmaf(cam[8], X[0], cam[9] * X[1])) + cam[11];
                    const float ptx = tx / tz, pty = ty / tz;
                    depth = tz;
                    if (ptx < 0 || ptx >= W || pty < 0 || pty >= H) continue;
                    float tmp_T[4];
                    fuse_tex_linear(images + (size_t)i * HW * 4, H, W, ptx + 0.5f, pty + 0.5f, tmp_T);
                    if ((double)tmp_T[3] <= 425.001) continue;
                    const float bx = ref[21] - cam[21], by = ref[22] - cam[22], bz = ref[23] - cam[23];
                    const float baseline = sqrtf(fmaf(bz, bz, fmaf(bx, bx, by * by)));
                    const float fb = ref[27] * baseline;
                    const float depth_disp = fb / depth, temp_disp = fb / tmp_T[3];
                    if (fabsf(depth_disp - temp_disp) < depth_threshold) {
                        float Y[3];
                        fuse_backproject(cam, (int)ptx, (int)pty, tmp_T[3], Y);
                        sum_X[0] += Y[0]; sum_X[1] += Y[1]; sum_X[2] += Y[2];
                        sum_T[0] += tmp_T[0]; sum_T[1] += tmp_T[1]; sum_T[2] += tmp_T[2]; sum_T[3] = 0.0f;
                        ++count;
                    }
                }
                if (count >= consistent_threshold) {
                    const float k = (float)count + 1.0f;
                    float *o = scratch + ((size_t)y * W + x) * 8;
                    o[0] = sum_X[0] / k; o[1] = sum_X[1] / k; o[2] = sum_X[2] / k; o[3] = 0.0f;
                    o[4] = sum_T[0] / k; o[5] = sum_T[1] / k; o[6] = sum_T[2] / k; o[7] = 0.0f;
                }
            }
        /* copy_pc_to_host: y-major, x-minor, all three coordinates non-zero */
        for (size_t p = 0; p < HW; ++p) {
            const float *s = scratch + p * 8;
            if (s[0] != 0 && s[1] != 0 && s[2] != 0) {
                if (n < capacity) memcpy(points + (size_t)n * 8, s, 8 * sizeof(float));
                ++n;
            }
        }
    }
    return n;
}

/* the filter emulation alone: out[j] = tex2D<float4>(img, uv[j]) as fuse_tex_linear models it */
void tmvs_oracle_tex_linear(const float *img, int H, int W, const float *uv, float *out, int n)
{
    for (int j = 0; j < n; ++j) fuse_tex_linear(img, H, W, uv[2 * j], uv[2 * j + 1], out + 4 * j);
}
