// Backward of the fused cost volume wrt the features, for sm_100a -- deterministic and atomic-free.
//
// Autograd of models/module.py:318-320 (grid_sample) and models/TransMVSNet.py:80 ((warped*ref).mean(1)),
// SURVEY.md 3.4.  Input: G_i = dL/d similarity_i  [Nsrc][B][D][H][W].
//
//   grad_ref[b,c,p]   = 1/C * sum_i sum_d G_i[d,p] * warped_i[c,d,p]               -- a gather
//   grad_src_i[b,c,q] = 1/C * sum_{(p,d,tap) -> q} G_i[d,p] * w_tap(p,d) * ref[c,p] -- the grid_sample scatter
//
// ATen does the scatter with float atomicAdd (cuda/GridSampler.cuh:250-260), so its result depends on the
// order the atomics retire.  Here the scatter is turned into a gather with an explicit order:
//   1. bwd_bbox_kernel: for every (view, reference tile T of 32x8 pixels, depth plane d) the bounding box of
//      the source pixels T's bilinear footprints touch -- exact, from the same coordinate code as the forward.
//   2. bwd_src_kernel: one CTA OWNS one 32x8 tile S of grad_src (one thread per source pixel, its C
//      accumulators in registers).  It scans the boxes in two levels (8x8-tile groups, then the (T, d) pairs of
//      the groups that touch S), and for every (T, d) that overlaps S, in (group, T, d) order, up to 4
//      consecutive planes of a tile per round:  phase 1 -- thread t recomputes the footprint of reference pixel t of T and registers itself in
//      a shared-memory cell grid indexed by its north-west source pixel (a few rounds of plain stores; which
//      thread wins a round is irrelevant because phase 2 orders the ids);  phase 2 -- each owner reads the <= 4 cells whose taps
//      hit its pixel, SORTS the registered thread ids, and accumulates k * ref[:, p] in that order.
//   Every output element is written exactly once by its owner, in a fixed summation order: no atomics on
//   data, bit-reproducible run to run.
// grad_ref is a plain gather (forward-shaped kernel), split over views and reduced in view order.
#include "tmvs_common.cuh"


// cell-table path of grad_src (tmvs_costvol_bwd_cells.cu)
size_t tmvs_bwd_cells_bytes_per_pair(int D, int H, int W);
int tmvs_bwd_src_cells(const float4 *refp, const float *depth, int per_pixel, const float *G, float *grad_src,
                       char *tables, int pairs_per_pass, int *flags, int *overflow, int *tile_overflow, int b_total,
                       int b_first, int bc,
                       int n_src, int C, int D, int H, int W, const TmvsGeom &geom, cudaStream_t st);

int tmvs_bwd_warp_cells(const float *depth, int per_pixel, const float *gout, float *grad_src, char *tables, int *flags,
                        int *overflow, int *tile_overflow, int b, int C, int D, int H, int W, const TmvsGeom &geom,
                        cudaStream_t st);

namespace {

constexpr int kTX = 32, kTY = 8, kThreads = kTX * kTY;
constexpr int kCellW = kTX + 1, kCellH = kTY + 1, kCells = kCellW * kCellH;
constexpr int kEmpty = 0x7fffffff;

// --------------------------------------------------------------------------------------------- grad_ref
// resident CTAs per SM (registers): the gather is a chain of dependent L1 loads, so occupancy is what hides its latency
#ifndef TMVS_BWDREF_MINB8
#define TMVS_BWDREF_MINB8 2
#endif
#ifndef TMVS_BWDREF_MINB4
#define TMVS_BWDREF_MINB4 4
#endif
#ifndef TMVS_BWDREF_MINB2
#define TMVS_BWDREF_MINB2 4
#endif
template <int C4T> struct BwdRefMinBlocks { static constexpr int value = C4T >= 8 ? TMVS_BWDREF_MINB8 : (C4T >= 4 ? TMVS_BWDREF_MINB4 : TMVS_BWDREF_MINB2); };

template <int C4T, bool EXACT, bool PER_PIXEL>
__global__ void __launch_bounds__(kThreads, BwdRefMinBlocks<C4T>::value)
bwd_ref_kernel(const float4 *__restrict__ packed, const float *__restrict__ depth, const float *__restrict__ G,
               float *__restrict__ partial, int b_total, int b_first, int b_chunk, int C, int c4, int D, int H, int W,
               const __grid_constant__ TmvsGeom geom)
{
    const int x = blockIdx.x * kTX + threadIdx.x;
    const int y = blockIdx.y * kTY + threadIdx.y;
    if (x >= W || y >= H) return;
    const int i = blockIdx.z / b_chunk, bl = blockIdx.z - i * b_chunk;
    const int b = b_first + bl;
    const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
    float rt[12];
    tmvs_geom_rt(geom, i, bl, b_chunk, rt);
    const TmvsRay ray = tmvs_ray(rt, (float)x, (float)y, geom.ray_unfused);
    const float inv_c = 1.0f / (float)C;
    const TmvsDims dims = tmvs_dims(H, W, geom.arith);
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    const float4 *img = packed + ((size_t)i * b_total + b) * pk.slice;
    const float *gp = G + ((size_t)i * b_total + b) * D * HW + pix;
    float4 acc[C4T];
#pragma unroll
    for (int g = 0; g < C4T; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 0; d < D; ++d) {
        const float dep = PER_PIXEL ? __ldg(depth + ((size_t)b * D + d) * HW + pix) : __ldg(depth + (size_t)b * D + d);
        const float gw = __ldg(gp + (size_t)d * HW) * inv_c;
        const TmvsTaps t = tmvs_taps(ray, rt, dep, dims);
        if (!t.any) continue;
        const float k00 = t.ok00 ? gw * t.w00 : 0.0f, k01 = t.ok01 ? gw * t.w01 : 0.0f;
        const float k10 = t.ok10 ? gw * t.w10 : 0.0f, k11 = t.ok11 ? gw * t.w11 : 0.0f;
        const int xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
        const int ra = min(max(t.y0, 0), H - 1) * pk.row, rb = min(max(t.y0 + 1, 0), H - 1) * pk.row;
        const float4 *p00 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xa, ra)), *p01 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xb, ra));
        const float4 *p10 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xa, rb)), *p11 = tmvs_pk_ptr(img, tmvs_pk_off(pk, xb, rb));
#pragma unroll
        for (int g = 0; g < C4T; ++g) {
            if (EXACT || g < c4) {
                const float4 a = ldg4(p00 + g * 8), bq = ldg4(p01 + g * 8);
                const float4 cq = ldg4(p10 + g * 8), dq = ldg4(p11 + g * 8);
                acc[g].x += k00 * a.x + k01 * bq.x + k10 * cq.x + k11 * dq.x;
                acc[g].y += k00 * a.y + k01 * bq.y + k10 * cq.y + k11 * dq.y;
                acc[g].z += k00 * a.z + k01 * bq.z + k10 * cq.z + k11 * dq.z;
                acc[g].w += k00 * a.w + k01 * bq.w + k10 * cq.w + k11 * dq.w;
            }
        }
    }
    float *o = partial + (((size_t)i * b_total + b) * C) * HW + pix;
#pragma unroll
    for (int g = 0; g < C4T; ++g) {
        const float v[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * g + j < C) o[(size_t)(4 * g + j) * HW] = v[j];
    }
}

__global__ void __launch_bounds__(256)
bwd_ref_reduce_kernel(const float *__restrict__ partial, float *__restrict__ grad_ref, size_t n, int n_src)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float s = __ldg(partial + k);
    for (int i = 1; i < n_src; ++i) s += __ldg(partial + (size_t)i * n + k);     // fixed view order
    grad_ref[k] = s;
}

// --------------------------------------------------------------------------------------------- grad_src
// Bounding box (inclusive, in source pixels, in-bounds taps only) of the footprints of tile T at plane d.
template <bool PER_PIXEL>
__global__ void __launch_bounds__(kThreads)
bwd_bbox_kernel(const float *__restrict__ depth, int4 *__restrict__ bbox, int b_first, int b_chunk, int D, int H,
                int W, int n_tx, int n_tiles, const int *__restrict__ gate, const __grid_constant__ TmvsGeom geom)
{
    __shared__ int red[4][kTY];
    if (gate && gate[blockIdx.z] == 0) return;   // the cell-table path (tmvs_costvol_bwd_cells.cu) served this pair
    const int x = blockIdx.x * kTX + threadIdx.x;
    const int y = blockIdx.y * kTY + threadIdx.y;
    const bool valid = x < W && y < H;
    const int i = blockIdx.z / b_chunk, bl = blockIdx.z - i * b_chunk;
    const int b = b_first + bl;
    const size_t HW = (size_t)H * W, pix = (size_t)min(y, H - 1) * W + min(x, W - 1);
    float rt[12];
    tmvs_geom_rt(geom, i, bl, b_chunk, rt);
    const TmvsRay ray = tmvs_ray(rt, (float)x, (float)y, geom.ray_unfused);
    const TmvsDims dims = tmvs_dims(H, W, geom.arith);
    const int tile = blockIdx.y * n_tx + blockIdx.x;
    int4 *out = bbox + ((size_t)blockIdx.z * n_tiles + tile) * D;
    const int warp = threadIdx.y, lane = threadIdx.x;
    for (int d = 0; d < D; ++d) {
        const float dep = PER_PIXEL ? __ldg(depth + ((size_t)b * D + d) * HW + pix) : __ldg(depth + (size_t)b * D + d);
        const TmvsTaps t = tmvs_taps(ray, rt, dep, dims);
        int lo_x = kEmpty, lo_y = kEmpty, hi_x = -1, hi_y = -1;
        if (valid && t.any) {
            lo_x = max(t.x0, 0); hi_x = min(t.x0 + 1, W - 1);
            lo_y = max(t.y0, 0); hi_y = min(t.y0 + 1, H - 1);
        }
        lo_x = __reduce_min_sync(0xffffffffu, lo_x);
        lo_y = __reduce_min_sync(0xffffffffu, lo_y);
        hi_x = __reduce_max_sync(0xffffffffu, hi_x);
        hi_y = __reduce_max_sync(0xffffffffu, hi_y);
        if (lane == 0) { red[0][warp] = lo_x; red[1][warp] = lo_y; red[2][warp] = hi_x; red[3][warp] = hi_y; }
        __syncthreads();
        if (warp == 0 && lane == 0) {
            int4 r = make_int4(kEmpty, kEmpty, -1, -1);
#pragma unroll
            for (int w = 0; w < kTY; ++w) {
                r.x = min(r.x, red[0][w]); r.y = min(r.y, red[1][w]);
                r.z = max(r.z, red[2][w]); r.w = max(r.w, red[3][w]);
            }
            out[d] = r;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void cswap(int &a, int &b)
{
    const int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

// Registration uses RELAXED (morally strong) shared-memory stores and loads: several threads may store their id
// to the same cell in the same round and exactly one value survives.  These are plain STS / LDS in SASS -- no
// read-modify-write atomic -- but, unlike weak accesses, concurrent relaxed stores are not a data race in the
// PTX memory model.
__device__ __forceinline__ void st_relaxed_s32(int *p, int v)
{
    asm volatile("st.relaxed.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_s32(const int *p)
{
    int v;
    asm volatile("ld.relaxed.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

constexpr int kGroup = 8;            // tiles per side of a scan group (8x8 tiles = 256x64 pixels)
#ifndef TMVS_BWD_PLANES
#define TMVS_BWD_PLANES 4
#endif
#ifndef TMVS_BWD_MINB8
#define TMVS_BWD_MINB8 4
#endif
#ifndef TMVS_BWD_MINB4
#define TMVS_BWD_MINB4 5
#endif
#ifndef TMVS_BWD_MINB2
#define TMVS_BWD_MINB2 5
#endif
constexpr int kPlanes = TMVS_BWD_PLANES;   // depth planes of one reference tile handled per registration round
constexpr int kSlots2 = 4;           // footprints per cell per plane on the fast path

// Union over ALL planes of the boxes of the 8x8 tiles of a group: the coarse level of the scan.
__global__ void __launch_bounds__(kThreads)
bwd_gbox_kernel(const int4 *__restrict__ bbox, int4 *__restrict__ gbox, int D, int n_tx, int n_ty, int n_gx,
                int n_tiles, int n_groups, const int *__restrict__ gate)
{
    __shared__ int red[4][kTY];
    if (gate && gate[blockIdx.y] == 0) return;
    const int group = blockIdx.x, vb = blockIdx.y;
    const int gy = group / n_gx, gx = group - gy * n_gx;
    const int tid = threadIdx.y * kTX + threadIdx.x;
    int4 r = make_int4(kEmpty, kEmpty, -1, -1);
    const int4 *boxes = bbox + (size_t)vb * n_tiles * D;
    for (int j = tid; j < kGroup * kGroup * D; j += kThreads) {
        const int tl = j / D, d = j - tl * D;
        const int ty = gy * kGroup + tl / kGroup, tx = gx * kGroup + tl % kGroup;
        if (tx < n_tx && ty < n_ty) {
            const int4 bb = __ldg(boxes + (size_t)(ty * n_tx + tx) * D + d);
            r.x = min(r.x, bb.x); r.y = min(r.y, bb.y); r.z = max(r.z, bb.z); r.w = max(r.w, bb.w);
        }
    }
    r.x = __reduce_min_sync(0xffffffffu, r.x); r.y = __reduce_min_sync(0xffffffffu, r.y);
    r.z = __reduce_max_sync(0xffffffffu, r.z); r.w = __reduce_max_sync(0xffffffffu, r.w);
    if (threadIdx.x == 0) { red[0][threadIdx.y] = r.x; red[1][threadIdx.y] = r.y; red[2][threadIdx.y] = r.z; red[3][threadIdx.y] = r.w; }
    __syncthreads();
    if (tid == 0) {
        int4 o = make_int4(kEmpty, kEmpty, -1, -1);
        for (int w = 0; w < kTY; ++w) { o.x = min(o.x, red[0][w]); o.y = min(o.y, red[1][w]); o.z = max(o.z, red[2][w]); o.w = max(o.w, red[3][w]); }
        gbox[(size_t)vb * n_groups + group] = o;
    }
}

template <int C4T> struct BwdMinBlocks { static constexpr int value = C4T >= 8 ? TMVS_BWD_MINB8 : (C4T >= 4 ? TMVS_BWD_MINB4 : TMVS_BWD_MINB2); };

template <int C4T, bool EXACT, bool PER_PIXEL>
__global__ void __launch_bounds__(kThreads, BwdMinBlocks<C4T>::value)
bwd_src_kernel(const float4 *__restrict__ refp, const float *__restrict__ depth, const float *__restrict__ G,
               const int4 *__restrict__ bbox, const int4 *__restrict__ gbox, float *__restrict__ grad_src, int b_total,
               int b_first, int b_chunk, int C, int c4, int D, int H, int W, int n_tx, int n_ty, int n_tiles,
               int n_gx, int n_groups, const int *__restrict__ gate, const __grid_constant__ TmvsGeom geom)
{
    __shared__ int cell[kPlanes][kSlots2][kCells];
    __shared__ float krec[kPlanes][4][kThreads];
    __shared__ int xyrec[kPlanes][kThreads];           // footprint origin, (y0 + 2) << 16 | (x0 + 2), for the exhaustive path
    __shared__ int pixrec[kThreads];
    __shared__ int hits[kThreads];
    __shared__ int ghits[kThreads];
    __shared__ int wcount[kTY];

    // the cell-table path served every tile but those a slot-less footprint touches
    if (gate && gate[(size_t)blockIdx.z * n_tiles + blockIdx.y * n_tx + blockIdx.x] == 0) return;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kTX + tx;
    const int s_x = blockIdx.x * kTX, s_y = blockIdx.y * kTY;      // the owned source tile
    const int qx = s_x + tx, qy = s_y + ty;
    const bool q_valid = qx < W && qy < H;
    const int i = blockIdx.z / b_chunk, bl = blockIdx.z - i * b_chunk;
    const int b = b_first + bl;
    const size_t HW = (size_t)H * W;
    float rt[12];
    tmvs_geom_rt(geom, i, bl, b_chunk, rt);
    const float inv_c = 1.0f / (float)C;
    const TmvsDims dims = tmvs_dims(H, W, geom.arith);
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    const float4 *rimg = refp + (size_t)b * pk.slice;
    const float *gview = G + ((size_t)i * b_total + b) * D * HW;
    const int4 *boxes = bbox + (size_t)blockIdx.z * n_tiles * D;
    const int4 *gboxes = gbox + (size_t)blockIdx.z * n_groups;

    float4 acc[C4T];
#pragma unroll
    for (int g = 0; g < C4T; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto accumulate = [&](float k, int off) {
#pragma unroll
        for (int g = 0; g < C4T; ++g) {
            if (EXACT || g < c4) {
                const float4 rv = ldg4(tmvs_pk_ptr(rimg, (unsigned)off) + g * 8);
                acc[g].x = fmaf(k, rv.x, acc[g].x);
                acc[g].y = fmaf(k, rv.y, acc[g].y);
                acc[g].z = fmaf(k, rv.z, acc[g].z);
                acc[g].w = fmaf(k, rv.w, acc[g].w);
            }
        }
    };
    auto overlaps = [&](const int4 &bb) {
        return bb.x <= s_x + kTX - 1 && bb.z >= s_x && bb.y <= s_y + kTY - 1 && bb.w >= s_y;
    };
    // ordered compaction of a per-thread flag into list[]: returns the number of set flags (uniform)
    auto compact = [&](bool flag, int value, int *list) {
        const unsigned ballot = __ballot_sync(0xffffffffu, flag);
        __syncthreads();                       // previous users of wcount / list are done
        if (tx == 0) wcount[ty] = __popc(ballot);
        __syncthreads();
        int prefix = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kTY; ++w) {
            const int c = wcount[w];
            if (w < ty) prefix += c;
            total += c;
        }
        if (flag) list[prefix + __popc(ballot & ((1u << tx) - 1u))] = value;
        __syncthreads();
        return total;
    };

    // ---- level 1: groups of 8x8 reference tiles whose all-plane box touches S (visited in group order)
    for (int gbase = 0; gbase < n_groups; gbase += kThreads) {
        const int gj = gbase + tid;
        const bool ghit = gj < n_groups && overlaps(__ldg(gboxes + gj));
        const int n_gh = compact(ghit, gj, ghits);
        for (int gh = 0; gh < n_gh; ++gh) {
            const int group = ghits[gh];
            const int g_y = group / n_gx, g_x = group - g_y * n_gx;
            // ---- level 2: the (tile, plane) pairs of this group, tile-major / plane-minor
            const int n_pairs = kGroup * kGroup * D;
            for (int base = 0; base < n_pairs; base += kThreads) {
                const int j = base + tid;
                bool hit = false;
                int code = 0;
                if (j < n_pairs) {
                    const int tl = j / D, d = j - tl * D;
                    const int t_y = g_y * kGroup + tl / kGroup, t_x = g_x * kGroup + tl % kGroup;
                    if (t_x < n_tx && t_y < n_ty) {
                        const int tile = t_y * n_tx + t_x;
                        hit = overlaps(__ldg(boxes + (size_t)tile * D + d));
                        code = tile * D + d;
                    }
                }
                const int total = compact(hit, code, hits);

                int h = 0;
                while (h < total) {
                    // a run of up to kPlanes consecutive planes of ONE reference tile
                    const int jj = hits[h];
                    const int tile = jj / D, d_first = jj - tile * D;
                    int run = 1;
                    while (run < kPlanes && h + run < total && hits[h + run] == jj + run && d_first + run < D) ++run;
                    h += run;
                    const int t_y = tile / n_tx, t_x = tile - t_y * n_tx;
                    // ---- phase 1: thread t = reference pixel t of tile T, for each plane of the run
                    const int px = t_x * kTX + tx, py = t_y * kTY + ty;
                    const bool p_valid = px < W && py < H;
                    int my_cell[kPlanes];
                    TmvsRay ray;
                    if (p_valid) ray = tmvs_ray(rt, (float)px, (float)py, geom.ray_unfused);
                    pixrec[tid] = tmvs_pk_off(pk, min(px, W - 1), min(py, H - 1) * pk.row);   // packed word of ref pixel p
#pragma unroll
                    for (int pl = 0; pl < kPlanes; ++pl) {
                        my_cell[pl] = -1;
                        float k00 = 0.f, k01 = 0.f, k10 = 0.f, k11 = 0.f;
                        int x0 = kEmpty, y0 = kEmpty;
                        if (pl < run && p_valid) {
                            const int d = d_first + pl;
                            const size_t pix = (size_t)py * W + px;
                            const float dep = PER_PIXEL ? __ldg(depth + ((size_t)b * D + d) * HW + pix) : __ldg(depth + (size_t)b * D + d);
                            const TmvsTaps t = tmvs_taps(ray, rt, dep, dims);
                            const int cx = t.x0 - s_x + 1, cy = t.y0 - s_y + 1;
                            if (t.any && cx >= 0 && cx < kCellW && cy >= 0 && cy < kCellH) {
                                const float gw = __ldg(gview + (size_t)d * HW + pix) * inv_c;
                                k00 = t.ok00 ? gw * t.w00 : 0.0f; k01 = t.ok01 ? gw * t.w01 : 0.0f;
                                k10 = t.ok10 ? gw * t.w10 : 0.0f; k11 = t.ok11 ? gw * t.w11 : 0.0f;
                                my_cell[pl] = cy * kCellW + cx;
                                x0 = t.x0; y0 = t.y0;
                            }
                        }
                        if (pl < run) {
                            krec[pl][0][tid] = k00; krec[pl][1][tid] = k01; krec[pl][2][tid] = k10; krec[pl][3][tid] = k11;
                            xyrec[pl][tid] = x0 == kEmpty ? -1 : (((y0 + 2) << 16) | (x0 + 2));
                        }
                    }
                    for (int c = tid; c < run * kSlots2 * kCells; c += kThreads) (&cell[0][0][0])[c] = kEmpty;
                    __syncthreads();
                    // registration rounds: plain stores; one registrant per cell per round survives, the rest
                    // retry in the next slot.  WHICH one survives does not matter: phase 2 orders the ids of a cell.
                    unsigned pend = 0;
#pragma unroll
                    for (int pl = 0; pl < kPlanes; ++pl)
                        if (my_cell[pl] >= 0) pend |= 1u << pl;
                    int left = 0;
#pragma unroll 1
                    for (int r = 0; r < kSlots2; ++r) {
#pragma unroll
                        for (int pl = 0; pl < kPlanes; ++pl)
                            if (pend & (1u << pl)) st_relaxed_s32(&cell[pl][r][my_cell[pl]], tid);
                        __syncthreads();
#pragma unroll
                        for (int pl = 0; pl < kPlanes; ++pl)
                            if ((pend & (1u << pl)) && ld_relaxed_s32(&cell[pl][r][my_cell[pl]]) == tid) pend &= ~(1u << pl);
                        left = __syncthreads_or(pend != 0);
                        if (!left) break;
                    }
                    // ---- phase 2: the owner of source pixel q gathers the taps that land on it, plane by plane
                    if (q_valid) {
                        const int cls_cell[4] = {(ty + 1) * kCellW + (tx + 1), (ty + 1) * kCellW + tx,
                                                 ty * kCellW + (tx + 1), ty * kCellW + tx};
                        for (int pl = 0; pl < run; ++pl) {
                            if (!left) {
#pragma unroll
                                for (int cls = 0; cls < 4; ++cls) {
                                    int i0 = cell[pl][0][cls_cell[cls]], i1 = cell[pl][1][cls_cell[cls]];
                                    int i2 = cell[pl][2][cls_cell[cls]], i3 = cell[pl][3][cls_cell[cls]];
                                    if (i1 != kEmpty) {           // several footprints share the cell: fix the order
                                        cswap(i0, i1); cswap(i2, i3); cswap(i0, i2); cswap(i1, i3); cswap(i1, i2);
                                    }
                                    const int ids[4] = {i0, i1, i2, i3};
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        if (ids[u] != kEmpty) {
                                            const float k = krec[pl][cls][ids[u]];
                                            if (k != 0.0f) accumulate(k, pixrec[ids[u]]);
                                        }
                                    }
                                }
                            } else {
                                // exhaustive, still ordered (class, then reference-pixel id)
#pragma unroll 1
                                for (int cls = 0; cls < 4; ++cls) {
                                    const int want = ((qy - (cls >> 1) + 2) << 16) | (qx - (cls & 1) + 2);
#pragma unroll 1
                                    for (int t2 = 0; t2 < kThreads; ++t2) {
                                        if (xyrec[pl][t2] == want) {
                                            const float k = krec[pl][cls][t2];
                                            if (k != 0.0f) accumulate(k, pixrec[t2]);
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        }
    }
    if (q_valid) {
        float *o = grad_src + (((size_t)i * b_total + b) * C) * HW + (size_t)qy * W + qx;
#pragma unroll
        for (int g = 0; g < C4T; ++g) {
            const float v[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2)
                if (4 * g + j2 < C) o[(size_t)(4 * g + j2) * HW] = v[j2];
        }
    }
}

inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

struct BwdWorkspace {
    size_t ref_packed, partial, bbox, gbox, flags, tables, total;
    int pairs_per_pass;      // (view, batch) pairs whose cell tables are resident at once
    bool cells;              // the cell-table path applies (ids pack y < 2^15, x < 2^16)
};

#ifndef TMVS_BWD_TABLE_BYTES
#define TMVS_BWD_TABLE_BYTES (3ull << 30)      // cap of the cell-table workspace; at least one pair always fits
#endif

inline BwdWorkspace bwd_layout(int B, int C, int D, int H, int W, int n_src, unsigned flags)
{
    const size_t HW = (size_t)H * W;
    const size_t n_tiles = (size_t)((W + kTX - 1) / kTX) * ((H + kTY - 1) / kTY);
    BwdWorkspace ws;
    ws.ref_packed = 0;
    ws.partial = align256((size_t)B * tmvs_packed_layout((C + 3) / 4, H, W).slice * 16);
    ws.bbox = ws.partial + align256((size_t)n_src * B * C * HW * 4);
    ws.gbox = ws.bbox + align256((size_t)n_src * B * n_tiles * D * 16);
    const size_t n_groups = (size_t)(((W + kTX - 1) / kTX + kGroup - 1) / kGroup) * (((H + kTY - 1) / kTY + kGroup - 1) / kGroup);
    ws.flags = ws.gbox + align256((size_t)n_src * B * n_groups * 16);
    ws.tables = ws.flags + align256((size_t)(4 + n_tiles) * n_src * B * sizeof(int));
    ws.cells = H <= 32767 && W <= 65535;
    const size_t per_pair = tmvs_bwd_cells_bytes_per_pair(D, H, W);
    const int b_group = B < TMVS_GEOM_SLOTS / n_src ? B : TMVS_GEOM_SLOTS / n_src;
    // TMVS_F_TABLE_MB(mb) in the call's flags overrides the cap, e.g. to exercise the multi-pass path on small inputs
    const size_t cap = (flags >> 16) ? (size_t)(flags >> 16) << 20 : (size_t)TMVS_BWD_TABLE_BYTES;
    size_t pairs = cap / per_pair;
    if (pairs < 1) pairs = 1;
    if (pairs > (size_t)n_src * b_group) pairs = (size_t)n_src * b_group;
    ws.pairs_per_pass = (int)pairs;
    ws.total = ws.tables + (ws.cells ? pairs * per_pair : 0);
    return ws;
}

template <bool PER_PIXEL>
int launch_bwd(int c4, bool want_ref, bool want_src, dim3 grid, cudaStream_t st, const float4 *packed,
               const float4 *refp, const float *depth, const float *G, float *partial, int4 *bbox, int4 *gbox,
               float *grad_src, int b_total, int b_first, int b_chunk, int C, int D, int H, int W, int n_tx,
               int n_tiles, const int *gate, const int *gate_tile, const TmvsGeom &geom)
{
    dim3 block(kTX, kTY);
    const int n_ty = n_tiles / n_tx;
    const int n_gx = (n_tx + kGroup - 1) / kGroup, n_gy = (n_ty + kGroup - 1) / kGroup, n_groups = n_gx * n_gy;
#define TMVS_BWD(C4T, EX)                                                                                          \
    do {                                                                                                           \
        if (want_ref) {                                                                                            \
            bwd_ref_kernel<C4T, EX, PER_PIXEL><<<grid, block, 0, st>>>(packed, depth, G, partial, b_total, b_first, \
                                                                       b_chunk, C, c4, D, H, W, geom);             \
        }                                                                                                          \
        if (want_src) {                                                                                            \
            bwd_bbox_kernel<PER_PIXEL><<<grid, block, 0, st>>>(depth, bbox, b_first, b_chunk, D, H, W, n_tx,       \
                                                               n_tiles, gate, geom);                               \
            bwd_gbox_kernel<<<dim3(n_groups, grid.z), block, 0, st>>>(bbox, gbox, D, n_tx, n_ty, n_gx, n_tiles,    \
                                                                      n_groups, gate);                             \
            bwd_src_kernel<C4T, EX, PER_PIXEL><<<grid, block, 0, st>>>(refp, depth, G, bbox, gbox, grad_src,       \
                                                                       b_total, b_first, b_chunk, C, c4, D, H, W,  \
                                                                       n_tx, n_ty, n_tiles, n_gx, n_groups,        \
                                                                       gate_tile, geom);                           \
        }                                                                                                          \
    } while (0)
    if (c4 == 2) TMVS_BWD(2, true);
    else if (c4 == 4) TMVS_BWD(4, true);
    else if (c4 == 8) TMVS_BWD(8, true);
    else if (c4 < 4) TMVS_BWD(4, false);
    else if (c4 < 8) TMVS_BWD(8, false);
    else TMVS_BWD(16, false);
#undef TMVS_BWD
    return tmvs_launch_status();
}

}  // namespace

extern "C" size_t tmvs_costvol_bwd_workspace_bytes(int B, int C, int D, int H, int W, int n_src, unsigned flags)
{
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || n_src <= 0) return 0;
    return bwd_layout(B, C, D, H, W, n_src, flags).total;
}

extern "C" int tmvs_costvol_bwd(const float *ref, int64_t rB, int64_t rC, int64_t rH, int64_t rW, const float *packed,
                                const float *rot_trans, const float *depth, int per_pixel, const float *grad_views,
                                float *grad_ref, float *grad_src, void *workspace, size_t workspace_bytes, int B,
                                int C, int D, int H, int W, int n_src, unsigned flags, tmvs_stream_t stream)
{
    if (!ref || !packed || !rot_trans || !depth || !grad_views || !workspace) return TMVS_E_NULL;
    if (!grad_ref && !grad_src) return TMVS_E_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || n_src <= 0) return TMVS_E_SHAPE;
    if (n_src > TMVS_MAX_SRC_VIEWS || D > TMVS_MAX_DEPTH || C > 64) return TMVS_E_SHAPE;
    if ((size_t)H * (W + 7) * ((C + 3) / 4) > 0x7fffffffu) return TMVS_E_SHAPE;
    if (((uintptr_t)packed & 15) != 0 || ((uintptr_t)workspace & 15) != 0) return TMVS_E_ALIGN;
    const BwdWorkspace ws = bwd_layout(B, C, D, H, W, n_src, flags);
    if (workspace_bytes < ws.total) return TMVS_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    char *wsp = (char *)workspace;
    float *refp = (float *)(wsp + ws.ref_packed);
    float *partial = (float *)(wsp + ws.partial);
    int4 *bbox = (int4 *)(wsp + ws.bbox);
    int4 *gbox = (int4 *)(wsp + ws.gbox);
    int *wflags = (int *)(wsp + ws.flags);
    int *overflow = wflags + 3 * (size_t)n_src * B;       // one per (view, batch) pair: some tile of it needs the tile scan
    int *tile_overflow = overflow + (size_t)n_src * B;   // one per (pair, 32x8 source tile): that tile needs the tile scan
    // TMVS_F_BWD_SCAN forces the tile-scan kernels (the robust path the cell tables fall back to)
    const bool use_cells = grad_src && ws.cells && !(flags & TMVS_F_BWD_SCAN);
    const int c4 = (C + 3) / 4;
    const int n_tx = (W + kTX - 1) / kTX, n_ty = (H + kTY - 1) / kTY, n_tiles = n_tx * n_ty;
    const size_t HW = (size_t)H * W;
    if (grad_src) {   // reference features in the packed layout: the scatter's "value" operand
        const float *one[1] = {ref};
        int rc = tmvs_pack_sources(one, 1, rB, rC, rH, rW, refp, B, C, H, W, flags & TMVS_F_PACK_LDG, stream);
        if (rc != TMVS_OK) return rc;
        if (use_cells) {
            cudaError_t e = cudaMemsetAsync(wflags, 0, (size_t)(4 + n_tiles) * n_src * B * sizeof(int), st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    const int b_per_launch = TMVS_GEOM_SLOTS / n_src;
    for (int b0 = 0; b0 < B; b0 += b_per_launch) {
        const int bc = (B - b0 < b_per_launch) ? B - b0 : b_per_launch;
        TmvsGeom geom;
        tmvs_geom_fill(geom, rot_trans, flags, n_src, B, b0, bc);
        dim3 grid(n_tx, n_ty, n_src * bc);
        int rc;
        if (use_cells) {
            // grad_src through the cell tables; the tile-scan kernels below then run only if a cell overflowed
            rc = tmvs_bwd_src_cells((const float4 *)refp, depth, per_pixel, grad_views, grad_src, wsp + ws.tables,
                                    ws.pairs_per_pass, wflags + 3 * (size_t)n_src * b0, overflow + (size_t)n_src * b0,
                                    tile_overflow + (size_t)n_src * b0 * n_tiles, B, b0, bc, n_src, C, D,
                                    H, W, geom, st);
            if (rc != TMVS_OK) return rc;
        }
        const int *gate = use_cells ? overflow + (size_t)n_src * b0 : nullptr;
        const int *gate_tile = use_cells ? tile_overflow + (size_t)n_src * b0 * n_tiles : nullptr;
        // the bbox table of this launch is indexed by blockIdx.z = i * bc + bl
        if (per_pixel)
            rc = launch_bwd<true>(c4, grad_ref != nullptr, grad_src != nullptr, grid, st, (const float4 *)packed,
                                  (const float4 *)refp, depth, grad_views, partial, bbox, gbox, grad_src, B, b0, bc, C, D, H,
                                  W, n_tx, n_tiles, gate, gate_tile, geom);
        else
            rc = launch_bwd<false>(c4, grad_ref != nullptr, grad_src != nullptr, grid, st, (const float4 *)packed,
                                   (const float4 *)refp, depth, grad_views, partial, bbox, gbox, grad_src, B, b0, bc, C, D, H,
                                   W, n_tx, n_tiles, gate, gate_tile, geom);
        if (rc != TMVS_OK) return rc;
    }
    if (grad_ref) {
        const size_t n = (size_t)B * C * HW;
        bwd_ref_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, grad_ref, n, n_src);
        int rc = tmvs_launch_status();
        if (rc != TMVS_OK) return rc;
    }
    return TMVS_OK;
}

// ------------------------------------------------------------------------------------- drop-in warp backward
namespace {

struct WarpBwdWorkspace {
    size_t ones, bbox, gbox, flags, tables, total;
};

inline WarpBwdWorkspace warp_bwd_layout(int D, int H, int W)
{
    const size_t n_tiles = (size_t)((W + kTX - 1) / kTX) * ((H + kTY - 1) / kTY);
    const size_t n_groups = (size_t)(((W + kTX - 1) / kTX + kGroup - 1) / kGroup) * (((H + kTY - 1) / kTY + kGroup - 1) / kGroup);
    WarpBwdWorkspace ws;
    ws.ones = 0;
    ws.bbox = align256(tmvs_packed_layout(1, H, W).slice * 16);
    ws.gbox = ws.bbox + align256(n_tiles * D * 16);
    ws.flags = ws.gbox + align256(n_groups * 16);
    ws.tables = ws.flags + align256((4 + n_tiles) * sizeof(int));
    ws.total = ws.tables + tmvs_bwd_cells_bytes_per_pair(D, H, W);
    return ws;
}

__global__ void __launch_bounds__(256) fill_ones_kernel(float4 *p, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = make_float4(1.f, 1.f, 1.f, 1.f);
}

}  // namespace

extern "C" size_t tmvs_homo_warp_bwd_workspace_bytes(int B, int C, int D, int H, int W)
{
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return warp_bwd_layout(D, H, W).total;
}

extern "C" int tmvs_homo_warp_bwd(const float *rot_trans, const float *depth, int per_pixel, const float *grad_out,
                                  float *grad_src, void *workspace, size_t workspace_bytes, int B, int C, int D, int H,
                                  int W, unsigned flags, tmvs_stream_t stream)
{
    if (!rot_trans || !depth || !grad_out || !grad_src || !workspace) return TMVS_E_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return TMVS_E_SHAPE;
    if (D > TMVS_MAX_DEPTH || C > 64 || H > 32767 || W > 65535) return TMVS_E_SHAPE;
    if (((uintptr_t)workspace & 15) != 0) return TMVS_E_ALIGN;
    const WarpBwdWorkspace ws = warp_bwd_layout(D, H, W);
    if (workspace_bytes < ws.total) return TMVS_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    char *wsp = (char *)workspace;
    float4 *ones = (float4 *)(wsp + ws.ones);
    int4 *bbox = (int4 *)(wsp + ws.bbox), *gbox = (int4 *)(wsp + ws.gbox);
    int *wflags = (int *)(wsp + ws.flags);                  // [0..2] level flags, [3] pair overflow, [4..] per-tile overflow
    const int n_tx = (W + kTX - 1) / kTX, n_ty = (H + kTY - 1) / kTY, n_tiles = n_tx * n_ty;
    const int n_gx = (n_tx + kGroup - 1) / kGroup, n_gy = (n_ty + kGroup - 1) / kGroup, n_groups = n_gx * n_gy;
    const size_t HW = (size_t)H * W;
    const size_t n_ones = tmvs_packed_layout(1, H, W).slice;
    fill_ones_kernel<<<(unsigned)((n_ones + 255) / 256), 256, 0, st>>>(ones, n_ones);
    dim3 block(kTX, kTY), grid(n_tx, n_ty, 1);
    for (int b = 0; b < B; ++b) {
        TmvsGeom geom;
        tmvs_geom_fill(geom, rot_trans, flags, 1, B, b, 1);
        cudaError_t e = cudaMemsetAsync(wflags, 0, (size_t)(4 + n_tiles) * sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
        int *overflow = wflags + 3, *tile_overflow = wflags + 4;
        const float *gout_b = grad_out + (size_t)b * C * D * HW;
        float *gsrc_b = grad_src + (size_t)b * C * HW;
        int rc = tmvs_bwd_warp_cells(depth, per_pixel, gout_b, gsrc_b, wsp + ws.tables, wflags, overflow, tile_overflow, b,
                                     C, D, H, W, geom, st);
        if (rc != TMVS_OK) return rc;
        // Tiles touched by a footprint that found no slot (folded / strongly minified geometry): the tile-scan kernels
        // redo exactly those, one channel at a time with ref := 1, C := 1 (k * 1 = the tap weight times the gradient).
        // Every launch below returns at once when nothing overflowed (device-side gates; no host synchronisation).
        const float *depth_b = per_pixel ? depth + (size_t)b * D * HW : depth + (size_t)b * D;
        if (per_pixel) bwd_bbox_kernel<true><<<grid, block, 0, st>>>(depth_b, bbox, 0, 1, D, H, W, n_tx, n_tiles, overflow, geom);
        else bwd_bbox_kernel<false><<<grid, block, 0, st>>>(depth_b, bbox, 0, 1, D, H, W, n_tx, n_tiles, overflow, geom);
        bwd_gbox_kernel<<<dim3(n_groups, 1), block, 0, st>>>(bbox, gbox, D, n_tx, n_ty, n_gx, n_tiles, n_groups, overflow);
        for (int c = 0; c < C; ++c) {
            const float *g_c = gout_b + (size_t)c * D * HW;
            float *o_c = gsrc_b + (size_t)c * HW;
            if (per_pixel)
                bwd_src_kernel<4, false, true><<<grid, block, 0, st>>>(ones, depth_b, g_c, bbox, gbox, o_c, 1, 0, 1, 1, 1, D, H,
                                                                       W, n_tx, n_ty, n_tiles, n_gx, n_groups,
                                                                       tile_overflow, geom);
            else
                bwd_src_kernel<4, false, false><<<grid, block, 0, st>>>(ones, depth_b, g_c, bbox, gbox, o_c, 1, 0, 1, 1, 1, D, H,
                                                                        W, n_tx, n_ty, n_tiles, n_gx, n_groups,
                                                                        tile_overflow, geom);
        }
        rc = tmvs_launch_status();
        if (rc != TMVS_OK) return rc;
    }
    return TMVS_OK;
}
