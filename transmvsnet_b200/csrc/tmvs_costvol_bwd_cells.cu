// grad_src (the grid_sample scatter of models/module.py:318-320 under autograd) as a gather through a global
// cell table -- deterministic, no atomics of any kind, every output element written once.
//
// ATen scatters k * ref[:, p] from every (reference pixel p, depth plane d) into the four source pixels of p's
// bilinear footprint with float atomicAdd (cuda/GridSampler.cuh:250-260).  Here the relation is inverted once per
// (view, batch item, plane):
//
//   1. register  one thread per (p, d): the reference's coordinate arithmetic (tmvs_coords), the sample position is
//                kept in pos[d][p], and p's id (y << 16 | x) is stored in the cell of its footprint's north-west
//                source pixel, in the slot given by the PARITY of p (x & 1, y & 1).  Two reference pixels of the
//                same parity land in the same cell only where the view is minified more than 2x or the hypothesis
//                surface folds over, so in the common case every footprint owns its slot after one plain store.
//   2. fix-up    (p, d) pairs that lost their slot to another pixel of the same parity (detected by reading the slot
//                back in the next kernel -- the kernel boundary is the only synchronisation) move to overflow slot
//                1 and are listed, compacted per CTA; the next level walks the lists and moves the losers of slot 1
//                to overflow slot 2; whoever is still homeless flags the 32x8 source tiles its footprint touches,
//                and exactly those tiles are redone by the tile-scan kernels of tmvs_costvol_bwd.cu
//                (bwd_src_kernel), which handle any multiplicity.  Each level returns at once when the level before
//                it had no loser.
//   3. gather    one thread OWNS one source pixel q of one view: for d = 0..D-1, for the four tap classes
//                (nw, ne, sw, se), it reads the cell whose footprints hit q with that tap, sorts the <= 6 ids, and
//                accumulates k * ref[:, p] in (plane, rank of the id within its cell, class) order from pos[d][p],
//                G[d][p] and the packed reference features.  WHICH thread wins a slot never matters: the gather
//                orders ids itself.
//
// Work is O(voxel-views) with no barrier and no re-projection per overlapping tile (the tile-scan kernel re-projects
// every (tile, plane) once per source tile its box overlaps, 4-6x, between block-wide barriers).
#include "tmvs_common.cuh"

namespace {

constexpr int kTX = 32, kTY = 8;
constexpr int kRegDC = 8;                       // planes per thread in the register / fix-up kernels
constexpr int kListCap = 128;                   // losers a register CTA (256 pixels x 8 planes) can list; more -> full pass
constexpr unsigned kEmptyId = 0xffffffffu;      // memset(0xff); ids are (y << 16 | x) < 2^31 (H <= 32767)

struct CellTables {
    uint4 *par;         // [nvb][D][ncell]  parity slots (x&1 | (y&1)<<1)
    uint2 *ovf;         // [nvb][D][ncell]  overflow slots 1, 2
    float2 *pos;        // [nvb][D][HW]     sample position (ix, iy) of (p, d)
    int *flags;         // [0] losers after the parity level, [1] losers after overflow slot 1, [2] some CTA's loser list overflowed
    uint2 *list;        // [register CTA][kListCap] losers of the parity level: (pixel, plane | pair << 16)
    int *list_count;    // [register CTA]
    int *overflow;      // [pair z of the launch group] raised when a footprint of that (view, batch) pair found no slot
    int *tile_overflow; // [pair z][32x8 source tile]: the tiles such a footprint touches -- the caller's tile-scan fallback
                        // redoes exactly those tiles, the gather below skips them
    int n_tx, n_tiles;
};

// Registration stores are RELAXED (morally strong) device-scope stores: several footprints may store to the same slot
// and exactly one value survives -- plain STG in SASS, no read-modify-write -- and, unlike weak stores, concurrent
// relaxed stores are not a data race in the PTX memory model.  They are read back only by a LATER kernel.
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned cell_of(const TmvsTaps &t, int W) { return (unsigned)((t.y0 + 1) * (W + 1) + (t.x0 + 1)); }

template <bool PER_PIXEL>
__global__ void __launch_bounds__(kTX * kTY)
cells_register_kernel(const float *__restrict__ depth, CellTables tb, int z0, int b_first, int b_chunk, int D, int H,
                      int W, int n_dchunks, const __grid_constant__ TmvsGeom geom)
{
    const int x = blockIdx.x * kTX + threadIdx.x, y = blockIdx.y * kTY + threadIdx.y;
    if (x >= W || y >= H) return;
    const int zl = blockIdx.z / n_dchunks, chunk = blockIdx.z - zl * n_dchunks;   // zl: (view, batch) pair of this chunk
    const int z = z0 + zl;
    const int bl = z % b_chunk, b = b_first + bl;
    const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
    const size_t ncell = (size_t)(H + 1) * (W + 1);
    float rt[12];
    tmvs_geom_rt(geom, z / b_chunk, bl, b_chunk, rt);
    const TmvsRay ray = tmvs_ray(rt, (float)x, (float)y, geom.ray_unfused);
    const TmvsDims dims = tmvs_dims(H, W, geom.arith);
    const unsigned id = ((unsigned)y << 16) | (unsigned)x;
    const int cls = (x & 1) | ((y & 1) << 1);
    const int d1 = min(D, (chunk + 1) * kRegDC);
    for (int d = chunk * kRegDC; d < d1; ++d) {
        const float dep = PER_PIXEL ? __ldg(depth + ((size_t)b * D + d) * HW + pix) : __ldg(depth + (size_t)b * D + d);
        const float2 c = tmvs_coords(ray, rt, dep, dims);
        const TmvsTaps t = tmvs_footprint(c.x, c.y, dims);
        const size_t plane = (size_t)zl * D + d;
        tb.pos[plane * HW + pix] = c;
        TMVS_ASSERT(!t.any || cell_of(t, W) < ncell);
        if (t.any) st_relaxed_u32(reinterpret_cast<unsigned *>(tb.par + plane * ncell + cell_of(t, W)) + cls, id);
    }
}

// LEVEL 1: losers of the parity slot -> overflow slot 1.  LEVEL 2: losers of slot 1 -> slot 2.  LEVEL 3: anybody
// still without a slot raises the overflow flag.
template <int LEVEL>
__global__ void __launch_bounds__(kTX * kTY)
cells_fixup_kernel(CellTables tb, int z0, int D, int H, int W, int n_dchunks)
{
    if (LEVEL >= 2 && tb.flags[LEVEL - 2] == 0) return;      // the level before had no loser: nothing to move
    if (LEVEL == 2 && tb.flags[2] == 0) return;              // every CTA's losers fit its list: cells_fixup_list_kernel did it
    if (LEVEL == 2 && tb.list_count[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] != -1)
        return;                                              // this CTA's did
    __shared__ unsigned warp_cnt[kTY];
    const int x = blockIdx.x * kTX + threadIdx.x, y = blockIdx.y * kTY + threadIdx.y;
    const bool valid = x < W && y < H;
    const int zl = blockIdx.z / n_dchunks, chunk = blockIdx.z - zl * n_dchunks;
    const size_t HW = (size_t)H * W, pix = (size_t)y * W + x;
    const size_t ncell = (size_t)(H + 1) * (W + 1);
    const TmvsDims dims = tmvs_dims(H, W, 0);               // footprint only: the arithmetic mode plays no part
    const unsigned id = ((unsigned)y << 16) | (unsigned)x;
    const int cls = (x & 1) | ((y & 1) << 1);
    const int d_first = chunk * kRegDC;
    const int d1 = min(D, d_first + kRegDC);
    unsigned lost = 0;                                       // LEVEL 1: planes of this chunk where the pixel lost its slot
    for (int d = d_first; valid && d < d1; ++d) {
        const size_t plane = (size_t)zl * D + d;
        const float2 c = tb.pos[plane * HW + pix];
        const TmvsTaps t = tmvs_footprint(c.x, c.y, dims);
        if (!t.any) continue;
        TMVS_ASSERT(cell_of(t, W) < ncell);
        const size_t cell = plane * ncell + cell_of(t, W);
        if (reinterpret_cast<const unsigned *>(tb.par + cell)[cls] == id) continue;
        if (LEVEL == 1) {
            st_relaxed_u32(&tb.ovf[cell].x, id);
            st_relaxed_u32(reinterpret_cast<unsigned *>(tb.flags), 1u);
            lost |= 1u << (d - d_first);
            continue;
        }
        if (tb.ovf[cell].x == id) continue;
        if (LEVEL == 2) {
            st_relaxed_u32(&tb.ovf[cell].y, id);
            st_relaxed_u32(reinterpret_cast<unsigned *>(tb.flags + 1), 1u);
            continue;
        }
        if (tb.ovf[cell].y != id) {
            // homeless: flag the pair and every source tile the footprint touches
            st_relaxed_u32(reinterpret_cast<unsigned *>(tb.overflow + z0 + zl), 1u);
            int *tf = tb.tile_overflow + (size_t)(z0 + zl) * tb.n_tiles;
            const int xa = max(t.x0, 0) / kTX, xb = min(t.x0 + 1, W - 1) / kTX;
            const int ya = max(t.y0, 0) / kTY, yb = min(t.y0 + 1, H - 1) / kTY;
            st_relaxed_u32(reinterpret_cast<unsigned *>(tf + ya * tb.n_tx + xa), 1u);
            st_relaxed_u32(reinterpret_cast<unsigned *>(tf + ya * tb.n_tx + xb), 1u);
            st_relaxed_u32(reinterpret_cast<unsigned *>(tf + yb * tb.n_tx + xa), 1u);
            st_relaxed_u32(reinterpret_cast<unsigned *>(tf + yb * tb.n_tx + xb), 1u);
        }
    }
    if (LEVEL == 1) {
        // the losers of this CTA, compacted in (thread, plane) order into its list: the next level then visits only them
        const unsigned mine = __popc(lost);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += v;
        }
        if (threadIdx.x == 31) warp_cnt[threadIdx.y] = incl;
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kTY; ++w) {
            const unsigned c = warp_cnt[w];
            if (w < (int)threadIdx.y) before += c;
            total += c;
        }
        const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        if (total > (unsigned)kListCap) {
            if (threadIdx.x == 0 && threadIdx.y == 0) {
                tb.list_count[cta] = -1;                                            // too many: level 2 revisits
                st_relaxed_u32(reinterpret_cast<unsigned *>(tb.flags + 2), 1u);     // this CTA's pixels in full
            }
        } else {
            if (threadIdx.x == 0 && threadIdx.y == 0) tb.list_count[cta] = (int)total;
            unsigned k = before + incl - mine;
            uint2 *out = tb.list + (size_t)cta * kListCap;
            while (lost) {
                const int j = __ffs(lost) - 1;
                lost &= lost - 1;
                out[k++] = make_uint2((unsigned)pix, (unsigned)(d_first + j) | ((unsigned)zl << 16));
            }
        }
    }
}

// Level 2 over the loser lists: losers of overflow slot 1 move to slot 2.
__global__ void __launch_bounds__(128)
cells_fixup_list_kernel(CellTables tb, int D, int H, int W)
{
    if (tb.flags[0] == 0) return;                            // no loser at all
    const int n = tb.list_count[blockIdx.x];                 // -1: listed too many, the full-pass kernel revisits that CTA
    const size_t HW = (size_t)H * W;
    const size_t ncell = (size_t)(H + 1) * (W + 1);
    const TmvsDims dims = tmvs_dims(H, W, 0);
    const uint2 *in = tb.list + (size_t)blockIdx.x * kListCap;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const uint2 ent = in[e];
        const unsigned pix = ent.x, d = ent.y & 0xffffu, zl = ent.y >> 16;
        const unsigned y = pix / (unsigned)W, x = pix - y * (unsigned)W;
        const size_t plane = (size_t)zl * D + d;
        const float2 c = tb.pos[plane * HW + pix];
        const TmvsTaps t = tmvs_footprint(c.x, c.y, dims);
        const size_t cell = plane * ncell + cell_of(t, W);
        const unsigned id = (y << 16) | x;
        if (tb.ovf[cell].x == id) continue;
        st_relaxed_u32(&tb.ovf[cell].y, id);
        st_relaxed_u32(reinterpret_cast<unsigned *>(tb.flags + 1), 1u);
    }
}

// The gather walks D planes of tables that are far larger than L2; every plane costs three dependent trips through
// the memory system (cell -> position / gradient of the registrant -> reference features).  The cell addresses of the
// next plane are known in advance and its registrants sit next to this plane's, so both are requested from L2 one
// plane ahead (CCTL.E.PF2, no registers).  Measured (profiles/r2_tune_bwd_prefetch_*.json): 2 = cells + registrants
// -1.9 / -1.5 / -2.3 % per grad_src call at DTU size, 1 = cells only +-0, two planes ahead slower; 0 = off.
#ifndef TMVS_GATHER_PF
#define TMVS_GATHER_PF 2
#endif
#ifndef TMVS_GATHER_PFD
#define TMVS_GATHER_PFD 1                       // planes ahead
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void cswap_u(unsigned &a, unsigned &b)
{
    const unsigned lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

#ifndef TMVS_GATHER_MINB8
#define TMVS_GATHER_MINB8 2
#endif
#ifndef TMVS_GATHER_MINB4
#define TMVS_GATHER_MINB4 3
#endif
#ifndef TMVS_GATHER_MINB2
#define TMVS_GATHER_MINB2 3
#endif
template <int C4T> struct GatherMinBlocks { static constexpr int value = C4T >= 8 ? TMVS_GATHER_MINB8 : (C4T >= 4 ? TMVS_GATHER_MINB4 : TMVS_GATHER_MINB2); };

// The gather is a chain of dependent loads (cell -> position and gradient of the registered pixel -> its reference
// features).  Rank u of the four classes' id lists goes through the chain together (four independent chains in
// flight); ranks beyond the first exist only where several footprints share a north-west source pixel (local
// minification) and are skipped by warps that have none.  Summation order per output element: plane, then rank (ids
// ascending within a cell), then class nw, ne, sw, se -- a fixed function of the registered SET, so the result does
// not depend on which store won a slot.  Index arithmetic is 32-bit offsets from per-plane base pointers.
template <int C4T, bool EXACT>
__global__ void __launch_bounds__(kTX * kTY, GatherMinBlocks<C4T>::value)
cells_gather_kernel(const float4 *__restrict__ refp, const float *__restrict__ G, CellTables tb,
                    float *__restrict__ grad_src, int z0, int b_total, int b_first, int b_chunk, int C, int c4, int D,
                    int H, int W)
{
    const int zl = blockIdx.z, z = z0 + zl;
    // a tile touched by a footprint that found no slot is redone, whole, by the tile-scan fallback
    if (tb.tile_overflow[(size_t)z * tb.n_tiles + blockIdx.y * tb.n_tx + blockIdx.x] != 0) return;
    const int qx = blockIdx.x * kTX + threadIdx.x, qy = blockIdx.y * kTY + threadIdx.y;
    if (qx >= W || qy >= H) return;
    const int i = z / b_chunk, bl = z - i * b_chunk, b = b_first + bl;
    const unsigned HW = (unsigned)H * (unsigned)W;
    const unsigned ncell = (unsigned)(H + 1) * (unsigned)(W + 1);
    const bool has_ovf = tb.flags[0] != 0;
    const float inv_c = 1.0f / (float)C;
    const TmvsPacked pk = tmvs_packed_layout(c4, H, W);
    const float4 *rimg = refp + (size_t)b * pk.slice;
    const uint4 *par_d = tb.par + (size_t)zl * D * ncell;
    const uint2 *ovf_d = tb.ovf + (size_t)zl * D * ncell;
    const float2 *pos_d = tb.pos + (size_t)zl * D * HW;
    const float *g_d = G + ((size_t)i * b_total + b) * D * HW;
    float qxf = (float)qx, qyf = (float)qy;
    // north-west-corner cell of the footprints that reach q with tap class 0 (nw); classes 1 (ne), 2 (sw), 3 (se)
    // are the cells one to the left, one up, and both
    unsigned c0 = (unsigned)((qy + 1) * (W + 1) + qx + 1);
    const unsigned wp1 = (unsigned)(W + 1);
    asm volatile("" : "+f"(qxf), "+f"(qyf), "+r"(c0), "+l"(rimg));      // keep them in registers (no re-derivation per tap)

    float4 acc[C4T];
#pragma unroll
    for (int g = 0; g < C4T; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int d = 0; d < D; ++d, par_d += ncell, ovf_d += ncell, pos_d += HW, g_d += HW) {
        if (TMVS_GATHER_PF >= 1 && d + TMVS_GATHER_PFD < D) {   // a later plane's cells (class 1 / 3 share the lines of 0 / 2)
            const size_t ahead = (size_t)TMVS_GATHER_PFD * ncell;
            prefetch_l2(par_d + ahead + c0);
            prefetch_l2(par_d + ahead + c0 - wp1);
            if (has_ovf) {
                prefetch_l2(ovf_d + ahead + c0);
                prefetch_l2(ovf_d + ahead + c0 - wp1);
            }
        }
        // ids of every class in ascending order (empties last)
        unsigned ids[4][6];
#pragma unroll
        for (int cls = 0; cls < 4; ++cls) {
            const unsigned cell = c0 - ((cls & 1) ? 1u : 0u) - ((cls & 2) ? wp1 : 0u);
            TMVS_ASSERT(cell < ncell);
            const uint4 pr = __ldg(par_d + cell);
            unsigned *v = ids[cls];
            v[0] = pr.x; v[1] = pr.y; v[2] = pr.z; v[3] = pr.w; v[4] = kEmptyId; v[5] = kEmptyId;
            if (has_ovf) {                                   // uniform branch
                const uint2 ov = __ldg(ovf_d + cell);
                v[4] = ov.x; v[5] = ov.y;
            }
        }
#pragma unroll
        for (int cls = 0; cls < 4; ++cls) {
            unsigned *v = ids[cls];
            if (has_ovf) {                                   // 6-input sorting network
                cswap_u(v[0], v[5]); cswap_u(v[1], v[3]); cswap_u(v[2], v[4]);
                cswap_u(v[1], v[2]); cswap_u(v[3], v[4]);
                cswap_u(v[0], v[3]); cswap_u(v[2], v[5]);
                cswap_u(v[0], v[1]); cswap_u(v[2], v[3]); cswap_u(v[4], v[5]);
                cswap_u(v[1], v[2]); cswap_u(v[3], v[4]);
            } else {
                cswap_u(v[0], v[1]); cswap_u(v[2], v[3]); cswap_u(v[0], v[2]); cswap_u(v[1], v[3]); cswap_u(v[1], v[2]);
            }
        }
        if (TMVS_GATHER_PF >= 2 && d + 1 < D) {             // next plane's registrants are this plane's neighbours
            const unsigned id = min(min(ids[0][0], ids[1][0]), min(ids[2][0], ids[3][0]));
            if (id != kEmptyId) {
                const unsigned pix = (id >> 16) * (unsigned)W + (id & 0xffffu);
                prefetch_l2(pos_d + HW + pix);
                prefetch_l2(g_d + HW + pix);
            }
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (u >= 4 && !has_ovf) break;
            if (u > 0 && (ids[0][u] & ids[1][u] & ids[2][u] & ids[3][u]) == kEmptyId)
                break;                                       // sorted: no class has a rank-u registrant, nor any later one
            // k = G * weight of the tap that lands on q, and the packed address of the registered reference pixel.
            // The footprint's north-west corner is known from its cell -- (qx - dx, qy - dy) -- and the tap is q
            // itself, in bounds: the ATen corner weight needs no floor and no bounds test, only the two differences
            // it is the product of ((x0 + 1) - ix or ix - x0, likewise in y: the operations of tmvs_footprint).
            float k[4];
            unsigned roff[4];
#pragma unroll
            for (int cls = 0; cls < 4; ++cls) {
                const unsigned id = ids[cls][u];
                k[cls] = 0.0f;
                roff[cls] = 0;
                if (id != kEmptyId) {
                    const unsigned px = id & 0xffffu, py = id >> 16;
                    const unsigned pix = py * (unsigned)W + px;
                    TMVS_ASSERT(px < (unsigned)W && py < (unsigned)H);
                    const float2 c = __ldg(pos_d + pix);
                    const float gw = __ldg(g_d + pix) * inv_c;
                    const float wx = (cls & 1) ? __fsub_rn(c.x, qxf - 1.0f) : __fsub_rn(qxf + 1.0f, c.x);
                    const float wy = (cls & 2) ? __fsub_rn(c.y, qyf - 1.0f) : __fsub_rn(qyf + 1.0f, c.y);
                    k[cls] = gw * __fmul_rn(wx, wy);
                    roff[cls] = py * (unsigned)pk.row + (px >> 3) * (unsigned)pk.c4x8 + (px & 7u);
                }
            }
#pragma unroll
            for (int cls = 0; cls < 4; ++cls) {
                if (k[cls] != 0.0f) {
                    const float4 *rp = tmvs_pk_ptr(rimg, roff[cls]);
#pragma unroll
                    for (int g = 0; g < C4T; ++g) {
                        if (EXACT || g < c4) {
                            const float4 rv = ldg4(rp + g * 8);
                            acc[g].x = fmaf(k[cls], rv.x, acc[g].x);
                            acc[g].y = fmaf(k[cls], rv.y, acc[g].y);
                            acc[g].z = fmaf(k[cls], rv.z, acc[g].z);
                            acc[g].w = fmaf(k[cls], rv.w, acc[g].w);
                        }
                    }
                }
            }
        }
    }
    float *o = grad_src + (((size_t)i * b_total + b) * C) * HW + (size_t)qy * W + qx;
#pragma unroll
    for (int g = 0; g < C4T; ++g) {
        const float v[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * g + j < C) __stcs(o + (size_t)(4 * g + j) * HW, v[j]);
    }
}

// The same gather for the drop-in homo_warping's backward (autograd of F.grid_sample wrt the source features,
// models/module.py:318-320, for an ARBITRARY upstream gradient): the value scattered by (p, d) is the C-vector
// grad_out[b, :, d, p] instead of the rank-1  G[d, p] * ref[:, p]  of the fused cost volume, so the owner of source
// pixel q accumulates  w_tap * grad_out[b, c, d, p]  for every channel, in the same (plane, rank, class) order.
template <int CT>
__global__ void __launch_bounds__(kTX * kTY)
cells_gather_warp_kernel(const float *__restrict__ gout, CellTables tb, float *__restrict__ grad_src, int C, int D, int H,
                         int W)
{
    if (tb.tile_overflow[blockIdx.y * tb.n_tx + blockIdx.x] != 0) return;      // redone by the tile-scan fallback
    const int qx = blockIdx.x * kTX + threadIdx.x, qy = blockIdx.y * kTY + threadIdx.y;
    if (qx >= W || qy >= H) return;
    const unsigned HW = (unsigned)H * (unsigned)W;
    const unsigned ncell = (unsigned)(H + 1) * (unsigned)(W + 1);
    const bool has_ovf = tb.flags[0] != 0;
    const uint4 *par_d = tb.par;
    const uint2 *ovf_d = tb.ovf;
    const float2 *pos_d = tb.pos;
    const float *g_d = gout;                                  // [C][D][HW] of this batch item
    const size_t c_stride = (size_t)D * HW;
    const float qxf = (float)qx, qyf = (float)qy;
    const unsigned c0 = (unsigned)((qy + 1) * (W + 1) + qx + 1);
    const unsigned wp1 = (unsigned)(W + 1);
    float acc[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = 0.0f;
    for (int d = 0; d < D; ++d, par_d += ncell, ovf_d += ncell, pos_d += HW, g_d += HW) {
        unsigned ids[4][6];
#pragma unroll
        for (int cls = 0; cls < 4; ++cls) {
            const unsigned cell = c0 - ((cls & 1) ? 1u : 0u) - ((cls & 2) ? wp1 : 0u);
            TMVS_ASSERT(cell < ncell);
            const uint4 pr = __ldg(par_d + cell);
            unsigned *v = ids[cls];
            v[0] = pr.x; v[1] = pr.y; v[2] = pr.z; v[3] = pr.w; v[4] = kEmptyId; v[5] = kEmptyId;
            if (has_ovf) {
                const uint2 ov = __ldg(ovf_d + cell);
                v[4] = ov.x; v[5] = ov.y;
            }
            if (has_ovf) {
                cswap_u(v[0], v[5]); cswap_u(v[1], v[3]); cswap_u(v[2], v[4]);
                cswap_u(v[1], v[2]); cswap_u(v[3], v[4]);
                cswap_u(v[0], v[3]); cswap_u(v[2], v[5]);
                cswap_u(v[0], v[1]); cswap_u(v[2], v[3]); cswap_u(v[4], v[5]);
                cswap_u(v[1], v[2]); cswap_u(v[3], v[4]);
            } else {
                cswap_u(v[0], v[1]); cswap_u(v[2], v[3]); cswap_u(v[0], v[2]); cswap_u(v[1], v[3]); cswap_u(v[1], v[2]);
            }
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (u >= 4 && !has_ovf) break;
            if ((ids[0][u] & ids[1][u] & ids[2][u] & ids[3][u]) == kEmptyId) break;
#pragma unroll
            for (int cls = 0; cls < 4; ++cls) {
                const unsigned id = ids[cls][u];
                if (id == kEmptyId) continue;
                const unsigned px = id & 0xffffu, py = id >> 16;
                const unsigned pix = py * (unsigned)W + px;
                TMVS_ASSERT(px < (unsigned)W && py < (unsigned)H);
                const float2 c = __ldg(pos_d + pix);
                const float wx = (cls & 1) ? __fsub_rn(c.x, qxf - 1.0f) : __fsub_rn(qxf + 1.0f, c.x);
                const float wy = (cls & 2) ? __fsub_rn(c.y, qyf - 1.0f) : __fsub_rn(qyf + 1.0f, c.y);
                const float k = __fmul_rn(wx, wy);
                const float *gp = g_d + pix;
#pragma unroll
                for (int ch = 0; ch < CT; ++ch)
                    if (ch < C) acc[ch] = fmaf(k, __ldg(gp + (size_t)ch * c_stride), acc[ch]);
            }
        }
    }
    float *o = grad_src + (size_t)qy * W + qx;
#pragma unroll
    for (int ch = 0; ch < CT; ++ch)
        if (ch < C) o[(size_t)ch * HW] = acc[ch];
}

inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace

// bytes of cell tables + positions for one (view, batch item) pair
size_t tmvs_bwd_cells_bytes_per_pair(int D, int H, int W)
{
    const size_t ncell = (size_t)(H + 1) * (W + 1), HW = (size_t)H * W;
    const size_t ctas = (size_t)((W + kTX - 1) / kTX) * ((H + kTY - 1) / kTY) * ((D + kRegDC - 1) / kRegDC);
    return align256((size_t)D * ncell * (sizeof(uint4) + sizeof(uint2))) + align256((size_t)D * HW * sizeof(float2)) +
           align256(ctas * kListCap * sizeof(uint2)) + align256(ctas * sizeof(int)) + 256;   // + the loser lists and their counts
}

// grad_src of the (view, batch) pairs z = 0 .. n_src*bc-1 of one launch group (geom.rt[z], z = view * bc + bl),
// `pairs_per_pass` pairs at a time through the table workspace `tables` (pairs_per_pass * bytes_per_pair bytes).
// flags: 3 ints per pass; overflow: one int per pair of the group; tile_overflow: one int per (pair, 32x8 source tile)
// -- all zeroed by the caller.
int tmvs_bwd_src_cells(const float4 *refp, const float *depth, int per_pixel, const float *G, float *grad_src,
                       char *tables, int pairs_per_pass, int *flags, int *overflow, int *tile_overflow, int b_total,
                       int b_first, int bc,
                       int n_src, int C, int D, int H, int W, const TmvsGeom &geom, cudaStream_t st)
{
    const int c4 = (C + 3) / 4;
    const int n_tx = (W + kTX - 1) / kTX, n_ty = (H + kTY - 1) / kTY;
    const int n_dchunks = (D + kRegDC - 1) / kRegDC;
    const size_t ncell = (size_t)(H + 1) * (W + 1), HW = (size_t)H * W;
    const int n_pairs = n_src * bc;
    dim3 block(kTX, kTY);
    int pass = 0;
    for (int z0 = 0; z0 < n_pairs; z0 += pairs_per_pass, ++pass) {
        const int nz = n_pairs - z0 < pairs_per_pass ? n_pairs - z0 : pairs_per_pass;
        const size_t cell_bytes = align256((size_t)D * ncell * (sizeof(uint4) + sizeof(uint2)));
        CellTables tb;
        // all parity tables of the pass first, then all overflow tables, then the positions
        tb.par = reinterpret_cast<uint4 *>(tables);
        tb.ovf = reinterpret_cast<uint2 *>(tables + (size_t)nz * D * ncell * sizeof(uint4));
        tb.pos = reinterpret_cast<float2 *>(tables + (size_t)pairs_per_pass * cell_bytes);
        const size_t pos_bytes = align256((size_t)D * HW * sizeof(float2));
        const size_t n_ctas = (size_t)n_tx * n_ty * n_dchunks * nz;
        tb.list = reinterpret_cast<uint2 *>(tables + (size_t)pairs_per_pass * (cell_bytes + pos_bytes));
        tb.list_count = reinterpret_cast<int *>(tables + (size_t)pairs_per_pass * (cell_bytes + pos_bytes) +
                                                align256(n_ctas * kListCap * sizeof(uint2)));
        tb.flags = flags + 3 * pass;
        tb.overflow = overflow;
        tb.tile_overflow = tile_overflow;
        tb.n_tx = n_tx;
        tb.n_tiles = n_tx * n_ty;
        cudaError_t e = cudaMemsetAsync(tables, 0xff, (size_t)nz * D * ncell * (sizeof(uint4) + sizeof(uint2)), st);
        if (e != cudaSuccess) return (int)e;
        dim3 grid_pd(n_tx, n_ty, nz * n_dchunks);
        if (per_pixel)
            cells_register_kernel<true><<<grid_pd, block, 0, st>>>(depth, tb, z0, b_first, bc, D, H, W, n_dchunks, geom);
        else
            cells_register_kernel<false><<<grid_pd, block, 0, st>>>(depth, tb, z0, b_first, bc, D, H, W, n_dchunks, geom);
        cells_fixup_kernel<1><<<grid_pd, block, 0, st>>>(tb, z0, D, H, W, n_dchunks);
        cells_fixup_list_kernel<<<(unsigned)n_ctas, 128, 0, st>>>(tb, D, H, W);          // level 2 over the loser lists ...
        cells_fixup_kernel<2><<<grid_pd, block, 0, st>>>(tb, z0, D, H, W, n_dchunks);    // ... or in full if one overflowed
        cells_fixup_kernel<3><<<grid_pd, block, 0, st>>>(tb, z0, D, H, W, n_dchunks);
        dim3 grid_q(n_tx, n_ty, nz);
#define TMVS_GATHER(C4T, EX)                                                                                        \
        cells_gather_kernel<C4T, EX><<<grid_q, block, 0, st>>>(refp, G, tb, grad_src, z0, b_total, b_first, bc, C, c4, \
                                                               D, H, W)
        if (c4 == 2) { TMVS_GATHER(2, true); }
        else if (c4 == 4) { TMVS_GATHER(4, true); }
        else if (c4 == 8) { TMVS_GATHER(8, true); }
        else if (c4 < 4) { TMVS_GATHER(4, false); }
        else if (c4 < 8) { TMVS_GATHER(8, false); }
        else { TMVS_GATHER(16, false); }
#undef TMVS_GATHER
        const int rc = tmvs_launch_status();
        if (rc != TMVS_OK) return rc;
    }
    return TMVS_OK;
}

// grad_src of the drop-in warp for ONE batch item (geom.rt slot 0): registration + fix-up as above with a single pair,
// then the all-channel gather.  gout = grad_out[b] ([C][D][HW]), grad_src = out[b] ([C][HW]).  flags (3 ints),
// overflow (1 int) and tile_overflow (n_tiles ints) are zeroed by the caller; flagged tiles are left to its fallback.
int tmvs_bwd_warp_cells(const float *depth, int per_pixel, const float *gout, float *grad_src, char *tables, int *flags,
                        int *overflow, int *tile_overflow, int b, int C, int D, int H, int W, const TmvsGeom &geom,
                        cudaStream_t st)
{
    const int n_tx = (W + kTX - 1) / kTX, n_ty = (H + kTY - 1) / kTY;
    const int n_dchunks = (D + kRegDC - 1) / kRegDC;
    const size_t ncell = (size_t)(H + 1) * (W + 1), HW = (size_t)H * W;
    const size_t cell_bytes = align256((size_t)D * ncell * (sizeof(uint4) + sizeof(uint2)));
    const size_t pos_bytes = align256((size_t)D * HW * sizeof(float2));
    const size_t n_ctas = (size_t)n_tx * n_ty * n_dchunks;
    CellTables tb;
    tb.par = reinterpret_cast<uint4 *>(tables);
    tb.ovf = reinterpret_cast<uint2 *>(tables + (size_t)D * ncell * sizeof(uint4));
    tb.pos = reinterpret_cast<float2 *>(tables + cell_bytes);
    tb.list = reinterpret_cast<uint2 *>(tables + cell_bytes + pos_bytes);
    tb.list_count = reinterpret_cast<int *>(tables + cell_bytes + pos_bytes + align256(n_ctas * kListCap * sizeof(uint2)));
    tb.flags = flags;
    tb.overflow = overflow;
    tb.tile_overflow = tile_overflow;
    tb.n_tx = n_tx;
    tb.n_tiles = n_tx * n_ty;
    cudaError_t e = cudaMemsetAsync(tables, 0xff, (size_t)D * ncell * (sizeof(uint4) + sizeof(uint2)), st);
    if (e != cudaSuccess) return (int)e;
    dim3 block(kTX, kTY), grid_pd(n_tx, n_ty, n_dchunks);
    // one pair: z0 = 0, b_chunk = 1 -> the kernels' (view, batch) slot is 0 and their batch item is b_first = b
    if (per_pixel)
        cells_register_kernel<true><<<grid_pd, block, 0, st>>>(depth, tb, 0, b, 1, D, H, W, n_dchunks, geom);
    else
        cells_register_kernel<false><<<grid_pd, block, 0, st>>>(depth, tb, 0, b, 1, D, H, W, n_dchunks, geom);
    cells_fixup_kernel<1><<<grid_pd, block, 0, st>>>(tb, 0, D, H, W, n_dchunks);
    cells_fixup_list_kernel<<<(unsigned)n_ctas, 128, 0, st>>>(tb, D, H, W);
    cells_fixup_kernel<2><<<grid_pd, block, 0, st>>>(tb, 0, D, H, W, n_dchunks);
    cells_fixup_kernel<3><<<grid_pd, block, 0, st>>>(tb, 0, D, H, W, n_dchunks);
    dim3 grid_q(n_tx, n_ty, 1);
    if (C <= 8) cells_gather_warp_kernel<8><<<grid_q, block, 0, st>>>(gout, tb, grad_src, C, D, H, W);
    else if (C <= 16) cells_gather_warp_kernel<16><<<grid_q, block, 0, st>>>(gout, tb, grad_src, C, D, H, W);
    else if (C <= 32) cells_gather_warp_kernel<32><<<grid_q, block, 0, st>>>(gout, tb, grad_src, C, D, H, W);
    else cells_gather_warp_kernel<64><<<grid_q, block, 0, st>>>(gout, tb, grad_src, C, D, H, W);
    return tmvs_launch_status();
}
