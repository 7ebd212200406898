// edx_kernels.cuh — CUDA kernels of the raster hot path, sm_100a.
//
// Frame = geom_kernel -> clip_kernel -> tile_kernel:
//   geom_kernel  one thread per submitted triangle: vertex transform (a1), clip codes (a2), setup (a5),
//                exact culls; triangles whose pixel-centre box is <= smallMax rasterise immediately with
//                64-bit atomicMin into the L2-resident visibility-key buffer, larger ones are appended
//                to the tile-path list, straddlers go to the clip queue.
//   clip_kernel  Sutherland-Hodgman in clip space (a3/a4) for the queued straddlers, fan, setup, same routing.
//   tile_kernel  one CTA per 64x64 bin, one warp per 16x16 tile: stages the bin's keys in shared memory,
//                culls the large-triangle list against the bin (edge tests + hierarchical Z), rasterises
//                survivors 16x16 -> 8x8 -> pixel with ballot masks, then resolves every pixel (depth,
//                perspective-correct interpolation, Blinn-Phong) and writes the tile to HBM once.
#pragma once
#include "edx_device.cuh"

namespace edx {

// ---------------------------------------------------------------------------------------------
// Mesh upload: 32-byte AoS vertices -> two float4 streams; uint32 x 3 indices -> three streams
// (Utils/InputBuffer.h:16-28,148-194). Runs once per mesh, not per frame.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_vertices_kernel(const float4* __restrict__ aos, float4* __restrict__ pos4,
                                                             float4* __restrict__ nrm4, uint32_t nVerts)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVerts) return;
    float4 a = __ldg(aos + 2 * (size_t)i);          // px py pz nx
    float4 b = __ldg(aos + 2 * (size_t)i + 1);      // ny nz u  v
    pos4[i] = make_float4(a.x, a.y, a.z, b.w);
    nrm4[i] = make_float4(a.w, b.x, b.y, b.z);
}

__global__ void __launch_bounds__(256) split_indices_kernel(const uint32_t* __restrict__ idx, uint32_t* __restrict__ i0,
                                                            uint32_t* __restrict__ i1, uint32_t* __restrict__ i2, uint32_t nTris)
{
    // 256 triangles = 768 consecutive words per CTA, staged through shared memory so both the
    // global read and the three global writes are fully coalesced.
    __shared__ uint32_t s[768];
    size_t base = (size_t)blockIdx.x * 256;
    size_t words = 3 * (size_t)nTris;
    for (int k = threadIdx.x; k < 768; k += 256) {
        size_t w = base * 3 + k;
        s[k] = w < words ? __ldg(idx + w) : 0u;
    }
    __syncthreads();
    size_t t = base + threadIdx.x;
    if (t < nTris) {
        i0[t] = s[3 * threadIdx.x];
        i1[t] = s[3 * threadIdx.x + 1];
        i2[t] = s[3 * threadIdx.x + 2];
    }
}

// Object-space bounding box of each cluster of 256 consecutive triangles (= one geom_kernel CTA). Built once
// per mesh upload; lets a frame skip clusters that lie entirely outside one clip plane.
__global__ void __launch_bounds__(256) cluster_bounds_kernel(const float4* __restrict__ pos4, const uint32_t* __restrict__ i0,
                                                             const uint32_t* __restrict__ i1, const uint32_t* __restrict__ i2,
                                                             uint32_t nTris, float4* __restrict__ boxes)
{
    __shared__ float red[6][8];
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
    bool bad = false;
    if (t < nTris) {
        const uint32_t id[3] = { __ldg(i0 + t), __ldg(i1 + t), __ldg(i2 + t) };
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            const float4 p = __ldg(pos4 + id[k]);
            const float v[3] = { p.x, p.y, p.z };
            #pragma unroll
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], v[a]); hi[a] = fmaxf(hi[a], v[a]); bad |= !(fabsf(v[a]) < 3.0e38f); }
        }
    }
    #pragma unroll
    for (int a = 0; a < 3; a++)
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
    const bool anyBad = __syncthreads_or(bad);             // NaN / inf coordinates: never cull this cluster
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; a++) { red[a][threadIdx.x >> 5] = lo[a]; red[3 + a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            for (int a = 0; a < 3; a++) { red[a][0] = fminf(red[a][0], red[a][w]); red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]); }
        // w of the min box carries a validity flag
        boxes[2 * blockIdx.x] = make_float4(red[0][0], red[1][0], red[2][0], anyBad ? 0.0f : 1.0f);
        boxes[2 * blockIdx.x + 1] = make_float4(red[3][0], red[4][0], red[5][0], 0.0f);
    }
}

// The same for each block of 256 consecutive VERTICES (= one vertex_kernel work item of the list front end).
__global__ void __launch_bounds__(256) vertex_cluster_bounds_kernel(const float4* __restrict__ pos4, uint32_t nVerts, float4* __restrict__ boxes)
{
    __shared__ float red[6][8];
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
    bool bad = false;
    if (v < nVerts) {
        const float4 p = __ldg(pos4 + v);
        const float c[3] = { p.x, p.y, p.z };
        #pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = hi[a] = c[a]; bad |= !(fabsf(c[a]) < 3.0e38f); }
    }
    #pragma unroll
    for (int a = 0; a < 3; a++)
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
    const bool anyBad = __syncthreads_or(bad);
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < 3; a++) { red[a][threadIdx.x >> 5] = lo[a]; red[3 + a][threadIdx.x >> 5] = hi[a]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            for (int a = 0; a < 3; a++) { red[a][0] = fminf(red[a][0], red[a][w]); red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]); }
        boxes[2 * blockIdx.x] = make_float4(red[0][0], red[1][0], red[2][0], anyBad ? 0.0f : 1.0f);
        boxes[2 * blockIdx.x + 1] = make_float4(red[3][0], red[4][0], red[5][0], 0.0f);
    }
}

__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, size_t n, float v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256) fill_keys_kernel(ulonglong2* __restrict__ keys, size_t nPairs)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nPairs) keys[i] = make_ulonglong2(KEY_EMPTY, KEY_EMPTY);
}

// Stage a1 on its own (Renderer::VertexProcessing, Core/Renderer.cpp:120-127): SoA position stream in,
// clip-space float4 out. The frame path fuses this arithmetic into geom_kernel; this kernel serves
// edx_debug_clip_vertices (stage parity) and callers that want the projected stream.
__global__ void __launch_bounds__(256) vertex_transform_kernel(const __grid_constant__ FrameParams P, float4* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nVerts) return;
    float4 p = __ldg(P.pos4 + i);
    V4 c = to_clip(P.mvp, p.x, p.y, p.z);
    out[i] = make_float4(c.x, c.y, c.z, c.w);
}

// ---------------------------------------------------------------------------------------------
// Routing of one post-setup triangle (shared by geom_kernel and clip_kernel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter)
{
    // warp-aggregated atomic: one atomicAdd per group of lanes that reached this point together
    uint32_t mask = __activemask();
    uint32_t lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void route_triangle(const FrameParams& P, const SetupTri& s, float z0, float z1, float z2,
                                               float iw0, float iw1, float iw2, uint32_t prim, int smallMax)
{
    if (P.dump) {
        uint32_t at = atomicAdd(&P.counters->nDump, 1u);
        if (at < P.dumpCap) {
            DumpRec& d = P.dumpBuf[at];
            d.i[0] = (int)prim; d.i[1] = s.v0x; d.i[2] = s.v0y; d.i[3] = s.v1x; d.i[4] = s.v1y; d.i[5] = s.v2x; d.i[6] = s.v2y;
            d.f[0] = z0; d.f[1] = z1; d.f[2] = z2; d.f[3] = iw0; d.f[4] = iw1; d.f[5] = iw2; d.f[6] = s.invDet;
        }
    }
    // Pixel centres inside the snapped bounding box, clamped to the screen. A covered centre always
    // lies inside the box, so this is the same pixel set the reference visits (Rasterizer.h:132-137)
    // minus pixels that cannot be covered; an empty range is an exact cull.
    const bool ms = P.samples > 1;
    int x0 = max(0, first_pixel(min3i(s.v0x, s.v1x, s.v2x), ms));
    int x1 = min(P.width - 1, last_pixel(max3i(s.v0x, s.v1x, s.v2x), ms));
    int y0 = max(0, first_pixel(min3i(s.v0y, s.v1y, s.v2y), ms));
    int y1 = min(P.height - 1, last_pixel(max3i(s.v0y, s.v1y, s.v2y), ms));
    if (x0 > x1 || y0 > y1) return;

    if (x1 - x0 < smallMax && y1 - y0 < smallMax) {
        // Small triangle: rasterise here. Stages a10/a12/a13 per pixel:
        // coverage (Rasterizer.h:162), barycentrics + depth (RasterTriangle.h:324-337), depth test as key-min.
        Edges e;
        e.init(s.v0x, s.v0y, s.v1x, s.v1y, s.v2x, s.v2y);
        const int cx = (x0 << 4) + 8, cy = (y0 << 4) + 8;
        uint32_t r0 = (uint32_t)e.e0(cx, cy), r1 = (uint32_t)e.e1(cx, cy), r2 = (uint32_t)e.e2(cx, cy);
        const uint32_t sB0 = e.B0 << 4, sB1 = e.B1 << 4, sB2 = e.B2 << 4;   // one pixel = 16 sub-pixels (RasterTriangle.h:53-58)
        const uint32_t sC0 = e.C0 << 4, sC1 = e.C1 << 4, sC2 = e.C2 << 4;
        if (ms) {
            // Multi-sample (Rasterizer.h:202-300): every sample of every pixel in the box is an independent
            // coverage + depth test at centre + offset; its key goes to the sample's own key plane.
            const int* off = c_sampleOffsets[P.msLevel];
            for (int y = y0; y <= y1; y++) {
                uint32_t a0 = r0, a1 = r1, a2 = r2;
                for (int x = x0; x <= x1; x++) {
                    const uint32_t ki = key_index(x, y, P.binsX);
                    const bool owned = owns_pixel(x, y, P.binsX, P.part, P.parts);
                    for (int sId = 0; sId < P.samples; sId++) {
                        const uint32_t ox = (uint32_t)off[2 * sId], oy = ox;      // DESIGN.md shim 17: Vector2i(int) sets both components
                        const uint32_t f0 = a0 + ox * e.B0 + oy * e.C0, f1 = a1 + ox * e.B1 + oy * e.C1, f2 = a2 + ox * e.B2 + oy * e.C2;
                        if ((int)(f0 | f1 | f2) >= 0) {
                            float l0, l1;
                            barycentric((int)(f1 - (uint32_t)e.bias1), (int)(f2 - (uint32_t)e.bias2), s.invDet, l0, l1);
                            const float d = depth_at(l0, l1, z0, z1, z2);
                            if (d <= 1.0f && owned) atomicMin(P.keys + (size_t)sId * P.keyStride + ki, make_key(d, prim));
                        }
                    }
                    a0 += sB0; a1 += sB1; a2 += sB2;
                }
                r0 += sC0; r1 += sC1; r2 += sC2;
            }
        } else if (x1 - x0 < 8 && y1 - y0 < 8) {
            // Phase 1: coverage only, into a 64-bit mask (bit = 8*dy + dx). Phase 2: one iteration per
            // covered pixel. Splitting keeps the long depth/key/atomic sequence out of the divergent
            // coverage loop, so lanes of a warp execute it together instead of one at a time.
            const uint32_t base1 = r1 - (uint32_t)e.bias1, base2 = r2 - (uint32_t)e.bias2;   // unbiased edge 1 / 2 at (x0, y0)
            const int wm1 = x1 - x0;
            uint32_t lo = 0, hi = 0;
            if (wm1 < 5 && y1 - y0 < 5) {
                // Tiny box (the bulk of a dense mesh): evaluate a fixed 5x5 block of pixels branch-free and
                // mask off what lies outside the real box. Every lane of the warp executes the same ~130
                // instructions instead of the warp iterating max(height) x max(width) times.
                const uint32_t colMask = (1u << (wm1 + 1)) - 1u;
                const int hgt = y1 - y0 + 1;
                // Walk the fixed 5x5 block backwards: funnel-shifting each coverage bit (sign of the OR clear) in from
                // the right then leaves pixel (dx, dy) at bit 5*dy + dx - two instructions per pixel besides the adds.
                uint32_t q0 = r0 + 4u * sB0 + 4u * sC0, q1 = r1 + 4u * sB1 + 4u * sC1, q2 = r2 + 4u * sB2 + 4u * sC2;
                uint32_t m = 0;
                #pragma unroll
                for (int dy = 4; dy >= 0; dy--) {
                    uint32_t a0 = q0, a1 = q1, a2 = q2;
                    #pragma unroll
                    for (int dx = 4; dx >= 0; dx--) {
                        m = __funnelshift_l(~(a0 | a1 | a2), m, 1);
                        a0 -= sB0; a1 -= sB1; a2 -= sB2;
                    }
                    q0 -= sC0; q1 -= sC1; q2 -= sC2;
                }
                m &= (colMask * 0x00108421u) & ((1u << (5 * hgt)) - 1u);    // the real box: wm1 + 1 columns of hgt rows
                while (m) {
                    const uint32_t b = (uint32_t)__ffs(m) - 1u;
                    m &= m - 1u;
                    const uint32_t dy = (b * 13u) >> 6, dx = b - 5u * dy;       // b / 5 for b < 25
                    float l0, l1;
                    barycentric((int)(base1 + dx * sB1 + dy * sC1), (int)(base2 + dx * sB2 + dy * sC2), s.invDet, l0, l1);
                    const float d = depth_at(l0, l1, z0, z1, z2);
                    if (d <= 1.0f && owns_pixel(x0 + (int)dx, y0 + (int)dy, P.binsX, P.part, P.parts))
                        atomicMin(P.keys + key_index(x0 + (int)dx, y0 + (int)dy, P.binsX), make_key(d, prim));
                }
                return;
            }
            for (int y = y0; y <= y1; y++) {
                uint32_t a0 = r0, a1 = r1, a2 = r2, row = 0;
                for (int x = x0; x <= x1; x++) {
                    // shift the coverage bit (sign of the OR clear) in from the right: pixel dx ends at bit wm1 - dx
                    row = __funnelshift_l(~(a0 | a1 | a2), row, 1);
                    a0 += sB0; a1 += sB1; a2 += sB2;
                }
                const int dy = y - y0;
                if (dy < 4) lo |= row << (8 * dy); else hi |= row << (8 * (dy - 4));
                r0 += sC0; r1 += sC1; r2 += sC2;
            }
            while (lo | hi) {
                int b;
                if (lo) { b = __ffs(lo) - 1; lo &= lo - 1; } else { b = 32 + __ffs(hi) - 1; hi &= hi - 1; }
                const uint32_t dx = (uint32_t)(wm1 - (b & 7)), dy = (uint32_t)(b >> 3);
                float l0, l1;
                barycentric((int)(base1 + dx * sB1 + dy * sC1), (int)(base2 + dx * sB2 + dy * sC2), s.invDet, l0, l1);
                const float d = depth_at(l0, l1, z0, z1, z2);
                if (d <= 1.0f && owns_pixel(x0 + (int)dx, y0 + (int)dy, P.binsX, P.part, P.parts))     // depth buffer is cleared to 1.0 and tested LESS_EQUAL (FrameBuffer.cpp:64,103)
                    atomicMin(P.keys + key_index(x0 + (int)dx, y0 + (int)dy, P.binsX), make_key(d, prim));
            }
        } else {
            for (int y = y0; y <= y1; y++) {
                uint32_t a0 = r0, a1 = r1, a2 = r2;
                for (int x = x0; x <= x1; x++) {
                    if ((int)(a0 | a1 | a2) >= 0) {
                        float l0, l1;
                        barycentric((int)(a1 - (uint32_t)e.bias1), (int)(a2 - (uint32_t)e.bias2), s.invDet, l0, l1);
                        float d = depth_at(l0, l1, z0, z1, z2);
                        if (d <= 1.0f && owns_pixel(x, y, P.binsX, P.part, P.parts))
                            atomicMin(P.keys + key_index(x, y, P.binsX), make_key(d, prim));
                    }
                    a0 += sB0; a1 += sB1; a2 += sB2;
                }
                r0 += sC0; r1 += sC1; r2 += sC2;
            }
        }
    } else if (P.midLaunched && ((x1 - x0 < P.midMax && y1 - y0 < P.midMax) || !P.tileLaunched)) {
        // Mid-size triangle: a lone thread would serialise hundreds of pixel tests (and stall its warp), a whole-bin
        // sweep is too heavy a machine for it. mid_kernel gives it one warp. (A frame without tile_kernel - the previous
        // frame of this mesh had no large triangle - sends a large one here as well: correct at any size, slow for a
        // huge one, and it tells the host to bring the tile kernel back.)
        if (!P.tileLaunched && !(x1 - x0 < P.midMax && y1 - y0 < P.midMax)) atomicAdd(&P.counters->nBigDiverted, 1u);
        const uint32_t at = warp_append(&P.counters->nMid);
        if (at < P.midCap) {
            int4* dst = reinterpret_cast<int4*>(P.mid + at);
            dst[0] = make_int4(s.v0x, s.v0y, s.v1x, s.v1y);
            dst[1] = make_int4(s.v2x, s.v2y, __float_as_int(z0), __float_as_int(z1));
            dst[2] = make_int4(__float_as_int(z2), __float_as_int(s.invDet), (int)prim, 0);
        }
    } else {
        if (!P.midLaunched && x1 - x0 < P.midMax && y1 - y0 < P.midMax) atomicAdd(&P.counters->nMidDiverted, 1u);   // tells the host to launch mid_kernel again
        const uint32_t at = warp_append(&P.counters->nBig);
        if (at < P.bigCap) {
            // Depth plane over pixel centres (sub-pixel 16*p + 8), fitted in fp32 together with a bound of
            // its own rounding error; it only feeds the conservative hierarchical-Z test.
            const float B1 = (float)(s.v1y - s.v2y), C1 = (float)(s.v2x - s.v1x);      // |B|,|C| < 2^24: exact
            const float B2 = (float)(s.v2y - s.v0y), C2 = (float)(s.v0x - s.v2x);
            const float dz0 = z0 - z2, dz1 = z1 - z2;
            const float t0 = B1 * dz0, t1 = B2 * dz1, t2 = C1 * dz0, t3 = C2 * dz1;
            const float ax = (t0 + t1) * s.invDet, ay = (t2 + t3) * s.invDet;            // per sub-pixel
            const float eA = 4e-7f * (fabsf(t0) + fabsf(t1)) * s.invDet, eB = 4e-7f * (fabsf(t2) + fabsf(t3)) * s.invDet;
            const float ox = 8.0f - (float)s.v2x, oy = 8.0f - (float)s.v2y;
            const float zref = z2 + ax * ox + ay * oy;
            const float perr = eA * fabsf(ox) + eB * fabsf(oy) + 2.5e-7f * (fabsf(z2) + fabsf(ax * ox) + fabsf(ay * oy)) +
                               16.0f * (eA * (float)P.width + eB * (float)P.height);
            int4* dst = reinterpret_cast<int4*>(P.big + at);
            dst[0] = make_int4(s.v0x, s.v0y, s.v1x, s.v1y);
            dst[1] = make_int4(s.v2x, s.v2y, __float_as_int(z0), __float_as_int(z1));
            dst[2] = make_int4(__float_as_int(z2), __float_as_int(s.invDet), (int)prim, __float_as_int(16.0f * ax));
            dst[3] = make_int4(__float_as_int(16.0f * ay), __float_as_int(zref), __float_as_int(perr), 0);
            P.bigBox[at] = (uint32_t)(x0 >> BIN_LOG2) | ((uint32_t)(x1 >> BIN_LOG2) << 8) |
                           ((uint32_t)(y0 >> BIN_LOG2) << 16) | ((uint32_t)(y1 >> BIN_LOG2) << 24);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// geom_kernel: stages a1, a2, a5, a6 and the small-triangle part of a10-a13
// ---------------------------------------------------------------------------------------------
__device__ void clip_single_plane(const FrameParams& P, uint32_t t, const V4& c0, const V4& c1, const V4& c2, uint32_t planes);

// hand a straddler to clip_kernel: single-plane ones to the front half of the queue, the others to the back half
__device__ __forceinline__ void queue_straddler(const FrameParams& P, uint32_t t, const V4& c0, const V4& c1, const V4& c2, uint32_t planes)
{
    const uint32_t half = P.clipQueueCap / 2u;
    const bool single = __popc(planes) == 1;
    uint32_t at;
    if (single) { at = warp_append(&P.counters->nClipQueue); if (at >= half) return; }
    else { at = warp_append(&P.counters->nClipMulti); if (at >= P.clipQueueCap - half) return; at += half; }
    float4* q = reinterpret_cast<float4*>(P.clipQueue + at);
    q[0] = make_float4(c0.x, c0.y, c0.z, c0.w); q[1] = make_float4(c1.x, c1.y, c1.z, c1.w);
    q[2] = make_float4(c2.x, c2.y, c2.z, c2.w); q[3] = make_float4(__uint_as_float(t), __uint_as_float(planes), 0.0f, 0.0f);
}

// Stages a1, a2, a5, a6 for one submitted triangle with its vertex work done per corner (front end 0 / 1), then routing.
__device__ __forceinline__ void geom_triangle(const FrameParams& P, uint32_t t)
{
    // streaming loads (evict-first): the geometry is read once per frame and must not push the visibility keys, which
    // every path hits with atomics and the final pass reads back, out of the L2
    uint32_t i0 = __ldcs(P.i0 + t), i1 = __ldcs(P.i1 + t), i2 = __ldcs(P.i2 + t);
    float4 p0 = __ldcs(P.pos4 + i0), p1 = __ldcs(P.pos4 + i1), p2 = __ldcs(P.pos4 + i2);
    V4 c0 = to_clip(P.mvp, p0.x, p0.y, p0.z);
    V4 c1 = to_clip(P.mvp, p1.x, p1.y, p1.z);
    V4 c2 = to_clip(P.mvp, p2.x, p2.y, p2.z);
    if (!(surely_inside(c0) && surely_inside(c1) && surely_inside(c2))) {
        const uint32_t k0 = clip_code(c0), k1 = clip_code(c1), k2 = clip_code(c2);
        if (k0 | k1 | k2) {
        if (!(k0 & k1 & k2)) {                       // Clipper.h:107-109: straddles the frustum
            const uint32_t planes = (k0 ^ k1) | (k1 ^ k2) | (k2 ^ k0);
            if (P.fuseClip && __popc(planes) == 1) {
                clip_single_plane(P, t, c0, c1, c2, planes);     // one plane: clip right here, no queue round trip
            } else {
                queue_straddler(P, t, c0, c1, c2, planes);
            }
        }
        return;
        }
    }
    SetupTri s;
    if (!P.dump) {
        // Exact early cull (most sub-pixel triangles end here): snap first, and drop the triangle if its
        // box holds no pixel centre before paying for the det / invDet / invW arithmetic.
        project_snap(P.raster, P.rasterAffineXY != 0, c0, s.v0x, s.v0y);
        project_snap(P.raster, P.rasterAffineXY != 0, c1, s.v1x, s.v1y);
        project_snap(P.raster, P.rasterAffineXY != 0, c2, s.v2x, s.v2y);
        const bool ms = P.samples > 1;
        if (max(0, first_pixel(min3i(s.v0x, s.v1x, s.v2x), ms)) > min(P.width - 1, last_pixel(max3i(s.v0x, s.v1x, s.v2x), ms)) ||
            max(0, first_pixel(min3i(s.v0y, s.v1y, s.v2y), ms)) > min(P.height - 1, last_pixel(max3i(s.v0y, s.v1y, s.v2y), ms)))
            return;
        if (!finish_setup(s)) return;
    } else if (!setup_tri(P.raster, P.rasterAffineXY != 0, c0, c1, c2, s)) return;
    // Renderer.cpp:139-147: invW = 1/w, z = z * invW
    const float iw0 = inv_w(c0.w), iw1 = inv_w(c1.w), iw2 = inv_w(c2.w);
    route_triangle(P, s, fmul(c0.z, iw0), fmul(c1.z, iw1), fmul(c2.z, iw2), iw0, iw1, iw2, t * 8u, P.smallMax);
}

__global__ void __launch_bounds__(256, 6) geom_kernel(const __grid_constant__ FrameParams P)
{
    // Programmatic dependent launch: the grid may be scheduled while the previous kernel in the stream is
    // still draining; everything before this point touches no global memory.
    cudaGridDependencySynchronize();
    if (P.clusterCull) {
        // Cluster cull: if all 8 corners of this CTA's 256-triangle bounding box are outside the SAME clip plane
        // by a margin that dominates fp32 rounding of the transform, every vertex of the cluster is outside that
        // plane too, i.e. every triangle would be rejected by Clipper.h:109. Exact, and the CTA loads nothing else.
        // Eight lanes of warp 0 take one corner each; the verdict reaches the CTA through shared memory.
        __shared__ uint32_t sOut;
        if (threadIdx.x < 32) {
            const float4 bl = __ldg(P.clusterBox + 2 * blockIdx.x), bh = __ldg(P.clusterBox + 2 * blockIdx.x + 1);
            const float* M = P.mvp;
            // Largest magnitude any vertex of the box can reach in each row's sum: bounds the fp32 rounding error
            // (4 roundings, < 2.5e-7 of this) of EVERY vertex in the box; the margin is 16x that.
            const float ax = fmaxf(fabsf(bl.x), fabsf(bh.x)), ay = fmaxf(fabsf(bl.y), fabsf(bh.y)), az = fmaxf(fabsf(bl.z), fabsf(bh.z));
            const float mx = fabsf(M[0]) * ax + fabsf(M[1]) * ay + fabsf(M[2]) * az + fabsf(M[3]);
            const float my = fabsf(M[4]) * ax + fabsf(M[5]) * ay + fabsf(M[6]) * az + fabsf(M[7]);
            const float mz = fabsf(M[8]) * ax + fabsf(M[9]) * ay + fabsf(M[10]) * az + fabsf(M[11]);
            const float mw = fabsf(M[12]) * ax + fabsf(M[13]) * ay + fabsf(M[14]) * az + fabsf(M[15]);
            const float ex = 4e-6f * (mx + mw), ey = 4e-6f * (my + mw), ez = 4e-6f * (mz + mw);
            const int k = threadIdx.x & 7;
            const V4 c = to_clip(M, (k & 1) ? bh.x : bl.x, (k & 2) ? bh.y : bl.y, (k & 4) ? bh.z : bl.z);
            uint32_t o = 0;
            if (c.x < -c.w - ex) o |= LEFT_BIT;
            if (c.x > c.w + ex) o |= RIGHT_BIT;
            if (c.y < -c.w - ey) o |= BOTTOM_BIT;
            if (c.y > c.w + ey) o |= TOP_BIT;
            if (c.z > c.w + ez) o |= FAR_BIT;
            if (c.z < -ez) o |= NEAR_BIT;
            o = __reduce_and_sync(0xFFFFFFFFu, o);          // lanes 8..31 repeat corners 0..7
            if (threadIdx.x == 0) sOut = bl.w != 0.0f ? o : 0u;
        }
        __syncthreads();
        if (sOut) return;
    }
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nTris) return;
    geom_triangle(P, t);
}

// ---------------------------------------------------------------------------------------------
// List front end (FrameParams::frontEnd 1 / 2), for large meshes:
//   cull_kernel       one THREAD per cluster: the frustum test geom_kernel's CTAs do for themselves, once, for every
//                     256-triangle cluster (survivors are compacted into workList) and every 256-vertex cluster (flag).
//                     geom_kernel pays a launch + a dependent load + a barrier per culled CTA (a fifth of its
//                     stall samples on the 10M-triangle grid); here culled clusters cost one thread each.
//   vertex_kernel     stages a1, a2, a5 (project, raster transform, snap), a6 once per vertex of every unflagged
//                     vertex cluster (Renderer::VertexProcessing is per vertex too, Renderer.cpp:120-127) -> vrec.
//   geom_list_kernel  persistent: CTA b walks surviving clusters b, b + gridDim.x, ... (no barrier: each warp its own slice).
// ---------------------------------------------------------------------------------------------
// Clip planes that ALL eight corners of the box are outside of, by a margin that dominates the fp32 rounding of the
// transform of any point in the box: every vertex inside the box then has that bit in its clip code. 0 for an
// invalid box (NaN / inf coordinates).
__device__ __forceinline__ uint32_t box_outside_planes(const float* M, const float4 bl, const float4 bh)
{
    if (bl.w == 0.0f) return 0u;
    const float ax = fmaxf(fabsf(bl.x), fabsf(bh.x)), ay = fmaxf(fabsf(bl.y), fabsf(bh.y)), az = fmaxf(fabsf(bl.z), fabsf(bh.z));
    const float mx = fabsf(M[0]) * ax + fabsf(M[1]) * ay + fabsf(M[2]) * az + fabsf(M[3]);
    const float my = fabsf(M[4]) * ax + fabsf(M[5]) * ay + fabsf(M[6]) * az + fabsf(M[7]);
    const float mz = fabsf(M[8]) * ax + fabsf(M[9]) * ay + fabsf(M[10]) * az + fabsf(M[11]);
    const float mw = fabsf(M[12]) * ax + fabsf(M[13]) * ay + fabsf(M[14]) * az + fabsf(M[15]);
    const float ex = 4e-6f * (mx + mw), ey = 4e-6f * (my + mw), ez = 4e-6f * (mz + mw);
    uint32_t all = 63u;
    #pragma unroll
    for (int k = 0; k < 8; k++) {
        const V4 c = to_clip(M, (k & 1) ? bh.x : bl.x, (k & 2) ? bh.y : bl.y, (k & 4) ? bh.z : bl.z);
        uint32_t o = 0;
        if (c.x < -c.w - ex) o |= LEFT_BIT;
        if (c.x > c.w + ex) o |= RIGHT_BIT;
        if (c.y < -c.w - ey) o |= BOTTOM_BIT;
        if (c.y > c.w + ey) o |= TOP_BIT;
        if (c.z > c.w + ez) o |= FAR_BIT;
        if (c.z < -ez) o |= NEAR_BIT;
        all &= o;
    }
    return all;
}

__global__ void __launch_bounds__(256) cull_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j < P.nTriClusters) {
        const uint32_t out = P.clusterCull ? box_outside_planes(P.mvp, __ldg(P.clusterBox + 2 * j), __ldg(P.clusterBox + 2 * j + 1)) : 0u;
        if (!out) P.workList[warp_append(&P.counters->nWork)] = j;
    } else if (j - P.nTriClusters < P.nVertClusters && P.frontEnd == 2) {
        const uint32_t v = j - P.nTriClusters;
        P.vcFlag[v] = P.clusterCull ? box_outside_planes(P.mvp, __ldg(P.vclusterBox + 2 * v), __ldg(P.vclusterBox + 2 * v + 1)) : 0u;
    }
}

__global__ void __launch_bounds__(256) vertex_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    for (uint32_t vc = blockIdx.x; vc < P.nVertClusters; vc += gridDim.x) {
        if (P.vcFlag[vc]) continue;                     // every vertex of it is outside one plane: the flag stands in for the records
        const uint32_t v = vc * 256u + threadIdx.x;
        if (v >= P.nVerts) continue;
        const float4 p = __ldcs(P.pos4 + v);          // (streaming: read once per frame)
        const V4 c = to_clip(P.mvp, p.x, p.y, p.z);                          // a1
        const uint32_t code = surely_inside(c) ? 0u : clip_code(c);         // a2
        int4 rec;
        if (code) {
            rec = make_int4(0, 0, 0, (int)(VREC_OUTSIDE | code));
        } else {
            int sx, sy;
            project_snap(P.raster, P.rasterAffineXY != 0, c, sx, sy);       // a5: the part of Setup that depends on one vertex only
            const float iw = inv_w(c.w);                                     // a6
            const float zw = fmul(c.z, iw);
            rec = make_int4(sx, sy, __float_as_int(zw), iw != iw ? 0x7FFFFFFF : __float_as_int(iw));
        }
        P.vrec[v] = rec;
    }
}

__device__ __forceinline__ int4 load_vrec(const FrameParams& P, uint32_t i)
{
    // both loads are issued together (a culled cluster's records are stale, not unmapped): index -> flag -> record as a
    // dependent chain was a third of geom_list_kernel's stall samples on the 10M-triangle grid
    const uint32_t f = __ldg(P.vcFlag + (i >> 8));
    const int4 r = __ldg(P.vrec + i);
    return f ? make_int4(0, 0, 0, (int)(VREC_OUTSIDE | f)) : r;
}

// One submitted triangle from the per-vertex records (front end 2). Same arithmetic as geom_triangle - what
// project_snap / inv_w / z*invW compute depends on the vertex only - so the results are bit-identical.
__device__ __forceinline__ void geom_triangle_vrec(const FrameParams& P, uint32_t t)
{
    const uint32_t i0 = __ldcs(P.i0 + t), i1 = __ldcs(P.i1 + t), i2 = __ldcs(P.i2 + t);      // (streaming: read once per frame)
    const int4 r0 = load_vrec(P, i0), r1 = load_vrec(P, i1), r2 = load_vrec(P, i2);
    const uint32_t k0 = vrec_code(r0.w), k1 = vrec_code(r1.w), k2 = vrec_code(r2.w);
    if (k0 | k1 | k2) {
        if (k0 & k1 & k2) return;                        // Clipper.h:109
        // Straddler (rare): its clip-space vertices travel with it, so transform the three corners here. A flagged
        // cluster only reports ONE plane bit per vertex, so the reject test is repeated with the exact codes.
        const float4 p0 = __ldg(P.pos4 + i0), p1 = __ldg(P.pos4 + i1), p2 = __ldg(P.pos4 + i2);
        const V4 c0 = to_clip(P.mvp, p0.x, p0.y, p0.z), c1 = to_clip(P.mvp, p1.x, p1.y, p1.z), c2 = to_clip(P.mvp, p2.x, p2.y, p2.z);
        const uint32_t e0 = clip_code(c0), e1 = clip_code(c1), e2 = clip_code(c2);
        if (e0 & e1 & e2) return;
        queue_straddler(P, t, c0, c1, c2, (e0 ^ e1) | (e1 ^ e2) | (e2 ^ e0));
        return;
    }
    SetupTri s;
    s.v0x = r0.x; s.v0y = r0.y; s.v1x = r1.x; s.v1y = r1.y; s.v2x = r2.x; s.v2y = r2.y;
    if (!P.dump) {
        const bool ms = P.samples > 1;               // exact early cull: no pixel centre in the box
        if (max(0, first_pixel(min3i(s.v0x, s.v1x, s.v2x), ms)) > min(P.width - 1, last_pixel(max3i(s.v0x, s.v1x, s.v2x), ms)) ||
            max(0, first_pixel(min3i(s.v0y, s.v1y, s.v2y), ms)) > min(P.height - 1, last_pixel(max3i(s.v0y, s.v1y, s.v2y), ms)))
            return;
    }
    if (!finish_setup(s)) return;
    route_triangle(P, s, __int_as_float(r0.z), __int_as_float(r1.z), __int_as_float(r2.z),
                   __int_as_float(r0.w), __int_as_float(r1.w), __int_as_float(r2.w), t * 8u, P.smallMax);
}

template <bool VREC>
__global__ void __launch_bounds__(256, 5) geom_list_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    const uint32_t n = P.counters->nWork;
    // Static round robin, no barrier: the surviving clusters of a frame are many per CTA and alike, and a ticket per
    // cluster meant two CTA barriers around it - the warps of a CTA waited for the slowest at every cluster (a fifth of
    // the kernel's stall samples). Each warp now walks its 32-triangle slice of the CTA's clusters on its own.
    for (uint32_t cur = blockIdx.x; cur < n; cur += gridDim.x) {
        const uint32_t t = __ldg(P.workList + cur) * 256u + threadIdx.x;
        if (t < P.nTris) {
            if (VREC) geom_triangle_vrec(P, t);
            else geom_triangle(P, t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// clip_kernel: stages a3/a4 (Clipper.h:71-288) for queued straddlers
// ---------------------------------------------------------------------------------------------
struct Poly {
    V4 p[10];
    float w[10][3];
    int n;
};

__device__ __forceinline__ bool plane_inside(int plane, const V4& v)
{
    switch (plane) {
    case LEFT_BIT:   return v.x >= -v.w;
    case RIGHT_BIT:  return v.x <= v.w;
    case BOTTOM_BIT: return v.y >= -v.w;
    case TOP_BIT:    return v.y <= v.w;
    case FAR_BIT:    return v.z <= v.w;
    default:         return v.z >= 0.0f;
    }
}

__device__ __forceinline__ float plane_param(int plane, const V4& a, const V4& b)
{
    switch (plane) {                                  // Clipper.h:241,248,255,262,269,276
    case LEFT_BIT:   return fdiv(fadd(a.w, a.x), fsub(fadd(a.x, a.w), fadd(b.x, b.w)));
    case RIGHT_BIT:  return fdiv(fsub(a.x, a.w), fsub(fsub(a.x, a.w), fsub(b.x, b.w)));
    case BOTTOM_BIT: return fdiv(fadd(a.w, a.y), fsub(fadd(a.y, a.w), fadd(b.y, b.w)));
    case TOP_BIT:    return fdiv(fsub(a.y, a.w), fsub(fsub(a.y, a.w), fsub(b.y, b.w)));
    case FAR_BIT:    return fdiv(fsub(a.z, a.w), fsub(fsub(a.z, a.w), fsub(b.z, b.w)));
    default:         return fdiv(a.z, fsub(a.z, b.z));
    }
}

// one new vertex where edge a->b crosses `plane` (Clipper.h:210-214): position, snap, clip weights
__device__ __forceinline__ void cut_vertex(int plane, const V4& a, const V4& b, const float* wa, const float* wb, V4& r, float* rw)
{
    const float t = plane_param(plane, a, b);
    const float s = fsub(1.0f, t);
    r.x = fadd(fmul(a.x, s), fmul(b.x, t));
    r.y = fadd(fmul(a.y, s), fmul(b.y, t));
    r.z = fadd(fmul(a.z, s), fmul(b.z, t));
    r.w = fadd(fmul(a.w, s), fmul(b.w, t));
    switch (plane) {                                  // snap onto the plane (:242,249,256,263,270,277)
    case LEFT_BIT:   r.x = -r.w; break;
    case RIGHT_BIT:  r.x = r.w; break;
    case BOTTOM_BIT: r.y = -r.w; break;
    case TOP_BIT:    r.y = r.w; break;
    case FAR_BIT:    r.z = r.w; break;
    default:         r.z = 0.0f; break;
    }
    #pragma unroll
    for (int k = 0; k < 3; k++) rw[k] = fadd(fmul(wa[k], s), fmul(wb[k], t));
}

__device__ __forceinline__ void cut_edge(int plane, const Poly& in, int i, int j, Poly& out)
{
    const int n = out.n++;
    cut_vertex(plane, in.p[i], in.p[j], in.w[i], in.w[j], out.p[n], out.w[n]);
}

__device__ __forceinline__ void copy_vertex(const Poly& in, int j, Poly& out)
{
    int n = out.n++;
    out.p[n] = in.p[j];
    out.w[n][0] = in.w[j][0]; out.w[n][1] = in.w[j][1]; out.w[n][2] = in.w[j][2];
}

__device__ void clip_by_plane(int plane, const Poly& in, Poly& out)     // Clipper.h:192-232
{
    out.n = 0;
    for (int i = 0; i < in.n; i++) {
        int j = (i + 1 == in.n) ? 0 : i + 1;
        bool in0 = plane_inside(plane, in.p[i]), in1 = plane_inside(plane, in.p[j]);
        if (in0) {
            if (in1) copy_vertex(in, j, out);
            else cut_edge(plane, in, i, j, out);
        } else if (in1) {
            cut_edge(plane, in, i, j, out);
            copy_vertex(in, j, out);
        }
    }
}

// Clipper.h:121-153: a polygon vertex whose weight is exactly 1 IS that original vertex
__device__ __forceinline__ uint32_t vertex_source(const float* w)
{
    if (w[0] == 1.0f) return 0;
    if (w[1] == 1.0f) return 1;
    if (w[2] == 1.0f) return 2;
    return 3;
}

#ifdef EDX_DEBUG_STATS
__device__ unsigned long long g_clipDbg[4];      // longest per-warp cycles: single-plane loop, multi-plane loop, whole kernel body; [3] multi-plane loop until the polygons are clipped
#endif
// One fan triangle (0, k-1, k) of a clipped polygon: setup, shading record, routing (Clipper.h:156-170).
// Deliberately not inlined: the clipper runs a few thousand threads, each serially; a small code
// footprint (instruction-cache hits) matters more than call overhead there.
__device__ __noinline__ void emit_fan(const FrameParams& P, uint32_t t, int fan, uint32_t slot, bool haveRecs,
                                         const V4& f0, const V4& f1, const V4& f2, float iwA, float zA,
                                         uint32_t srcBits, const float* w0, const float* w1, const float* w2)
{
    SetupTri s;
    const bool ok = setup_tri(P.raster, P.rasterAffineXY != 0, f0, f1, f2, s);
    const float iwB = inv_w(f1.w), iwC = inv_w(f2.w);
    if (haveRecs) {
        ClipRec r;
        r.v0x = s.v0x; r.v0y = s.v0y; r.v1x = s.v1x; r.v1y = s.v1y; r.v2x = s.v2x; r.v2y = s.v2y;
        r.invDet = ok ? s.invDet : 0.0f;
        r.src = srcBits;
        r.invW0 = iwA; r.invW1 = iwB; r.invW2 = iwC; r.valid = ok ? 1u : 0u;
        #pragma unroll
        for (int m = 0; m < 3; m++) { r.wt[0][m] = w0[m]; r.wt[1][m] = w1[m]; r.wt[2][m] = w2[m]; }
        r.pad[0] = r.pad[1] = r.pad[2] = 0.0f;
        int4* dst = reinterpret_cast<int4*>(P.clipRecs + slot + fan);
        const int4* src = reinterpret_cast<const int4*>(&r);
        #pragma unroll
        for (int m = 0; m < 6; m++) dst[m] = src[m];
    }
    // without a record the resolve pass could not shade this fan triangle: the frame is incomplete and will be re-run
    // with larger queues (finish_frame), so do not leave keys that point at a record that was never written
    if (ok && haveRecs) route_triangle(P, s, zA, fmul(f1.z, iwB), fmul(f2.z, iwC), iwA, iwB, iwC, t * 8u + (uint32_t)fan, P.smallMaxClip);
}

__device__ __forceinline__ V4 pick3(const V4* c, uint32_t i) { return i == 0 ? c[0] : (i == 1 ? c[1] : c[2]); }

// One clip plane as data instead of control flow, so that straddlers of DIFFERENT planes run the same instructions
// side by side in a warp. Clipper.h:237-278 per plane: inside test, t = d0 / (d0 - d1), snap.
//   LEFT / BOTTOM   inside c >= -w   d(v) = c + w   ((w + c) in the reference: commutative)     snap c = -w
//   RIGHT / TOP / FAR  inside c <= w   d(v) = c - w   ((-w + c) in the reference: the same sum)   snap c = w
//   NEAR            inside z >= 0    d(v) = z                                                     snap z = 0
struct PlaneSel {
    int comp;        // 0 x, 1 y, 2 z
    bool plus;       // LEFT / BOTTOM
    bool nearp;
    __device__ __forceinline__ explicit PlaneSel(uint32_t plane)
    {
        comp = (plane & (LEFT_BIT | RIGHT_BIT)) ? 0 : ((plane & (BOTTOM_BIT | TOP_BIT)) ? 1 : 2);
        plus = (plane & (LEFT_BIT | BOTTOM_BIT)) != 0;
        nearp = plane == NEAR_BIT;
    }
    __device__ __forceinline__ float coord(const V4& v) const { return comp == 0 ? v.x : (comp == 1 ? v.y : v.z); }
    __device__ __forceinline__ bool inside(const V4& v) const
    {
        const float c = coord(v);
        return nearp ? c >= 0.0f : (plus ? c >= -v.w : c <= v.w);
    }
    __device__ __forceinline__ float dist(const V4& v) const
    {
        const float c = coord(v);
        return nearp ? c : fadd(c, plus ? v.w : -v.w);
    }
    // one new vertex where edge a->b crosses the plane (Clipper.h:210-214): position, snap, clip weights
    __device__ __forceinline__ void cut(const V4& a, const V4& b, const float* wa, const float* wb, V4& r, float* rw) const
    {
        const float da = dist(a), db = dist(b);
        const float t = fdiv(da, fsub(da, db));
        const float s = fsub(1.0f, t);
        r.x = fadd(fmul(a.x, s), fmul(b.x, t));
        r.y = fadd(fmul(a.y, s), fmul(b.y, t));
        r.z = fadd(fmul(a.z, s), fmul(b.z, t));
        r.w = fadd(fmul(a.w, s), fmul(b.w, t));
        const float snap = nearp ? 0.0f : (plus ? -r.w : r.w);
        r.x = comp == 0 ? snap : r.x; r.y = comp == 1 ? snap : r.y; r.z = comp == 2 ? snap : r.z;
        #pragma unroll
        for (int k = 0; k < 3; k++) rw[k] = fadd(fmul(wa[k], s), fmul(wb[k], t));
    }
};

// One straddler that crosses exactly ONE clip plane (the common case at screen edges): the polygon has 3 or
// 4 vertices in an order fixed by which vertices are inside (Clipper.h:196-229), so it is built with static
// indices and lives in registers. Plane and inside pattern only enter through selects: a warp full of such
// straddlers - whatever their planes - executes one instruction stream. Not inlined: it is shared by clip_kernel and
// (optionally) geom_kernel, where it must not raise the hot path's register count.
__device__ __noinline__ void clip_single_plane(const FrameParams& P, uint32_t t, const V4& c0, const V4& c1, const V4& c2, uint32_t planes)
{
    const V4 c[3] = { c0, c1, c2 };
    const PlaneSel pl(planes);
    const uint32_t in = (pl.inside(c[0]) ? 1u : 0u) | (pl.inside(c[1]) ? 2u : 0u) | (pl.inside(c[2]) ? 4u : 0u);
    // Exactly two edges cross the plane. In the order Clipper.h:196-229 visits them they are
    //   in = 1:(0>1),(2>0)  2:(0>1),(1>2)  4:(1>2),(2>0)  6:(0>1),(2>0)  5:(0>1),(1>2)  3:(1>2),(2>0)
    const bool firstIs01 = (in != 4u && in != 3u), secondIs20 = (in == 1u || in == 4u || in == 6u || in == 3u);
    const uint32_t ai = firstIs01 ? 0u : 1u, aj = firstIs01 ? 1u : 2u;
    const uint32_t bi = secondIs20 ? 2u : 1u, bj = secondIs20 ? 0u : 2u;
    V4 cutA, cutB; float wA[3], wB[3];
    {
        const float ua[3] = { ai == 0u ? 1.0f : 0.0f, ai == 1u ? 1.0f : 0.0f, 0.0f };
        const float uaj[3] = { 0.0f, aj == 1u ? 1.0f : 0.0f, aj == 2u ? 1.0f : 0.0f };
        pl.cut(pick3(c, ai), pick3(c, aj), ua, uaj, cutA, wA);
        const float ub[3] = { 0.0f, bi == 1u ? 1.0f : 0.0f, bi == 2u ? 1.0f : 0.0f };
        const float ubj[3] = { bj == 0u ? 1.0f : 0.0f, 0.0f, bj == 2u ? 1.0f : 0.0f };
        pl.cut(pick3(c, bi), pick3(c, bj), ub, ubj, cutB, wB);
    }
    // assemble the polygon: slot pattern per inside-mask (3 = cut A, 4 = cut B, digits = original vertex), three bits
    // per slot:   1:[A,B,0]  2:[A,1,B]  4:[A,2,B]  6:[A,1,2,B]  5:[A,B,2,0]  3:[1,A,B,0]
    const uint32_t pattern = in == 1u ? 00043u : (in == 2u ? 04413u : (in == 4u ? 04423u : (in == 6u ? 04213u : (in == 5u ? 00243u : 00431u))));
    V4 v[4]; float w[4][3]; int nv = (in == 1u || in == 2u || in == 4u) ? 3 : 4;
    if (in == 0u || in == 7u) nv = 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        const int kd = (int)((pattern >> (3 * k)) & 7u);
        v[k] = kd == 3 ? cutA : (kd == 4 ? cutB : pick3(c, (uint32_t)kd));
        #pragma unroll
        for (int m = 0; m < 3; m++) w[k][m] = kd == 3 ? wA[m] : (kd == 4 ? wB[m] : (kd == m ? 1.0f : 0.0f));
    }
    bool drop = nv == 0;
    #pragma unroll
    for (int k = 0; k < 4; k++) if (k < nv && v[k].w <= 0.0f) drop = true;                  // Clipper.h:280-287
    if (drop) return;
    uint32_t src[4];
    #pragma unroll
    for (int k = 0; k < 4; k++) { src[k] = vertex_source(w[k]); if (src[k] < 3) v[k] = pick3(c, src[k]); }
    const uint32_t nFan = (uint32_t)(nv - 2);
    const uint32_t slot = atomicAdd(&P.counters->nClipRecs, nFan);
    const bool haveRecs = slot + nFan <= P.clipRecCap;
    if (haveRecs) P.clipSlot[t] = slot;
    const float iwA = inv_w(v[0].w), zA = fmul(v[0].z, iwA);
    emit_fan(P, t, 0, slot, haveRecs, v[0], v[1], v[2], iwA, zA, src[0] | (src[1] << 2) | (src[2] << 4), w[0], w[1], w[2]);
    if (nv == 4)
        emit_fan(P, t, 1, slot, haveRecs, v[0], v[2], v[3], iwA, zA, src[0] | (src[2] << 2) | (src[3] << 4), w[0], w[2], w[3]);
}

__global__ void __launch_bounds__(128) clip_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
#ifdef EDX_DEBUG_STATS
    const long long tc0 = clock64();
#endif
    // The queue has two halves. Front: straddlers of exactly one plane - one instruction stream whatever the plane,
    // so they are packed 32 to a warp. Back: straddlers of several planes - long, data-dependent loops over
    // local-memory polygons - spread one per warp first so a short queue costs one item's latency, not 32.
    const uint32_t half = P.clipQueueCap / 2u;
    const uint32_t n1 = min(P.counters->nClipQueue, half), nN = min(P.counters->nClipMulti, P.clipQueueCap - half);
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, nThreads = gridDim.x * blockDim.x;
    for (uint32_t q = gtid; q < n1; q += nThreads) {
        const float4* item = reinterpret_cast<const float4*>(P.clipQueue + q);
        const float4 q0 = __ldg(item), q1 = __ldg(item + 1), q2 = __ldg(item + 2), q3 = __ldg(item + 3);
        V4 c0, c1, c2;
        c0.x = q0.x; c0.y = q0.y; c0.z = q0.z; c0.w = q0.w;
        c1.x = q1.x; c1.y = q1.y; c1.z = q1.z; c1.w = q1.w;
        c2.x = q2.x; c2.y = q2.y; c2.z = q2.z; c2.w = q2.w;
        clip_single_plane(P, __float_as_uint(q3.x), c0, c1, c2, __float_as_uint(q3.y));
    }
#ifdef EDX_DEBUG_STATS
    const long long tc1 = clock64();
#endif
    // Work item q goes to lane q / W of warp q % W (W = warps in the grid)
    const uint32_t W = gridDim.x * (blockDim.x >> 5);
    // (counted from the END of the grid: the single-plane items above fill the warps from the front, and a warp that
    // had both would run the two latency chains one after the other)
    const uint32_t gw = W - 1u - (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    const uint32_t lane = threadIdx.x & 31u;
    // Warp-uniform loop (lane 0 has the smallest q): after the per-lane clip the fan triangles of a straddler are
    // emitted by DIFFERENT LANES when the warp holds only a few straddlers - the usual case, one per warp - because setup
    // + routing of up to seven fan triangles one after the other was most of this path's latency, with 31 lanes idle.
    for (uint32_t q0w = gw; q0w < nN; q0w += 32u * W) {
        const uint32_t q = q0w + lane * W;
        const bool has = q < nN;
        Poly a, b;
        Poly* cur = &a; Poly* buf = &b;
        a.n = 0;
        uint32_t t = 0, srcs = 0, slot = 0;
        int nv = 0;
        bool haveRecs = false;
        V4 c[3];
        if (has) {
            const float4* item = reinterpret_cast<const float4*>(P.clipQueue + half + q);
            const float4 q0 = __ldg(item), q1 = __ldg(item + 1), q2 = __ldg(item + 2), q3 = __ldg(item + 3);
            t = __float_as_uint(q3.x);
            c[0].x = q0.x; c[0].y = q0.y; c[0].z = q0.z; c[0].w = q0.w;
            c[1].x = q1.x; c[1].y = q1.y; c[1].z = q1.z; c[1].w = q1.w;
            c[2].x = q2.x; c[2].y = q2.y; c[2].z = q2.z; c[2].w = q2.w;
            const uint32_t k0 = clip_code(c[0]), k1 = clip_code(c[1]), k2 = clip_code(c[2]);
            const uint32_t planes = (k0 ^ k1) | (k1 ^ k2) | (k2 ^ k0);           // Clipper.h:119

            // General path: several planes, up to 9 vertices, polygon in local memory.
            a.n = 3;
            for (int k = 0; k < 3; k++) {
                a.p[k] = c[k];
                a.w[k][0] = k == 0 ? 1.0f : 0.0f; a.w[k][1] = k == 1 ? 1.0f : 0.0f; a.w[k][2] = k == 2 ? 1.0f : 0.0f;
            }
            const int order[6] = { LEFT_BIT, RIGHT_BIT, BOTTOM_BIT, TOP_BIT, FAR_BIT, NEAR_BIT };   // Clipper.h:237-278
            for (int k = 0; k < 6; k++) {
                if (planes & order[k]) { clip_by_plane(order[k], *cur, *buf); Poly* tmp = cur; cur = buf; buf = tmp; }
            }
            nv = cur->n;
            for (int k = 0; k < cur->n; k++)
                if (cur->p[k].w <= 0.0f) nv = 0;                           // Clipper.h:280-287
            if (nv < 3) nv = 0;
            for (int k = 0; k < nv; k++) {
                const uint32_t src = vertex_source(cur->w[k]);
                if (src < 3) cur->p[k] = pick3(c, src);
                srcs |= src << (2 * k);
            }
            if (nv) {
                const uint32_t nFan = (uint32_t)(nv - 2);
                slot = atomicAdd(&P.counters->nClipRecs, nFan);
                haveRecs = slot + nFan <= P.clipRecCap;
                if (haveRecs) P.clipSlot[t] = slot;
            }
        }
#ifdef EDX_DEBUG_STATS
        if (lane == 0) atomicMax(&g_clipDbg[3], (unsigned long long)(clock64() - tc1));      // multi-plane: until the polygons are clipped (first iteration)
#endif
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, nv != 0);
        if (__popc(act) > 4) {
            // a long queue: every lane has its own straddler, the lanes already run in parallel
            if (nv) {
                const V4 f0 = cur->p[0];
                const float iwA = inv_w(f0.w), zA = fmul(f0.z, iwA);
                for (int k = 2; k < nv; k++) {                                  // Clipper.h:156-170 fan (0, k-1, k)
                    const uint32_t sb = (srcs & 3u) | (((srcs >> (2 * (k - 1))) & 3u) << 2) | (((srcs >> (2 * k)) & 3u) << 4);
                    emit_fan(P, t, k - 2, slot, haveRecs, f0, cur->p[k - 1], cur->p[k], iwA, zA, sb, cur->w[0], cur->w[k - 1], cur->w[k]);
                }
            }
            continue;
        }
        for (uint32_t rest = act; rest; rest &= rest - 1u) {
            const int src = __ffs(rest) - 1;
            const int nvS = __shfl_sync(0xFFFFFFFFu, nv, src);
            const uint32_t tS = __shfl_sync(0xFFFFFFFFu, t, src), slotS = __shfl_sync(0xFFFFFFFFu, slot, src), srcsS = __shfl_sync(0xFFFFFFFFu, srcs, src);
            const bool recsS = __shfl_sync(0xFFFFFFFFu, haveRecs ? 1 : 0, src) != 0;
            // lane k takes fan triangle (0, k + 1, k + 2): the owner broadcasts its vertices one by one
            V4 f0, fa, fb; float w0[3], wa[3], wb[3];
            f0.x = f0.y = f0.z = f0.w = 0.0f; fa = f0; fb = f0;
            #pragma unroll
            for (int m = 0; m < 3; m++) { w0[m] = 0.0f; wa[m] = 0.0f; wb[m] = 0.0f; }
            for (int j = 0; j < nvS; j++) {
                const int jj = min(j, 8);
                V4 v; float w[3];
                v.x = __shfl_sync(0xFFFFFFFFu, cur->p[jj].x, src); v.y = __shfl_sync(0xFFFFFFFFu, cur->p[jj].y, src);
                v.z = __shfl_sync(0xFFFFFFFFu, cur->p[jj].z, src); v.w = __shfl_sync(0xFFFFFFFFu, cur->p[jj].w, src);
                #pragma unroll
                for (int m = 0; m < 3; m++) w[m] = __shfl_sync(0xFFFFFFFFu, cur->w[jj][m], src);
                if (j == 0) { f0 = v; for (int m = 0; m < 3; m++) w0[m] = w[m]; }
                if (j == (int)lane + 1) { fa = v; for (int m = 0; m < 3; m++) wa[m] = w[m]; }
                if (j == (int)lane + 2) { fb = v; for (int m = 0; m < 3; m++) wb[m] = w[m]; }
            }
            if ((int)lane + 2 < nvS) {
                const int k = (int)lane + 2;
                const float iwA = inv_w(f0.w), zA = fmul(f0.z, iwA);
                const uint32_t sb = (srcsS & 3u) | (((srcsS >> (2 * (k - 1))) & 3u) << 2) | (((srcsS >> (2 * k)) & 3u) << 4);
                emit_fan(P, tS, k - 2, slotS, recsS, f0, fa, fb, iwA, zA, sb, w0, wa, wb);
            }
            __syncwarp();
        }
    }
#ifdef EDX_DEBUG_STATS
    if ((threadIdx.x & 31u) == 0) {
        const long long tc2 = clock64();
        atomicMax(&g_clipDbg[0], (unsigned long long)(tc1 - tc0)); atomicMax(&g_clipDbg[1], (unsigned long long)(tc2 - tc1));
        atomicMax(&g_clipDbg[2], (unsigned long long)(tc2 - tc0));
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// mid_kernel: stages a10, a12, a13 for mid-size triangles, one WARP per triangle. The lanes form an 8 x 4 pixel
// block that steps over the triangle's pixel-centre box; each lane evaluates the three biased edge functions of its
// pixel (Rasterizer.h:162) from integer steps, and covered pixels go through the same barycentric / depth / key
// arithmetic as everywhere else into the L2-resident key buffer (atomicMin = the sequential depth test).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mid_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    const uint32_t n = min(P.counters->nMid, P.midCap);
    const uint32_t lane = threadIdx.x & 31u, nWarps = gridDim.x * (blockDim.x >> 5), gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool ms = P.samples > 1;
    const int lx = (int)(lane & 7u), ly = (int)(lane >> 3);
    unsigned long long boxArea = 0;          // pixels this warp's triangles' boxes span: the fragment load of a path that has no occlusion culling
    for (uint32_t q = gw; q < n; q += nWarps) {
        const int4* rp = reinterpret_cast<const int4*>(P.mid + q);
        const int4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);      // same address in every lane: one broadcast transaction each
        const float z0 = __int_as_float(b.z), z1 = __int_as_float(b.w), z2 = __int_as_float(c.x), invDet = __int_as_float(c.y);
        const uint32_t prim = (uint32_t)c.z;
        Edges e;
        e.init(a.x, a.y, a.z, a.w, b.x, b.y);
        const int x0 = max(0, first_pixel(min3i(a.x, a.z, b.x), ms)), x1 = min(P.width - 1, last_pixel(max3i(a.x, a.z, b.x), ms));
        const int y0 = max(0, first_pixel(min3i(a.y, a.w, b.y), ms)), y1 = min(P.height - 1, last_pixel(max3i(a.y, a.w, b.y), ms));
        if (x1 >= x0 && y1 >= y0) boxArea += (unsigned long long)((x1 - x0 + 1) * (y1 - y0 + 1));
        // biased edge values at this lane's pixel of the first block; a block step is 8 pixels in x, 4 in y
        const int cx = ((x0 + lx) << 4) + 8, cy = ((y0 + ly) << 4) + 8;
        uint32_t r0 = (uint32_t)e.e0(cx, cy), r1 = (uint32_t)e.e1(cx, cy), r2 = (uint32_t)e.e2(cx, cy);
        const uint32_t sB0 = e.B0 << 7, sB1 = e.B1 << 7, sB2 = e.B2 << 7;     // 8 pixels = 128 sub-pixels
        const uint32_t sC0 = e.C0 << 6, sC1 = e.C1 << 6, sC2 = e.C2 << 6;     // 4 pixels = 64 sub-pixels
        for (int by = y0; by <= y1; by += 4) {
            uint32_t a0 = r0, a1 = r1, a2 = r2;
            const int y = by + ly;
            for (int bx = x0; bx <= x1; bx += 8) {
                const int x = bx + lx;
                if (x <= x1 && y <= y1) {
                    if (!ms) {
                        if ((int)(a0 | a1 | a2) >= 0) {
                            float l0, l1;
                            barycentric((int)(a1 - (uint32_t)e.bias1), (int)(a2 - (uint32_t)e.bias2), invDet, l0, l1);
                            const float d = depth_at(l0, l1, z0, z1, z2);
                            if (d <= 1.0f && owns_pixel(x, y, P.binsX, P.part, P.parts))
                                atomicMin(P.keys + key_index(x, y, P.binsX), make_key(d, prim));
                        }
                    } else {
                        const int* off = c_sampleOffsets[P.msLevel];
                        const uint32_t ki = key_index(x, y, P.binsX);
                        const bool owned = owns_pixel(x, y, P.binsX, P.part, P.parts);
                        for (int sId = 0; sId < P.samples; sId++) {
                            const uint32_t ox = (uint32_t)off[2 * sId], oy = ox;      // DESIGN.md shim 17
                            const uint32_t f0 = a0 + ox * e.B0 + oy * e.C0, f1 = a1 + ox * e.B1 + oy * e.C1, f2 = a2 + ox * e.B2 + oy * e.C2;
                            if ((int)(f0 | f1 | f2) >= 0) {
                                float l0, l1;
                                barycentric((int)(f1 - (uint32_t)e.bias1), (int)(f2 - (uint32_t)e.bias2), invDet, l0, l1);
                                const float d = depth_at(l0, l1, z0, z1, z2);
                                if (d <= 1.0f && owned) atomicMin(P.keys + (size_t)sId * P.keyStride + ki, make_key(d, prim));
                            }
                        }
                    }
                }
                a0 += sB0; a1 += sB1; a2 += sB2;
            }
            r0 += sC0; r1 += sC1; r2 += sC2;
        }
    }
    if (lane == 0 && boxArea) atomicAdd(&P.counters->midAreaSlot[gw & 63u], boxArea);
}

// ---------------------------------------------------------------------------------------------
// tile_kernel helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_f32(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Exact classification of a triangle against the pixel centres of rect [px0, px0+size) x [py0, py0+size)
// (clamped to the screen), using the same biased edge functions as per-pixel coverage. `reject`: no
// centre can be covered; `full`: every centre is covered. This is the role of the reject / accept
// corners of RasterTriangle.h:65-150, Renderer.cpp:189-224 and Rasterizer.h:42-85, evaluated at the
// extreme pixel centres instead of the tile corners (so it is exact rather than conservative).
// With wantZ it also returns conservative bounds of the depth the triangle can produce in the rect.
__device__ __forceinline__ void classify_rect(const BigRec& r, int px0, int py0, int size, int W, int H, bool wantZ, bool ms,
                                              uint32_t cullU, bool& reject, bool& full, float& zmin, float& zmax)
{
    reject = true; full = false; zmin = 0.0f; zmax = 0.0f;
    // (W and H used as they are: operands straight from the constant bank, nothing loop-invariant to keep in a register)
    const int rx1 = min(px0 + size, W) - 1, ry1 = min(py0 + size, H) - 1;
    if (px0 > rx1 || py0 > ry1) return;
    const int tx0 = first_pixel(min3i(r.v0x, r.v1x, r.v2x), ms), tx1 = last_pixel(max3i(r.v0x, r.v1x, r.v2x), ms);
    const int ty0 = first_pixel(min3i(r.v0y, r.v1y, r.v2y), ms), ty1 = last_pixel(max3i(r.v0y, r.v1y, r.v2y), ms);
    if (tx0 > rx1 || tx1 < px0 || ty0 > ry1 || ty1 < py0) return;
    if (wantZ) {
        // Depth is affine in the pixel position up to fp32 rounding. Bound the stored plane over the rect's
        // pixel centres, intersect with the vertex range (covered pixels lie inside the triangle), and widen by
        //   E  = the plane fit's own error bound (perr) + rounding of this evaluation (<= 2.5e-7 of the summed magnitudes)
        //   Ed = fp32 error of barycentric()/depth_at() at a covered pixel (<= ~1.3e-6 * max|z|)
        const float half = ms ? 0.5f : 0.0f;          // samples sit up to half a pixel away from the centre
        const float x0f = (float)px0 - half, x1f = (float)rx1 + half, y0f = (float)py0 - half, y1f = (float)ry1 + half;
        const float ax0 = r.gx * x0f, ax1 = r.gx * x1f, ay0 = r.gy * y0f, ay1 = r.gy * y1f;
        const float lo = r.zref + fminf(ax0, ax1) + fminf(ay0, ay1);
        const float hi = r.zref + fmaxf(ax0, ax1) + fmaxf(ay0, ay1);
        const float vlo = fminf(r.z0, fminf(r.z1, r.z2)), vhi = fmaxf(r.z0, fmaxf(r.z1, r.z2));
        const float E = r.perr + 2.5e-7f * (fabsf(r.zref) + fmaxf(fabsf(ax0), fabsf(ax1)) + fmaxf(fabsf(ay0), fabsf(ay1)));
        const float Ed = 1e-5f * fmaxf(fabsf(vlo), fabsf(vhi)) + 1e-30f;
        zmin = fmaxf(lo - E, vlo) - Ed;
        zmax = fminf(hi + E, vhi) + Ed;
        // Depth cull first: a triangle whose nearest possible depth in the rect is behind the rect's current
        // upper bound cannot own a pixel there; it is dropped before the (dearer) edge tests.
        if (order_f32(zmin) > cullU) return;
    }
    Edges e;
    e.init(r.v0x, r.v0y, r.v1x, r.v1y, r.v2x, r.v2y);
    // 1x: the extreme pixel CENTRES of the rect (exact). MSAA: the extreme sub-pixel positions any sample of
    // the rect's pixels can take, [16p, 16p + 15] (conservative: reject / full still imply the same for
    // every sample, they are just no longer "if and only if").
    const int subLo = ms ? 0 : 8, subHi = ms ? 15 : 8;
    const int cx0 = (px0 << 4) + subLo, cx1 = (rx1 << 4) + subHi, cy0 = (py0 << 4) + subLo, cy1 = (ry1 << 4) + subHi;
    const bool b0 = (int)e.B0 > 0, c0 = (int)e.C0 > 0, b1 = (int)e.B1 > 0, c1 = (int)e.C1 > 0, b2 = (int)e.B2 > 0, c2 = (int)e.C2 > 0;
    if (e.e0(b0 ? cx1 : cx0, c0 ? cy1 : cy0) < 0) return;
    if (e.e1(b1 ? cx1 : cx0, c1 ? cy1 : cy0) < 0) return;
    if (e.e2(b2 ? cx1 : cx0, c2 ? cy1 : cy0) < 0) return;
    reject = false;
    full = e.e0(b0 ? cx0 : cx1, c0 ? cy0 : cy1) >= 0 && e.e1(b1 ? cx0 : cx1, c1 ? cy0 : cy1) >= 0 &&
           e.e2(b2 ? cx0 : cx1, c2 ? cy0 : cy1) >= 0;
}

// One warp rasterises one triangle into its 16x16 tile: 8x8 block masks by ballot, then pixels.
// k[8] = this lane's eight keys of the tile, held in REGISTERS for the whole survivor walk (the warp owns
// the tile, so no atomics and no shared-memory round trip per pixel). Slot j = block (j >> 1), half (j & 1):
// pixel (tx0 + 8*(q&1) + lane%8, ty0 + 8*(q>>1) + lane/8 + 4*h). The eight pixels are independent, so the
// fully unrolled body gives the scheduler eight interleaved dependency chains.
__device__ __forceinline__ void raster_tile_tri(unsigned long long (&k)[8], const BigRec& r, int tx0, int ty0,
                                                int W, int H, bool full, bool hierarchical, bool ms, int offX, int offY)
{
    const int lane = threadIdx.x & 31;
    Edges e;
    e.init(r.v0x, r.v0y, r.v1x, r.v1y, r.v2x, r.v2y);
    uint32_t rejMask = 0, accMask = full ? 0xFu : 0u;
    if (!full && hierarchical) {
        bool rej, acc; float zl, zh;
        const int q = lane & 3;
        classify_rect(r, tx0 + (q & 1) * BLOCK_PX, ty0 + (q >> 1) * BLOCK_PX, BLOCK_PX, W, H, false, ms, 0xFFFFFFFFu, rej, acc, zl, zh);
        rejMask = __ballot_sync(0xFFFFFFFFu, rej) & 0xFu;
        accMask = __ballot_sync(0xFFFFFFFFu, acc) & 0xFu;
    }
    // biased edge values at this lane's pixel of block 0 / half 0; other slots are integer steps away
    const int bx = tx0 + (lane & 7), by = ty0 + (lane >> 3);
    const int cx = (bx << 4) + 8 + offX, cy = (by << 4) + 8 + offY;     // sample position (offset 0 at 1x)
    const uint32_t e0b = (uint32_t)e.e0(cx, cy), e1b = (uint32_t)e.e1(cx, cy), e2b = (uint32_t)e.e2(cx, cy);
    const uint32_t lowBias1 = (uint32_t)e.bias1, lowBias2 = (uint32_t)e.bias2;
    #pragma unroll
    for (int j = 0; j < 8; j++) {
        const int q = j >> 1, h = j & 1;
        const uint32_t sx = (uint32_t)((q & 1) * BLOCK_PX * 16), sy = (uint32_t)(((q >> 1) * BLOCK_PX + 4 * h) * 16);   // sub-pixel steps
        const bool live = !((rejMask >> q) & 1u) && (bx + (q & 1) * BLOCK_PX) < W && (by + (q >> 1) * BLOCK_PX + 4 * h) < H;
        const uint32_t a0 = e0b + e.B0 * sx + e.C0 * sy, a1 = e1b + e.B1 * sx + e.C1 * sy, a2 = e2b + e.B2 * sx + e.C2 * sy;
        // branch-free: everything is computed for every slot and committed with one select, so the eight
        // chains really interleave
        const bool covered = live && (((accMask >> q) & 1u) || (int)(a0 | a1 | a2) >= 0);
        float l0, l1;
        barycentric((int)(a1 - lowBias1), (int)(a2 - lowBias2), r.invDet, l0, l1);
        const float d = depth_at(l0, l1, r.z0, r.z1, r.z2);
        const unsigned long long key = make_key(d, r.prim);
        k[j] = (covered && d <= 1.0f && key < k[j]) ? key : k[j];
    }
}

__device__ __forceinline__ uint8_t to_u8(float c)       // Color4b::FromFloats (Renderer.cpp:296-299; DESIGN.md shim 13)
{
    float t = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    float s = fadd(fmul(t, 255.0f), 0.5f);
    if (!(s >= 0.0f)) return 0;
    return (uint8_t)__float2int_rz(s);
}

__device__ __forceinline__ float rsqrt_exact(float x) { return frcp(__fsqrt_rn(x)); }   // DESIGN.md shim 9

// ---------------------------------------------------------------------------------------------
// Texture2D<Color>::Sample (EDXUtil, absent): DESIGN.md shims 18-24, operation for operation as defined there.
// Texels are RGBA8; every float operation is a correctly rounded IEEE one except log2f,
// which only feeds a blend weight (continuous in its error).
// ---------------------------------------------------------------------------------------------
struct TexLevelRef { const uchar4* px; int w, h; };

__device__ __forceinline__ TexLevelRef tex_level(const FrameParams& P, const TexDesc* t, int l)
{
    TexLevelRef r;
    r.px = P.texels + __ldg(&t->off[l]);
    r.w = max(1, (int)(__ldg(&t->w) >> l));
    r.h = max(1, (int)(__ldg(&t->h) >> l));
    return r;
}

__device__ __forceinline__ int tex_wrap(int i, int n) { const int m = i % n; return m < 0 ? m + n : m; }   // shim 20

__device__ __forceinline__ void tex_texel(const TexLevelRef& L, int x, int y, float& r, float& g, float& b)
{
    const uchar4 c = __ldg(L.px + (size_t)tex_wrap(y, L.h) * L.w + tex_wrap(x, L.w));
    const float k = 1.0f / 255.0f;                                                        // shim 18
    r = fmul((float)c.x, k); g = fmul((float)c.y, k); b = fmul((float)c.z, k);
}

__device__ __forceinline__ float tex_coord(float u, int n, float bias)
{
    const float x = fsub(fmul(u, (float)n), bias);
    return fabsf(x) < 1.0e9f ? x : 0.0f;
}

__device__ __noinline__ void tex_bilinear(const TexLevelRef L, float u, float v, float& r, float& g, float& b)    // shim 22
{
    const float x = tex_coord(u, L.w, 0.5f), y = tex_coord(v, L.h, 0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = fsub(x, x0), fy = fsub(y, y0);
    const int ix = (int)x0, iy = (int)y0;
    float r00, g00, b00, r10, g10, b10, r01, g01, b01, r11, g11, b11;
    tex_texel(L, ix, iy, r00, g00, b00); tex_texel(L, ix + 1, iy, r10, g10, b10);
    tex_texel(L, ix, iy + 1, r01, g01, b01); tex_texel(L, ix + 1, iy + 1, r11, g11, b11);
    const float gx = fsub(1.0f, fx), gy = fsub(1.0f, fy);
    const float w00 = fmul(gx, gy), w10 = fmul(fx, gy), w01 = fmul(gx, fy), w11 = fmul(fx, fy);
    r = fadd(fadd(fadd(fmul(w00, r00), fmul(w10, r10)), fmul(w01, r01)), fmul(w11, r11));
    g = fadd(fadd(fadd(fmul(w00, g00), fmul(w10, g10)), fmul(w01, g01)), fmul(w11, g11));
    b = fadd(fadd(fadd(fmul(w00, b00), fmul(w10, b10)), fmul(w01, b01)), fmul(w11, b11));
}

__device__ __forceinline__ void tex_trilinear(const FrameParams& P, const TexDesc* t, float u, float v, float width, float& r, float& g, float& b)   // shim 23
{
    const int L = (int)__ldg(&t->levels);
    const float level = fadd((float)(L - 1), log2f(fmaxf(width, 1.0e-8f)));
    if (!(level >= 0.0f)) { tex_bilinear(tex_level(P, t, 0), u, v, r, g, b); return; }
    if (level >= (float)(L - 1)) { tex_texel(tex_level(P, t, L - 1), 0, 0, r, g, b); return; }
    const int i = (int)floorf(level);
    const float d = fsub(level, (float)i);
    float r1, g1, b1;
    tex_bilinear(tex_level(P, t, i), u, v, r, g, b);
    tex_bilinear(tex_level(P, t, i + 1), u, v, r1, g1, b1);
    const float e = fsub(1.0f, d);
    r = fadd(fmul(e, r), fmul(d, r1)); g = fadd(fmul(e, g), fmul(d, g1)); b = fadd(fmul(e, b), fmul(d, b1));
}

// Shader.h:234-236: one Sample per pixel with the quad's two differentials
__device__ __forceinline__ void tex_sample(const FrameParams& P, const TexDesc* t, float u, float v, float du0, float dv0, float du1, float dv1,
                                           float& r, float& g, float& b)
{
    if (__ldg(&t->kind) == 0u) { r = __ldg(&t->r); g = __ldg(&t->g); b = __ldg(&t->b); return; }
    const int filter = P.texFilter;
    if (filter == 0) {                                                                     // shim 21
        const TexLevelRef L0 = tex_level(P, t, 0);
        tex_texel(L0, (int)floorf(tex_coord(u, L0.w, 0.0f)), (int)floorf(tex_coord(v, L0.h, 0.0f)), r, g, b);
    } else if (filter == 1) {
        tex_bilinear(tex_level(P, t, 0), u, v, r, g, b);
    } else if (filter == 2) {
        const float width = fmul(2.0f, fmaxf(fmaxf(fabsf(du0), fabsf(dv0)), fmaxf(fabsf(du1), fabsf(dv1))));
        tex_trilinear(P, t, u, v, width, r, g, b);
    } else {                                                                               // shim 24
        const int N = filter == 3 ? 4 : (filter == 4 ? 8 : 16);
        const float l0 = __fsqrt_rn(fadd(fmul(du0, du0), fmul(dv0, dv0))), l1 = __fsqrt_rn(fadd(fmul(du1, du1), fmul(dv1, dv1)));
        const bool first = l0 >= l1;
        const float lmaj = first ? l0 : l1, lmin = first ? l1 : l0;
        const float mu = first ? du0 : du1, mv = first ? dv0 : dv1;
        int n = 1;
        if (lmaj > 0.0f) n = (fmul(lmin, (float)N) <= lmaj) ? N : min(N, max(1, (int)ceilf(fdiv(lmaj, lmin))));
        if (!(lmaj < 3.0e38f)) n = 1;
        const float width = fmul(2.0f, fdiv(lmaj, (float)n));
        float ar = 0.0f, ag = 0.0f, ab = 0.0f;
        for (int i = 0; i < n; i++) {
            const float s = fsub(fdiv(fadd((float)i, 0.5f), (float)n), 0.5f);
            float cr, cg, cb;
            tex_trilinear(P, t, fadd(u, fmul(mu, s)), fadd(v, fmul(mv, s)), width, cr, cg, cb);
            ar = fadd(ar, cr); ag = fadd(ag, cg); ab = fadd(ab, cb);
        }
        const float inv = frcp((float)n);
        r = fmul(ar, inv); g = fmul(ag, inv); b = fmul(ab, inv);
    }
}

// one mip level from the previous one (shim 19): 2x2 box in float, re-quantised like Color4b::FromFloats
__global__ void __launch_bounds__(256) mip_kernel(uchar4* __restrict__ pool, uint32_t srcOff, int sw, int sh, uint32_t dstOff, int dw, int dh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw * dh) return;
    const int x = i % dw, y = i / dw;
    const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    const uchar4 c00 = pool[srcOff + (size_t)y0 * sw + x0], c10 = pool[srcOff + (size_t)y0 * sw + x1];
    const uchar4 c01 = pool[srcOff + (size_t)y1 * sw + x0], c11 = pool[srcOff + (size_t)y1 * sw + x1];
    const float k = 1.0f / 255.0f;
    auto box = [&](uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
        return to_u8(fmul(fadd(fadd(fadd(fmul((float)a, k), fmul((float)b, k)), fmul((float)c, k)), fmul((float)d, k)), 0.25f));
    };
    pool[dstOff + i] = make_uchar4(box(c00.x, c10.x, c01.x, c11.x), box(c00.y, c10.y, c01.y, c11.y), box(c00.z, c10.z, c01.z, c11.z),
                                   box(c00.w, c10.w, c01.w, c11.w));
}

// Shader.h:228-241. The reference shades 2x2 quads (even-aligned, lane k = (x + (k & 1), y + (k >> 1)),
// Rasterizer.h:23) and takes the texture differentials from lanes 1 and 2 against lane 0, whether or not those
// pixels are covered: evaluate this triangle's texcoord there with the same arithmetic, then Sample once for
// this pixel. Not inlined: its registers must not weigh on the untextured shaders.
struct TexQuad { uint32_t B1, C1, B2, C2; int v2x, v2y; float invDet, iw0, iw1, iw2; float tu[3], tv[3]; };

__device__ __noinline__ float3 albedo_textured(const FrameParams& P, const TexQuad& q, uint32_t tri, int px, int py, float b0, float b1, float b2)
{
    auto uv_at = [&](int x, int y, float& u, float& v) {
        const uint32_t ex = (uint32_t)((x << 4) + 8 - q.v2x), ey = (uint32_t)((y << 4) + 8 - q.v2y);
        float q0, q1;
        barycentric((int)(q.B1 * ex + q.C1 * ey), (int)(q.B2 * ex + q.C2 * ey), q.invDet, q0, q1);
        float q2 = fsub(fsub(1.0f, q0), q1);
        q0 = fmul(q0, q.iw0); q1 = fmul(q1, q.iw1); q2 = fmul(q2, q.iw2);
        const float iq = frcp(fadd(fadd(q0, q1), q2));
        q0 = fmul(q0, iq); q1 = fmul(q1, iq);
        q2 = fsub(fsub(1.0f, q0), q1);
        u = blend3(q0, q1, q2, q.tu[0], q.tu[1], q.tu[2]);
        v = blend3(q0, q1, q2, q.tv[0], q.tv[1], q.tv[2]);
    };
    const int qx = px & ~1, qy = py & ~1;
    float u0, v0, u1, v1, u2, v2;
    uv_at(qx, qy, u0, v0); uv_at(qx + 1, qy, u1, v1); uv_at(qx, qy + 1, u2, v2);
    const float uu = blend3(b0, b1, b2, q.tu[0], q.tu[1], q.tu[2]), vv = blend3(b0, b1, b2, q.tv[0], q.tv[1], q.tv[2]);   // Shader.h:167-169
    uint32_t slot = P.texIds ? __ldg(P.texIds + tri) : 0u;
    if (slot >= P.nTex) slot = 0u;
    float3 c;
    tex_sample(P, P.tex + slot, uu, vv, fsub(u1, u0), fsub(v1, v0), fsub(u2, u0), fsub(v2, v0), c.x, c.y, c.z);
    return c;
}

// What interpolation and shading need of the (fan) triangle that owns a pixel: the coefficients of edges 1 and 2,
// vertex 2, 1/det, the three 1/w, and the attributes of its three vertices.
struct OwnerSetup {
    uint32_t B1, C1, B2, C2;
    int v2x, v2y;
    float invDet, iw0, iw1, iw2;
    float A[3][6];                     // position.xyz, normal.xyz of the three (fan) vertices
    float TU[3], TV[3];                // their texture coordinates (only the textured shader reads them)
    uint32_t ok;                       // 0: the owner's record is missing (overflowed frame, re-run by the host)
};

// Re-derive the owning triangle from its prim id (visibility-buffer style: nothing per-triangle was stored for
// triangles that did not go through the clipper): indices -> vertices -> clip / per-vertex records -> setup, or the
// ClipRec of a fan triangle with the attribute re-weighting of Clipper.h:139-147.
template <bool TEX>
__device__ __forceinline__ void derive_owner(const FrameParams& P, uint32_t prim, OwnerSetup& o)
{
    const uint32_t t = prim >> 3, fan = prim & 7u;
    const uint32_t i0 = __ldg(P.i0 + t), i1 = __ldg(P.i1 + t), i2 = __ldg(P.i2 + t);
    const float4 p0 = __ldg(P.pos4 + i0), p1 = __ldg(P.pos4 + i1), p2 = __ldg(P.pos4 + i2);
    const float4 n0 = __ldg(P.nrm4 + i0), n1 = __ldg(P.nrm4 + i1), n2 = __ldg(P.nrm4 + i2);

    int v1x, v1y, v2x, v2y, v0y, v0x;
    float invDet, iw0, iw1, iw2;
    const bool textured = TEX;
    bool unclipped;
    o.ok = 1u;
    if (P.vrec) {
        // front end 2: the owner's vertices were set up once per vertex this frame; gather the records instead of
        // transforming, projecting, snapping and inverting w again
        const int4 r0 = load_vrec(P, i0), r1 = load_vrec(P, i1), r2 = load_vrec(P, i2);
        unclipped = (vrec_code(r0.w) | vrec_code(r1.w) | vrec_code(r2.w)) == 0u;
        if (unclipped) {
            SetupTri s;
            s.v0x = r0.x; s.v0y = r0.y; s.v1x = r1.x; s.v1y = r1.y; s.v2x = r2.x; s.v2y = r2.y;
            finish_setup(s);
            v0x = s.v0x; v0y = s.v0y; v1x = s.v1x; v1y = s.v1y; v2x = s.v2x; v2y = s.v2y; invDet = s.invDet;
            iw0 = __int_as_float(r0.w); iw1 = __int_as_float(r1.w); iw2 = __int_as_float(r2.w);
        }
    } else {
        const V4 c0 = to_clip(P.mvp, p0.x, p0.y, p0.z), c1 = to_clip(P.mvp, p1.x, p1.y, p1.z), c2 = to_clip(P.mvp, p2.x, p2.y, p2.z);
        unclipped = (clip_code(c0) | clip_code(c1) | clip_code(c2)) == 0;
        if (unclipped) {
            SetupTri s;
            setup_tri(P.raster, P.rasterAffineXY != 0, c0, c1, c2, s);
            v0x = s.v0x; v0y = s.v0y; v1x = s.v1x; v1y = s.v1y; v2x = s.v2x; v2y = s.v2y; invDet = s.invDet;
            iw0 = inv_w(c0.w); iw1 = inv_w(c1.w); iw2 = inv_w(c2.w);
        }
    }
    if (unclipped) {
        o.A[0][0] = p0.x; o.A[0][1] = p0.y; o.A[0][2] = p0.z; o.A[0][3] = n0.x; o.A[0][4] = n0.y; o.A[0][5] = n0.z;
        o.A[1][0] = p1.x; o.A[1][1] = p1.y; o.A[1][2] = p1.z; o.A[1][3] = n1.x; o.A[1][4] = n1.y; o.A[1][5] = n1.z;
        o.A[2][0] = p2.x; o.A[2][1] = p2.y; o.A[2][2] = p2.z; o.A[2][3] = n2.x; o.A[2][4] = n2.y; o.A[2][5] = n2.z;
        o.TU[0] = n0.w; o.TU[1] = n1.w; o.TU[2] = n2.w; o.TV[0] = p0.w; o.TV[1] = p1.w; o.TV[2] = p2.w;
    } else {
        const uint32_t recAt = __ldg(P.clipSlot + t) + fan;
        if (recAt >= P.clipRecCap) {               // overflowed frame (re-run by finish_frame): never read past the records
            o.ok = 0u; o.B1 = o.C1 = o.B2 = o.C2 = 0u; o.v2x = o.v2y = 0; o.invDet = o.iw0 = o.iw1 = o.iw2 = 0.0f;
            #pragma unroll
            for (int k = 0; k < 3; k++) { o.TU[k] = o.TV[k] = 0.0f; for (int m = 0; m < 6; m++) o.A[k][m] = 0.0f; }
            return;
        }
        const ClipRec* rp = P.clipRecs + recAt;
        const int4 w0 = __ldg(reinterpret_cast<const int4*>(rp));
        const int4 w1 = __ldg(reinterpret_cast<const int4*>(rp) + 1);
        const float4 w2 = __ldg(reinterpret_cast<const float4*>(rp) + 2);
        const float4 w3 = __ldg(reinterpret_cast<const float4*>(rp) + 3);
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(rp) + 4);
        const float4 w5 = __ldg(reinterpret_cast<const float4*>(rp) + 5);
        v0x = w0.x; v0y = w0.y; v1x = w0.z; v1y = w0.w; v2x = w1.x; v2y = w1.y;
        invDet = __int_as_float(w1.z);
        const uint32_t src = (uint32_t)w1.w;
        iw0 = w2.x; iw1 = w2.y; iw2 = w2.z;
        const float wt[3][3] = { { w3.x, w3.y, w3.z }, { w3.w, w4.x, w4.y }, { w4.z, w4.w, w5.x } };
        const float O[3][6] = { { p0.x, p0.y, p0.z, n0.x, n0.y, n0.z }, { p1.x, p1.y, p1.z, n1.x, n1.y, n1.z }, { p2.x, p2.y, p2.z, n2.x, n2.y, n2.z } };
        #pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t sk = (src >> (2 * k)) & 3u;
            #pragma unroll
            for (int m = 0; m < 6; m++) {
                // Clipper.h:141-146: weight.x*a + weight.y*b + weight.z*c for new vertices
                float blended = blend3(wt[k][0], wt[k][1], wt[k][2], O[0][m], O[1][m], O[2][m]);
                o.A[k][m] = sk == 0 ? O[0][m] : (sk == 1 ? O[1][m] : (sk == 2 ? O[2][m] : blended));
            }
            if (textured) {
                const float bu = blend3(wt[k][0], wt[k][1], wt[k][2], n0.w, n1.w, n2.w), bv = blend3(wt[k][0], wt[k][1], wt[k][2], p0.w, p1.w, p2.w);
                o.TU[k] = sk == 0 ? n0.w : (sk == 1 ? n1.w : (sk == 2 ? n2.w : bu));
                o.TV[k] = sk == 0 ? p0.w : (sk == 1 ? p1.w : (sk == 2 ? p2.w : bv));
            }
        }
    }
    if (!textured) { o.TU[0] = o.TU[1] = o.TU[2] = o.TV[0] = o.TV[1] = o.TV[2] = 0.0f; }
    o.B1 = (uint32_t)v1y - (uint32_t)v2y; o.C1 = (uint32_t)v2x - (uint32_t)v1x;
    o.B2 = (uint32_t)v2y - (uint32_t)v0y; o.C2 = (uint32_t)v0x - (uint32_t)v2x;
    o.v2x = v2x; o.v2y = v2y; o.invDet = invDet; o.iw0 = iw0; o.iw1 = iw1; o.iw2 = iw2;
}

// Stages a15-a17 for one pixel of a known owner: perspective-correct interpolation (Shader.h:142-170), shading
// (Shader.h:185-282) and packing (Renderer.cpp:295-301).
template <bool TEX>
__device__ __forceinline__ uchar4 shade_owned(const FrameParams& P, const OwnerSetup& o, uint32_t prim, int px, int py)
{
    if (!o.ok) return make_uchar4(0, 0, 0, 255);
    const uint32_t dx = (uint32_t)((px << 4) + 8 - o.v2x), dy = (uint32_t)((py << 4) + 8 - o.v2y);
    float b0, b1;
    barycentric((int)(o.B1 * dx + o.C1 * dy), (int)(o.B2 * dx + o.C2 * dy), o.invDet, b0, b1);
    // Fragment::Interpolate, Shader.h:151-159
    float b2 = fsub(fsub(1.0f, b0), b1);
    b0 = fmul(b0, o.iw0); b1 = fmul(b1, o.iw1); b2 = fmul(b2, o.iw2);
    const float invB = frcp(fadd(fadd(b0, b1), b2));
    b0 = fmul(b0, invB); b1 = fmul(b1, invB);
    b2 = fsub(fsub(1.0f, b0), b1);
    const float posx = blend3(b0, b1, b2, o.A[0][0], o.A[1][0], o.A[2][0]);
    const float posy = blend3(b0, b1, b2, o.A[0][1], o.A[1][1], o.A[2][1]);
    const float posz = blend3(b0, b1, b2, o.A[0][2], o.A[1][2], o.A[2][2]);
    float nx = blend3(b0, b1, b2, o.A[0][3], o.A[1][3], o.A[2][3]);
    float ny = blend3(b0, b1, b2, o.A[0][4], o.A[1][4], o.A[2][4]);
    float nz = blend3(b0, b1, b2, o.A[0][5], o.A[1][5], o.A[2][5]);
    // Shader.h:256-264
    float w = rsqrt_exact(dot3(nx, ny, nz, nx, ny, nz));
    nx = fmul(nx, w); ny = fmul(ny, w); nz = fmul(nz, w);
    float dA = dot3(P.light[0], P.light[1], P.light[2], nx, ny, nz);
    if (dA < 0.0f) dA = 0.0f;
    const float diffuse = fmul(fmul(fadd(dA, 0.2f), 3.0f), 0.31830988618f);
    float cr = diffuse, cg = diffuse, cb = diffuse;
    if (P.shader == SH_BLINN_PHONG) {
        // Shader.h:266-280
        float ex = fsub(P.eye[0], posx), ey = fsub(P.eye[1], posy), ez = fsub(P.eye[2], posz);
        w = rsqrt_exact(dot3(ex, ey, ez, ex, ey, ez));
        ex = fmul(ex, w); ey = fmul(ey, w); ez = fmul(ez, w);
        float hx = fadd(P.light[0], ex), hy = fadd(P.light[1], ey), hz = fadd(P.light[2], ez);
        w = rsqrt_exact(dot3(hx, hy, hz, hx, hy, hz));
        hx = fmul(hx, w); hy = fmul(hy, w); hz = fmul(hz, w);
        const float spec = fmul(powf(dot3(nx, ny, nz, hx, hy, hz), 200.0f), 3.0f);
        cr = cg = cb = fadd(diffuse, spec);
    } else if (P.shader == SH_LAMBERT_ALBEDO) {
        float ar = P.albedo[0], ag = P.albedo[1], ab = P.albedo[2];
        if (TEX) {
            const TexQuad q = { o.B1, o.C1, o.B2, o.C2, o.v2x, o.v2y, o.invDet, o.iw0, o.iw1, o.iw2, { o.TU[0], o.TU[1], o.TU[2] }, { o.TV[0], o.TV[1], o.TV[2] } };
            const float3 c = albedo_textured(P, q, prim >> 3, px, py, b0, b1, b2);
            ar = c.x; ag = c.y; ab = c.z;
        }
        cr = fmul(diffuse, ar); cg = fmul(diffuse, ag); cb = fmul(diffuse, ab);
    }
    return make_uchar4(to_u8(cr), to_u8(cg), to_u8(cb), 255);
}

template <bool TEX>
__device__ __forceinline__ uchar4 shade_impl(const FrameParams& P, uint32_t prim, int px, int py)
{
    OwnerSetup o;
    derive_owner<TEX>(P, prim, o);
    return shade_owned<TEX>(P, o, prim, px, py);
}

__device__ __noinline__ uchar4 shade_pixel_textured(const FrameParams& P, uint32_t prim, int px, int py) { return shade_impl<true>(P, prim, px, py); }

__device__ __forceinline__ uchar4 shade_pixel_any(const FrameParams& P, uint32_t prim, int px, int py)
{
    if (P.shader == SH_LAMBERT_ALBEDO && P.nTex != 0u) return shade_pixel_textured(P, prim, px, py);
    return shade_impl<false>(P, prim, px, py);
}

// ---------------------------------------------------------------------------------------------
// shade_kernel: stages a15-a17 of a single-sample frame as a pass of its own over the visibility buffer. tile_kernel
// has written depth and the owning prim id of every pixel. The reference shades every Z-passing fragment and lets
// the last one win (Renderer.cpp:272-350); only the owner is shaded here.
//
// One CTA per 16 x 16 tile, one thread per pixel. Deriving an owner (indices, six vertex attributes, per-vertex or
// clip records, setup) costs more than shading a pixel with it, and a tile usually has far fewer owners than pixels.
// So the CTA first builds the set of distinct owners of its tile in a 64-slot table in shared memory (one lane per
// group of equal ids in a warp inserts it), then thread s derives the owner in slot s - up to 64 DIFFERENT owners in
// one pass of the code - and parks the result in shared memory, and every pixel shades from its owner's slot. A tile
// whose pixels mostly have owners of their own (sub-pixel triangles) skips the table and derives per pixel.
// ---------------------------------------------------------------------------------------------
constexpr int SHADE_SLOTS = 64;

template <bool TEX>
__device__ __forceinline__ void owner_to_smem(const OwnerSetup& o, uint32_t* d)
{
    d[0] = o.B1; d[1] = o.C1; d[2] = o.B2; d[3] = o.C2; d[4] = (uint32_t)o.v2x; d[5] = (uint32_t)o.v2y;
    d[6] = __float_as_uint(o.invDet); d[7] = __float_as_uint(o.iw0); d[8] = __float_as_uint(o.iw1); d[9] = __float_as_uint(o.iw2); d[10] = o.ok;
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        #pragma unroll
        for (int m = 0; m < 6; m++) d[11 + 6 * k + m] = __float_as_uint(o.A[k][m]);
        if (TEX) { d[29 + k] = __float_as_uint(o.TU[k]); d[32 + k] = __float_as_uint(o.TV[k]); }
    }
}

template <bool TEX>
__device__ __forceinline__ void owner_from_smem(OwnerSetup& o, const uint32_t* d)
{
    o.B1 = d[0]; o.C1 = d[1]; o.B2 = d[2]; o.C2 = d[3]; o.v2x = (int)d[4]; o.v2y = (int)d[5];
    o.invDet = __uint_as_float(d[6]); o.iw0 = __uint_as_float(d[7]); o.iw1 = __uint_as_float(d[8]); o.iw2 = __uint_as_float(d[9]); o.ok = d[10];
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        #pragma unroll
        for (int m = 0; m < 6; m++) o.A[k][m] = __uint_as_float(d[11 + 6 * k + m]);
        if (TEX) { o.TU[k] = __uint_as_float(d[29 + k]); o.TV[k] = __uint_as_float(d[32 + k]); }
        else { o.TU[k] = 0.0f; o.TV[k] = 0.0f; }
    }
}

// FROM_KEYS: the frame has no tile_kernel (nothing on the tile path; skip_tile): the pass reads the visibility keys
// itself - depth out, key reset, owner from the key - instead of the ids a resolve kernel would have written for it:
// one kernel and one 4-byte round trip per pixel less in frames of small triangles.
template <bool TEX, bool FROM_KEYS>
__global__ void __launch_bounds__(256) shade_kernel(const __grid_constant__ FrameParams P)
{
    constexpr int OWN_WORDS = TEX ? 37 : 31;                       // odd strides: distinct slots fall into distinct banks
    __shared__ uint32_t sTab[SHADE_SLOTS];
    __shared__ uint32_t sOwn[SHADE_SLOTS * OWN_WORDS];
    __shared__ uint32_t sLeaders;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t tilesX = (uint32_t)(P.width + TILE_PX - 1) >> TILE_LOG2;
    const int tx0 = (int)(blockIdx.x % tilesX) << TILE_LOG2, ty0 = (int)(blockIdx.x / tilesX) << TILE_LOG2;
    const int px = tx0 + (int)(tid & 15u), py = ty0 + (int)(tid >> 4);      // a warp = two rows of the tile
    if (tid < SHADE_SLOTS) sTab[tid] = 0xFFFFFFFFu;
    if (tid == 0) sLeaders = 0;
    cudaGridDependencySynchronize();
    if (!owns_pixel(tx0, ty0, P.binsX, P.part, P.parts)) return;           // (the whole CTA: a tile lies inside one bin)
    const bool inside = px < P.width && py < P.height;
    const size_t at = inside ? (size_t)px + (size_t)P.width * (size_t)(P.height - 1 - py) : 0;
    uint32_t prim = 0xFFFFFFFFu;
    if (FROM_KEYS) {
        if (inside) {
            unsigned long long* kp = P.keys + key_index(px, py, P.binsX);
            const unsigned long long key = *kp;
            if (key != KEY_EMPTY) { *kp = KEY_EMPTY; prim = key_prim(key); }
            P.depth[at] = key != KEY_EMPTY ? key_depth(key) : 1.0f;              // clear value, FrameBuffer.cpp:103
            if (P.captureIds) P.ids[at] = prim;
        }
    } else if (inside) prim = __ldg(P.ids + at);
    const bool hit = prim != 0xFFFFFFFFu;
    const uint32_t group = __match_any_sync(0xFFFFFFFFu, prim);
    const int leader = __ffs(group) - 1;
    const bool isLeader = hit && (int)lane == leader;
    const uint32_t leaders = __ballot_sync(0xFFFFFFFFu, isLeader);
    __syncthreads();
    if (lane == 0 && leaders) atomicAdd(&sLeaders, (uint32_t)__popc(leaders));
    __syncthreads();
    const uint32_t nLeaders = sLeaders;
    if (nLeaders == 0) {                                           // nothing drawn in this tile: cleared colour (FrameBuffer.cpp:91-95)
        if (inside) P.color[at] = make_uchar4(0, 0, 0, 0);
        return;
    }
    uint32_t sl = 0xFFu;
    if (nLeaders <= 96u) {                                         // (at most 96 groups of equal ids in the 8 warps: the table pays)
        if (isLeader) {
            uint32_t h = (prim * 2654435761u) >> 26;
            #pragma unroll 1
            for (int probe = 0; probe < 8; probe++) {
                const uint32_t old = atomicCAS(sTab + h, 0xFFFFFFFFu, prim);
                if (old == 0xFFFFFFFFu || old == prim) { sl = h; break; }
                h = (h + 1u) & (SHADE_SLOTS - 1);
            }
        }
        sl = __shfl_sync(0xFFFFFFFFu, sl, leader);
        if (!hit) sl = 0xFFu;
        __syncthreads();
        if (tid < SHADE_SLOTS) {
            const uint32_t p = sTab[tid];
            if (p != 0xFFFFFFFFu) {
                OwnerSetup o;
                derive_owner<TEX>(P, p, o);
                owner_to_smem<TEX>(o, sOwn + tid * OWN_WORDS);
            }
        }
        __syncthreads();
    }
    if (!inside) return;
    uchar4 c = make_uchar4(0, 0, 0, 0);
    if (hit) {
        OwnerSetup o;
        if (sl != 0xFFu) owner_from_smem<TEX>(o, sOwn + sl * OWN_WORDS);
        else derive_owner<TEX>(P, prim, o);
        c = shade_owned<TEX>(P, o, prim, px, py);
    }
    P.color[at] = c;
}

// ---------------------------------------------------------------------------------------------
// sort_big_kernel: puts the frame's tile-path list in nearest-first order (one CTA; lists shorter than SORT_MIN or
// longer than SORT_MAX stay as appended). key = ordered( min vertex depth - the slack classify_rect allows itself ):
// a lower bound of the near depth classify_rect computes for ANY rect. A counting sort over 1024 depth buckets is
// enough: what tile_kernel needs at position p is a lower bound of the keys of ALL later entries - the lower edge of
// p's bucket - so that a bin whose depth bound is below it can stop reading the list there. Screen-sized triangles
// made every bin classify the whole list (C3: 4000 candidates per bin, two thirds of the tile kernel's time) although
// the nearest few settle it.
// ---------------------------------------------------------------------------------------------
constexpr int SORT_BUCKETS = 1024;

__device__ __forceinline__ uint32_t big_sort_key(const BigRec* r)
{
    const float4 a = __ldg(reinterpret_cast<const float4*>(r) + 1);      // v2x v2y z0 z1
    const float z2 = __ldg(reinterpret_cast<const float*>(r) + 8);
    const float vlo = fminf(a.z, fminf(a.w, z2)), vhi = fmaxf(a.z, fmaxf(a.w, z2));
    if (!(vlo == vlo) || !(vhi == vhi)) return 0u;                        // NaN depths: never used to stop a bin
    const float Ed = 1e-5f * fmaxf(fabsf(vlo), fabsf(vhi)) + 1e-30f;     // as classify_rect
    return min(order_f32(vlo - 2.0f * Ed), 0xFFFFFFFEu);
}

__global__ void __launch_bounds__(1024) sort_big_kernel(const __grid_constant__ FrameParams P)
{
    __shared__ uint32_t sCount[SORT_BUCKETS];
    __shared__ uint32_t sMin, sMax;
    const uint32_t tid = threadIdx.x;
    sCount[tid] = 0;
    if (tid == 0) { sMin = 0xFFFFFFFFu; sMax = 0u; }
    cudaGridDependencySynchronize();
    const uint32_t n = min(P.counters->nBig, P.bigCap);
    if (n < (uint32_t)SORT_MIN || n > (uint32_t)SORT_MAX) { if (tid == 0) P.counters->bigSorted = 0; return; }
    __syncthreads();
    // keys (parked in bigKey), their range
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (uint32_t i = tid; i < n; i += 1024u) { const uint32_t k = big_sort_key(P.big + i); P.bigKey[i] = k; lo = min(lo, k); hi = max(hi, k); }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((tid & 31u) == 0) { atomicMin(&sMin, lo); atomicMax(&sMax, hi); }
    __syncthreads();
    const uint32_t kmin = sMin;
    const unsigned long long range = (unsigned long long)(sMax - kmin) + 1ull;
    // histogram -> exclusive scan -> scatter
    for (uint32_t i = tid; i < n; i += 1024u)
        atomicAdd(&sCount[(uint32_t)(((unsigned long long)(P.bigKey[i] - kmin) * SORT_BUCKETS) / range)], 1u);
    __syncthreads();
    {
        const uint32_t c = sCount[tid];
        uint32_t incl = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((tid & 31u) >= (uint32_t)o) incl += v; }
        __shared__ uint32_t sWarp[32];
        if ((tid & 31u) == 31u) sWarp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            uint32_t w = sWarp[tid];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, w, o); if (tid >= (uint32_t)o) w += v; }
            sWarp[tid] = w;
        }
        __syncthreads();
        sCount[tid] = incl - c + ((tid >> 5) ? sWarp[(tid >> 5) - 1] : 0u);       // first position of bucket tid
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 1024u) {
        const uint32_t k = P.bigKey[i];
        const uint32_t b = (uint32_t)(((unsigned long long)(k - kmin) * SORT_BUCKETS) / range);
        const uint32_t at = atomicAdd(&sCount[b], 1u);
        P.bigOrder[at] = i;
        P.bigBoxSorted[at] = __ldg(P.bigBox + i);
        // every key of bucket b and of the buckets after it is >= this
        P.bigBound[at] = kmin + (uint32_t)(((unsigned long long)b * range) / SORT_BUCKETS);
    }
    if (tid == 0) P.counters->bigSorted = 1;
}

// ---------------------------------------------------------------------------------------------
// Per-bin lists for LONG tile-path lists (stage a7: Renderer::TiledRasterization bins every triangle into the tiles
// its bounding box overlaps, Renderer.cpp:162-229). tile_kernel's bins otherwise all read the whole list (a 4-byte box
// per entry): fine for a few thousand large triangles, quadratic for hundreds of thousands. Four small kernels,
// launched only when the previous frame of an equal-sized mesh had such a list and doing nothing unless this frame's
// is at least P.binMin long:
//   bin_keys_kernel     depth key of every entry (as sort_big_kernel's), their range; zeroes the counters
//   bin_fill_kernel<0>  count per (bin, depth level): level = position of the key in the frame's key range, 16 levels
//   bin_scan_kernel     exclusive scan of the counts -> first slot of every (bin, level) run; decides `binned`
//   bin_fill_kernel<1>  scatter: list entry i goes to the run of every bin of its box, at its level
// A bin's list is then the concatenation of its 16 runs, NEAREST LEVEL FIRST, so what sort_big_kernel's order gives
// short lists holds here too: tile_kernel reads the bin's list a few thousand entries at a time and stops at the first
// level that lies behind the bin's depth bound. Order inside a run is whatever the atomics made it - the frame does
// not depend on it (the visibility key does the ordering), only on the set.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool bin_wanted(const FrameParams& P, uint32_t n) { return P.binMin > 0 && n >= (uint32_t)P.binMin; }
__device__ __forceinline__ uint32_t bin_level(uint32_t k, uint32_t kmin, unsigned long long range)
{
    return (uint32_t)(((unsigned long long)(k - kmin) * BIN_LEVELS) / range);
}

__global__ void __launch_bounds__(256) bin_keys_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    const uint32_t n = min(P.counters->nBig, P.bigCap);
    if (!bin_wanted(P, n)) return;
    const uint32_t stride = gridDim.x * 256u, g = blockIdx.x * 256u + threadIdx.x;
    const uint32_t nRuns = (uint32_t)(P.binsX * P.binsY) * BIN_LEVELS;
    for (uint32_t i = g; i < nRuns; i += stride) P.binCursor[i] = 0;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (uint32_t i = g; i < n; i += stride) { const uint32_t k = big_sort_key(P.big + i); P.binKey[i] = k; lo = min(lo, k); hi = max(hi, k); }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    // (the minimum is kept complemented so that the counters' zero state is its identity)
    if ((threadIdx.x & 31u) == 0 && lo <= hi) { atomicMax(&P.counters->binKeyMin, ~lo); atomicMax(&P.counters->binKeyMax, hi); }
}

template <bool SCATTER>
__global__ void __launch_bounds__(256) bin_fill_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    const uint32_t n = min(P.counters->nBig, P.bigCap);
    if (!bin_wanted(P, n)) return;
    if (SCATTER && P.counters->binned == 0) return;
    const uint32_t kmin = ~P.counters->binKeyMin;
    const unsigned long long range = (unsigned long long)(P.counters->binKeyMax - kmin) + 1ull;
    const uint32_t stride = gridDim.x * 256u;
    unsigned long long pairs = 0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += stride) {
        const uint32_t box = __ldg(P.bigBox + i);
        const uint32_t lvl = bin_level(P.binKey[i], kmin, range);
        const uint32_t x0 = box & 255u, x1 = min((box >> 8) & 255u, (uint32_t)P.binsX - 1u), y0 = (box >> 16) & 255u, y1 = min(box >> 24, (uint32_t)P.binsY - 1u);
        if (x0 > x1 || y0 > y1) continue;
        pairs += (unsigned long long)(x1 - x0 + 1u) * (y1 - y0 + 1u);
        for (uint32_t y = y0; y <= y1; y++)
            for (uint32_t x = x0; x <= x1; x++) {
                uint32_t* run = P.binCursor + (y * (uint32_t)P.binsX + x) * BIN_LEVELS + lvl;
                if (SCATTER) P.binList[atomicAdd(run, 1u)] = i;
                else atomicAdd(run, 1u);
            }
    }
    if (!SCATTER) {
        // the pair total in 64 bits (the 32-bit counts of a frame of a million screen-sized triangles would wrap)
        // (one atomic per CTA: thousands of warps adding to one address serialise in L2)
        __shared__ unsigned long long sPairs;
        if (threadIdx.x == 0) sPairs = 0;
        __syncthreads();
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o);
        if ((threadIdx.x & 31u) == 0 && pairs) atomicAdd(&sPairs, pairs);
        __syncthreads();
        if (threadIdx.x == 0 && sPairs) atomicAdd(&P.counters->binPairs64, sPairs);
    }
}

__global__ void __launch_bounds__(1024) bin_scan_kernel(const __grid_constant__ FrameParams P)
{
    __shared__ uint32_t sWarp[32];
    cudaGridDependencySynchronize();
    const uint32_t n = min(P.counters->nBig, P.bigCap);
    if (!bin_wanted(P, n)) return;
    const uint32_t tid = threadIdx.x;
    const uint32_t nBins = (uint32_t)(P.binsX * P.binsY), nRuns = nBins * BIN_LEVELS;
    const uint32_t per = (nRuns + 1023u) / 1024u, first = min(tid * per, nRuns), last = min(first + per, nRuns);
    uint32_t sum = 0;
    for (uint32_t i = first; i < last; i++) sum += P.binCursor[i];
    uint32_t incl = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((tid & 31u) >= (uint32_t)o) incl += v; }
    if ((tid & 31u) == 31u) sWarp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        uint32_t w = sWarp[tid];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, w, o); if (tid >= (uint32_t)o) w += v; }
        sWarp[tid] = w;
    }
    __syncthreads();
    uint32_t at = incl - sum + ((tid >> 5) ? sWarp[(tid >> 5) - 1] : 0u);
    for (uint32_t i = first; i < last; i++) { const uint32_t c = P.binCursor[i]; P.binCursor[i] = at; at += c; }
    if (tid == 0) {
        // Lists pay when a bin's list is much shorter than the whole list; a frame of screen-sized triangles (every
        // triangle in most bins) keeps the shared list, and so does a frame whose pairs do not fit this frame's
        // allocation (the host sees the demand and the next frame has the room).
        const unsigned long long total = P.counters->binPairs64;
        const bool pays = P.binForce || 4ull * total <= (unsigned long long)n * nBins;
        P.counters->nBinPairs = pays ? (uint32_t)min(total, 0xFFFFFFFFull) : 0u;
        P.counters->binned = (pays && total <= (unsigned long long)P.binListCap) ? 1u : 0u;
    }
}

// ---------------------------------------------------------------------------------------------
struct TileShared {
    unsigned long long keys[KEYS_PER_BIN];     // 32 KB: [tile 4x4][block 2x2][8x8]
    BigRec surv[SURV_CAP];                     // 40 KB
    uint32_t cand[CAND_CAP];                   // 16 KB: candidate indices of the current chunk of the tile-path list ...
    uint32_t candZ[CAND_CAP];                  // 16 KB: ... and their ordered near depth (0xFFFFFFFF = rejected)
    uint32_t survCount, candCount;
    uint32_t binU;                             // order_f32 of the bin's depth upper bound
    uint32_t keyMax;                           // scratch: max ordered depth currently stored in the bin
    uint32_t admitCount, admitNear, admitFar;  // scratch: candidates the bin's bound admits, the nearest and the farthest of them
    unsigned long long keysReady;              // mbarrier: the bulk copy of the bin's keys has landed
};

// --- bulk asynchronous copy (TMA engine, no tensor map: the bin's key block is one contiguous 32 KB run) ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: arm the barrier with the byte count, then hand the copy to the copy engine
__device__ __forceinline__ void bulk_load(void* dstSmem, const void* srcGlobal, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EDX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra EDX_DONE;\n"
        "bra EDX_WAIT;\n"
        "EDX_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// pixel of key slot j (0..7) of this lane inside the warp's tile
__device__ __forceinline__ void slot_pixel(int tx0, int ty0, int lane, int j, int& px, int& py)
{
    const int q = j >> 1, h = j & 1;
    px = tx0 + (q & 1) * BLOCK_PX + (lane & 7);
    py = ty0 + (q >> 1) * BLOCK_PX + (lane >> 3) + 4 * h;
}

// Largest ordered depth held by the on-screen pixels of the warp's tile; 0xFFFFFFFF while any is empty.
__device__ __forceinline__ uint32_t tile_key_max(const unsigned long long (&k)[8], int tx0, int ty0, int W, int H)
{
    const int lane = threadIdx.x & 31;
    uint32_t m = 0;
    #pragma unroll
    for (int j = 0; j < 8; j++) {
        int px, py;
        slot_pixel(tx0, ty0, lane, j, px, py);
        if (px < W && py < H) m = max(m, (uint32_t)(k[j] >> 32));
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    return m;
}

// Rasterise the bin's survivor list: every warp walks the list for its own 16x16 tile, one lane per
// triangle for the tile-level test (exact reject / full cover / hierarchical Z), then the whole warp per
// surviving triangle. The Z bound tightens as the tile fills: it is the smaller of (a) the far side of
// any triangle covering the whole tile and (b) the largest depth currently stored in the tile.
template <bool MS>
__device__ __forceinline__ void raster_survivors(const FrameParams& P, TileShared& S, int ox, int oy, bool hiz, int offX, int offY)
{
    const bool ms = MS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx0 = ox + (warp & 3) * TILE_PX, ty0 = oy + (warp >> 2) * TILE_PX;
    const int n = (int)S.survCount;
    if (tx0 >= P.width || ty0 >= P.height || n == 0) return;
    unsigned long long* tkeys = S.keys + warp * 256;
    unsigned long long k[8];
    #pragma unroll
    for (int j = 0; j < 8; j++) k[j] = tkeys[j * 32 + lane];
    uint32_t U = order_f32(1.0f);
    if (hiz) {
        // Bound pass: the tile's depth upper bound is a pure minimum - the nearest far side of any survivor that
        // covers the whole tile (and what the tile already holds) - so it is taken over the WHOLE list before anything
        // is rasterised. A bound that only tightens as the list is walked makes the work depend on the order the
        // clipper happened to append the triangles in (C3: 550K or 670K surviving pairs, a 3x swing in time).
        U = min(U, tile_key_max(k, tx0, ty0, P.width, P.height));
        uint32_t zhiAll = 0xFFFFFFFFu;
        for (int b = 0; b < n; b += 32) {
            const int j = b + lane;
            if (j < n) {
                bool rej, full; float zl, zh;
                classify_rect(S.surv[j], tx0, ty0, TILE_PX, P.width, P.height, true, ms, min(U, S.binU), rej, full, zl, zh);
                if (!rej && full && zh <= 1.0f) zhiAll = min(zhiAll, order_f32(zh));
            }
        }
        U = min(U, __reduce_min_sync(0xFFFFFFFFu, zhiAll));
    }
    for (int b = 0; b < n; b += 32) {
        if (hiz) U = min(U, tile_key_max(k, tx0, ty0, P.width, P.height));
        const int j = b + lane;
        bool keep = false, full = false;
        uint32_t zlo = 0, zhi = 0xFFFFFFFFu;
        if (j < n) {
            bool rej; float zl, zh;
            classify_rect(S.surv[j], tx0, ty0, TILE_PX, P.width, P.height, hiz, ms, hiz ? min(U, S.binU) : 0xFFFFFFFFu, rej, full, zl, zh);
            keep = !rej;
            if (hiz) { zlo = order_f32(zl); if (keep && full && zh <= 1.0f) zhi = order_f32(zh); }
            if (!P.hierarchical) full = false;
        }
        if (hiz) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) zhi = min(zhi, __shfl_xor_sync(0xFFFFFFFFu, zhi, o));
            U = min(U, zhi);
            keep = keep && zlo <= U;
        }
        uint32_t keepMask = __ballot_sync(0xFFFFFFFFu, keep);
        const uint32_t fullMask = __ballot_sync(0xFFFFFFFFu, keep && full);
        while (keepMask) {
            const int l = __ffs(keepMask) - 1;
            keepMask &= keepMask - 1;
            raster_tile_tri(k, S.surv[b + l], tx0, ty0, P.width, P.height, (fullMask >> l) & 1u, P.hierarchical != 0, ms, offX, offY);
        }
    }
    #pragma unroll
    for (int j = 0; j < 8; j++) tkeys[j * 32 + lane] = k[j];
    __syncwarp();
}

__device__ __forceinline__ bool bin_in_box(uint32_t box, uint32_t bx, uint32_t by)
{
    return bx >= (box & 255u) && bx <= ((box >> 8) & 255u) && by >= ((box >> 16) & 255u) && by <= (box >> 24);
}

__device__ __forceinline__ void load_big(const BigRec* src, BigRec& r)
{
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(&r);
    d4[0] = __ldg(s4); d4[1] = __ldg(s4 + 1); d4[2] = __ldg(s4 + 2); d4[3] = __ldg(s4 + 3);
}

// Output of one pixel: depth always, owner ids on request or when the frame is shaded (stage a13)
__device__ __forceinline__ void resolve_pixel(const FrameParams& P, unsigned long long key, int px, int py)
{
    const size_t at = (size_t)px + (size_t)P.width * (size_t)(P.height - 1 - py);   // bottom-up, FrameBuffer.cpp:41
    const bool hit = key != KEY_EMPTY;
    P.depth[at] = hit ? key_depth(key) : 1.0f;                                      // clear value, FrameBuffer.cpp:103
    if (P.captureIds) P.ids[at] = hit ? key_prim(key) : 0xFFFFFFFFu;       // colour is shade_kernel's job (it reads these ids)
}

// End of frame: a one-thread kernel behind the frame's last kernel publishes the counters to pinned host memory
// (the host checks them for queue overflow) and zeroes them for the next frame, so a frame needs no memset before
// and no copy after it. (It used to be a ticket taken by every tile_kernel CTA - fence, atomic, last one publishes;
// that kept each CTA, i.e. half an SM, alive 2-3 us longer than its work.)
__global__ void frame_end_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    unsigned long long midArea = 0;
    {
        volatile unsigned long long* slot = P.counters->midAreaSlot;
        midArea = slot[threadIdx.x] + slot[threadIdx.x + 32];
        if (midArea) { slot[threadIdx.x] = 0; slot[threadIdx.x + 32] = 0; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) midArea += __shfl_xor_sync(0xFFFFFFFFu, midArea, o);
    }
    if (threadIdx.x != 0) return;
    // every load first (nine independent L2 reads in flight), then the stores from registers; no fence: the host
    // reads the pinned copy only after the stream has completed this kernel
    volatile Counters* d = P.counters;
    const uint32_t nBig = d->nBig, nClip1 = d->nClipQueue, nClipN = d->nClipMulti, nClipRecs = d->nClipRecs, nDump = d->nDump, tilePairs = d->tilePairs, nMid = d->nMid, nMidDiverted = d->nMidDiverted, nBigDiverted = d->nBigDiverted, serial = d->frameSerial + 1u, nBinPairs = d->nBinPairs, binned = d->binned;
    // the clip queue is split in two halves (clip_kernel): publish a demand that exceeds the capacity exactly when a half overflowed
    const uint32_t halfQ = P.clipQueueCap / 2u;
    const uint32_t nClipQueue = (nClip1 > halfQ || nClipN > P.clipQueueCap - halfQ) ? 2u * max(nClip1, nClipN) + 2u : nClip1 + nClipN;
    uint32_t overFrames = d->overFrames;
    const uint32_t maxBig = max(d->maxBig, nBig), maxClipQueue = max(d->maxClipQueue, nClipQueue), maxClipRecs = max(d->maxClipRecs, nClipRecs), maxMid = max(d->maxMid, nMid);
#ifdef EDX_DEBUG_STATS
    for (int i = 0; i < 8; i++) { P.hostCounters->dbg[i] = P.counters->dbg[i]; P.counters->dbg[i] = 0; }
#endif
    if (nBig > P.bigCap || nClipQueue > P.clipQueueCap || nClipRecs > P.clipRecCap || nMid > P.midCap) overFrames++;
    d->nBig = 0; d->nClipQueue = 0; d->nClipMulti = 0; d->nClipRecs = 0; d->nDump = 0; d->nMid = 0; d->tilePairs = 0; d->nWork = 0; d->nMidDiverted = 0; d->nBigDiverted = 0; d->bigSorted = 0; d->frameSerial = serial;
    d->binned = 0; d->binKeyMin = 0; d->binKeyMax = 0; d->nBinPairs = 0; d->binPairs64 = 0;
    d->overFrames = overFrames; d->maxBig = maxBig; d->maxClipQueue = maxClipQueue; d->maxClipRecs = maxClipRecs; d->maxMid = maxMid;
    volatile Counters* h = P.hostCounters;
    h->nBig = nBig; h->nClipQueue = nClipQueue; h->nClipRecs = nClipRecs; h->nDump = nDump; h->tilePairs = tilePairs; h->nMid = nMid; h->maxMid = maxMid; h->nMidDiverted = nMidDiverted; h->nBigDiverted = nBigDiverted; h->frameSerial = serial; h->nBinPairs = nBinPairs; h->binned = binned; h->midArea = midArea;
    h->overFrames = overFrames; h->maxBig = maxBig; h->maxClipQueue = maxClipQueue; h->maxClipRecs = maxClipRecs;
}

// edx_set_frame_sink_signal: runs behind the copy-engine pushes of a finished frame (same stream) and publishes the
// number of frames pushed so far where the consumer - usually another GPU - can poll it.
__global__ void sink_signal_kernel(uint32_t* word, uint32_t serial)
{
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(word) = serial;
    __threadfence_system();
}

// A 16x16 tile whose pixels were all decided by the direct (small-triangle) path: turn its keys, already in
// registers, into depth / ids / colour and leave the key buffer clean for the next frame. One warp per tile.
template <bool DEPTH_ONLY>             // true: the caller guarantees a depth-only frame without id capture (no shading code at all)
__device__ __forceinline__ void resolve_tile_direct(const FrameParams& P, ulonglong2* gk2, const ulonglong2 (&kk)[4], int tx0, int ty0, int lane)
{
    if (tx0 >= P.width || ty0 >= P.height) return;
    const ulonglong2 empty2 = make_ulonglong2(KEY_EMPTY, KEY_EMPTY);
    #pragma unroll
    for (int b4 = 0; b4 < 4; b4++)
        if (kk[b4].x != KEY_EMPTY || kk[b4].y != KEY_EMPTY) gk2[b4 * 32 + lane] = empty2;
    #pragma unroll 1
    for (int b4 = 0; b4 < 4; b4++) {
        // keys 2*lane and 2*lane + 1 of block b4 are two horizontally adjacent pixels
        const int px = tx0 + (b4 & 1) * BLOCK_PX + ((2 * lane) & 7), py = ty0 + (b4 >> 1) * BLOCK_PX + ((2 * lane) >> 3);
        if (py >= P.height || px >= P.width) continue;
        const ulonglong2 kc = kk[b4];      // (dynamic index: kk[] has a local-memory copy; selecting from registers spills more)
        if ((DEPTH_ONLY || (P.shader == SH_DEPTH_ONLY && !P.captureIds)) && px + 1 < P.width) {
            const size_t at = (size_t)px + (size_t)P.width * (size_t)(P.height - 1 - py);
            const float d0 = kc.x != KEY_EMPTY ? key_depth(kc.x) : 1.0f, d1 = kc.y != KEY_EMPTY ? key_depth(kc.y) : 1.0f;
            if ((at & 1) == 0) *reinterpret_cast<float2*>(P.depth + at) = make_float2(d0, d1);
            else { P.depth[at] = d0; P.depth[at + 1] = d1; }
        } else if (DEPTH_ONLY) {
            const size_t at = (size_t)px + (size_t)P.width * (size_t)(P.height - 1 - py);      // last column of an odd width
            P.depth[at] = kc.x != KEY_EMPTY ? key_depth(kc.x) : 1.0f;
        } else {
            resolve_pixel(P, kc.x, px, py);
            if (px + 1 < P.width) resolve_pixel(P, kc.y, px + 1, py);
        }
    }
}

// The same resolve as a kernel of its own, launched before tile_kernel when the previous frame had nothing on the
// tile path (dense meshes of small triangles: C1, C2, most of C4). It does the work only if this frame has nothing
// there either; tile_kernel then finds leanResolve set and just ends the frame. Why: tile_kernel's CTAs own half an
// SM each (512 threads x 64 registers, 97 KB of shared memory) even when all they do is this latency-bound pass, so
// nothing else - in particular the geometry kernel of another frame in flight - can share the SM with them. This
// kernel has no shared memory and small CTAs. MEASURED (C2, B200): one frame at a time 59.5 -> 59.5 us, C1 46 -> 43.8 us,
// but with three frames in flight 45.0 -> 49.0 us - the extra launch and tile_kernel's 510 empty CTAs cost more
// than the co-residency wins. Off by default (`edx_set_option("lean_resolve", 1 | 2)`).
template <bool DEPTH_ONLY>
__global__ void __launch_bounds__(256) lean_resolve_kernel(const __grid_constant__ FrameParams P)
{
    const int lane = threadIdx.x & 31;
    const uint32_t tileG = blockIdx.x * 8u + (threadIdx.x >> 5);
    const uint32_t bin = tileG >> 4, warp = tileG & 15u;
    cudaGridDependencySynchronize();
    if (bin >= (uint32_t)(P.binsX * P.binsY)) return;
    if (P.parts > 1 && bin % (uint32_t)P.parts != (uint32_t)P.part) return;
    const uint32_t bx = bin % (uint32_t)P.binsX, by = bin / (uint32_t)P.binsX;
    const int tx0 = ((int)bx << BIN_LOG2) + (int)(warp & 3u) * TILE_PX, ty0 = ((int)by << BIN_LOG2) + (int)(warp >> 2) * TILE_PX;
    ulonglong2* gk2 = reinterpret_cast<ulonglong2*>(P.keys + (size_t)bin * KEYS_PER_BIN + warp * 256u);
    ulonglong2 kk[4];
    #pragma unroll
    for (int b4 = 0; b4 < 4; b4++) kk[b4] = gk2[b4 * 32 + lane];
    if (P.tileLaunched && min(P.counters->nBig, P.bigCap) != 0) return;      // the tile path has work: tile_kernel does everything
    resolve_tile_direct<DEPTH_ONLY>(P, gk2, kk, tx0, ty0, lane);
}

#ifdef EDX_DEBUG_STATS
__device__ unsigned long long g_binDbg[8192][10];   // per bin: cycles cand / sweep / flush+final raster / resolve, candidates, survivors
__device__ uint32_t g_tileResident[512];       // [sm] live tile_kernel CTAs, [256 + sm] the most seen at once
struct ResidentScope {
    uint32_t sm;
    __device__ ResidentScope() {
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        if (threadIdx.x == 0) atomicMax(&g_tileResident[256u + (sm & 255u)], atomicAdd(&g_tileResident[sm & 255u], 1u) + 1u);
    }
    __device__ ~ResidentScope() { if (threadIdx.x == 0) atomicSub(&g_tileResident[sm & 255u], 1u); }
};
#endif

// MS = false is the single-sample instantiation: sample id and offsets are the constants 0, which keeps them out of
// registers (the kernel is at its 64-register limit; spilled values made its speed depend on the context's
// local-memory layout - C3 ran 0.8 or 1.65 ms depending on which other CUDA modules the process had used).
template <bool MS>
__global__ void __launch_bounds__(TILE_THREADS, 2) tile_kernel(const __grid_constant__ FrameParams P)
{
#ifdef EDX_DEBUG_STATS
    ResidentScope residentScope;
#endif
    extern __shared__ __align__(16) unsigned char smemRaw[];
    TileShared& S = *reinterpret_cast<TileShared*>(smemRaw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int bin = blockIdx.x;
    // MSAA: blockIdx.y is the sample; each (bin, sample) CTA works on that sample's key plane, evaluates
    // coverage and depth at centre + offset, and leaves the keys in place for msaa_resolve_kernel.
    const bool ms = MS;
    const int sId = MS ? (int)blockIdx.y : 0;
    const int offX = MS ? c_sampleOffsets[P.msLevel][2 * sId] : 0, offY = offX;      // DESIGN.md shim 17: Vector2i(int) sets both components
    const uint32_t bx = (uint32_t)(bin % P.binsX), by = (uint32_t)(bin / P.binsX);
    const int ox = (int)bx << BIN_LOG2, oy = (int)by << BIN_LOG2;
    const int tx0 = ox + (warp & 3) * TILE_PX, ty0 = oy + (warp >> 2) * TILE_PX;
    if (tid == 0) mbar_init(&S.keysReady, 1);                 // (before the dependency wait: touches shared memory only)
    __syncthreads();
    cudaGridDependencySynchronize();
    if (P.parts > 1 && (uint32_t)bin % (uint32_t)P.parts != (uint32_t)P.part) {   // sort-first: not this context's bin
        return;
    }
    unsigned long long* gbin = P.keys + (size_t)sId * P.keyStride + (size_t)bin * KEYS_PER_BIN;     // the bin's 4096 keys: one contiguous run
    unsigned long long* gkeys = gbin + warp * 256;                                                   // this warp's tile, [block][8x8]
    if (ms && min(P.counters->nBig, P.bigCap) == 0) return;      // nothing on the tile path: the keys are already final
    if (!MS && P.leanResolve && min(P.counters->nBig, P.bigCap) == 0) return;   // lean_resolve_kernel has resolved the frame
    // The bin's keys are wanted on every path. One thread hands the whole 32 KB block to the copy engine
    // (cp.async.bulk, completion counted on an mbarrier) and the copy runs under the dependent counter read.
    if (tid == 0) bulk_load(S.keys, gbin, KEYS_PER_BIN * (uint32_t)sizeof(unsigned long long), &S.keysReady);
    ulonglong2* gk2 = reinterpret_cast<ulonglong2*>(gkeys);
    const uint32_t nBig = min(P.counters->nBig, P.bigCap);
    const ulonglong2 empty2 = make_ulonglong2(KEY_EMPTY, KEY_EMPTY);
    mbar_wait(&S.keysReady, 0);
    ulonglong2* sk2 = reinterpret_cast<ulonglong2*>(S.keys + warp * 256);

    if (nBig == 0) {
        // Nothing on the tile path: resolve straight from the staged keys (unless lean_resolve_kernel has already
        // done exactly that for this frame).
        if (!P.leanResolve) {
            ulonglong2 kk[4];
            #pragma unroll
            for (int b4 = 0; b4 < 4; b4++) kk[b4] = sk2[b4 * 32 + lane];          // keys b4*64 + 2*lane, +1
            resolve_tile_direct<false>(P, gk2, kk, tx0, ty0, lane);
        }
        return;
    }

    // reset the bin's keys in L2 for the next frame (what the small-triangle paths left there is now in shared memory)
    if (!ms) {
        #pragma unroll
        for (int b4 = 0; b4 < 4; b4++) {
            const ulonglong2 k2 = sk2[b4 * 32 + lane];
            if (k2.x != KEY_EMPTY || k2.y != KEY_EMPTY) gk2[b4 * 32 + lane] = empty2;
        }
    }
    if (tid == 0) { S.survCount = 0; S.candCount = 0; S.binU = order_f32(1.0f); S.keyMax = 0; S.admitCount = 0; S.admitNear = 0xFFFFFFFFu; S.admitFar = 0; }
    __syncthreads();
#ifdef EDX_DEBUG_STATS
    long long tMark = clock64(), tCand = 0, tSweep = 0, tFlush = 0, nFlush = 0, nSurvTot = 0, nIter = 0, nCandTot = 0, tClass = 0, tCount = 0, tSlab = 0, tTop = 0, tFlushSlab = 0;
#endif

    const bool hizOn = P.hiz && P.hierarchical;
    auto flush_bin = [&](uint32_t haveSurv, bool hiz) {
        if (tid == 0) atomicAdd(&P.counters->tilePairs, haveSurv);
#ifdef EDX_DEBUG_STATS
        long long tF = clock64(); nFlush++; nSurvTot += haveSurv;
#endif
        raster_survivors<MS>(P, S, ox, oy, hiz, offX, offY);
        __syncthreads();
#ifdef EDX_DEBUG_STATS
        tFlush += clock64() - tF;
#endif
        if (hiz) {
            // what is now stored in the bin bounds everything still to come
            if (tx0 < P.width && ty0 < P.height) {
                unsigned long long kk[8];
                #pragma unroll
                for (int j = 0; j < 8; j++) kk[j] = S.keys[warp * 256 + j * 32 + lane];
                const uint32_t m = tile_key_max(kk, tx0, ty0, P.width, P.height);
                if (lane == 0) atomicMax(&S.keyMax, m);
            }
            __syncthreads();
            if (tid == 0) { S.binU = min(S.binU, S.keyMax); S.keyMax = 0; }
        }
        if (tid == 0) S.survCount = 0;
        __syncthreads();
    };
    // Nearest-first view of the list (sort_big_kernel): the bin reads it 512 entries at a time and stops at the first
    // position whose key - a lower bound of every later triangle's depth - is behind the bin's depth bound.
    // Long lists come binned (bin_*_kernel): the bin reads its own list, nearest depth level first, instead of
    // filtering the whole list by bin box.
    const bool binned = P.counters->binned != 0;
    const bool sorted = !binned && hizOn && P.counters->bigSorted != 0;
    const uint32_t* boxes = sorted ? P.bigBoxSorted : P.bigBox;
    const uint32_t chunk = sorted ? 512u : 4u * TILE_THREADS;
    const uint32_t* runEnd = P.binCursor + (size_t)bin * BIN_LEVELS;          // binned: end of each of the bin's runs in binList
    const uint32_t listLo = (binned && bin) ? runEnd[-1] : 0u;
    const uint32_t listN = binned ? runEnd[BIN_LEVELS - 1] - listLo : nBig;
    uint32_t cursor = 0;                                       // next entry of the (bin's) tile-path list (uniform)
    while (cursor < listN) {
        if ((sorted || (binned && hizOn)) && cursor) {
            // (uniform: every thread reads the same words after the barrier that closed the previous chunk)
            uint32_t bound;                                          // no later triangle is nearer than this
            if (sorted) bound = __ldg(P.bigBound + cursor);
            else {
                uint32_t l = 0;                                      // depth level of the next unread entry: all later ones are at or behind it
                while (l < (uint32_t)BIN_LEVELS - 1u && runEnd[l] <= listLo + cursor) l++;
                const uint32_t kmin = ~P.counters->binKeyMin;
                bound = kmin + (uint32_t)((((unsigned long long)(P.counters->binKeyMax - kmin) + 1ull) * l) / BIN_LEVELS);
            }
            bool behind = bound > S.binU;
            if (!behind) {
                // not behind the bound yet: the bin may still get bounded by what the admitted triangles draw
                const uint32_t haveSurv = S.survCount;
                __syncthreads();
#ifdef EDX_DEBUG_STATS
                const long long t0 = clock64();
#endif
                if (haveSurv) { flush_bin(haveSurv, true); behind = bound > S.binU; }
#ifdef EDX_DEBUG_STATS
                tTop += clock64() - t0;
#endif
            }
            if (behind) break;
        }
        // 1. candidates: a 4-byte bin box per triangle filters the list before any record is loaded
        for (;;) {
            // The loop decision must be the same for every thread: read the count, then a barrier, so no
            // thread can start appending (and change the count) before all threads have read it.
            const uint32_t have = S.candCount;
            __syncthreads();
            if (!(cursor < listN && have <= (sorted ? 0u : (uint32_t)(CAND_CAP - 4 * TILE_THREADS)))) break;
            const uint32_t i = cursor + 4u * tid;
            uint32_t hits = 0;
            uint32_t box[4];                                         // binned: the four list entries themselves
            if (binned) {
                #pragma unroll
                for (int k = 0; k < 4; k++)
                    if (i + k < listN) { box[k] = P.binList[listLo + i + k]; hits |= 1u << k; }
            } else if (4u * tid < chunk && i < nBig) {
                if (i + 3 < nBig) {
                    const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(boxes + i));
                    box[0] = b4.x; box[1] = b4.y; box[2] = b4.z; box[3] = b4.w;
                } else {
                    #pragma unroll
                    for (int k = 0; k < 4; k++) box[k] = (i + k < nBig) ? __ldg(boxes + i + k) : 0xFFu;   // x0=255 > x1=0: never matches
                }
                hits = (bin_in_box(box[0], bx, by) ? 1u : 0u) | (bin_in_box(box[1], bx, by) ? 2u : 0u) |
                       (bin_in_box(box[2], bx, by) ? 4u : 0u) | (bin_in_box(box[3], bx, by) ? 8u : 0u);
            }
            {
                // warp-aggregated append: one shared-memory atomic per warp instead of one per hit (a frame of
                // screen-sized triangles made every thread of every bin hit the same counter: C3, 19 M bank conflicts)
                const uint32_t cnt = (uint32_t)__popc(hits);
                uint32_t incl = cnt;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                uint32_t base = 0;
                if (lane == 31 && total) base = atomicAdd(&S.candCount, total);
                base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - cnt;
                #pragma unroll
                for (int k = 0; k < 4; k++)
                    if (hits & (1u << k)) S.cand[base++] = binned ? box[k] : i + k;
            }
            cursor += chunk;
            __syncthreads();
        }
#ifdef EDX_DEBUG_STATS
        tCand += clock64() - tMark; tMark = clock64();
#endif
        const uint32_t nCand = S.candCount;
#ifdef EDX_DEBUG_STATS
        nIter++; nCandTot += nCand;
#endif
        const bool hiz = hizOn && nCand >= HIZ_MIN_CAND;
        // 3a. classify every candidate against the bin: exact reject, and (hierarchical Z) its conservative near depth;
        //     a candidate that covers the whole bin lowers the bin's depth upper bound S.binU to its far side. The bound
        //     is a pure minimum over the candidates, so it is complete before anything is admitted (3b): which
        //     triangles survive does not depend on the order the list was appended in.
        for (uint32_t base = 0; base < nCand; base += TILE_THREADS) {
            const uint32_t j = base + tid;
            BigRec r;
            bool rej, full; float zl, zh;
            // (threads past the end classify the last candidate again and drop the result: uniform control flow
            // keeps the record in registers)
            { const uint32_t pos = S.cand[min(j, nCand - 1u)]; load_big(P.big + (sorted ? __ldg(P.bigOrder + pos) : pos), r); }
            classify_rect(r, ox, oy, BIN, P.width, P.height, hiz, ms, hiz ? S.binU : 0xFFFFFFFFu, rej, full, zl, zh);
            rej = rej || j >= nCand;
            if (hiz) {
                const uint32_t u = (!rej && full && zh <= 1.0f) ? order_f32(zh) : 0xFFFFFFFFu;
                const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, u);
                if (lane == 0 && m != 0xFFFFFFFFu) atomicMin(&S.binU, m);      // one shared-memory atomic per warp
            }
            if (j < nCand) S.candZ[j] = rej ? 0xFFFFFFFFu : (hiz ? min(order_f32(zl), 0xFFFFFFFEu) : 0u);
        }
        __syncthreads();
        // 3b. admit what can still be visible; survivors go to shared memory and are rasterised whenever the list fills
        //     up. A bin that no candidate covers completely (the diagonal of a screen-sized quad's two fan triangles)
        //     gets no bound from 3a; what bounds it is the depth already drawn. So a long admission list is walked
        //     NEAREST FIRST, in eight slabs of near depth with a raster flush after each: once the near triangles are
        //     drawn, the bin's (and in raster_survivors each tile's) stored depth culls the slabs behind them - again
        //     whatever order the list was appended in.
#ifdef EDX_DEBUG_STATS
        tClass += clock64() - tMark;
#endif
        auto flush = [&](uint32_t haveSurv) { flush_bin(haveSurv, hiz); };
        uint32_t nSlabs = 1, zNear = 0, zFar = 0xFFFFFFFEu;
        if (hiz) {
            const uint32_t U0 = S.binU;
            uint32_t cnt = 0, zmin = 0xFFFFFFFFu, zmax = 0u;
            for (uint32_t j = tid; j < nCand; j += TILE_THREADS) {
                const uint32_t z = S.candZ[j];
                if (z <= U0) { cnt++; zmin = min(zmin, z); zmax = max(zmax, z); }
            }
            cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
            zmin = __reduce_min_sync(0xFFFFFFFFu, zmin);
            zmax = __reduce_max_sync(0xFFFFFFFFu, zmax);
            if (lane == 0 && cnt) { atomicAdd(&S.admitCount, cnt); atomicMin(&S.admitNear, zmin); atomicMax(&S.admitFar, zmax); }
            __syncthreads();
            // The slabs divide the depth range the admitted candidates actually occupy. (They used to run up to the bin's
            // bound - 1.0 where nothing covers the bin - and the few hundred triangles of a diagonal bin, whose depths
            // lie within a sliver of that range, all fell into the first slab: 280 admitted and rasterised per pixel
            // where the nearest handful settles the bin - those bins took 8x the mean and set the kernel's tail.)
            if (S.admitCount > (uint32_t)(SURV_CAP - TILE_THREADS)) { nSlabs = 8; zNear = S.admitNear; zFar = S.admitFar; }
            __syncthreads();
            if (tid == 0) { S.admitCount = 0; S.admitNear = 0xFFFFFFFFu; S.admitFar = 0; }
        }
#ifdef EDX_DEBUG_STATS
        tCount += clock64() - tMark;
        const long long tFlushBefore = tFlush;
#endif
        for (uint32_t slab = 0; slab < nSlabs; slab++) {
            const uint32_t span = zFar - zNear;
            const uint32_t lo = slab == 0 ? 0u : zNear + (uint32_t)(((unsigned long long)span * slab) / nSlabs) + 1u;
            const uint32_t hi = slab + 1 == nSlabs ? 0xFFFFFFFEu : zNear + (uint32_t)(((unsigned long long)span * (slab + 1)) / nSlabs);
            for (uint32_t base = 0; base < nCand; base += TILE_THREADS) {
                const uint32_t haveSurv = S.survCount;               // same rule: read, barrier, then decide
                __syncthreads();
                if (haveSurv > SURV_CAP - TILE_THREADS) flush(haveSurv);
                const uint32_t j = base + tid;
                const uint32_t z = j < nCand ? S.candZ[j] : 0xFFFFFFFFu;
                const bool keep = z >= lo && z <= hi && z <= (hiz ? S.binU : 0xFFFFFFFEu);
                const uint32_t keepMask = __ballot_sync(0xFFFFFFFFu, keep);
                uint32_t at = 0;
                if (lane == 0 && keepMask) at = atomicAdd(&S.survCount, (uint32_t)__popc(keepMask));     // one shared-memory atomic per warp
                at = __shfl_sync(0xFFFFFFFFu, at, 0) + (uint32_t)__popc(keepMask & ((1u << lane) - 1u));
                if (keep) {
                    const uint32_t pos = S.cand[j];
                    const int4* s2 = reinterpret_cast<const int4*>(P.big + (sorted ? __ldg(P.bigOrder + pos) : pos));
                    int4* d2 = reinterpret_cast<int4*>(&S.surv[at]);
                    d2[0] = __ldg(s2); d2[1] = __ldg(s2 + 1); d2[2] = __ldg(s2 + 2); d2[3] = __ldg(s2 + 3);
                }
                __syncthreads();
            }
            if (nSlabs > 1 && slab + 1 < nSlabs) {
                const uint32_t haveSurv = S.survCount;
                __syncthreads();
                if (haveSurv) flush(haveSurv);
            }
        }
        if (tid == 0) S.candCount = 0;
        __syncthreads();
#ifdef EDX_DEBUG_STATS
        tSlab += clock64() - tMark; tFlushSlab += tFlush - tFlushBefore;
        tSweep += clock64() - tMark; tMark = clock64();
#endif
    }
#ifdef EDX_DEBUG_STATS
    nSurvTot += S.survCount;
#endif
    if (tid == 0) atomicAdd(&P.counters->tilePairs, S.survCount);      // load of the tile path, for the host's tuning
    raster_survivors<MS>(P, S, ox, oy, hizOn && S.survCount >= HIZ_MIN_CAND, offX, offY);
#ifdef EDX_DEBUG_STATS
    __syncthreads();
    if (tid == 0) {
        const long long tFinal = clock64() - tMark;
        atomicAdd(&P.counters->dbg[0], (unsigned long long)tCand); atomicAdd(&P.counters->dbg[1], (unsigned long long)(tSweep - tFlush));
        atomicAdd(&P.counters->dbg[2], (unsigned long long)tFlush); atomicAdd(&P.counters->dbg[3], (unsigned long long)tFinal);
        atomicAdd(&P.counters->dbg[4], (unsigned long long)nFlush); atomicAdd(&P.counters->dbg[5], (unsigned long long)nSurvTot);
        atomicAdd(&P.counters->dbg[6], 1ull);
        if (bin < 8192) { g_binDbg[bin][0] = tCand; g_binDbg[bin][1] = tSweep - tFlush; g_binDbg[bin][2] = tFlush + tFinal; g_binDbg[bin][5] = nSurvTot; g_binDbg[bin][4] = (unsigned long long)(nIter * 1000000ll + nFlush * 10000ll) + (unsigned long long)min(nCandTot, 9999ll) + ((unsigned long long)tClass << 32); g_binDbg[bin][6] = tCount; g_binDbg[bin][7] = tSlab; g_binDbg[bin][8] = tTop; g_binDbg[bin][9] = tFlushSlab; }
    }
    tMark = clock64();
#endif
    __syncwarp();

    if (ms) {
        // hand the sample's keys back; msaa_resolve_kernel shades, averages and ends the frame
        const unsigned long long* tkeys = S.keys + warp * 256;
        #pragma unroll
        for (int j = 0; j < 8; j++) gkeys[j * 32 + lane] = tkeys[j * 32 + lane];
        return;
    }
    // resolve: every warp finishes its own tile; each pixel is written to HBM exactly once
    if (tx0 < P.width && ty0 < P.height) {
        const unsigned long long* tkeys = S.keys + warp * 256;
        #pragma unroll 1
        for (int j = 0; j < 8; j++) {
            const int q = j >> 1, h = j & 1;
            const int px = tx0 + (q & 1) * BLOCK_PX + (lane & 7), py = ty0 + (q >> 1) * BLOCK_PX + (lane >> 3) + 4 * h;
            if (px < P.width && py < P.height) resolve_pixel(P, tkeys[q * 64 + lane + 32 * h], px, py);
        }
    }
#ifdef EDX_DEBUG_STATS
    __syncthreads();
    if (tid == 0) { atomicAdd(&P.counters->dbg[7], (unsigned long long)(clock64() - tMark)); if (bin < 8192) g_binDbg[bin][3] = clock64() - tMark; }
#endif
}

// ---------------------------------------------------------------------------------------------
// msaa_resolve_kernel: last kernel of a multi-sample frame (Renderer::UpdateFrameBuffer per sample,
// Renderer.cpp:309-345, then FrameBuffer::Resolve, FrameBuffer.cpp:70-87). One thread per pixel walks the
// pixel's S keys: the owner of each sample is shaded ONCE PER FRAGMENT at the pixel centre (the reference
// shades a fragment once and writes that colour to every sample it covers, Rasterizer.h:273-288), samples
// of the same triangle reuse the colour; the box filter sums Color(byte)/255 in sample order and rounds
// with FromFloats. Keys are reset for the next frame; per-sample depth and owner ids are kept for parity.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msaa_resolve_kernel(const __grid_constant__ FrameParams P)
{
    cudaGridDependencySynchronize();
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;            // key index inside one sample plane
    // invert key_index: [bin][tile 4x4][block 2x2][8x8]
    const uint32_t bin = i >> 12, tile = (i >> 8) & 15u, block = (i >> 6) & 3u, in = i & 63u;
    const int px = (int)((bin % (uint32_t)P.binsX) * BIN + (tile & 3u) * TILE_PX + (block & 1u) * BLOCK_PX + (in & 7u));
    const int py = (int)((bin / (uint32_t)P.binsX) * BIN + (tile >> 2) * TILE_PX + (block >> 1) * BLOCK_PX + (in >> 3));
    if (i < P.keyStride && px < P.width && py < P.height && owns_pixel(px, py, P.binsX, P.part, P.parts)) {
        const size_t at = (size_t)px + (size_t)P.width * (size_t)(P.height - 1 - py);
        const size_t plane = (size_t)P.width * P.height;
        float acc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
        uint32_t lastPrim = 0xFFFFFFFFu;
        uchar4 lastColor = make_uchar4(0, 0, 0, 0);
        for (int sId = 0; sId < P.samples; sId++) {
            unsigned long long* kp = P.keys + (size_t)sId * P.keyStride + i;
            const unsigned long long key = *kp;
            const bool hit = key != KEY_EMPTY;
            if (hit) *kp = KEY_EMPTY;
            P.depth[sId * plane + at] = hit ? key_depth(key) : 1.0f;
            if (P.captureIds) P.ids[sId * plane + at] = hit ? key_prim(key) : 0xFFFFFFFFu;
            if (hit && P.shader != SH_DEPTH_ONLY) {
                const uint32_t prim = key_prim(key);
                if (prim != lastPrim) { lastColor = shade_pixel_any(P, prim, px, py); lastPrim = prim; }
                acc[0] = fadd(acc[0], fmul((float)lastColor.x, 1.0f / 255.0f));
                acc[1] = fadd(acc[1], fmul((float)lastColor.y, 1.0f / 255.0f));
                acc[2] = fadd(acc[2], fmul((float)lastColor.z, 1.0f / 255.0f));
                acc[3] = fadd(acc[3], fmul((float)lastColor.w, 1.0f / 255.0f));
            }
        }
        if (P.shader != SH_DEPTH_ONLY) {
            const float inv = frcp((float)P.samples);
            P.color[at] = make_uchar4(to_u8(fmul(acc[0], inv)), to_u8(fmul(acc[1], inv)), to_u8(fmul(acc[2], inv)), to_u8(fmul(acc[3], inv)));
        }
    }
}


// Diagnostic: how many CTAs of tile_kernel's shape (512 threads, sizeof(TileShared) dynamic shared memory) does an SM
// really hold at once? perSm[sm] = live CTAs, perSm[256 + sm] = the most seen.
__global__ void __launch_bounds__(TILE_THREADS, 2) residency_kernel(uint32_t* perSm)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    if (threadIdx.x == 0) {
        uint32_t sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        const uint32_t now = atomicAdd(&perSm[sm & 255u], 1u) + 1u;
        atomicMax(&perSm[256u + (sm & 255u)], now);
        unsigned long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        const long long t0 = clock64();
        while (clock64() - t0 < 100000) { }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        perSm[512] = (uint32_t)(g1 - g0);                   // ns that 100,000 SM cycles took
        atomicSub(&perSm[sm & 255u], 1u);
        smemRaw[0] = 1;
    }
}


} // namespace edx
