// edx_device.cuh — device-side types and arithmetic of the raster hot path (sm_100a).
//
// Every expression that feeds coverage, the depth test or the depth value is written with explicit
// round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn never contract into FMA) and
// wrapping unsigned integer arithmetic, in the operation order of the reference
// (/root/reference/EDXRaster/Core/*.h, cited per function) so results are bit-identical to its SSE path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edx {

// ---------------------------------------------------------------------------------------------
// Layout constants
// ---------------------------------------------------------------------------------------------
constexpr int BIN_LOG2 = 6;              // a CTA owns a 64x64-pixel bin ...
constexpr int BIN = 1 << BIN_LOG2;
constexpr int TILE_LOG2 = 4;             // ... one warp per 16x16 tile ...
constexpr int TILE_PX = 1 << TILE_LOG2;
constexpr int BLOCK_PX = 8;              // ... tested as four 8x8 blocks, then pixels
constexpr int TILES_PER_BIN = (BIN / TILE_PX) * (BIN / TILE_PX);   // 16
constexpr int KEYS_PER_BIN = BIN * BIN;                            // 4096
constexpr int TILE_THREADS = TILES_PER_BIN * 32;                   // 512
constexpr int SURV_CAP = 640;            // per-bin survivor list held in shared memory (flushed to the rasteriser when it passes SURV_CAP - 512)
constexpr int CAND_CAP = 4096;           // per-bin candidate indices (bin-box filter hits) held in shared memory
constexpr int SORT_MIN = 1024, SORT_MAX = 65536;   // tile-path lists of this length are put in nearest-first order (sort_big_kernel)
constexpr int BIN_LEVELS = 16;           // depth levels a bin's list is ordered by (nearest first), uniform over the frame's key range
constexpr int HIZ_MIN_CAND = 48;         // below this many candidates a bin skips hierarchical Z

constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;

enum Shader { SH_DEPTH_ONLY = 0, SH_BLINN_PHONG = 1, SH_LAMBERT = 2, SH_LAMBERT_ALBEDO = 3 };

// Post-setup triangle routed to the tile path (64 bytes, four 16-byte loads).
// (gx, gy, zref): the triangle's depth plane over pixel centres, Z(px, py) = zref + gx*px + gy*py, fitted
// in fp32 at record creation; perr bounds the fit's error anywhere on screen. Used only for the
// conservative hierarchical-Z bounds, never for a depth value.
struct __align__(16) BigRec {
    int v0x, v0y, v1x, v1y;
    int v2x, v2y; float z0, z1;
    float z2, invDet; uint32_t prim; float gx;
    float gy, zref, perr; uint32_t pad;
};

// Post-setup triangle routed to the warp-per-triangle path (48 bytes, three 16-byte loads): its pixel-centre box is
// larger than the direct path takes and smaller than midMax on both sides.
struct __align__(16) MidRec {
    int v0x, v0y, v1x, v1y;
    int v2x, v2y; float z0, z1;
    float z2, invDet; uint32_t prim; uint32_t pad;
};

// What the resolve pass needs to shade a pixel owned by a fan triangle of a clipped polygon (96 bytes).
struct __align__(16) ClipRec {
    int v0x, v0y, v1x, v1y;
    int v2x, v2y; float invDet; uint32_t src;      // src: 2 bits per vertex, 0..2 = original vertex, 3 = blended
    float invW0, invW1, invW2; uint32_t valid;
    float wt[3][3];                                 // clip weights of the three fan vertices (Clipper.h:16-17)
    float pad[3];
};

// One straddling triangle handed from geom_kernel to clip_kernel: its clip-space vertices travel with it
// (64 bytes, one round trip) so the clipper does not have to chase index -> position -> transform again.
struct __align__(16) ClipItem {
    float c[12];             // c0.xyzw, c1.xyzw, c2.xyzw
    uint32_t tri; uint32_t pad[3];
};

// Debug dump of stages a3-a6 (edx_debug_raster_triangles).
struct DumpRec { int i[7]; float f[7]; };

struct Counters {
    uint32_t nBig;           // triangles appended to the tile path
    uint32_t nClipQueue;     // straddling triangles queued for the clipper
    uint32_t nClipRecs;      // fan-triangle records written by the clipper
    uint32_t nDump;
    uint32_t nMid;           // triangles appended to the warp-per-triangle path
    // Never reset on the device: frames whose queues overflowed, and the largest demand seen. A caller may
    // submit several frames before the next synchronising call; that call learns from these whether any of
    // them (not just the last) was incomplete.
    uint32_t overFrames;
    uint32_t maxBig, maxClipQueue, maxClipRecs;
    uint32_t tilePairs;      // (triangle, bin) pairs that survived the bin-level culls this frame: the tile path's load
    uint32_t maxMid;
    uint32_t bigSorted;      // this frame's tile-path list has a nearest-first order (sort_big_kernel)
    uint32_t nMidDiverted;   // mid-size triangles sent down the tile path because mid_kernel was not launched this frame
    uint32_t nBigDiverted;   // large triangles sent down the warp-per-triangle path because tile_kernel was not launched this frame
    uint32_t frameSerial;    // frames completed on this context (never reset)
    uint32_t nClipMulti;     // straddlers of several planes (back half of the clip queue; nClipQueue counts the single-plane front half)
    uint32_t nWork;          // list front end: triangle clusters that survived cull_kernel this frame
    // per-bin lists of the tile path (bin_*_kernel): valid this frame, range of the depth keys, (triangle, bin) pairs wanted
    uint32_t binned, binKeyMin, binKeyMax, nBinPairs;      // (binKeyMin is kept complemented: zero is its identity)
    unsigned long long binPairs64;
    unsigned long long midArea;  // pixels spanned by the boxes of the triangles mid_kernel rasterised this frame (host copy: the sum)
    unsigned long long midAreaSlot[64];   // ... accumulated in 64 slots: ten thousand warps adding to ONE address cost mid_kernel 14 us
    unsigned long long dbg[8];   // EDX_DEBUG_STATS builds: summed per-CTA cycle counts of the tile kernel's phases
};

// One texture slot of a mesh (Utils/Mesh.h:23; DESIGN.md shims 19-24): a constant colour, or an RGBA8 image with
// its 2x2 box-filtered mip chain stored level after level in the mesh's texel pool.
struct TexDesc {
    uint32_t kind;           // 0 ConstantTexture2D, 1 ImageTexture
    float r, g, b;           // the constant colour
    uint32_t w, h, levels, pad;
    uint32_t off[16];        // first texel of each level in the pool
};

struct FrameParams {
    float mvp[16];           // Core/Renderer.cpp:90
    float raster[16];        // Core/Renderer.cpp:91
    float eye[3];            // Core/Renderer.cpp:289
    float light[3];          // normalised (1,1,-1), Core/Renderer.cpp:290 + Shader.h:258
    float albedo[3];
    int width, height, binsX, binsY;
    int shader, smallMax, smallMaxClip, hiz, hierarchical, captureIds, dump;
    int midMax;              // pixel-centre boxes below this (and not small) are rasterised one warp per triangle; 0 = none
    int midLaunched;         // mid_kernel is part of this frame (else mid-size triangles take the tile path)
    int tileLaunched;        // tile_kernel is part of this frame (else large triangles take the warp-per-triangle path and lean_resolve_kernel ends the frame)
    int part, parts;         // sort-first split: this context owns the bins b with b % parts == part (parts = 1: all)
    int clusterCull;         // skip whole 256-triangle clusters whose bounding box is outside one clip plane
    int fuseClip;            // clip single-plane straddlers inside geom_kernel instead of queueing them
    int msLevel, samples;    // Renderer::SetMSAAMode (Renderer.cpp:94-98): samples = 1 << msLevel
    uint32_t keyStride;      // keys per sample plane
    int leanResolve;         // lean_resolve_kernel runs before tile_kernel this frame
    int rasterAffineXY;      // raster matrix has no z column and w' == 1: skip the unused z/w and w' arithmetic
    // Front end of the frame: 0 = geom_kernel, one CTA per 256-triangle cluster, vertex work per
    // triangle corner; 1 = cull_kernel compacts the clusters that survive the frustum test into `workList` and a
    // persistent geom_list_kernel walks it; 2 = as 1 plus vertex_kernel: stages a1/a2/a5/a6 once per VERTEX
    // (Renderer.cpp:120-127,139-147) into `vrec`, which the geometry kernel and the resolve pass then gather.
    int frontEnd;
    const float4* vclusterBox;           // per 256-vertex cluster: object-space AABB, built at upload
    uint32_t nTriClusters, nVertClusters;
    uint32_t* workList;                  // surviving triangle clusters (any order: submission order lives in the prim id)
    uint32_t* vcFlag;                    // per vertex cluster: 0 = processed by vertex_kernel, else a clip-plane bit every vertex of it is outside of
    int4* vrec;                          // per vertex: snapped x, y (28.4), z*invW, invW - or 0, 0, 0, VREC_OUTSIDE | clip code
    // mesh (SoA streams built at upload)
    const float4* pos4;      // x, y, z, texcoord.v
    const float4* nrm4;      // nx, ny, nz, texcoord.u
    const uint32_t* i0; const uint32_t* i1; const uint32_t* i2;
    const float4* clusterBox;            // per 256-triangle cluster: object-space AABB (min.xyz, max.xyz), built at upload
    uint32_t nTris, nVerts;
    // Mesh::mTextures / GetTextureIds (Utils/Mesh.h:23,54-59); nTex == 0: the constant `albedo`
    const TexDesc* tex; const uchar4* texels; const uint32_t* texIds;
    uint32_t nTex; int texFilter;        // RenderStates::TexFilter (RenderStates.h:23,60)
    // frame state
    unsigned long long* keys;            // 64-bit visibility keys, bin/tile/block-tiled, L2 resident
    MidRec* mid; uint32_t midCap;
    BigRec* big; uint32_t bigCap;
    uint32_t* bigBox;                    // per tile-path triangle: its bin bounding box, 4 x u8 (x0, x1, y0, y1)
    // nearest-first view of the tile-path list (sort_big_kernel): position -> record index, bin box, and a lower bound of
    // the depth any triangle at or after the position can produce (bigKey: scratch, keys in list order)
    uint32_t* bigOrder; uint32_t* bigKey; uint32_t* bigBoxSorted; uint32_t* bigBound;
    // Per-bin lists of the tile path for LONG lists (a7: Renderer.cpp:162-229 bins every triangle into the tiles its box
    // overlaps): count -> exclusive scan -> scatter over (bin, depth level); binCursor[bin * BIN_LEVELS + level] is the END
    // of that run of binList after the scatter (= the start of the next run), binKey the depth key of every list entry.
    uint32_t* binCursor; uint32_t* binList; uint32_t* binKey; uint32_t binListCap;
    int binMin;                          // lists at least this long are binned (0 = never)
    int binForce;                        // tests: bin even where the shared list would be read as cheaply
    ClipItem* clipQueue; uint32_t clipQueueCap;
    ClipRec* clipRecs; uint32_t clipRecCap;
    uint32_t* clipSlot;                  // per submitted triangle: first ClipRec of its polygon
    Counters* counters;
    Counters* hostCounters;              // pinned, device-mapped: frame_end_kernel publishes the counters here
    uchar4* color; float* depth; uint32_t* ids;
    DumpRec* dumpBuf; uint32_t dumpCap;
};

struct V4 { float x, y, z, w; };

// Per-vertex record of a vertex outside the frustum: its invW word is a quiet NaN whose low six bits are the clip code
// (never zero here). A real invW is 1/w of a vertex with code 0; if that is NaN it is stored as 0x7FFFFFFF.
constexpr uint32_t VREC_OUTSIDE = 0x7FC00000u;
__device__ __forceinline__ uint32_t vrec_code(int w)
{
    const uint32_t b = (uint32_t)w;
    return (b & 0x7FFFFFC0u) == VREC_OUTSIDE ? (b & 63u) : 0u;
}

// ---------------------------------------------------------------------------------------------
// fp32 helpers that can never be contracted
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// 1.0f / x: the correctly rounded reciprocal IS the correctly rounded quotient of 1 and x (IEEE 754), in about half the
// instructions of the general division
__device__ __forceinline__ float frcp(float x) { return __frcp_rn(x); }
// ((a*x + b*y) + c*z) + d : the row . vector order of Matrix::TransformPoint (SURVEY.md §8c shim 1/2)
__device__ __forceinline__ float row_dot(float a, float b, float c, float d, float x, float y, float z)
{
    return fadd(fadd(fadd(fmul(a, x), fmul(b, y)), fmul(c, z)), d);
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
}
// b0*a0 + b1*a1 + b2*a2, left to right (Shader.h:161-169)
__device__ __forceinline__ float blend3(float b0, float b1, float b2, float a0, float a1, float a2)
{
    return fadd(fadd(fmul(b0, a0), fmul(b1, a1)), fmul(b2, a2));
}

// Stage a1: DefaultVertexShader::Execute, Core/Shader.h:40-49 (w_in = 1)
__device__ __forceinline__ V4 to_clip(const float* M, float x, float y, float z)
{
    V4 r;
    r.x = row_dot(M[0], M[1], M[2], M[3], x, y, z);
    r.y = row_dot(M[4], M[5], M[6], M[7], x, y, z);
    r.z = row_dot(M[8], M[9], M[10], M[11], x, y, z);
    r.w = row_dot(M[12], M[13], M[14], M[15], x, y, z);
    return r;
}

// Stage a2: Clipper::ComputeClipCode, Core/Clipper.h:48-68
enum { LEFT_BIT = 1, RIGHT_BIT = 2, BOTTOM_BIT = 4, TOP_BIT = 8, NEAR_BIT = 16, FAR_BIT = 32 };
// true  =>  clip_code(v) == 0 (for NaN components it returns false and the caller computes the code)
__device__ __forceinline__ bool surely_inside(const V4& v)
{
    return fabsf(v.x) <= v.w && fabsf(v.y) <= v.w && v.z >= 0.0f && v.z <= v.w;
}
// 1/w of Renderer.cpp:144 (1.0f / 1.0f == 1.0f exactly)
__device__ __forceinline__ float inv_w(float w) { return w == 1.0f ? 1.0f : frcp(w); }
__device__ __forceinline__ uint32_t clip_code(const V4& v)
{
    uint32_t c = 0;
    if (v.x < -v.w) c |= LEFT_BIT;
    if (v.x > v.w) c |= RIGHT_BIT;
    if (v.y < -v.w) c |= BOTTOM_BIT;
    if (v.y > v.w) c |= TOP_BIT;
    if (v.z > v.w) c |= FAR_BIT;
    if (v.z < 0.0f) c |= NEAR_BIT;
    return c;
}

// `int = float * 16.0` (RasterTriangle.h:35-40): exact scale, truncate toward zero; anything outside
// int32 (or NaN) becomes the x86 integer-indefinite value the reference's cvttsd2si produces.
__device__ __forceinline__ int snap_28_4(float f)
{
    float s = fmul(f, 16.0f);
    if (!(s >= -2147483648.0f && s < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(s);
}

struct SetupTri {
    int v0x, v0y, v1x, v1y, v2x, v2y;
    float invDet;
};

// Vector4::HomogeneousProject + Matrix::TransformPoint(Vector3, raster) + snap of x and y
// (Clipper.h:161-163, RasterTriangle.h:29-40).
// affineXY (decided on the host): the raster matrix has zero z-column entries in rows 0/1 and its last
// row is (0,0,0,1). Then z/w only ever contributes +-0 to x and y and w' is exactly 1, so the three
// operations are skipped; x and y are bit-identical for every finite input.
__device__ __forceinline__ void project_snap(const float* R, bool affineXY, const V4& c, int& sx, int& sy)
{
    // x / 1.0f == x exactly, so orthographic / pre-projected geometry (w == 1) skips the IEEE divisions
    const bool unitW = c.w == 1.0f;
    const float ax = unitW ? c.x : fdiv(c.x, c.w), ay = unitW ? c.y : fdiv(c.y, c.w);
    float x, y;
    if (affineXY) {
        x = fadd(fadd(fmul(R[0], ax), fmul(R[1], ay)), R[3]);
        y = fadd(fadd(fmul(R[4], ax), fmul(R[5], ay)), R[7]);
    } else {
        const float az = fdiv(c.z, c.w);
        x = row_dot(R[0], R[1], R[2], R[3], ax, ay, az);
        y = row_dot(R[4], R[5], R[6], R[7], ax, ay, az);
        const float w = row_dot(R[12], R[13], R[14], R[15], ax, ay, az);
        if (w != 1.0f) { x = fdiv(x, w); y = fdiv(y, w); }
    }
    sx = snap_28_4(x);
    sy = snap_28_4(y);
}

// det, back-face / degenerate cull and 1/det from the snapped vertices (RasterTriangle.h:42-60)
__device__ __forceinline__ bool finish_setup(SetupTri& s)
{
    uint32_t B1 = (uint32_t)s.v1y - (uint32_t)s.v2y, C1 = (uint32_t)s.v2x - (uint32_t)s.v1x;
    uint32_t B2 = (uint32_t)s.v2y - (uint32_t)s.v0y, C2 = (uint32_t)s.v0x - (uint32_t)s.v2x;
    int det = (int)(C2 * B1 - C1 * B2);
    if (det <= 0) return false;
    s.invDet = frcp(__int2float_rn(det));
    return true;
}

// Stage a5: RasterTriangle::Setup, RasterTriangle.h:27-60. false = culled (det <= 0).
__device__ __forceinline__ bool setup_tri(const float* R, bool affineXY, const V4& c0, const V4& c1, const V4& c2, SetupTri& s)
{
    project_snap(R, affineXY, c0, s.v0x, s.v0y);
    project_snap(R, affineXY, c1, s.v1x, s.v1y);
    project_snap(R, affineXY, c2, s.v2x, s.v2y);
    return finish_setup(s);
}

// Edge equations of one triangle, ready for per-pixel evaluation.
// Fill rule: the SSE TopLeftEdge of RasterTriangle.h:296-299 (mask bit-cast to -1).
struct Edges {
    uint32_t B0, C0, B1, C1, B2, C2;
    int v0x, v0y, v1x, v1y, v2x, v2y;
    int bias0, bias1, bias2;

    __device__ __forceinline__ static int top_left(int ax, int ay, int bx, int by)
    {
        return ((by > ay) || (ay == by && ax > bx)) ? -1 : 0;
    }
    __device__ __forceinline__ void init(int a0x, int a0y, int a1x, int a1y, int a2x, int a2y)
    {
        v0x = a0x; v0y = a0y; v1x = a1x; v1y = a1y; v2x = a2x; v2y = a2y;
        B0 = (uint32_t)v0y - (uint32_t)v1y; C0 = (uint32_t)v1x - (uint32_t)v0x;     // RasterTriangle.h:42-47
        B1 = (uint32_t)v1y - (uint32_t)v2y; C1 = (uint32_t)v2x - (uint32_t)v1x;
        B2 = (uint32_t)v2y - (uint32_t)v0y; C2 = (uint32_t)v0x - (uint32_t)v2x;
        bias0 = top_left(v0x, v0y, v1x, v1y);
        bias1 = top_left(v1x, v1y, v2x, v2y);
        bias2 = top_left(v2x, v2y, v0x, v0y);
    }
    // biased edge functions at a sub-pixel position (RasterTriangle.h:301-312)
    __device__ __forceinline__ int e0(int px, int py) const { return (int)(B0 * (uint32_t)(px - v0x) + C0 * (uint32_t)(py - v0y) + (uint32_t)bias0); }
    __device__ __forceinline__ int e1(int px, int py) const { return (int)(B1 * (uint32_t)(px - v1x) + C1 * (uint32_t)(py - v1y) + (uint32_t)bias1); }
    __device__ __forceinline__ int e2(int px, int py) const { return (int)(B2 * (uint32_t)(px - v2x) + C2 * (uint32_t)(py - v2y) + (uint32_t)bias2); }
};

// Barycentrics and depth from the UNBIASED edge values of edges 1 and 2
// (TriangleSSE::CalcBarycentricCoord / GetDepth, RasterTriangle.h:324-337)
__device__ __forceinline__ void barycentric(int raw1, int raw2, float invDet, float& l0, float& l1)
{
    l0 = fmul(__int2float_rn(raw1), invDet);
    l1 = fmul(__int2float_rn(raw2), invDet);
}
__device__ __forceinline__ float depth_at(float l0, float l1, float z0, float z1, float z2)
{
    float l2 = fsub(fsub(1.0f, l0), l1);
    return fadd(fadd(fmul(l0, z0), fmul(l1, z1)), fmul(l2, z2));
}

// ---------------------------------------------------------------------------------------------
// 64-bit visibility key. Sequential LESS_EQUAL testing with immediate write
// (FrameBuffer.cpp:54-68) leaves, per pixel, the minimum depth and — among fragments at that
// depth — the LAST one in submission order. min() over
// (orderedDepth << 32) | ~prim reproduces that in any processing order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(float d, uint32_t prim)
{
    uint32_t b = __float_as_uint(fadd(d, 0.0f));               // -0 -> +0, so equal depths tie exactly
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - prim);
}
__device__ __forceinline__ float key_depth(unsigned long long k)
{
    uint32_t b = (uint32_t)(k >> 32);
    b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint32_t key_prim(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)k; }

// key address of pixel (x, y): [bin][tile 4x4][block 2x2][8x8]
__device__ __forceinline__ uint32_t key_index(int x, int y, int binsX)
{
    uint32_t bin = (uint32_t)(y >> BIN_LOG2) * (uint32_t)binsX + (uint32_t)(x >> BIN_LOG2);
    uint32_t tile = (uint32_t)(((y >> TILE_LOG2) & 3) * 4 + ((x >> TILE_LOG2) & 3));
    uint32_t block = (uint32_t)(((y >> 3) & 1) * 2 + ((x >> 3) & 1));
    return ((bin * 16u + tile) * 4u + block) * 64u + (uint32_t)((y & 7) * 8 + (x & 7));
}

// FrameBuffer::MultiSampleOffsets (FrameBuffer.cpp:107-191) as written: pairs in 1/16 pixel relative to the pixel
// centre, one row per SetMSAAMode level. The reference reads ONE int per sample, [2 * sampleId], into a Vector2i
// (Rasterizer.h:247,382), so sample s sits at (table[2s], table[2s]) and the second column is unused — DESIGN.md
// shim 17, established by compiling the reference itself.
__constant__ int c_sampleOffsets[6][64] = {
    { 0, 0 },
    { 4, 4, -4, -4 },
    { -2, -6, 6, -2, -6, 2, 2, 6 },
    { 1, -3, -1, 3, 5, 1, -3, -5, -5, 5, -7, -1, 3, 7, 7, -7 },
    { 1, 1, -1, -3, -3, 2, 4, -1, -5, -2, 2, 5, 5, 3, 3, -5, -2, 6, 0, -7, -4, -6, -6, 4, -8, 0, 7, -4, 6, 7, -7, -8 },
    { 1, 1, -1, -3, -3, 2, 4, -1, -5, -2, 2, 5, 5, 3, 3, -5, -2, 6, 0, -7, -4, -6, -6, 4, -8, 0, 7, -4, 6, 7, -7, -8,
      1, 3, -3, -3, -3, 0, 6, -2, -7, -1, 3, 4, 7, 3, 3, -6, -2, 7, 0, -4, -2, -5, -7, 6, -8, 3, 4, -1, 2, 7, 4, -8 },
};

// sort-first split: does this context own the 64x64 bin that holds pixel (x, y)?
__device__ __forceinline__ bool owns_pixel(int x, int y, int binsX, int part, int parts)
{
    return parts == 1 || (uint32_t)((y >> BIN_LOG2) * binsX + (x >> BIN_LOG2)) % (uint32_t)parts == (uint32_t)part;
}

// pixel-centre range covered by a snapped bounding box: centres sit at 16*i + 8 (Rasterizer.h:23)
__device__ __forceinline__ int first_centre(int lo) { return (lo + 7) >> 4; }     // ceil((lo - 8) / 16)
__device__ __forceinline__ int last_centre(int hi) { return (hi - 8) >> 4; }      // floor((hi - 8) / 16)
// Pixel range that can hold a covered SAMPLE: with MSAA the samples of pixel p lie anywhere in
// [16p, 16p + 15], so the range is the one the reference walks (Rasterizer.h:211-214); at 1x it is the
// tighter pixel-centre range.
__device__ __forceinline__ int first_pixel(int lo, bool ms) { return ms ? (lo >> 4) : first_centre(lo); }
__device__ __forceinline__ int last_pixel(int hi, bool ms) { return ms ? (hi >> 4) : last_centre(hi); }

__device__ __forceinline__ int min3i(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int max3i(int a, int b, int c) { return max(a, max(b, c)); }

} // namespace edx
