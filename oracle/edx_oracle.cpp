// edx_oracle.cpp — TEST INFRASTRUCTURE ONLY (never linked, imported or called by the product).
//
// CPU restatement of the EDXRaster raster hot path (behindthepixels/EDXRaster), written from
// the reference's sources as the parity oracle for the CUDA implementation in edxraster_b200/.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this library.
//
// PARITY UNPINNED: the reference has no tests, golden images or known-answer vectors
//  and cannot be compiled here (its EDXUtil dependency is absent and it is
// MSVC/Win32-only). Every EDXUtil semantic this file needs is DEFINED in
// the "EDXUtil shim" section below and listed in DESIGN.md; each definition cites the call site
// it was inferred from. All file:line citations are relative to /root/reference/EDXRaster/.
//
// Stage structure follows Core/Renderer.cpp:100-118 (RenderMesh):
//   VertexProcessing -> Clipping -> TiledRasterization -> FragmentProcessing -> UpdateFrameBuffer
// with the reference's 32x32 tiles, four 16x16 coarse quadrants and 2x2 SSE quads.
//
// Build: g++ -O2 -msse4.1 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
//   -DORC_FAST_RSQRT selects rsqrtps + one Newton step (timing build); default is the exact
//   1/sqrt definition used for parity.

#include <smmintrin.h>
#include <omp.h>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <climits>
#include <vector>
#include <chrono>
#include <algorithm>

namespace orc {

// ------------------------------------------------------------------------------------------
// EDXUtil shim. Plain fp32, no FMA contraction, x86 semantics.
// ------------------------------------------------------------------------------------------
struct V2i { int x, y; };
struct V2f { float x, y; };
struct V3f { float x, y, z; };
struct V4f { float x, y, z, w; };
struct Mat4 { float m[4][4]; };   // m[row][col], column-vector convention (Renderer.cpp:90 MVP = P*MV)

static inline V3f operator*(float s, const V3f& v) { return { s * v.x, s * v.y, s * v.z }; }
static inline V3f operator+(const V3f& a, const V3f& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static inline V2f operator*(float s, const V2f& v) { return { s * v.x, s * v.y }; }
static inline V2f operator+(const V2f& a, const V2f& b) { return { a.x + b.x, a.y + b.y }; }

// shim 4: Matrix operator* (Renderer.cpp:90). Row-times-column, summed left to right.
static Mat4 mat_mul(const Mat4& a, const Mat4& b)
{
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = ((a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j]) + a.m[i][3] * b.m[3][j];
    return r;
}

// shim 4: Matrix::Inverse (Renderer.cpp:88). Cofactor expansion through 2x2 sub-determinants.
// Host-only; it feeds the eye position (Renderer.cpp:289) and nothing on the coverage path.
static Mat4 mat_inverse(const Mat4& a)
{
    const float (*m)[4] = a.m;
    float s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1];
    float s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
    float s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3];
    float s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
    float s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3];
    float s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
    float c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    float id = 1.0f / det;
    Mat4 r;
    r.m[0][0] = ((m[1][1] * c5 - m[1][2] * c4) + m[1][3] * c3) * id;
    r.m[0][1] = ((-m[0][1] * c5 + m[0][2] * c4) - m[0][3] * c3) * id;
    r.m[0][2] = ((m[3][1] * s5 - m[3][2] * s4) + m[3][3] * s3) * id;
    r.m[0][3] = ((-m[2][1] * s5 + m[2][2] * s4) - m[2][3] * s3) * id;
    r.m[1][0] = ((-m[1][0] * c5 + m[1][2] * c2) - m[1][3] * c1) * id;
    r.m[1][1] = ((m[0][0] * c5 - m[0][2] * c2) + m[0][3] * c1) * id;
    r.m[1][2] = ((-m[3][0] * s5 + m[3][2] * s2) - m[3][3] * s1) * id;
    r.m[1][3] = ((m[2][0] * s5 - m[2][2] * s2) + m[2][3] * s1) * id;
    r.m[2][0] = ((m[1][0] * c4 - m[1][1] * c2) + m[1][3] * c0) * id;
    r.m[2][1] = ((-m[0][0] * c4 + m[0][1] * c2) - m[0][3] * c0) * id;
    r.m[2][2] = ((m[3][0] * s4 - m[3][1] * s2) + m[3][3] * s0) * id;
    r.m[2][3] = ((-m[2][0] * s4 + m[2][1] * s2) - m[2][3] * s0) * id;
    r.m[3][0] = ((-m[1][0] * c3 + m[1][1] * c1) - m[1][2] * c0) * id;
    r.m[3][1] = ((m[0][0] * c3 - m[0][1] * c1) + m[0][2] * c0) * id;
    r.m[3][2] = ((-m[3][0] * s3 + m[3][1] * s1) - m[3][2] * s0) * id;
    r.m[3][3] = ((m[2][0] * s3 - m[2][1] * s1) + m[2][2] * s0) * id;
    return r;
}

// shim 1: Matrix::TransformPoint(Vector4, M) (Shader.h:45): full 4x4, row . vector, left to right.
static inline V4f transform_point4(const Mat4& M, float x, float y, float z, float w)
{
    V4f r;
    r.x = ((M.m[0][0] * x + M.m[0][1] * y) + M.m[0][2] * z) + M.m[0][3] * w;
    r.y = ((M.m[1][0] * x + M.m[1][1] * y) + M.m[1][2] * z) + M.m[1][3] * w;
    r.z = ((M.m[2][0] * x + M.m[2][1] * y) + M.m[2][2] * z) + M.m[2][3] * w;
    r.w = ((M.m[3][0] * x + M.m[3][1] * y) + M.m[3][2] * z) + M.m[3][3] * w;
    return r;
}

// shim 2: Matrix::TransformPoint(Vector3, M) (RasterTriangle.h:30-32, Renderer.cpp:289):
// same sum with w = 1; divide by w' only when w' != 1.
static inline V3f transform_point3(const Mat4& M, const V3f& p)
{
    float x = ((M.m[0][0] * p.x + M.m[0][1] * p.y) + M.m[0][2] * p.z) + M.m[0][3];
    float y = ((M.m[1][0] * p.x + M.m[1][1] * p.y) + M.m[1][2] * p.z) + M.m[1][3];
    float z = ((M.m[2][0] * p.x + M.m[2][1] * p.y) + M.m[2][2] * p.z) + M.m[2][3];
    float w = ((M.m[3][0] * p.x + M.m[3][1] * p.y) + M.m[3][2] * p.z) + M.m[3][3];
    if (w != 1.0f) { x = x / w; y = y / w; z = z / w; }
    return { x, y, z };
}

// shim 3: Vector4::HomogeneousProject (Clipper.h:161-163,178-180): true division.
static inline V3f homogeneous_project(const V4f& v) { return { v.x / v.w, v.y / v.w, v.z / v.w }; }

// shim 16: `int = float * 16.0` (RasterTriangle.h:35-40). The product is formed in double
// (exact) and truncated toward zero; out-of-range / NaN gives the x86 "integer indefinite".
static inline int snap_28_4(float f)
{
    double d = (double)f * 16.0;
    if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
    return (int)d;
}

// shim 10: Math::Normalize (Shader.h:258) and the eye position (Renderer.cpp:289) are per-frame
// constants; they are computed once here and handed to the shader.
static inline V3f normalize3(const V3f& v)
{
    float len = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
    return { v.x / len, v.y / len, v.z / len };
}

// shim 13: Color4b::FromFloats (Renderer.cpp:296-299): clamp to [0,1], *255, +0.5, truncate; a = 255.
static inline uint8_t to_u8(float c)
{
    float t = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    float s = t * 255.0f + 0.5f;
    if (!(s >= 0.0f)) return 0;      // NaN -> 0 (low byte of the x86 integer indefinite)
    return (uint8_t)(int)s;
}

// shim 9: SSE::Rsqrt (Shader.h:256,267,271)
static inline __m128 rsqrt4(__m128 x)
{
#ifdef ORC_FAST_RSQRT
    __m128 r = _mm_rsqrt_ps(x);
    __m128 h = _mm_mul_ps(_mm_set1_ps(0.5f), x);
    return _mm_mul_ps(r, _mm_sub_ps(_mm_set1_ps(1.5f), _mm_mul_ps(h, _mm_mul_ps(r, r))));
#else
    return _mm_div_ps(_mm_set1_ps(1.0f), _mm_sqrt_ps(x));
#endif
}

static const float kInvPi = 0.31830988618f;   // shim 11: Math::EDX_INV_PI

// ------------------------------------------------------------------------------------------
// Records (Shader.h:13-20, RasterTriangle.h:11-24, Tile.h:15-30, Shader.h:105-114)
// ------------------------------------------------------------------------------------------
struct ProjVertex {
    V4f   proj;
    float invW;
    V3f   position;
    V3f   normal;
    V2f   uv;
};

struct RasterTri {
    V2i v0, v1, v2;
    int B0, C0, B1, C1, B2, C2;
    float invDet;
    uint32_t vId0, vId1, vId2, coreId;
    uint32_t primId;                       // ours: submitted triangle * 8 + fan index
    uint8_t rej0, rej1, rej2, acc0, acc1, acc2;

    // scalar top-left bias, RasterTriangle.h:155-158 (used only by conservative corner tests)
    static inline int tl_scalar(const V2i& a, const V2i& b)
    {
        return ((b.y > a.y) || (a.y == b.y && a.x > b.x)) ? 0 : -1;
    }
    // RasterTriangle.h:160-171; int arithmetic wraps (computed in unsigned)
    inline int edge0(int px, int py) const { return (int)((uint32_t)B0 * (uint32_t)(px - v0.x) + (uint32_t)C0 * (uint32_t)(py - v0.y) + (uint32_t)tl_scalar(v0, v1)); }
    inline int edge1(int px, int py) const { return (int)((uint32_t)B1 * (uint32_t)(px - v1.x) + (uint32_t)C1 * (uint32_t)(py - v1.y) + (uint32_t)tl_scalar(v1, v2)); }
    inline int edge2(int px, int py) const { return (int)((uint32_t)B2 * (uint32_t)(px - v2.x) + (uint32_t)C2 * (uint32_t)(py - v2.y) + (uint32_t)tl_scalar(v2, v0)); }
};

static inline void pick_corners(int B, int C, uint8_t& rej, uint8_t& acc)
{
    // RasterTriangle.h:66-150
    float slope = (float)B / (float)C;
    if (slope >= 0.0f) {
        if (C >= 0) { rej = 3; acc = 0; } else { rej = 0; acc = 3; }
    } else {
        if (C >= 0) { rej = 2; acc = 1; } else { rej = 1; acc = 2; }
    }
}

// RasterTriangle.h:27-153
static bool setup_triangle(RasterTri& t, const Mat4& raster, V3f a, V3f b, V3f c,
                           uint32_t i0, uint32_t i1, uint32_t i2, uint32_t core, uint32_t prim)
{
    a = transform_point3(raster, a);
    b = transform_point3(raster, b);
    c = transform_point3(raster, c);
    t.v0.x = snap_28_4(a.x); t.v0.y = snap_28_4(a.y);
    t.v1.x = snap_28_4(b.x); t.v1.y = snap_28_4(b.y);
    t.v2.x = snap_28_4(c.x); t.v2.y = snap_28_4(c.y);
    t.B0 = (int)((uint32_t)t.v0.y - (uint32_t)t.v1.y);
    t.C0 = (int)((uint32_t)t.v1.x - (uint32_t)t.v0.x);
    t.B1 = (int)((uint32_t)t.v1.y - (uint32_t)t.v2.y);
    t.C1 = (int)((uint32_t)t.v2.x - (uint32_t)t.v1.x);
    t.B2 = (int)((uint32_t)t.v2.y - (uint32_t)t.v0.y);
    t.C2 = (int)((uint32_t)t.v0.x - (uint32_t)t.v2.x);
    int det = (int)((uint32_t)t.C2 * (uint32_t)t.B1 - (uint32_t)t.C1 * (uint32_t)t.B2);
    if (det <= 0) return false;
    t.invDet = 1.0f / (float)det;          // :60 (det > 0 so |det| == det)
    t.vId0 = i0; t.vId1 = i1; t.vId2 = i2; t.coreId = core; t.primId = prim;
    pick_corners(t.B0, t.C0, t.rej0, t.acc0);
    pick_corners(t.B1, t.C1, t.rej1, t.acc1);
    pick_corners(t.B2, t.C2, t.rej2, t.acc2);
    return true;
}

struct TriRef {              // Tile.h:15-30
    uint32_t triId;
    bool acc0, acc1, acc2, trivialAccept, big;
};

struct Fragment {            // Shader.h:105-114 (single-sample coverage only)
    __m128 l0, l1;
    uint32_t vId0, vId1, vId2, coreId;
    uint32_t primId;
    uint16_t x, y;
    uint32_t mask[4];        // CoverageMask (Shader.h:52-103): bit (4*sample + lane) of a 128-bit mask
};

struct Tile {                // Tile.h:10-41
    V2i minC, maxC;
    std::vector<std::vector<TriRef>> refs;   // one list per core (F8: sized dynamically)
    std::vector<Fragment> frags;
};

struct Stats {
    uint64_t nRasterTris, nFragments, nCoveredSamples, nBinRefs;
    double ms[6];            // vertex, clip+setup, bin, raster, shade, fbupdate
};

// ------------------------------------------------------------------------------------------
// Texture2D<Color> (EDXUtil Graphics/Texture.h, absent): DESIGN.md shims 19-24. ImageTexture<Color, Color4b>
// keeps 8-bit texels and returns float colours (Mesh.cpp:27); ConstantTexture2D returns its colour (:29,47,66).
// ------------------------------------------------------------------------------------------
struct TexLevel { int w = 0, h = 0; std::vector<uint8_t> px; };      // RGBA8, row-major, row 0 = v 0
struct Texture {
    int kind = 0;                      // 0 constant, 1 image
    float color[3] = { 0, 0, 0 };
    std::vector<TexLevel> levels;      // shim 19: 2x2 box-filtered chain down to 1x1
};

static const float kInv255 = 1.0f / 255.0f;        // shim 18: Color(Color4b) = byte * (1/255)

static void build_mips(Texture& t)                 // shim 19
{
    while (t.levels.back().w > 1 || t.levels.back().h > 1) {
        const TexLevel& s = t.levels.back();
        TexLevel d;
        d.w = std::max(1, s.w >> 1); d.h = std::max(1, s.h >> 1);
        d.px.resize((size_t)d.w * d.h * 4);
        for (int y = 0; y < d.h; y++)
            for (int x = 0; x < d.w; x++) {
                const int x0 = std::min(2 * x, s.w - 1), x1 = std::min(2 * x + 1, s.w - 1);
                const int y0 = std::min(2 * y, s.h - 1), y1 = std::min(2 * y + 1, s.h - 1);
                for (int c = 0; c < 4; c++) {
                    const float c00 = s.px[4 * ((size_t)y0 * s.w + x0) + c] * kInv255, c10 = s.px[4 * ((size_t)y0 * s.w + x1) + c] * kInv255;
                    const float c01 = s.px[4 * ((size_t)y1 * s.w + x0) + c] * kInv255, c11 = s.px[4 * ((size_t)y1 * s.w + x1) + c] * kInv255;
                    d.px[4 * ((size_t)y * d.w + x) + c] = to_u8((((c00 + c10) + c01) + c11) * 0.25f);
                }
            }
        t.levels.push_back(std::move(d));
    }
}

static inline int tex_wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }       // shim 20: repeat

static inline void tex_texel(const TexLevel& L, int x, int y, float out[3])
{
    const uint8_t* p = &L.px[4 * ((size_t)tex_wrap(y, L.h) * L.w + tex_wrap(x, L.w))];
    out[0] = p[0] * kInv255; out[1] = p[1] * kInv255; out[2] = p[2] * kInv255;
}

// texel-space coordinate of a uv component; anything a float -> int conversion cannot hold becomes 0
static inline float tex_coord(float u, int n, float bias)
{
    const float x = u * (float)n - bias;
    return fabsf(x) < 1.0e9f ? x : 0.0f;
}

static void tex_bilinear(const TexLevel& L, float u, float v, float out[3])                  // shim 22
{
    const float x = tex_coord(u, L.w, 0.5f), y = tex_coord(v, L.h, 0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix = (int)x0, iy = (int)y0;
    float c00[3], c10[3], c01[3], c11[3];
    tex_texel(L, ix, iy, c00); tex_texel(L, ix + 1, iy, c10); tex_texel(L, ix, iy + 1, c01); tex_texel(L, ix + 1, iy + 1, c11);
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
    for (int c = 0; c < 3; c++) out[c] = ((w00 * c00[c] + w10 * c10[c]) + w01 * c01[c]) + w11 * c11[c];
}

static void tex_trilinear(const Texture& t, float u, float v, float width, float out[3])    // shim 23
{
    const int L = (int)t.levels.size();
    const float level = (float)(L - 1) + log2f(std::max(width, 1.0e-8f));
    if (!(level >= 0.0f)) { tex_bilinear(t.levels[0], u, v, out); return; }
    if (level >= (float)(L - 1)) { tex_texel(t.levels[L - 1], 0, 0, out); return; }
    const int i = (int)floorf(level);
    const float d = level - (float)i;
    float a[3], b[3];
    tex_bilinear(t.levels[i], u, v, a);
    tex_bilinear(t.levels[i + 1], u, v, b);
    for (int c = 0; c < 3; c++) out[c] = (1.0f - d) * a[c] + d * b[c];
}

// Texture2D::Sample(uv, differentials) under RenderStates' filter (Shader.h:234-236; Renderer.h:48)
static void tex_sample(const Texture& t, int filter, float u, float v, float du0, float dv0, float du1, float dv1, float out[3])
{
    if (t.kind == 0) { out[0] = t.color[0]; out[1] = t.color[1]; out[2] = t.color[2]; return; }
    const TexLevel& L0 = t.levels[0];
    if (filter == 0) {                                                                          // shim 21: nearest
        tex_texel(L0, (int)floorf(tex_coord(u, L0.w, 0.0f)), (int)floorf(tex_coord(v, L0.h, 0.0f)), out);
    } else if (filter == 1) {
        tex_bilinear(L0, u, v, out);
    } else if (filter == 2) {
        const float width = 2.0f * std::max(std::max(fabsf(du0), fabsf(dv0)), std::max(fabsf(du1), fabsf(dv1)));
        tex_trilinear(t, u, v, width, out);
    } else {                                                                                    // shim 24: N taps along the major axis
        const int N = filter == 3 ? 4 : (filter == 4 ? 8 : 16);
        const float l0 = sqrtf(du0 * du0 + dv0 * dv0), l1 = sqrtf(du1 * du1 + dv1 * dv1);
        const bool first = l0 >= l1;
        const float lmaj = first ? l0 : l1, lmin = first ? l1 : l0;
        const float mu = first ? du0 : du1, mv = first ? dv0 : dv1;
        int n = 1;
        if (lmaj > 0.0f) n = (lmin * (float)N <= lmaj) ? N : std::min(N, std::max(1, (int)ceilf(lmaj / lmin)));
        if (!(lmaj < 3.0e38f)) n = 1;                                   // inf / NaN footprints: one tap
        const float width = 2.0f * (lmaj / (float)n);
        float acc[3] = { 0.0f, 0.0f, 0.0f };
        for (int i = 0; i < n; i++) {
            const float s = ((float)i + 0.5f) / (float)n - 0.5f;
            float c[3];
            tex_trilinear(t, u + mu * s, v + mv * s, width, c);
            acc[0] += c[0]; acc[1] += c[1]; acc[2] += c[2];
        }
        const float inv = 1.0f / (float)n;
        out[0] = acc[0] * inv; out[1] = acc[1] * inv; out[2] = acc[2] * inv;
    }
}

static const int TILE = 32, TILE_LOG2 = 5;

struct Oracle {
    int W = 0, H = 0, tilesX = 0, tilesY = 0, cores = 1;
    int shader = 1;                 // 0 depth-only, 1 Blinn-Phong, 2 Lambertian, 3 Lambertian * constant albedo
    int msLevel = 0, samples = 1;   // FrameBuffer.cpp:14-15
    bool hierarchical = true;
    float albedo[3] = { 0.9f, 0.9f, 0.9f };
    int texFilter = 2;              // RenderStates.h:60: TextureFilter::TriLinear
    std::vector<struct Texture> textures;        // Mesh::mTextures (Mesh.h:23); empty = the constant `albedo`
    std::vector<uint32_t> texIds;                // Mesh::GetTextureIds (Mesh.h:58): one slot index per submitted triangle
    Mat4 MV, MVinv, P, MVP, R;
    std::vector<Tile> tiles;
    std::vector<__m128> depth;      // per tile 16x16 quads, FrameBuffer.cpp:25-27
    std::vector<uint8_t> color;     // mColorBufferMS: [S][W][H] Color4b, sample fastest, rows bottom-up (FrameBuffer.cpp:19,41)
    std::vector<uint8_t> resolved;  // mColorBuffer: W*H*4 after Resolve (FrameBuffer.cpp:70-87); == color when S == 1
    std::vector<uint32_t> winner;   // ours: prim id of the last colour writer per sample, same layout as color
    std::vector<ProjVertex> projected;
    std::vector<std::vector<ProjVertex>> coreVerts;
    std::vector<std::vector<RasterTri>> coreTris;
    std::vector<Fragment> frags;
    std::vector<uint32_t> fragTile, fragSlot;
    std::vector<std::vector<uint32_t>> shaded;   // per tile, 4 x RGBA8 per fragment
    Stats stats;
};

// ------------------------------------------------------------------------------------------
// Stage a1: vertex processing (Renderer.cpp:120-127, Shader.h:40-49)
// ------------------------------------------------------------------------------------------
static void vertex_processing(Oracle& o, const float* vtx, uint32_t nv)
{
    o.projected.resize(nv);
    #pragma omp parallel for schedule(static) num_threads(o.cores)
    for (int64_t i = 0; i < (int64_t)nv; i++) {
        const float* v = vtx + 8 * i;        // InputBuffer.h:16-28: pos3, normal3, uv2
        ProjVertex& p = o.projected[i];
        p.proj = transform_point4(o.MVP, v[0], v[1], v[2], 1.0f);
        p.invW = 0.0f;
        p.position = { v[0], v[1], v[2] };
        p.normal = { v[3], v[4], v[5] };
        p.uv = { v[6], v[7] };
    }
}

// ------------------------------------------------------------------------------------------
// Stages a2-a6: clipping + setup (Clipper.h:48-288, Renderer.cpp:129-148)
// ------------------------------------------------------------------------------------------
enum { LEFT_BIT = 1, RIGHT_BIT = 2, BOTTOM_BIT = 4, TOP_BIT = 8, NEAR_BIT = 16, FAR_BIT = 32 };

static inline uint32_t clip_code(const V4f& v)      // Clipper.h:48-68
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

struct PolyVert { V4f pos; V3f wt; };
struct Poly { PolyVert v[16]; int n; };

static inline bool plane_inside(int plane, const V4f& v)
{
    switch (plane) {
    case LEFT_BIT:   return v.x >= -v.w;      // Clipper.h:240
    case RIGHT_BIT:  return v.x <= v.w;       // :247
    case BOTTOM_BIT: return v.y >= -v.w;      // :254
    case TOP_BIT:    return v.y <= v.w;       // :261
    case FAR_BIT:    return v.z <= v.w;       // :268
    default:         return v.z >= 0.0f;      // :275 NEAR
    }
}
static inline float plane_t(int plane, const V4f& a, const V4f& b)
{
    switch (plane) {
    case LEFT_BIT:   return (a.w + a.x) / ((a.x + a.w) - (b.x + b.w));      // :241
    case RIGHT_BIT:  return (-a.w + a.x) / ((a.x - a.w) - (b.x - b.w));     // :248
    case BOTTOM_BIT: return (a.w + a.y) / ((a.y + a.w) - (b.y + b.w));      // :255
    case TOP_BIT:    return (-a.w + a.y) / ((a.y - a.w) - (b.y - b.w));     // :262
    case FAR_BIT:    return (-a.w + a.z) / ((a.z - a.w) - (b.z - b.w));     // :269
    default:         return a.z / (a.z - b.z);                               // :276
    }
}
static inline void plane_snap(int plane, V4f& v)
{
    switch (plane) {
    case LEFT_BIT:   v.x = -v.w; break;
    case RIGHT_BIT:  v.x = v.w; break;
    case BOTTOM_BIT: v.y = -v.w; break;
    case TOP_BIT:    v.y = v.w; break;
    case FAR_BIT:    v.z = v.w; break;
    default:         v.z = 0.0f; break;
    }
}

static inline PolyVert cut(int plane, const PolyVert& a, const PolyVert& b)   // Clipper.h:210-214
{
    float t = plane_t(plane, a.pos, b.pos);
    float s = 1.0f - t;
    PolyVert r;
    r.pos = { a.pos.x * s + b.pos.x * t, a.pos.y * s + b.pos.y * t, a.pos.z * s + b.pos.z * t, a.pos.w * s + b.pos.w * t };
    plane_snap(plane, r.pos);
    r.wt = { a.wt.x * s + b.wt.x * t, a.wt.y * s + b.wt.y * t, a.wt.z * s + b.wt.z * t };
    return r;
}

static void clip_by_plane(int plane, const Poly& in, Poly& out)   // Clipper.h:192-232
{
    out.n = 0;
    for (int i = 0; i < in.n; i++) {
        int j = (i + 1 == in.n) ? 0 : i + 1;
        bool in0 = plane_inside(plane, in.v[i].pos), in1 = plane_inside(plane, in.v[j].pos);
        if (in0) {
            if (in1) out.v[out.n++] = in.v[j];
            else out.v[out.n++] = cut(plane, in.v[i], in.v[j]);
        } else if (in1) {
            out.v[out.n++] = cut(plane, in.v[i], in.v[j]);
            out.v[out.n++] = in.v[j];
        }
    }
}

// Clipper.h:235-288. Returns the final polygon (n == 0 if dropped).
static void clip_polygon(Poly& a, Poly& b, uint32_t planes, Poly*& result)
{
    static const int order[6] = { LEFT_BIT, RIGHT_BIT, BOTTOM_BIT, TOP_BIT, FAR_BIT, NEAR_BIT };
    Poly* cur = &a; Poly* buf = &b;
    for (int k = 0; k < 6; k++) {
        if (planes & order[k]) { clip_by_plane(order[k], *cur, *buf); std::swap(cur, buf); }
    }
    for (int i = 0; i < cur->n; i++)
        if (cur->v[i].pos.w <= 0.0f) { cur->n = 0; break; }     // :280-287
    result = cur;
}

static void clip_and_setup(Oracle& o, const uint32_t* idx, uint32_t nt)
{
    const int cores = o.cores;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(o.cores)
    for (int core = 0; core < cores; core++) {
        o.coreVerts[core].clear();
        o.coreTris[core].clear();
    }
    #pragma omp parallel for schedule(dynamic, 1) num_threads(o.cores)
    for (int core = 0; core < cores; core++) {
        // Clipper.h:80-82: contiguous ascending chunks
        uint64_t interval = ((uint64_t)nt + cores - 1) / cores;
        uint64_t begin = core * interval, end = std::min<uint64_t>((core + 1) * interval, nt);
        auto& verts = o.coreVerts[core];
        auto& tris = o.coreTris[core];
        for (uint64_t i = begin; i < end; i++) {
            const uint32_t* id = idx + 3 * i;
            const V4f v0 = o.projected[id[0]].proj, v1 = o.projected[id[1]].proj, v2 = o.projected[id[2]].proj;
            uint32_t i0 = (uint32_t)verts.size(); verts.push_back(o.projected[id[0]]);   // :96-101 (de-index)
            uint32_t i1 = (uint32_t)verts.size(); verts.push_back(o.projected[id[1]]);
            uint32_t i2 = (uint32_t)verts.size(); verts.push_back(o.projected[id[2]]);
            uint32_t c0 = clip_code(v0), c1 = clip_code(v1), c2 = clip_code(v2);
            if (c0 | c1 | c2) {
                if (!(c0 & c1 & c2)) {                                             // :109
                    Poly pa, pb; Poly* poly;
                    pa.n = 3;
                    pa.v[0] = { v0, { 1.0f, 0.0f, 0.0f } };
                    pa.v[1] = { v1, { 0.0f, 1.0f, 0.0f } };
                    pa.v[2] = { v2, { 0.0f, 0.0f, 1.0f } };
                    clip_polygon(pa, pb, (c0 ^ c1) | (c1 ^ c2) | (c2 ^ c0), poly);
                    uint32_t ids[16];
                    for (int j = 0; j < poly->n; j++) {                           // :121-153
                        const V3f wt = poly->v[j].wt;
                        if (wt.x == 1.0f) ids[j] = i0;
                        else if (wt.y == 1.0f) ids[j] = i1;
                        else if (wt.z == 1.0f) ids[j] = i2;
                        else {
                            ids[j] = (uint32_t)verts.size();
                            ProjVertex nv;
                            nv.proj = poly->v[j].pos;
                            nv.invW = 0.0f;
                            const ProjVertex a = verts[i0], b = verts[i1], c = verts[i2];
                            nv.position = wt.x * a.position + wt.y * b.position + wt.z * c.position;
                            nv.normal = wt.x * a.normal + wt.y * b.normal + wt.z * c.normal;
                            nv.uv = wt.x * a.uv + wt.y * b.uv + wt.z * c.uv;
                            verts.push_back(nv);
                        }
                    }
                    for (int k = 2; k < poly->n; k++) {                           // :156-170 fan (0,k-1,k)
                        RasterTri t;
                        if (setup_triangle(t, o.R,
                                homogeneous_project(verts[ids[0]].proj),
                                homogeneous_project(verts[ids[k - 1]].proj),
                                homogeneous_project(verts[ids[k]].proj),
                                ids[0], ids[k - 1], ids[k], core, (uint32_t)(i * 8 + (k - 2))))
                            tris.push_back(t);
                    }
                }
                continue;
            }
            RasterTri t;                                                           // :176-186
            if (setup_triangle(t, o.R, homogeneous_project(verts[i0].proj), homogeneous_project(verts[i1].proj),
                               homogeneous_project(verts[i2].proj), i0, i1, i2, core, (uint32_t)(i * 8)))
                tris.push_back(t);
        }
    }
    // Renderer.cpp:139-147: invW = 1/w; z *= invW (reciprocal-multiply, not z/w)
    #pragma omp parallel for schedule(dynamic, 1) num_threads(o.cores)
    for (int core = 0; core < cores; core++) {
        for (auto& v : o.coreVerts[core]) {
            v.invW = 1.0f / v.proj.w;
            v.proj.z *= v.invW;
        }
    }
}

// ------------------------------------------------------------------------------------------
// SSE triangle (RasterTriangle.h:258-338)
// ------------------------------------------------------------------------------------------
struct TriSSE {
    __m128i v0x, v0y, v1x, v1y, v2x, v2y;
    __m128i B0, C0, B1, C1, B2, C2;
    __m128i sB0, sC0, sB1, sC1, sB2, sC2;     // 2 * 16 * B/C : one 2x2 quad step
    __m128i tl0, tl1, tl2;
    __m128 invDet, l0, l1;

    // shim 5 / F5: BoolSSE -> IntSSE is a bit-cast, so the SSE bias is -1 where the predicate holds.
    static inline __m128i tl(__m128i ax, __m128i ay, __m128i bx, __m128i by)
    {
        return _mm_or_si128(_mm_cmpgt_epi32(by, ay), _mm_and_si128(_mm_cmpeq_epi32(ay, by), _mm_cmpgt_epi32(ax, bx)));
    }
    explicit TriSSE(const RasterTri& t)
    {
        v0x = _mm_set1_epi32(t.v0.x); v0y = _mm_set1_epi32(t.v0.y);
        v1x = _mm_set1_epi32(t.v1.x); v1y = _mm_set1_epi32(t.v1.y);
        v2x = _mm_set1_epi32(t.v2.x); v2y = _mm_set1_epi32(t.v2.y);
        B0 = _mm_set1_epi32(t.B0); C0 = _mm_set1_epi32(t.C0);
        B1 = _mm_set1_epi32(t.B1); C1 = _mm_set1_epi32(t.C1);
        B2 = _mm_set1_epi32(t.B2); C2 = _mm_set1_epi32(t.C2);
        sB0 = _mm_slli_epi32(B0, 5); sC0 = _mm_slli_epi32(C0, 5);
        sB1 = _mm_slli_epi32(B1, 5); sC1 = _mm_slli_epi32(C1, 5);
        sB2 = _mm_slli_epi32(B2, 5); sC2 = _mm_slli_epi32(C2, 5);
        tl0 = tl(v0x, v0y, v1x, v1y); tl1 = tl(v1x, v1y, v2x, v2y); tl2 = tl(v2x, v2y, v0x, v0y);
        invDet = _mm_set1_ps(t.invDet);
        l0 = l1 = _mm_setzero_ps();
    }
    inline __m128i e0(__m128i px, __m128i py) const { return _mm_add_epi32(_mm_add_epi32(_mm_mullo_epi32(B0, _mm_sub_epi32(px, v0x)), _mm_mullo_epi32(C0, _mm_sub_epi32(py, v0y))), tl0); }
    inline __m128i e1(__m128i px, __m128i py) const { return _mm_add_epi32(_mm_add_epi32(_mm_mullo_epi32(B1, _mm_sub_epi32(px, v1x)), _mm_mullo_epi32(C1, _mm_sub_epi32(py, v1y))), tl1); }
    inline __m128i e2(__m128i px, __m128i py) const { return _mm_add_epi32(_mm_add_epi32(_mm_mullo_epi32(B2, _mm_sub_epi32(px, v2x)), _mm_mullo_epi32(C2, _mm_sub_epi32(py, v2y))), tl2); }
    // RasterTriangle.h:333-337
    inline void bary(__m128i px, __m128i py)
    {
        __m128i dx = _mm_sub_epi32(px, v2x), dy = _mm_sub_epi32(py, v2y);
        l0 = _mm_mul_ps(_mm_cvtepi32_ps(_mm_add_epi32(_mm_mullo_epi32(B1, dx), _mm_mullo_epi32(C1, dy))), invDet);
        l1 = _mm_mul_ps(_mm_cvtepi32_ps(_mm_add_epi32(_mm_mullo_epi32(B2, dx), _mm_mullo_epi32(C2, dy))), invDet);
    }
    // RasterTriangle.h:324-331
    inline __m128 depth(float z0, float z1, float z2) const
    {
        __m128 l2 = _mm_sub_ps(_mm_sub_ps(_mm_set1_ps(1.0f), l0), l1);
        return _mm_add_ps(_mm_add_ps(_mm_mul_ps(l0, _mm_set1_ps(z0)), _mm_mul_ps(l1, _mm_set1_ps(z1))), _mm_mul_ps(l2, _mm_set1_ps(z2)));
    }
};

// FrameBuffer.cpp:54-68. LESS_EQUAL, masked write; returns the UNMASKED compare.
static inline __m128 ztest_quad(Oracle& o, __m128 d, int x, int y, int sId, __m128 mask)
{
    int tx = x >> TILE_LOG2, ty = y >> TILE_LOG2;
    int ix = x & (TILE - 1), iy = y & (TILE - 1);
    __m128& cur = o.depth[((size_t)(ty * o.tilesX + tx) * o.samples + sId) * 256 + (size_t)(iy >> 1) * 16 + (ix >> 1)];
    __m128 ret = _mm_cmple_ps(d, cur);
    cur = _mm_blendv_ps(cur, d, _mm_and_ps(ret, mask));
    return ret;
}

// FrameBuffer.cpp:107-191. Offsets in 1/16 pixel relative to the pixel centre, (x, y) pairs.
// shim 17: the reference binds `const Vector2i&` to ONE element, [2*sampleId], of this int table
// (Rasterizer.h:247,382). That only compiles through a converting constructor Vector2i(int), which sets
// both components to the value, so sample s sits at (table[2s], table[2s]) — the table's second column is
// never read. Established by compiling the reference itself (oracle/_ref); round 1 read the written pairs.
static const int kSampleOffsets[6][64] = {
    { 0, 0 },
    { 4, 4, -4, -4 },
    { -2, -6, 6, -2, -6, 2, 2, 6 },
    { 1, -3, -1, 3, 5, 1, -3, -5, -5, 5, -7, -1, 3, 7, 7, -7 },
    { 1, 1, -1, -3, -3, 2, 4, -1, -5, -2, 2, 5, 5, 3, 3, -5, -2, 6, 0, -7, -4, -6, -6, 4, -8, 0, 7, -4, 6, 7, -7, -8 },
    { 1, 1, -1, -3, -3, 2, 4, -1, -5, -2, 2, 5, 5, 3, 3, -5, -2, 6, 0, -7, -4, -6, -6, 4, -8, 0, 7, -4, 6, 7, -7, -8,
      1, 3, -3, -3, -3, 0, 6, -2, -7, -1, 3, 4, 7, 3, 3, -6, -2, 7, 0, -4, -2, -5, -7, 6, -8, 3, 4, -1, 2, 7, 4, -8 },
};

static inline int min3(int a, int b, int c) { return std::min(a, std::min(b, c)); }
static inline int max3(int a, int b, int c) { return std::max(a, std::max(b, c)); }

static inline void emit(Tile& tile, const TriSSE& s, const RasterTri& t, int x, int y, const uint32_t mask[4])
{
    Fragment f;
    f.l0 = s.l0; f.l1 = s.l1;
    f.vId0 = t.vId0; f.vId1 = t.vId1; f.vId2 = t.vId2; f.coreId = t.coreId; f.primId = t.primId;
    f.x = (uint16_t)x; f.y = (uint16_t)y;
    f.mask[0] = mask[0]; f.mask[1] = mask[1]; f.mask[2] = mask[2]; f.mask[3] = mask[3];
    tile.frags.push_back(f);
}
static inline void mask_set(uint32_t mask[4], int lanes4, int sampleId)      // CoverageMask::SetBit, Shader.h:73-92
{
    int bit = sampleId << 2;
    mask[bit >> 5] |= (uint32_t)lanes4 << (bit & 31);
}

// Rasterizer.h:126-200
static void fine_rasterize(Oracle& o, Tile& tile, V2i bmin, V2i bmax, const RasterTri& t, uint64_t& covered)
{
    int minX = std::max(bmin.x, min3(t.v0.x, t.v1.x, t.v2.x) >> 4);
    int maxX = std::min(bmax.x - 1, max3(t.v0.x, t.v1.x, t.v2.x) >> 4);
    int minY = std::max(bmin.y, min3(t.v0.y, t.v1.y, t.v2.y) >> 4);
    int maxY = std::min(bmax.y - 1, max3(t.v0.y, t.v1.y, t.v2.y) >> 4);
    minX -= minX % 2;
    minY -= minY % 2;
    if (maxX < minX || maxY < minY) return;

    TriSSE s(t);
    const ProjVertex* vb = o.coreVerts[t.coreId].data();
    const float z0 = vb[t.vId0].proj.z, z1 = vb[t.vId1].proj.z, z2 = vb[t.vId2].proj.z;
    const __m128i offX = _mm_setr_epi32(8, 24, 8, 24), offY = _mm_setr_epi32(8, 8, 24, 24);   // Rasterizer.h:23
    __m128i cy = _mm_add_epi32(_mm_set1_epi32(minY << 4), offY);
    __m128i cx0 = _mm_add_epi32(_mm_set1_epi32(minX << 4), offX);
    __m128i e0 = s.e0(cx0, cy), e1 = s.e1(cx0, cy), e2 = s.e2(cx0, cy);
    const __m128i step32 = _mm_set1_epi32(32);
    for (int y = minY; y <= maxY; y += 2) {
        __m128i r0 = e0, r1 = e1, r2 = e2;
        __m128i cx = cx0;
        for (int x = minX; x <= maxX; x += 2) {
            // covered = (e0|e1|e2) >= 0  <=>  sign bit of the OR is clear
            int cov = (~_mm_movemask_ps(_mm_castsi128_ps(_mm_or_si128(_mm_or_si128(e0, e1), e2)))) & 15;
            if (cov) {
                covered += (uint64_t)__builtin_popcount(cov);
                s.bary(cx, cy);
                __m128 covMask = _mm_castsi128_ps(_mm_cmpgt_epi32(_mm_set1_epi32(0), _mm_or_si128(_mm_or_si128(e0, e1), e2)));
                covMask = _mm_xor_ps(covMask, _mm_castsi128_ps(_mm_set1_epi32(-1)));
                __m128 zt = ztest_quad(o, s.depth(z0, z1, z2), x, y, 0, covMask);
                int vis = _mm_movemask_ps(zt) & cov;
                if (vis) { uint32_t m[4] = { (uint32_t)vis, 0, 0, 0 }; emit(tile, s, t, x, y, m); }
            }
            e0 = _mm_add_epi32(e0, s.sB0); e1 = _mm_add_epi32(e1, s.sB1); e2 = _mm_add_epi32(e2, s.sB2);
            cx = _mm_add_epi32(cx, step32);
        }
        e0 = _mm_add_epi32(r0, s.sC0); e1 = _mm_add_epi32(r1, s.sC1); e2 = _mm_add_epi32(r2, s.sC2);
        cy = _mm_add_epi32(cy, step32);
    }
}

// Rasterizer.h:310-353
static void trivial_accept(Oracle& o, Tile& tile, V2i bmin, V2i bmax, const RasterTri& t, uint64_t& covered)
{
    int minX = bmin.x, maxX = bmax.x - 1, minY = bmin.y, maxY = bmax.y - 1;
    minX -= minX % 2;
    minY -= minY % 2;
    TriSSE s(t);
    const ProjVertex* vb = o.coreVerts[t.coreId].data();
    const float z0 = vb[t.vId0].proj.z, z1 = vb[t.vId1].proj.z, z2 = vb[t.vId2].proj.z;
    const __m128i offX = _mm_setr_epi32(8, 24, 8, 24), offY = _mm_setr_epi32(8, 8, 24, 24);
    const __m128 all = _mm_castsi128_ps(_mm_set1_epi32(-1));
    for (int y = minY; y <= maxY; y += 2) {
        __m128i cy = _mm_add_epi32(_mm_set1_epi32(y << 4), offY);
        for (int x = minX; x <= maxX; x += 2) {
            __m128i cx = _mm_add_epi32(_mm_set1_epi32(x << 4), offX);
            s.bary(cx, cy);
            covered += 4;
            int vis = _mm_movemask_ps(ztest_quad(o, s.depth(z0, z1, z2), x, y, 0, all));
            if (vis) { uint32_t m[4] = { (uint32_t)vis, 0, 0, 0 }; emit(tile, s, t, x, y, m); }
        }
    }
}

// Rasterizer.h:202-300
static void fine_rasterize_ms(Oracle& o, Tile& tile, V2i bmin, V2i bmax, const RasterTri& t, uint64_t& covered)
{
    int minX = std::max(bmin.x, min3(t.v0.x, t.v1.x, t.v2.x) >> 4);
    int maxX = std::min(bmax.x - 1, max3(t.v0.x, t.v1.x, t.v2.x) >> 4);
    int minY = std::max(bmin.y, min3(t.v0.y, t.v1.y, t.v2.y) >> 4);
    int maxY = std::min(bmax.y - 1, max3(t.v0.y, t.v1.y, t.v2.y) >> 4);
    minX -= minX % 2;
    minY -= minY % 2;
    if (maxX < minX || maxY < minY) return;
    TriSSE s(t);
    const ProjVertex* vb = o.coreVerts[t.coreId].data();
    const float z0 = vb[t.vId0].proj.z, z1 = vb[t.vId1].proj.z, z2 = vb[t.vId2].proj.z;
    const __m128i offX = _mm_setr_epi32(8, 24, 8, 24), offY = _mm_setr_epi32(8, 8, 24, 24);
    const int* table = kSampleOffsets[o.msLevel];
    __m128i cy = _mm_add_epi32(_mm_set1_epi32(minY << 4), offY);
    __m128i cx0 = _mm_add_epi32(_mm_set1_epi32(minX << 4), offX);
    __m128i e0 = s.e0(cx0, cy), e1 = s.e1(cx0, cy), e2 = s.e2(cx0, cy);
    const __m128i step32 = _mm_set1_epi32(32);
    for (int y = minY; y <= maxY; y += 2) {
        __m128i r0 = e0, r1 = e1, r2 = e2;
        __m128i cx = cx0;
        for (int x = minX; x <= maxX; x += 2) {
            uint32_t mask[4] = { 0, 0, 0, 0 };
            bool gen = false;
            for (int sId = 0; sId < o.samples; sId++) {
                const int ox = table[2 * sId], oy = table[2 * sId];      // shim 17
                const __m128i vx = _mm_set1_epi32(ox), vy = _mm_set1_epi32(oy);
                // e = edgeVal + off.x * B + off.y * C   (Rasterizer.h:248-250)
                __m128i f0 = _mm_add_epi32(_mm_add_epi32(e0, _mm_mullo_epi32(vx, s.B0)), _mm_mullo_epi32(vy, s.C0));
                __m128i f1 = _mm_add_epi32(_mm_add_epi32(e1, _mm_mullo_epi32(vx, s.B1)), _mm_mullo_epi32(vy, s.C1));
                __m128i f2 = _mm_add_epi32(_mm_add_epi32(e2, _mm_mullo_epi32(vx, s.B2)), _mm_mullo_epi32(vy, s.C2));
                __m128i orv = _mm_or_si128(_mm_or_si128(f0, f1), f2);
                int cov = (~_mm_movemask_ps(_mm_castsi128_ps(orv))) & 15;
                if (cov) {
                    covered += (uint64_t)__builtin_popcount(cov);
                    s.bary(_mm_add_epi32(cx, vx), _mm_add_epi32(cy, vy));
                    __m128 covMask = _mm_xor_ps(_mm_castsi128_ps(_mm_cmpgt_epi32(_mm_set1_epi32(0), orv)), _mm_castsi128_ps(_mm_set1_epi32(-1)));
                    int vis = _mm_movemask_ps(ztest_quad(o, s.depth(z0, z1, z2), x, y, sId, covMask)) & cov;
                    if (vis) { mask_set(mask, vis, sId); gen = true; }
                }
            }
            if (gen) {
                s.bary(cx, cy);                  // shading barycentrics at the pixel centres (Rasterizer.h:275)
                emit(tile, s, t, x, y, mask);
            }
            e0 = _mm_add_epi32(e0, s.sB0); e1 = _mm_add_epi32(e1, s.sB1); e2 = _mm_add_epi32(e2, s.sB2);
            cx = _mm_add_epi32(cx, step32);
        }
        e0 = _mm_add_epi32(r0, s.sC0); e1 = _mm_add_epi32(r1, s.sC1); e2 = _mm_add_epi32(r2, s.sC2);
        cy = _mm_add_epi32(cy, step32);
    }
}

// Rasterizer.h:355-415
static void trivial_accept_ms(Oracle& o, Tile& tile, V2i bmin, V2i bmax, const RasterTri& t, uint64_t& covered)
{
    int minX = bmin.x, maxX = bmax.x - 1, minY = bmin.y, maxY = bmax.y - 1;
    minX -= minX % 2;
    minY -= minY % 2;
    TriSSE s(t);
    const ProjVertex* vb = o.coreVerts[t.coreId].data();
    const float z0 = vb[t.vId0].proj.z, z1 = vb[t.vId1].proj.z, z2 = vb[t.vId2].proj.z;
    const __m128i offX = _mm_setr_epi32(8, 24, 8, 24), offY = _mm_setr_epi32(8, 8, 24, 24);
    const __m128 all = _mm_castsi128_ps(_mm_set1_epi32(-1));
    const int* table = kSampleOffsets[o.msLevel];
    for (int y = minY; y <= maxY; y += 2) {
        __m128i cy = _mm_add_epi32(_mm_set1_epi32(y << 4), offY);
        for (int x = minX; x <= maxX; x += 2) {
            __m128i cx = _mm_add_epi32(_mm_set1_epi32(x << 4), offX);
            uint32_t mask[4] = { 0, 0, 0, 0 };
            bool gen = false;
            for (int sId = 0; sId < o.samples; sId++) {
                s.bary(_mm_add_epi32(cx, _mm_set1_epi32(table[2 * sId])), _mm_add_epi32(cy, _mm_set1_epi32(table[2 * sId])));      // shim 17
                covered += 4;
                int vis = _mm_movemask_ps(ztest_quad(o, s.depth(z0, z1, z2), x, y, sId, all));
                if (vis) { mask_set(mask, vis, sId); gen = true; }
            }
            if (gen) {
                s.bary(cx, cy);
                emit(tile, s, t, x, y, mask);
            }
        }
    }
}

// Rasterizer.h:113-124, 302-308: single- vs multi-sample dispatch
static inline void fine_dispatch(Oracle& o, Tile& tile, V2i bmin, V2i bmax, const RasterTri& t, uint64_t& covered)
{
    if (o.samples == 1) fine_rasterize(o, tile, bmin, bmax, t, covered);
    else fine_rasterize_ms(o, tile, bmin, bmax, t, covered);
}
static inline void accept_dispatch(Oracle& o, Tile& tile, V2i bmin, V2i bmax, const RasterTri& t, uint64_t& covered)
{
    if (o.samples == 1) trivial_accept(o, tile, bmin, bmax, t, covered);
    else trivial_accept_ms(o, tile, bmin, bmax, t, covered);
}

// Rasterizer.h:32-111 with RasterTriangle.h:192-255 (step vectors)
static void coarse_rasterize(Oracle& o, Tile& tile, const TriRef& ref, const RasterTri& t, uint64_t& covered)
{
    const V2i bmin = tile.minC, bmax = tile.maxC;
    const int baseX = bmin.x << 4, baseY = bmin.y << 4;
    const int far = (TILE << 4) - 1;                 // 511
    const int step = TILE << 3;                      // 256 sub-pixels = one 16-px quadrant
    auto corner_vals = [&](int corner, int B, int C, int which, int out[4]) {
        int cx = corner & 1, cy = corner >> 1;
        int px = baseX + cx * far, py = baseY + cy * far;
        int e = which == 0 ? t.edge0(px, py) : (which == 1 ? t.edge1(px, py) : t.edge2(px, py));
        for (int q = 0; q < 4; q++) {
            int qx = q & 1, qy = q >> 1;
            out[q] = (int)((uint32_t)e + (uint32_t)step * ((uint32_t)((qx - cx) * B) + (uint32_t)((qy - cy) * C)));
        }
    };
    int rej0[4], rej1[4], rej2[4], acc0[4], acc1[4], acc2[4];
    corner_vals(t.rej0, t.B0, t.C0, 0, rej0);
    corner_vals(t.rej1, t.B1, t.C1, 1, rej1);
    corner_vals(t.rej2, t.B2, t.C2, 2, rej2);
    for (int q = 0; q < 4; q++) acc0[q] = acc1[q] = acc2[q] = INT_MAX;   // shim 12
    if (!ref.acc0) corner_vals(t.acc0, t.B0, t.C0, 0, acc0);
    if (!ref.acc1) corner_vals(t.acc1, t.B1, t.C1, 1, acc1);
    if (!ref.acc2) corner_vals(t.acc2, t.B2, t.C2, 2, acc2);
    for (int q = 0; q < 4; q++) {
        if (rej0[q] < 0 || rej1[q] < 0 || rej2[q] < 0) continue;
        const int half = TILE >> 1;
        V2i qmin, qmax;
        qmin.x = !(q & 1) ? bmin.x : bmin.x + half;
        qmax.x = !(q & 1) ? bmin.x + half : bmax.x;
        qmin.y = !(q >> 1) ? bmin.y : bmin.y + half;
        qmax.y = !(q >> 1) ? bmin.y + half : bmax.y;
        // The reference does not clamp a quadrant to a partial tile shorter than 16 px
        // (Rasterizer.h:92-95, a latent out-of-bounds write). We clamp; none of the five
        // benchmark configs reaches this case.
        qmax.x = std::min(qmax.x, bmax.x);
        qmax.y = std::min(qmax.y, bmax.y);
        if (acc0[q] >= 0 && acc1[q] >= 0 && acc2[q] >= 0) accept_dispatch(o, tile, qmin, qmax, t, covered);
        else fine_dispatch(o, tile, qmin, qmax, t, covered);
    }
}

// ------------------------------------------------------------------------------------------
// Stages a7-a8: binning + per-tile rasterization (Renderer.cpp:150-270)
// ------------------------------------------------------------------------------------------
static void tiled_rasterization(Oracle& o, double& msBin, double& msRaster)
{
    auto t0 = std::chrono::steady_clock::now();
    const int nTiles = (int)o.tiles.size();
    #pragma omp parallel for schedule(dynamic, 16) num_threads(o.cores)
    for (int i = 0; i < nTiles; i++) {
        for (auto& r : o.tiles[i].refs) r.clear();
        o.tiles[i].frags.clear();
    }
    const int Shift = TILE_LOG2 + 4;
    uint64_t nRefs = 0;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(o.cores) reduction(+:nRefs)
    for (int core = 0; core < o.cores; core++) {
        const auto& tris = o.coreTris[core];
        for (uint32_t i = 0; i < (uint32_t)tris.size(); i++) {
            const RasterTri& t = tris[i];
            int minX = std::max(0, min3(t.v0.x, t.v1.x, t.v2.x) >> Shift);
            int maxX = std::min(o.tilesX - 1, max3(t.v0.x, t.v1.x, t.v2.x) >> Shift);
            int minY = std::max(0, min3(t.v0.y, t.v1.y, t.v2.y) >> Shift);
            int maxY = std::min(o.tilesY - 1, max3(t.v0.y, t.v1.y, t.v2.y) >> Shift);
            if (maxX - minX < 2 && maxY - minY < 2) {                      // Renderer.cpp:173-180
                for (int y = minY; y <= maxY; y++)
                    for (int x = minX; x <= maxX; x++) {
                        o.tiles[y * o.tilesX + x].refs[core].push_back({ i, false, false, false, false, false });
                        nRefs++;
                    }
            } else {                                                        // :181-226
                for (int y = minY; y <= maxY; y++)
                    for (int x = minX; x <= maxX; x++) {
                        auto cx = [&](int c) { return (x + (c & 1)) << Shift; };
                        auto cy = [&](int c) { return (y + (c >> 1)) << Shift; };
                        if (t.edge0(cx(t.rej0), cy(t.rej0)) < 0 || t.edge1(cx(t.rej1), cy(t.rej1)) < 0 ||
                            t.edge2(cx(t.rej2), cy(t.rej2)) < 0)
                            continue;
                        bool a0 = t.edge0(cx(t.acc0), cy(t.acc0)) >= 0;
                        bool a1 = t.edge1(cx(t.acc1), cy(t.acc1)) >= 0;
                        bool a2 = t.edge2(cx(t.acc2), cy(t.acc2)) >= 0;
                        o.tiles[y * o.tilesX + x].refs[core].push_back({ i, a0, a1, a2, a0 && a1 && a2, true });
                        nRefs++;
                    }
            }
        }
    }
    o.stats.nBinRefs = nRefs;
    auto t1 = std::chrono::steady_clock::now();

    uint64_t covered = 0;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(o.cores) reduction(+:covered)
    for (int i = 0; i < nTiles; i++) {
        Tile& tile = o.tiles[i];
        for (int core = 0; core < o.cores; core++) {                        // Renderer.cpp:248-270
            for (const TriRef& ref : tile.refs[core]) {
                const RasterTri& t = o.coreTris[core][ref.triId];
                if (ref.trivialAccept) { accept_dispatch(o, tile, tile.minC, tile.maxC, t, covered); continue; }
                if (o.hierarchical && ref.big) coarse_rasterize(o, tile, ref, t, covered);
                else fine_dispatch(o, tile, tile.minC, tile.maxC, t, covered);
            }
        }
    }
    o.stats.nCoveredSamples = covered;
    // Renderer.cpp:238-245: serial concatenation of every tile's fragments
    o.frags.clear(); o.fragTile.clear(); o.fragSlot.clear();
    for (int i = 0; i < nTiles; i++) {
        o.shaded[i].resize(o.tiles[i].frags.size() * 4);
        for (uint32_t j = 0; j < (uint32_t)o.tiles[i].frags.size(); j++) {
            o.frags.push_back(o.tiles[i].frags[j]);
            o.fragTile.push_back((uint32_t)i);
            o.fragSlot.push_back(j);
        }
    }
    o.stats.nFragments = o.frags.size();
    auto t2 = std::chrono::steady_clock::now();
    msBin = std::chrono::duration<double, std::milli>(t1 - t0).count();
    msRaster = std::chrono::duration<double, std::milli>(t2 - t1).count();
}

// ------------------------------------------------------------------------------------------
// Stages a15-a17: interpolate, shade, pack, frame-buffer update (Renderer.cpp:272-350)
// ------------------------------------------------------------------------------------------
static inline __m128 dot3(__m128 ax, __m128 ay, __m128 az, __m128 bx, __m128 by, __m128 bz)
{
    return _mm_add_ps(_mm_add_ps(_mm_mul_ps(ax, bx), _mm_mul_ps(ay, by)), _mm_mul_ps(az, bz));
}
static inline __m128 lerp3(__m128 b0, __m128 b1, __m128 b2, float a0, float a1, float a2)
{
    return _mm_add_ps(_mm_add_ps(_mm_mul_ps(b0, _mm_set1_ps(a0)), _mm_mul_ps(b1, _mm_set1_ps(a1))), _mm_mul_ps(b2, _mm_set1_ps(a2)));
}

static void fragment_processing(Oracle& o)
{
    if (o.shader == 0) return;
    const V3f eye = transform_point3(o.MVinv, { 0.0f, 0.0f, 0.0f });       // Renderer.cpp:289
    const V3f L = normalize3({ 1.0f, 1.0f, -1.0f });                        // Renderer.cpp:290, Shader.h:258
    const int64_t n = (int64_t)o.frags.size();
    const __m128 one = _mm_set1_ps(1.0f);
    #pragma omp parallel for schedule(dynamic, 256) num_threads(o.cores)
    for (int64_t i = 0; i < n; i++) {
        const Fragment& f = o.frags[i];
        const ProjVertex* vb = o.coreVerts[f.coreId].data();
        const ProjVertex& v0 = vb[f.vId0]; const ProjVertex& v1 = vb[f.vId1]; const ProjVertex& v2 = vb[f.vId2];
        // Shader.h:142-170 perspective-correct interpolation
        __m128 b0 = f.l0, b1 = f.l1;
        __m128 b2 = _mm_sub_ps(_mm_sub_ps(one, b0), b1);
        b0 = _mm_mul_ps(b0, _mm_set1_ps(v0.invW));
        b1 = _mm_mul_ps(b1, _mm_set1_ps(v1.invW));
        b2 = _mm_mul_ps(b2, _mm_set1_ps(v2.invW));
        __m128 invB = _mm_div_ps(one, _mm_add_ps(_mm_add_ps(b0, b1), b2));
        b0 = _mm_mul_ps(b0, invB);
        b1 = _mm_mul_ps(b1, invB);
        b2 = _mm_sub_ps(_mm_sub_ps(one, b0), b1);
        __m128 px = lerp3(b0, b1, b2, v0.position.x, v1.position.x, v2.position.x);
        __m128 py = lerp3(b0, b1, b2, v0.position.y, v1.position.y, v2.position.y);
        __m128 pz = lerp3(b0, b1, b2, v0.position.z, v1.position.z, v2.position.z);
        __m128 nx = lerp3(b0, b1, b2, v0.normal.x, v1.normal.x, v2.normal.x);
        __m128 ny = lerp3(b0, b1, b2, v0.normal.y, v1.normal.y, v2.normal.y);
        __m128 nz = lerp3(b0, b1, b2, v0.normal.z, v1.normal.z, v2.normal.z);
        const __m128 tu = lerp3(b0, b1, b2, v0.uv.x, v1.uv.x, v2.uv.x);      // Shader.h:167-169
        const __m128 tv = lerp3(b0, b1, b2, v0.uv.y, v1.uv.y, v2.uv.y);

        // Shader.h:256-264 (shared by the Lambertian and Blinn-Phong shaders)
        __m128 w = rsqrt4(dot3(nx, ny, nz, nx, ny, nz));
        nx = _mm_mul_ps(nx, w); ny = _mm_mul_ps(ny, w); nz = _mm_mul_ps(nz, w);
        const __m128 Lx = _mm_set1_ps(L.x), Ly = _mm_set1_ps(L.y), Lz = _mm_set1_ps(L.z);
        __m128 d = dot3(Lx, Ly, Lz, nx, ny, nz);
        d = _mm_blendv_ps(d, _mm_setzero_ps(), _mm_cmplt_ps(d, _mm_setzero_ps()));
        __m128 diffuse = _mm_mul_ps(_mm_mul_ps(_mm_add_ps(d, _mm_set1_ps(0.2f)), _mm_set1_ps(3.0f)), _mm_set1_ps(kInvPi));
        __m128 r = diffuse, g = diffuse, b = diffuse;
        if (o.shader == 1) {
            // Shader.h:266-280
            __m128 ex = _mm_sub_ps(_mm_set1_ps(eye.x), px), ey = _mm_sub_ps(_mm_set1_ps(eye.y), py), ez = _mm_sub_ps(_mm_set1_ps(eye.z), pz);
            w = rsqrt4(dot3(ex, ey, ez, ex, ey, ez));
            ex = _mm_mul_ps(ex, w); ey = _mm_mul_ps(ey, w); ez = _mm_mul_ps(ez, w);
            __m128 hx = _mm_add_ps(Lx, ex), hy = _mm_add_ps(Ly, ey), hz = _mm_add_ps(Lz, ez);
            w = rsqrt4(dot3(hx, hy, hz, hx, hy, hz));
            hx = _mm_mul_ps(hx, w); hy = _mm_mul_ps(hy, w); hz = _mm_mul_ps(hz, w);
            alignas(16) float sp[4];
            _mm_store_ps(sp, dot3(nx, ny, nz, hx, hy, hz));
            for (int k = 0; k < 4; k++) sp[k] = powf(sp[k], 200.0f);      // shim 10: Math::Pow = powf
            __m128 spec = _mm_mul_ps(_mm_load_ps(sp), _mm_set1_ps(3.0f));
            r = g = b = _mm_add_ps(diffuse, spec);
        } else if (o.shader == 3 && o.textures.empty()) {
            // Shader.h:209-244 with the constant-colour texture Mesh.cpp:48,67 installs
            r = _mm_mul_ps(diffuse, _mm_set1_ps(o.albedo[0]));
            g = _mm_mul_ps(diffuse, _mm_set1_ps(o.albedo[1]));
            b = _mm_mul_ps(diffuse, _mm_set1_ps(o.albedo[2]));
        } else if (o.shader == 3) {
            // Shader.h:228-241: differentials from quad lanes 1 and 2 against lane 0, one Sample per lane
            alignas(16) float uu[4], vv[4], ar[4], ag[4], ab[4];
            _mm_store_ps(uu, tu); _mm_store_ps(vv, tv);
            const uint32_t tri = f.primId >> 3;
            const uint32_t slot = tri < o.texIds.size() ? o.texIds[tri] : 0u;
            const Texture& tex = o.textures[slot < o.textures.size() ? slot : 0u];
            const float du0 = uu[1] - uu[0], dv0 = vv[1] - vv[0], du1 = uu[2] - uu[0], dv1 = vv[2] - vv[0];
            for (int k = 0; k < 4; k++) {
                float c[3];
                tex_sample(tex, o.texFilter, uu[k], vv[k], du0, dv0, du1, dv1, c);
                ar[k] = c[0]; ag[k] = c[1]; ab[k] = c[2];
            }
            r = _mm_mul_ps(diffuse, _mm_load_ps(ar));
            g = _mm_mul_ps(diffuse, _mm_load_ps(ag));
            b = _mm_mul_ps(diffuse, _mm_load_ps(ab));
        }
        alignas(16) float rr[4], gg[4], bb[4];
        _mm_store_ps(rr, r); _mm_store_ps(gg, g); _mm_store_ps(bb, b);
        uint32_t* out = &o.shaded[o.fragTile[i]][(size_t)o.fragSlot[i] * 4];
        for (int k = 0; k < 4; k++)                                         // Renderer.cpp:295-301
            out[k] = (uint32_t)to_u8(rr[k]) | ((uint32_t)to_u8(gg[k]) << 8) | ((uint32_t)to_u8(bb[k]) << 16) | 0xFF000000u;
    }
}

static void update_frame_buffer(Oracle& o)
{
    const int nTiles = (int)o.tiles.size();
    const int S = o.samples;
    #pragma omp parallel for schedule(dynamic, 4) num_threads(o.cores)
    for (int i = 0; i < nTiles; i++) {
        const Tile& tile = o.tiles[i];
        for (size_t j = 0; j < tile.frags.size(); j++) {                    // Renderer.cpp:309-345
            const Fragment& f = tile.frags[j];
            for (int sId = 0; sId < S; sId++) {
                const int shift = sId << 2;
                const uint32_t lanes = (f.mask[shift >> 5] >> (shift & 31)) & 15u;
                for (int k = 0; k < 4; k++) {
                    if (!(lanes & (1u << k))) continue;
                    int x = f.x + (k & 1), y = f.y + (k >> 1);
                    if (x >= o.W || y >= o.H) continue;                      // odd sizes only; see DESIGN.md
                    size_t at = ((size_t)x + (size_t)o.W * (size_t)(o.H - 1 - y)) * S + sId;   // FrameBuffer.cpp:41
                    if (o.shader != 0) memcpy(&o.color[at * 4], &o.shaded[i][j * 4 + k], 4);
                    o.winner[at] = f.primId;
                }
            }
        }
    }
    // FrameBuffer::Resolve, FrameBuffer.cpp:70-87. shim 18: Color(Color4b) = bytes / 255, Color4b(Color) = the
    // FromFloats rounding of shim 13 on every channel (alpha included).
    if (S == 1) { o.resolved = o.color; return; }
    const float inv = 1.0f / (float)S;
    const int64_t nPix = (int64_t)o.W * o.H;
    #pragma omp parallel for schedule(static) num_threads(o.cores)
    for (int64_t p = 0; p < nPix; p++) {
        float acc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
        for (int sId = 0; sId < S; sId++)
            for (int c = 0; c < 4; c++) acc[c] = acc[c] + (float)o.color[((size_t)p * S + sId) * 4 + c] * (1.0f / 255.0f);
        for (int c = 0; c < 4; c++) o.resolved[(size_t)p * 4 + c] = to_u8(acc[c] * inv);
    }
}

static double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

static void render(Oracle& o, const float* vtx, uint32_t nv, const uint32_t* idx, uint32_t nt)
{
    memset(&o.stats, 0, sizeof(o.stats));
    // FrameBuffer.cpp:89-105: colour := 0, depth := 1.0
    auto t0 = std::chrono::steady_clock::now();
    std::fill(o.color.begin(), o.color.end(), (uint8_t)0);
    std::fill(o.winner.begin(), o.winner.end(), 0xFFFFFFFFu);
    {
        const __m128 one = _mm_set1_ps(1.0f);
        const int64_t n = (int64_t)o.depth.size();
        #pragma omp parallel for schedule(static) num_threads(o.cores)
        for (int64_t i = 0; i < n; i++) o.depth[i] = one;
    }
    vertex_processing(o, vtx, nv);
    o.stats.ms[0] = ms_since(t0); t0 = std::chrono::steady_clock::now();
    clip_and_setup(o, idx, nt);
    o.stats.ms[1] = ms_since(t0);
    for (auto& v : o.coreTris) o.stats.nRasterTris += v.size();
    tiled_rasterization(o, o.stats.ms[2], o.stats.ms[3]);
    t0 = std::chrono::steady_clock::now();
    fragment_processing(o);
    o.stats.ms[4] = ms_since(t0); t0 = std::chrono::steady_clock::now();
    update_frame_buffer(o);
    o.stats.ms[5] = ms_since(t0);
}

static void resize(Oracle& o, int w, int h)
{
    o.W = w; o.H = h;
    o.tilesX = (w + TILE - 1) >> TILE_LOG2;       // Renderer.cpp:26-27
    o.tilesY = (h + TILE - 1) >> TILE_LOG2;
    o.tiles.clear();
    for (int y = 0; y < h; y += TILE)
        for (int x = 0; x < w; x += TILE) {        // Renderer.cpp:43-53
            Tile t;
            t.minC = { x, y };
            t.maxC = { std::min(x + TILE, w), std::min(y + TILE, h) };
            t.refs.resize(o.cores);
            o.tiles.push_back(std::move(t));
        }
    o.depth.assign((size_t)o.tilesX * o.tilesY * 256 * o.samples, _mm_set1_ps(1.0f));
    o.color.assign((size_t)w * h * 4 * o.samples, 0);
    o.resolved.assign((size_t)w * h * 4, 0);
    o.winner.assign((size_t)w * h * o.samples, 0xFFFFFFFFu);
    o.shaded.assign(o.tiles.size(), {});
    o.coreVerts.assign(o.cores, {});
    o.coreTris.assign(o.cores, {});
}

} // namespace orc

// ------------------------------------------------------------------------------------------
// C entry points (ctypes-friendly). Test infrastructure only.
// ------------------------------------------------------------------------------------------
using namespace orc;

extern "C" {

void* orc_create(int w, int h, int threads)
{
    Oracle* o = new Oracle;
    o->cores = threads > 0 ? threads : omp_get_max_threads();
    Mat4 I; memset(&I, 0, sizeof(I)); for (int i = 0; i < 4; i++) I.m[i][i] = 1.0f;
    o->MV = o->MVinv = o->P = o->MVP = o->R = I;
    resize(*o, w, h);
    return o;
}
void orc_destroy(void* h) { delete (Oracle*)h; }
int orc_threads(void* h) { return ((Oracle*)h)->cores; }
int orc_max_threads() { return omp_get_max_threads(); }
void orc_resize(void* h, int w, int ht) { resize(*(Oracle*)h, w, ht); }

// Renderer.cpp:85-92
void orc_set_transform(void* h, const float* mv, const float* proj, const float* raster)
{
    Oracle& o = *(Oracle*)h;
    memcpy(o.MV.m, mv, 64); memcpy(o.P.m, proj, 64); memcpy(o.R.m, raster, 64);
    o.MVinv = mat_inverse(o.MV);
    o.MVP = mat_mul(o.P, o.MV);
}
void orc_get_derived(void* h, float* mvp16, float* eye3, float* light3)
{
    Oracle& o = *(Oracle*)h;
    memcpy(mvp16, o.MVP.m, 64);
    V3f e = transform_point3(o.MVinv, { 0.0f, 0.0f, 0.0f });
    V3f L = normalize3({ 1.0f, 1.0f, -1.0f });
    eye3[0] = e.x; eye3[1] = e.y; eye3[2] = e.z;
    light3[0] = L.x; light3[1] = L.y; light3[2] = L.z;
}
void orc_set_shader(void* h, int mode) { ((Oracle*)h)->shader = mode; }
void orc_set_albedo(void* h, float r, float g, float b) { Oracle& o = *(Oracle*)h; o.albedo[0] = r; o.albedo[1] = g; o.albedo[2] = b; }
void orc_set_texture_filter(void* h, int filter) { ((Oracle*)h)->texFilter = filter; }
void orc_clear_textures(void* h) { Oracle& o = *(Oracle*)h; o.textures.clear(); o.texIds.clear(); }
void orc_add_constant_texture(void* h, float r, float g, float b)
{
    Texture t; t.kind = 0; t.color[0] = r; t.color[1] = g; t.color[2] = b;
    ((Oracle*)h)->textures.push_back(std::move(t));
}
void orc_add_image_texture(void* h, const uint8_t* rgba8, int w, int ht)
{
    Texture t; t.kind = 1;
    TexLevel l; l.w = w; l.h = ht; l.px.assign(rgba8, rgba8 + (size_t)w * ht * 4);
    t.levels.push_back(std::move(l));
    build_mips(t);
    ((Oracle*)h)->textures.push_back(std::move(t));
}
void orc_set_texture_ids(void* h, const uint32_t* ids, uint32_t n) { ((Oracle*)h)->texIds.assign(ids, ids + n); }
// sampler alone, for known-answer tests: texture `slot`, filter, uv and the two differentials
void orc_tex_sample(void* h, uint32_t slot, int filter, float u, float v, float du0, float dv0, float du1, float dv1, float* out3)
{
    tex_sample(((Oracle*)h)->textures[slot], filter, u, v, du0, dv0, du1, dv1, out3);
}
int orc_tex_levels(void* h, uint32_t slot) { return (int)((Oracle*)h)->textures[slot].levels.size(); }
void orc_tex_level(void* h, uint32_t slot, int level, uint8_t* out)
{
    const TexLevel& l = ((Oracle*)h)->textures[slot].levels[level];
    memcpy(out, l.px.data(), l.px.size());
}
void orc_set_hierarchical(void* h, int on) { ((Oracle*)h)->hierarchical = on != 0; }

// Renderer.cpp:100-118. vtx: nv x 32-byte (pos3, normal3, uv2); idx: nt x 3 uint32.
void orc_render(void* h, const float* vtx, uint32_t nv, const uint32_t* idx, uint32_t nt)
{
    render(*(Oracle*)h, vtx, nv, idx, nt);
}

// Renderer::SetMSAAMode, Renderer.cpp:94-98 (re-creates the frame buffer)
void orc_set_msaa(void* h, int log2)
{
    Oracle& o = *(Oracle*)h;
    o.msLevel = log2; o.samples = 1 << log2;
    resize(o, o.W, o.H);
}
int orc_samples(void* h) { return ((Oracle*)h)->samples; }
// Renderer::GetBackBuffer, Renderer.cpp:360-363 / FrameBuffer.h:55-58: the resolved buffer (== sample buffer at 1x)
const uint8_t* orc_color(void* h) { return ((Oracle*)h)->resolved.data(); }
// per-sample owner ids: sample sId of every pixel, bottom-up
void orc_get_winner_sample(void* h, int sId, uint32_t* out)
{
    Oracle& o = *(Oracle*)h;
    const size_t n = (size_t)o.W * o.H;
    for (size_t p = 0; p < n; p++) out[p] = o.winner[p * o.samples + sId];
}
void orc_get_winner(void* h, uint32_t* out) { orc_get_winner_sample(h, 0, out); }
// depth of one sample, linearised bottom-up like the colour buffer
void orc_get_depth_sample(void* h, int sId, float* out)
{
    Oracle& o = *(Oracle*)h;
    for (int y = 0; y < o.H; y++)
        for (int x = 0; x < o.W; x++) {
            int tx = x >> TILE_LOG2, ty = y >> TILE_LOG2, ix = x & (TILE - 1), iy = y & (TILE - 1);
            const float* q = (const float*)&o.depth[((size_t)(ty * o.tilesX + tx) * o.samples + sId) * 256 + (size_t)(iy >> 1) * 16 + (ix >> 1)];
            out[(size_t)x + (size_t)o.W * (size_t)(o.H - 1 - y)] = q[(ix & 1) + 2 * (iy & 1)];
        }
}
void orc_get_depth(void* h, float* out) { orc_get_depth_sample(h, 0, out); }
void orc_get_clip_verts(void* h, float* out)
{
    Oracle& o = *(Oracle*)h;
    for (size_t i = 0; i < o.projected.size(); i++) memcpy(out + 4 * i, &o.projected[i].proj, 16);
}
uint64_t orc_num_raster_tris(void* h) { return ((Oracle*)h)->stats.nRasterTris; }
// per raster triangle, in submission order: ints[7] = prim, v0x,v0y,v1x,v1y,v2x,v2y ; floats[7] = z0,z1,z2,invW0,invW1,invW2,invDet
void orc_get_raster_tris(void* h, int32_t* ints, float* floats)
{
    Oracle& o = *(Oracle*)h;
    size_t k = 0;
    for (int c = 0; c < o.cores; c++)
        for (const RasterTri& t : o.coreTris[c]) {
            const ProjVertex* vb = o.coreVerts[c].data();
            int32_t* I = ints + 7 * k; float* F = floats + 7 * k;
            I[0] = (int32_t)t.primId; I[1] = t.v0.x; I[2] = t.v0.y; I[3] = t.v1.x; I[4] = t.v1.y; I[5] = t.v2.x; I[6] = t.v2.y;
            F[0] = vb[t.vId0].proj.z; F[1] = vb[t.vId1].proj.z; F[2] = vb[t.vId2].proj.z;
            F[3] = vb[t.vId0].invW; F[4] = vb[t.vId1].invW; F[5] = vb[t.vId2].invW; F[6] = t.invDet;
            k++;
        }
}
void orc_get_stats(void* h, uint64_t* counts4, double* ms6)
{
    Oracle& o = *(Oracle*)h;
    counts4[0] = o.stats.nRasterTris; counts4[1] = o.stats.nFragments; counts4[2] = o.stats.nCoveredSamples; counts4[3] = o.stats.nBinRefs;
    memcpy(ms6, o.stats.ms, sizeof(o.stats.ms));
}

// small known-answer helpers
int orc_snap(float f) { return snap_28_4(f); }
uint32_t orc_clip_code(float x, float y, float z, float w) { return clip_code({ x, y, z, w }); }
// clips one clip-space triangle; out_pos: up to 16 x 4 floats, out_wt: up to 16 x 3 floats; returns vertex count
int orc_clip_triangle(const float* tri12, float* out_pos, float* out_wt)
{
    V4f v0 = { tri12[0], tri12[1], tri12[2], tri12[3] }, v1 = { tri12[4], tri12[5], tri12[6], tri12[7] }, v2 = { tri12[8], tri12[9], tri12[10], tri12[11] };
    uint32_t c0 = clip_code(v0), c1 = clip_code(v1), c2 = clip_code(v2);
    if (!(c0 | c1 | c2)) return -1;          // not clipped
    if (c0 & c1 & c2) return 0;              // rejected
    Poly a, b; Poly* r;
    a.n = 3; a.v[0] = { v0, { 1, 0, 0 } }; a.v[1] = { v1, { 0, 1, 0 } }; a.v[2] = { v2, { 0, 0, 1 } };
    clip_polygon(a, b, (c0 ^ c1) | (c1 ^ c2) | (c2 ^ c0), r);
    for (int i = 0; i < r->n; i++) { memcpy(out_pos + 4 * i, &r->v[i].pos, 16); memcpy(out_wt + 3 * i, &r->v[i].wt, 12); }
    return r->n;
}
void orc_mat_mul(const float* a, const float* b, float* out) { Mat4 A, B; memcpy(A.m, a, 64); memcpy(B.m, b, 64); Mat4 r = mat_mul(A, B); memcpy(out, r.m, 64); }
void orc_mat_inverse(const float* a, float* out) { Mat4 A; memcpy(A.m, a, 64); Mat4 r = mat_inverse(A); memcpy(out, r.m, 64); }

} // extern "C"
