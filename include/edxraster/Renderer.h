// edxraster/Renderer.h — C++ host API with the reference's renderer-facing names, over the C ABI.
//
// Mirrors EDX::RasterRenderer::Renderer (EDXRaster/Core/Renderer.h:36-50), Mesh (Utils/Mesh.h:17-69),
// IVertexBuffer / VertexBuffer / IndexBuffer (Utils/InputBuffer.h:43-205) and the parts of EDXUtil the
// viewer touches (Matrix, Vector3, Camera; RealtimeViewer/Main.cpp:37-42,69-75). A program written
// against the reference's Renderer compiles against this header by switching the namespace; everything
// executes in libedxraster_b200.so (CUDA, sm_100a). Header-only; link with -ledxraster_b200.
//
// Differences, all additive: SetPixelShader (the reference hard-codes its shader, Renderer.cpp:41),
// GetDepthBuffer, a device ordinal in the constructor, and status codes through LastStatus()/LastError()
// (the reference has void returns and no error reporting).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <tuple>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../edxraster_c.h"

namespace edx_b200 {

typedef unsigned int uint;
typedef unsigned char _byte;

struct Vector2 { float x, y; Vector2(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct Vector3 {
    float x, y, z;
    Vector3(float x_ = 0, float y_ = 0, float z_ = 0) : x(x_), y(y_), z(z_) {}
    Vector3 operator-(const Vector3& b) const { return Vector3(x - b.x, y - b.y, z - b.z); }
    Vector3 operator+(const Vector3& b) const { return Vector3(x + b.x, y + b.y, z + b.z); }
    Vector3 operator*(float s) const { return Vector3(x * s, y * s, z * s); }
    static float Dot(const Vector3& a, const Vector3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    static Vector3 Cross(const Vector3& a, const Vector3& b) { return Vector3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
    static Vector3 Normalize(const Vector3& v) { float l = std::sqrt(Dot(v, v)); return Vector3(v.x / l, v.y / l, v.z / l); }
};

// 4x4, m[row][col], column-vector convention (clip = Proj * ModelView * p; Renderer.cpp:90)
class Matrix {
public:
    float m[4][4];
    Matrix() { std::memset(m, 0, sizeof(m)); m[0][0] = m[1][1] = m[2][2] = m[3][3] = 1.0f; }
    const float* Data() const { return &m[0][0]; }
    Matrix operator*(const Matrix& b) const
    {
        Matrix r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                r.m[i][j] = ((m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j]) + m[i][2] * b.m[2][j]) + m[i][3] * b.m[3][j];
        return r;
    }
    static Matrix LookAt(const Vector3& eye, const Vector3& target, const Vector3& up)      // left-handed
    {
        Vector3 z = Vector3::Normalize(target - eye), x = Vector3::Normalize(Vector3::Cross(up, z)), y = Vector3::Cross(z, x);
        Matrix r;
        r.m[0][0] = x.x; r.m[0][1] = x.y; r.m[0][2] = x.z; r.m[0][3] = -Vector3::Dot(x, eye);
        r.m[1][0] = y.x; r.m[1][1] = y.y; r.m[1][2] = y.z; r.m[1][3] = -Vector3::Dot(y, eye);
        r.m[2][0] = z.x; r.m[2][1] = z.y; r.m[2][2] = z.z; r.m[2][3] = -Vector3::Dot(z, eye);
        return r;
    }
    static Matrix Perspective(float fovYDeg, float aspect, float zn, float zf)            // D3D depth: z in [0, w]
    {
        const float ys = 1.0f / std::tan(fovYDeg * 3.14159265358979323846f / 360.0f);
        Matrix r;
        r.m[0][0] = ys / aspect; r.m[1][1] = ys; r.m[2][2] = zf / (zf - zn); r.m[2][3] = -zn * zf / (zf - zn);
        r.m[3][2] = 1.0f; r.m[3][3] = 0.0f;
        return r;
    }
    static Matrix Raster(int w, int h)                                                    // NDC -> pixels, y down
    {
        Matrix r;
        r.m[0][0] = w * 0.5f; r.m[0][3] = w * 0.5f; r.m[1][1] = -h * 0.5f; r.m[1][3] = h * 0.5f;
        return r;
    }
};

// Stand-in for EDXUtil's Camera as the viewer uses it (Main.cpp:39,69-71,146)
class Camera {
public:
    void Init(const Vector3& pos, const Vector3& target, const Vector3& up, int w, int h, float fov = 65.0f, float zn = 0.01f, float zf = 100.0f)
    {
        mPos = pos; mTarget = target; mUp = up; mFov = fov; mNear = zn; mFar = zf;
        Resize(w, h);
    }
    void Resize(int w, int h)
    {
        mW = w; mH = h;
        mView = Matrix::LookAt(mPos, mTarget, mUp);
        mProj = Matrix::Perspective(mFov, float(w) / float(h), mNear, mFar);
        mRaster = Matrix::Raster(w, h);
    }
    void Transform() {}
    const Matrix& GetViewMatrix() const { return mView; }
    const Matrix& GetProjMatrix() const { return mProj; }
    const Matrix& GetRasterMatrix() const { return mRaster; }
    Vector3 mPos, mTarget, mUp;
private:
    float mFov = 65, mNear = 0.01f, mFar = 100; int mW = 0, mH = 0;
    Matrix mView, mProj, mRaster;
};

// ---- Utils/InputBuffer.h ---------------------------------------------------------------------
struct Vertex_PositionNormalTex { Vector3 Position; Vector3 Normal; Vector2 TexCoord; static const int Size = 32; };
static_assert(sizeof(Vertex_PositionNormalTex) == 32, "submission format is 32 bytes per vertex (InputBuffer.h:16-28)");

class IVertexBuffer {
public:
    virtual ~IVertexBuffer() {}
    virtual void* GetBuffer() const = 0;
    virtual int GetVertexSize() const = 0;
    virtual size_t GetBufferSize() const = 0;
    virtual Vector3 GetPosition(uint idx) const = 0;
    virtual Vector3 GetNormal(uint idx) const = 0;
    virtual Vector2 GetTexCoord(uint idx) const = 0;
    uint GetVertexCount() const { return mVertexCount; }
protected:
    uint mVertexCount = 0;
};

template <typename VertexType = Vertex_PositionNormalTex>
class VertexBuffer : public IVertexBuffer {
public:
    void NewBuffer(uint n) { mVertexCount = n; mData.resize(n); }
    void* GetBuffer() const override { return (void*)mData.data(); }
    int GetVertexSize() const override { return VertexType::Size; }
    size_t GetBufferSize() const override { return mData.size() * sizeof(VertexType); }
    Vector3 GetPosition(uint i) const override { return mData[i].Position; }
    Vector3 GetNormal(uint i) const override { return mData[i].Normal; }
    Vector2 GetTexCoord(uint i) const override { return mData[i].TexCoord; }
private:
    std::vector<VertexType> mData;
};

template <typename VertexType = Vertex_PositionNormalTex>
inline IVertexBuffer* CreateVertexBuffer(const void* pData, size_t vertexCount)     // InputBuffer.h:136-146
{
    auto* vb = new VertexBuffer<VertexType>;
    vb->NewBuffer((uint)vertexCount);
    std::memcpy(vb->GetBuffer(), pData, vb->GetBufferSize());
    return vb;
}

class IndexBuffer {                                                                  // InputBuffer.h:148-194
public:
    void ResizeBuffer(uint triCount) { mBuffer.resize(3 * (size_t)triCount); }
    uint* GetBuffer() { return mBuffer.data(); }
    const uint* GetBuffer() const { return mBuffer.data(); }
    uint GetTriangleCount() const { return (uint)(mBuffer.size() / 3); }
    size_t GetBufferSize() const { return mBuffer.size(); }
    const uint* GetIndex(uint idx) const { return &mBuffer[3 * (size_t)idx]; }
    void AppendTriangle(int a, int b, int c) { mBuffer.push_back(a); mBuffer.push_back(b); mBuffer.push_back(c); }
private:
    std::vector<uint> mBuffer;
};

inline IndexBuffer* CreateIndexBuffer(const void* pData, size_t triCount)           // InputBuffer.h:196-205
{
    auto* ib = new IndexBuffer;
    ib->ResizeBuffer((uint)triCount);
    std::memcpy(ib->GetBuffer(), pData, ib->GetBufferSize() * sizeof(uint));
    return ib;
}

// EDXUtil's BoundingBox as far as Mesh::GetBounds needs it (Utils/Mesh.h:26,63-66)
struct BoundingBox { Vector3 mMin, mMax; };

// ---- Utils/Mesh.h ----------------------------------------------------------------------------
class Mesh {
public:
    ~Mesh() { Release(); }
    // The reference delegates these to EDXUtil's ObjMesh (Mesh.cpp:36-70), which is unavailable; the
    // generators below are ours (UV sphere, slices x stacks quads; unit plane in the xz-plane).
    void LoadSphere(const Vector3& pos, const Vector3& scl, const Vector3& rot, float radius, int slices = 64, int stacks = 64)
    {
        (void)rot;
        std::vector<Vertex_PositionNormalTex> v;
        std::vector<uint> idx;
        const double PI = 3.14159265358979323846;
        for (int i = 0; i <= stacks; i++)
            for (int j = 0; j <= slices; j++) {
                double t = PI * i / stacks, p = 2.0 * PI * j / slices;
                Vector3 n((float)(std::sin(t) * std::cos(p)), (float)std::cos(t), (float)(std::sin(t) * std::sin(p)));
                Vertex_PositionNormalTex o;
                o.Position = Vector3(pos.x + scl.x * radius * n.x, pos.y + scl.y * radius * n.y, pos.z + scl.z * radius * n.z);
                o.Normal = n;
                o.TexCoord = Vector2((float)(p / (2.0 * PI)), (float)(t / PI));
                v.push_back(o);
            }
        for (int i = 0; i < stacks; i++)
            for (int j = 0; j < slices; j++) {
                uint a = i * (slices + 1) + j, b = a + 1, c = a + slices + 1, d = c + 1;
                idx.insert(idx.end(), { a, b, c, b, d, c });
            }
        SetBuffers(v.data(), v.size(), idx.data(), idx.size() / 3);
        AddConstantTexture(0.9f, 0.9f, 0.9f);          // Mesh.cpp:47
    }
    void LoadPlane(const Vector3& pos, const Vector3& scl, const Vector3& rot, float length)
    {
        (void)rot;
        const float h = 0.5f * length;
        Vertex_PositionNormalTex v[4];
        const float sx[4] = { -h, h, -h, h }, sz[4] = { -h, -h, h, h };
        for (int k = 0; k < 4; k++) {
            v[k].Position = Vector3(pos.x + scl.x * sx[k], pos.y, pos.z + scl.z * sz[k]);
            v[k].Normal = Vector3(0, 1, 0);
            v[k].TexCoord = Vector2(k & 1 ? 1.0f : 0.0f, k & 2 ? 1.0f : 0.0f);
        }
        const uint idx[6] = { 0, 2, 1, 1, 2, 3 };
        SetBuffers(v, 4, idx, 2);
        AddConstantTexture(0.9f, 0.9f, 0.9f);          // Mesh.cpp:66
    }
    // Mesh::LoadMesh (Utils/Mesh.cpp:11-34): Wavefront OBJ. The reference delegates to EDXUtil's ObjMesh (absent);
    // this reader handles v / vt / vn / f (polygons fanned, negative indices, v, v/vt, v//vn, v/vt/vn), builds one
    // vertex per distinct index triple, computes area-weighted normals when the file has none, and applies
    // scale, then rotation (degrees, about x then y then z), then translation. Materials (Mesh.cpp:22-31): `mtllib` files
    // are read for `newmtl` / `Kd` / `map_Kd`; every material becomes one texture slot in file order - an
    // ImageTexture when map_Kd names a readable uncompressed 24/32-bit BMP (the only decoder here; EDXUtil's image
    // loaders are absent), else a ConstantTexture2D of Kd - and every face carries the slot of the `usemtl` in force
    // (Mesh::GetTextureIds). A file without materials gets one constant 0.9 white slot like LoadSphere.
    bool LoadMesh(const Vector3& pos, const Vector3& scl, const Vector3& rot, const char* path)
    {
        FILE* f = std::fopen(path, "r");
        if (!f) return false;
        std::vector<Vector3> P, N;
        std::vector<Vector2> T;
        std::vector<Vertex_PositionNormalTex> verts;
        std::vector<uint> idx;
        std::map<std::tuple<int, int, int>, uint> seen;
        struct Material { std::string name; float kd[3] = { 0.9f, 0.9f, 0.9f }; std::string map; };
        std::vector<Material> mats;
        std::vector<uint> faceSlot;                        // per triangle
        uint curSlot = 0;
        const std::string dir = DirOf(path);
        char line[1024];
        while (std::fgets(line, sizeof(line), f)) {
            if (!std::strncmp(line, "mtllib ", 7)) {
                FILE* mf = std::fopen((dir + Trim(line + 7)).c_str(), "r");
                if (mf) {
                    char ml[1024];
                    while (std::fgets(ml, sizeof(ml), mf)) {
                        char* q = ml; while (*q == ' ' || *q == '\t') q++;
                        if (!std::strncmp(q, "newmtl ", 7)) { mats.emplace_back(); mats.back().name = Trim(q + 7); }
                        else if (!mats.empty() && !std::strncmp(q, "Kd ", 3)) std::sscanf(q + 3, "%f %f %f", &mats.back().kd[0], &mats.back().kd[1], &mats.back().kd[2]);
                        else if (!mats.empty() && !std::strncmp(q, "map_Kd ", 7)) mats.back().map = Trim(q + 7);
                    }
                    std::fclose(mf);
                }
            }
            else if (!std::strncmp(line, "usemtl ", 7)) {
                const std::string name = Trim(line + 7);
                for (size_t k = 0; k < mats.size(); k++) if (mats[k].name == name) curSlot = (uint)k;
            }
            else if (line[0] == 'v' && line[1] == ' ') { Vector3 v; if (std::sscanf(line + 2, "%f %f %f", &v.x, &v.y, &v.z) == 3) P.push_back(v); }
            else if (line[0] == 'v' && line[1] == 'n') { Vector3 v; if (std::sscanf(line + 3, "%f %f %f", &v.x, &v.y, &v.z) == 3) N.push_back(v); }
            else if (line[0] == 'v' && line[1] == 't') { Vector2 v; if (std::sscanf(line + 3, "%f %f", &v.x, &v.y) >= 1) T.push_back(v); }
            else if (line[0] == 'f' && line[1] == ' ') {
                std::vector<uint> poly;
                char* p = line + 2;
                while (*p) {
                    while (*p == ' ' || *p == '\t') p++;
                    if (*p == '\0' || *p == '\n' || *p == '\r') break;
                    int vi = 0, ti = 0, ni = 0;
                    vi = (int)std::strtol(p, &p, 10);
                    if (*p == '/') { p++; if (*p != '/') ti = (int)std::strtol(p, &p, 10); if (*p == '/') { p++; ni = (int)std::strtol(p, &p, 10); } }
                    if (vi < 0) vi = (int)P.size() + vi + 1;
                    if (ti < 0) ti = (int)T.size() + ti + 1;
                    if (ni < 0) ni = (int)N.size() + ni + 1;
                    if (vi < 1 || vi > (int)P.size()) continue;
                    auto key = std::make_tuple(vi, ti, ni);
                    auto it = seen.find(key);
                    if (it == seen.end()) {
                        Vertex_PositionNormalTex o;
                        o.Position = P[vi - 1];
                        o.Normal = (ni >= 1 && ni <= (int)N.size()) ? N[ni - 1] : Vector3(0, 0, 0);
                        o.TexCoord = (ti >= 1 && ti <= (int)T.size()) ? T[ti - 1] : Vector2(0, 0);
                        it = seen.emplace(key, (uint)verts.size()).first;
                        verts.push_back(o);
                    }
                    poly.push_back(it->second);
                }
                for (size_t k = 2; k < poly.size(); k++) { idx.push_back(poly[0]); idx.push_back(poly[k - 1]); idx.push_back(poly[k]); faceSlot.push_back(curSlot); }
            }
        }
        std::fclose(f);
        if (verts.empty() || idx.empty()) return false;
        if (N.empty()) {                                   // area-weighted vertex normals
            for (size_t t = 0; t + 2 < idx.size(); t += 3) {
                Vector3 a = verts[idx[t]].Position, b = verts[idx[t + 1]].Position, c = verts[idx[t + 2]].Position;
                Vector3 n = Vector3::Cross(b - a, c - a);
                for (int k = 0; k < 3; k++) verts[idx[t + k]].Normal = verts[idx[t + k]].Normal + n;
            }
            for (auto& v : verts) { float l = std::sqrt(Vector3::Dot(v.Normal, v.Normal)); if (l > 0) v.Normal = v.Normal * (1.0f / l); }
        }
        const float d2r = 3.14159265358979323846f / 180.0f;
        const float cx = std::cos(rot.x * d2r), sx = std::sin(rot.x * d2r), cy = std::cos(rot.y * d2r), sy = std::sin(rot.y * d2r);
        const float cz = std::cos(rot.z * d2r), sz = std::sin(rot.z * d2r);
        auto rotate = [&](Vector3 v) {
            v = Vector3(v.x, cx * v.y - sx * v.z, sx * v.y + cx * v.z);
            v = Vector3(cy * v.x + sy * v.z, v.y, -sy * v.x + cy * v.z);
            return Vector3(cz * v.x - sz * v.y, sz * v.x + cz * v.y, v.z);
        };
        for (auto& v : verts) {
            v.Position = rotate(Vector3(v.Position.x * scl.x, v.Position.y * scl.y, v.Position.z * scl.z)) + pos;
            v.Normal = rotate(v.Normal);
        }
        SetBuffers(verts.data(), verts.size(), idx.data(), idx.size() / 3);
        if (mats.empty()) AddConstantTexture(0.9f, 0.9f, 0.9f);
        for (const Material& m : mats) {
            std::vector<_byte> texels; uint tw = 0, th = 0;
            if (!m.map.empty() && LoadBmp((dir + m.map).c_str(), texels, tw, th)) AddImageTexture(texels.data(), tw, th);
            else AddConstantTexture(m.kd[0], m.kd[1], m.kd[2]);
        }
        SetTextureIds(faceSlot);
        return true;
    }
    // Uncompressed 24- or 32-bit BMP -> RGBA8, first row = top of the picture (v = 0 at the top, the D3D convention
    // the reference's raster space uses). Stands in for EDXUtil's bitmap loader behind ImageTexture (Mesh.cpp:27).
    static bool LoadBmp(const char* path, std::vector<_byte>& rgba, uint& w, uint& h)
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) return false;
        unsigned char hd[54];
        bool ok = std::fread(hd, 1, 54, f) == 54 && hd[0] == 'B' && hd[1] == 'M';
        auto u32 = [&](int o) { return (uint32_t)hd[o] | ((uint32_t)hd[o + 1] << 8) | ((uint32_t)hd[o + 2] << 16) | ((uint32_t)hd[o + 3] << 24); };
        const uint32_t off = ok ? u32(10) : 0;
        const int32_t bw = ok ? (int32_t)u32(18) : 0, bh = ok ? (int32_t)u32(22) : 0;
        const int bpp = ok ? (hd[28] | (hd[29] << 8)) : 0;
        ok = ok && u32(30) == 0 && (bpp == 24 || bpp == 32) && bw > 0 && bh != 0 && bw <= 32768 && std::abs(bh) <= 32768;
        if (ok) {
            w = (uint)bw; h = (uint)std::abs(bh);
            const size_t stride = ((size_t)w * (bpp / 8) + 3) & ~(size_t)3;
            std::vector<unsigned char> row(stride);
            rgba.assign((size_t)w * h * 4, 255);
            ok = std::fseek(f, (long)off, SEEK_SET) == 0;
            for (uint y = 0; ok && y < h; y++) {
                ok = std::fread(row.data(), 1, stride, f) == stride;
                const uint dy = bh > 0 ? h - 1 - y : y;                   // positive height = bottom-up file
                for (uint x = 0; ok && x < w; x++) {
                    const unsigned char* p = &row[(size_t)x * (bpp / 8)];
                    _byte* o = &rgba[4 * ((size_t)dy * w + x)];
                    o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; o[3] = bpp == 32 ? p[3] : 255;
                }
            }
        }
        std::fclose(f);
        return ok;
    }
    static std::string Trim(const char* s)
    {
        std::string t(s);
        while (!t.empty() && (t.back() == '\n' || t.back() == '\r' || t.back() == ' ' || t.back() == '\t')) t.pop_back();
        size_t b = 0; while (b < t.size() && (t[b] == ' ' || t[b] == '\t')) b++;
        return t.substr(b);
    }
    static std::string DirOf(const char* path)
    {
        const std::string p(path);
        const size_t k = p.find_last_of("/\\");
        return k == std::string::npos ? std::string() : p.substr(0, k + 1);
    }
    // raw submission in the reference's wire format (CreateVertexBuffer / CreateIndexBuffer)
    void SetBuffers(const void* vertices, size_t vertexCount, const uint* indices, size_t triCount)
    {
        Release();
        mpVertexBuf.reset(CreateVertexBuffer<>(vertices, vertexCount));
        mpIndexBuf.reset(CreateIndexBuffer(indices, triCount));
        mTexIdx.assign(triCount, 0u);
    }
    const IVertexBuffer* GetVertexBuffer() const { return mpVertexBuf.get(); }
    IndexBuffer* GetIndexBuffer() const { return mpIndexBuf.get(); }
    // Mesh::GetBounds (Mesh.h:63-66): object-space bounds of the submitted vertices
    BoundingBox GetBounds() const
    {
        BoundingBox b;
        if (!mpVertexBuf || !mpVertexBuf->GetVertexCount()) return b;
        const float* v = (const float*)mpVertexBuf->GetBuffer();
        b.mMin = b.mMax = Vector3(v[0], v[1], v[2]);
        for (uint i = 1; i < mpVertexBuf->GetVertexCount(); i++) {
            const float* p = v + 8 * (size_t)i;
            b.mMin = Vector3(std::min(b.mMin.x, p[0]), std::min(b.mMin.y, p[1]), std::min(b.mMin.z, p[2]));
            b.mMax = Vector3(std::max(b.mMax.x, p[0]), std::max(b.mMax.y, p[1]), std::max(b.mMax.z, p[2]));
        }
        return b;
    }
    const std::vector<uint>& GetTextureIds() const { return mTexIdx; }
    // Mesh::mTextures (Mesh.h:23): ConstantTexture2D<Color>(colour) or ImageTexture<Color, Color4b> (Mesh.cpp:27,29).
    // The reference decodes image files through EDXUtil; here the caller hands over decoded RGBA8 texels (row 0 at
    // v = 0). Returns the slot index that SetTextureIds refers to.
    uint AddConstantTexture(float r, float g, float b)
    {
        Texture t; t.kind = EDX_TEXTURE_CONSTANT; t.color[0] = r; t.color[1] = g; t.color[2] = b;
        mTextures.push_back(t); mTexDirty = true;
        return (uint)mTextures.size() - 1;
    }
    uint AddImageTexture(const _byte* rgba8, uint width, uint height)
    {
        Texture t; t.kind = EDX_TEXTURE_IMAGE; t.width = width; t.height = height;
        t.texels.assign(rgba8, rgba8 + (size_t)width * height * 4);
        mTextures.push_back(t); mTexDirty = true;
        return (uint)mTextures.size() - 1;
    }
    void SetTextureIds(const std::vector<uint>& perTriangle) { mTexIdx = perTriangle; mTexDirty = true; }
    size_t GetTextureCount() const { return mTextures.size(); }
    void GetTexture(size_t slot, int& kind, float color[3], uint& width, uint& height, const _byte*& texels) const
    {
        const Texture& t = mTextures[slot];
        kind = t.kind; color[0] = t.color[0]; color[1] = t.color[1]; color[2] = t.color[2];
        width = t.width; height = t.height; texels = t.texels.data();
    }
    void Release()
    {
        if (mDevice) ReleaseDevice();
        mDevice = nullptr; mOwner = nullptr;
        mpVertexBuf.reset(); mpIndexBuf.reset(); mTexIdx.clear(); mTextures.clear(); mTexDirty = false;
    }
private:
    friend class Renderer;
    inline void ReleaseDevice() const;           // defined after Renderer (needs its registry of live contexts)
    struct Texture { int kind = 0; float color[3] = { 0, 0, 0 }; uint width = 0, height = 0; std::vector<_byte> texels; };
    std::vector<Texture> mTextures;
    mutable bool mTexDirty = false;
    std::unique_ptr<IVertexBuffer> mpVertexBuf;
    std::unique_ptr<IndexBuffer> mpIndexBuf;
    std::vector<uint> mTexIdx;
    mutable edx_mesh* mDevice = nullptr;        // device copy, made on first RenderMesh
    mutable edx_context* mOwner = nullptr;
};

// ---- Core/Scene.h: an array of meshes the Renderer constructs and never reads (Renderer.cpp:35-38); kept API-shaped
class Scene {
public:
    void AddMesh(Mesh* pMesh) { mMeshes.emplace_back(pMesh); }
    size_t GetMeshCount() const { return mMeshes.size(); }
private:
    std::vector<std::unique_ptr<Mesh>> mMeshes;
};

enum class TextureFilter { Nearest = 0, Linear = 1, TriLinear = 2, Anisotropic4x = 3, Anisotropic8x = 4, Anisotropic16x = 5 };
enum class PixelShaderKind { DepthOnly = EDX_SHADER_DEPTH_ONLY, BlinnPhong = EDX_SHADER_BLINN_PHONG, Lambertian = EDX_SHADER_LAMBERT, LambertianAlbedo = EDX_SHADER_LAMBERT_ALBEDO };

// ---- Core/Renderer.h -------------------------------------------------------------------------
class Renderer {
public:
    explicit Renderer(int device = 0) : mDeviceId(device) { mStatus = edx_create(device, &mCtx); if (mCtx) { LiveContexts().insert(mCtx); DeviceOf()[mCtx] = device; } }
    // contexts that are still alive, so a Mesh that outlives its Renderer releases its device copy safely
    static std::set<edx_context*>& LiveContexts() { static std::set<edx_context*> s; return s; }
    static std::map<edx_context*, int>& DeviceOf() { static std::map<edx_context*, int> m; return m; }
    ~Renderer() { if (mCtx) { LiveContexts().erase(mCtx); DeviceOf().erase(mCtx); edx_destroy(mCtx); } }
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    void Initialize(uint w, uint h) { mW = w; mH = h; Call(edx_initialize(mCtx, w, h)); }
    void Resize(uint w, uint h) { mW = w; mH = h; Call(edx_resize(mCtx, w, h)); }
    void SetTransform(const Matrix& modelView, const Matrix& proj, const Matrix& toRaster)
    {
        Call(edx_set_transform(mCtx, modelView.Data(), proj.Data(), toRaster.Data()));
    }
    void RenderMesh(const Mesh& mesh)
    {
        if (!mCtx) return;
        // a device copy made by another Renderer on the same GPU is shared (a mesh is read-only while it renders)
        bool upload = !mesh.mDevice || !LiveContexts().count(mesh.mOwner);
        if (!upload && mesh.mOwner != mCtx && !SameDevice(mesh.mOwner)) upload = true;
        if (upload) {
            if (mesh.mDevice) mesh.ReleaseDevice();
            mesh.mDevice = nullptr;
            const IVertexBuffer* vb = mesh.GetVertexBuffer();
            IndexBuffer* ib = mesh.GetIndexBuffer();
            if (!vb || !ib) { mStatus = EDX_ERR_INVALID; return; }
            Call(edx_mesh_create(mCtx, vb->GetBuffer(), vb->GetVertexCount(), ib->GetBuffer(), ib->GetTriangleCount(),
                                 mesh.GetTextureIds().data(), &mesh.mDevice));
            mesh.mOwner = mCtx;
            if (mStatus != EDX_OK) return;
            mesh.mTexDirty = !mesh.mTextures.empty();
        }
        if (mesh.mTexDirty) {                    // Renderer.cpp:106: RenderStates::TextureSlots = &mesh.GetTextures()
            std::vector<edx_texture_desc> descs(mesh.mTextures.size());
            for (size_t i = 0; i < descs.size(); i++) {
                const Mesh::Texture& t = mesh.mTextures[i];
                descs[i].kind = t.kind; descs[i].rgba8 = t.texels.data(); descs[i].width = t.width; descs[i].height = t.height;
                for (int k = 0; k < 3; k++) descs[i].color[k] = t.color[k];
            }
            const bool ids = mesh.mTexIdx.size() == mesh.GetIndexBuffer()->GetTriangleCount();
            Call(edx_mesh_set_textures(mCtx, mesh.mDevice, descs.data(), (uint32_t)descs.size(), ids ? mesh.mTexIdx.data() : nullptr));
            mesh.mTexDirty = false;
            if (mStatus != EDX_OK) return;
        }
        Call(edx_render_mesh(mCtx, mesh.mDevice));
        if (mWriteFrames) WriteFrameToFile();
        mFrameCount++;
    }
    void WriteFrameToFile() const
    {
        char name[64];
        std::snprintf(name, sizeof(name), "Frame%05i.bmp", mFrameCount);      // Renderer.cpp:355
        edx_write_frame_to_file(mCtx, name);
    }
    const _byte* GetBackBuffer() const { return mCtx ? edx_get_back_buffer(mCtx) : nullptr; }
    void SetMSAAMode(int msaaCountLog2) { Call(edx_set_msaa_mode(mCtx, msaaCountLog2)); }
    void SetTextureFilter(TextureFilter f) { Call(edx_set_texture_filter(mCtx, (int)f)); }
    void SetHierarchicalRasterize(bool h) { Call(edx_set_hierarchical_rasterize(mCtx, h ? 1 : 0)); }
    void SetWriteFrames(bool wf) { mWriteFrames = wf; }

    // extensions
    void SetPixelShader(PixelShaderKind k) { Call(edx_set_pixel_shader(mCtx, (int)k)); }
    bool GetDepthBuffer(float* out) const { return mCtx && edx_read_depth(mCtx, out) == EDX_OK; }
    bool WriteFrame(const char* path) const { return mCtx && edx_write_frame_to_file(mCtx, path) == EDX_OK; }
    void Synchronize() { Call(edx_synchronize(mCtx)); }
    // frame-parallel gather: every finished frame is pushed to these device addresses (own or peer-mapped) by the copy engine
    void SetFrameSink(void* remoteColor, void* remoteDepth) { Call(edx_set_frame_sink(mCtx, remoteColor, remoteDepth)); }
    // ... and the count of frames pushed so far, stored behind every frame's pushes where the consumer can poll it
    void SetFrameSinkSignal(void* remoteWord) { Call(edx_set_frame_sink_signal(mCtx, remoteWord)); }
    void FlushFrameSink() { Call(edx_flush_frame_sink(mCtx)); }     // the context's stream waits for the pushes issued so far
    int LastStatus() const { return mStatus; }
    const char* LastError() const { return mCtx ? edx_last_error(mCtx) : "no CUDA device (edx_create failed)"; }
    edx_context* Handle() const { return mCtx; }
    int Device() const { return mDeviceId; }
    uint Width() const { return mW; }
    uint Height() const { return mH; }
private:
    void Call(int rc) { if (!mCtx) { mStatus = EDX_ERR_NO_DEVICE; return; } mStatus = rc; }
    bool SameDevice(edx_context* other) const { auto it = DeviceOf().find(other); return it != DeviceOf().end() && it->second == mDeviceId; }
    edx_context* mCtx = nullptr;
    int mDeviceId = 0;
    int mStatus = EDX_OK;
    uint mW = 0, mH = 0;
    bool mWriteFrames = false;
    int mFrameCount = 0;
};

// Several frames in flight on one GPU: `depth` Renderers on their own streams that share the meshes. One frame
// is three dependent kernels of very different shapes and leaves much of a B200 idle; 3-4 overlapping frames
// raise rendering throughput 1.3-3x. Submit() = SetTransform + RenderMesh on the next lane
// and returns a ticket; GetBackBuffer(ticket) waits for that frame only. A ticket is valid until `depth` more
// frames have been submitted. (The reference renders one frame at a time, Core/Renderer.cpp:100-118.)
class FrameRing {
public:
    explicit FrameRing(int depth = 3, int device = 0) { for (int i = 0; i < (depth < 1 ? 1 : depth); i++) mLanes.emplace_back(new Renderer(device)); }
    void Initialize(uint w, uint h) { for (auto& r : mLanes) r->Initialize(w, h); }
    void Resize(uint w, uint h) { for (auto& r : mLanes) r->Resize(w, h); }
    void SetMSAAMode(int log2) { for (auto& r : mLanes) r->SetMSAAMode(log2); }
    void SetTextureFilter(TextureFilter f) { for (auto& r : mLanes) r->SetTextureFilter(f); }
    void SetHierarchicalRasterize(bool h) { for (auto& r : mLanes) r->SetHierarchicalRasterize(h); }
    void SetPixelShader(PixelShaderKind k) { for (auto& r : mLanes) r->SetPixelShader(k); }
    void Synchronize() { for (auto& r : mLanes) r->Synchronize(); }
    size_t Submit(const Mesh& mesh, const Matrix& modelView, const Matrix& proj, const Matrix& toRaster)
    {
        Renderer& r = *mLanes[mNext % mLanes.size()];
        r.SetTransform(modelView, proj, toRaster);
        r.RenderMesh(mesh);                      // the first lane to see a mesh uploads it; the others share that copy
        return mNext++;
    }
    bool InRing(size_t ticket) const { return ticket < mNext && ticket + mLanes.size() >= mNext; }
    const _byte* GetBackBuffer(size_t ticket) { return InRing(ticket) ? Lane(ticket).GetBackBuffer() : nullptr; }
    bool GetDepthBuffer(size_t ticket, float* out) { return InRing(ticket) && Lane(ticket).GetDepthBuffer(out); }
    Renderer& Lane(size_t ticket) { return *mLanes[ticket % mLanes.size()]; }
    Renderer& NextLane() { return *mLanes[mNext % mLanes.size()]; }      // the lane the next Submit renders on
    size_t Depth() const { return mLanes.size(); }
    int LastStatus() const { for (auto& r : mLanes) if (r->LastStatus() != EDX_OK) return r->LastStatus(); return EDX_OK; }
private:
    std::vector<std::unique_ptr<Renderer>> mLanes;
    size_t mNext = 0;
};

// Frame-parallel rendering on the GPUs of one box (BASELINE config 5): the C++ counterpart of
// RealtimeViewer's per-frame loop (Main.cpp:71-75) run once per GPU. One FrameRing per GPU, view i on GPU i mod N;
// every finished colour buffer is pushed by the copy engine straight into the root GPU's frame store over NVLink
// (edx_set_frame_sink) - no collective, no host copy per frame. One Mesh per GPU (a Mesh owns one device copy).
struct ViewTransform { Matrix modelView, proj, toRaster; };

class FrameFarm {
public:
    explicit FrameFarm(int gpus = 0, int depth = 3)
    {
        const int have = edx_device_count();
        mGpus = gpus <= 0 || gpus > have ? have : gpus;
        for (int g = 0; g < mGpus; g++) mRings.emplace_back(new FrameRing(depth, g));
        for (int g = 0; g < mGpus; g++) mMeshes.emplace_back(new Mesh);
    }
    ~FrameFarm() { ReleaseStore(); }
    int Gpus() const { return mGpus; }
    void Initialize(uint w, uint h)
    {
        mW = w; mH = h;
        for (auto& r : mRings) r->Initialize(w, h);
        // every GPU may write the root's memory
        for (int g = 1; g < mGpus; g++)
            for (size_t l = 0; l < mRings[g]->Depth(); l++)
                if (edx_enable_peer_access(mRings[g]->Lane(l).Handle(), 0) != EDX_OK) mStatus = EDX_ERR_UNSUPPORTED;
    }
    void SetPixelShader(PixelShaderKind k) { for (auto& r : mRings) r->SetPixelShader(k); }
    void SetTextureFilter(TextureFilter f) { for (auto& r : mRings) r->SetTextureFilter(f); }
    // build(mesh) fills the Mesh of one GPU; called once per GPU
    template <class Build> void LoadMeshes(Build build) { for (auto& m : mMeshes) build(*m); }
    // Renders every view; returns false on error. Frames stay in the root GPU's store until the next Render.
    bool Render(const std::vector<ViewTransform>& views)
    {
        if (mGpus == 0 || mStatus != EDX_OK) return false;
        const size_t frameBytes = (size_t)mW * mH * 4;
        if (views.size() > mStoreFrames) {
            ReleaseStore();
            if (edx_device_alloc(Root(), views.size() * frameBytes, &mStore) != EDX_OK) return false;
            mStoreFrames = views.size();
        }
        // one submitting thread per GPU (a frame costs 10-30 us of host time to submit; contexts of different GPUs are independent)
        std::vector<std::thread> workers;
        for (int g = 0; g < mGpus; g++)
            workers.emplace_back([&, g]() {
                FrameRing& ring = *mRings[g];
                for (size_t i = (size_t)g; i < views.size(); i += (size_t)mGpus) {
                    Renderer& lane = ring.NextLane();
                    lane.SetFrameSink((_byte*)mStore + i * frameBytes, nullptr);
                    ring.Submit(*mMeshes[g], views[i].modelView, views[i].proj, views[i].toRaster);
                }
                ring.Synchronize();
            });
        for (auto& w : workers) w.join();
        for (auto& r : mRings) if (r->LastStatus() != EDX_OK) return false;
        mFrames = views.size();
        return true;
    }
    // frame `view` of the last Render, copied from the root GPU to `out` (w*h*4 bytes, RGBA8 bottom-up)
    bool GetFrame(size_t view, _byte* out) const
    {
        if (view >= mFrames) return false;
        const size_t frameBytes = (size_t)mW * mH * 4;
        return edx_read_device(Root(), out, (const _byte*)mStore + view * frameBytes, frameBytes) == EDX_OK;
    }
    FrameRing& Ring(int gpu) { return *mRings[gpu]; }
    Mesh& MeshOf(int gpu) { return *mMeshes[gpu]; }
private:
    edx_context* Root() const { return mRings[0]->Lane(0).Handle(); }
    void ReleaseStore() { if (mStore) { edx_device_free(Root(), mStore); mStore = nullptr; mStoreFrames = 0; } }
    int mGpus = 0, mStatus = EDX_OK;
    uint mW = 0, mH = 0;
    std::vector<std::unique_ptr<FrameRing>> mRings;
    std::vector<std::unique_ptr<Mesh>> mMeshes;
    void* mStore = nullptr;
    size_t mStoreFrames = 0, mFrames = 0;
};

inline void Mesh::ReleaseDevice() const
{
    // pass the context only if it is still alive; edx_mesh_destroy(NULL, mesh) does a device-wide wait instead
    edx_mesh_destroy(Renderer::LiveContexts().count(mOwner) ? mOwner : nullptr, mDevice);
}

} // namespace edx_b200
