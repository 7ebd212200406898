// ref_driver.cpp — TEST INFRASTRUCTURE ONLY. C entry points around the reference's own, unmodified
// EDX::RasterRenderer::Renderer / Mesh (compiled from /root/reference/EDXRaster where the sources lie, against the
// EDXUtil stand-in in this directory) so that tests/ and bench.py's reference leg can drive it through ctypes.
// Nothing here re-implements a raster stage: every pixel, depth and triangle record it returns was produced by the
// reference's code. Private members are read (never written, except the pixel-shader slot, which the reference
// hard-codes at Renderer.cpp:41 and offers no setter for) with g++ -fno-access-control.
#include "Core/Renderer.h"
#include "Core/FrameBuffer.h"
#include "Core/Rasterizer.h"
#include "Core/Scene.h"
#include "Utils/Mesh.h"
#include "Utils/InputBuffer.h"
#include "Windows/Threading.h"

#include <chrono>
#include <string>
#include <vector>

using namespace EDX;
using namespace EDX::RasterRenderer;

namespace
{
	struct Ref
	{
		Renderer* renderer = nullptr;
		Mesh* mesh = nullptr;
		int w = 0, h = 0;
		int shader = 3;                       // 1 Blinn-Phong, 3 LambertianAlbedo (the reference's default)
		// backing store of the "mem:" blobs the shim's ObjMesh / ImageTexture read during LoadMesh
		std::vector<ShimImage> images;
		std::vector<std::vector<unsigned char>> imageData;
		double lastMs = 0.0;
	};
	Ref* gLive = nullptr;                     // RenderStates is a process-wide singleton (RenderStates.h:35-44): one at a time
}

extern "C" {

void* ref_create(int width, int height, int threads)
{
	if (gLive) return nullptr;
	ShimNumCoresOverride() = threads;
	if (threads > 0) omp_set_num_threads(threads);
	Ref* r = new Ref;
	r->w = width; r->h = height;
	r->renderer = new Renderer;
	r->renderer->Initialize((uint)width, (uint)height);          // Renderer.cpp:22-62
	gLive = r;
	return r;
}

void ref_destroy(void* h)
{
	Ref* r = (Ref*)h;
	if (!r) return;
	delete r->mesh;
	delete r->renderer;
	if (gLive == r) gLive = nullptr;
	delete r;
}

int ref_threads(void* h) { return ((Ref*)h)->renderer->mNumCores; }

void ref_resize(void* h, int width, int height)
{
	Ref* r = (Ref*)h;
	r->w = width; r->h = height;
	r->renderer->Resize((uint)width, (uint)height);              // Renderer.cpp:64-83
}

void ref_set_transform(void* h, const float* mv, const float* proj, const float* raster)
{
	((Ref*)h)->renderer->SetTransform(Matrix(mv), Matrix(proj), Matrix(raster));      // Renderer.cpp:85-92
}

// ModelViewProj and the eye position exactly as FragmentProcessing derives it (Renderer.cpp:289)
void ref_get_derived(void* h, float* mvp16, float* eye3)
{
	memcpy(mvp16, RenderStates::Instance()->GetModelViewProjMatrix().m, 64);
	const Vector3 e = Matrix::TransformPoint(Vector3::ZERO, RenderStates::Instance()->GetModelViewInvMatrix());
	eye3[0] = e.x; eye3[1] = e.y; eye3[2] = e.z;
}

int ref_set_shader(void* h, int mode)
{
	Ref* r = (Ref*)h;
	if (mode == 1) r->renderer->mpPixelShader = MakeUnique<BlinnPhongPixelShader>();            // Shader.h:246-282
	else if (mode == 3) r->renderer->mpPixelShader = MakeUnique<LambertianAlbedoPixelShader>(); // Shader.h:209-244
	else return -1;
	r->shader = mode;
	return 0;
}

void ref_set_msaa(void* h, int log2) { ((Ref*)h)->renderer->SetMSAAMode(log2); }                       // Renderer.cpp:94-98
void ref_set_hierarchical(void* h, int on) { ((Ref*)h)->renderer->SetHierarchicalRasterize(on != 0); }  // Renderer.h:49
void ref_set_texture_filter(void* h, int f) { ((Ref*)h)->renderer->SetTextureFilter(TextureFilter(f)); } // Renderer.h:48
int ref_samples(void* h) { return (int)((Ref*)h)->renderer->mpFrameBuffer->GetSampleCount(); }

// Builds the reference Mesh through Mesh::LoadMesh (Utils/Mesh.cpp:11-34). kinds[i]: 0 constant colour (colors[3i..]),
// 1 RGBA8 image (images[i], dims[2i] x dims[2i+1]); texIds: one slot per triangle or null.
void ref_set_mesh(void* h, const float* vtx, unsigned nv, const unsigned* idx, unsigned nt,
                  unsigned nMat, const int* kinds, const float* colors, const unsigned char* const* images, const int* dims,
                  const unsigned* texIds)
{
	Ref* r = (Ref*)h;
	delete r->mesh;
	r->mesh = new Mesh;
	r->images.assign(nMat, ShimImage());
	r->imageData.assign(nMat, std::vector<unsigned char>());
	std::vector<std::string> paths(nMat);
	std::vector<const char*> pathPtrs(nMat);
	for (unsigned i = 0; i < nMat; i++)
	{
		if (kinds[i] == 1)
		{
			r->imageData[i].assign(images[i], images[i] + (size_t)dims[2 * i] * dims[2 * i + 1] * 4);
			r->images[i].width = dims[2 * i]; r->images[i].height = dims[2 * i + 1];
			r->images[i].rgba = r->imageData[i].data();
			char buf[64];
			snprintf(buf, sizeof(buf), "mem:%llx", (unsigned long long)(uintptr_t)&r->images[i]);
			paths[i] = buf;
		}
		pathPtrs[i] = paths[i].c_str();
	}
	ShimMeshData d;
	d.vertices = vtx; d.nVertices = nv; d.indices = idx; d.nTriangles = nt;
	d.nMaterials = nMat; d.materialColors = colors; d.imagePaths = pathPtrs.data(); d.materialIds = texIds;
	char path[64];
	snprintf(path, sizeof(path), "mem:%llx", (unsigned long long)(uintptr_t)&d);
	r->mesh->LoadMesh(Vector3::ZERO, Vector3::UNIT_SCALE, Vector3::ZERO, path);
}

// Renderer::RenderMesh, Renderer.cpp:100-118; returns the wall time of the call in ms
double ref_render(void* h)
{
	Ref* r = (Ref*)h;
	const auto t0 = std::chrono::steady_clock::now();
	r->renderer->RenderMesh(*r->mesh);
	r->lastMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	return r->lastMs;
}

// Renderer::GetBackBuffer, Renderer.cpp:360-363: W x H RGBA8, bottom-up
const unsigned char* ref_color(void* h) { return ((Ref*)h)->renderer->GetBackBuffer(); }

// colour of one sample plane (FrameBuffer::mColorBufferMS, [sample][x][row]), bottom-up
void ref_get_color_sample(void* h, int sId, unsigned char* out)
{
	Ref* r = (Ref*)h;
	FrameBuffer* fb = r->renderer->mpFrameBuffer.Get();
	for (int row = 0; row < r->h; row++)
		for (int x = 0; x < r->w; x++)
			memcpy(out + ((size_t)row * r->w + x) * 4, &fb->mColorBufferMS[Vector3i(sId, x, row)], 4);
}

// depth of one sample, read from the tiled quad-layout depth buffer the way ZTestQuad addresses it
// (FrameBuffer.cpp:54-68), linearised bottom-up like the colour buffer
void ref_get_depth_sample(void* h, int sId, float* out)
{
	Ref* r = (Ref*)h;
	FrameBuffer* fb = r->renderer->mpFrameBuffer.Get();
	for (int y = 0; y < r->h; y++)
		for (int x = 0; x < r->w; x++)
		{
			const int tileX = x >> Tile::SIZE_LOG_2, tileY = y >> Tile::SIZE_LOG_2;
			const int qx = x & ~1, qy = y & ~1;
			const int ix = qx & (Tile::SIZE - 1), iy = qy & (Tile::SIZE - 1);
			const FloatSSE& q = fb->mTiledDepthBuffer[tileY * fb->mTileDimX + tileX][Vector3i(sId, ix >> 1, (Tile::SIZE - 1 - iy) >> 1)];
			out[(size_t)(r->h - 1 - y) * r->w + x] = q[(x & 1) + 2 * (y & 1)];
		}
}

// clip-space position of every submitted vertex (mProjectedVertexBuf, Renderer.cpp:120-127)
void ref_get_clip_verts(void* h, float* out)
{
	Renderer* R = ((Ref*)h)->renderer;
	for (size_t i = 0; i < R->mProjectedVertexBuf.Size(); i++)
		memcpy(out + 4 * i, &R->mProjectedVertexBuf[i].projectedPos, 16);
}

unsigned long long ref_num_raster_tris(void* h)
{
	Renderer* R = ((Ref*)h)->renderer;
	unsigned long long n = 0;
	for (int c = 0; c < R->mNumCores; c++) n += R->mpRasterTriangleBuf[c].Size();
	return n;
}

// every set-up triangle in submission order (core chunks ascending, Clipper.h:80-82):
// ints[6] = v0x, v0y, v1x, v1y, v2x, v2y (28.4); floats[7] = z0, z1, z2, invW0, invW1, invW2, invDet
void ref_get_raster_tris(void* h, int* ints, float* floats)
{
	Renderer* R = ((Ref*)h)->renderer;
	size_t at = 0;
	for (int c = 0; c < R->mNumCores; c++)
		for (size_t i = 0; i < R->mpRasterTriangleBuf[c].Size(); i++, at++)
		{
			const RasterTriangle& t = R->mpRasterTriangleBuf[c][i];
			int* I = ints + 6 * at; float* F = floats + 7 * at;
			I[0] = t.v0.x; I[1] = t.v0.y; I[2] = t.v1.x; I[3] = t.v1.y; I[4] = t.v2.x; I[5] = t.v2.y;
			const ProjectedVertex* vb = R->mpDistributedProjVertexBuf[c].Data();
			F[0] = vb[t.vId0].projectedPos.z; F[1] = vb[t.vId1].projectedPos.z; F[2] = vb[t.vId2].projectedPos.z;
			F[3] = vb[t.vId0].invW; F[4] = vb[t.vId1].invW; F[5] = vb[t.vId2].invW;
			F[6] = t.invDet;
		}
}

// per pixel of one sample: ordinal (position in ref_get_raster_tris order) of the triangle whose fragment wrote the
// pixel last (UpdateFrameBuffer's order, Renderer.cpp:305-345), 0xFFFFFFFF where nothing was written. Bottom-up.
void ref_get_winner_sample(void* h, int sId, unsigned* out)
{
	Ref* r = (Ref*)h;
	Renderer* R = r->renderer;
	// (core, vId2) identifies a set-up triangle: vertex ids are per-core append positions and every fan triangle of a
	// polygon ends in a different vertex (Clipper.h:156-170)
	std::vector<std::vector<unsigned>> ordinalOfV2(R->mNumCores);
	unsigned at = 0;
	for (int c = 0; c < R->mNumCores; c++)
	{
		ordinalOfV2[c].assign(R->mpDistributedProjVertexBuf[c].Size(), 0xFFFFFFFFu);
		for (size_t i = 0; i < R->mpRasterTriangleBuf[c].Size(); i++, at++)
			ordinalOfV2[c][R->mpRasterTriangleBuf[c][i].vId2] = at;
	}
	for (size_t i = 0; i < (size_t)r->w * r->h; i++) out[i] = 0xFFFFFFFFu;
	for (size_t i = 0; i < R->mTiles.Size(); i++)
		for (size_t j = 0; j < R->mTiles[i].fragmentBuf.Size(); j++)
		{
			const Fragment& f = R->mTiles[i].fragmentBuf[j];
			const int shift = sId << 2;
			for (int k = 0; k < 4; k++)
			{
				if (!f.coverageMask.GetBit(shift + k)) continue;
				const int x = f.x + (k & 1), y = f.y + (k >> 1);
				if (x >= r->w || y >= r->h) continue;
				out[(size_t)(r->h - 1 - y) * r->w + x] = ordinalOfV2[f.coreId][f.vId2];
			}
		}
}

unsigned long long ref_num_fragments(void* h) { return ((Ref*)h)->renderer->mFragmentBuf.Size(); }

} // extern "C"
