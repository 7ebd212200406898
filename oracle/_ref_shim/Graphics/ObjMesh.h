// EDXUtil stand-in (oracle/_ref_shim): ObjMesh, the geometry source behind Mesh::LoadMesh / LoadPlane / LoadSphere
// (Utils/Mesh.cpp:11-70). EDXUtil's OBJ parser and generators are absent; the raster path only needs the arrays they
// would have produced, so LoadFromObj accepts a "mem:<address>" path naming a ShimMeshData (below) — this is how the
// test driver feeds arbitrary vertex / index / material data through the reference's own, unmodified Mesh class.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "../EDXPrerequisites.h"
#include "../Math/BoundingBox.h"
#include "Texture.h"
namespace EDX
{
	struct MeshVertex            // == RasterRenderer::Vertex_PositionNormalTex, 32 bytes (InputBuffer.h:16-28)
	{
		Vector3 position;
		Vector3 normal;
		float fU, fV;
	};

	struct ObjMaterial
	{
		char strName[MAX_PATH];
		char strTexturePath[MAX_PATH];
		Color color;
		ObjMaterial() : color(0.9f, 0.9f, 0.9f) { strName[0] = 0; strTexturePath[0] = 0; }
	};

	// What a "mem:" path points at. Materials: `imagePath[i]` non-empty -> ImageTexture(imagePath[i]) else constant colour.
	struct ShimMeshData
	{
		const float* vertices;       // nVertices x 8 floats: position, normal, texcoord
		unsigned nVertices;
		const unsigned* indices;     // nTriangles x 3
		unsigned nTriangles;
		unsigned nMaterials;
		const float* materialColors; // nMaterials x 3
		const char* const* imagePaths; // nMaterials entries, "" or "mem:<ShimImage address>"
		const unsigned* materialIds; // nTriangles entries, or null for all 0
	};

	class ObjMesh
	{
	private:
		Array<MeshVertex> mVertices;
		Array<uint> mIndices;
		Array<ObjMaterial> mMaterialInfo;
		Array<uint> mMaterialIdx;
		BoundingBox mBounds;

	public:
		bool LoadFromObj(const Vector3& pos, const Vector3& scl, const Vector3& rot, const char* path)
		{
			if (!path || strncmp(path, "mem:", 4) != 0)
				return false;
			const ShimMeshData* d = (const ShimMeshData*)(uintptr_t)strtoull(path + 4, nullptr, 16);
			mVertices.Resize(d->nVertices);
			memcpy((void*)mVertices.Data(), d->vertices, (size_t)d->nVertices * 32);
			mIndices.Resize((size_t)d->nTriangles * 3);
			memcpy(mIndices.Data(), d->indices, (size_t)d->nTriangles * 12);
			for (unsigned i = 0; i < d->nMaterials; i++)
			{
				ObjMaterial m;
				m.color = Color(d->materialColors[3 * i], d->materialColors[3 * i + 1], d->materialColors[3 * i + 2]);
				if (d->imagePaths && d->imagePaths[i])
					strncpy(m.strTexturePath, d->imagePaths[i], MAX_PATH - 1), m.strTexturePath[MAX_PATH - 1] = 0;
				mMaterialInfo.Add(m);
			}
			mMaterialIdx.Resize(d->nTriangles);
			for (unsigned i = 0; i < d->nTriangles; i++)
				mMaterialIdx[i] = d->materialIds ? d->materialIds[i] : 0u;
			for (size_t i = 0; i < mVertices.Size(); i++)
				mBounds.Grow(mVertices[i].position);
			return true;
		}

		// Plane of side `length` in the xz plane, +y normal, two triangles (stand-in generator; EDXUtil's is absent)
		void LoadPlane(const Vector3& pos, const Vector3& scl, const Vector3& rot, const float length)
		{
			const float h = length * 0.5f;
			const float p[4][2] = { { -h, -h }, { -h, h }, { h, h }, { h, -h } };
			for (int i = 0; i < 4; i++)
			{
				MeshVertex v;
				v.position = Vector3(p[i][0] * scl.x + pos.x, pos.y, p[i][1] * scl.z + pos.z);
				v.normal = Vector3(0.0f, 1.0f, 0.0f);
				v.fU = p[i][0] / length + 0.5f; v.fV = p[i][1] / length + 0.5f;
				mVertices.Add(v);
				mBounds.Grow(v.position);
			}
			const uint idx[6] = { 0, 1, 2, 0, 2, 3 };
			for (int i = 0; i < 6; i++) mIndices.Add(idx[i]);
			mMaterialIdx.Resize(2);
			mMaterialIdx[0] = mMaterialIdx[1] = 0;
		}

		// UV sphere, (slices + 1) x (stacks + 1) vertices, 2 x slices x stacks triangles (stand-in generator)
		void LoadSphere(const Vector3& pos, const Vector3& scl, const Vector3& rot, const float radius, const int slices = 64, const int stacks = 64)
		{
			for (int j = 0; j <= stacks; j++)
				for (int i = 0; i <= slices; i++)
				{
					const float theta = Math::EDX_PI * (float)j / (float)stacks, phi = 2.0f * Math::EDX_PI * (float)i / (float)slices;
					const Vector3 n(sinf(theta) * cosf(phi), cosf(theta), sinf(theta) * sinf(phi));
					MeshVertex v;
					v.position = Vector3(radius * n.x * scl.x + pos.x, radius * n.y * scl.y + pos.y, radius * n.z * scl.z + pos.z);
					v.normal = n;
					v.fU = (float)i / (float)slices; v.fV = (float)j / (float)stacks;
					mVertices.Add(v);
					mBounds.Grow(v.position);
				}
			for (int j = 0; j < stacks; j++)
				for (int i = 0; i < slices; i++)
				{
					const uint a = j * (slices + 1) + i, b = a + slices + 1;
					mIndices.Add(a); mIndices.Add(a + 1); mIndices.Add(b);
					mIndices.Add(a + 1); mIndices.Add(b + 1); mIndices.Add(b);
				}
			mMaterialIdx.Resize(mIndices.Size() / 3);
			for (size_t i = 0; i < mMaterialIdx.Size(); i++) mMaterialIdx[i] = 0;
		}

		const MeshVertex& GetVertexAt(const size_t i) const { return mVertices[i]; }
		const uint* GetIndexAt(const size_t i) const { return &mIndices[3 * i]; }
		uint GetVertexCount() const { return (uint)mVertices.Size(); }
		uint GetTriangleCount() const { return (uint)(mIndices.Size() / 3); }
		const Array<ObjMaterial>& GetMaterialInfo() const { return mMaterialInfo; }
		const Array<uint>& GetMaterialIdxBuf() const { return mMaterialIdx; }
		const BoundingBox& GetBounds() const { return mBounds; }
	};
}
