// EDXUtil stand-in (oracle/_ref_shim): BoundingBox (Mesh.h:26,61-64; Main.cpp:55-57). Not on the raster path.
#pragma once
#include "Vector.h"
namespace EDX
{
	class BoundingBox
	{
	public:
		Vector3 mMin, mMax;
		BoundingBox() : mMin(1e30f, 1e30f, 1e30f), mMax(-1e30f, -1e30f, -1e30f) {}
		void Grow(const Vector3& p)
		{
			mMin = Vector3(Math::Min(mMin.x, p.x), Math::Min(mMin.y, p.y), Math::Min(mMin.z, p.z));
			mMax = Vector3(Math::Max(mMax.x, p.x), Math::Max(mMax.y, p.y), Math::Max(mMax.z, p.z));
		}
		void BoundingSphere(Vector3* pCenter, float* pRadius) const
		{
			*pCenter = (mMin + mMax) * 0.5f;
			*pRadius = Math::Length(mMax - *pCenter);
		}
	};
}
