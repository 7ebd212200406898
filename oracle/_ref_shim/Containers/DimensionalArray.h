// EDXUtil stand-in (oracle/_ref_shim): DimensionalArray<N, T>. DESIGN.md shim 14: the FIRST index varies fastest, so
// at one sample per pixel FrameBuffer's [sample][x][row] colour array is the row-major bottom-up RGBA image that
// RealtimeViewer/Main.cpp:75 hands to glDrawPixels.
#pragma once
#include "../ShimBase.h"
#include <cstdlib>
#include <cstring>
#include <new>
#include "../Math/Vector.h"
namespace EDX
{
	template<int N, class T>
	class DimensionalArray
	{
	private:
		int mDim[N];
		size_t mSize;
		T* mpData;

	public:
		DimensionalArray() : mSize(0), mpData(nullptr) { for (int i = 0; i < N; i++) mDim[i] = 0; }
		DimensionalArray(const DimensionalArray& o) : mSize(0), mpData(nullptr) { *this = o; }
		DimensionalArray& operator=(const DimensionalArray& o)
		{
			if (this == &o) return *this;
			Free();
			for (int i = 0; i < N; i++) mDim[i] = o.mDim[i];
			mSize = o.mSize;
			if (mSize) { Alloc(); memcpy((void*)mpData, (const void*)o.mpData, mSize * sizeof(T)); }
			return *this;
		}
		~DimensionalArray() { Free(); }

		void Init(const Vec<N, int>& size)
		{
			Free();
			Dims(size, mDim);
			mSize = 1;
			for (int i = 0; i < N; i++) mSize *= (size_t)mDim[i];
			Alloc();
			Clear();
		}
		void Free() { if (mpData) { free(mpData); mpData = nullptr; } mSize = 0; }
		void Clear() { if (mSize) memset((void*)mpData, 0, mSize * sizeof(T)); }                 // FrameBuffer.cpp:93-94
		size_t LinearSize() const { return mSize; }
		size_t LinearIndex(const Vec<N, int>& idx) const
		{
			int c[N];
			Dims(idx, c);
			size_t at = 0;
			for (int i = N - 1; i >= 0; i--) at = at * (size_t)mDim[i] + (size_t)c[i];
			return at;
		}
		Vec<N, int> Index(size_t linear) const                                                    // FrameBuffer.cpp:78
		{
			int c[N];
			for (int i = 0; i < N; i++) { c[i] = (int)(linear % (size_t)mDim[i]); linear /= (size_t)mDim[i]; }
			return Make(c, (const Vec<N, int>*)nullptr);
		}
		T& operator[](const Vec<N, int>& idx) { return mpData[LinearIndex(idx)]; }
		const T& operator[](const Vec<N, int>& idx) const { return mpData[LinearIndex(idx)]; }
		T& operator[](const size_t i) { return mpData[i]; }
		const T& operator[](const size_t i) const { return mpData[i]; }
		T* Data() { return mpData; }
		const T* Data() const { return mpData; }

	private:
		void Alloc() { mpData = (T*)aligned_alloc(64, ((mSize * sizeof(T) + 63) / 64) * 64 + 64); }
		static void Dims(const Vec<2, int>& v, int* d) { d[0] = v.x; d[1] = v.y; }
		static void Dims(const Vec<3, int>& v, int* d) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
		static Vec<2, int> Make(const int* c, const Vec<2, int>*) { return Vec<2, int>(c[0], c[1]); }
		static Vec<3, int> Make(const int* c, const Vec<3, int>*) { return Vec<3, int>(c[0], c[1], c[2]); }
	};
}
