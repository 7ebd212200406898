// EDXUtil stand-in (oracle/_ref_shim): FloatSSE / IntSSE / BoolSSE and the Vec*_SSE typedefs.
// EDXUtil's wrappers follow Embree's ssef / ssei / sseb: lane 0 is the first constructor argument, integer
// arithmetic wraps mod 2^32 (_mm_mullo_epi32), comparisons yield all-ones lanes, int -> float is cvtepi32_ps.
// DESIGN.md shims 5-9.
#pragma once
#include "../ShimBase.h"
#include <smmintrin.h>
#include "../Math/EDXMath.h"
#include "../Math/Vector.h"
namespace EDX
{
	class FloatSSE;
	class IntSSE;

	class BoolSSE
	{
	public:
		__m128 m128;
		BoolSSE() = default;
		BoolSSE(const __m128 a) : m128(a) {}
		BoolSSE(const __m128i a) : m128(_mm_castsi128_ps(a)) {}
		BoolSSE(Constants::TrueTy) : m128(_mm_castsi128_ps(_mm_set1_epi32(-1))) {}
		BoolSSE(Constants::FalseTy) : m128(_mm_setzero_ps()) {}
		operator __m128() const { return m128; }
		int operator[](const int i) const { return (_mm_movemask_ps(m128) >> i) & 1; }     // Shader.h:76-91, Rasterizer.h:89,97
	};
	__forceinline BoolSSE operator&(const BoolSSE& a, const BoolSSE& b) { return _mm_and_ps(a.m128, b.m128); }
	__forceinline BoolSSE operator|(const BoolSSE& a, const BoolSSE& b) { return _mm_or_ps(a.m128, b.m128); }
	__forceinline BoolSSE operator^(const BoolSSE& a, const BoolSSE& b) { return _mm_xor_ps(a.m128, b.m128); }
	__forceinline BoolSSE operator!(const BoolSSE& a) { return _mm_xor_ps(a.m128, BoolSSE(Constants::EDX_TRUE).m128); }

	class IntSSE
	{
	public:
		// Renderer.cpp:319-342 reads `m128.m128i_u8[k]`: MSVC's __m128i is a union with that member
		union M128
		{
			__m128i v;
			unsigned char m128i_u8[16];
			int m128i_i32[4];
			unsigned int m128i_u32[4];
			M128() = default;
			M128(const __m128i a) : v(a) {}
			operator __m128i() const { return v; }
		} m128;

		IntSSE() = default;
		IntSSE(const __m128i a) : m128(a) {}
		IntSSE(const int a) : m128(_mm_set1_epi32(a)) {}
		IntSSE(const int a, const int b, const int c, const int d) : m128(_mm_set_epi32(d, c, b, a)) {}   // shim 6: lane 0 = a
		IntSSE(Math::ZeroTy) : m128(_mm_setzero_si128()) {}
		IntSSE(Math::OneTy) : m128(_mm_set1_epi32(1)) {}
		IntSSE(Math::PosInfTy) : m128(_mm_set1_epi32(INT_MAX)) {}                                          // shim 12
		// shim 5: a comparison mask becomes an integer by BIT-CAST (true = -1), as Embree's ssei(const sseb&);
		// RasterTriangle.h:296-299 returns one from TriangleSSE::TopLeftEdge
		IntSSE(const BoolSSE& a) : m128(_mm_castps_si128(a.m128)) {}
		explicit IntSSE(const FloatSSE& a);
		operator __m128i() const { return m128.v; }
		int operator[](const int i) const { return m128.m128i_i32[i]; }
		int& operator[](const int i) { return m128.m128i_i32[i]; }
	};
	__forceinline IntSSE operator+(const IntSSE& a, const IntSSE& b) { return _mm_add_epi32(a, b); }
	__forceinline IntSSE operator-(const IntSSE& a, const IntSSE& b) { return _mm_sub_epi32(a, b); }
	__forceinline IntSSE operator*(const IntSSE& a, const IntSSE& b) { return _mm_mullo_epi32(a, b); }         // shim 6
	// scalar-on-one-side forms the reference writes (`stepSize * IntSSE(...)`, RasterTriangle.h:199; `offset.x * B0`, Rasterizer.h:248)
	__forceinline IntSSE operator*(const int a, const IntSSE& b) { return _mm_mullo_epi32(_mm_set1_epi32(a), b); }
	__forceinline IntSSE operator*(const IntSSE& a, const int b) { return _mm_mullo_epi32(a, _mm_set1_epi32(b)); }
	__forceinline IntSSE operator+(const IntSSE& a, const int b) { return _mm_add_epi32(a, _mm_set1_epi32(b)); }
	__forceinline IntSSE operator-(const IntSSE& a, const int b) { return _mm_sub_epi32(a, _mm_set1_epi32(b)); }
	__forceinline IntSSE operator-(const IntSSE& a) { return _mm_sub_epi32(_mm_setzero_si128(), a); }
	__forceinline IntSSE operator&(const IntSSE& a, const IntSSE& b) { return _mm_and_si128(a, b); }
	__forceinline IntSSE operator|(const IntSSE& a, const IntSSE& b) { return _mm_or_si128(a, b); }
	__forceinline IntSSE& operator+=(IntSSE& a, const IntSSE& b) { return a = a + b; }
	__forceinline IntSSE& operator-=(IntSSE& a, const IntSSE& b) { return a = a - b; }
	__forceinline BoolSSE operator==(const IntSSE& a, const IntSSE& b) { return _mm_cmpeq_epi32(a, b); }
	__forceinline BoolSSE operator<(const IntSSE& a, const IntSSE& b) { return _mm_cmplt_epi32(a, b); }
	__forceinline BoolSSE operator>(const IntSSE& a, const IntSSE& b) { return _mm_cmpgt_epi32(a, b); }
	__forceinline BoolSSE operator>=(const IntSSE& a, const IntSSE& b) { return !(a < b); }
	__forceinline BoolSSE operator<=(const IntSSE& a, const IntSSE& b) { return !(a > b); }

	class FloatSSE
	{
	public:
		__m128 m128;
		FloatSSE() = default;
		FloatSSE(const __m128 a) : m128(a) {}
		FloatSSE(const float a) : m128(_mm_set1_ps(a)) {}
		FloatSSE(const float a, const float b, const float c, const float d) : m128(_mm_set_ps(d, c, b, a)) {}
		FloatSSE(Math::ZeroTy) : m128(_mm_setzero_ps()) {}
		FloatSSE(Math::OneTy) : m128(_mm_set1_ps(1.0f)) {}
		FloatSSE(const IntSSE& a) : m128(_mm_cvtepi32_ps(a)) {}                                            // shim 7
		operator __m128() const { return m128; }
		const float& operator[](const int i) const { return ((const float*)&m128)[i]; }
		float& operator[](const int i) { return ((float*)&m128)[i]; }
	};
	inline IntSSE::IntSSE(const FloatSSE& a) : m128(_mm_cvtps_epi32(a.m128)) {}       // Rasterizer.h:327,374: exact multiples of 16

	__forceinline FloatSSE operator+(const FloatSSE& a, const FloatSSE& b) { return _mm_add_ps(a, b); }
	__forceinline FloatSSE operator-(const FloatSSE& a, const FloatSSE& b) { return _mm_sub_ps(a, b); }
	__forceinline FloatSSE operator*(const FloatSSE& a, const FloatSSE& b) { return _mm_mul_ps(a, b); }
	__forceinline FloatSSE operator/(const FloatSSE& a, const FloatSSE& b) { return _mm_div_ps(a, b); }         // shim 8
	__forceinline FloatSSE operator+(const FloatSSE& a, const float b) { return _mm_add_ps(a, _mm_set1_ps(b)); }
	__forceinline FloatSSE operator-(const FloatSSE& a, const float b) { return _mm_sub_ps(a, _mm_set1_ps(b)); }
	__forceinline FloatSSE operator*(const FloatSSE& a, const float b) { return _mm_mul_ps(a, _mm_set1_ps(b)); }
	__forceinline FloatSSE operator*(const float a, const FloatSSE& b) { return _mm_mul_ps(_mm_set1_ps(a), b); }
	__forceinline FloatSSE& operator+=(FloatSSE& a, const FloatSSE& b) { return a = a + b; }
	__forceinline FloatSSE& operator*=(FloatSSE& a, const FloatSSE& b) { return a = a * b; }
	__forceinline BoolSSE operator<(const FloatSSE& a, const FloatSSE& b) { return _mm_cmplt_ps(a, b); }
	__forceinline BoolSSE operator<=(const FloatSSE& a, const FloatSSE& b) { return _mm_cmple_ps(a, b); }
	__forceinline BoolSSE operator>(const FloatSSE& a, const FloatSSE& b) { return _mm_cmpgt_ps(a, b); }
	__forceinline BoolSSE operator>=(const FloatSSE& a, const FloatSSE& b) { return _mm_cmpge_ps(a, b); }

	namespace SSE
	{
		__forceinline bool Any(const BoolSSE& a) { return _mm_movemask_ps(a.m128) != 0; }
		__forceinline bool All(const BoolSSE& a) { return _mm_movemask_ps(a.m128) == 15; }
		// Select(mask, a, b) = mask ? a : b per lane (FrameBuffer.cpp:65, Shader.h:262)
		__forceinline FloatSSE Select(const BoolSSE& m, const FloatSSE& a, const FloatSSE& b) { return _mm_blendv_ps(b, a, m); }
		__forceinline IntSSE Select(const BoolSSE& m, const IntSSE& a, const IntSSE& b) { return _mm_castps_si128(_mm_blendv_ps(_mm_castsi128_ps(b), _mm_castsi128_ps(a), m)); }
		// shim 9: parity build = exact 1 / sqrt(x); -DEDX_SHIM_FAST_RSQRT = rsqrtps + one Newton step (Embree's rsqrt)
		__forceinline FloatSSE Rsqrt(const FloatSSE& a)
		{
#ifdef EDX_SHIM_FAST_RSQRT
			const __m128 r = _mm_rsqrt_ps(a);
			const __m128 h = _mm_mul_ps(_mm_set1_ps(0.5f), a);
			return _mm_mul_ps(r, _mm_sub_ps(_mm_set1_ps(1.5f), _mm_mul_ps(h, _mm_mul_ps(r, r))));
#else
			return _mm_div_ps(_mm_set1_ps(1.0f), _mm_sqrt_ps(a));
#endif
		}
	}

	typedef Vec<2, FloatSSE> Vec2f_SSE;
	typedef Vec<3, FloatSSE> Vec3f_SSE;
	typedef Vec<2, IntSSE> Vec2i_SSE;
	typedef Vec<3, IntSSE> Vec3i_SSE;
}
