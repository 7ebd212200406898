// EDXUtil stand-in (oracle/_ref_shim): Vec<N, T> and the Vector* typedefs.
// Arithmetic is component-wise, plain fp32, evaluated left to right as the reference's expressions are written
// (`v0 * (1 - t) + v1 * t`, Clipper.h:211; `w.x * a + w.y * b + w.z * c`, Clipper.h:139-147).
#pragma once
#include "EDXMath.h"
namespace EDX
{
	template<int N, class T> class Vec;

	template<class T>
	class Vec<2, T>
	{
	public:
		union { struct { T x, y; }; struct { T u, v; }; };        // Shader.h:228-229 reads texCoord.u / .v

		Vec() : x(Math::EDX_ZERO), y(Math::EDX_ZERO) {}
		Vec(const Vec& o) : x(o.x), y(o.y) {}
		// Rasterizer.h:247,382 bind `const Vector2i&` to ONE int of FrameBuffer::MultiSampleOffsets, which only compiles
		// through a converting constructor from a scalar: both components take the value (DESIGN.md shim 17).
		Vec(const T& s) : x(s), y(s) {}
		Vec(const T& a, const T& b) : x(a), y(b) {}
		template<class U> Vec(const Vec<2, U>& o) : x(o.x), y(o.y) {}
		Vec& operator=(const Vec& o) { x = o.x; y = o.y; return *this; }

		Vec operator+(const Vec& o) const { return Vec(x + o.x, y + o.y); }
		Vec operator-(const Vec& o) const { return Vec(x - o.x, y - o.y); }
		Vec operator*(const Vec& o) const { return Vec(x * o.x, y * o.y); }
		Vec operator*(const T& s) const { return Vec(x * s, y * s); }
		friend Vec operator*(const T& s, const Vec& a) { return Vec(s * a.x, s * a.y); }
		Vec& operator+=(const Vec& o) { x = x + o.x; y = y + o.y; return *this; }
		Vec& operator*=(const T& s) { x = x * s; y = y * s; return *this; }

		static const Vec ZERO;
	};
	template<class T> const Vec<2, T> Vec<2, T>::ZERO = Vec<2, T>(T(Math::EDX_ZERO), T(Math::EDX_ZERO));

	template<class T>
	class Vec<3, T>
	{
	public:
		T x, y, z;

		Vec() : x(Math::EDX_ZERO), y(Math::EDX_ZERO), z(Math::EDX_ZERO) {}
		Vec(const Vec& o) : x(o.x), y(o.y), z(o.z) {}
		Vec(const T& s) : x(s), y(s), z(s) {}                     // Shader.h:205,281 return a FloatSSE as a Vec3f_SSE
		Vec(const T& a, const T& b, const T& c) : x(a), y(b), z(c) {}
		template<class U> Vec(const Vec<3, U>& o) : x(o.x), y(o.y), z(o.z) {}
		Vec& operator=(const Vec& o) { x = o.x; y = o.y; z = o.z; return *this; }

		Vec operator+(const Vec& o) const { return Vec(x + o.x, y + o.y, z + o.z); }
		Vec operator-(const Vec& o) const { return Vec(x - o.x, y - o.y, z - o.z); }
		Vec operator*(const Vec& o) const { return Vec(x * o.x, y * o.y, z * o.z); }
		Vec operator*(const T& s) const { return Vec(x * s, y * s, z * s); }
		friend Vec operator*(const T& s, const Vec& a) { return Vec(s * a.x, s * a.y, s * a.z); }
		Vec& operator+=(const Vec& o) { x = x + o.x; y = y + o.y; z = z + o.z; return *this; }
		Vec& operator*=(const T& s) { x = x * s; y = y * s; z = z * s; return *this; }

		static const Vec ZERO, UNIT_SCALE, UNIT_X, UNIT_Y, UNIT_Z;
	};
	template<class T> const Vec<3, T> Vec<3, T>::ZERO = Vec<3, T>(T(Math::EDX_ZERO), T(Math::EDX_ZERO), T(Math::EDX_ZERO));
	template<class T> const Vec<3, T> Vec<3, T>::UNIT_SCALE = Vec<3, T>(T(Math::EDX_ONE), T(Math::EDX_ONE), T(Math::EDX_ONE));
	template<class T> const Vec<3, T> Vec<3, T>::UNIT_X = Vec<3, T>(T(Math::EDX_ONE), T(Math::EDX_ZERO), T(Math::EDX_ZERO));
	template<class T> const Vec<3, T> Vec<3, T>::UNIT_Y = Vec<3, T>(T(Math::EDX_ZERO), T(Math::EDX_ONE), T(Math::EDX_ZERO));
	template<class T> const Vec<3, T> Vec<3, T>::UNIT_Z = Vec<3, T>(T(Math::EDX_ZERO), T(Math::EDX_ZERO), T(Math::EDX_ONE));

	template<class T>
	class Vec<4, T>
	{
	public:
		T x, y, z, w;

		Vec() : x(Math::EDX_ZERO), y(Math::EDX_ZERO), z(Math::EDX_ZERO), w(Math::EDX_ZERO) {}
		Vec(const Vec& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
		Vec(const T& a, const T& b, const T& c, const T& d) : x(a), y(b), z(c), w(d) {}
		Vec& operator=(const Vec& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }

		Vec operator+(const Vec& o) const { return Vec(x + o.x, y + o.y, z + o.z, w + o.w); }
		Vec operator-(const Vec& o) const { return Vec(x - o.x, y - o.y, z - o.z, w - o.w); }
		Vec operator*(const T& s) const { return Vec(x * s, y * s, z * s, w * s); }
		friend Vec operator*(const T& s, const Vec& a) { return Vec(s * a.x, s * a.y, s * a.z, s * a.w); }

		// DESIGN.md shim 3: true division by w (Clipper.h:161-163,178-180). The reference passes the result to
		// `RasterTriangle::Setup(Vector3& a, ...)` (RasterTriangle.h:27): MSVC binds the temporary to the non-const
		// reference and Setup overwrites it. g++ needs an lvalue, so the value is returned through a small per-thread
		// ring (three live results per Setup call); what Setup writes into it is discarded, as with MSVC's temporary.
		Vec<3, T>& HomogeneousProject() const
		{
			static thread_local Vec<3, T> ring[4];
			static thread_local unsigned at = 0;
			Vec<3, T>& r = ring[at++ & 3u];
			r = Vec<3, T>(x / w, y / w, z / w);
			return r;
		}

		static const Vec ZERO;
	};
	template<class T> const Vec<4, T> Vec<4, T>::ZERO = Vec<4, T>(T(Math::EDX_ZERO), T(Math::EDX_ZERO), T(Math::EDX_ZERO), T(Math::EDX_ZERO));

	typedef Vec<2, float> Vector2;
	typedef Vec<3, float> Vector3;
	typedef Vec<4, float> Vector4;
	typedef Vec<2, int> Vector2i;
	typedef Vec<3, int> Vector3i;

	namespace Math
	{
		// DESIGN.md shim 10: (x*x + y*y) + z*z, left to right; Normalize divides by the length
		template<class T> inline T Dot(const Vec<3, T>& a, const Vec<3, T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
		inline float Length(const Vector3& v) { return sqrtf(Dot(v, v)); }
		inline Vector3 Normalize(const Vector3& v) { const float len = Length(v); return Vector3(v.x / len, v.y / len, v.z / len); }
	}
}
