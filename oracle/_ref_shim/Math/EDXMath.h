// EDXUtil stand-in (oracle/_ref_shim): Math:: scalars and constants.
// EDXUtil's SIMD wrappers descend from Embree's ssef/ssei/sseb, whose `zero`, `one`, `pos_inf` tag objects convert to
// every arithmetic type (pos_inf -> INT_MAX for int, +inf for float); EDX_ZERO / EDX_ONE / EDX_INFINITY are used the
// same way (Rasterizer.h:56 `IntSSE acptEdgeFunc0 = Math::EDX_INFINITY;`, Shader.h:151 `FloatSSE(Math::EDX_ONE)`).
#pragma once
#include "../ShimBase.h"
#include <climits>
#include <cmath>
#include <type_traits>
namespace EDX
{
	namespace Math
	{
		struct ZeroTy { operator float() const { return 0.0f; } operator double() const { return 0.0; } operator int() const { return 0; } operator unsigned() const { return 0u; } };
		struct OneTy { operator float() const { return 1.0f; } operator double() const { return 1.0; } operator int() const { return 1; } operator unsigned() const { return 1u; } };
		struct PosInfTy { operator float() const { return INFINITY; } operator double() const { return (double)INFINITY; } operator int() const { return INT_MAX; } };   // DESIGN.md shim 12
		static const ZeroTy EDX_ZERO = ZeroTy();
		static const OneTy EDX_ONE = OneTy();
		static const PosInfTy EDX_INFINITY = PosInfTy();
		static const float EDX_PI = 3.14159265358979323846f;
		static const float EDX_INV_PI = 0.31830988618f;                  // DESIGN.md shim 11 (Shader.h:264)

		// Renderer.cpp:48 mixes int and uint arguments, so two type parameters; arguments by value (Tile::SIZE is a
		// static const int without an out-of-class definition and must not be odr-used).
		template<class A, class B> inline typename std::common_type<A, B>::type Min(A a, B b) { typedef typename std::common_type<A, B>::type R; return (R)a < (R)b ? (R)a : (R)b; }
		template<class A, class B> inline typename std::common_type<A, B>::type Max(A a, B b) { typedef typename std::common_type<A, B>::type R; return (R)a > (R)b ? (R)a : (R)b; }
		inline int Abs(int a) { return a < 0 ? -a : a; }
		inline float Abs(float a) { return fabsf(a); }
		inline float Pow(float a, float b) { return powf(a, b); }          // DESIGN.md shim 10 (Shader.h:275-278)
		inline float Sqrt(float a) { return sqrtf(a); }
	}
	namespace Constants
	{
		struct TrueTy { operator bool() const { return true; } };
		struct FalseTy { operator bool() const { return false; } };
		static const TrueTy EDX_TRUE = TrueTy();
		static const FalseTy EDX_FALSE = FalseTy();
	}
}
