// EDXUtil stand-in (oracle/_ref_shim): Color (4 x fp32) and Color4b (RGBA8). DESIGN.md shims 13, 18.
#pragma once
#include "../ShimBase.h"
#include "../Math/EDXMath.h"
namespace EDX
{
	class Color4b;

	class Color
	{
	public:
		float r, g, b, a;
		Color() : r(0.0f), g(0.0f), b(0.0f), a(1.0f) {}
		Color(const float v) : r(v), g(v), b(v), a(v) {}                 // FrameBuffer.cpp:80 `Color c = 0;`
		Color(const float R, const float G, const float B, const float A = 1.0f) : r(R), g(G), b(B), a(A) {}
		Color(const Color4b& c);                                        // shim 18: byte * (1 / 255)
		Color& operator+=(const Color& o) { r = r + o.r; g = g + o.g; b = b + o.b; a = a + o.a; return *this; }
		Color& operator*=(const float s) { r = r * s; g = g * s; b = b * s; a = a * s; return *this; }
		Color operator*(const float s) const { return Color(r * s, g * s, b * s, a * s); }
		friend Color operator*(const float s, const Color& c) { return Color(s * c.r, s * c.g, s * c.b, s * c.a); }
		static const Color WHITE, BLACK;
	};

	class Color4b
	{
	public:
		_byte r, g, b, a;
		Color4b() : r(0), g(0), b(0), a(0) {}
		Color4b(const _byte R, const _byte G, const _byte B, const _byte A = 255) : r(R), g(G), b(B), a(A) {}   // Renderer.cpp:319-321
		// shim 13: clamp to [0, 1], * 255, + 0.5, truncate; NaN -> 0
		static _byte Quantize(const float c)
		{
			const float t = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
			const float s = t * 255.0f + 0.5f;
			if (!(s >= 0.0f)) return 0;
			return (_byte)(int)s;
		}
		Color4b(const Color& c) : r(Quantize(c.r)), g(Quantize(c.g)), b(Quantize(c.b)), a(Quantize(c.a)) {}   // shim 18 (FrameBuffer.cpp:85)
		void FromFloats(const float R, const float G, const float B, const float A = 1.0f)                       // Renderer.cpp:296-299
		{
			r = Quantize(R); g = Quantize(G); b = Quantize(B); a = Quantize(A);
		}
	};

	inline const Color Color::WHITE = Color(1.0f, 1.0f, 1.0f, 1.0f);
	inline const Color Color::BLACK = Color(0.0f, 0.0f, 0.0f, 1.0f);
	inline Color::Color(const Color4b& c) : r(c.r * (1.0f / 255.0f)), g(c.g * (1.0f / 255.0f)), b(c.b * (1.0f / 255.0f)), a(c.a * (1.0f / 255.0f)) {}
}
