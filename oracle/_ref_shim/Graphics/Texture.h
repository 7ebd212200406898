// EDXUtil stand-in (oracle/_ref_shim): Texture2D<T>, ConstantTexture2D<T>, ImageTexture<TRet, TMem>, TextureFilter.
// EDXUtil's sampler cannot be read, so its behaviour is DEFINED here (DESIGN.md shims 19-24) — the one part of a
// `_ref` frame that is this repository's definition rather than the reference's code: storage and mip chain (19),
// repeat addressing (20), nearest (21), bilinear (22), trilinear LOD from the quad differentials (23), anisotropic
// taps along the major axis (24). Filter ids follow RealtimeViewer/Main.cpp:99-108.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "Color.h"
#include "../Math/Vector.h"
namespace EDX
{
	enum class TextureFilter
	{
		Nearest = 0,
		Linear = 1,
		TriLinear = 2,
		Anisotropic4x = 3,
		Anisotropic8x = 4,
		Anisotropic16x = 5
	};

	// In-memory image handed to ImageTexture through a "mem:<address>" path (the reference only knows file paths,
	// Mesh.cpp:27; there are no asset files offline).
	struct ShimImage
	{
		int width, height;
		const unsigned char* rgba;          // row 0 at v = 0
	};

	template<class T>
	class Texture2D
	{
	public:
		virtual ~Texture2D() {}
		virtual T Sample(const Vector2& texCoord, const Vector2 differentials[2]) const = 0;
		virtual void SetFilter(const TextureFilter filter) {}
	};

	template<class T>
	class ConstantTexture2D : public Texture2D<T>
	{
	private:
		T mVal;
	public:
		ConstantTexture2D(const T& val) : mVal(val) {}
		T Sample(const Vector2& texCoord, const Vector2 differentials[2]) const { return mVal; }
	};

	template<class TRet, class TMem>
	class ImageTexture : public Texture2D<TRet>
	{
	private:
		struct Level { int w, h; std::vector<TMem> texels; };
		std::vector<Level> mLevels;
		TextureFilter mFilter;

	public:
		ImageTexture(const char* path, const float gamma = 1.0f) : mFilter(TextureFilter::TriLinear)
		{
			Level base;
			const ShimImage* img = nullptr;
			if (path && !strncmp(path, "mem:", 4))
				img = (const ShimImage*)(uintptr_t)strtoull(path + 4, nullptr, 16);
			if (img)
			{
				base.w = img->width; base.h = img->height;
				base.texels.resize((size_t)base.w * base.h);
				for (size_t i = 0; i < base.texels.size(); i++)
					base.texels[i] = TMem(img->rgba[4 * i], img->rgba[4 * i + 1], img->rgba[4 * i + 2], img->rgba[4 * i + 3]);
			}
			else
			{
				base.w = base.h = 1;
				base.texels.assign(1, TMem(255, 255, 255, 255));
			}
			mLevels.push_back(base);
			// shim 19: box-filtered chain down to 1x1; odd edges clamp; each level is re-quantised to TMem
			while (mLevels.back().w > 1 || mLevels.back().h > 1)
			{
				const Level& s = mLevels.back();
				Level d;
				d.w = s.w >> 1 > 1 ? s.w >> 1 : 1;
				d.h = s.h >> 1 > 1 ? s.h >> 1 : 1;
				d.texels.resize((size_t)d.w * d.h);
				for (int y = 0; y < d.h; y++)
					for (int x = 0; x < d.w; x++)
					{
						const int x0 = 2 * x < s.w - 1 ? 2 * x : s.w - 1, x1 = 2 * x + 1 < s.w - 1 ? 2 * x + 1 : s.w - 1;
						const int y0 = 2 * y < s.h - 1 ? 2 * y : s.h - 1, y1 = 2 * y + 1 < s.h - 1 ? 2 * y + 1 : s.h - 1;
						const TRet c00(s.texels[(size_t)y0 * s.w + x0]), c10(s.texels[(size_t)y0 * s.w + x1]);
						const TRet c01(s.texels[(size_t)y1 * s.w + x0]), c11(s.texels[(size_t)y1 * s.w + x1]);
						TRet sum = c00;
						sum += c10; sum += c01; sum += c11;
						sum *= 0.25f;
						d.texels[(size_t)y * d.w + x] = TMem(sum);
					}
				mLevels.push_back(d);
			}
		}

		void SetFilter(const TextureFilter filter) { mFilter = filter; }

		TRet Sample(const Vector2& uv, const Vector2 diff[2]) const
		{
			const Level& L0 = mLevels[0];
			switch (mFilter)
			{
			case TextureFilter::Nearest:                                                        // shim 21
				return Texel(L0, (int)floorf(Coord(uv.u, L0.w, 0.0f)), (int)floorf(Coord(uv.v, L0.h, 0.0f)));
			case TextureFilter::Linear:
				return Bilinear(L0, uv.u, uv.v);
			case TextureFilter::TriLinear:                                                      // shim 23
			{
				const float a = fmaxf(fabsf(diff[0].u), fabsf(diff[0].v)), b = fmaxf(fabsf(diff[1].u), fabsf(diff[1].v));
				return Trilinear(uv.u, uv.v, 2.0f * fmaxf(a, b));
			}
			default:                                                                            // shim 24
			{
				const int N = mFilter == TextureFilter::Anisotropic4x ? 4 : (mFilter == TextureFilter::Anisotropic8x ? 8 : 16);
				const float l0 = sqrtf(diff[0].u * diff[0].u + diff[0].v * diff[0].v);
				const float l1 = sqrtf(diff[1].u * diff[1].u + diff[1].v * diff[1].v);
				const bool first = l0 >= l1;
				const float lmaj = first ? l0 : l1, lmin = first ? l1 : l0;
				const float mu = first ? diff[0].u : diff[1].u, mv = first ? diff[0].v : diff[1].v;
				int n = 1;
				if (lmaj > 0.0f)
				{
					if (lmin * (float)N <= lmaj) n = N;
					else { n = (int)ceilf(lmaj / lmin); if (n < 1) n = 1; if (n > N) n = N; }
				}
				if (!(lmaj < 3.0e38f)) n = 1;
				const float width = 2.0f * (lmaj / (float)n);
				float acc[3] = { 0.0f, 0.0f, 0.0f };
				for (int i = 0; i < n; i++)
				{
					const float s = ((float)i + 0.5f) / (float)n - 0.5f;
					const TRet c = Trilinear(uv.u + mu * s, uv.v + mv * s, width);
					acc[0] += c.r; acc[1] += c.g; acc[2] += c.b;
				}
				const float inv = 1.0f / (float)n;
				return TRet(acc[0] * inv, acc[1] * inv, acc[2] * inv);
			}
			}
		}

	private:
		static int Wrap(const int i, const int n) { const int m = i % n; return m < 0 ? m + n : m; }        // shim 20
		static float Coord(const float u, const int n, const float bias)
		{
			const float x = u * (float)n - bias;
			return fabsf(x) < 1.0e9f ? x : 0.0f;
		}
		static TRet Texel(const Level& L, const int x, const int y)
		{
			const TRet c(L.texels[(size_t)Wrap(y, L.h) * L.w + Wrap(x, L.w)]);
			return TRet(c.r, c.g, c.b);
		}
		static TRet Bilinear(const Level& L, const float u, const float v)                                  // shim 22
		{
			const float x = Coord(u, L.w, 0.5f), y = Coord(v, L.h, 0.5f);
			const float x0 = floorf(x), y0 = floorf(y);
			const float fx = x - x0, fy = y - y0;
			const int ix = (int)x0, iy = (int)y0;
			const TRet c00 = Texel(L, ix, iy), c10 = Texel(L, ix + 1, iy), c01 = Texel(L, ix, iy + 1), c11 = Texel(L, ix + 1, iy + 1);
			const float gx = 1.0f - fx, gy = 1.0f - fy;
			const float w00 = gx * gy, w10 = fx * gy, w01 = gx * fy, w11 = fx * fy;
			return TRet(((w00 * c00.r + w10 * c10.r) + w01 * c01.r) + w11 * c11.r,
				((w00 * c00.g + w10 * c10.g) + w01 * c01.g) + w11 * c11.g,
				((w00 * c00.b + w10 * c10.b) + w01 * c01.b) + w11 * c11.b);
		}
		TRet Trilinear(const float u, const float v, const float width) const
		{
			const int L = (int)mLevels.size();
			const float level = (float)(L - 1) + log2f(width > 1.0e-8f ? width : 1.0e-8f);
			if (!(level >= 0.0f)) return Bilinear(mLevels[0], u, v);
			if (level >= (float)(L - 1)) return Texel(mLevels[L - 1], 0, 0);
			const int i = (int)floorf(level);
			const float d = level - (float)i;
			const TRet a = Bilinear(mLevels[i], u, v), b = Bilinear(mLevels[i + 1], u, v);
			return TRet((1.0f - d) * a.r + d * b.r, (1.0f - d) * a.g + d * b.g, (1.0f - d) * a.b + d * b.b);
		}
	};
}
