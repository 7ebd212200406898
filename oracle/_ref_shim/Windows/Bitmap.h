// EDXUtil stand-in (oracle/_ref_shim): Bitmap::SaveBitmapFile (Renderer.cpp:357) — 32-bit uncompressed BMP, bottom-up
#pragma once
#include <cstdio>
#include <cstdint>
#include <vector>
namespace EDX
{
	class Bitmap
	{
	public:
		static bool SaveBitmapFile(const char* path, const unsigned char* rgba, const int w, const int h)
		{
			FILE* f = fopen(path, "wb");
			if (!f) return false;
			const uint32_t bytes = (uint32_t)w * h * 4, off = 54, size = off + bytes;
			unsigned char hdr[54] = { 'B', 'M' };
			auto put = [&](int at, uint32_t v) { hdr[at] = v & 255; hdr[at + 1] = (v >> 8) & 255; hdr[at + 2] = (v >> 16) & 255; hdr[at + 3] = (v >> 24) & 255; };
			put(2, size); put(10, off); put(14, 40); put(18, (uint32_t)w); put(22, (uint32_t)h);
			hdr[26] = 1; hdr[28] = 32; put(34, bytes);
			fwrite(hdr, 1, 54, f);
			std::vector<unsigned char> row((size_t)w * 4);
			for (int y = 0; y < h; y++)
			{
				for (int x = 0; x < w; x++)
				{
					const unsigned char* p = rgba + ((size_t)y * w + x) * 4;
					row[4 * x] = p[2]; row[4 * x + 1] = p[1]; row[4 * x + 2] = p[0]; row[4 * x + 3] = p[3];
				}
				fwrite(row.data(), 1, row.size(), f);
			}
			fclose(f);
			return true;
		}
	};
}
