// EDXUtil stand-in (oracle/_ref_shim): GetNumberOfCores (Renderer.cpp:55). Tile::triangleRefs is a fixed [12] array
// indexed by core id (Tile.h:34), so the count is capped at 12; `gShimNumCores` lets a test pin it.
#pragma once
#include <omp.h>
namespace EDX
{
	inline int& ShimNumCoresOverride() { static int n = 0; return n; }
	inline int GetNumberOfCores()
	{
		int n = ShimNumCoresOverride() > 0 ? ShimNumCoresOverride() : omp_get_max_threads();
		return n < 1 ? 1 : (n > 12 ? 12 : n);
	}
}
