// EDXUtil stand-in, part of oracle/_ref_shim (TEST INFRASTRUCTURE ONLY — see README.md in this directory).
//
// behindthepixels/EDXUtil is the reference's only dependency and is absent from /root/reference (an unpinned
// sibling checkout, EDXRaster.sln:11). These headers define the ~25 EDXUtil symbols the raster path uses so that
// g++ can compile the reference's OWN, UNMODIFIED sources (Core/*.h, Core/*.cpp, Utils/*) where they lie.
// Every definition here is listed, with the call site it was inferred from, in DESIGN.md section 2.
#pragma once

#include <smmintrin.h>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <memory>
#include <utility>
#include <vector>

#include "ShimBase.h"

namespace EDX
{
	template<class T>
	inline void Swap(T& a, T& b) { T t = a; a = b; b = t; }      // Clipper.h:231
}

#include "Core/Memory.h"
#include "Core/SmartPointer.h"
#include "Containers/Array.h"
#include "Math/EDXMath.h"
#include "Math/Vector.h"
#include "Math/Matrix.h"
#include "Graphics/Color.h"
