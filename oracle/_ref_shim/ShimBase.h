// EDXUtil stand-in (oracle/_ref_shim): MSVC spellings and basic typedefs every shim header needs
#pragma once
#include <cassert>
#include <cstddef>
#include <cstdint>

#ifndef __forceinline
#define __forceinline inline __attribute__((always_inline))
#endif
#ifndef MAX_PATH
#define MAX_PATH 260
#endif
#ifndef sprintf_s
#define sprintf_s snprintf                    // Renderer.cpp:355: (buffer, size, format, ...) — snprintf's order
#endif
#ifndef Assert
#define Assert(expr) assert(expr)             // InputBuffer.h:177
#endif

namespace EDX
{
	typedef unsigned int uint;
	typedef unsigned char _byte;
	typedef unsigned char uint8;
}
