// EDXUtil stand-in (oracle/_ref_shim): Memory::SafeDelete / SafeDeleteArray (InputBuffer.h:130, Renderer.cpp:367-368)
#pragma once
#include "../ShimBase.h"
namespace EDX
{
	namespace Memory
	{
		template<class T> inline void SafeDelete(T*& p) { if (p) { delete p; p = nullptr; } }
		template<class T> inline void SafeDeleteArray(T*& p) { if (p) { delete[] p; p = nullptr; } }
	}
}
