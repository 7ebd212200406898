// EDXUtil stand-in (oracle/_ref_shim): Application::GetBaseDirectory (Renderer.cpp:355)
#pragma once
namespace EDX
{
	class Application
	{
	public:
		static const char* GetBaseDirectory() { return "."; }
	};
}
