// EDXUtil stand-in (oracle/_ref_shim): UniquePtr / MakeUnique (Renderer.h:18-22, Mesh.h:20-23)
#pragma once
#include "../ShimBase.h"
#include <memory>
#include <utility>
namespace EDX
{
	template<class T>
	class UniquePtr : public std::unique_ptr<T>
	{
	public:
		typedef std::unique_ptr<T> Base;
		UniquePtr() {}
		UniquePtr(std::nullptr_t) {}
		explicit UniquePtr(T* p) : Base(p) {}
		UniquePtr(UniquePtr&& o) : Base(std::move(o)) {}
		template<class U> UniquePtr(UniquePtr<U>&& o) : Base(std::move(o)) {}
		UniquePtr& operator=(UniquePtr&& o) { Base::operator=(std::move(o)); return *this; }
		template<class U> UniquePtr& operator=(UniquePtr<U>&& o) { Base::operator=(std::move(o)); return *this; }
		T* Get() const { return Base::get(); }
		void Reset(T* p = nullptr) { Base::reset(p); }
	};

	template<class T, class... Args>
	inline UniquePtr<T> MakeUnique(Args&&... args) { return UniquePtr<T>(new T(std::forward<Args>(args)...)); }
}
