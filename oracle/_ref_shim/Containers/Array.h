// EDXUtil stand-in (oracle/_ref_shim): Array<T>, a growable contiguous container (std::vector semantics).
// Members used by the reference: Resize, Clear, Size, Add, Insert, Data, operator[], range-for, copy assignment.
#pragma once
#include "../ShimBase.h"
#include <vector>
#include <utility>
#include <cstddef>
namespace EDX
{
	template<class T>
	class Array
	{
	private:
		std::vector<T> mData;
	public:
		void Resize(const size_t n) { mData.resize(n); }
		void Clear() { mData.clear(); }
		size_t Size() const { return mData.size(); }
		bool Empty() const { return mData.empty(); }
		void Add(const T& v) { mData.push_back(v); }
		void Add(T&& v) { mData.push_back(std::move(v)); }
		// Renderer.cpp:243: Insert(pointer, count, position)
		void Insert(const T* p, const size_t count, const size_t at) { mData.insert(mData.begin() + at, p, p + count); }
		T* Data() { return mData.data(); }
		const T* Data() const { return mData.data(); }
		T& operator[](const size_t i) { return mData[i]; }
		const T& operator[](const size_t i) const { return mData[i]; }
		typename std::vector<T>::iterator begin() { return mData.begin(); }
		typename std::vector<T>::iterator end() { return mData.end(); }
		typename std::vector<T>::const_iterator begin() const { return mData.begin(); }
		typename std::vector<T>::const_iterator end() const { return mData.end(); }
	};
}
