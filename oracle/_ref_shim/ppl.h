// MSVC Parallel Patterns Library stand-in (oracle/_ref_shim): concurrency::parallel_for(first, last, body) over
// OpenMP with dynamic scheduling (DESIGN.md shim 15). The reference's bodies write disjoint outputs.
#pragma once
#include <omp.h>
#include <algorithm>
namespace concurrency
{
	template<class Index, class Func>
	inline void parallel_for(const Index first, const Index last, const Func& body)
	{
		// PPL splits the range into stolen sub-ranges; a chunked dynamic schedule is the closest OpenMP shape
		const long long n = (long long)last - (long long)first;
		const int chunk = (int)std::max<long long>(1, std::min<long long>(4096, n / (64LL * omp_get_max_threads())));
		#pragma omp parallel for schedule(dynamic, chunk)
		for (Index i = first; i < last; i++)
			body(i);
	}
}
