// Force-included before the reference's WellTemperedEnsemble.cc (oracle/Makefile): names its GPU branch mentions.  That branch
// never runs here (exec_mode is CPU); these only have to compile.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <vector>
#include <hoomd/ForceCompute.h>
#define TAG_ALLOCATION(x)
#define CHECK_CUDA_ERROR()
inline int cudaMemsetAsync(void*, int, size_t) { return 0; }
template <class T> struct ScopedAllocation {
    template <class A> ScopedAllocation(const A&, size_t n) : store(n), data(store.data()) {}
    std::vector<T> store;
    T* data;
};
