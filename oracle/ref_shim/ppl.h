// Test infrastructure only. Stand-in for MSVC's <ppl.h> so that the reference's
// TensorOpCpuMt.cpp (which uses concurrency::parallel_for, TensorOpCpuMt.cpp:1,7)
// compiles under g++. parallel_for is mapped onto an OpenMP dynamic loop; nested
// calls become nested OpenMP regions (enable with omp_set_max_active_levels(2)).
#pragma once
#include <omp.h>

namespace concurrency
{
    template <typename Index, typename Func>
    void parallel_for(Index first, Index last, const Func& f)
    {
        long long lo = (long long)first, hi = (long long)last;
#pragma omp parallel for schedule(dynamic)
        for (long long i = lo; i < hi; ++i)
            f((Index)i);
    }
}
