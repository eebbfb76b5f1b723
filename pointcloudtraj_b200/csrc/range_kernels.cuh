// range_kernels.cuh -- fixed-radius range queries (kd_nearest_range3).
#pragma once
#include "query_kernels.cuh"
