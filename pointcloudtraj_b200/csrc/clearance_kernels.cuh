// clearance_kernels.cuh -- Bezier sampling fused with the nearest-obstacle query (checkSafeTrajectory).
#pragma once
#include "query_kernels.cuh"
