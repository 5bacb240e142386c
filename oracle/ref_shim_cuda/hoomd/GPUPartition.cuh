// Stand-in for hoomd/GPUPartition.cuh (declaration only; see Autotuner.h in this directory).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <hoomd/ForceCompute.h>
