// Stand-in for hoomd/GlobalArray.h (the types live in the ForceCompute.h stand-in).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <hoomd/ForceCompute.h>
