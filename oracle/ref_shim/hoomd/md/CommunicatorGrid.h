// Stand-in (MPI is off in the oracle cross-check build): nothing to declare.
#pragma once
