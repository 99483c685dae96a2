// Tiled TMA kernels for K1 (fast path).  Placeholder until the first specialisation lands:
// the planner reports "not applicable" and pg_fd_lib_gram runs the generic kernel.
#include "common.cuh"
#include "launch.h"

namespace pg {

bool tiled_plan(const K1Params &, int, int64_t, int, TiledPlan &) { return false; }

int tiled_launch(const K1Params &, int, const TiledPlan &, double *, char *, cudaStream_t) {
    PG_FAIL(PG_EUNSUPPORTED, "tiled kernel not built");
}

}  // namespace pg
