#include "mb_internal.h"
namespace mb {
int launch_tile2d(Ctx *, const DevPlan &, const void *, void *, const OuterRange *) { return -1; }
int run_sep2_fused(Ctx *, const DevPlan &, const DevPlan &, const void *, void *) { return -1; }
}
