#include "mb_internal.h"
namespace mb {
int launch_vol3d(Ctx *, const DevPlan &, const void *, void *, const OuterRange *) { return -1; }
}
