#include "mb_internal.h"
namespace mb {
int launch_life(Ctx *, const DevPlan &, const void *, void *) { return -1; }
int life_iterate(Ctx *, const DevPlan &, void *, void *, int64_t, void **) { return -1; }
}
