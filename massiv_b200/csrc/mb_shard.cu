#include "mb_internal.h"
namespace mb {
void shard_destroy(Ctx *) {}
}
extern "C" {
int mb_nccl_unique_id(uint8_t *) { return MB_ERR_UNSUPPORTED; }
int mb_shard_init(mb_ctx *, int, int, const uint8_t *) { return MB_ERR_UNSUPPORTED; }
int mb_shard_halo(const mb_stencil_desc *, int64_t *, int64_t *) { return MB_ERR_UNSUPPORTED; }
int mb_iterate_sharded_dev(mb_ctx *, const mb_stencil_desc *, const int64_t *, const int64_t *, void *, void *, int64_t, void **) { return MB_ERR_UNSUPPORTED; }
}
