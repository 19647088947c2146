// temporary stubs until quant.cu / insert.cu land
#include "index.cuh"
namespace sdb {
int insert_batch_locked(sdb_index*, uint64_t, const uint64_t*, const float*) { return fail(SDB_ERR_STATE, "not implemented"); }
int fit_locked(sdb_index*, uint64_t, int32_t*) { return fail(SDB_ERR_STATE, "not implemented"); }
}
