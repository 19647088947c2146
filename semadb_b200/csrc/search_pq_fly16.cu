// search_pq_fly16.cu — on-the-fly PQ evaluator instantiations for sub-vectors of 16 floats (see search_launch.cuh).
#include "search_launch.cuh"

namespace sdb {
namespace launch {
template int launch_pq_fly<16>(sdb_index*, const SearchArgs&, bool, cudaStream_t);
}  // namespace launch
}  // namespace sdb
