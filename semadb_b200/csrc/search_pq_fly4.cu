// search_pq_fly4.cu — on-the-fly PQ evaluator instantiations for sub-vectors of 4 floats (see search_launch.cuh).
#include "search_launch.cuh"

namespace sdb {
namespace launch {
template int launch_pq_fly<4>(sdb_index*, const SearchArgs&, bool, cudaStream_t);
}  // namespace launch
}  // namespace sdb
