// search_l2.cu — beam-search kernel instantiations for f32 rows, METRIC_EUCLIDEAN (see search_launch.cuh).
#include "search_launch.cuh"

namespace sdb {
namespace launch {
template int launch_float<METRIC_EUCLIDEAN>(sdb_index*, const SearchArgs&, bool, cudaStream_t);
}  // namespace launch
}  // namespace sdb
