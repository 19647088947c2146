// search_dot.cu — beam-search kernel instantiations for f32 rows, METRIC_DOT (see search_launch.cuh).
#include "search_launch.cuh"

namespace sdb {
namespace launch {
template int launch_float<METRIC_DOT>(sdb_index*, const SearchArgs&, bool, cudaStream_t);
}  // namespace launch
}  // namespace sdb
