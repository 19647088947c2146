// search_cos.cu — beam-search kernel instantiations for f32 rows, METRIC_COSINE (see search_launch.cuh).
#include "search_launch.cuh"

namespace sdb {
namespace launch {
template int launch_float<METRIC_COSINE>(sdb_index*, const SearchArgs&, bool, cudaStream_t);
}  // namespace launch
}  // namespace sdb
