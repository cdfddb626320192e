// Kernel 1, tensor-core form (tcgen05 / TMEM / TMA).  Placeholder until the sm_100a
// pipeline lands: reports "unsupported" so TOME_MATCH_AUTO routes to the exact kernel.
#include "common.cuh"

namespace tome {
size_t match_tc_workspace(int, int, int) { return 0; }
bool match_tc_supported(int, int, int, int, const View&, const void*) { return false; }
int launch_match_tc(const void*, int, int, int, int, const View&, int, int, float*, int*, void*, size_t, cudaStream_t) {
  return set_error(TOME_ERR_UNSUPPORTED, "tome_match: tcgen05 path not built");
}
}  // namespace tome
