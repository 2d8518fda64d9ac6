// trunk.cu — placeholder until the conv-net kernels land.
#include "common.cuh"
namespace iago {
void trunk_destroy(iago_ctx *) {}
}  // namespace iago
