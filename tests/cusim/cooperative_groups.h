// cusim: <cooperative_groups.h> reduced to a thread-block cluster of exactly one block — TEST INFRASTRUCTURE.
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct cluster_group {
  unsigned block_rank() const { return 0; }
  unsigned num_blocks() const { return 1; }
  void sync() const { __syncthreads(); }
  template <class T> T* map_shared_rank(T* p, unsigned) const { return p; }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups
