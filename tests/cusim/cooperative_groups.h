// cusim: <cooperative_groups.h> reduced to thread-block clusters — TEST INFRASTRUCTURE.  Clusters of more than one
// block only exist in the CUSIM_CLUSTERS build variant (see cuda_runtime.h).
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct cluster_group {
  unsigned block_rank() const { return cusim::cluster_rank(); }
  unsigned num_blocks() const { return cusim::cluster_size(); }
  void sync() const { cusim::cluster_sync(); }
  template <class T> T* map_shared_rank(T* p, unsigned r) const {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(const_cast<typename std::remove_const<T>::type*>(p)) + cusim::cluster_delta(r));
  }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups
