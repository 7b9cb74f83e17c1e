// forest_host.cpp — TEST ONLY: compiles pico_tree_b200/csrc/forest.cuh (the `__host__ __device__` traversal core
// of the kd_forest path) for the host and drives it like kd_forest::search_nearest does
// (examples/pico_understory/pico_understory/kd_forest.hpp:91-120): for every tree, reflect the query, run the
// best-bin-first search, all trees sharing one search_knn visitor (internal/search_visitor.hpp:82-123).
// Built by tests/cpp/Makefile into tests/_bin/libforest_host.so with -ffp-contract=off and used by
// tests/test_forest_core.py against the reference fixtures. Not part of the product.
#include <cstdint>
#include <limits>
#include <vector>

#include "../../pico_tree_b200/csrc/forest.cuh"

namespace {

template <typename T>
struct Neighbor {
  int32_t index;
  T distance;
};

// search_knn::operator() + insert_sorted (search_visitor.hpp:20-38,106-114)
template <typename T>
struct KnnVisitor {
  Neighbor<T>* out;
  size_t k, active;
  T max() const { return out[k - 1].distance; }
  void visit(int32_t idx, T d) {
    if (!(max() > d)) return;
    if (active < k) ++active;
    size_t end = active - 1;
    for (; end > 0 && d < out[end - 1].distance; --end) out[end] = out[end - 1];
    out[end].index = idx;
    out[end].distance = d;
  }
};

template <typename T>
int run(size_t n_trees, void const* const* nodes, T const* const* outer, int32_t const* const* indices,
        T const* const* points, uint32_t sdim, uint32_t height, T const* rotations, T const* q, size_t nq, size_t k,
        size_t max_leaves, uint32_t queue_cap, Neighbor<T>* out) {
  std::vector<pico::forest::Entry<T>> storage(queue_cap), path(height + 2);
  std::vector<T> rq(sdim);
  int overflow = 0;
  for (size_t i = 0; i < nq; ++i) {
    Neighbor<T>* row = out + i * k;
    for (size_t j = 0; j < k; ++j) {
      row[j].index = -1;
      row[j].distance = std::numeric_limits<T>::max();
    }
    KnnVisitor<T> vis{row, k, 0};
    for (size_t t = 0; t < n_trees; ++t) {
      pico::forest::TreeView<T> view{static_cast<typename pico::forest::NodeOfT<T>::type const*>(nodes[t]), outer[t],
                                     indices[t], points[t], sdim, sdim};
      pico::forest::householder(rotations + t * sdim, sdim, q + i * sdim, rq.data());
      pico::forest::MinQueue<T> queue{storage.data(), 0, queue_cap, false};
      pico::forest::priority_search_tree(view, rq.data(), max_leaves, vis, queue, path.data());
      overflow |= queue.overflow;
    }
  }
  return overflow;
}

}  // namespace

extern "C" {

int forest_host_knn_f32(size_t n_trees, void const* const* nodes, float const* const* outer,
                        int32_t const* const* indices, float const* const* points, uint32_t sdim, uint32_t height,
                        float const* rotations, float const* q, size_t nq, size_t k, size_t max_leaves,
                        uint32_t queue_cap, void* out) {
  return run<float>(n_trees, nodes, outer, indices, points, sdim, height, rotations, q, nq, k, max_leaves, queue_cap,
                    static_cast<Neighbor<float>*>(out));
}
int forest_host_knn_f64(size_t n_trees, void const* const* nodes, double const* const* outer,
                        int32_t const* const* indices, double const* const* points, uint32_t sdim, uint32_t height,
                        double const* rotations, double const* q, size_t nq, size_t k, size_t max_leaves,
                        uint32_t queue_cap, void* out) {
  return run<double>(n_trees, nodes, outer, indices, points, sdim, height, rotations, q, nq, k, max_leaves, queue_cap,
                     static_cast<Neighbor<double>*>(out));
}

}  // extern "C"
