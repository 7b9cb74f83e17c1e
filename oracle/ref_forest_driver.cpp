// ref_forest_driver.cpp — C entry points around the UNMODIFIED kd_forest of the reference
// (/root/reference/examples/pico_understory/pico_understory/kd_forest.hpp:15-138), compiled into
// oracle/_ref/libpico_ref_forest.so by oracle/Makefile. TEST INFRASTRUCTURE ONLY: it pins the C restatement
// of the forest (pico_oracle_impl.inc, po_forest_*) and generates tests/golden/forest_*.npz.
//
// kd_forest draws its Householder vectors from std::random_device (internal/rkd_tree_hh_data.hpp:16-31),
// so the restatement can only be compared on the same vectors: the class is built and searched exactly as
// a user would, and the vectors it drew are then READ OUT of its private `data_` member. That is the one
// liberty taken here (`#define private public` around the single include of kd_forest.hpp, after every
// header it depends on has been parsed normally); no reference source is copied or changed.
#include <cstdint>
#include <cstring>
#include <memory>
#include <queue>
#include <random>
#include <vector>

#include <pico_tree/internal/kd_tree_builder.hpp>
#include <pico_tree/internal/kd_tree_data.hpp>
#include <pico_tree/internal/kd_tree_node.hpp>
#include <pico_tree/internal/point.hpp>
#include <pico_tree/internal/point_wrapper.hpp>
#include <pico_tree/internal/search_visitor.hpp>
#include <pico_tree/internal/space_wrapper.hpp>
#include <pico_tree/map_traits.hpp>
#include <pico_tree/metric.hpp>
#include <pico_understory/internal/kd_tree_priority_search.hpp>
#include <pico_understory/internal/rkd_tree_builder.hpp>

#define private public
#include <pico_understory/kd_forest.hpp>
#undef private

namespace {

struct forest_base {
  virtual ~forest_base() = default;
  virtual size_t n_trees() const = 0;
  virtual void rotations(void* out) const = 0;
  virtual void knn(void const* q, size_t nq, size_t k, size_t max_leaves, void* out, int threads) const = 0;
};

template <typename T>
struct forest_impl final : forest_base {
  static constexpr size_t Dim = pico_tree::dynamic_extent;
  using point_type = pico_tree::point_map<T const, Dim>;
  using space_type = pico_tree::space_map<point_type>;
  using forest_type = pico_tree::kd_forest<space_type, pico_tree::metric_l2_squared, int>;
  using neighbor_type = typename forest_type::neighbor_type;

  forest_impl(T const* pts, size_t n, size_t sdim, size_t max_leaf_size, size_t forest_size)
      : sdim_(sdim), forest_(space_type(pts, n, sdim), max_leaf_size, forest_size) {}

  size_t n_trees() const override { return forest_.data_.size(); }

  void rotations(void* outv) const override {
    T* out = static_cast<T*>(outv);
    for (size_t i = 0; i < forest_.data_.size(); ++i)
      for (size_t j = 0; j < sdim_; ++j) out[i * sdim_ + j] = forest_.data_[i].rotation[j];
  }

  // kd_forest::search_nn (kd_forest.hpp:80-85) for k == 1, kd_forest::search_nearest (:69-76) with the
  // reference's search_knn visitor (internal/search_visitor.hpp:82-123) for k > 1.
  void knn(void const* qv, size_t nq, size_t k, size_t max_leaves, void* outv, int threads) const override {
    T const* q = static_cast<T const*>(qv);
    neighbor_type* out = static_cast<neighbor_type*>(outv);
    auto one = [&](size_t i) {
      point_type p(q + i * sdim_, sdim_);
      if (k == 1) {
        forest_.search_nn(p, max_leaves, out[i]);
      } else {
        pico_tree::internal::search_knn<neighbor_type*> v(out + i * k, out + (i + 1) * k);
        forest_.search_nearest(p, max_leaves, v);
      }
    };
    std::ptrdiff_t const cnt = static_cast<std::ptrdiff_t>(nq);
    if (threads <= 1) {
      for (std::ptrdiff_t i = 0; i < cnt; ++i) one(size_t(i));
    } else {
#pragma omp parallel for schedule(dynamic, 128) num_threads(threads)
      for (std::ptrdiff_t i = 0; i < cnt; ++i) one(size_t(i));
    }
  }

  size_t sdim_;
  forest_type forest_;
};

}  // namespace

extern "C" {

void* ref_forest_create(
    void const* pts, size_t n, size_t sdim, int scalar, size_t max_leaf_size, size_t forest_size) {
  if (scalar == 0)
    return new forest_impl<float>(static_cast<float const*>(pts), n, sdim, max_leaf_size, forest_size);
  return new forest_impl<double>(static_cast<double const*>(pts), n, sdim, max_leaf_size, forest_size);
}
void ref_forest_free(void* f) { delete static_cast<forest_base*>(f); }
size_t ref_forest_size(void const* f) { return static_cast<forest_base const*>(f)->n_trees(); }
void ref_forest_rotations(void const* f, void* out) { static_cast<forest_base const*>(f)->rotations(out); }
void ref_forest_knn(
    void const* f, void const* q, size_t nq, size_t k, size_t max_leaves, void* out, int threads) {
  static_cast<forest_base const*>(f)->knn(q, nq, k, max_leaves, out, threads);
}

}  // extern "C"
