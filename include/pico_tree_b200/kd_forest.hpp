// kd_forest.hpp — pico_tree::kd_forest<Space_, Metric_, Index_> on top of libpico_b200.so.
//
// Stand-in for examples/pico_understory/pico_understory/kd_forest.hpp:15-138 of the reference: same template
// parameters and member types, the constructor kd_forest(space, max_leaf_size, forest_size) (:44-52),
// search_nn(x, max_leaves_visited, nn) (:83-88), the deduction guide and make_kd_forest (:127-138). Building
// (Householder reflections + one tree per copy) and the best-bin-first search run on the device
// (pico_tree_b200/csrc/forest.cu) behind pico_b200_forest_* of include/pico_b200.h; the header holds no distance
// arithmetic and no tree walk.
//
// search_nearest(x, max_leaves_visited, visitor) of the reference takes any visitor; on the device the visitor is
// the reference's search_nn / search_knn (one sorted list shared by all trees), reachable here as search_nn and
// search_knn. Additions: search_nn_batch / search_knn_batch (a whole query set per call — the way the device is
// meant to be used), and an optional last constructor argument with the reflection vectors (the reference draws
// them from std::random_device, so its forests are not reproducible).
#pragma once

#include "kd_tree.hpp"

namespace pico_tree {

namespace b200 {
struct forest_deleter {
  void operator()(pico_b200_forest* f) const { pico_b200_forest_destroy(f); }
};
}  // namespace b200

template <typename Space_, typename Metric_ = metric_l2_squared, typename Index_ = int>
class kd_forest {
  static_assert(std::is_same_v<Metric_, metric_l2_squared>,
                "KD_FOREST_SUPPORTS_METRIC_L2_SQUARED_ONLY");  // like priority_search_nearest_euclidean's users
  static_assert(std::is_integral_v<Index_> && sizeof(Index_) == 4, "KD_FOREST_NEEDS_A_32_BIT_INDEX_TYPE");
  using unwrapped_space = b200::unwrap_t<Space_>;
  using traits = space_traits<unwrapped_space>;
  using rows_type = b200::rows_of<Space_>;

 public:
  using size_type = size_t;
  using index_type = Index_;
  using scalar_type = typename traits::scalar_type;
  static constexpr size_type dim = traits::dim;
  using space_type = Space_;
  using metric_type = Metric_;
  using neighbor_type = neighbor<index_type, scalar_type>;

  kd_forest(space_type space, size_type max_leaf_size, size_type forest_size,
            std::vector<scalar_type> const& rotations = {})
      : space_(std::move(space)), metric_() {
    unwrapped_space const& s = space_;
    rows_type rows(s);
    if (!rotations.empty() && rotations.size() != forest_size * rows.sdim())
      throw std::invalid_argument("rotations must hold forest_size vectors of the space's dimension");
    pico_b200_forest* h = nullptr;
    b200::check(pico_b200_forest_create(rows.data(), rows.size(), rows.sdim(), rows.stride(),
                                        b200::scalar_id<scalar_type>::value, max_leaf_size,
                                        rotations.empty() ? nullptr : rotations.data(), forest_size,
                                        b200::default_device(), &h));
    handle_.reset(h);
    n_ = rows.size();
    sdim_ = rows.sdim();
  }

  kd_forest(kd_forest const&) = delete;
  kd_forest(kd_forest&&) = default;
  kd_forest& operator=(kd_forest const&) = delete;
  kd_forest& operator=(kd_forest&&) = default;

  // kd_forest::search_nn (kd_forest.hpp:83-88): at most max_leaves_visited leaves per tree.
  template <typename P_>
  void search_nn(P_ const& x, size_type max_leaves_visited, neighbor_type& nn) const {
    call(query_data(x), 1, sdim_, 1, max_leaves_visited, &nn);
  }

  // search_nearest with a search_knn visitor: the k best of everything the budgeted searches saw, ascending.
  template <typename P_>
  void search_knn(P_ const& x, size_type k, size_type max_leaves_visited, std::vector<neighbor_type>& knn) const {
    knn.resize(std::min(k, n_));
    call(query_data(x), 1, sdim_, knn.size(), max_leaves_visited, knn.data());
  }

  template <typename Queries_>
  void search_knn_batch(Queries_ const& queries, size_type k, size_type max_leaves_visited,
                        std::vector<neighbor_type>& knn) const {
    b200::unwrap_t<Queries_> const& q = queries;
    b200::rows_of<Queries_> rows(q);
    if (rows.size() != 0 && rows.sdim() != sdim_) throw std::invalid_argument("query and forest dimensions differ");
    k = std::min(k, n_);
    knn.resize(rows.size() * k);
    call(rows.data(), rows.size(), rows.stride(), k, max_leaves_visited, knn.data());
  }

  template <typename Queries_>
  void search_nn_batch(Queries_ const& queries, size_type max_leaves_visited, std::vector<neighbor_type>& nns) const {
    search_knn_batch(queries, 1, max_leaves_visited, nns);
  }

  // The reflection vectors in use (rkd_tree_hh_data::rotation of every tree), forest_size x sdim.
  std::vector<scalar_type> rotations() const {
    std::vector<scalar_type> r(info().n_trees * sdim_);
    b200::check(pico_b200_forest_rotations(handle_.get(), r.data()));
    return r;
  }

  pico_b200_forest_info info() const {
    pico_b200_forest_info i;
    b200::check(pico_b200_forest_info_get(handle_.get(), &i));
    return i;
  }

  space_type const& space() const { return space_; }
  metric_type const& metric() const { return metric_; }
  pico_b200_forest const* native_handle() const { return handle_.get(); }

 private:
  static_assert(sizeof(neighbor_type) == (sizeof(scalar_type) == 4 ? 8 : 16), "NEIGHBOR_LAYOUT_DIFFERS_FROM_THE_ABI");

  template <typename P_>
  scalar_type const* query_data(P_ const& x) const {
    using pt = b200::point_traits_of<P_>;
    static_assert(std::is_same_v<scalar_type, typename pt::scalar_type>, "POINT_AND_TREE_SCALAR_TYPES_DIFFER");
    if constexpr (pt::dim == dynamic_extent) {
      if (pt::size(x) != sdim_) throw std::invalid_argument("point and forest dimensions differ");
    }
    return pt::data(x);
  }

  void call(scalar_type const* q, size_type nq, size_type stride, size_type k, size_type max_leaves,
            neighbor_type* out) const {
    if (nq == 0 || k == 0) return;
    b200::check(pico_b200_forest_knn(handle_.get(), q, nq, stride, k, max_leaves, out, 0, nullptr));
  }

  space_type space_;
  metric_type metric_;
  std::unique_ptr<pico_b200_forest, b200::forest_deleter> handle_;
  size_type n_ = 0;
  size_type sdim_ = 0;
};

template <typename Space_>
kd_forest(Space_, size_t, size_t) -> kd_forest<Space_, metric_l2_squared, int>;

template <typename Metric_ = metric_l2_squared, typename Index_ = int, typename Space_>
kd_forest<std::decay_t<Space_>, Metric_, Index_> make_kd_forest(Space_&& space, size_t max_leaf_size,
                                                                size_t forest_size) {
  return kd_forest<std::decay_t<Space_>, Metric_, Index_>(std::forward<Space_>(space), max_leaf_size, forest_size);
}

}  // namespace pico_tree
