// host_search.hpp — nearest-neighbour descent over a HOST MIRROR of the device tree, for the two cases in which the
// work cannot be handed to the GPU as it is:
//
//   * the caller's own code has to run inside the search — a user-defined `Metric_` (the reference accepts any type
//     with `space_category`, `operator()(begin1, end1, begin2)` and `operator()(x)`, kd_tree.hpp:19-36,
//     examples/kd_tree/kd_tree_custom_metric.cpp) or a user-defined visitor passed to `search_nearest`
//     (kd_tree.hpp:106-120, examples/kd_tree/kd_tree_custom_search_visitor.cpp): a C++ functor of the caller
//     cannot be called from a kernel of a prebuilt library;
//   * ONE query at a time when the caller opted in with pico_tree::b200::single_query_on_host(true): a device call
//     costs ~20 us of launch + synchronisation whatever it computes, the descent below ~0.5 us. Off by default —
//     everything goes to the device unless asked otherwise, and batches always do.
//
// The tree is still built on the device (pico_b200_tree_create; a user-defined metric only changes which bounds
// the nodes keep); this header walks a copy of its flat node array, fetched once per tree through
// pico_b200_tree_export, and reads the points from the caller's own space like the reference does. It restates
// internal/kd_tree_search.hpp:46-105 (euclidean) and :122-229 (topological) and the visitors of
// internal/search_visitor.hpp:20-288 of the reference; it shares no code with oracle/ (test infrastructure).
#pragma once

#include <algorithm>
#include <cstdint>
#include <iterator>
#include <limits>
#include <type_traits>
#include <vector>

#include "../pico_b200.h"

namespace pico_tree {
namespace b200 {

// Opt-in switch for single-query calls of trees with a device metric (see above). Process-wide.
inline bool& single_query_on_host_flag() {
  static bool on = false;
  return on;
}
inline void single_query_on_host(bool on) { single_query_on_host_flag() = on; }

template <typename Scalar_>
struct host_node_of {
  using type = std::conditional_t<sizeof(Scalar_) == 4, pico_b200_node_f32, pico_b200_node_f64>;
};

// Copy of the device tree's structure.
template <typename Scalar_>
struct host_tree {
  using node_type = typename host_node_of<Scalar_>::type;
  std::vector<node_type> nodes;        // pre-order: left child of i is i + 1
  std::vector<std::int32_t> indices;   // leaf position -> index of the point in the space
  std::vector<Scalar_> outer;          // {left_min, right_max} per node, topological spaces only
  std::vector<Scalar_> root_box;       // min[sdim] then max[sdim]
};

// ---------------------------------------------------------------- visitors (internal/search_visitor.hpp)
// search_nn (:41-65)
template <typename Neighbor_>
class visit_nn {
 public:
  using scalar_type = typename Neighbor_::scalar_type;
  explicit visit_nn(Neighbor_& nn) : nn_(nn) { nn_.distance = std::numeric_limits<scalar_type>::max(); }
  template <typename Index_>
  void operator()(Index_ idx, scalar_type d) {
    if (nn_.distance > d) nn_ = Neighbor_(static_cast<typename Neighbor_::index_type>(idx), d);
  }
  scalar_type max() const { return nn_.distance; }

 private:
  Neighbor_& nn_;
};

// search_knn (:82-123) with insert_sorted (:20-38): the range is kept sorted; among equal distances the earlier
// visited point stays in front.
template <typename It_>
class visit_knn {
 public:
  using neighbor_type = typename std::iterator_traits<It_>::value_type;
  using scalar_type = typename neighbor_type::scalar_type;
  visit_knn(It_ begin, It_ end) : begin_(begin), end_(end), active_end_(begin) {
    if (begin_ != end_) std::prev(end_)->distance = std::numeric_limits<scalar_type>::max();
  }
  template <typename Index_>
  void operator()(Index_ idx, scalar_type d) {
    if (!(max() > d)) return;
    if (active_end_ < end_) ++active_end_;
    It_ it = std::prev(active_end_);
    for (; it > begin_ && std::prev(it)->distance > d; --it) *it = *std::prev(it);
    *it = neighbor_type(static_cast<typename neighbor_type::index_type>(idx), d);
  }
  scalar_type max() const { return std::prev(end_)->distance; }

 private:
  It_ begin_, end_, active_end_;
};

// search_radius (:126-156): everything strictly inside the radius, in visit order
template <typename Neighbor_>
class visit_radius {
 public:
  using scalar_type = typename Neighbor_::scalar_type;
  visit_radius(scalar_type radius, std::vector<Neighbor_>& n) : radius_(radius), n_(n) { n_.clear(); }
  template <typename Index_>
  void operator()(Index_ idx, scalar_type d) {
    if (radius_ > d) n_.emplace_back(static_cast<typename Neighbor_::index_type>(idx), d);
  }
  scalar_type max() const { return radius_; }
  void sort() const {
    std::sort(n_.begin(), n_.end(), [](Neighbor_ const& a, Neighbor_ const& b) { return a.distance < b.distance; });
  }

 private:
  scalar_type radius_;
  std::vector<Neighbor_>& n_;
};

// search_approximate_* (:164-288): every distance is scaled by 1 / e before the inner visitor sees it
template <typename Inner_>
class visit_scaled {
 public:
  using scalar_type = typename Inner_::scalar_type;
  visit_scaled(scalar_type e, Inner_& inner) : e_inv_(scalar_type(1.0) / e), inner_(inner) {}
  template <typename Index_>
  void operator()(Index_ idx, scalar_type d) {
    inner_(idx, d * e_inv_);
  }
  scalar_type max() const { return inner_.max(); }

 private:
  scalar_type e_inv_;
  Inner_& inner_;
};

// ---------------------------------------------------------------- the descent
// `point_of(i)` returns a pointer to the sdim scalars of point i of the caller's space.
template <typename Scalar_, typename Metric_, typename PointOf_, typename Visitor_>
class host_descent {
  using node_type = typename host_tree<Scalar_>::node_type;

 public:
  host_descent(host_tree<Scalar_> const& tree, Metric_ const& metric, PointOf_ point_of, Scalar_ const* query,
               std::size_t sdim, Visitor_& visitor)
      : tree_(tree), metric_(metric), point_of_(point_of), q_(query), sdim_(sdim), visitor_(visitor) {
    if (sdim > kInline) {
      big_.resize(sdim);
      offset_ = big_.data();
    }
  }

  void operator()() {
    if (tree_.nodes.empty()) return;
    std::fill(offset_, offset_ + sdim_, Scalar_(0));
    if constexpr (std::is_same_v<typename Metric_::space_category, euclidean_space_tag>)
      walk<true>(0, Scalar_(0));
    else
      walk<false>(0, Scalar_(0));
  }

 private:
  template <bool Euclidean_>
  void walk(std::uint32_t node, Scalar_ box_distance) {
    node_type const& nd = tree_.nodes[node];
    if (nd.split_dim == PICO_B200_LEAF) {
      // kd_tree_search.hpp:54-59
      auto const begin = static_cast<std::size_t>(nd.a.begin_idx), end = static_cast<std::size_t>(nd.b.end_idx);
      for (std::size_t i = begin; i < end; ++i) {
        std::int32_t const idx = tree_.indices[i];
        visitor_(idx, metric_(q_, q_ + sdim_, point_of_(static_cast<std::size_t>(idx))));
      }
      return;
    }
    std::size_t const sd = nd.split_dim;
    Scalar_ const v = q_[sd];
    std::uint32_t first, second;
    Scalar_ new_offset;
    if constexpr (Euclidean_) {
      // :76-88 — the nearer child first; the other one lies at least metric(bound - v) away on this dimension
      if ((nd.a.left_max + nd.b.right_min - v - v) > 0) {
        first = node + 1;
        second = nd.right;
        new_offset = metric_(nd.b.right_min - v);
      } else {
        first = nd.right;
        second = node + 1;
        new_offset = metric_(nd.a.left_max - v);
      }
    } else {
      // :166-186 — distances to the boxes of both children, on the line or on the circle
      Scalar_ const d1 = segment_distance(tree_.outer[2 * node], nd.a.left_max, v, static_cast<int>(sd));
      Scalar_ const d2 = segment_distance(nd.b.right_min, tree_.outer[2 * node + 1], v, static_cast<int>(sd));
      if (d1 < d2) {
        first = node + 1;
        second = nd.right;
        new_offset = d2;
      } else {
        first = nd.right;
        second = node + 1;
        new_offset = d1;
      }
    }
    walk<Euclidean_>(first, box_distance);
    // :93-103 — incremental box distance, one rounding per operation, left to right
    Scalar_ const old_offset = offset_[sd];
    box_distance = box_distance - old_offset + new_offset;
    if (visitor_.max() >= box_distance) {
      offset_[sd] = new_offset;
      walk<Euclidean_>(second, box_distance);
      offset_[sd] = old_offset;
    }
  }

  // search_nearest_topological::box_distance (:205-229) with segment_r1 / segment_s1::distance (segment.hpp:34-100)
  Scalar_ segment_distance(Scalar_ mn, Scalar_ mx, Scalar_ v, int dim) const {
    Scalar_ d(0);
    metric_.apply_dim_space(dim, [&](auto one_space) {
      if constexpr (std::is_same_v<decltype(one_space), one_space_s1>) {
        auto const s1 = [](Scalar_ x, Scalar_ y) {
          Scalar_ const a = x > y ? x - y : y - x;
          return std::min(a, Scalar_(1.0) - a);
        };
        bool const outside = (mn <= mx) ? (v < mn || v > mx) : !(v < mx || v > mn);
        d = outside ? std::min(s1(v, mn), s1(v, mx)) : Scalar_(0);
      } else {
        d = v < mn ? mn - v : (v > mx ? v - mx : Scalar_(0));
      }
    });
    return metric_(d);
  }

  host_tree<Scalar_> const& tree_;
  Metric_ const& metric_;
  PointOf_ point_of_;
  Scalar_ const* q_;
  std::size_t sdim_;
  Visitor_& visitor_;
  static constexpr std::size_t kInline = 16;
  Scalar_ small_[kInline];
  std::vector<Scalar_> big_;
  Scalar_* offset_ = small_;  // node_box_offset_
};

}  // namespace b200
}  // namespace pico_tree
