// traits.hpp — the vocabulary types a pico_tree::kd_tree<> user needs, for builds that do
// NOT have the reference's headers on the include path.
//
// pico_tree_b200/kd_tree.hpp includes this file unless PICO_TREE_B200_USE_REFERENCE_TRAITS is
// defined, in which case the reference's own, unmodified headers provide the same names (and
// bring the Eigen / OpenCV adaptors with them). The names, template parameters and member
// names follow the reference so that user code and adaptors written against it compile:
//   neighbor, dynamic_extent          core.hpp:10-54
//   point_traits / space_traits       point_traits.hpp:8-9, space_traits.hpp:12-27
//   C arrays, std::array, std::vector array_traits.hpp:10-44, vector_traits.hpp:14-41
//   point_map / space_map (+ traits)  map.hpp:101-169, map_traits.hpp:10-44
//   metric tags                       metric.hpp:56-257
//   build tags                        internal/kd_tree_builder.hpp:22-138
// The metric functors are host conveniences (tree.metric()(x), brute-force checks in tests);
// searches never call them — distances come from the CUDA kernels.
#pragma once

#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <functional>
#include <iterator>
#include <limits>
#include <type_traits>
#include <vector>

namespace pico_tree {

using size_t = std::size_t;

// Marks a dimension only known at run time.
inline constexpr size_t dynamic_extent = static_cast<size_t>(-1);

// One search result. Layout {index, distance} is what libpico_b200.so writes for
// Index_ = 32-bit integers (pico_b200.h "neighbours").
template <typename Index_, typename Scalar_>
struct neighbor {
  static_assert(std::is_integral_v<Index_>, "INDEX_NOT_AN_INTEGRAL_TYPE");
  static_assert(std::is_arithmetic_v<Scalar_>, "SCALAR_NOT_AN_ARITHMETIC_TYPE");
  using index_type = Index_;
  using scalar_type = Scalar_;

  constexpr neighbor() = default;
  constexpr neighbor(Index_ i, Scalar_ d) noexcept : index(i), distance(d) {}

  Index_ index;
  Scalar_ distance;
};

template <typename Index_, typename Scalar_>
constexpr bool operator<(neighbor<Index_, Scalar_> const& a, neighbor<Index_, Scalar_> const& b) noexcept {
  return a.distance < b.distance;
}

// ------------------------------------------------------------------ traits: primary templates
template <typename Point_>
struct point_traits;

template <typename Space_>
struct space_traits;

// A std::reference_wrapper<Space> is a space too (the tree then does not copy the point set).
template <typename Space_>
struct space_traits<std::reference_wrapper<Space_>> : space_traits<std::remove_const_t<Space_>> {
  using space_type = std::reference_wrapper<Space_>;
};

// ------------------------------------------------------------------ points: Scalar[Dim], std::array
template <typename Scalar_, std::size_t Dim_>
struct point_traits<Scalar_[Dim_]> {
  using point_type = Scalar_[Dim_];
  using scalar_type = Scalar_;
  using size_type = size_t;
  static constexpr size_type dim = Dim_;
  static constexpr scalar_type const* data(point_type const& p) { return p; }
  static constexpr size_type size(point_type const&) { return Dim_; }
};

template <typename Scalar_, std::size_t Dim_>
struct point_traits<std::array<Scalar_, Dim_>> {
  using point_type = std::array<Scalar_, Dim_>;
  using scalar_type = Scalar_;
  using size_type = size_t;
  static constexpr size_type dim = Dim_;
  static constexpr scalar_type const* data(point_type const& p) { return p.data(); }
  static constexpr size_type size(point_type const&) { return Dim_; }
};

// ------------------------------------------------------------------ spaces: std::vector<Point>
template <typename Point_, typename Allocator_>
struct space_traits<std::vector<Point_, Allocator_>> {
  using space_type = std::vector<Point_, Allocator_>;
  using point_type = Point_;
  using scalar_type = typename point_traits<Point_>::scalar_type;
  using size_type = size_t;
  static constexpr size_type dim = point_traits<Point_>::dim;
  static_assert(dim != dynamic_extent, "VECTOR_OF_POINT_DOES_NOT_SUPPORT_DYNAMIC_DIM");

  template <typename Index_>
  static Point_ const& point_at(space_type const& s, Index_ i) {
    return s[static_cast<size_type>(i)];
  }
  static size_type size(space_type const& s) { return s.size(); }
  static constexpr size_type sdim(space_type const&) { return dim; }
};

// ------------------------------------------------------------------ maps over raw memory
// point_map<Scalar, Dim>: a view of Dim (or a run-time number of) contiguous scalars.
template <typename Scalar_, size_t Dim_>
class point_map {
  static_assert(std::is_arithmetic_v<Scalar_>, "SCALAR_NOT_AN_ARITHMETIC_TYPE");
  static_assert(Dim_ == dynamic_extent || Dim_ > 0, "DIM_MUST_BE_DYNAMIC_OR_>_0");

 public:
  using element_type = Scalar_;
  using scalar_type = std::remove_cv_t<Scalar_>;
  using size_type = size_t;
  static constexpr size_type dim = Dim_;

  explicit constexpr point_map(Scalar_* data) noexcept : data_(data), size_(Dim_) {}
  constexpr point_map(Scalar_* data, size_type size) noexcept
      : data_(data), size_(Dim_ == dynamic_extent ? size : Dim_) {}
  template <typename ContiguousIt_>
  constexpr point_map(ContiguousIt_ begin, ContiguousIt_ end) noexcept
      : data_(&*begin), size_(Dim_ == dynamic_extent ? static_cast<size_type>(end - begin) : Dim_) {}

  constexpr Scalar_& operator[](size_type i) const { return data_[i]; }
  constexpr Scalar_* data() const noexcept { return data_; }
  constexpr size_type size() const noexcept { return size_; }

 private:
  Scalar_* data_;
  size_type size_;
};

// space_map<Point>: a view of `size` contiguous points of a fixed-size point type.
template <typename Point_>
class space_map {
 public:
  using point_type = std::remove_cv_t<Point_>;
  using point_element_type = Point_;
  using scalar_type = typename point_traits<point_type>::scalar_type;
  using size_type = size_t;
  static constexpr size_type dim = point_traits<point_type>::dim;
  static_assert(dim != dynamic_extent, "SPACE_MAP_OF_POINT_DOES_NOT_SUPPORT_DYNAMIC_DIM");

  constexpr space_map(Point_* data, size_type size) noexcept : data_(data), size_(size) {}
  constexpr Point_& operator[](size_type i) const { return data_[i]; }
  constexpr Point_* data() const noexcept { return data_; }
  constexpr size_type size() const noexcept { return size_; }
  constexpr size_type sdim() const { return dim; }

 private:
  Point_* data_;
  size_type size_;
};

// space_map<point_map<Scalar, Dim>>: a row-major matrix of scalars, `sdim` per point.
template <typename Scalar_, size_t Dim_>
class space_map<point_map<Scalar_, Dim_>> {
 public:
  using point_type = point_map<Scalar_, Dim_>;
  using scalar_type = typename point_type::scalar_type;
  using scalar_element_type = Scalar_;
  using size_type = size_t;
  static constexpr size_type dim = Dim_;

  constexpr space_map(Scalar_* data, size_type size) noexcept : data_(data), size_(size), sdim_(Dim_) {}
  constexpr space_map(Scalar_* data, size_type size, size_type sdim) noexcept
      : data_(data), size_(size), sdim_(Dim_ == dynamic_extent ? sdim : Dim_) {}

  constexpr point_type operator[](size_type i) const noexcept { return point_type(data(i), sdim_); }
  constexpr Scalar_* data() const noexcept { return data_; }
  constexpr Scalar_* data(size_type i) const noexcept { return data_ + i * sdim_; }
  constexpr size_type size() const noexcept { return size_; }
  constexpr size_type sdim() const noexcept { return sdim_; }

 private:
  Scalar_* data_;
  size_type size_;
  size_type sdim_;
};

template <typename Scalar_, size_t Dim_>
struct point_traits<point_map<Scalar_, Dim_>> {
  using point_type = point_map<Scalar_, Dim_>;
  using scalar_type = typename point_type::scalar_type;
  using size_type = size_t;
  static constexpr size_type dim = Dim_;
  static scalar_type const* data(point_type const& p) { return p.data(); }
  static size_type size(point_type const& p) { return p.size(); }
};

template <typename Point_>
struct space_traits<space_map<Point_>> {
  using space_type = space_map<Point_>;
  using point_type = typename space_type::point_type;
  using scalar_type = typename space_type::scalar_type;
  using size_type = size_t;
  static constexpr size_type dim = space_type::dim;

  template <typename Index_>
  static decltype(auto) point_at(space_type const& s, Index_ i) {
    return s[static_cast<size_type>(i)];
  }
  static size_type size(space_type const& s) { return s.size(); }
  static size_type sdim(space_type const& s) { return s.sdim(); }
};

// ------------------------------------------------------------------ metrics
class topological_space_tag {};
class euclidean_space_tag : public topological_space_tag {};

// One-dimensional spaces a metric reports per dimension through apply_dim_space (distance.hpp:5-7 of the
// reference), and the distance helpers user-defined metrics are written with (distance.hpp:17-48).
class one_space_r1 {};
class one_space_s1 {};

template <typename Scalar_>
constexpr Scalar_ r1_distance(Scalar_ x, Scalar_ y) {
  return x > y ? x - y : y - x;
}
template <typename Scalar_>
constexpr Scalar_ s1_distance(Scalar_ x, Scalar_ y) {
  Scalar_ const d = r1_distance(x, y);
  return d < Scalar_(1.0) - d ? d : Scalar_(1.0) - d;
}
template <typename Scalar_>
constexpr Scalar_ squared(Scalar_ x) {
  return x * x;
}
template <typename Scalar_>
constexpr Scalar_ squared_r1_distance(Scalar_ x, Scalar_ y) {
  return squared(x - y);
}
template <typename Scalar_>
constexpr Scalar_ squared_s1_distance(Scalar_ x, Scalar_ y) {
  return squared(s1_distance(x, y));
}

namespace b200_detail {

template <typename It1_, typename End1_, typename It2_, typename Term_, typename Fold_, typename Scalar_>
constexpr Scalar_ fold_terms(It1_ a, End1_ a_end, It2_ b, Scalar_ init, Term_ term, Fold_ fold) {
  for (; a != a_end; ++a, ++b) init = fold(init, term(*a, *b));
  return init;
}

// distance on the unit circle [0, 1): the shorter way round (distance.hpp:19-22)
template <typename Scalar_>
constexpr Scalar_ circle_distance(Scalar_ x, Scalar_ y) {
  Scalar_ const d = x > y ? x - y : y - x;
  return d < Scalar_(1.0) - d ? d : Scalar_(1.0) - d;
}

}  // namespace b200_detail

// sum |x_i - y_i|
struct metric_l1 {
  using space_category = euclidean_space_tag;
  template <typename It1_, typename End1_, typename It2_>
  constexpr auto operator()(It1_ a, End1_ a_end, It2_ b) const {
    using S = typename std::iterator_traits<It1_>::value_type;
    return b200_detail::fold_terms(
        a, a_end, b, S(0), [](S x, S y) { return x > y ? x - y : y - x; }, [](S d, S t) { return d + t; });
  }
  template <typename Scalar_>
  constexpr Scalar_ operator()(Scalar_ x) const {
    return x < Scalar_(0) ? -x : x;
  }
};

// sum (x_i - y_i)^2, accumulated in dimension order
struct metric_l2_squared {
  using space_category = euclidean_space_tag;
  template <typename It1_, typename End1_, typename It2_>
  constexpr auto operator()(It1_ a, End1_ a_end, It2_ b) const {
    using S = typename std::iterator_traits<It1_>::value_type;
    return b200_detail::fold_terms(
        a, a_end, b, S(0), [](S x, S y) { return (x - y) * (x - y); }, [](S d, S t) { return d + t; });
  }
  template <typename Scalar_>
  constexpr Scalar_ operator()(Scalar_ x) const {
    return x * x;
  }
};

// max |x_i - y_i|
struct metric_lpinf {
  using space_category = euclidean_space_tag;
  template <typename It1_, typename End1_, typename It2_>
  constexpr auto operator()(It1_ a, End1_ a_end, It2_ b) const {
    using S = typename std::iterator_traits<It1_>::value_type;
    return b200_detail::fold_terms(
        a, a_end, b, S(0), [](S x, S y) { return x > y ? x - y : y - x; }, [](S d, S t) { return d < t ? t : d; });
  }
  template <typename Scalar_>
  constexpr Scalar_ operator()(Scalar_ x) const {
    return x < Scalar_(0) ? -x : x;
  }
};

// min |x_i - y_i|
struct metric_lninf {
  using space_category = euclidean_space_tag;
  template <typename It1_, typename End1_, typename It2_>
  constexpr auto operator()(It1_ a, End1_ a_end, It2_ b) const {
    using S = typename std::iterator_traits<It1_>::value_type;
    return b200_detail::fold_terms(
        a, a_end, b, std::numeric_limits<S>::max(), [](S x, S y) { return x > y ? x - y : y - x; },
        [](S d, S t) { return t < d ? t : d; });
  }
  template <typename Scalar_>
  constexpr Scalar_ operator()(Scalar_ x) const {
    return x < Scalar_(0) ? -x : x;
  }
};

// the circle S1 = [0, 1) with wrap-around
struct metric_so2 {
  using space_category = topological_space_tag;
  template <typename It1_, typename End1_, typename It2_>
  constexpr auto operator()(It1_ a, End1_, It2_ b) const {
    return b200_detail::circle_distance(*a, *b);
  }
  template <typename Scalar_>
  constexpr Scalar_ operator()(Scalar_ x) const {
    return x < Scalar_(0) ? -x : x;
  }
  template <typename UnaryPredicate_>
  void apply_dim_space(int, UnaryPredicate_ p) const {
    p(one_space_s1{});
  }
};

// R2 x S1: squared euclidean distance on (x, y) plus squared circle distance on the angle
struct metric_se2_squared {
  using space_category = topological_space_tag;
  template <typename It1_, typename End1_, typename It2_>
  constexpr auto operator()(It1_ a, End1_, It2_ b) const {
    using S = typename std::iterator_traits<It1_>::value_type;
    S d(0);
    for (int j = 0; j < 2; ++j, ++a, ++b) d += (*a - *b) * (*a - *b);
    S const c = b200_detail::circle_distance(*a, *b);
    return d + c * c;
  }
  template <typename Scalar_>
  constexpr Scalar_ operator()(Scalar_ x) const {
    return x * x;
  }
  template <typename UnaryPredicate_>
  void apply_dim_space(int dim, UnaryPredicate_ p) const {
    if (dim < 2)
      p(one_space_r1{});
    else
      p(one_space_s1{});
  }
};

// ------------------------------------------------------------------ build tags
// CRTP bases so that the kd_tree constructor can name "any rule / stop / bounds" in its
// signature (kd_tree.hpp:72-81 of the reference).
template <typename Derived_>
struct splitter_rule_t {
  Derived_ const& derived() const { return static_cast<Derived_ const&>(*this); }

 protected:
  constexpr splitter_rule_t() = default;
};
template <typename Derived_>
struct splitter_stop_condition_t {
  Derived_ const& derived() const { return static_cast<Derived_ const&>(*this); }

 protected:
  constexpr splitter_stop_condition_t() = default;
};
template <typename Derived_>
struct splitter_start_bounds_t {
  Derived_ const& derived() const { return static_cast<Derived_ const&>(*this); }

 protected:
  constexpr splitter_start_bounds_t() = default;
};

// split at the median of the longest side of the node's box
struct median_max_side_t : splitter_rule_t<median_max_side_t> {
  constexpr explicit median_max_side_t() = default;
};
// split at the middle of the longest side; children may be empty
struct midpoint_max_side_t : splitter_rule_t<midpoint_max_side_t> {
  constexpr explicit midpoint_max_side_t() = default;
};
// like midpoint, but the split slides to the nearest point when a side would be empty
struct sliding_midpoint_max_side_t : splitter_rule_t<sliding_midpoint_max_side_t> {
  constexpr explicit sliding_midpoint_max_side_t() = default;
};
inline constexpr median_max_side_t median_max_side{};
inline constexpr midpoint_max_side_t midpoint_max_side{};
inline constexpr sliding_midpoint_max_side_t sliding_midpoint_max_side{};

struct max_leaf_size_t : splitter_stop_condition_t<max_leaf_size_t> {
  constexpr max_leaf_size_t(size_t v) : value(v) { assert(value > 0); }
  size_t value;
};
struct max_leaf_depth_t : splitter_stop_condition_t<max_leaf_depth_t> {
  constexpr max_leaf_depth_t(size_t v) : value(v) {}
  size_t value;
};

struct bounds_from_space_t : splitter_start_bounds_t<bounds_from_space_t> {
  constexpr explicit bounds_from_space_t() = default;
};
inline constexpr bounds_from_space_t bounds_from_space{};

template <typename Point_>
struct bounds_t : splitter_start_bounds_t<bounds_t<Point_>> {
  constexpr explicit bounds_t(Point_ const& min, Point_ const& max) : min_(min), max_(max) {}
  constexpr Point_ const& min() const { return min_; }
  constexpr Point_ const& max() const { return max_; }

 private:
  Point_ min_;
  Point_ max_;
};

}  // namespace pico_tree
