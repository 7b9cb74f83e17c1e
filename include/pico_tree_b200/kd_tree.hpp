// kd_tree.hpp — pico_tree::kd_tree<Space_, Metric_, Index_> on top of libpico_b200.so.
//
// Source-compatible stand-in for the reference's class (src/pico_tree/pico_tree/kd_tree.hpp:19-435):
// same template parameters, member types, constructor tags, search_* overloads, leaf_ranges,
// space(), metric(), load / save, deduction guide and make_kd_tree. Where the reference runs
// internal::build_kd_tree and the internal::search_* recursions on the calling thread, this class
// makes ONE call into the C-ABI of include/pico_b200.h; the tree lives in HBM as a flat array of
// nodes next to leaf-ordered points. There is no CPU fallback: without a B200 every call throws
// std::runtime_error, and every search goes to the device — except where the CALLER'S OWN CODE has to run
// inside the search (a user-defined Metric_, a user-defined visitor passed to search_nearest) or where the
// caller opted in with b200::single_query_on_host(true): those walk a host mirror of the device-built tree
// (host_search.hpp).
//
// Additions the reference does not have (its Python binding loops over single queries,
// src/pyco_tree/pico_tree/_pyco_tree/kd_tree.hpp:117-268): search_nn_batch, search_knn_batch,
// search_radius_batch, search_box_batch — a whole query set per call, which is how the device is
// meant to be used. Single-query overloads are one-query batches (tens of microseconds of launch
// latency each): correct, but not the fast path.
//
// Vocabulary types: by default from "traits.hpp" next to this file. Define
// PICO_TREE_B200_USE_REFERENCE_TRAITS to take them from the reference's unmodified headers
// instead (pico_tree/core.hpp, map.hpp, metric.hpp, ... must then be on the include path); the
// Eigen / OpenCV adaptors of the reference then work unchanged. See INTEGRATION.md.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iterator>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../pico_b200.h"

#if defined(PICO_TREE_B200_USE_REFERENCE_TRAITS)
#include <pico_tree/core.hpp>
#include <pico_tree/internal/kd_tree_builder.hpp>  // build tags (max_leaf_size_t, bounds_t, rules)
#include <pico_tree/map.hpp>
#include <pico_tree/map_traits.hpp>
#include <pico_tree/metric.hpp>
#include <pico_tree/point_traits.hpp>
#include <pico_tree/space_traits.hpp>
#else
#include "traits.hpp"
#endif

#include "host_search.hpp"

namespace pico_tree {

namespace b200 {

// Device the next kd_tree is created on. Defaults to $PICO_B200_DEVICE or 0.
inline int& default_device() {
  static int device = [] {
    char const* e = std::getenv("PICO_B200_DEVICE");
    return e ? std::atoi(e) : 0;
  }();
  return device;
}

inline void check(int rc) {
  if (rc == PICO_B200_OK) return;
  std::string msg = std::string("pico_b200: ") + pico_b200_last_error();
  if (rc == PICO_B200_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

template <typename T_>
struct unwrap {
  using type = T_;
};
template <typename T_>
struct unwrap<std::reference_wrapper<T_>> {
  using type = std::remove_const_t<T_>;
};
template <typename T_>
using unwrap_t = typename unwrap<T_>::type;

template <typename Scalar_>
struct scalar_id;
template <>
struct scalar_id<float> : std::integral_constant<int, PICO_B200_F32> {};
template <>
struct scalar_id<double> : std::integral_constant<int, PICO_B200_F64> {};

// Re-encodes the index width of a reference tree stream (kd_tree_data.hpp:89-135): [sdim u64][n u64][n indices]
// [root box][nodes in pre-order: a leaf flag byte, then {begin, end} as indices for a leaf or a branch record of
// `branch_bytes`]. `consumed` receives the bytes of the INPUT stream that belong to the tree.
template <typename From_, typename To_>
inline std::vector<char> recode_stream_index(char const* src, std::size_t size, std::size_t branch_bytes,
                                             std::size_t scalar_bytes, std::uint64_t* consumed) {
  char const* p = src;
  char const* const end = src + size;
  std::vector<char> out;
  auto need = [&](std::size_t n) {
    if (static_cast<std::size_t>(end - p) < n) throw std::runtime_error("tree stream ends early");
  };
  auto copy = [&](std::size_t n) {
    need(n);
    out.insert(out.end(), p, p + n);
    p += n;
  };
  auto index = [&] {
    need(sizeof(From_));
    From_ v;
    std::memcpy(&v, p, sizeof(From_));
    p += sizeof(From_);
    To_ const w = static_cast<To_>(v);
    if (static_cast<From_>(w) != v || ((v < From_(0)) != (w < To_(0))))
      throw std::runtime_error("tree stream holds an index the other index type cannot represent");
    char b[sizeof(To_)];
    std::memcpy(b, &w, sizeof(To_));
    out.insert(out.end(), b, b + sizeof(To_));
  };
  need(16);
  std::uint64_t sdim = 0, n = 0;
  std::memcpy(&sdim, p, 8);
  std::memcpy(&n, p + 8, 8);
  copy(16);
  if (n > size / sizeof(From_)) throw std::runtime_error("tree stream ends early");
  out.reserve(size / sizeof(From_) * sizeof(To_) + 64);
  for (std::uint64_t i = 0; i < n; ++i) index();
  copy(2 * sdim * scalar_bytes);  // the root box: min, then max
  for (std::uint64_t pending = 1; pending > 0;) {
    need(1);
    bool const leaf = *p != 0;
    copy(1);
    if (leaf) {
      index();
      index();
      --pending;
    } else {
      copy(branch_bytes);
      ++pending;
    }
  }
  if (consumed) *consumed = static_cast<std::uint64_t>(p - src);
  return out;
}

// Metrics the device implements. A user-defined metric type has no device functor: the
// static_assert in kd_tree points here.
template <typename Metric_>
struct metric_id : std::integral_constant<int, -1> {};
template <>
struct metric_id<metric_l1> : std::integral_constant<int, PICO_B200_METRIC_L1> {};
template <>
struct metric_id<metric_l2_squared> : std::integral_constant<int, PICO_B200_METRIC_L2_SQUARED> {};
template <>
struct metric_id<metric_lpinf> : std::integral_constant<int, PICO_B200_METRIC_LPINF> {};
template <>
struct metric_id<metric_lninf> : std::integral_constant<int, PICO_B200_METRIC_LNINF> {};
template <>
struct metric_id<metric_so2> : std::integral_constant<int, PICO_B200_METRIC_SO2> {};
template <>
struct metric_id<metric_se2_squared> : std::integral_constant<int, PICO_B200_METRIC_SE2_SQUARED> {};

// metric id handed to pico_b200_tree_create: a user-defined metric builds the same tree (splits only look at
// coordinates); its category decides whether the nodes keep two bounds or four
template <typename Metric_>
constexpr int build_metric_id() {
  if constexpr (metric_id<Metric_>::value >= 0)
    return metric_id<Metric_>::value;
  else if constexpr (std::is_same_v<typename Metric_::space_category, euclidean_space_tag>)
    return PICO_B200_METRIC_CUSTOM_EUCLIDEAN;
  else
    return PICO_B200_METRIC_CUSTOM_TOPOLOGICAL;
}

template <typename Rule_>
struct rule_id;
template <>
struct rule_id<sliding_midpoint_max_side_t>
    : std::integral_constant<int, PICO_B200_RULE_SLIDING_MIDPOINT_MAX_SIDE> {};
template <>
struct rule_id<midpoint_max_side_t> : std::integral_constant<int, PICO_B200_RULE_MIDPOINT_MAX_SIDE> {};
template <>
struct rule_id<median_max_side_t> : std::integral_constant<int, PICO_B200_RULE_MEDIAN_MAX_SIDE> {};

inline std::pair<int, size_t> stop_of(max_leaf_size_t const& s) { return {PICO_B200_STOP_MAX_LEAF_SIZE, s.value}; }
inline std::pair<int, size_t> stop_of(max_leaf_depth_t const& s) { return {PICO_B200_STOP_MAX_LEAF_DEPTH, s.value}; }

// A space seen as rows of scalars: either the caller's memory (all points lie `stride` scalars
// apart, which is what space_map, std::vector<std::array>, Eigen and cv::Mat give) or, for a
// space whose points are scattered, a gathered copy.
template <typename Space_>
class rows_of {
  using space_type = unwrap_t<Space_>;
  using traits = space_traits<space_type>;
  using ptraits = point_traits<std::remove_cv_t<std::remove_reference_t<typename traits::point_type>>>;

 public:
  using scalar_type = typename traits::scalar_type;

  explicit rows_of(space_type const& s) : n_(traits::size(s)), sdim_(dim_of(s)), stride_(sdim_) {
    if (n_ == 0) return;
    scalar_type const* const p0 = at(s, 0);
    data_ = p0;
    bool regular = true;
    if (n_ > 1) {
      auto const a = reinterpret_cast<std::uintptr_t>(p0);
      auto const b = reinterpret_cast<std::uintptr_t>(at(s, 1));
      regular = b > a && (b - a) % sizeof(scalar_type) == 0 && (b - a) / sizeof(scalar_type) >= sdim_;
      if (regular) {
        stride_ = (b - a) / sizeof(scalar_type);
        if constexpr (one_block<space_type>::value) {
          // one allocation with a fixed pitch by construction (space_map, Eigen matrices, cv::Mat, vectors of
          // fixed-size points): the first, second and last point pin the layout — no O(n) walk per batch call
          regular = at(s, n_ - 1) == p0 + (n_ - 1) * stride_;
        } else {
          for (size_t i = 2; i < n_ && regular; ++i) regular = at(s, i) == p0 + i * stride_;
        }
      }
    }
    if (!regular) {
      copy_.resize(n_ * sdim_);
      for (size_t i = 0; i < n_; ++i) std::copy_n(at(s, i), sdim_, copy_.data() + i * sdim_);
      data_ = copy_.data();
      stride_ = sdim_;
    }
  }

  scalar_type const* data() const { return data_; }
  size_t size() const { return n_; }
  size_t sdim() const { return sdim_; }
  size_t stride() const { return stride_; }

  static size_t dim_of(space_type const& s) {
    if constexpr (traits::dim != dynamic_extent) {
      return traits::dim;
    } else {
      return traits::sdim(s);
    }
  }

 private:
  static scalar_type const* at(space_type const& s, size_t i) { return ptraits::data(traits::point_at(s, i)); }

  // Spaces whose points live in ONE block at a fixed pitch: anything with a data() member that yields scalars
  // (space_map, Eigen::Matrix / Eigen::Map) and std::vector of fixed-size points.
  template <typename S_, typename = void>
  struct one_block : std::false_type {};
  template <typename S_>
  struct one_block<S_, std::enable_if_t<std::is_convertible_v<decltype(std::declval<S_ const&>().data()),
                                                              scalar_type const*>>> : std::true_type {};
  template <typename P_, typename A_>
  struct one_block<std::vector<P_, A_>, std::enable_if_t<point_traits<P_>::dim != dynamic_extent &&
                                                          sizeof(P_) == point_traits<P_>::dim * sizeof(scalar_type)>>
      : std::true_type {};

  size_t n_;
  size_t sdim_;
  size_t stride_;
  scalar_type const* data_ = nullptr;
  std::vector<scalar_type> copy_;
};

template <typename Point_>
using point_traits_of = point_traits<std::remove_cv_t<std::remove_reference_t<Point_>>>;

struct tree_deleter {
  void operator()(pico_b200_tree* t) const { pico_b200_tree_destroy(t); }
};

// frees a buffer handed out by the library (ragged radius / box results)
struct lib_buffer {
  void* p = nullptr;
  ~lib_buffer() { pico_b200_free(p); }
};

}  // namespace b200

template <typename Space_, typename Metric_ = metric_l2_squared, typename Index_ = int>
class kd_tree {
  static_assert(std::is_same_v<std::remove_cv_t<Space_>, Space_>, "SPACE_TYPE_MUST_BE_NON-CONST_NON-VOLATILE");
  // metric_l1, metric_l2_squared, metric_lpinf, metric_lninf, metric_so2 and metric_se2_squared are searched on the
  // device; any other Metric_ is the caller's code and is searched over the host mirror (host_search.hpp)
  static constexpr bool device_metric = b200::metric_id<Metric_>::value >= 0;
  static_assert(std::is_integral_v<Index_>, "INDEX_NOT_AN_INTEGRAL_TYPE");

  using unwrapped_space = b200::unwrap_t<Space_>;
  using traits = space_traits<unwrapped_space>;
  using rows_type = b200::rows_of<Space_>;

  template <typename It_>
  class iterator_range {
   public:
    using iterator_type = It_;
    using difference_type = typename std::iterator_traits<It_>::difference_type;
    constexpr iterator_range(It_ b, It_ e) : begin_(b), end_(e) {}
    constexpr It_ begin() const { return begin_; }
    constexpr It_ end() const { return end_; }

   private:
    It_ begin_, end_;
  };

 public:
  using size_type = size_t;
  using index_type = Index_;
  using scalar_type = typename traits::scalar_type;
  static constexpr size_type dim = traits::dim;
  using space_type = Space_;
  using metric_type = Metric_;
  using neighbor_type = neighbor<index_type, scalar_type>;
  using leaf_range_type = iterator_range<typename std::vector<index_type>::const_iterator>;

  static_assert(std::is_same_v<scalar_type, float> || std::is_same_v<scalar_type, double>,
                "SCALAR_TYPE_MUST_BE_FLOAT_OR_DOUBLE");

  // Builds the tree on the device (pico_b200_tree_create). The space is taken by value like in
  // the reference: move it in, or pass a std::reference_wrapper, to avoid a host copy. The
  // device always gets its own copy of the coordinates.
  template <typename Stop_, typename Bounds_ = bounds_from_space_t, typename Rule_ = sliding_midpoint_max_side_t>
  kd_tree(space_type space, splitter_stop_condition_t<Stop_> const& stop_condition,
          splitter_start_bounds_t<Bounds_> const& start_bounds = Bounds_{},
          splitter_rule_t<Rule_> const& = Rule_{})
      : space_(std::move(space)), metric_() {
    rows_type rows(unwrapped());
    auto const stop = b200::stop_of(stop_condition.derived());
    scalar_type const* bmin = nullptr;
    scalar_type const* bmax = nullptr;
    if constexpr (!std::is_same_v<Bounds_, bounds_from_space_t>) {
      bmin = b200::point_traits_of<decltype(start_bounds.derived().min())>::data(start_bounds.derived().min());
      bmax = b200::point_traits_of<decltype(start_bounds.derived().max())>::data(start_bounds.derived().max());
    }
    pico_b200_tree* h = nullptr;
    b200::check(pico_b200_tree_create(rows.data(), rows.size(), rows.sdim(), rows.stride(),
                                      b200::scalar_id<scalar_type>::value, b200::build_metric_id<Metric_>(),
                                      b200::rule_id<Rule_>::value, stop.first, stop.second, bmin, bmax,
                                      b200::default_device(), &h));
    handle_.reset(h);
    n_ = rows.size();
    sdim_ = rows.sdim();
  }

  kd_tree(kd_tree const&) = delete;
  kd_tree(kd_tree&&) = default;
  kd_tree& operator=(kd_tree const&) = delete;
  kd_tree& operator=(kd_tree&&) = default;

  // ---------------------------------------------------------------- single query
  // The visitor is the caller's code: it is called for every point of every leaf the depth-first search reaches,
  // in the reference's order (kd_tree.hpp:106-120, internal/kd_tree_search.hpp:52-105 / :122-229), over the host
  // mirror of the device-built tree.
  template <typename P_, typename V_>
  void search_nearest(P_ const& x, V_& visitor) const {
    host_walk(query_data(x), visitor);
  }

  template <typename P_>
  void search_nn(P_ const& x, neighbor_type& nn) const {
    if (on_host()) {
      b200::visit_nn<neighbor_type> v(nn);
      host_walk(query_data(x), v);
      return;
    }
    knn_call(query_data(x), 1, sdim_, 1, 0.0, &nn);
  }

  // Approximate nearest neighbour: at most a factor e farther than the true one; the returned
  // distance is scaled by 1/e like the reference's (search_visitor.hpp:178-183).
  template <typename P_>
  void search_nn(P_ const& x, scalar_type const e, neighbor_type& nn) const {
    if (on_host()) {
      b200::visit_nn<neighbor_type> v(nn);
      b200::visit_scaled<b200::visit_nn<neighbor_type>> a(e, v);
      host_walk(query_data(x), a);
      return;
    }
    knn_call(query_data(x), 1, sdim_, 1, static_cast<double>(e), &nn);
  }

  template <typename P_, typename RandomAccessIterator_>
  void search_knn(P_ const& x, RandomAccessIterator_ begin, RandomAccessIterator_ end) const {
    knn_range(x, 0.0, begin, end);
  }

  template <typename P_>
  void search_knn(P_ const& x, size_type const k, std::vector<neighbor_type>& knn) const {
    knn.resize(std::min(k, n_));  // fewer points than k: all of them (kd_tree.hpp:190-195)
    if (on_host()) {
      host_knn(query_data(x), 0.0, knn.begin(), knn.end());
      return;
    }
    knn_call(query_data(x), 1, sdim_, knn.size(), 0.0, knn.data());
  }

  template <typename P_, typename RandomAccessIterator_>
  void search_knn(P_ const& x, scalar_type const e, RandomAccessIterator_ begin, RandomAccessIterator_ end) const {
    knn_range(x, static_cast<double>(e), begin, end);
  }

  template <typename P_>
  void search_knn(P_ const& x, size_type const k, scalar_type const e, std::vector<neighbor_type>& knn) const {
    knn.resize(std::min(k, n_));
    if (on_host()) {
      host_knn(query_data(x), static_cast<double>(e), knn.begin(), knn.end());
      return;
    }
    knn_call(query_data(x), 1, sdim_, knn.size(), static_cast<double>(e), knn.data());
  }

  template <typename P_>
  void search_radius(P_ const& x, scalar_type const radius, std::vector<neighbor_type>& n,
                     bool const sort = false) const {
    radius_one(query_data(x), radius, 0.0, n, sort);
  }

  template <typename P_>
  void search_radius(P_ const& x, scalar_type const radius, scalar_type const e, std::vector<neighbor_type>& n,
                     bool const sort = false) const {
    radius_one(query_data(x), radius, static_cast<double>(e), n, sort);
  }

  template <typename P_>
  void search_box(P_ const& min, P_ const& max, std::vector<index_type>& idxs) const {
    std::uint64_t offsets[2] = {0, 0};
    b200::lib_buffer out;
    std::int32_t* raw = nullptr;
    b200::check(pico_b200_box(handle_.get(), query_data(min), query_data(max), 1, sdim_, offsets, &raw, 0, nullptr));
    out.p = raw;
    idxs.assign(raw, raw + offsets[1]);
  }

  // ---------------------------------------------------------------- batches (one device call)
  // queries: any space (space_map, std::vector<std::array>, Eigen matrix, ...) of the tree's
  // scalar type and dimension. knn receives queries.size() rows of min(k, n) neighbours.
  template <typename Queries_>
  void search_knn_batch(Queries_ const& queries, size_type const k, std::vector<neighbor_type>& knn) const {
    knn_batch(queries, k, 0.0, knn);
  }

  template <typename Queries_>
  void search_knn_batch(Queries_ const& queries, size_type const k, scalar_type const e,
                        std::vector<neighbor_type>& knn) const {
    knn_batch(queries, k, static_cast<double>(e), knn);
  }

  template <typename Queries_>
  void search_nn_batch(Queries_ const& queries, std::vector<neighbor_type>& nns) const {
    knn_batch(queries, 1, 0.0, nns);
  }

  // Into caller memory: queries.size() * min(k, n) records (numpy arrays of the Python binding, pinned buffers).
  // e == 0: exact.
  template <typename Queries_>
  void search_knn_batch(Queries_ const& queries, size_type const k, neighbor_type* out,
                        scalar_type const e = scalar_type(0)) const {
    b200::rows_of<Queries_> rows(unwrap_queries(queries));
    check_queries(rows);
    if (k <= n_) {
      knn_rows(rows.data(), rows.size(), rows.stride(), k, static_cast<double>(e), out);
      return;
    }
    // rows of k slots but only n points: like the iterator overloads, the tail of every row keeps an
    // infinite distance (search_visitor.hpp:98-103)
    std::vector<neighbor_type> tmp(rows.size() * n_);
    knn_rows(rows.data(), rows.size(), rows.stride(), n_, static_cast<double>(e), tmp.data());
    for (size_type i = 0; i < rows.size(); ++i) {
      neighbor_type* row = std::copy_n(tmp.data() + i * n_, n_, out + i * k);
      std::fill(row, out + (i + 1) * k, neighbor_type(index_type(0), std::numeric_limits<scalar_type>::max()));
    }
  }

  // Ragged results: the neighbours of query i are flat[offsets[i] .. offsets[i + 1]).
  template <typename Queries_>
  void search_radius_batch(Queries_ const& queries, scalar_type const radius, std::vector<size_type>& offsets,
                           std::vector<neighbor_type>& flat, bool const sort = false,
                           scalar_type const e = scalar_type(0)) const {
    b200::rows_of<Queries_> rows(unwrap_queries(queries));
    check_queries(rows);
    if constexpr (!device_metric) {
      offsets.assign(1, 0);
      flat.clear();
      std::vector<neighbor_type> one;
      for (size_type i = 0; i < rows.size(); ++i) {
        radius_one(rows.data() + i * rows.stride(), radius, static_cast<double>(e), one, sort);
        flat.insert(flat.end(), one.begin(), one.end());
        offsets.push_back(flat.size());
      }
      return;
    }
    std::vector<std::uint64_t> offs(rows.size() + 1, 0);
    b200::lib_buffer out;
    b200::check(pico_b200_radius(handle_.get(), rows.data(), rows.size(), rows.stride(), static_cast<double>(radius),
                                 static_cast<double>(e), offs.data(), &out.p, sort ? unsigned(PICO_B200_SORT_RESULTS) : 0u,
                                 nullptr));
    offsets.assign(offs.begin(), offs.end());
    flat.resize(offs.back());
    take_neighbors(out.p, flat.size(), flat.data());
  }

  // Same, as one vector per query (what the reference's binding returns through its DArray).
  template <typename Queries_>
  void search_radius_batch(Queries_ const& queries, scalar_type const radius,
                           std::vector<std::vector<neighbor_type>>& nns, bool const sort = false,
                           scalar_type const e = scalar_type(0)) const {
    std::vector<size_type> offsets;
    std::vector<neighbor_type> flat;
    search_radius_batch(queries, radius, offsets, flat, sort, e);
    split_rows(offsets, flat, nns);
  }

  // Box i is [mins[i], maxs[i]] (inclusive); indices of box i are flat[offsets[i] .. offsets[i+1]).
  template <typename Queries_>
  void search_box_batch(Queries_ const& mins, Queries_ const& maxs, std::vector<size_type>& offsets,
                        std::vector<index_type>& flat) const {
    b200::rows_of<Queries_> lo(unwrap_queries(mins)), hi(unwrap_queries(maxs));
    check_queries(lo);
    check_queries(hi);
    if (lo.size() != hi.size()) throw std::invalid_argument("query min and max don't have equal size");
    if (lo.stride() != hi.stride()) throw std::invalid_argument("query min and max have different strides");
    std::vector<std::uint64_t> offs(lo.size() + 1, 0);
    b200::lib_buffer out;
    std::int32_t* raw = nullptr;
    b200::check(pico_b200_box(handle_.get(), lo.data(), hi.data(), lo.size(), lo.stride(), offs.data(), &raw, 0,
                              nullptr));
    out.p = raw;
    offsets.assign(offs.begin(), offs.end());
    flat.assign(raw, raw + offs.back());
  }

  // Same, as one vector per box.
  template <typename Queries_>
  void search_box_batch(Queries_ const& mins, Queries_ const& maxs, std::vector<std::vector<index_type>>& idxs) const {
    std::vector<size_type> offsets;
    std::vector<index_type> flat;
    search_box_batch(mins, maxs, offsets, flat);
    split_rows(offsets, flat, idxs);
  }

  // Boxes given as consecutive (min, max) rows of ONE space of 2 * n points — the layout of the reference's
  // Python binding (_pyco_tree/kd_tree.hpp:245-268).
  template <typename Queries_>
  void search_box_batch_pairs(Queries_ const& boxes, std::vector<std::vector<index_type>>& idxs) const {
    b200::rows_of<Queries_> rows(unwrap_queries(boxes));
    check_queries(rows);
    if (rows.size() % 2 != 0) throw std::invalid_argument("query min and max don't have equal size");
    size_type const nb = rows.size() / 2;
    std::vector<std::uint64_t> offs(nb + 1, 0);
    b200::lib_buffer out;
    std::int32_t* raw = nullptr;
    b200::check(pico_b200_box(handle_.get(), rows.data(), rows.data() + rows.stride(), nb, 2 * rows.stride(),
                              offs.data(), &raw, 0, nullptr));
    out.p = raw;
    std::vector<size_type> offsets(offs.begin(), offs.end());
    idxs.resize(nb);
    split_rows_raw(offsets, raw, idxs);
  }

  // ---------------------------------------------------------------- structure
  // Index ranges of all non-empty leaves in depth-first order (kd_tree.hpp:325). The ranges
  // point into a host copy of the index permutation fetched from the device on first use.
  std::vector<leaf_range_type> leaf_ranges() const {
    host_mirror const& m = mirror();
    std::vector<leaf_range_type> ranges;
    ranges.reserve(m.leaves.size());
    for (auto const& be : m.leaves)
      ranges.emplace_back(m.indices.begin() + be.first, m.indices.begin() + be.second);
    return ranges;
  }

  space_type const& space() const { return space_; }
  metric_type const& metric() const { return metric_; }

  // Shape of the device tree and the device time of its build.
  pico_b200_tree_info info() const {
    pico_b200_tree_info i;
    b200::check(pico_b200_tree_info_get(handle_.get(), &i));
    return i;
  }

  // The C handle, for callers that drive the C-ABI directly (device-resident batches, streams).
  pico_b200_tree const* native_handle() const { return handle_.get(); }

  // ---------------------------------------------------------------- load / save
  // Byte-compatible with the reference's stream (kd_tree.hpp:336-370): trees saved by either
  // library load in the other. The points are not stored.
  static kd_tree load(space_type space, std::string const& filename) {
    std::fstream stream(filename, std::ios::in | std::ios::binary);
    if (!stream.is_open()) throw std::runtime_error("unable to open file: " + filename);
    return load(std::move(space), stream);
  }

  static kd_tree load(space_type space, std::iostream& stream) {
    auto const pos = stream.tellg();
    std::vector<char> bytes((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
    std::uint64_t consumed = 0;
    if constexpr (sizeof(index_type) == 4) {
      kd_tree tree(std::move(space), bytes.data(), bytes.size(), &consumed);
      stream.clear();
      stream.seekg(pos + static_cast<std::streamoff>(consumed));
      return tree;
    } else {
      // the reference writes the permutation and the leaf ranges as Index_ (kd_tree_data.hpp:89-135); the device
      // keeps 32-bit indices, so a stream of another index width is re-encoded on the way in and out
      std::vector<char> const narrow =
          b200::recode_stream_index<index_type, std::int32_t>(bytes.data(), bytes.size(), branch_record_bytes(),
                                                              sizeof(scalar_type), &consumed);
      kd_tree tree(std::move(space), narrow.data(), narrow.size(), nullptr);
      stream.clear();
      stream.seekg(pos + static_cast<std::streamoff>(consumed));
      return tree;
    }
  }

  static void save(kd_tree const& tree, std::string const& filename) {
    std::fstream stream(filename, std::ios::out | std::ios::binary);
    if (!stream.is_open()) throw std::runtime_error("unable to open file: " + filename);
    save(tree, stream);
  }

  static void save(kd_tree const& tree, std::iostream& stream) {
    std::uint64_t bytes = 0;
    b200::check(pico_b200_tree_save_size(tree.handle_.get(), &bytes));
    std::vector<char> buf(bytes);
    b200::check(pico_b200_tree_save(tree.handle_.get(), buf.data()));
    if constexpr (sizeof(index_type) != 4)
      buf = b200::recode_stream_index<std::int32_t, index_type>(buf.data(), buf.size(), branch_record_bytes(),
                                                                sizeof(scalar_type), nullptr);
    stream.write(buf.data(), static_cast<std::streamsize>(buf.size()));
  }

 private:
  // bytes of one branch record of the reference's stream: kd_tree_branch_split (euclidean) or
  // kd_tree_branch_double (topological spaces) of kd_tree_node.hpp:28-60, written as the struct lies in memory
  static constexpr std::size_t branch_record_bytes() {
    struct split { int split_dim; scalar_type left_max, right_min; };
    struct twice { int split_dim; scalar_type left_min, left_max, right_min, right_max; };
    return std::is_same_v<typename Metric_::space_category, euclidean_space_tag> ? sizeof(split) : sizeof(twice);
  }

  struct host_mirror {
    b200::host_tree<scalar_type> tree;  // flat nodes, permutation, outer bounds (host_search.hpp)
    std::vector<index_type> indices;    // the permutation as index_type (leaf_ranges)
    std::vector<std::pair<std::ptrdiff_t, std::ptrdiff_t>> leaves;
  };
  struct lazy_mirror {
    std::once_flag once;
    host_mirror data;
  };

  kd_tree(space_type space, char const* bytes, size_type size, std::uint64_t* consumed)
      : space_(std::move(space)), metric_() {
    rows_type rows(unwrapped());
    pico_b200_tree* h = nullptr;
    b200::check(pico_b200_tree_load(rows.data(), rows.size(), rows.sdim(), rows.stride(),
                                    b200::scalar_id<scalar_type>::value, b200::build_metric_id<Metric_>(), bytes, size,
                                    b200::default_device(), &h, consumed));
    handle_.reset(h);
    n_ = rows.size();
    sdim_ = rows.sdim();
  }

  unwrapped_space const& unwrapped() const { return space_; }

  template <typename Q_>
  static b200::unwrap_t<Q_> const& unwrap_queries(Q_ const& q) {
    return q;
  }

  template <typename P_>
  scalar_type const* query_data(P_ const& x) const {
    using pt = b200::point_traits_of<P_>;
    static_assert(std::is_same_v<scalar_type, typename pt::scalar_type>, "POINT_AND_TREE_SCALAR_TYPES_DIFFER");
    static_assert(dim == pt::dim || dim == dynamic_extent || pt::dim == dynamic_extent, "POINT_AND_TREE_DIMS_DIFFER");
    if constexpr (pt::dim == dynamic_extent) {
      if (pt::size(x) != sdim_) throw std::invalid_argument("point and tree dimensions differ");
    }
    return pt::data(x);
  }

  template <typename Rows_>
  void check_queries(Rows_ const& rows) const {
    static_assert(std::is_same_v<scalar_type, typename Rows_::scalar_type>, "POINT_AND_TREE_SCALAR_TYPES_DIFFER");
    if (rows.size() != 0 && rows.sdim() != sdim_) throw std::invalid_argument("query and tree dimensions differ");
  }

  static constexpr bool abi_layout =
      sizeof(index_type) == 4 && sizeof(neighbor_type) == (sizeof(scalar_type) == 4 ? 8 : 16) &&
      offsetof(neighbor_type, distance) == sizeof(scalar_type);

  // library records {int32 index; Scalar distance} -> neighbor_type
  static void take_neighbors(void const* src, size_type count, neighbor_type* dst) {
    if constexpr (abi_layout) {
      if (count) std::memcpy(dst, src, count * sizeof(neighbor_type));
    } else {
      struct rec {
        std::int32_t index;
        scalar_type distance;
      };
      rec const* r = static_cast<rec const*>(src);
      for (size_type i = 0; i < count; ++i) dst[i] = neighbor_type(static_cast<index_type>(r[i].index), r[i].distance);
    }
  }

  // flat ragged result -> one vector per row (the filling is independent per row: OpenMP where it is enabled)
  template <typename Item_>
  static void split_rows(std::vector<size_type> const& offsets, std::vector<Item_> const& flat,
                         std::vector<std::vector<Item_>>& rows) {
    rows.resize(offsets.empty() ? 0 : offsets.size() - 1);
    split_rows_raw(offsets, flat.data(), rows);
  }
  template <typename Src_, typename Item_>
  static void split_rows_raw(std::vector<size_type> const& offsets, Src_ const* flat,
                             std::vector<std::vector<Item_>>& rows) {
    std::ptrdiff_t const n = static_cast<std::ptrdiff_t>(rows.size());
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (std::ptrdiff_t i = 0; i < n; ++i)
      rows[static_cast<size_type>(i)].assign(flat + offsets[static_cast<size_type>(i)],
                                             flat + offsets[static_cast<size_type>(i) + 1]);
  }

  void knn_call(scalar_type const* q, size_type nq, size_type stride, size_type k, double e,
                neighbor_type* out) const {
    if (nq == 0 || k == 0) return;
    if constexpr (abi_layout) {
      b200::check(pico_b200_knn(handle_.get(), q, nq, stride, k, e, out, 0, nullptr));
    } else {
      std::vector<unsigned char> tmp(nq * k * (sizeof(scalar_type) == 4 ? 8 : 16));
      b200::check(pico_b200_knn(handle_.get(), q, nq, stride, k, e, tmp.data(), 0, nullptr));
      take_neighbors(tmp.data(), nq * k, out);
    }
  }

  template <typename P_, typename It_>
  void knn_range(P_ const& x, double e, It_ begin, It_ end) const {
    static_assert(std::is_same_v<typename std::iterator_traits<It_>::value_type, neighbor_type>,
                  "ITERATOR_VALUE_TYPE_DOES_NOT_EQUAL_NEIGHBOR_TYPE");
    if (on_host()) {
      host_knn(query_data(x), e, begin, end);
      return;
    }
    size_type const k = static_cast<size_type>(std::distance(begin, end));
    size_type const kk = std::min(k, n_);
    std::vector<neighbor_type> tmp(kk);
    knn_call(query_data(x), 1, sdim_, kk, e, tmp.data());
    It_ it = std::copy(tmp.begin(), tmp.end(), begin);
    // range longer than the point set: the reference leaves the tail unspecified apart from
    // an infinite last distance (search_visitor.hpp:98-103)
    for (; it != end; ++it) *it = neighbor_type(index_type(0), std::numeric_limits<scalar_type>::max());
  }

  template <typename Queries_>
  void knn_batch(Queries_ const& queries, size_type k, double e, std::vector<neighbor_type>& knn) const {
    b200::rows_of<Queries_> rows(unwrap_queries(queries));
    check_queries(rows);
    k = std::min(k, n_);
    knn.resize(rows.size() * k);
    knn_rows(rows.data(), rows.size(), rows.stride(), k, e, knn.data());
  }

  // rows of queries -> rows of k neighbours: one device call, or (user-defined metric) the host descent per row
  void knn_rows(scalar_type const* q, size_type nq, size_type stride, size_type k, double e, neighbor_type* out) const {
    if constexpr (device_metric) {
      knn_call(q, nq, stride, k, e, out);
    } else {
      mirror();  // fetch once, outside the parallel loop
      std::ptrdiff_t const n = static_cast<std::ptrdiff_t>(nq);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 128)
#endif
      for (std::ptrdiff_t i = 0; i < n; ++i)
        host_knn(q + static_cast<size_type>(i) * stride, e, out + static_cast<size_type>(i) * k,
                 out + (static_cast<size_type>(i) + 1) * k);
    }
  }

  void radius_one(scalar_type const* q, scalar_type radius, double e, std::vector<neighbor_type>& n, bool sort) const {
    if (on_host()) {
      // search_approximate_radius compares against radius / e as well (search_visitor.hpp:261-267)
      b200::visit_radius<neighbor_type> v(e > 0 ? radius * (scalar_type(1.0) / static_cast<scalar_type>(e)) : radius, n);
      if (e > 0) {
        b200::visit_scaled<b200::visit_radius<neighbor_type>> a(static_cast<scalar_type>(e), v);
        host_walk(q, a);
      } else {
        host_walk(q, v);
      }
      if (sort) v.sort();
      return;
    }
    std::uint64_t offsets[2] = {0, 0};
    b200::lib_buffer out;
    b200::check(pico_b200_radius(handle_.get(), q, 1, sdim_, static_cast<double>(radius), e, offsets, &out.p,
                                 sort ? unsigned(PICO_B200_SORT_RESULTS) : 0u, nullptr));
    n.resize(offsets[1]);
    take_neighbors(out.p, n.size(), n.data());
  }

  host_mirror const& mirror() const {
    std::call_once(mirror_->once, [this] {
      pico_b200_tree_info const i = info();
      host_mirror& m = mirror_->data;
      m.tree.nodes.resize(i.n_nodes);
      m.tree.indices.resize(i.n_points);
      m.tree.root_box.resize(2 * sdim_);
      b200::check(pico_b200_tree_export(handle_.get(), m.tree.nodes.data(), m.tree.indices.data(),
                                        m.tree.root_box.data()));
      if constexpr (!std::is_same_v<typename Metric_::space_category, euclidean_space_tag>) {
        m.tree.outer.resize(2 * i.n_nodes);
        b200::check(pico_b200_tree_export_outer_bounds(handle_.get(), m.tree.outer.data()));
      }
      for (auto const& nd : m.tree.nodes)  // pre-order == depth-first order
        if (nd.split_dim == PICO_B200_LEAF && nd.b.end_idx > nd.a.begin_idx)
          m.leaves.emplace_back(static_cast<std::ptrdiff_t>(nd.a.begin_idx), static_cast<std::ptrdiff_t>(nd.b.end_idx));
      m.indices.assign(m.tree.indices.begin(), m.tree.indices.end());
    });
    return mirror_->data;
  }

  // ---- host descent (host_search.hpp): user-defined metric / visitor, or single queries on request
  static bool on_host() {
    if constexpr (device_metric)
      return b200::single_query_on_host_flag();
    else
      return true;
  }

  template <typename V_>
  void host_walk(scalar_type const* q, V_& visitor) const {
    host_mirror const& m = mirror();
    unwrapped_space const& s = unwrapped();
    auto point_of = [&s](size_type i) {
      return b200::point_traits_of<typename traits::point_type>::data(traits::point_at(s, i));
    };
    b200::host_descent<scalar_type, Metric_, decltype(point_of), V_>(m.tree, metric_, point_of, q, sdim_, visitor)();
  }

  template <typename It_>
  void host_knn(scalar_type const* q, double e, It_ begin, It_ end) const {
    static_assert(std::is_same_v<typename std::iterator_traits<It_>::value_type, neighbor_type>,
                  "ITERATOR_VALUE_TYPE_DOES_NOT_EQUAL_NEIGHBOR_TYPE");
    if (begin == end) return;
    b200::visit_knn<It_> v(begin, end);
    if (e > 0) {
      b200::visit_scaled<b200::visit_knn<It_>> a(static_cast<scalar_type>(e), v);
      host_walk(q, a);
    } else {
      host_walk(q, v);
    }
  }

  space_type space_;
  metric_type metric_;
  std::unique_ptr<pico_b200_tree, b200::tree_deleter> handle_;
  size_type n_ = 0;
  size_type sdim_ = 0;
  std::unique_ptr<lazy_mirror> mirror_ = std::make_unique<lazy_mirror>();
};

template <typename Space_, typename... Args>
kd_tree(Space_, Args...) -> kd_tree<Space_, metric_l2_squared, int>;

template <typename Metric_ = metric_l2_squared, typename Index_ = int, typename Bounds_ = bounds_from_space_t,
          typename Rule_ = sliding_midpoint_max_side_t, typename Space_, typename Stop_>
kd_tree<std::decay_t<Space_>, Metric_, Index_> make_kd_tree(
    Space_&& space, splitter_stop_condition_t<Stop_> const& stop_condition,
    splitter_start_bounds_t<Bounds_> const& start_bounds = Bounds_{}, splitter_rule_t<Rule_> const& rule = Rule_{}) {
  return kd_tree<std::decay_t<Space_>, Metric_, Index_>(std::forward<Space_>(space), stop_condition, start_bounds,
                                                        rule);
}

}  // namespace pico_tree
