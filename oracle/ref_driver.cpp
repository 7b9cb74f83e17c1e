// ref_driver.cpp — C entry points around the UNMODIFIED reference headers
// (/root/reference/src/pico_tree), compiled into oracle/_ref/libpico_ref.so by
// oracle/Makefile. TEST INFRASTRUCTURE ONLY: it validates the C restatement in
// pico_oracle.c, generates tests/golden/ fixtures and serves as the
// "reference" CPU baseline of bench.py. No reference source is copied here;
// this file only *uses* the public API (kd_tree.hpp:72-88,125-318,336-370).
//
// Flags (oracle/Makefile): -std=c++17 -O3 -DNDEBUG -ffp-contract=off -fopenmp.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <pico_tree/kd_tree.hpp>
#include <pico_tree/map_traits.hpp>

namespace {

enum { METRIC_L1 = 0, METRIC_L2SQ = 1, METRIC_LPINF = 2, METRIC_LNINF = 3, METRIC_SO2 = 4, METRIC_SE2SQ = 5 };
enum { RULE_SLIDING = 0, RULE_MIDPOINT = 1, RULE_MEDIAN = 2 };
enum { STOP_SIZE = 0, STOP_DEPTH = 1 };

struct tree_base {
  virtual ~tree_base() = default;
  virtual std::string save() const = 0;
  virtual void knn(void const* q, size_t nq, size_t k, double e, void* out, int threads) const = 0;
  virtual void radius(
      void const* q, size_t nq, double radius, double e, int sort, uint64_t* offsets,
      void** out, int threads) const = 0;
  virtual void box(
      void const* mins, void const* maxs, size_t nb, uint64_t* offsets, int32_t** out,
      int threads) const = 0;
};

template <typename T, size_t Dim, typename Metric>
struct tree_impl final : tree_base {
  using space_type = pico_tree::space_map<pico_tree::point_map<T const, Dim>>;
  using point_type = pico_tree::point_map<T const, Dim>;
  using tree_type = pico_tree::kd_tree<space_type, Metric, int>;
  using neighbor_type = typename tree_type::neighbor_type;

  template <typename Stop, typename Rule>
  static tree_type make2(
      space_type s, size_t sdim, Stop stop, Rule rule, T const* bmin, T const* bmax) {
    if (bmin && bmax) {
      return tree_type(
          s, stop, pico_tree::bounds_t<point_type>(point_type(bmin, sdim), point_type(bmax, sdim)),
          rule);
    }
    return tree_type(s, stop, pico_tree::bounds_from_space, rule);
  }

  template <typename Stop>
  static tree_type make1(
      space_type s, size_t sdim, Stop stop, int rule, T const* bmin, T const* bmax) {
    switch (rule) {
      case RULE_MIDPOINT:
        return make2(s, sdim, stop, pico_tree::midpoint_max_side, bmin, bmax);
      case RULE_MEDIAN:
        return make2(s, sdim, stop, pico_tree::median_max_side, bmin, bmax);
      default:
        return make2(s, sdim, stop, pico_tree::sliding_midpoint_max_side, bmin, bmax);
    }
  }

  static tree_type make(
      space_type s, size_t sdim, int rule, int stop_kind, size_t stop_value, T const* bmin,
      T const* bmax) {
    if (stop_kind == STOP_DEPTH) {
      return make1(s, sdim, pico_tree::max_leaf_depth_t(stop_value), rule, bmin, bmax);
    }
    return make1(s, sdim, pico_tree::max_leaf_size_t(stop_value), rule, bmin, bmax);
  }

  tree_impl(
      T const* pts, size_t n, size_t sdim, int rule, int stop_kind, size_t stop_value,
      T const* bmin, T const* bmax)
      : sdim_(sdim),
        tree_(make(space_type(pts, n, sdim), sdim, rule, stop_kind, stop_value, bmin, bmax)) {}

  std::string save() const override {
    std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
    tree_type::save(tree_, ss);
    return ss.str();
  }

  // Batch loop exactly like _pyco_tree/kd_tree.hpp:117-135 (threads>1) or
  // examples/benchmark/bm_pico_kd_tree.cpp:70-77 (threads==1, serial).
  void knn(void const* qv, size_t nq, size_t k, double e, void* outv, int threads) const override {
    T const* q = static_cast<T const*>(qv);
    neighbor_type* out = static_cast<neighbor_type*>(outv);
    auto one = [&](size_t i) {
      point_type p(q + i * sdim_, sdim_);
      if (e > 0) {
        tree_.search_knn(p, T(e), out + i * k, out + (i + 1) * k);
      } else {
        tree_.search_knn(p, out + i * k, out + (i + 1) * k);
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

  // The reference's binding runs these loops under OpenMP (dynamic,128) into per-query vectors
  // (_pyco_tree/kd_tree.hpp:158-167,193-199,260-267); here each thread keeps the results of a block of
  // queries and the blocks are concatenated in query order afterwards.
  template <typename Item, typename One>
  static void ragged(size_t n, int threads, uint64_t* offsets, void** outv, One&& one) {
    constexpr size_t kBlock = 1024;
    size_t const nblocks = (n + kBlock - 1) / kBlock;
    std::vector<std::vector<Item>> parts(nblocks);
    std::vector<uint32_t> counts(n);
    std::ptrdiff_t const cnt = static_cast<std::ptrdiff_t>(nblocks);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
    for (std::ptrdiff_t b = 0; b < cnt; ++b) {
      std::vector<Item> tmp;
      auto& part = parts[size_t(b)];
      size_t const end = std::min(n, (size_t(b) + 1) * kBlock);
      for (size_t i = size_t(b) * kBlock; i < end; ++i) {
        one(i, tmp);
        counts[i] = static_cast<uint32_t>(tmp.size());
        part.insert(part.end(), tmp.begin(), tmp.end());
      }
    }
    offsets[0] = 0;
    for (size_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + counts[i];
    auto* buf = static_cast<Item*>(std::malloc(sizeof(Item) * (offsets[n] + 1)));
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
    for (std::ptrdiff_t b = 0; b < cnt; ++b) {
      auto const& part = parts[size_t(b)];
      if (!part.empty())
        std::memcpy(buf + offsets[size_t(b) * kBlock], part.data(), sizeof(Item) * part.size());
    }
    *outv = buf;
  }

  void radius(
      void const* qv, size_t nq, double radius, double e, int sort, uint64_t* offsets,
      void** outv, int threads) const override {
    T const* q = static_cast<T const*>(qv);
    ragged<neighbor_type>(nq, threads, offsets, outv, [&](size_t i, std::vector<neighbor_type>& one) {
      point_type p(q + i * sdim_, sdim_);
      if (e > 0) {
        tree_.search_radius(p, T(radius), T(e), one, sort != 0);
      } else {
        tree_.search_radius(p, T(radius), one, sort != 0);
      }
    });
  }

  void box(void const* mins, void const* maxs, size_t nb, uint64_t* offsets, int32_t** outv, int threads)
      const override {
    T const* mn = static_cast<T const*>(mins);
    T const* mx = static_cast<T const*>(maxs);
    void* out = nullptr;
    ragged<int>(nb, threads, offsets, &out, [&](size_t i, std::vector<int>& one) {
      tree_.search_box(point_type(mn + i * sdim_, sdim_), point_type(mx + i * sdim_, sdim_), one);
    });
    *outv = static_cast<int32_t*>(out);
  }

  size_t sdim_;
  tree_type tree_;
};

template <typename T, size_t Dim>
tree_base* make_metric(
    T const* pts, size_t n, size_t sdim, int metric, int rule, int stop_kind, size_t stop_value,
    T const* bmin, T const* bmax) {
  switch (metric) {
    case METRIC_L1:
      return new tree_impl<T, Dim, pico_tree::metric_l1>(
          pts, n, sdim, rule, stop_kind, stop_value, bmin, bmax);
    case METRIC_LPINF:
      return new tree_impl<T, Dim, pico_tree::metric_lpinf>(
          pts, n, sdim, rule, stop_kind, stop_value, bmin, bmax);
    case METRIC_LNINF:
      return new tree_impl<T, Dim, pico_tree::metric_lninf>(
          pts, n, sdim, rule, stop_kind, stop_value, bmin, bmax);
    case METRIC_SO2:  // topological spaces: search_nearest_topological, kd_tree_search.hpp:122-229
      return new tree_impl<T, Dim, pico_tree::metric_so2>(
          pts, n, sdim, rule, stop_kind, stop_value, bmin, bmax);
    case METRIC_SE2SQ:
      return new tree_impl<T, Dim, pico_tree::metric_se2_squared>(
          pts, n, sdim, rule, stop_kind, stop_value, bmin, bmax);
    default:
      return new tree_impl<T, Dim, pico_tree::metric_l2_squared>(
          pts, n, sdim, rule, stop_kind, stop_value, bmin, bmax);
  }
}

template <typename T>
tree_base* make_dim(
    T const* pts, size_t n, size_t sdim, int force_dynamic, int metric, int rule, int stop_kind,
    size_t stop_value, T const* bmin, T const* bmax) {
  if (!force_dynamic && sdim == 2) {
    return make_metric<T, 2>(pts, n, sdim, metric, rule, stop_kind, stop_value, bmin, bmax);
  }
  if (!force_dynamic && sdim == 3) {
    return make_metric<T, 3>(pts, n, sdim, metric, rule, stop_kind, stop_value, bmin, bmax);
  }
  return make_metric<T, pico_tree::dynamic_extent>(
      pts, n, sdim, metric, rule, stop_kind, stop_value, bmin, bmax);
}

}  // namespace

extern "C" {

// scalar: 0 = float, 1 = double. Points are borrowed (row-major, packed) and
// must outlive the handle, like the reference's space_map.
void* ref_build(
    void const* pts, size_t n, size_t sdim, int scalar, int force_dynamic, int metric, int rule,
    int stop_kind, size_t stop_value, void const* bmin, void const* bmax) {
  if (scalar == 1) {
    return make_dim<double>(
        static_cast<double const*>(pts), n, sdim, force_dynamic, metric, rule, stop_kind,
        stop_value, static_cast<double const*>(bmin), static_cast<double const*>(bmax));
  }
  return make_dim<float>(
      static_cast<float const*>(pts), n, sdim, force_dynamic, metric, rule, stop_kind, stop_value,
      static_cast<float const*>(bmin), static_cast<float const*>(bmax));
}

void ref_free(void* h) { delete static_cast<tree_base*>(h); }

// kd_tree::save stream (kd_tree_data.hpp:43-58,89-135). Returns malloc'd bytes.
void* ref_save(void const* h, size_t* size) {
  std::string s = static_cast<tree_base const*>(h)->save();
  void* buf = std::malloc(s.size() + 1);
  std::memcpy(buf, s.data(), s.size());
  *size = s.size();
  return buf;
}

void ref_knn(void const* h, void const* q, size_t nq, size_t k, double e, void* out, int threads) {
  static_cast<tree_base const*>(h)->knn(q, nq, k, e, out, threads);
}

void ref_radius(
    void const* h, void const* q, size_t nq, double radius, double e, int sort, uint64_t* offsets,
    void** out) {
  static_cast<tree_base const*>(h)->radius(q, nq, radius, e, sort, offsets, out, 1);
}

void ref_box(
    void const* h, void const* mins, void const* maxs, size_t nb, uint64_t* offsets,
    int32_t** out) {
  static_cast<tree_base const*>(h)->box(mins, maxs, nb, offsets, out, 1);
}

// the same loops on `threads` host threads (results still in query order)
void ref_radius_mt(
    void const* h, void const* q, size_t nq, double radius, double e, int sort, uint64_t* offsets,
    void** out, int threads) {
  static_cast<tree_base const*>(h)->radius(q, nq, radius, e, sort, offsets, out, threads);
}

void ref_box_mt(
    void const* h, void const* mins, void const* maxs, size_t nb, uint64_t* offsets, int32_t** out,
    int threads) {
  static_cast<tree_base const*>(h)->box(mins, maxs, nb, offsets, out, threads);
}

void ref_free_buffer(void* p) { std::free(p); }
}
