// kd_tree_test.cpp — the reference's KdTreeTest suite (test/pico_tree/kd_tree_test.cpp:60-178 with the
// helpers of test/pico_tree/common.hpp:54-215) restated against pico_tree_b200's kd_tree.hpp, plus
// cases for what the drop-in adds (batches) or must keep (tags, spaces, index types, visitors).
// Every search below runs on the GPU through libpico_b200.so; the expected values come from a brute
// force over the point set, like in the reference's own tests.
#include <pico_tree/array_traits.hpp>
#include <pico_tree/kd_tree.hpp>
#include <pico_tree/vector_traits.hpp>
#include <pico_understory/kd_forest.hpp>

#include <array>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <numeric>
#include <random>
#include <sstream>

#include "mini_test.hpp"

namespace {

using point_1f = std::array<float, 1>;
using point_2f = std::array<float, 2>;
using point_3f = std::array<float, 3>;
using point_3d = std::array<double, 3>;

template <typename Point_>
std::vector<Point_> generate_random_n(std::size_t n, typename Point_::value_type lo, typename Point_::value_type hi,
                                      unsigned seed = 42) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<typename Point_::value_type> dist(lo, hi);
  std::vector<Point_> pts(n);
  for (auto& p : pts)
    for (auto& c : p) c = dist(gen);
  return pts;
}

template <typename Point_>
std::vector<Point_> generate_random_n(std::size_t n, typename Point_::value_type size) {
  return generate_random_n<Point_>(n, 0, size);
}

template <typename Point_>
using space = std::reference_wrapper<std::vector<Point_>>;

template <typename Point_>
using kd_tree = pico_tree::kd_tree<space<Point_>>;

inline void float_eq(float a, float b) { EXPECT_FLOAT_EQ(a, b); }
inline void float_eq(double a, double b) { EXPECT_DOUBLE_EQ(a, b); }

// A view of point i of the tree's space as `sdim` scalars.
template <typename Tree_>
struct space_view {
  using scalar_type = typename Tree_::scalar_type;
  explicit space_view(Tree_ const& tree) : rows(tree.space()) {}
  scalar_type const* operator[](std::size_t i) const { return rows.data() + i * rows.stride(); }
  std::size_t size() const { return rows.size(); }
  std::size_t sdim() const { return rows.sdim(); }
  pico_tree::b200::rows_of<typename Tree_::space_type> rows;
};

// brute-force k nearest neighbours (common.hpp:54-80)
template <typename Tree_>
std::vector<typename Tree_::neighbor_type> brute_knn(Tree_ const& tree, typename Tree_::scalar_type const* p,
                                                     std::size_t k) {
  space_view<Tree_> sp(tree);
  std::vector<typename Tree_::neighbor_type> all(sp.size());
  for (std::size_t i = 0; i < sp.size(); ++i)
    all[i] = {static_cast<typename Tree_::index_type>(i), tree.metric()(p, p + sp.sdim(), sp[i])};
  k = std::min(k, all.size());
  std::partial_sort(all.begin(), all.begin() + static_cast<std::ptrdiff_t>(k), all.end());
  all.resize(k);
  return all;
}

// common.hpp:82-130
template <typename Tree_>
void test_box(Tree_ const& tree, typename Tree_::scalar_type min_v, typename Tree_::scalar_type max_v) {
  using scalar_type = typename Tree_::scalar_type;
  space_view<Tree_> sp(tree);
  std::vector<scalar_type> lo(sp.sdim(), min_v), hi(sp.sdim(), max_v);
  pico_tree::point_map<scalar_type const, Tree_::dim> pmin(lo.data(), lo.size()), pmax(hi.data(), hi.size());
  std::vector<typename Tree_::index_type> idxs;
  tree.search_box(pmin, pmax, idxs);
  auto inside = [&](scalar_type const* x) {
    for (std::size_t d = 0; d < sp.sdim(); ++d)
      if (x[d] < min_v || x[d] > max_v) return false;
    return true;
  };
  for (auto j : idxs) EXPECT_TRUE(inside(sp[static_cast<std::size_t>(j)]));
  std::size_t count = 0;
  for (std::size_t j = 0; j < sp.size(); ++j) count += inside(sp[j]);
  EXPECT_EQ(count, idxs.size());
}

// common.hpp:132-180
template <typename Tree_>
void test_radius(Tree_ const& tree, typename Tree_::scalar_type radius) {
  using scalar_type = typename Tree_::scalar_type;
  space_view<Tree_> sp(tree);
  pico_tree::point_map<scalar_type const, Tree_::dim> p(sp[sp.size() / 2], sp.sdim());
  auto const& metric = tree.metric();
  scalar_type const lp_radius = metric(radius);
  scalar_type const lp_scale = metric(scalar_type(1.5));
  std::vector<typename Tree_::neighbor_type> exact, apprx;
  tree.search_radius(p, lp_radius, exact);
  tree.search_radius(p, lp_radius, lp_scale, apprx);
  for (auto const& r : exact) {
    scalar_type const d = metric(p.data(), p.data() + p.size(), sp[static_cast<std::size_t>(r.index)]);
    EXPECT_LE(d, lp_radius);
    EXPECT_EQ(d, r.distance);
  }
  for (auto const& r : apprx) {
    scalar_type const d = metric(p.data(), p.data() + p.size(), sp[static_cast<std::size_t>(r.index)]);
    EXPECT_LE(d, lp_radius);
    float_eq(d, r.distance * lp_scale);
  }
  std::size_t count = 0;
  for (std::size_t j = 0; j < sp.size(); ++j) count += metric(p.data(), p.data() + p.size(), sp[j]) < lp_radius;
  EXPECT_EQ(count, exact.size());
  EXPECT_GE(count, apprx.size());
  // sorted variant: same set, ascending
  std::vector<typename Tree_::neighbor_type> sorted;
  tree.search_radius(p, lp_radius, sorted, true);
  EXPECT_EQ(sorted.size(), exact.size());
  EXPECT_TRUE(std::is_sorted(sorted.begin(), sorted.end()));
}

// common.hpp:182-215
template <typename Tree_, typename Point_>
void test_knn(Tree_ const& tree, std::size_t k, Point_ const& p) {
  using scalar_type = typename Tree_::scalar_type;
  scalar_type const lp_scale = tree.metric()(scalar_type(1.5));
  std::vector<typename Tree_::neighbor_type> exact, apprx;
  tree.search_knn(p, k, exact);
  tree.search_knn(p, k, lp_scale, apprx);
  auto const compare = brute_knn(tree, pico_tree::b200::point_traits_of<Point_>::data(p), k);
  ASSERT_EQ(compare.size(), exact.size());
  ASSERT_EQ(compare.size(), apprx.size());
  for (std::size_t i = 0; i < compare.size(); ++i) {
    // the device keeps the reference's operation order, so the brute force over the same
    // metric functor matches bit for bit; indices are not compared on equal distances
    EXPECT_EQ(exact[i].distance, compare[i].distance);
    EXPECT_LE(apprx[i].distance, exact[i].distance);
  }
}

template <typename Tree_>
void test_knn(Tree_ const& tree, std::size_t k) {
  space_view<Tree_> sp(tree);
  pico_tree::point_map<typename Tree_::scalar_type const, Tree_::dim> p(sp[sp.size() / 2], sp.sdim());
  test_knn(tree, k, p);
}

template <typename Point_>
void query_range(std::size_t n, typename Point_::value_type area, typename Point_::value_type min_v,
                 typename Point_::value_type max_v) {
  std::vector<Point_> random = generate_random_n<Point_>(n, area);
  kd_tree<Point_> tree(random, pico_tree::max_leaf_size_t(8));
  test_box(tree, min_v, max_v);
}

template <typename Point_>
void query_radius(std::size_t n, typename Point_::value_type area, typename Point_::value_type radius) {
  std::vector<Point_> random = generate_random_n<Point_>(n, area);
  kd_tree<Point_> tree(random, pico_tree::max_leaf_size_t(8));
  test_radius(tree, radius);
}

template <typename Point_>
void query_knn(std::size_t n, typename Point_::value_type area, std::size_t k) {
  std::vector<Point_> random = generate_random_n<Point_>(n, area);
  kd_tree<Point_> tree1(random, pico_tree::max_leaf_size_t(8));
  auto tree2 = std::move(tree1);  // move constructor
  tree1 = std::move(tree2);       // move assignment
  test_knn(tree1, k);
}

}  // namespace

// ------------------------------------------------------------------ the reference's suite
TEST(KdTreeTest, QueryRangeSubset2d) { query_range<point_2f>(1024 * 1024, 100.0f, 15.1f, 34.9f); }

TEST(KdTreeTest, QueryRangeAll2d) { query_range<point_2f>(1024, 10.0f, 0.0f, 10.0f); }

TEST(KdTreeTest, QueryRadiusSubset2d) { query_radius<point_2f>(1024 * 1024, 100.0f, 2.5f); }

TEST(KdTreeTest, QueryKnn1) { query_knn<point_2f>(1024 * 1024, 100.0f, 1); }

TEST(KdTreeTest, QueryKnn10) { query_knn<point_2f>(1024 * 1024, 100.0f, 10); }

TEST(KdTreeTest, QuerySo2Knn4) {
  using space_type = space<point_1f>;
  std::vector<point_1f> random = generate_random_n<point_1f>(256 * 256, 0.0f, 1.0f);
  pico_tree::kd_tree<space_type, pico_tree::metric_so2> tree(random, pico_tree::max_leaf_size_t(10));
  test_knn(tree, 8, point_1f{1.0f});
  test_knn(tree, 8, point_1f{0.02f});
  test_knn(tree, 1, point_1f{0.999f});
  // boxes on the circle: [0.90, 1.00] and the wrapping [0.95, 0.05] (kd_tree_test.cpp:95-97)
  for (auto const& mm : {std::pair<float, float>{0.90f, 1.00f}, std::pair<float, float>{0.95f, 0.05f}}) {
    point_1f lo{mm.first}, hi{mm.second};
    std::vector<int> idxs;
    tree.search_box(lo, hi, idxs);
    auto inside = [&](float x) {
      return mm.first <= mm.second ? (mm.first <= x && x <= mm.second) : (x >= mm.first || x <= mm.second);
    };
    std::size_t count = 0;
    for (auto const& p : random) count += inside(p[0]);
    for (int j : idxs) EXPECT_TRUE(inside(random[static_cast<std::size_t>(j)][0]));
    EXPECT_EQ(count, idxs.size());
    EXPECT_GE(count, 1000u);
  }
}

TEST(KdTreeTest, QuerySe2Knn) {
  using space_type = space<point_3f>;
  std::vector<point_3f> random = generate_random_n<point_3f>(100 * 1000, 0.0f, 1.0f);
  pico_tree::kd_tree<space_type, pico_tree::metric_se2_squared> tree(random, pico_tree::max_leaf_size_t(10));
  test_knn(tree, 8, point_3f{0.5f, 0.5f, 0.99f});
  test_knn(tree, 3, point_3f{0.1f, 0.9f, 0.01f});
  test_radius(tree, 0.05f);
}

TEST(KdTreeTest, WriteRead) {
  std::vector<point_2f> random = generate_random_n<point_2f>(100, 2.0f);
  std::string const filename = "tree_b200.bin";
  // compile time known dimensions
  {
    kd_tree<point_2f> tree(random, pico_tree::max_leaf_size_t(1));
    kd_tree<point_2f>::save(tree, filename);
  }
  {
    kd_tree<point_2f> tree = kd_tree<point_2f>::load(random, filename);
    test_knn(tree, 20);
  }
  EXPECT_EQ(std::remove(filename.c_str()), 0);
  // run time known dimensions
  using dyn_space = pico_tree::space_map<pico_tree::point_map<float const, pico_tree::dynamic_extent>>;
  dyn_space drandom(random.data()->data(), random.size(), 2);
  {
    static_assert(pico_tree::kd_tree<dyn_space>::dim == pico_tree::dynamic_extent, "KD_TREE_DIM_NOT_DYNAMIC");
    pico_tree::kd_tree<dyn_space> tree(drandom, pico_tree::max_leaf_size_t(1));
    pico_tree::kd_tree<dyn_space>::save(tree, filename);
  }
  {
    auto tree = pico_tree::kd_tree<dyn_space>::load(drandom, filename);
    test_knn(tree, 20);
  }
  EXPECT_EQ(std::remove(filename.c_str()), 0);
  // through a stream, with other content around the tree
  std::stringstream ss;
  ss.write("head", 4);
  {
    kd_tree<point_2f> tree(random, pico_tree::max_leaf_size_t(3));
    kd_tree<point_2f>::save(tree, ss);
  }
  ss.write("tail", 4);
  ss.seekg(4);
  kd_tree<point_2f> tree = kd_tree<point_2f>::load(random, ss);
  char tail[5] = {0, 0, 0, 0, 0};
  ss.read(tail, 4);
  EXPECT_EQ(std::string(tail), std::string("tail"));
  test_knn(tree, 5);
  EXPECT_THROW(kd_tree<point_2f>::load(random, "/nonexistent/dir/tree.bin"), std::runtime_error);
}

TEST(KdTreeTest, LeafRanges) {
  std::size_t const point_count = 100;
  std::vector<point_2f> random = generate_random_n<point_2f>(point_count, 2.0f);
  kd_tree<point_2f> tree(random, pico_tree::max_leaf_depth_t(2));
  auto leaf_ranges = tree.leaf_ranges();
  EXPECT_EQ(leaf_ranges.size(), 4u);
  if (!leaf_ranges.empty()) {
    std::ptrdiff_t sum_range_point_count = 0;
    int index_sum = 0;
    for (auto const& r : leaf_ranges) {
      sum_range_point_count += std::distance(r.begin(), r.end());
      for (int i : r) index_sum += i;
    }
    EXPECT_EQ(static_cast<std::size_t>(sum_range_point_count), point_count);
    EXPECT_EQ(static_cast<std::size_t>(index_sum), (point_count - 1) * point_count / 2);
    EXPECT_EQ(static_cast<std::size_t>(std::distance(leaf_ranges.front().begin(), leaf_ranges.back().end())),
              point_count);
  }
}

// ------------------------------------------------------------------ drop-in surface
TEST(KdTreeDropIn, SpaceMapAndDeductionGuide) {
  // BASELINE's space: kd_tree<space_map<point_map<float const, 3>>>
  std::vector<float> raw(3 * 50000);
  std::mt19937 gen(7);
  std::uniform_real_distribution<float> dist(0.0f, 1.0f);
  for (auto& v : raw) v = dist(gen);
  pico_tree::space_map<pico_tree::point_map<float const, 3>> sp(raw.data(), raw.size() / 3);
  pico_tree::kd_tree tree(sp, pico_tree::max_leaf_size_t(10));  // deduction guide
  static_assert(std::is_same_v<decltype(tree)::metric_type, pico_tree::metric_l2_squared>);
  static_assert(std::is_same_v<decltype(tree)::index_type, int>);
  static_assert(decltype(tree)::dim == 3);
  test_knn(tree, 1);
  test_knn(tree, 16);
  test_radius(tree, 0.05f);
  test_box(tree, 0.25f, 0.5f);
  pico_tree::neighbor<int, float> nn;
  float const q[3] = {0.5f, 0.5f, 0.5f};
  tree.search_nn(q, nn);  // Scalar[Dim] points (array_traits)
  EXPECT_EQ(nn.distance, brute_knn(tree, q, 1)[0].distance);
  tree.search_nn(q, 2.0f, nn);
  EXPECT_LE(nn.distance, brute_knn(tree, q, 1)[0].distance);
}

TEST(KdTreeDropIn, OwnedSpaceMoveAndMakeKdTree) {
  std::vector<point_3d> pts = generate_random_n<point_3d>(20000, 1.0);
  auto copy = pts;
  auto tree = pico_tree::make_kd_tree<pico_tree::metric_l1>(std::move(copy), pico_tree::max_leaf_size_t(12));
  static_assert(std::is_same_v<decltype(tree)::scalar_type, double>);
  static_assert(sizeof(decltype(tree)::neighbor_type) == 16);
  EXPECT_EQ(tree.space().size(), pts.size());
  EXPECT_EQ(tree.metric()(-2.0), 2.0);
  test_knn(tree, 7);
  test_radius(tree, 0.1);
  test_box(tree, 0.4, 0.6);
}

TEST(KdTreeDropIn, RulesBoundsAndStopConditions) {
  std::vector<point_2f> pts = generate_random_n<point_2f>(30000, 1.0f);
  pico_tree::kd_tree<space<point_2f>, pico_tree::metric_lpinf> t1(pts, pico_tree::max_leaf_size_t(5),
                                                                  pico_tree::bounds_from_space,
                                                                  pico_tree::median_max_side);
  test_knn(t1, 9);
  pico_tree::kd_tree<space<point_2f>> t2(pts, pico_tree::max_leaf_depth_t(9),
                                         pico_tree::bounds_t<point_2f>({-1.0f, -1.0f}, {2.0f, 2.0f}),
                                         pico_tree::midpoint_max_side);
  test_knn(t2, 9);
  test_box(t2, 0.1f, 0.3f);
  EXPECT_LE(t2.info().height, 9u);
  pico_tree::kd_tree<space<point_2f>> t3(pts, pico_tree::max_leaf_size_t(1), pico_tree::bounds_from_space,
                                         pico_tree::sliding_midpoint_max_side);
  EXPECT_EQ(t3.info().n_leaves, pts.size());
  test_knn(t3, 3);
}

TEST(KdTreeDropIn, IteratorOverloadsAndWideIndex) {
  std::vector<point_3f> pts = generate_random_n<point_3f>(5000, 1.0f);
  pico_tree::kd_tree<space<point_3f>, pico_tree::metric_l2_squared, long> tree(pts, pico_tree::max_leaf_size_t(10));
  using neighbor_type = decltype(tree)::neighbor_type;
  static_assert(std::is_same_v<neighbor_type::index_type, long>);
  point_3f q{0.3f, 0.6f, 0.9f};
  std::array<neighbor_type, 6> out;
  tree.search_knn(q, out.begin(), out.end());
  auto want = brute_knn(tree, q.data(), 6);
  for (std::size_t i = 0; i < 6; ++i) EXPECT_EQ(out[i].distance, want[i].distance);
  EXPECT_EQ(out[0].index, want[0].index);
  tree.search_knn(q, 2.25f, out.begin(), out.end());
  EXPECT_LE(out[0].distance, want[0].distance);
  // k larger than the point set returns every point (kd_tree.hpp:190-195)
  std::vector<neighbor_type> all;
  tree.search_knn(q, 6000, all);
  EXPECT_EQ(all.size(), pts.size());
  EXPECT_TRUE(std::is_sorted(all.begin(), all.end()));
}

// save / load with an index type that is not 32 bits wide: the stream must be the one the reference writes for that
// Index_ (golden files made by tests/golden/make_wide_index_streams.cpp from the unmodified reference headers), and a
// stream written by the reference must load.
namespace {
template <typename Metric_>
void wide_index_round_trip(char const* file) {
  std::ifstream in(std::string(GOLDEN_DIR) + "/" + file, std::ios::binary);
  ASSERT_EQ(in.is_open(), true);
  auto u64 = [&] {
    std::uint64_t v = 0;
    in.read(reinterpret_cast<char*>(&v), 8);
    return v;
  };
  std::uint64_t const n = u64(), dim = u64();
  ASSERT_EQ(dim, std::uint64_t(3));
  std::vector<point_3f> pts(n);
  in.read(reinterpret_cast<char*>(pts.data()), static_cast<std::streamsize>(n * sizeof(point_3f)));
  std::string s_int(u64(), '\0');
  in.read(s_int.data(), static_cast<std::streamsize>(s_int.size()));
  std::string s_long(u64(), '\0');
  in.read(s_long.data(), static_cast<std::streamsize>(s_long.size()));

  using tree_long = pico_tree::kd_tree<space<point_3f>, Metric_, long>;
  using tree_int = pico_tree::kd_tree<space<point_3f>, Metric_, int>;
  tree_long built(pts, pico_tree::max_leaf_size_t(8));
  std::stringstream ours;
  tree_long::save(built, ours);
  EXPECT_TRUE(ours.str() == s_long);
  std::stringstream ours_int;
  tree_int built_int(pts, pico_tree::max_leaf_size_t(8));
  tree_int::save(built_int, ours_int);
  EXPECT_TRUE(ours_int.str() == s_int);

  std::stringstream theirs;
  theirs.write("xy", 2);
  theirs.write(s_long.data(), static_cast<std::streamsize>(s_long.size()));
  theirs.write("z", 1);
  theirs.seekg(2);
  tree_long loaded = tree_long::load(pts, theirs);
  char z = 0;
  theirs.read(&z, 1);
  EXPECT_EQ(z, 'z');
  std::vector<typename tree_long::neighbor_type> a, b;
  for (std::size_t i = 0; i < 50; ++i) {
    built.search_knn(pts[i * 7], 5, a);
    loaded.search_knn(pts[i * 7], 5, b);
    ASSERT_EQ(a.size(), b.size());
    for (std::size_t j = 0; j < a.size(); ++j) {
      EXPECT_EQ(a[j].index, b[j].index);
      EXPECT_EQ(a[j].distance, b[j].distance);
    }
  }
}
}  // namespace

TEST(KdTreeDropIn, SaveLoadWithWideIndexMatchesReferenceStream) {
  wide_index_round_trip<pico_tree::metric_l2_squared>("l2.bin");
  wide_index_round_trip<pico_tree::metric_se2_squared>("se2.bin");
}

namespace {
// examples/kd_tree/kd_tree_custom_search_visitor.cpp:10-45 — keeps the nearest point only
template <typename Neighbor_>
struct search_nn_counter {
  explicit search_nn_counter(Neighbor_& nn) : nn_(nn) { nn_.distance = std::numeric_limits<float>::max(); }
  void operator()(int i, float d) {
    ++count;
    if (nn_.distance > d) nn_ = {i, d};
  }
  float max() const { return nn_.distance; }
  std::size_t count = 0;
  Neighbor_& nn_;
};
// collects everything closer than a fixed bound
struct within {
  float bound;
  std::vector<int> hits;
  void operator()(int i, float d) {
    if (d < bound) hits.push_back(i);
  }
  float max() const { return bound; }
};
}  // namespace

TEST(KdTreeDropIn, CustomVisitor) {
  std::vector<point_3f> pts = generate_random_n<point_3f>(40000, 1.0f);
  kd_tree<point_3f> tree(pts, pico_tree::max_leaf_size_t(10));
  point_3f q{0.5f, 0.5f, 0.5f};
  pico_tree::neighbor<int, float> nn;
  search_nn_counter<pico_tree::neighbor<int, float>> v(nn);
  tree.search_nearest(q, v);
  auto want = brute_knn(tree, q.data(), 1);
  EXPECT_EQ(nn.index, want[0].index);
  EXPECT_EQ(nn.distance, want[0].distance);
  EXPECT_GE(v.count, 1u);
  within w{0.01f, {}};
  tree.search_nearest(q, w);
  std::vector<pico_tree::neighbor<int, float>> rad;
  tree.search_radius(q, 0.01f, rad);
  EXPECT_EQ(w.hits.size(), rad.size());
  EXPECT_GE(w.hits.size(), 100u);
  // the visitor sees the points in the reference's visit order: exactly the unsorted search_radius sequence
  for (std::size_t i = 0; i < rad.size() && i < w.hits.size(); ++i) EXPECT_EQ(w.hits[i], rad[i].index);
}

// ---------------------------------------------------------------- user-defined metrics (host_search.hpp)
// examples/kd_tree/kd_tree_custom_metric.cpp of the reference: the L_p^p metric (euclidean) and the squared distance
// on the torus S1 x S1 (topological). The tree is built on the device, the searches call the user's functors.
namespace {
template <std::size_t P_>
struct metric_lp_p {
  using space_category = pico_tree::euclidean_space_tag;
  template <typename It_>
  auto operator()(It_ b1, It_ e1, It_ b2) const {
    using scalar = typename std::iterator_traits<It_>::value_type;
    scalar d{};
    for (; b1 != e1; ++b1, ++b2) d += operator()(*b1 - *b2);
    return d;
  }
  template <typename S_>
  S_ operator()(S_ x) const {
    return std::pow(std::abs(x), static_cast<S_>(P_));
  }
};

struct metric_t2_squared {
  using space_category = pico_tree::topological_space_tag;
  template <typename It_>
  auto operator()(It_ b1, It_ e1, It_ b2) const {
    using scalar = typename std::iterator_traits<It_>::value_type;
    scalar d{};
    for (; b1 != e1; ++b1, ++b2) d += pico_tree::squared_s1_distance(*b1, *b2);
    return d;
  }
  template <typename S_>
  S_ operator()(S_ x) const {
    return x * x;
  }
  template <typename P_>
  void apply_dim_space(int, P_ p) const {
    p(pico_tree::one_space_s1{});
  }
};

template <typename Tree_, typename Metric_>
std::vector<typename Tree_::neighbor_type> brute_metric(Tree_ const& tree, Metric_ const& m, float const* q,
                                                         std::size_t k) {
  space_view<Tree_> sv(tree);
  std::vector<typename Tree_::neighbor_type> all;
  for (std::size_t i = 0; i < sv.size(); ++i)
    all.emplace_back(static_cast<int>(i), m(q, q + sv.sdim(), sv[i]));
  std::stable_sort(all.begin(), all.end(), [](auto const& a, auto const& b) { return a.distance < b.distance; });
  all.resize(std::min(k, all.size()));
  return all;
}
}  // namespace

TEST(KdTreeDropIn, UserDefinedEuclideanMetric) {
  std::vector<point_2f> pts = generate_random_n<point_2f>(50000, 10.0f);
  pico_tree::kd_tree<space<point_2f>, metric_lp_p<3>> tree(pts, pico_tree::max_leaf_size_t(12));
  auto queries = generate_random_n<point_2f>(200, -1.0f, 11.0f, 5);
  for (auto const& q : queries) {
    pico_tree::neighbor<int, float> nn;
    tree.search_nn(q, nn);
    auto want = brute_metric(tree, tree.metric(), q.data(), 6);
    EXPECT_EQ(nn.index, want[0].index);
    EXPECT_EQ(nn.distance, want[0].distance);
    std::vector<pico_tree::neighbor<int, float>> knn;
    tree.search_knn(q, 6, knn);
    EXPECT_EQ(knn.size(), 6u);
    for (std::size_t i = 0; i < knn.size(); ++i) EXPECT_EQ(knn[i].distance, want[i].distance);
    std::vector<pico_tree::neighbor<int, float>> rad;
    tree.search_radius(q, want[5].distance, rad, true);  // strictly inside the 6th distance
    EXPECT_TRUE(rad.size() <= 5u);
  }
  // batches loop over the host descent; box search needs no metric and runs on the device
  std::vector<pico_tree::neighbor<int, float>> flat;
  tree.search_knn_batch(queries, 3, flat);
  EXPECT_EQ(flat.size(), queries.size() * 3);
  for (std::size_t i = 0; i < queries.size(); ++i) {
    auto want = brute_metric(tree, tree.metric(), queries[i].data(), 3);
    for (std::size_t j = 0; j < 3; ++j) EXPECT_EQ(flat[i * 3 + j].distance, want[j].distance);
  }
  test_box(tree, 2.0f, 4.0f);
  // save / load keeps working (the stream does not depend on the metric)
  std::stringstream ss;
  pico_tree::kd_tree<space<point_2f>, metric_lp_p<3>>::save(tree, ss);
  auto loaded = pico_tree::kd_tree<space<point_2f>, metric_lp_p<3>>::load(pts, ss);
  pico_tree::neighbor<int, float> a, b;
  tree.search_nn(queries[0], a);
  loaded.search_nn(queries[0], b);
  EXPECT_EQ(a.index, b.index);
}

TEST(KdTreeDropIn, UserDefinedTopologicalMetric) {
  std::vector<point_2f> pts = generate_random_n<point_2f>(30000, 1.0f);
  pico_tree::kd_tree<space<point_2f>, metric_t2_squared> tree(pts, pico_tree::max_leaf_size_t(12));
  auto queries = generate_random_n<point_2f>(200, 0.0f, 1.0f, 6);
  queries.push_back({1.0f, 1.0f});  // the corner: neighbours wrap around in both dimensions
  queries.push_back({0.0f, 0.5f});
  for (auto const& q : queries) {
    std::array<pico_tree::neighbor<int, float>, 8> knn;
    tree.search_knn(q, knn.begin(), knn.end());
    auto want = brute_metric(tree, tree.metric(), q.data(), 8);
    for (std::size_t i = 0; i < 8; ++i) EXPECT_EQ(knn[i].distance, want[i].distance);
  }
}

// One query at a time on the host mirror when the caller asks for it: same tree, same arithmetic order, so the
// answers are the device's bit for bit (ties included).
TEST(KdTreeDropIn, SingleQueriesOnHostWhenAskedFor) {
  std::vector<point_3f> pts = generate_random_n<point_3f>(80000, 1.0f);
  for (auto& p : pts) p[2] = std::round(p[2] * 16.0f) / 16.0f;  // planes: plenty of equal distances
  kd_tree<point_3f> tree(pts, pico_tree::max_leaf_size_t(10));
  auto queries = generate_random_n<point_3f>(500, -0.2f, 1.2f, 11);
  using neighbor_type = kd_tree<point_3f>::neighbor_type;
  std::vector<neighbor_type> dev_knn, host_knn, dev_aknn, host_aknn, dev_rad, host_rad;
  for (auto const& q : queries) {
    neighbor_type dev_nn, host_nn;
    pico_tree::b200::single_query_on_host(false);
    tree.search_nn(q, dev_nn);
    tree.search_knn(q, 7, dev_knn);
    tree.search_knn(q, 7, 1.5f, dev_aknn);
    tree.search_radius(q, 0.002f, dev_rad);
    pico_tree::b200::single_query_on_host(true);
    tree.search_nn(q, host_nn);
    tree.search_knn(q, 7, host_knn);
    tree.search_knn(q, 7, 1.5f, host_aknn);
    tree.search_radius(q, 0.002f, host_rad);
    pico_tree::b200::single_query_on_host(false);
    EXPECT_EQ(dev_nn.index, host_nn.index);
    EXPECT_EQ(dev_nn.distance, host_nn.distance);
    EXPECT_EQ(dev_rad.size(), host_rad.size());
    for (std::size_t i = 0; i < 7; ++i) {
      EXPECT_EQ(dev_knn[i].index, host_knn[i].index);
      EXPECT_EQ(dev_knn[i].distance, host_knn[i].distance);
      EXPECT_EQ(dev_aknn[i].index, host_aknn[i].index);
      EXPECT_EQ(dev_aknn[i].distance, host_aknn[i].distance);
    }
    for (std::size_t i = 0; i < dev_rad.size() && i < host_rad.size(); ++i) EXPECT_EQ(dev_rad[i].index, host_rad[i].index);
  }
}

TEST(KdTreeDropIn, BatchesEqualSingles) {
  std::vector<point_3f> pts = generate_random_n<point_3f>(60000, 1.0f);
  std::vector<point_3f> queries = generate_random_n<point_3f>(3000, -0.1f, 1.1f, 99);
  kd_tree<point_3f> tree(pts, pico_tree::max_leaf_size_t(10));
  using neighbor_type = kd_tree<point_3f>::neighbor_type;
  std::vector<neighbor_type> flat, one;
  tree.search_knn_batch(queries, 5, flat);
  ASSERT_EQ(flat.size(), queries.size() * 5);
  for (std::size_t i = 0; i < queries.size(); i += 37) {
    tree.search_knn(queries[i], 5, one);
    for (std::size_t j = 0; j < 5; ++j) {
      EXPECT_EQ(flat[i * 5 + j].index, one[j].index);
      EXPECT_EQ(flat[i * 5 + j].distance, one[j].distance);
    }
  }
  std::vector<neighbor_type> nns;
  tree.search_nn_batch(queries, nns);
  ASSERT_EQ(nns.size(), queries.size());
  for (std::size_t i = 0; i < queries.size(); ++i) EXPECT_EQ(nns[i].index, flat[i * 5].index);
  // a strided query set: every other point of a space_map
  std::vector<std::vector<neighbor_type>> ragged;
  tree.search_radius_batch(queries, 0.002f, ragged, true);
  ASSERT_EQ(ragged.size(), queries.size());
  for (std::size_t i = 0; i < queries.size(); i += 41) {
    tree.search_radius(queries[i], 0.002f, one, true);
    ASSERT_EQ(ragged[i].size(), one.size());
    for (std::size_t j = 0; j < one.size(); ++j) EXPECT_EQ(ragged[i][j].distance, one[j].distance);
  }
  std::vector<point_3f> mins(queries), maxs(queries);
  for (auto& p : mins)
    for (auto& c : p) c -= 0.03f;
  for (auto& p : maxs)
    for (auto& c : p) c += 0.03f;
  std::vector<std::size_t> offsets;
  std::vector<int> idx, idx_one;
  tree.search_box_batch(mins, maxs, offsets, idx);
  ASSERT_EQ(offsets.size(), queries.size() + 1);
  for (std::size_t i = 0; i < queries.size(); i += 53) {
    tree.search_box(mins[i], maxs[i], idx_one);
    ASSERT_EQ(offsets[i + 1] - offsets[i], idx_one.size());
    EXPECT_TRUE(std::equal(idx_one.begin(), idx_one.end(), idx.begin() + static_cast<std::ptrdiff_t>(offsets[i])));
  }
}

namespace {
// a space whose points do not lie at a regular stride: the tree gathers a copy
struct scattered_space {
  std::vector<std::unique_ptr<point_2f>> pts;
};
}  // namespace
namespace pico_tree {
template <>
struct space_traits<scattered_space> {
  using space_type = scattered_space;
  using point_type = point_2f;
  using scalar_type = float;
  using size_type = size_t;
  static constexpr size_type dim = 2;
  template <typename Index_>
  static point_type const& point_at(space_type const& s, Index_ i) {
    return *s.pts[static_cast<size_type>(i)];
  }
  static size_type size(space_type const& s) { return s.pts.size(); }
  static constexpr size_type sdim(space_type const&) { return dim; }
};
}  // namespace pico_tree

TEST(KdTreeDropIn, CustomSpaceTraits) {
  scattered_space s;
  auto pts = generate_random_n<point_2f>(5000, 1.0f);
  std::vector<std::unique_ptr<char[]>> gaps;
  for (auto const& p : pts) {
    s.pts.push_back(std::make_unique<point_2f>(p));
    gaps.push_back(std::make_unique<char[]>(1 + (s.pts.size() % 7)));  // irregular addresses
  }
  pico_tree::kd_tree<std::reference_wrapper<scattered_space>> tree(s, pico_tree::max_leaf_size_t(4));
  test_knn(tree, 4);
  test_box(tree, 0.2f, 0.4f);
}

// ---------------------------------------------------------------- kd_forest (examples/pico_understory)
// examples/kd_forest/kd_forest.cpp:20-41 drives the forest with search_nn per test point and counts how often the
// exact neighbour is found; with an unbounded leaf budget every tree is searched exhaustively, so the answer must
// be the exact one (distances are measured in the reflected spaces: equal up to rounding).
TEST(KdForest, UnboundedBudgetIsExactAndBatchesEqualSingles) {
  using point_8f = std::array<float, 8>;
  auto pts = generate_random_n<point_8f>(20000, 1.0f);
  auto queries = generate_random_n<point_8f>(300, 0.0f, 1.0f, 7);
  pico_tree::kd_tree<space<point_8f>> tree(pts, pico_tree::max_leaf_size_t(10));
  pico_tree::kd_forest<space<point_8f>> forest(pts, 10, 4);
  EXPECT_EQ(forest.info().n_trees, 4u);
  auto const rot = forest.rotations();
  EXPECT_EQ(rot.size(), 4u * 8u);
  for (std::size_t t = 0; t < 4; ++t) {
    float s = 0;
    for (std::size_t j = 0; j < 8; ++j) s += rot[t * 8 + j] * rot[t * 8 + j];
    EXPECT_TRUE(std::abs(s - 1.0f) < 1e-4f);
  }
  std::vector<pico_tree::neighbor<int, float>> batch;
  forest.search_nn_batch(queries, std::size_t(1) << 40, batch);
  EXPECT_EQ(batch.size(), queries.size());
  std::size_t found_small = 0;
  for (std::size_t i = 0; i < queries.size(); ++i) {
    pico_tree::neighbor<int, float> exact, nn, few;
    tree.search_nn(queries[i], exact);
    forest.search_nn(queries[i], std::size_t(1) << 40, nn);
    EXPECT_EQ(nn.index, exact.index);
    EXPECT_TRUE(std::abs(nn.distance - exact.distance) <= 1e-5f * (1.0f + exact.distance));
    EXPECT_EQ(batch[i].index, nn.index);
    float_eq(batch[i].distance, nn.distance);
    forest.search_nn(queries[i], 8, few);
    found_small += few.index == exact.index;
  }
  EXPECT_TRUE(found_small > queries.size() / 2);  // 4 trees x 8 leaves already find most neighbours in 8-D
  // the same forest again from its own reflection vectors: reproducible
  pico_tree::kd_forest<space<point_8f>> again(pts, 10, 4, rot);
  std::vector<pico_tree::neighbor<int, float>> a, b;
  forest.search_knn_batch(queries, 5, 16, a);
  again.search_knn_batch(queries, 5, 16, b);
  EXPECT_EQ(a.size(), b.size());
  for (std::size_t i = 0; i < a.size(); ++i) {
    EXPECT_EQ(a[i].index, b[i].index);
    float_eq(a[i].distance, b[i].distance);
  }
  auto moved = pico_tree::make_kd_forest(std::ref(pts), 10, 2);
  pico_tree::neighbor<int, float> nn;
  moved.search_nn(queries[0], 4, nn);
  EXPECT_TRUE(nn.index >= 0);
}

int main(int argc, char** argv) { return mini_test::run_all(argc, argv); }
