// bm_kd_tree.cpp — the reference's benchmark loops (examples/benchmark/bm_pico_kd_tree.cpp: BuildCtSldMid
// :29-36, KnnCtSldMid :63-78, RadiusCtSldMid :110-127) through pico_tree_b200's kd_tree.hpp, on a
// synthetic cloud (the reference loads scans*.bin, which are not available offline).
//
//   bm_kd_tree [n_tree] [n_query]      defaults 7733372 7200863 (the sizes of README.md:14-16)
//
// Prints build time, the batch entry points (one device call per batch, host memory in and out, pageable
// std::vector storage) and, for comparison, the reference-style loop of single-query calls on a sample.
#include <pico_tree/array_traits.hpp>
#include <pico_tree/kd_tree.hpp>
#include <pico_tree/vector_traits.hpp>

#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using point_type = std::array<float, 3>;
using space_type = std::reference_wrapper<std::vector<point_type>>;
using tree_type = pico_tree::kd_tree<space_type>;
using neighbor_type = tree_type::neighbor_type;

static double seconds_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// points on the walls / floor of a 50 x 50 x 8 m room with 1 cm noise, in a coherent order
static std::vector<point_type> make_cloud(std::size_t n, unsigned seed) {
  std::mt19937 gen(seed);
  std::uniform_real_distribution<float> u(0.0f, 1.0f);
  std::normal_distribution<float> noise(0.0f, 0.01f);
  std::vector<point_type> pts(n);
  for (std::size_t i = 0; i < n; ++i) {
    float const t = static_cast<float>(i) / static_cast<float>(n);
    float x = 50.0f * u(gen), y = 50.0f * t, z = 8.0f * u(gen);
    switch (i % 3) {
      case 0: z = 0.0f; break;
      case 1: x = (i % 2) ? 0.0f : 50.0f; break;
      default: break;
    }
    pts[i] = {x + noise(gen), y + noise(gen), z + noise(gen)};
  }
  return pts;
}

int main(int argc, char** argv) {
  std::size_t const n_tree = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 7733372;
  std::size_t const n_query = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 7200863;
  std::vector<point_type> points_tree = make_cloud(n_tree, 1), points_test = make_cloud(n_query, 2);

  // BuildCtSldMid
  auto t0 = std::chrono::steady_clock::now();
  tree_type tree(points_tree, pico_tree::max_leaf_size_t(10));
  double const first_build = seconds_since(t0);
  t0 = std::chrono::steady_clock::now();
  { tree_type again(points_tree, pico_tree::max_leaf_size_t(10)); }
  std::printf("build: first %.1f ms (creates the CUDA context), again %.1f ms wall, %.2f ms on the device, %llu nodes\n",
              first_build * 1e3, seconds_since(t0) * 1e3, tree.info().build_ms,
              static_cast<unsigned long long>(tree.info().n_nodes));

  // KnnCtSldMid as ONE batch call per k
  std::vector<neighbor_type> flat;
  for (std::size_t k : {1, 4, 8, 12}) {
    tree.search_knn_batch(points_test, k, flat);  // warm-up (also sizes `flat`)
    tree.search_knn_batch(points_test, k, flat);  // second big pageable batch: the pinned mirrors are allocated here
    t0 = std::chrono::steady_clock::now();
    tree.search_knn_batch(points_test, k, flat);
    double const s = seconds_since(t0);
    std::printf("knn k=%-2zu batch: %8.2f ms  %8.1f Mq/s\n", k, s * 1e3, static_cast<double>(n_query) / s / 1e6);
  }

  // RadiusCtSldMid (r = 0.1 m, squared): ragged batch
  {
    std::vector<std::size_t> offsets;
    std::vector<neighbor_type> hits;
    std::size_t const nq = std::min<std::size_t>(n_query, 1000000);
    std::vector<point_type> sample(points_test.begin(), points_test.begin() + static_cast<std::ptrdiff_t>(nq));
    tree.search_radius_batch(sample, 0.01f, offsets, hits);
    t0 = std::chrono::steady_clock::now();
    tree.search_radius_batch(sample, 0.01f, offsets, hits);
    double const s = seconds_since(t0);
    std::printf("radius r^2=0.01 batch of %zu: %8.2f ms  %8.1f Mq/s  %.1f hits/query\n", nq, s * 1e3,
                static_cast<double>(nq) / s / 1e6, static_cast<double>(hits.size()) / static_cast<double>(nq));
  }

  // the reference's loop shape (bm_pico_kd_tree.cpp:63-78: one search_knn per point, the result vector resized
  // per query), answered from the host mirror of the device-built tree: b200::single_query_on_host(true)
  {
    pico_tree::b200::single_query_on_host(true);
    std::size_t const nq = std::min<std::size_t>(n_query, 1000000);
    std::vector<neighbor_type> results;
    std::size_t sum = 0;
    tree.search_knn(points_test[0], 1, results);  // fetches the mirror (once per tree)
    for (std::size_t k : {1, 4, 8, 12}) {
      t0 = std::chrono::steady_clock::now();
      for (std::size_t i = 0; i < nq; ++i) {
        tree.search_knn(points_test[i], k, results);
        sum += results.size();
      }
      double const s = seconds_since(t0);
      std::printf("knn k=%-2zu single-query loop on the host mirror over %zu queries: %.3f us per call, %.2f Mq/s (1 thread)\n",
                  k, nq, s / nq * 1e6, static_cast<double>(nq) / s / 1e6);
    }
    pico_tree::b200::single_query_on_host(false);
    if (sum == 0) std::printf("?\n");
  }

  // the same loop with every call going to the device (the default): each a one-query device batch
  {
    std::size_t const nq = std::min<std::size_t>(n_query, 20000);
    std::vector<neighbor_type> results;
    std::size_t sum = 0;
    t0 = std::chrono::steady_clock::now();
    for (std::size_t i = 0; i < nq; ++i) {
      tree.search_knn(points_test[i], 1, results);
      sum += results.size();
    }
    double const s = seconds_since(t0);
    std::printf("knn k=1 single-query loop over %zu queries: %.1f us per call (%zu results)\n", nq, s / nq * 1e6, sum);
    // RadiusCtSldMid / box, one call per query
    sum = 0;
    t0 = std::chrono::steady_clock::now();
    for (std::size_t i = 0; i < nq; ++i) {
      tree.search_radius(points_test[i], 0.01f, results);
      sum += results.size();
    }
    double const sr = seconds_since(t0);
    std::printf("radius r^2=0.01 single-query loop: %.1f us per call (%.1f hits/query)\n", sr / nq * 1e6,
                static_cast<double>(sum) / static_cast<double>(nq));
    std::vector<int> idxs;
    sum = 0;
    t0 = std::chrono::steady_clock::now();
    for (std::size_t i = 0; i < nq; ++i) {
      point_type lo = points_test[i], hi = points_test[i];
      for (int d = 0; d < 3; ++d) {
        lo[d] -= 0.1f;
        hi[d] += 0.1f;
      }
      tree.search_box(lo, hi, idxs);
      sum += idxs.size();
    }
    double const sb = seconds_since(t0);
    std::printf("box 0.2 m edge single-query loop: %.1f us per call (%.1f hits/box)\n", sb / nq * 1e6,
                static_cast<double>(sum) / static_cast<double>(nq));
  }
  return 0;
}
