// make_wide_index_streams.cpp — golden streams for the index-width handling of kd_tree::save / load
// (include/pico_tree_b200/kd_tree.hpp recode_stream_index). Compiled against the UNMODIFIED reference headers in
// the development container (tests/cpp/Makefile, target `golden-wide-index`); writes, for an euclidean and a
// topological metric, the stream the reference saves with Index_ = int and with Index_ = long over the same points:
//
//   tests/golden/wide_index/<name>.bin = [n u64][dim u64][n * dim floats][bytes_int u64][stream, int]
//                                        [bytes_long u64][stream, long]
#include <pico_tree/array_traits.hpp>
#include <pico_tree/kd_tree.hpp>
#include <pico_tree/vector_traits.hpp>

#include <array>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using point = std::array<float, 3>;

template <typename Metric_, typename Index_>
std::string stream_of(std::vector<point> const& pts) {
  using tree_type = pico_tree::kd_tree<std::reference_wrapper<std::vector<point> const>, Metric_, Index_>;
  tree_type tree(std::cref(pts), pico_tree::max_leaf_size_t(8));
  std::stringstream ss;
  tree_type::save(tree, ss);
  return ss.str();
}

template <typename Metric_>
void emit(std::string const& path, std::uint64_t seed) {
  std::vector<point> pts(3000);
  std::uint64_t x = seed;
  for (auto& p : pts)
    for (auto& c : p) {
      x = x * 6364136223846793005ull + 1442695040888963407ull;  // every coordinate in [0, 1): valid for SE(2) too
      c = static_cast<float>((x >> 40) & 0xffffff) / 16777216.0f;
    }
  std::ofstream out(path, std::ios::binary);
  auto u64 = [&](std::uint64_t v) { out.write(reinterpret_cast<char const*>(&v), 8); };
  u64(pts.size());
  u64(3);
  out.write(reinterpret_cast<char const*>(pts.data()), static_cast<std::streamsize>(pts.size() * sizeof(point)));
  for (std::string const& s : {stream_of<Metric_, int>(pts), stream_of<Metric_, long>(pts)}) {
    u64(s.size());
    out.write(s.data(), static_cast<std::streamsize>(s.size()));
  }
}

int main(int argc, char** argv) {
  std::string const dir = argc > 1 ? argv[1] : ".";
  emit<pico_tree::metric_l2_squared>(dir + "/l2.bin", 11);
  emit<pico_tree::metric_se2_squared>(dir + "/se2.bin", 12);
  return 0;
}
