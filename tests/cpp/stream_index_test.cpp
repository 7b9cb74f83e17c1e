// stream_index_test.cpp — CPU check of b200::recode_stream_index (include/pico_tree_b200/kd_tree.hpp): the golden
// files under tests/golden/wide_index hold, for the same points, the stream the UNMODIFIED reference saves with
// Index_ = int and with Index_ = long (tests/golden/make_wide_index_streams.cpp). Re-encoding one must give the
// other, byte for byte. No device call is made.
#include <pico_tree/array_traits.hpp>
#include <pico_tree/kd_tree.hpp>

#include <fstream>
#include <iterator>

#include "mini_test.hpp"

namespace {

struct golden {
  std::vector<float> points;
  std::vector<char> stream_int, stream_long;
};

golden read_golden(std::string const& name) {
  std::string const path = std::string(GOLDEN_DIR) + "/" + name;
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) throw std::runtime_error("missing " + path);
  auto u64 = [&] {
    std::uint64_t v = 0;
    in.read(reinterpret_cast<char*>(&v), 8);
    return v;
  };
  golden g;
  std::uint64_t const n = u64(), dim = u64();
  g.points.resize(n * dim);
  in.read(reinterpret_cast<char*>(g.points.data()), static_cast<std::streamsize>(g.points.size() * 4));
  g.stream_int.resize(u64());
  in.read(g.stream_int.data(), static_cast<std::streamsize>(g.stream_int.size()));
  g.stream_long.resize(u64());
  in.read(g.stream_long.data(), static_cast<std::streamsize>(g.stream_long.size()));
  return g;
}

void both_ways(std::string const& name, std::size_t branch_bytes) {
  golden const g = read_golden(name);
  EXPECT_TRUE(g.stream_long.size() > g.stream_int.size());
  std::uint64_t used = 0;
  auto const wide =
      pico_tree::b200::recode_stream_index<int, long>(g.stream_int.data(), g.stream_int.size(), branch_bytes, 4, &used);
  EXPECT_EQ(used, g.stream_int.size());
  EXPECT_TRUE(wide == g.stream_long);
  // with other content behind the tree, like a caller's own stream
  std::vector<char> padded = g.stream_long;
  padded.insert(padded.end(), {'t', 'a', 'i', 'l'});
  auto const narrow = pico_tree::b200::recode_stream_index<long, int>(padded.data(), padded.size(), branch_bytes, 4, &used);
  EXPECT_EQ(used, g.stream_long.size());
  EXPECT_TRUE(narrow == g.stream_int);
  // a stream that ends early is an error, not a read past the end
  EXPECT_THROW((pico_tree::b200::recode_stream_index<long, int>(g.stream_long.data(), g.stream_long.size() - 5,
                                                                 branch_bytes, 4, nullptr)),
               std::runtime_error);
  // an index the narrow type cannot hold
  std::vector<char> big = g.stream_long;
  long const huge = 1L << 40;
  std::memcpy(big.data() + 16, &huge, sizeof(long));
  EXPECT_THROW((pico_tree::b200::recode_stream_index<long, int>(big.data(), big.size(), branch_bytes, 4, nullptr)),
               std::runtime_error);
}

}  // namespace

TEST(StreamIndex, EuclideanBranchRecords) { both_ways("l2.bin", 12); }      // {int, float, float}
TEST(StreamIndex, TopologicalBranchRecords) { both_ways("se2.bin", 20); }   // {int, float x 4}

int main(int argc, char** argv) { return mini_test::run_all(argc, argv); }
