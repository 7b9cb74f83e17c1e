// mini_test.hpp — a few gtest-shaped macros (gtest is not installed in this image), enough for
// tests/cpp/*.cpp to read like the reference's test/pico_tree/*.cpp.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace mini_test {

struct test_case {
  std::string name;
  std::function<void()> body;
};

inline std::vector<test_case>& registry() {
  static std::vector<test_case> r;
  return r;
}
inline int& failures() {
  static int f = 0;
  return f;
}
struct fatal_failure {};

struct registrar {
  registrar(char const* suite, char const* name, std::function<void()> body) {
    registry().push_back({std::string(suite) + "." + name, std::move(body)});
  }
};

template <typename A, typename B>
void report(char const* file, int line, char const* what, A const& a, B const& b) {
  ++failures();
  std::ostringstream s;
  s << file << ":" << line << ": Failure: " << what << " (" << a << " vs " << b << ")";
  std::cout << s.str() << std::endl;
}

// 4-ULP closeness like EXPECT_FLOAT_EQ / EXPECT_DOUBLE_EQ
template <typename T>
bool almost_equal(T a, T b) {
  if (a == b) return true;
  T const scale = std::max(std::fabs(a), std::fabs(b));
  return std::fabs(a - b) <= scale * 4 * std::numeric_limits<T>::epsilon();
}

inline int run_all(int argc, char** argv) {
  char const* filter = argc > 1 ? argv[1] : nullptr;
  int ran = 0, failed = 0;
  for (auto const& t : registry()) {
    if (filter && t.name.find(filter) == std::string::npos) continue;
    std::cout << "[ RUN      ] " << t.name << std::endl;
    int const before = failures();
    try {
      t.body();
    } catch (fatal_failure const&) {
    } catch (std::exception const& e) {
      ++failures();
      std::cout << "unexpected exception: " << e.what() << std::endl;
    }
    ++ran;
    if (failures() != before) {
      ++failed;
      std::cout << "[  FAILED  ] " << t.name << std::endl;
    } else {
      std::cout << "[       OK ] " << t.name << std::endl;
    }
  }
  std::cout << "[==========] " << ran << " tests ran, " << failed << " failed" << std::endl;
  return failed == 0 && ran > 0 ? 0 : 1;
}

}  // namespace mini_test

#define TEST(suite, name)                                                                  \
  static void suite##_##name##_body();                                                     \
  static mini_test::registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);  \
  static void suite##_##name##_body()

#define MT_CHECK(cond, what, a, b) \
  do {                             \
    if (!(cond)) mini_test::report(__FILE__, __LINE__, what, a, b); \
  } while (0)

#define EXPECT_TRUE(x) MT_CHECK((x), #x " is false", 0, 1)
#define EXPECT_FALSE(x) MT_CHECK(!(x), #x " is true", 1, 0)
#define EXPECT_EQ(a, b) MT_CHECK((a) == (b), #a " == " #b, (a), (b))
#define EXPECT_NE(a, b) MT_CHECK((a) != (b), #a " != " #b, (a), (b))
#define EXPECT_LE(a, b) MT_CHECK((a) <= (b), #a " <= " #b, (a), (b))
#define EXPECT_LT(a, b) MT_CHECK((a) < (b), #a " < " #b, (a), (b))
#define EXPECT_GE(a, b) MT_CHECK((a) >= (b), #a " >= " #b, (a), (b))
#define EXPECT_FLOAT_EQ(a, b) MT_CHECK(mini_test::almost_equal<float>((a), (b)), #a " ~= " #b, (a), (b))
#define EXPECT_DOUBLE_EQ(a, b) MT_CHECK(mini_test::almost_equal<double>((a), (b)), #a " ~= " #b, (a), (b))
#define ASSERT_EQ(a, b)                                            \
  do {                                                             \
    if (!((a) == (b))) {                                           \
      mini_test::report(__FILE__, __LINE__, #a " == " #b, (a), (b)); \
      throw mini_test::fatal_failure{};                            \
    }                                                              \
  } while (0)
#define EXPECT_THROW(stmt, ex)                                                    \
  do {                                                                            \
    bool caught__ = false;                                                        \
    try {                                                                         \
      stmt;                                                                       \
    } catch (ex const&) {                                                         \
      caught__ = true;                                                            \
    }                                                                             \
    MT_CHECK(caught__, #stmt " throws " #ex, 0, 1);                               \
  } while (0)
