// Stand-in for <gtest/gtest.h> -- TEST INFRASTRUCTURE ONLY (see ../Eigen/Dense).  GoogleTest is not in the image; this is
// the subset the reference's se_core unit tests use (TEST, TEST_F with SetUp/TearDown, ASSERT_/EXPECT_ EQ NE TRUE FALSE LT
// with streamed messages), so that those tests compile UNMODIFIED from /root/reference/se_core/test and run against the
// reference's headers built on the stand-in Eigen.  Output mimics gtest's summary lines; exit code 1 on any failure.
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

class Test {
 public:
  virtual ~Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

struct Registry {
  struct Entry { std::string name; std::function<Test*()> make; };
  std::vector<Entry> tests;
  int failures_in_current = 0;
  static Registry& get() { static Registry r; return r; }
};
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<Test*()> make) { Registry::get().tests.push_back({std::string(suite) + "." + name, make}); }
};

// collects the `<< message` tail of a failed assertion and reports when it goes out of scope
struct Failure {
  std::ostringstream msg;
  const char* file; int line; std::string what;
  Failure(const char* f, int l, std::string w) : file(f), line(l), what(std::move(w)) {}
  ~Failure() { std::printf("%s:%d: Failure\n%s\n%s\n", file, line, what.c_str(), msg.str().c_str()); ++Registry::get().failures_in_current; }
  template <class T> Failure& operator<<(const T& v) { msg << v; return *this; }
};
struct Voidify { void operator&(const Failure&) {} };      // lets `return Voidify() & Failure(...) << ...;` type-check in a void function

inline void InitGoogleTest(int*, char**) {}
inline int RunAllTests() {
  Registry& r = Registry::get();
  int failed = 0;
  std::printf("[==========] Running %zu tests.\n", r.tests.size());
  for (auto& e : r.tests) {
    std::printf("[ RUN      ] %s\n", e.name.c_str());
    r.failures_in_current = 0;
    Test* t = e.make();
    t->SetUp(); t->TestBody(); t->TearDown();
    delete t;
    if (r.failures_in_current) { ++failed; std::printf("[  FAILED  ] %s\n", e.name.c_str()); }
    else std::printf("[       OK ] %s\n", e.name.c_str());
  }
  std::printf("[==========] %zu tests ran.\n[  PASSED  ] %zu tests.\n", r.tests.size(), r.tests.size() - failed);
  if (failed) std::printf("[  FAILED  ] %d tests.\n", failed);
  return failed ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()

#define SE_GTEST_DEFINE(suite, name, base)                                                                      \
  class suite##_##name##_Test : public base { public: void TestBody() override; };                              \
  static ::testing::Registrar suite##_##name##_registrar(#suite, #name, [] { return (::testing::Test*)new suite##_##name##_Test; }); \
  void suite##_##name##_Test::TestBody()
#define TEST(suite, name) SE_GTEST_DEFINE(suite, name, ::testing::Test)
#define TEST_F(fixture, name) SE_GTEST_DEFINE(fixture, name, fixture)

#define SE_GTEST_CHECK(cond, text, fatal)                                          \
  if (cond) ; else fatal ::testing::Voidify() & ::testing::Failure(__FILE__, __LINE__, text)
#define SE_GTEST_FATAL return
#define SE_GTEST_NONFATAL

#define ASSERT_TRUE(c) SE_GTEST_CHECK((c), "Value of: " #c "\n  Actual: false\nExpected: true", SE_GTEST_FATAL)
#define ASSERT_FALSE(c) SE_GTEST_CHECK(!(c), "Value of: " #c "\n  Actual: true\nExpected: false", SE_GTEST_FATAL)
#define EXPECT_TRUE(c) SE_GTEST_CHECK((c), "Value of: " #c "\n  Actual: false\nExpected: true", SE_GTEST_NONFATAL)
#define EXPECT_FALSE(c) SE_GTEST_CHECK(!(c), "Value of: " #c "\n  Actual: true\nExpected: false", SE_GTEST_NONFATAL)
#define ASSERT_EQ(a, b) SE_GTEST_CHECK(((a) == (b)), "Expected equality of these values:\n  " #a "\n  " #b, SE_GTEST_FATAL)
#define EXPECT_EQ(a, b) SE_GTEST_CHECK(((a) == (b)), "Expected equality of these values:\n  " #a "\n  " #b, SE_GTEST_NONFATAL)
#define ASSERT_NE(a, b) SE_GTEST_CHECK(((a) != (b)), "Expected: (" #a ") != (" #b ")", SE_GTEST_FATAL)
#define EXPECT_NE(a, b) SE_GTEST_CHECK(((a) != (b)), "Expected: (" #a ") != (" #b ")", SE_GTEST_NONFATAL)
#define ASSERT_LT(a, b) SE_GTEST_CHECK(((a) < (b)), "Expected: (" #a ") < (" #b ")", SE_GTEST_FATAL)
#define EXPECT_LT(a, b) SE_GTEST_CHECK(((a) < (b)), "Expected: (" #a ") < (" #b ")", SE_GTEST_NONFATAL)
#define ASSERT_GT(a, b) SE_GTEST_CHECK(((a) > (b)), "Expected: (" #a ") > (" #b ")", SE_GTEST_FATAL)
#define ASSERT_LE(a, b) SE_GTEST_CHECK(((a) <= (b)), "Expected: (" #a ") <= (" #b ")", SE_GTEST_FATAL)
#define ASSERT_GE(a, b) SE_GTEST_CHECK(((a) >= (b)), "Expected: (" #a ") >= (" #b ")", SE_GTEST_FATAL)
#define ASSERT_FLOAT_EQ(a, b) SE_GTEST_CHECK((std::fabs((double)(a) - (double)(b)) <= 4 * 1.1920929e-7 * std::fmax(std::fabs((double)(a)), std::fabs((double)(b)))), "Expected float equality of " #a " and " #b, SE_GTEST_FATAL)
#define EXPECT_FLOAT_EQ(a, b) SE_GTEST_CHECK((std::fabs((double)(a) - (double)(b)) <= 4 * 1.1920929e-7 * std::fmax(std::fabs((double)(a)), std::fabs((double)(b)))), "Expected float equality of " #a " and " #b, SE_GTEST_NONFATAL)
#define ASSERT_NEAR(a, b, tol) SE_GTEST_CHECK((std::fabs((double)(a) - (double)(b)) <= (tol)), "Expected |" #a " - " #b "| <= " #tol, SE_GTEST_FATAL)
