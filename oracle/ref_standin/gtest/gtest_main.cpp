// main() for the reference's unit tests built on the stand-in gtest (the role of libgtest_main).  Test infrastructure only.
#include "gtest.h"
int main(int argc, char** argv) { ::testing::InitGoogleTest(&argc, argv); return RUN_ALL_TESTS(); }
