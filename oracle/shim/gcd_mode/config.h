// What the reference's CMake would generate from include/common/config.h.in with -DIS_RSA_SEMIPRIME=OFF
// (reference CMakeLists.txt:7): the "has common factor" build, getSmoothNumbersIteration takes the gcd branch
// (include/qimcifa.hpp:373-379).  Everything else as in ../config.h.
#pragma once
/* #undef IS_RSA_SEMIPRIME */
#define IS_DISTRIBUTED 1
#define USE_BOOST 1
#define BIG_INT_BITS 64
#include <thread>
