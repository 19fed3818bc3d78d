// What the reference's CMake would generate from include/common/config.h.in with its
// default options (reference CMakeLists.txt:7-12): IS_RSA_SEMIPRIME=ON, IS_DISTRIBUTED=ON,
// IS_SQUARES_CONGRUENCE_CHECK=OFF, USE_GMP=OFF, USE_BOOST=ON, BIG_INT_BITS=64.
#pragma once
#define IS_RSA_SEMIPRIME 1
#define IS_DISTRIBUTED 1
#define USE_BOOST 1
#define BIG_INT_BITS 64
#include <thread>
