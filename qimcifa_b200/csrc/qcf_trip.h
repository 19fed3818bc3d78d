// Guesses per trip of the filter kernel, by polynomial degree: ONE definition for the host planner (qcf_api.cu), the kernel
// (qcf_filter.cu), the chain arithmetic (qcf_filter_math.cuh) and the host build of it (tests/device_math_host.cpp).
//   degree 1 and 2: 32;  degree 3: 16.  (64 at degree 1 -- 32 registers of two packed offsets -- was measured on the B200:
//   36.3 against 37.3 ms at 128 bit, 44.6 against 41.6 ms at 112 bit, where the segments are short: not kept.)
#pragma once
#define QCF_FILTER_UNROLL_FOR(D) (((D) >= 3) ? 16 : 32)
