// hd.h -- host/device portability helpers for the core (math-only) headers.
//
// The *_core.h headers hold the arithmetic of the hot path as plain inline functions so that the same source is
// (a) inlined into the sm_100a kernels (kernels.cu) and (b) compiled by g++ into tests/hostcheck, a test-only
// harness that replays the per-column math on the CPU against the oracle.  The product library never executes
// them on the host: every entry point of libecrad_b200.so launches CUDA kernels or fails.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace ecb {

HD double dmin(double a, double b) { return a < b ? a : b; }
HD double dmax(double a, double b) { return a > b ? a : b; }
HD int imin(int a, int b) { return a < b ? a : b; }
HD int imax(int a, int b) { return a > b ? a : b; }

// Round-to-nearest operations that the compiler may NOT contract into FMAs.  Used where a floating-point result
// feeds a discrete decision that must agree bit for bit with the reference (McICA cloud generator: 30-bit random
// numbers are compared with products of cloud fractions, radiation_cloud_generator.F90:337,349).
#if defined(__CUDA_ARCH__)
HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
HD double sub_rn(double a, double b) { return __dadd_rn(a, -b); }
HD double div_rn(double a, double b) { return __ddiv_rn(a, b); }
#else
// host build of the harness uses -ffp-contract=off
HD double mul_rn(double a, double b) { return a * b; }
HD double add_rn(double a, double b) { return a + b; }
HD double sub_rn(double a, double b) { return a - b; }
HD double div_rn(double a, double b) { return a / b; }
#endif

}  // namespace ecb
