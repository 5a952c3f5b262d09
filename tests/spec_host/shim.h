// shim.h -- TEST INFRASTRUCTURE ONLY: lets g++ compile the text that symbanafis_b200/csrc/spec.cpp generates for NVRTC, so that
// the generator's output can be EXECUTED on the CPU by tests/test_specialize_cpu.py (no GPU in the build container).
// The product never includes this file and has no CPU path; the shim maps the CUDA vocabulary the generated kernel and
// the device math headers use onto the host: explicit-rounding intrinsics -> IEEE operations (compile with
// -ffp-contract=off; fma() is the correctly rounded libm one), bit casts -> memcpy, __constant__ -> const,
// threadIdx / blockIdx / gridDim -> thread-local variables the driver loop in tests/spec_host/driver.cpp sets.
#pragma once
#define SAF_HOST_SHIM 1
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __constant__ const
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)

struct SafDim3 { unsigned x = 0, y = 0, z = 0; };
extern thread_local SafDim3 threadIdx, blockIdx, gridDim, blockDim;
inline void __syncthreads() {} // threads of a block run one after the other here; the kernels share nothing

inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
inline int __double2loint(double d) { return (int)(__double_as_longlong(d) & 0xffffffffLL); }
inline double __hiloint2double(int hi, int lo) {
    return __longlong_as_double((long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo));
}
using std::max;
using std::min;
// the CUDA math library's global-namespace functions the device headers call (glibc's: <= 1 ulp)
using std::isnan; using std::isinf; using std::signbit;
inline void sincos(double x, double *s, double *c) { *s = std::sin(x); *c = std::cos(x); }
