// driver.cpp -- TEST INFRASTRUCTURE ONLY (see shim.h): runs the generated kernel `saf_spec` / `saf_spec_pf` on the CPU,
// block by block and thread by thread, for tests/test_specialize_cpu.py.  Compiled together with the generated text:
//   g++ -O1 -ffp-contract=off -include shim.h -I <csrc> -DSAF_GENERATED='"<file>"' driver.cpp
#include "shim.h"
thread_local SafDim3 threadIdx, blockIdx, gridDim, blockDim;
#include SAF_GENERATED

extern "C" int saf_spec_host_run(int entry, unsigned grid, unsigned long long n_points, unsigned long long iters,
                                 const double *const *cols, const unsigned long long *col_len, int n_cols, double *const *outs,
                                 int n_outs) {
    saf::sp::SpecDesc d;
    std::memset(&d, 0, sizeof d);
    d.n_points = n_points;
    d.iters = iters;
    for (int j = 0; j < n_cols; ++j) { d.cols[j] = cols[j]; d.col_len[j] = col_len[j]; }
    for (int k = 0; k < n_outs; ++k) d.outs[k] = outs[k];
    gridDim.x = grid;
    blockDim.x = SPEC_BLOCK;
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        for (unsigned t = 0; t < SPEC_BLOCK; ++t) {
            threadIdx.x = t;
            if (entry == 0) saf_spec(d);
#ifdef SAF_HAS_PF
            else saf_spec_pf(d);
#else
            else return 1;
#endif
        }
    }
    return 0;
}
extern "C" int saf_spec_host_shape(int *p, int *block) { *p = SPEC_P; *block = SPEC_BLOCK; return 0; }
