// fvp_fused.cu -- placeholder until the fused DMMA kernel lands: nothing is eligible, the GEMM-chain path runs.
#include "trpo_internal.cuh"
bool fused_eligible(const NetDesc &) { return false; }
int fused_partial_rows() { return 0; }
int fused_fvp_accumulate(const NetDesc &, const double *, const double *, const double *, const double *, size_t,
                         double *, double *, const int *, cudaStream_t, long long *) { return 1; }
