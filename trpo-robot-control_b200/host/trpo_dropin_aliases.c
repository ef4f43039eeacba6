/*
 * trpo_dropin_aliases.c -- the reference's own FPGA symbol names (TRPO.h:98,101,113) bound to the GPU implementation, so
 * Test_FVP_FPGA / Test_CG_FPGA / Test_TRPO_Lightweight_FPGA (TRPOCpuCode.c:189,273,457) link against
 * libtrpo_b200_dropin.so unchanged.
 */
#include "../../include/trpo_b200.h"

double FVP_FPGA(TRPOparam param, double *Result, double *Input) { return FVP_GPU(param, Result, Input); }

double CG_FPGA(TRPOparam param, double *Result, double *b, size_t MaxIter, double ResidualTh, size_t NumThreads) {
    return CG_GPU(param, Result, b, MaxIter, ResidualTh, NumThreads);
}

double TRPO_Lightweight_FPGA(TRPOparam param, const int NumIter, const size_t NumThreads) {
    return TRPO_Lightweight_GPU(param, NumIter, NumThreads);
}
