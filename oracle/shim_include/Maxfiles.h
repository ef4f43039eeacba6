/* Stand-in for the MaxCompiler-generated Maxfiles.h, which the reference's test harness includes (TRPOCpuCode.c:11)
 * but does not need: Test_FVP_FPGA / Test_CG_FPGA only call FVP_FPGA / CG_FPGA (TRPO.h:98,101). Empty on purpose --
 * it lets the checker build compile TRPOCpuCode.c UNMODIFIED against libtrpo_b200_dropin.so. Test infrastructure only. */
