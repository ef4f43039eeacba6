/* ref_harness_main.c -- entry point for the reference's OWN test harness linked against the drop-in library.
 *
 * Test infrastructure only (never linked by the product). The checker build (oracle/Makefile, target `ref`) compiles
 * /root/reference/src/TRPOCpuCode.c unmodified (its main() renamed on the command line, its two MaxCompiler headers
 * replaced by the empty stand-ins in oracle/shim_include/) and links it with
 *   - the reference's CPU FVP() / CG() (oracle/_ref/libtrpo_ref.so: the unmodified TRPO_FVP.c / TRPO_CG.c / TRPO_Util.c), and
 *   - FVP_FPGA / CG_FPGA from trpo-robot-control_b200/libtrpo_b200_dropin.so,
 * so Test_FVP_FPGA (TRPOCpuCode.c:138-223) and Test_CG_FPGA (:225-311) run exactly as shipped, with the GPU in the
 * FPGA's place. They read ArmTestModel.txt / ArmTestData.txt / ArmTestFVP.txt / ArmTestCG.txt from the working directory
 * and print "[INFO] Mean Absolute Percentage Error = ...%". */
#include <stddef.h>
#include <string.h>

void Test_FVP_FPGA(void);
void Test_CG_FPGA(size_t NumThreads);

int main(int argc, char **argv) {
    const char *what = argc > 1 ? argv[1] : "all";
    if (strcmp(what, "fvp") == 0 || strcmp(what, "all") == 0) Test_FVP_FPGA();
    if (strcmp(what, "cg") == 0 || strcmp(what, "all") == 0) Test_CG_FPGA(1);
    return 0;
}
