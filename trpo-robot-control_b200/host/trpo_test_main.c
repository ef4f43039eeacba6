/*
 * trpo_test_main.c -- C driver in the shape of the reference's own harness (TRPOCpuCode.c): the commented CPU tests
 * Test_FVP / Test_CG / Test_TRPO_Update (:14-135,:313-373) and the FPGA tests Test_FVP_FPGA / Test_CG_FPGA (:138-311),
 * with the GPU entry points in the device role. Expected values come from a two-column "input expected" text file
 * (ArmTestFVP.txt / ArmTestCG.txt format) or a one-column model file (ArmTestModelUpdated.txt).
 *
 *   trpo_test_gpu fvp    <model> <data> <N> <vectors.txt>
 *   trpo_test_gpu cg     <model> <data> <N> <vectors.txt>
 *   trpo_test_gpu update <model> <data> <N> <expected_model.txt>
 *   trpo_test_gpu lightweight <model> <baseline> <result-prefix> <NumIter>     (Test_TRPO_Lightweight_FPGA, :433-465)
 * Network: ArmDOF_0-v0, 15-16-16-3, {'l','t','t','l'}, CG_Damping 0.1 (TRPOCpuCode.c:142-160) unless
 * TRPO_LAYERS="17,64,64,6" TRPO_ACFUNC="lttl" are set in the environment.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/trpo_b200.h"

static void report(const char *what, const double *got, const double *expect, size_t n) {
    /* reference-style per-element percentage error (TRPOCpuCode.c:58-66) plus norm-relative metrics */
    double mape = 0, maxabs = 0, maxref = 0, num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) {
        const double d = got[i] - expect[i];
        if (expect[i] != 0) {
            const double e = fabs(d / expect[i]) * 100.0;
            mape += e;
            if (e > 1) printf("%s[%zu]=%e, Expect=%e. %.4f%% Difference\n", what, i, got[i], expect[i], e);
        }
        if (fabs(d) > maxabs) maxabs = fabs(d);
        if (fabs(expect[i]) > maxref) maxref = fabs(expect[i]);
        num += d * d; den += expect[i] * expect[i];
    }
    printf("[INFO] %s Mean Absolute Percentage Error = %.12f%%\n", what, mape / (double)n);
    printf("[INFO] %s max|d|/max|ref| = %.3e, rel-L2 = %.3e\n", what, maxabs / maxref, sqrt(num / den));
}

/* Test_TRPO_Lightweight_FPGA (TRPOCpuCode.c:433-465): the whole training loop, NumIter iterations */
static int run_lightweight(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s lightweight <model> <baseline> <result-prefix> <NumIter>   (TRPO_LBFGS_LIB names libLBFGS)\n", argv[0]);
        return 2;
    }
    size_t LayerSize[4] = {15, 16, 16, 3};
    char AcFunc[4] = {'l', 't', 't', 'l'};
    TRPOparam Param;
    memset(&Param, 0, sizeof(Param));
    Param.ModelFile = argv[2];
    Param.BaselineFile = argv[3];
    Param.ResultFile = argv[4];
    Param.NumLayers = 4;
    Param.AcFunc = AcFunc;
    Param.LayerSize = LayerSize;
    Param.CG_Damping = 0.1;
    const int NumIter = atoi(argv[5]);
    const double compTime = TRPO_Lightweight_GPU(Param, NumIter, 1);
    if (compTime < 0) fprintf(stderr, "[ERROR] TRPO Lightweight Failed.\n");
    else printf("[INFO] TRPO Lightweight GPU: %d iterations in %f seconds.\n", NumIter, compTime);
    return compTime < 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "lightweight") == 0) return run_lightweight(argc, argv);
    if (argc < 6) {
        fprintf(stderr, "usage: %s fvp|cg|update <model> <data> <N> <vectors>\n       %s lightweight <model> <baseline> <result-prefix> <NumIter>\n", argv[0], argv[0]);
        return 2;
    }
    size_t LayerSize[16] = {15, 16, 16, 3};
    char AcFunc[17] = {'l', 't', 't', 'l'};
    size_t NumLayers = 4;
    const char *ls = getenv("TRPO_LAYERS"), *af = getenv("TRPO_ACFUNC");
    if (ls && af) {
        NumLayers = 0;
        char *copy = strdup(ls);
        for (char *tok = strtok(copy, ","); tok && NumLayers < 16; tok = strtok(NULL, ",")) LayerSize[NumLayers++] = strtoul(tok, NULL, 10);
        free(copy);
        strncpy(AcFunc, af, 16);
    }
    TRPOparam Param;
    memset(&Param, 0, sizeof(Param));
    Param.ModelFile = argv[2];
    Param.DataFile = argv[3];
    Param.NumLayers = NumLayers;
    Param.AcFunc = AcFunc;
    Param.LayerSize = LayerSize;
    Param.NumSamples = strtoul(argv[4], NULL, 10);
    Param.CG_Damping = 0.1;

    const size_t P = trpo_num_params(LayerSize, NumLayers);
    double *input = (double *)calloc(P, sizeof(double));
    double *result = (double *)calloc(P, sizeof(double));
    double *expect = (double *)calloc(P, sizeof(double));
    FILE *f = fopen(argv[5], "r");
    if (f == NULL) { fprintf(stderr, "[ERROR] Cannot open Data File [%s]. \n", argv[5]); return 1; }
    const int is_update = strcmp(argv[1], "update") == 0;
    for (size_t i = 0; i < P; ++i) {
        if (is_update) { if (fscanf(f, "%lf", &expect[i]) != 1) break; }
        else if (fscanf(f, "%lf %lf", &input[i], &expect[i]) != 2) break;
    }
    fclose(f);

    double t;
    if (strcmp(argv[1], "fvp") == 0) {
        t = FVP_GPU(Param, result, input);
        if (t < 0) { fprintf(stderr, "[ERROR] Fisher Vector Product Calculation Failed.\n"); return 1; }
        report("FVP_GPU", result, expect, P);
    } else if (strcmp(argv[1], "cg") == 0) {
        printf("---------------------- CG Test GPU -----------------------\n");
        t = CG_GPU(Param, result, input, 10, 1e-10, 1);
        if (t < 0) { fprintf(stderr, "[ERROR] GPU-based Conjugate Gradient Calculation Failed.\n"); return 1; }
        report("CG_GPU", result, expect, P);
    } else if (is_update) {
        t = TRPO_Update_GPU(Param, result, 1);
        if (t < 0) { fprintf(stderr, "[ERROR] TRPO Update Failed.\n"); return 1; }
        report("TRPO_Update_GPU", result, expect, P);
    } else {
        fprintf(stderr, "unknown test %s\n", argv[1]);
        return 2;
    }
    printf("[INFO] GPU Computing Time = %f seconds\n", t);
    free(input); free(result); free(expect);
    return 0;
}
