/*
 * trpo_lightweight_host.c -- TRPO_Lightweight_GPU: the reference's all-in-one training loop on the lightweight arm
 * simulator (TRPO_Lightweight.c:12-1533; TRPO_Lightweight_FPGA.c is the same loop with the simulator and the CG on the
 * FPGA) with every compute step on the GPU:
 *
 *   rollouts            trpo_ctx_rollout_arm   (fed with the rand() stream the reference's loop consumes, srand(0))
 *   return / GAE        trpo_vf_advantage
 *   baseline fit        the CALLER's libLBFGS driving trpo_vf_evaluate (the reference vendors it as src/lbfgs.c; this
 *                       library does not contain an L-BFGS): `lbfgs` and `lbfgs_parameter_init` are looked up among the
 *                       symbols already loaded in the process, or in the shared object named by $TRPO_LBFGS_LIB
 *   TRPO update         trpo_ctx_update
 *
 * Same files in, same log lines and result files out (`<ResultFile>%03d.txt`, one number per line, %.14f).
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "../../include/trpo_b200.h"

/* lbfgs_parameter_t of libLBFGS 1.10 with LBFGS_FLOAT = 64 (the public interface in lbfgs.h, restated so that this file
 * builds without the header) */
typedef struct {
    int m; double epsilon; int past; double delta; int max_iterations; int linesearch; int max_linesearch;
    double min_step, max_step, ftol, wolfe, gtol, xtol, orthantwise_c;
    int orthantwise_start, orthantwise_end;
} lw_lbfgs_parameter_t;
typedef double (*lw_evaluate_t)(void *, const double *, double *, const int, const double);
typedef int (*lw_lbfgs_t)(int, double *, double *, lw_evaluate_t, void *, void *, lw_lbfgs_parameter_t *);
typedef void (*lw_lbfgs_init_t)(lw_lbfgs_parameter_t *);

static int resolve_lbfgs(lw_lbfgs_t *fn, lw_lbfgs_init_t *init) {
    *fn = (lw_lbfgs_t)dlsym(RTLD_DEFAULT, "lbfgs");
    *init = (lw_lbfgs_init_t)dlsym(RTLD_DEFAULT, "lbfgs_parameter_init");
    if (*fn && *init) return 0;
    const char *path = getenv("TRPO_LBFGS_LIB");
    if (path) {
        void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            *fn = (lw_lbfgs_t)dlsym(h, "lbfgs");
            *init = (lw_lbfgs_init_t)dlsym(h, "lbfgs_parameter_init");
        }
    }
    return (*fn && *init) ? 0 : -1;
}

static int read_numbers(const char *path, double *dst, size_t n) {
    FILE *f = fopen(path, "r");
    if (f == NULL) return -1;
    for (size_t i = 0; i < n; ++i)
        if (fscanf(f, "%lf", &dst[i]) != 1) break;
    fclose(f);
    return 0;
}

double TRPO_Lightweight_GPU_ex(TRPOparam param, const int NumIter, size_t NumEpBatch, size_t EpLen, double gamma, double lam) {
    const size_t NumLayers = param.NumLayers;
    if (NumLayers < 2 || NumLayers > 16 || NumEpBatch == 0 || EpLen == 0) return -1;
    const size_t O = param.LayerSize[0], A = param.LayerSize[NumLayers - 1];
    const size_t NumSamples = NumEpBatch * EpLen;
    const size_t NumParams = trpo_num_params(param.LayerSize, NumLayers);
    /* baseline network: the policy's hidden layers on [observation, step/EpLen] with one output (TRPO_Lightweight.c:36) */
    size_t LayerSizeBase[16];
    for (size_t i = 0; i < NumLayers; ++i) LayerSizeBase[i] = param.LayerSize[i];
    LayerSizeBase[0] = O + 1;
    LayerSizeBase[NumLayers - 1] = 1;
    const size_t NumParamsBase = trpo_num_params(LayerSizeBase, NumLayers) - 1;
    const int PaddedParamsBase = (int)ceil((double)NumParamsBase / 16.0) * 16;

    lw_lbfgs_t lbfgs_fn; lw_lbfgs_init_t lbfgs_init;
    if (resolve_lbfgs(&lbfgs_fn, &lbfgs_init)) {
        fprintf(stderr, "[ERROR] libLBFGS not found: link the application with lbfgs.c or set TRPO_LBFGS_LIB. \n");
        return -1;
    }
    double *theta = (double *)calloc(NumParams, sizeof(double));
    double *theta_new = (double *)calloc(NumParams, sizeof(double));
    double *Reward = (double *)calloc(NumSamples, sizeof(double));
    const size_t ndraws = NumEpBatch * (3 + 2 * A * EpLen);
    int *draws = (int *)calloc(ndraws, sizeof(int));
    double *x = NULL;
    if (posix_memalign((void **)&x, 64, (size_t)PaddedParamsBase * sizeof(double))) x = NULL;
    trpo_ctx *ctx = NULL;
    trpo_vf *vf = NULL;
    double runtime = -1;
    if (!theta || !theta_new || !Reward || !draws || !x) goto done;
    memset(x, 0, (size_t)PaddedParamsBase * sizeof(double));
    if (read_numbers(param.ModelFile, theta, NumParams)) {
        fprintf(stderr, "[ERROR] Cannot open Model File [%s]. \n", param.ModelFile);
        goto done;
    }
    if (read_numbers(param.BaselineFile, x, NumParamsBase)) {
        fprintf(stderr, "[ERROR] Cannot open BaselineFile [%s]. \n", param.BaselineFile);
        goto done;
    }
    ctx = trpo_ctx_create(param.LayerSize, param.AcFunc, NumLayers, -1, TRPO_PRECISION_FP64);
    if (ctx) vf = trpo_vf_create(ctx, LayerSizeBase, param.AcFunc, NumLayers);
    if (!ctx || !vf) {
        fprintf(stderr, "[ERROR] Cannot open the GPU: %s\n", trpo_last_error());
        goto done;
    }
    lw_lbfgs_parameter_t prm;
    lbfgs_init(&prm);
    prm.max_iterations = 25;                                  /* TRPO_Lightweight.c:327 */
    srand(0);                                                 /* TRPO_Lightweight.c:66 */

    struct timeval tv1, tv2;
    gettimeofday(&tv1, NULL);
    int failed = 0;
    for (int iter = 0; iter < NumIter && !failed; ++iter) {
        for (size_t i = 0; i < ndraws; ++i) draws[i] = rand();
        failed = trpo_ctx_set_model(ctx, theta) || trpo_ctx_rollout_arm(ctx, NumEpBatch, EpLen, draws, 0) ||
                 trpo_ctx_get_rollout(ctx, NULL, NULL, NULL, Reward);
        if (failed) break;
        /* reward statistics (TRPO_Lightweight.c:545-560) */
        double EpRewMean = 0, EpRewStd = 0;
        for (size_t i = 0; i < NumSamples; ++i) EpRewMean += Reward[i];
        EpRewMean = EpRewMean / (double)NumEpBatch;
        for (size_t ep = 0; ep < NumEpBatch; ++ep) {
            double r = 0;
            for (size_t j = 0; j < EpLen; ++j) r += Reward[ep * EpLen + j];
            EpRewStd += (r - EpRewMean) * (r - EpRewMean);
        }
        EpRewStd = sqrt(EpRewStd / (double)NumEpBatch);
        printf("[INFO] Iteration %d, Episode Rewards Mean = %f, Std = %f\n", iter, EpRewMean, EpRewStd);

        failed = trpo_vf_advantage(vf, x, gamma, lam, NULL, NULL);
        if (failed) break;
        double fx = 0;
        const int lb = lbfgs_fn(PaddedParamsBase, x, &fx, trpo_vf_evaluate, NULL, vf, &prm);   /* TRPO_Lightweight.c:675 */
        /* the reference ignores lbfgs()'s status (a line-search or iteration-limit stop is normal); a failed objective
         * evaluation on the GPU is not: the fit would continue from a corrupted x */
        (void)lb;
        if (trpo_vf_failed(vf)) { failed = 1; break; }
        failed = trpo_ctx_update(ctx, theta_new, param.CG_Damping);
        if (failed) break;
        trpo_info info;
        trpo_ctx_get_info(ctx, &info);
        for (int i = 0; i <= info.cg_iters; ++i)
            printf("CG Iter[%d] Residual Norm=%.12e, Soln Norm=%.12e\n", i, info.cg_rdotr[i], info.cg_xnorm[i]);
        printf("shs: %.14f\n", info.shs);
        printf("lagrange multiplier: %.14f, gnorm: %.14f\n", info.lm, info.gnorm);
        printf("fval before %.14e\n", info.fval);
        for (int i = 0; i < info.ls_steps; ++i)
            printf("a/e/r %.14f / %.14f / %.14f\n", info.ls_actual[i], info.ls_expected[i], info.ls_ratio[i]);
        memcpy(theta, theta_new, NumParams * sizeof(double));

        if (iter % 100 == 0 || iter == NumIter - 1) {         /* TRPO_Lightweight.c:1466-1505 */
            char name[4096];
            snprintf(name, sizeof(name), "%s%03d.txt", param.ResultFile, iter);
            FILE *rf = fopen(name, "w");
            if (rf == NULL) {
                fprintf(stderr, "[ERROR] Cannot open Result File [%s]. \n", name);
                failed = 1;
                break;
            }
            for (size_t i = 0; i < NumParams; ++i) fprintf(rf, "%.14f\n", theta[i]);
            fclose(rf);
        }
    }
    gettimeofday(&tv2, NULL);
    if (failed) {
        if (trpo_last_error()[0]) fprintf(stderr, "[ERROR] TRPO on the GPU failed: %s\n", trpo_last_error());
    } else {
        runtime = ((tv2.tv_sec - tv1.tv_sec) * (double)1E6 + (tv2.tv_usec - tv1.tv_usec)) / (double)1E6;
    }
done:
    if (vf) trpo_vf_destroy(vf);
    if (ctx) trpo_ctx_destroy(ctx);
    free(theta); free(theta_new); free(Reward); free(draws); free(x);
    return runtime;
}

double TRPO_Lightweight_GPU(TRPOparam param, const int NumIter, const size_t NumThreads) {
    (void)NumThreads;
    /* 20 episodes of 150 steps, gamma 0.995, lambda 0.98: the constants of TRPO_Lightweight.c:32-33,52-55 */
    return TRPO_Lightweight_GPU_ex(param, NumIter, 20, 150, 0.995, 0.98);
}
