/*
 * trpo_gpu_host.c -- C host side of the drop-in entry points (include/trpo_b200.h).
 *
 * Mirrors what the reference's FPGA host files do around the device (TRPO_FVP_FPGA.c:13-425, TRPO_CG_FPGA.c:13-638):
 * read the model and rollout text files named in TRPOparam, stage them on the device once, run, de-stage the P-length
 * result, release the device. The numerical work happens in the CUDA kernels behind the trpo_ctx_* shim; there is no
 * CPU fallback: without a GPU every entry point prints [ERROR] and returns -1.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "../../include/trpo_b200.h"

static double now_s(void) {
    struct timeval tv; gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

typedef struct {
    size_t N, O, A, P;
    double *theta, *Mean, *Std, *Observ, *Action, *Advantage;
    int observ_pinned;        /* Observ is page-locked: trpo_ctx_set_batch streams it in chunks under the first FVP */
} HostBatch;

static void host_batch_free(HostBatch *hb) {
    free(hb->theta); free(hb->Mean); free(hb->Std); free(hb->Action); free(hb->Advantage);
    if (hb->observ_pinned) trpo_host_free_pinned(hb->Observ); else free(hb->Observ);
    memset(hb, 0, sizeof(*hb));
}

/* Model file: P numbers (TRPO_FVP.c:670-699). Data file: N rows of Mean[A] Std[A] Observ[O] Action[A] Advantage (:731-762). */
static int host_batch_load(const TRPOparam *param, HostBatch *hb) {
    memset(hb, 0, sizeof(*hb));
    hb->N = param->NumSamples;
    hb->O = param->LayerSize[0];
    hb->A = param->LayerSize[param->NumLayers - 1];
    hb->P = trpo_num_params(param->LayerSize, param->NumLayers);
    FILE *mf = fopen(param->ModelFile, "r");
    if (mf == NULL) {
        fprintf(stderr, "[ERROR] Cannot open Model File [%s]. \n", param->ModelFile);
        return -1;
    }
    hb->theta = (double *)calloc(hb->P, sizeof(double));
    for (size_t i = 0; i < hb->P; ++i)
        if (fscanf(mf, "%lf", &hb->theta[i]) != 1) break;
    fclose(mf);
    FILE *df = fopen(param->DataFile, "r");
    if (df == NULL) {
        fprintf(stderr, "[ERROR] Cannot open Data File [%s]. \n", param->DataFile);
        host_batch_free(hb);
        return -1;
    }
    const size_t N = hb->N, O = hb->O, A = hb->A;
    hb->Mean = (double *)calloc(N * A, sizeof(double));
    hb->Std = (double *)calloc(A, sizeof(double));
    hb->Observ = trpo_host_alloc_pinned(N * O);          /* NULL without a GPU: the calls fail later with the proper message */
    hb->observ_pinned = hb->Observ != NULL;
    if (hb->Observ == NULL) hb->Observ = (double *)calloc(N * O, sizeof(double));
    else memset(hb->Observ, 0, N * O * sizeof(double));
    hb->Action = (double *)calloc(N * A, sizeof(double));
    hb->Advantage = (double *)calloc(N, sizeof(double));
    trpo_batch_file_header bh;
    if (trpo_batch_file_probe(param->DataFile, &bh) == 0) {
        /* binary batch file (trpo_batch_file.c): same content, no decimal parsing */
        fclose(df);
        if (bh.ObservSpaceDim != O || bh.ActionSpaceDim != A || !(bh.flags & 1u) ||
            trpo_batch_file_read(param->DataFile, N, hb->Observ, hb->Std, hb->Mean, hb->Action, hb->Advantage)) {
            fprintf(stderr, "[ERROR] Data File [%s] does not match the network or holds fewer than %zu samples. \n", param->DataFile, N);
            host_batch_free(hb);
            return -1;
        }
        return 0;
    }
    int ok = 1;
    for (size_t n = 0; n < N && ok; ++n) {
        for (size_t j = 0; j < A; ++j) ok &= fscanf(df, "%lf", &hb->Mean[n * A + j]) == 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(df, "%lf", &hb->Std[j]) == 1;      /* last row wins */
        for (size_t j = 0; j < O; ++j) ok &= fscanf(df, "%lf", &hb->Observ[n * O + j]) == 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(df, "%lf", &hb->Action[n * A + j]) == 1;
        ok &= fscanf(df, "%lf", &hb->Advantage[n]) == 1;
    }
    fclose(df);
    if (!ok) {
        /* the reference ignores short files and computes on zeros; a GPU batch with missing rows is an error here */
        fprintf(stderr, "[ERROR] Data File [%s] holds fewer than %zu samples. \n", param->DataFile, N);
        host_batch_free(hb);
        return -1;
    }
    return 0;
}

static trpo_ctx *open_device(const TRPOparam *param, const HostBatch *hb) {
    trpo_ctx *ctx = trpo_ctx_create(param->LayerSize, param->AcFunc, param->NumLayers, -1, TRPO_PRECISION_FP64);
    if (ctx == NULL) {
        fprintf(stderr, "[ERROR] Cannot open the GPU: %s\n", trpo_last_error());
        return NULL;
    }
    if (trpo_ctx_set_model(ctx, hb->theta) ||
        trpo_ctx_set_batch(ctx, hb->N, hb->Observ, hb->Std, hb->Mean, hb->Action, hb->Advantage)) {
        fprintf(stderr, "[ERROR] Staging the batch on the GPU failed: %s\n", trpo_last_error());
        trpo_ctx_destroy(ctx);
        return NULL;
    }
    return ctx;
}

double FVP_GPU(TRPOparam param, double *Result, double *Input) {
    HostBatch hb;
    if (host_batch_load(&param, &hb)) return -1;
    trpo_ctx *ctx = open_device(&param, &hb);
    if (ctx == NULL) { host_batch_free(&hb); return -1; }
    const double t0 = now_s();
    const int rc = trpo_ctx_fvp(ctx, Input, Result, param.CG_Damping);
    const double t1 = now_s();
    if (rc) fprintf(stderr, "[ERROR] Fisher Vector Product Calculation Failed: %s\n", trpo_last_error());
    trpo_ctx_destroy(ctx);
    host_batch_free(&hb);
    return rc ? -1 : t1 - t0;
}

static void print_cg_trace(const trpo_ctx *ctx, const trpo_info *info) {
    /* same line as TRPO_CG.c:56, one per executed iteration plus the terminating one */
    const size_t n = (size_t)info->cg_iters + 1;
    double *rd = (double *)calloc(2 * n, sizeof(double));
    if (rd == NULL) return;
    const int got = trpo_ctx_get_cg_trace(ctx, rd, rd + n, n);
    for (int i = 0; i < got; ++i)
        printf("CG Iter[%d] Residual Norm=%.12e, Soln Norm=%.12e\n", i, rd[i], rd[n + i]);
    free(rd);
}

double CG_GPU(TRPOparam param, double *Result, double *b, size_t MaxIter, double ResidualTh, size_t NumThreads) {
    (void)NumThreads;
    HostBatch hb;
    if (host_batch_load(&param, &hb)) return -1;
    trpo_ctx *ctx = open_device(&param, &hb);
    if (ctx == NULL) { host_batch_free(&hb); return -1; }
    const double t0 = now_s();
    const int rc = trpo_ctx_cg(ctx, b, Result, MaxIter, ResidualTh, param.CG_Damping);
    const double t1 = now_s();
    if (rc) {
        fprintf(stderr, "[ERROR] Fisher Vector Product Calculation Failed: %s\n", trpo_last_error());
    } else {
        trpo_info info;
        trpo_ctx_get_info(ctx, &info);
        print_cg_trace(ctx, &info);
    }
    trpo_ctx_destroy(ctx);
    host_batch_free(&hb);
    return rc ? -1 : t1 - t0;
}

double TRPO_Update_GPU(TRPOparam param, double *Result, size_t NumThreads) {
    (void)NumThreads;
    HostBatch hb;
    if (host_batch_load(&param, &hb)) return -1;
    trpo_ctx *ctx = open_device(&param, &hb);
    if (ctx == NULL) { host_batch_free(&hb); return -1; }
    const double t0 = now_s();
    const int rc = trpo_ctx_update(ctx, Result, param.CG_Damping);
    const double t1 = now_s();
    if (rc) {
        fprintf(stderr, "[ERROR] TRPO Update Failed: %s\n", trpo_last_error());
    } else {
        /* the reference's log lines (TRPO_Update.c:410,819,832,890,998) */
        trpo_info info;
        trpo_ctx_get_info(ctx, &info);
        print_cg_trace(ctx, &info);
        printf("shs: %.14f\n", info.shs);
        printf("lagrange multiplier: %.14f, gnorm: %.14f\n", info.lm, info.gnorm);
        printf("fval before %.14e\n", info.fval);
        for (int i = 0; i < info.ls_steps; ++i)
            printf("a/e/r %.14f / %.14f / %.14f\n", info.ls_actual[i], info.ls_expected[i], info.ls_ratio[i]);
    }
    trpo_ctx_destroy(ctx);
    host_batch_free(&hb);
    return rc ? -1 : t1 - t0;
}
