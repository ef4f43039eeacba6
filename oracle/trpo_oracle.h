/*
 * oracle/trpo_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, in-memory) of the reference's natural-gradient solve:
 *   FVPFast      /root/reference/src/TRPO_FVP.c:548-949
 *   FVP (4-pass) /root/reference/src/TRPO_FVP.c:11-545
 *   CG           /root/reference/src/TRPO_CG.c:11-113
 *   TRPO_Update  /root/reference/src/TRPO_Update.c:10-1056
 *   NumParamsCalc/root/reference/src/TRPO_Util.c:7-17
 *   evaluate (baseline objective)  /root/reference/src/TRPO_Baseline.c:29-237
 *   rollout / return / GAE blocks  /root/reference/src/TRPO_Lightweight.c:349-653
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker. The product (libtrpo_b200.so) never links or calls it.
 *
 * Parity pinning: bit-exact against the compiled, unmodified reference (oracle/_ref/, built by
 * oracle/Makefile with -ffp-contract=off) on the ArmTest vectors and on synthetic cases; the
 * %.17g outputs of that reference run are committed under tests/golden/ (see tests/golden/make_golden.py).
 * The rollout / GAE / baseline-objective restatements are pinned against the reference's exported `evaluate` callback
 * (bit-exact) and against the result files the unmodified TRPO_Lightweight writes after 1-3 iterations
 * (tests/golden/make_golden_lightweight.py, tests/test_oracle_rollout.py).
 */
#ifndef TRPO_ORACLE_H
#define TRPO_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Network description shared by all oracle calls (mirrors the TRPOparam fields the path uses,
 * /root/reference/src/include/TRPO.h:21-37). AcFunc has NumLayers chars, AcFunc[0] unused. */
typedef struct {
    size_t        NumLayers;
    const char   *AcFunc;
    const size_t *LayerSize;
} OracleNet;

/* Diagnostics of one TRPO_Update (the values the reference printf's, TRPO_Update.c:819,832,890,998). */
typedef struct {
    int    cg_iters;           /* number of FVPs executed inside CG */
    double cg_rdotr[16];       /* "Residual Norm" printed at each CG iteration */
    double cg_xnorm[16];       /* "Soln Norm" printed at each CG iteration */
    double shs, lm, gnorm, fval;
    int    ls_steps;           /* number of line-search trials evaluated */
    int    ls_accepted;        /* 1 if a step was accepted */
    double ls_actual[16], ls_expected[16], ls_ratio[16];
} OracleUpdateInfo;

size_t oracle_num_params(const size_t *LayerSize, size_t NumLayers);

/* Text loaders (formats: TRPO_FVP.c:670-699 model, :731-762 data). Return 0 on success, -1 on open failure. */
int oracle_load_model(const char *path, const OracleNet *net, double *theta);
int oracle_load_data(const char *path, const OracleNet *net, size_t NumSamples,
                     double *Mean, double *Std, double *Observ, double *Action, double *Advantage);

/* Policy mean for every sample (ordinary forward pass, TRPO_Update.c:259-291). Mean is N x A. */
int oracle_forward(const OracleNet *net, const double *theta, const double *Observ, size_t NumSamples, double *Mean);

/* FVPFast restated: Result = (1/N) sum_n FVP_n(Input) + damping*Input.  Std has A entries (TRPO_FVP.c:746-748,853). */
int oracle_fvp_fast(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
                    size_t NumSamples, double CG_Damping, const double *Input, double *Result);

/* FVP (4-pass) restated (TRPO_FVP.c:266-521). */
int oracle_fvp_4pass(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
                     size_t NumSamples, double CG_Damping, const double *Input, double *Result);

/* CG restated (TRPO_CG.c:25-107), serial dot products (== reference at NumThreads=1).
 * rdotr_trace/xnorm_trace (may be NULL) receive the values printed per iteration (MaxIter+1 slots max).
 * Returns the number of FVPs executed, or -1 on error. */
int oracle_cg(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
              size_t NumSamples, double CG_Damping, const double *b, size_t MaxIter, double ResidualTh,
              double *Result, double *rdotr_trace, double *xnorm_trace);

/* Policy gradient b = (1/N) sum_n grad(ratio*A) (TRPO_Update.c:254-378). LogStd is theta's tail. */
int oracle_policy_gradient(const OracleNet *net, const double *theta, const double *Observ, const double *Mean,
                           const double *Action, const double *Advantage, size_t NumSamples, double *b);

/* Whole TRPO_Update restated (TRPO_Update.c:254-1011). Std = data-file Std. */
int oracle_update(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
                  const double *Mean, const double *Action, const double *Advantage, size_t NumSamples,
                  double CG_Damping, double *Result, OracleUpdateInfo *info);

/* ---- rows f-3 / f-4: the steps either side of the update in the training loop -------------------------------------
 * vfnet describes the value-function ("baseline") network: LayerSize[0] = ObservSpaceDim + 1 (the observation followed
 * by step/EpLen), last layer 1; x = [W0,B0,...] without a LogStd tail (TRPO_Lightweight.c:655-672). */

/* Baseline prediction of every step (TRPO_Lightweight.c:582-625). */
int oracle_vf_predict(const OracleNet *vfnet, size_t NumEpBatch, size_t EpLen, const double *Observ, const double *x,
                      double *Baseline);

/* libLBFGS objective callback restated (TRPO_Baseline.c:29-237): returns 0.01*MSE + 0.001*|x|^2, writes the gradient
 * (n >= number of parameters; the padding is zeroed) and the predictions. */
double oracle_vf_evaluate(const OracleNet *vfnet, size_t NumEpBatch, size_t EpLen, const double *Observ,
                          const double *Target, const double *x, double *g, int n, double *Predict);

/* Episode reward statistics printed per iteration (TRPO_Lightweight.c:545-558). */
void oracle_reward_stats(size_t NumEpBatch, size_t EpLen, const double *Reward, double *EpRewMean, double *EpRewStd);

/* Discounted return, GAE(gamma, lam) advantage and its standardisation (TRPO_Lightweight.c:565-653). Reward is
 * overwritten with the TD residuals exactly as the reference does. */
int oracle_gae(size_t NumEpBatch, size_t EpLen, double gamma, double lam, double *Reward, const double *Baseline,
               double *Return, double *Advantage);

/* One batch of rollouts of the lightweight arm simulator (TRPO_Lightweight.c:349-540); consumes rand(). */
int oracle_arm_rollout(const OracleNet *net, const double *theta, size_t NumEpBatch, size_t EpLen,
                       double *Observ, double *Mean, double *Std, double *Action, double *Reward);

#ifdef __cplusplus
}
#endif
#endif
