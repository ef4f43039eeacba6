/*
 * oracle/trpo_oracle.c -- TEST INFRASTRUCTURE ONLY (see trpo_oracle.h).
 *
 * In-memory CPU restatement of the reference's FVP / CG / TRPO_Update arithmetic. Every loop keeps the
 * reference's operation order so that, compiled with -ffp-contract=off, it reproduces the compiled reference
 * bit for bit (tests/test_oracle.py checks this against oracle/_ref when present and against tests/golden/).
 * It is a restatement, not a copy: one generic "sample pass" routine serves FVPFast, the CG loop and the
 * update, where the reference pastes the loop three times.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "trpo_oracle.h"

#define MAX_LAYERS 16

size_t oracle_num_params(const size_t *LayerSize, size_t NumLayers) {
    /* TRPO_Util.c:7-17 */
    size_t n = 0;
    for (size_t i = 0; i + 1 < NumLayers; ++i) n += LayerSize[i] * LayerSize[i + 1] + LayerSize[i + 1];
    return n + LayerSize[NumLayers - 1];
}

/* Offsets of W[i], B[i] and LogStd inside the flat vector (TRPO_FVP.c:704-725). */
typedef struct {
    size_t K;                     /* number of weight layers */
    size_t L[MAX_LAYERS];
    size_t w[MAX_LAYERS], b[MAX_LAYERS], logstd, P;
} Layout;

static int make_layout(const OracleNet *net, Layout *lo) {
    if (net->NumLayers < 2 || net->NumLayers > MAX_LAYERS) return -1;
    lo->K = net->NumLayers - 1;
    size_t pos = 0;
    for (size_t i = 0; i < net->NumLayers; ++i) lo->L[i] = net->LayerSize[i];
    for (size_t i = 0; i < lo->K; ++i) {
        lo->w[i] = pos; pos += lo->L[i] * lo->L[i + 1];
        lo->b[i] = pos; pos += lo->L[i + 1];
    }
    lo->logstd = pos;
    lo->P = pos + lo->L[lo->K];
    return 0;
}

static int ac_supported(char c) { return c == 'l' || c == 't' || c == 'o' || c == 's'; }

static int check_acfunc(const OracleNet *net) {
    for (size_t i = 1; i < net->NumLayers; ++i)
        if (!ac_supported(net->AcFunc[i])) {
            fprintf(stderr, "[ERROR] AC Function for Layer[%zu] is %c. Unsupported.\n", i, net->AcFunc[i]);
            return -1;
        }
    return 0;
}

int oracle_load_model(const char *path, const OracleNet *net, double *theta) {
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "[ERROR] Cannot open Model File [%s]. \n", path); return -1; }
    size_t P = oracle_num_params(net->LayerSize, net->NumLayers);
    for (size_t i = 0; i < P; ++i)
        if (fscanf(f, "%lf", &theta[i]) != 1) theta[i] = 0;
    fclose(f);
    return 0;
}

int oracle_load_data(const char *path, const OracleNet *net, size_t N,
                     double *Mean, double *Std, double *Observ, double *Action, double *Advantage) {
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "[ERROR] Cannot open Data File [%s]. \n", path); return -1; }
    const size_t O = net->LayerSize[0], A = net->LayerSize[net->NumLayers - 1];
    for (size_t n = 0; n < N; ++n) {
        int ok = 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(f, "%lf", &Mean[n * A + j]) == 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(f, "%lf", &Std[j]) == 1;   /* last row wins, TRPO_FVP.c:746-748 */
        for (size_t j = 0; j < O; ++j) ok &= fscanf(f, "%lf", &Observ[n * O + j]) == 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(f, "%lf", &Action[n * A + j]) == 1;
        ok &= fscanf(f, "%lf", &Advantage[n]) == 1;
        (void)ok;
    }
    fclose(f);
    return 0;
}

/* Scratch for one sample. */
typedef struct {
    double *y[MAX_LAYERS], *rx[MAX_LAYERS], *ry[MAX_LAYERS], *g[MAX_LAYERS], *rg[MAX_LAYERS];
} Scratch;

static void scratch_alloc(Scratch *s, const Layout *lo) {
    for (size_t i = 0; i <= lo->K; ++i) {
        s->y[i]  = (double *)calloc(lo->L[i], sizeof(double));
        s->rx[i] = (double *)calloc(lo->L[i], sizeof(double));
        s->ry[i] = (double *)calloc(lo->L[i], sizeof(double));
        s->g[i]  = (double *)calloc(lo->L[i], sizeof(double));
        s->rg[i] = (double *)calloc(lo->L[i], sizeof(double));
    }
}
static void scratch_free(Scratch *s, const Layout *lo) {
    for (size_t i = 0; i <= lo->K; ++i) { free(s->y[i]); free(s->rx[i]); free(s->ry[i]); free(s->g[i]); free(s->rg[i]); }
}

/* Ordinary forward pass of one sample (TRPO_Update.c:259-291): y[0] must hold the observation. */
static void forward_one(const Layout *lo, const char *ac, const double *theta, Scratch *s) {
    for (size_t i = 0; i < lo->K; ++i) {
        const size_t cur = lo->L[i], nxt = lo->L[i + 1];
        const double *W = theta + lo->w[i], *B = theta + lo->b[i];
        for (size_t j = 0; j < nxt; ++j) {
            double x = B[j];
            for (size_t k = 0; k < cur; ++k) x += s->y[i][k] * W[k * nxt + j];
            switch (ac[i + 1]) {
                case 't': x = tanh(x); break;
                case 'o': x = 0.1 * x; break;
                case 's': x = 1.0 / (1 + exp(-x)); break;
                default: break;
            }
            s->y[i + 1][j] = x;
        }
    }
}

int oracle_forward(const OracleNet *net, const double *theta, const double *Observ, size_t N, double *Mean) {
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    Scratch s; scratch_alloc(&s, &lo);
    const size_t O = lo.L[0], A = lo.L[lo.K];
    for (size_t n = 0; n < N; ++n) {
        memcpy(s.y[0], Observ + n * O, O * sizeof(double));
        forward_one(&lo, net->AcFunc, theta, &s);
        memcpy(Mean + n * A, s.y[lo.K], A * sizeof(double));
    }
    scratch_free(&s, &lo);
    return 0;
}

/* Un-normalised FVPFast sum over all samples: acc[0..P) += per-sample [RGW,RGB,...,2*VLogStd]
 * (TRPO_FVP.c:771-924). acc must be zeroed by the caller. */
static void fvp_fast_accumulate(const Layout *lo, const char *ac, const double *theta, const double *Std,
                                const double *Observ, size_t N, const double *V, double *acc, Scratch *s) {
    const size_t K = lo->K, O = lo->L[0], A = lo->L[K];
    double *rgw = (double *)calloc(lo->P, sizeof(double));   /* per-sample RGW/RGB image, rewritten every sample */
    for (size_t n = 0; n < N; ++n) {
        for (size_t k = 0; k < O; ++k) { s->y[0][k] = Observ[n * O + k]; s->rx[0][k] = 0; s->ry[0][k] = 0; }
        /* combined forward + R-forward (TRPO_FVP.c:783-836) */
        for (size_t i = 0; i < K; ++i) {
            const size_t cur = lo->L[i], nxt = lo->L[i + 1];
            const double *W = theta + lo->w[i], *B = theta + lo->b[i];
            const double *VW = V + lo->w[i], *VB = V + lo->b[i];
            for (size_t j = 0; j < nxt; ++j) {
                double x = B[j], rx = VB[j];
                for (size_t k = 0; k < cur; ++k) {
                    x  += s->y[i][k]  * W[k * nxt + j];
                    rx += s->ry[i][k] * W[k * nxt + j];
                    rx += s->y[i][k]  * VW[k * nxt + j];
                }
                double y = x, ry = rx;
                switch (ac[i + 1]) {
                    case 'l': break;
                    case 't': y = tanh(x); ry = rx * (1 - y * y); break;
                    case 'o': y = 0.1 * x; ry = 0.1 * rx; break;
                    case 's': y = 1.0 / (1 + exp(-x)); ry = rx * y * (1 - y); break;
                    default: break;
                }
                s->y[i + 1][j] = y; s->rx[i + 1][j] = rx; s->ry[i + 1][j] = ry;
            }
        }
        /* R-gradient seed (TRPO_FVP.c:852-854) */
        for (size_t j = 0; j < A; ++j) s->rg[K][j] = s->ry[K][j] / Std[j] / Std[j];
        /* R-backward (TRPO_FVP.c:857-900) */
        for (size_t i = K; i > 0; --i) {
            const size_t cur = lo->L[i], prv = lo->L[i - 1];
            const double *W = theta + lo->w[i - 1];
            for (size_t j = 0; j < cur; ++j) {
                const double y = s->y[i][j];
                switch (ac[i]) {
                    case 't': s->rg[i][j] = (1 - y * y) * s->rg[i][j]; break;
                    case 'o': s->rg[i][j] = 0.1 * s->rg[i][j]; break;
                    case 's': s->rg[i][j] = s->rg[i][j] * y * (1 - y); break;
                    default: break;
                }
                rgw[lo->b[i - 1] + j] = s->rg[i][j];
            }
            for (size_t j = 0; j < prv; ++j) {
                double t = 0;
                for (size_t k = 0; k < cur; ++k) {
                    rgw[lo->w[i - 1] + j * cur + k] = s->y[i - 1][j] * s->rg[i][k];
                    t += W[j * cur + k] * s->rg[i][k];
                }
                s->rg[i - 1][j] = t;
            }
        }
        /* accumulate (TRPO_FVP.c:903-921) */
        for (size_t q = 0; q < lo->logstd; ++q) acc[q] += rgw[q];
        for (size_t j = 0; j < A; ++j) acc[lo->logstd + j] += 2 * V[lo->logstd + j];
    }
    free(rgw);
}

static void fvp_finalise(size_t P, size_t N, double damping, const double *V, double *acc) {
    /* TRPO_FVP.c:928-931 */
    for (size_t q = 0; q < P; ++q) acc[q] = acc[q] / (double)N + damping * V[q];
}

int oracle_fvp_fast(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
                    size_t N, double damping, const double *Input, double *Result) {
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    Scratch s; scratch_alloc(&s, &lo);
    memset(Result, 0, lo.P * sizeof(double));
    fvp_fast_accumulate(&lo, net->AcFunc, theta, Std, Observ, N, Input, Result, &s);
    fvp_finalise(lo.P, N, damping, Input, Result);
    scratch_free(&s, &lo);
    return 0;
}

int oracle_fvp_4pass(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
                     size_t N, double damping, const double *Input, double *Result) {
    /* TRPO_FVP.c:266-527: ordinary fwd, ordinary bwd with an all-zero seed, R-fwd, R-bwd with the G cross terms. */
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    const char *ac = net->AcFunc;
    const size_t K = lo.K, O = lo.L[0], A = lo.L[K];
    Scratch sc; scratch_alloc(&sc, &lo); Scratch *s = &sc;
    double *rgw = (double *)calloc(lo.P, sizeof(double));
    double *RStd = (double *)calloc(A, sizeof(double));
    memset(Result, 0, lo.P * sizeof(double));
    for (size_t n = 0; n < N; ++n) {
        memcpy(s->y[0], Observ + n * O, O * sizeof(double));
        forward_one(&lo, ac, theta, s);
        /* ordinary backward, zero seed (TRPO_FVP.c:325-372) */
        for (size_t j = 0; j < A; ++j) s->g[K][j] = 0;
        for (size_t i = K; i > 0; --i) {
            const size_t cur = lo.L[i], prv = lo.L[i - 1];
            const double *W = theta + lo.w[i - 1];
            for (size_t j = 0; j < cur; ++j) {
                const double y = s->y[i][j];
                switch (ac[i]) {
                    case 't': s->g[i][j] = s->g[i][j] * (1 - y * y); break;
                    case 'o': s->g[i][j] = 0.1 * s->g[i][j]; break;
                    case 's': s->g[i][j] = s->g[i][j] * y * (1 - y); break;
                    default: break;
                }
            }
            for (size_t j = 0; j < prv; ++j) {
                s->g[i - 1][j] = 0;
                for (size_t k = 0; k < cur; ++k) s->g[i - 1][j] += s->g[i][k] * W[j * cur + k];
            }
        }
        /* R-forward (TRPO_FVP.c:376-417) */
        for (size_t k = 0; k < O; ++k) { s->rx[0][k] = 0; s->ry[0][k] = 0; }
        for (size_t i = 0; i < K; ++i) {
            const size_t cur = lo.L[i], nxt = lo.L[i + 1];
            const double *W = theta + lo.w[i];
            const double *VW = Input + lo.w[i], *VB = Input + lo.b[i];
            for (size_t j = 0; j < nxt; ++j) {
                double rx = VB[j];
                for (size_t k = 0; k < cur; ++k) {
                    rx += s->ry[i][k] * W[k * nxt + j];
                    rx += s->y[i][k] * VW[k * nxt + j];
                }
                const double y = s->y[i + 1][j];
                double ry = rx;
                switch (ac[i + 1]) {
                    case 't': ry = rx * (1 - y * y); break;
                    case 'o': ry = 0.1 * rx; break;
                    case 's': ry = rx * y * (1 - y); break;
                    default: break;
                }
                s->rx[i + 1][j] = rx; s->ry[i + 1][j] = ry;
            }
        }
        for (size_t j = 0; j < A; ++j) RStd[j] = Std[j] * Input[lo.logstd + j];
        /* R-backward (TRPO_FVP.c:426-490) */
        for (size_t j = 0; j < A; ++j) {
            const double sq = Std[j] * Std[j];
            s->rg[K][j] = s->ry[K][j] / sq - 2 * s->g[K][j] / Std[j] * RStd[j];
            rgw[lo.logstd + j] = 2 * RStd[j] / Std[j];
        }
        for (size_t i = K; i > 0; --i) {
            const size_t cur = lo.L[i], prv = lo.L[i - 1];
            const double *W = theta + lo.w[i - 1], *VW = Input + lo.w[i - 1];
            for (size_t j = 0; j < cur; ++j) {
                const double y = s->y[i][j];
                switch (ac[i]) {
                    case 't': s->rg[i][j] = (1 - y * y) * s->rg[i][j] - 2 * y * s->g[i][j] * s->rx[i][j]; break;
                    case 'o': s->rg[i][j] = 0.1 * s->rg[i][j]; break;
                    case 's': s->rg[i][j] = s->rg[i][j] * y * (1 - y) + s->g[i][j] * (1 - 2 * y) * s->rx[i][j]; break;
                    default: break;
                }
                rgw[lo.b[i - 1] + j] = s->rg[i][j];
            }
            for (size_t j = 0; j < prv; ++j)
                for (size_t k = 0; k < cur; ++k)
                    rgw[lo.w[i - 1] + j * cur + k] = s->y[i - 1][j] * s->rg[i][k] + s->ry[i - 1][j] * s->g[i][k];
            for (size_t j = 0; j < prv; ++j) {
                s->rg[i - 1][j] = 0;
                for (size_t k = 0; k < cur; ++k) {
                    s->rg[i - 1][j] += VW[j * cur + k] * s->g[i][k];
                    s->rg[i - 1][j] += W[j * cur + k] * s->rg[i][k];
                }
            }
        }
        for (size_t q = 0; q < lo.P; ++q) Result[q] += rgw[q];
    }
    for (size_t q = 0; q < lo.P; ++q) {   /* TRPO_FVP.c:524-527 */
        Result[q] = Result[q] / (double)N;
        Result[q] += damping * Input[q];
    }
    free(rgw); free(RStd);
    scratch_free(s, &lo);
    return 0;
}

/* CG core shared by oracle_cg and oracle_update (TRPO_CG.c:32-107 == TRPO_Update.c:391-628). */
static int cg_core(const Layout *lo, const char *ac, const double *theta, const double *Std, const double *Observ,
                   size_t N, double damping, const double *b, size_t MaxIter, double ResidualTh,
                   double *x, double *rdotr_trace, double *xnorm_trace, Scratch *s) {
    const size_t P = lo->P;
    double *p = (double *)calloc(P, sizeof(double));
    double *r = (double *)calloc(P, sizeof(double));
    double *z = (double *)calloc(P, sizeof(double));
    memset(x, 0, P * sizeof(double));
    double rdotr = 0;
    for (size_t i = 0; i < P; ++i) { p[i] = b[i]; r[i] = b[i]; rdotr += r[i] * r[i]; }
    int nfvp = 0;
    for (size_t it = 0; it <= MaxIter; ++it) {
        double nrm = 0;
        for (size_t i = 0; i < P; ++i) nrm += x[i] * x[i];
        nrm = sqrt(nrm);
        if (rdotr_trace) rdotr_trace[it] = rdotr;
        if (xnorm_trace) xnorm_trace[it] = nrm;
        if (rdotr < ResidualTh || it == MaxIter) break;
        memset(z, 0, P * sizeof(double));
        fvp_fast_accumulate(lo, ac, theta, Std, Observ, N, p, z, s);
        fvp_finalise(P, N, damping, p, z);
        ++nfvp;
        double pdotz = 0;
        for (size_t i = 0; i < P; ++i) pdotz += p[i] * z[i];
        const double v = rdotr / pdotz;
        for (size_t i = 0; i < P; ++i) { x[i] += v * p[i]; r[i] -= v * z[i]; }
        double newrdotr = 0;
        for (size_t i = 0; i < P; ++i) newrdotr += r[i] * r[i];
        const double mu = newrdotr / rdotr;
        for (size_t i = 0; i < P; ++i) p[i] = r[i] + mu * p[i];
        rdotr = newrdotr;
    }
    free(p); free(r); free(z);
    return nfvp;
}

int oracle_cg(const OracleNet *net, const double *theta, const double *Std, const double *Observ,
              size_t N, double damping, const double *b, size_t MaxIter, double ResidualTh,
              double *Result, double *rdotr_trace, double *xnorm_trace) {
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    Scratch s; scratch_alloc(&s, &lo);
    int n = cg_core(&lo, net->AcFunc, theta, Std, Observ, N, damping, b, MaxIter, ResidualTh, Result,
                    rdotr_trace, xnorm_trace, &s);
    scratch_free(&s, &lo);
    return n;
}

int oracle_policy_gradient(const OracleNet *net, const double *theta, const double *Observ, const double *Mean,
                           const double *Action, const double *Advantage, size_t N, double *b) {
    /* TRPO_Update.c:254-378 */
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    const char *ac = net->AcFunc;
    const size_t K = lo.K, O = lo.L[0], A = lo.L[K];
    const double *LogStd = theta + lo.logstd;
    Scratch sc; scratch_alloc(&sc, &lo); Scratch *s = &sc;
    double *gw = (double *)calloc(lo.P, sizeof(double));
    memset(b, 0, lo.P * sizeof(double));
    for (size_t n = 0; n < N; ++n) {
        memcpy(s->y[0], Observ + n * O, O * sizeof(double));
        forward_one(&lo, ac, theta, s);
        for (size_t j = 0; j < A; ++j) {
            const double t = (Action[n * A + j] - Mean[n * A + j]) / exp(LogStd[j]);
            s->g[K][j] = Advantage[n] * t / exp(LogStd[j]);
            gw[lo.logstd + j] = Advantage[n] * (t * t - 1);
        }
        for (size_t i = K; i > 0; --i) {
            const size_t cur = lo.L[i], prv = lo.L[i - 1];
            const double *W = theta + lo.w[i - 1];
            for (size_t j = 0; j < cur; ++j) {
                const double y = s->y[i][j];
                switch (ac[i]) {
                    case 't': s->g[i][j] = s->g[i][j] * (1 - y * y); break;
                    case 'o': s->g[i][j] = 0.1 * s->g[i][j]; break;
                    case 's': s->g[i][j] = s->g[i][j] * y * (1 - y); break;
                    default: break;
                }
                gw[lo.b[i - 1] + j] = s->g[i][j];
            }
            for (size_t j = 0; j < prv; ++j)
                for (size_t k = 0; k < cur; ++k) gw[lo.w[i - 1] + j * cur + k] = s->g[i][k] * s->y[i - 1][j];
            for (size_t j = 0; j < prv; ++j) {
                s->g[i - 1][j] = 0;
                for (size_t k = 0; k < cur; ++k) s->g[i - 1][j] += s->g[i][k] * W[j * cur + k];
            }
        }
        for (size_t q = 0; q < lo.P; ++q) b[q] += gw[q];
    }
    for (size_t q = 0; q < lo.P; ++q) b[q] = b[q] / (double)N;
    free(gw);
    scratch_free(s, &lo);
    return 0;
}

int oracle_update(const OracleNet *net, const double *theta0, const double *Std, const double *Observ,
                  const double *Mean, const double *Action, const double *Advantage, size_t N,
                  double damping, double *Result, OracleUpdateInfo *info) {
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    const char *ac = net->AcFunc;
    const size_t K = lo.K, O = lo.L[0], A = lo.L[K], P = lo.P;
    /* constants, TRPO_Update.c:29-33 */
    const double ResidualTh = 1e-10, MaxKL = 0.01, AcceptRatio = 0.1;
    const size_t MaxIter = 10, MaxBackTracks = 10;
    OracleUpdateInfo local; if (!info) info = &local;
    memset(info, 0, sizeof(*info));

    double *b = (double *)calloc(P, sizeof(double));
    double *x = (double *)calloc(P, sizeof(double));
    double *z = (double *)calloc(P, sizeof(double));
    double *fullstep = (double *)calloc(P, sizeof(double));
    double *theta = (double *)calloc(P, sizeof(double));
    double *xnew = (double *)calloc(P, sizeof(double));
    Scratch sc; scratch_alloc(&sc, &lo); Scratch *s = &sc;

    oracle_policy_gradient(net, theta0, Observ, Mean, Action, Advantage, N, b);
    info->cg_iters = cg_core(&lo, ac, theta0, Std, Observ, N, damping, b, MaxIter, ResidualTh, x,
                             info->cg_rdotr, info->cg_xnorm, s);
    /* one more FVP on the solution (TRPO_Update.c:633-810) */
    fvp_fast_accumulate(&lo, ac, theta0, Std, Observ, N, x, z, s);
    fvp_finalise(P, N, damping, x, z);
    double shs = 0;
    for (size_t i = 0; i < P; ++i) shs += z[i] * x[i];
    shs = shs * 0.5;
    const double lm = sqrt(shs / MaxKL);
    double gnorm = 0;
    for (size_t i = 0; i < P; ++i) gnorm += b[i] * b[i];
    gnorm = sqrt(gnorm);
    for (size_t i = 0; i < P; ++i) fullstep[i] = x[i] / lm;
    double neggdotstepdir = 0;
    for (size_t i = 0; i < P; ++i) neggdotstepdir += b[i] * x[i];
    for (size_t i = 0; i < P; ++i) theta[i] = x[i];     /* the reference's fallback quirk, TRPO_Update.c:852 */
    const double rate = neggdotstepdir / lm;
    memcpy(x, theta0, P * sizeof(double));              /* x <- current parameters, :862-880 */
    double fval = 0;
    for (size_t n = 0; n < N; ++n) fval += Advantage[n];
    fval = -fval / (double)N;
    info->shs = shs; info->lm = lm; info->gnorm = gnorm; info->fval = fval;

    for (size_t t = 0; t < MaxBackTracks; ++t) {
        const double stepfrac = pow(0.5, (double)t);
        for (size_t i = 0; i < P; ++i) xnew[i] = x[i] + stepfrac * fullstep[i];
        const double *LogStd = xnew + lo.logstd;
        double surr = 0;
        for (size_t n = 0; n < N; ++n) {
            memcpy(s->y[0], Observ + n * O, O * sizeof(double));
            forward_one(&lo, ac, xnew, s);
            double lld = 0;
            for (size_t j = 0; j < A; ++j) {
                const double tx = (Action[n * A + j] - Mean[n * A + j]) / Std[j];
                const double tn = (Action[n * A + j] - s->y[K][j]) / exp(LogStd[j]);
                lld += tx * tx - tn * tn + log(Std[j]) - LogStd[j];
            }
            lld = lld * 0.5;
            surr += exp(lld) * Advantage[n];
        }
        const double newfval = -surr / (double)N;
        const double actual = fval - newfval;
        const double expected = rate * stepfrac;
        const double ratio = actual / expected;
        info->ls_actual[t] = actual; info->ls_expected[t] = expected; info->ls_ratio[t] = ratio;
        info->ls_steps = (int)t + 1;
        if (ratio > AcceptRatio && actual > 0) {
            memcpy(theta, xnew, P * sizeof(double));
            info->ls_accepted = 1;
            break;
        }
    }
    memcpy(Result, theta, P * sizeof(double));
    free(b); free(x); free(z); free(fullstep); free(theta); free(xnew);
    scratch_free(s, &lo);
    return 0;
}
