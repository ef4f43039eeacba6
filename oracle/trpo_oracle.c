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

/* ------------------------------------------------------------------------------------------------------------------
 * Rows f-3 / f-4 of SURVEY.md section 8: what sits either side of the update inside the reference's training loop
 * (TRPO_Lightweight.c:349-694, TRPO_Baseline.c:29-237).
 * ------------------------------------------------------------------------------------------------------------------ */

/* Value-function ("baseline") input of step `t`: the observation followed by t / EpLen (TRPO_Baseline.c:98-103). */
static void vf_input(const Layout *lo, const double *Observ, size_t pos, size_t step, size_t EpLen, Scratch *s) {
    const size_t O = lo->L[0] - 1;
    for (size_t i = 0; i < O; ++i) s->y[0][i] = Observ[pos * O + i];
    s->y[0][O] = (double)step / (double)EpLen;
}

int oracle_vf_predict(const OracleNet *vfnet, size_t NumEpBatch, size_t EpLen, const double *Observ, const double *x,
                      double *Baseline) {
    /* TRPO_Lightweight.c:582-625 */
    Layout lo; if (make_layout(vfnet, &lo) || check_acfunc(vfnet)) return -1;
    Scratch s; scratch_alloc(&s, &lo);
    for (size_t ep = 0; ep < NumEpBatch; ++ep)
        for (size_t t = 0; t < EpLen; ++t) {
            const size_t pos = ep * EpLen + t;
            vf_input(&lo, Observ, pos, t, EpLen, &s);
            forward_one(&lo, vfnet->AcFunc, x, &s);
            Baseline[pos] = s.y[lo.K][0];
        }
    scratch_free(&s, &lo);
    return 0;
}

double oracle_vf_evaluate(const OracleNet *vfnet, size_t NumEpBatch, size_t EpLen, const double *Observ,
                          const double *Target, const double *x, double *g, int n, double *Predict) {
    /* TRPO_Baseline.c:29-237: 0.01*MSE + 0.001*|x|^2 and its gradient; x holds W,B per layer (no LogStd). */
    Layout lo; if (make_layout(vfnet, &lo) || check_acfunc(vfnet)) return -1;
    const char *ac = vfnet->AcFunc;
    const size_t K = lo.K, NP = lo.logstd, N = NumEpBatch * EpLen;
    Scratch sc; scratch_alloc(&sc, &lo); Scratch *s = &sc;
    double *gw = (double *)calloc(NP, sizeof(double));
    for (int i = 0; i < n; ++i) g[i] = 0;
    for (size_t ep = 0; ep < NumEpBatch; ++ep)
        for (size_t t = 0; t < EpLen; ++t) {
            const size_t pos = ep * EpLen + t;
            vf_input(&lo, Observ, pos, t, EpLen, s);
            forward_one(&lo, ac, x, s);
            Predict[pos] = s->y[K][0];
            s->g[K][0] = 0.02 * (Predict[pos] - Target[pos]);
            for (size_t i = K; i > 0; --i) {
                const size_t cur = lo.L[i], prv = lo.L[i - 1];
                const double *W = x + lo.w[i - 1];
                for (size_t j = 0; j < cur; ++j) {
                    if (ac[i] == 't') s->g[i][j] = s->g[i][j] * (1 - s->y[i][j] * s->y[i][j]);
                    gw[lo.b[i - 1] + j] = s->g[i][j];
                }
                for (size_t j = 0; j < prv; ++j)
                    for (size_t k = 0; k < cur; ++k) gw[lo.w[i - 1] + j * cur + k] = s->g[i][k] * s->y[i - 1][j];
                for (size_t j = 0; j < prv; ++j) {
                    s->g[i - 1][j] = 0;
                    for (size_t k = 0; k < cur; ++k) s->g[i - 1][j] += s->g[i][k] * W[j * cur + k];
                }
            }
            for (size_t q = 0; q < NP; ++q) g[q] += gw[q];
        }
    for (size_t q = 0; q < NP; ++q) g[q] = g[q] / (double)N + 0.002 * x[q];
    double mse = 0;
    for (size_t i = 0; i < N; ++i) mse += 0.01 * (Predict[i] - Target[i]) * (Predict[i] - Target[i]);
    mse = mse / (double)N;
    double l2 = 0;
    for (size_t q = 0; q < NP; ++q) l2 += x[q] * x[q];
    free(gw);
    scratch_free(s, &lo);
    return mse + 0.001 * l2;
}

void oracle_reward_stats(size_t NumEpBatch, size_t EpLen, const double *Reward, double *EpRewMean, double *EpRewStd) {
    /* TRPO_Lightweight.c:545-558 */
    double mean = 0;
    for (size_t i = 0; i < NumEpBatch * EpLen; ++i) mean += Reward[i];
    mean = mean / (double)NumEpBatch;
    double var = 0;
    for (size_t ep = 0; ep < NumEpBatch; ++ep) {
        double r = 0;
        for (size_t t = 0; t < EpLen; ++t) r += Reward[ep * EpLen + t];
        var += (r - mean) * (r - mean);
    }
    *EpRewMean = mean;
    *EpRewStd = sqrt(var / (double)NumEpBatch);
}

int oracle_gae(size_t NumEpBatch, size_t EpLen, double gamma, double lam, double *Reward, const double *Baseline,
               double *Return, double *Advantage) {
    /* TRPO_Lightweight.c:565-653, the reference's O(EpLen^2) pow() sums kept as they are; Reward is left holding the
     * TD residuals, as in the reference. */
    const size_t N = NumEpBatch * EpLen;
    for (size_t ep = 0; ep < NumEpBatch; ++ep) {
        double *R = Reward + ep * EpLen;
        const double *V = Baseline + ep * EpLen;
        for (size_t t = 0; t < EpLen; ++t) {
            double ret = R[t];
            for (size_t f = t + 1; f < EpLen; ++f) ret += R[f] * pow(gamma, (double)(int)(f - t));
            Return[ep * EpLen + t] = ret;
        }
        for (size_t t = 0; t + 1 < EpLen; ++t) R[t] += gamma * V[t + 1] - V[t];
        R[EpLen - 1] += (-1) * V[EpLen - 1];
        for (size_t t = 0; t < EpLen; ++t) {
            double adv = R[t];
            for (size_t f = t + 1; f < EpLen; ++f) adv += R[f] * pow(gamma * lam, (double)(int)(f - t));
            Advantage[ep * EpLen + t] = adv;
        }
    }
    double mean = 0;
    for (size_t i = 0; i < N; ++i) mean += Advantage[i];
    mean = mean / (double)N;
    double sd = 0;
    for (size_t i = 0; i < N; ++i) sd += (Advantage[i] - mean) * (Advantage[i] - mean);
    sd = sqrt(sd / (double)N);
    for (size_t i = 0; i < N; ++i) Advantage[i] = (Advantage[i] - mean) / sd;
    return 0;
}

/* Lightweight arm simulator (TRPO_Lightweight.c:349-540): three joint angles integrate the sampled action, forward
 * kinematics give the link positions, reward = -100*|object - grip|^2 - |action|^2. Draws from rand() in the
 * reference's order (3 per episode for the object, 2 per action component per step). */
typedef struct { double dof2[3], wrist[3], grip[3]; } ArmPose;

static void arm_kinematics(double t1, double t2, double t3, ArmPose *p) {
    const double s1 = sin(t1), c1 = cos(t1), s2 = sin(t2), c2 = cos(t2), s3 = sin(t3), c3 = cos(t3);
    const double c2c3 = c2 * c3, s2s3 = s2 * s3, c2s3 = c2 * s3, s2c3 = s2 * c3;
    p->dof2[0] = 0.0575 * c1 * c2;
    p->dof2[1] = 0.0575 * s1 * c2;
    p->dof2[2] = 0.01768 - 0.0575 * s2;
    p->wrist[0] = p->dof2[0] + 0.07375 * c1 * (c2c3 - s2s3);
    p->wrist[1] = p->dof2[1] + 0.07375 * s1 * (c2c3 - s2s3);
    p->wrist[2] = p->dof2[2] - 0.07375 * (c2s3 + s2c3);
    p->grip[0] = p->dof2[0] + 0.11315 * c1 * (c2c3 - s2s3) - 0.0125 * c1 * (c2s3 + s2c3);
    p->grip[1] = p->dof2[1] + 0.11315 * s1 * (c2c3 - s2s3) - 0.0125 * s1 * (c2s3 + s2c3);
    p->grip[2] = p->dof2[2] - 0.11315 * (c2s3 + s2c3) + 0.0125 * (s2s3 - c2c3);
}

int oracle_arm_rollout(const OracleNet *net, const double *theta, size_t NumEpBatch, size_t EpLen,
                       double *Observ, double *Mean, double *Std, double *Action, double *Reward) {
    Layout lo; if (make_layout(net, &lo) || check_acfunc(net)) return -1;
    const size_t K = lo.K, O = lo.L[0], A = lo.L[K];
    if (O != 15 || A != 3) return -1;
    const double pi = 3.1415926535897931, TimeStepLen = 0.02, coeff = 1;
    const double *LogStd = theta + lo.logstd;
    Scratch s; scratch_alloc(&s, &lo);
    for (size_t ep = 0; ep < NumEpBatch; ++ep) {
        double t1 = 0, t2 = -pi / 2.0, t3 = pi / 2.0;
        const double dof1[3] = {0, 0, 0.01768};
        ArmPose p = {{0, 0, 0.07518}, {0.07375, 0, 0.07518}, {0.11315, 0, 0.06268}};
        double obj[3];
        obj[0] = ((double)rand() / (double)RAND_MAX) * 0.076 + 0.084;
        obj[1] = ((double)rand() / (double)RAND_MAX) * 0.100 - 0.05;
        obj[2] = ((double)rand() / (double)RAND_MAX) * 0.100;
        for (size_t t = 0; t < EpLen; ++t) {
            const size_t row = ep * EpLen + t;
            double *ob = Observ + row * O;
            for (int i = 0; i < 3; ++i) {
                ob[i] = dof1[i]; ob[3 + i] = p.dof2[i]; ob[6 + i] = p.wrist[i]; ob[9 + i] = p.grip[i]; ob[12 + i] = obj[i];
            }
            memcpy(s.y[0], ob, O * sizeof(double));
            forward_one(&lo, net->AcFunc, theta, &s);
            double ac[3];
            for (size_t i = 0; i < A; ++i) Mean[row * A + i] = s.y[K][i];
            for (size_t i = 0; i < A; ++i) Std[i] = exp(LogStd[i]);
            for (size_t i = 0; i < A; ++i) {
                const double u1 = ((double)rand() + 1.0) / ((double)RAND_MAX + 1.0);
                const double u2 = ((double)rand() + 1.0) / ((double)RAND_MAX + 1.0);
                const double z0 = sqrt(-2.0 * log(u1)) * cos(2 * pi * u2);
                ac[i] = z0 * Std[i] + s.y[K][i];
                Action[row * A + i] = ac[i];
            }
            t1 += ac[0] * coeff * TimeStepLen;
            t2 += ac[1] * coeff * TimeStepLen;
            t3 += ac[2] * coeff * TimeStepLen;
            arm_kinematics(t1, t2, t3, &p);
            double re = 0;
            for (int i = 0; i < 3; ++i) {
                re -= 100 * (obj[i] - p.grip[i]) * (obj[i] - p.grip[i]);
                re -= ac[i] * ac[i];
            }
            Reward[row] = re;
        }
    }
    scratch_free(&s, &lo);
    return 0;
}
