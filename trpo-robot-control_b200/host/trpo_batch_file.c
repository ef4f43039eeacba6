/*
 * trpo_batch_file.c -- binary rollout-batch file (include/trpo_b200.h), the replacement for the reference's text data
 * file: TRPO_FVP.c:731-762 fscanf's N x (3A + O + 1) decimal numbers on EVERY FVP call (8x per CG on ArmTest), which
 * dominates wall time for large N. Here every array is one contiguous section of little-endian doubles behind a
 * 64-byte header, so staging is a handful of large reads straight into (pinned) buffers.
 *
 *   header | Std[A] | Observ[N*O] | Mean[N*A] | Action[N*A] | Advantage[N]      (the last three only with flag bit 0)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/trpo_b200.h"

static const char MAGIC[8] = {'T', 'R', 'P', 'O', 'B', '2', '0', '0'};

int trpo_batch_file_probe(const char *path, trpo_batch_file_header *hdr) {
    memset(hdr, 0, sizeof(*hdr));
    FILE *f = fopen(path, "rb");
    if (f == NULL) return -1;
    trpo_batch_file_header h;
    const size_t got = fread(&h, 1, sizeof(h), f);
    fclose(f);
    if (got != sizeof(h) || memcmp(h.magic, MAGIC, 8) != 0 || h.version != 1) return -1;
    *hdr = h;
    return 0;
}

int trpo_batch_file_write(const char *path, size_t N, size_t O, size_t A, const double *Observ, const double *Std,
                          const double *Mean, const double *Action, const double *Advantage) {
    if (!path || !Observ || !Std || N == 0 || O == 0 || A == 0) return -1;
    const int full = Mean && Action && Advantage;
    FILE *f = fopen(path, "wb");
    if (f == NULL) {
        fprintf(stderr, "[ERROR] Cannot open Data File [%s]. \n", path);
        return -1;
    }
    trpo_batch_file_header h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, MAGIC, 8);
    h.version = 1;
    h.flags = full ? 1u : 0u;
    h.NumSamples = N; h.ObservSpaceDim = O; h.ActionSpaceDim = A;
    int ok = fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && fwrite(Std, sizeof(double), A, f) == A;
    ok = ok && fwrite(Observ, sizeof(double), N * O, f) == N * O;
    if (full) {
        ok = ok && fwrite(Mean, sizeof(double), N * A, f) == N * A;
        ok = ok && fwrite(Action, sizeof(double), N * A, f) == N * A;
        ok = ok && fwrite(Advantage, sizeof(double), N, f) == N;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : -1;
}

int trpo_batch_file_read(const char *path, size_t N, double *Observ, double *Std, double *Mean, double *Action,
                         double *Advantage) {
    trpo_batch_file_header h;
    if (trpo_batch_file_probe(path, &h)) return -1;
    if (N == 0) N = (size_t)h.NumSamples;
    if (N > h.NumSamples) {
        fprintf(stderr, "[ERROR] Data File [%s] holds fewer than %zu samples. \n", path, N);
        return -1;
    }
    const size_t O = (size_t)h.ObservSpaceDim, A = (size_t)h.ActionSpaceDim, NF = (size_t)h.NumSamples;
    FILE *f = fopen(path, "rb");
    if (f == NULL) return -1;
    long off = (long)sizeof(h);
    int ok = fseek(f, off, SEEK_SET) == 0 && fread(Std, sizeof(double), A, f) == A;
    off += (long)(A * sizeof(double));
    ok = ok && fseek(f, off, SEEK_SET) == 0 && fread(Observ, sizeof(double), N * O, f) == N * O;
    off += (long)(NF * O * sizeof(double));
    if (Mean && Action && Advantage) {
        if (!(h.flags & 1u)) { fclose(f); return -1; }
        ok = ok && fseek(f, off, SEEK_SET) == 0 && fread(Mean, sizeof(double), N * A, f) == N * A;
        off += (long)(NF * A * sizeof(double));
        ok = ok && fseek(f, off, SEEK_SET) == 0 && fread(Action, sizeof(double), N * A, f) == N * A;
        off += (long)(NF * A * sizeof(double));
        ok = ok && fseek(f, off, SEEK_SET) == 0 && fread(Advantage, sizeof(double), N, f) == N;
    }
    fclose(f);
    return ok ? 0 : -1;
}

int trpo_batch_file_from_text(const char *text_path, const char *bin_path, size_t N, size_t O, size_t A) {
    /* row format of TRPO_FVP.c:739-760: Mean[A] Std[A] Observ[O] Action[A] Advantage; the last row's Std wins */
    FILE *df = fopen(text_path, "r");
    if (df == NULL) {
        fprintf(stderr, "[ERROR] Cannot open Data File [%s]. \n", text_path);
        return -1;
    }
    double *Mean = (double *)calloc(N * A, sizeof(double)), *Std = (double *)calloc(A, sizeof(double));
    double *Observ = (double *)calloc(N * O, sizeof(double)), *Action = (double *)calloc(N * A, sizeof(double));
    double *Advantage = (double *)calloc(N, sizeof(double));
    int ok = Mean && Std && Observ && Action && Advantage;
    for (size_t n = 0; n < N && ok; ++n) {
        for (size_t j = 0; j < A; ++j) ok &= fscanf(df, "%lf", &Mean[n * A + j]) == 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(df, "%lf", &Std[j]) == 1;
        for (size_t j = 0; j < O; ++j) ok &= fscanf(df, "%lf", &Observ[n * O + j]) == 1;
        for (size_t j = 0; j < A; ++j) ok &= fscanf(df, "%lf", &Action[n * A + j]) == 1;
        ok &= fscanf(df, "%lf", &Advantage[n]) == 1;
    }
    fclose(df);
    if (!ok) fprintf(stderr, "[ERROR] Data File [%s] holds fewer than %zu samples. \n", text_path, N);
    const int rc = ok ? trpo_batch_file_write(bin_path, N, O, A, Observ, Std, Mean, Action, Advantage) : -1;
    free(Mean); free(Std); free(Observ); free(Action); free(Advantage);
    return rc;
}
