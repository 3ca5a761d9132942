/*
 * slam_oracle.c -- TEST INFRASTRUCTURE ONLY (see slam_oracle.h).  PARITY UNPINNED.
 *
 * Plain-C restatement of the reference's EKF-SLAM / UKF-SLAM Filter::update() and of the
 * simulator's measurement generator.  Every float/double rounding of the reference's C++
 *  is reproduced; "dense" mode evaluates the same O(n^3) matrix
 * products the reference evaluates with Eigen (naive k-ascending inner products), and
 * "structured" mode evaluates the same arithmetic with the exactly-zero terms skipped.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -pthread (oracle/Makefile).
 *
 * Deviations from the reference that cannot be avoided here (documented in DESIGN.md):
 *   D-1  cos/sin of a *float* argument (libstdc++ overload -> cosf/sinf) is evaluated as
 *        (float)cos((double)x) unless oracle_set_trig_mode(1) selects libm cosf/sinf.
 *   D-3  Eigen's unsupported MatrixFunctions sqrt (real Schur) is restated as a symmetric
 *        eigen-decomposition sqrt of the same matrix.
 *   D-4  std::pow(x,2) is restated as x*x; Python's (..)**(1/2) as sqrt().
 */
#include "slam_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PI_REF 3.14159265358979323846 /* filter.h:42 */

static int g_trig_mode = 0;
void oracle_set_trig_mode(int mode) { g_trig_mode = mode; }
static float cos_f(float x) { return g_trig_mode ? cosf(x) : (float)cos((double)x); }
static float sin_f(float x) { return g_trig_mode ? sinf(x) : (float)sin((double)x); }

struct oracle_filter {
    int kind, base;              /* base = 3 (EKF) or 4 (UKF) */
    oracle_params p;
    double V[2][2], W[2][2];     /* filter.h:87,91 after readCommonParams */
    int cap_lm, ncap;            /* capacity in landmarks / state entries */
    int M, timestep, status;
    int M_commit;                /* landmarks in the committed x_t/P_t */
    int in_step;                 /* predict done, measure pending (split form) */
    double *x_t, *P_t, *x_pred, *P_pred;   /* ld = ncap, row-major */
    int *lm_ids;
    /* work (dense mode + UKF) */
    double *w1, *w2, *w3, *w4;   /* ncap x ncap each */
    double *X, *X_pred;          /* UKF sigma points, column j at X + j*ncap   (n x (2n+1)) */
    double *Wts;                 /* 2*ncap+1 */
    double *Qd;                  /* UKF process noise diagonal (only entries 0..3 ever non-zero) */
    int X_rows;                  /* current row count of X (ukf.cpp:169) */
    double *vec1, *vec2;
    int *assoc; int n_assoc, assoc_cap;
    float *map; int n_map;       /* UKF_LOC: the true map, [id, x, y]* float32 (filter.h:68, /truth/landmarks) */
};

/* ------------------------------------------------------------------ helpers */
static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

/* C(m x n) = A(m x kk) * B(kk x n), naive k-ascending inner products */
static void mm(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int kk, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < kk; ++k) s += A[i * lda + k] * B[k * ldb + j];
            C[i * ldc + j] = s;
        }
}
/* C(m x n) = A(m x kk) * B^T, B is n x kk */
static void mmT(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int kk, int n) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k < kk; ++k) s += A[i * lda + k] * B[j * ldb + k];
            C[i * ldc + j] = s;
        }
}

/* Eigen dynamic-size inverse() = PartialPivLU (ekf.cpp:135, ukf.cpp:339), 2x2 case */
static void inv2_lu(const double S[2][2], double Si[2][2]) {
    double a[2][2] = {{S[0][0], S[0][1]}, {S[1][0], S[1][1]}};
    int perm0 = 0, perm1 = 1;
    if (fabs(a[1][0]) > fabs(a[0][0])) { /* first maximum wins */
        double t0 = a[0][0], t1 = a[0][1];
        a[0][0] = a[1][0]; a[0][1] = a[1][1]; a[1][0] = t0; a[1][1] = t1;
        perm0 = 1; perm1 = 0;
    }
    double l10 = a[1][0] / a[0][0];
    double u00 = a[0][0], u01 = a[0][1], u11 = a[1][1] - l10 * a[0][1];
    for (int c = 0; c < 2; ++c) {
        double b0 = (perm0 == c) ? 1.0 : 0.0, b1 = (perm1 == c) ? 1.0 : 0.0;
        double y0 = b0, y1 = b1 - l10 * y0;
        double x1 = y1 / u11;
        double x0 = (y0 - u01 * x1) / u00;
        Si[0][c] = x0; Si[1][c] = x1;
    }
}

/* ------------------------------------------------------------------ eigh: Householder tridiagonalisation + implicit QL
 * (the classic EISPACK tred2/tql2 pair; Eigen's SelfAdjointEigenSolver, ukf.cpp:116-118, is the same family:
 *  Householder tridiagonalisation followed by implicit symmetric QR).  V is n x n row-major, ld = n. */
static void tred2(int n, double* V, double* d, double* e) {
#define VV(i, j) V[(i) * n + (j)]
    for (int j = 0; j < n; ++j) d[j] = VV(n - 1, j);
    for (int i = n - 1; i > 0; --i) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; ++k) scale += fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; ++j) { d[j] = VV(i - 1, j); VV(i, j) = 0.0; VV(j, i) = 0.0; }
        } else {
            for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1];
            double g = sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; ++j) e[j] = 0.0;
            for (int j = 0; j < i; ++j) {
                f = d[j];
                VV(j, i) = f;
                g = e[j] + VV(j, j) * f;
                for (int k = j + 1; k <= i - 1; ++k) { g += VV(k, j) * d[k]; e[k] += VV(k, j) * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
            double hh = f / (h + h);
            for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
            for (int j = 0; j < i; ++j) {
                f = d[j]; g = e[j];
                for (int k = j; k <= i - 1; ++k) VV(k, j) -= (f * e[k] + g * d[k]);
                d[j] = VV(i - 1, j);
                VV(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; ++i) {
        VV(n - 1, i) = VV(i, i);
        VV(i, i) = 1.0;
        double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; ++k) d[k] = VV(k, i + 1) / h;
            for (int j = 0; j <= i; ++j) {
                double g = 0.0;
                for (int k = 0; k <= i; ++k) g += VV(k, i + 1) * VV(k, j);
                for (int k = 0; k <= i; ++k) VV(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) VV(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; ++j) { d[j] = VV(n - 1, j); VV(n - 1, j) = 0.0; }
    VV(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
}

static void tql2(int n, double* V, double* d, double* e) {
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; ++l) {
        double t = fabs(d[l]) + fabs(e[l]);
        if (t > tst1) tst1 = t;
        int m = l;
        while (m < n) { if (fabs(e[m]) <= eps * tst1) break; ++m; }
        if (m > l) {
            int iter = 0;
            do {
                ++iter;
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; ++i) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; --i) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; ++k) {
                        h = VV(k, i + 1);
                        VV(k, i + 1) = s * VV(k, i) + c * h;
                        VV(k, i) = c * VV(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (fabs(e[l]) > eps * tst1 && iter < 200);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    /* ascending order (SelfAdjointEigenSolver sorts ascending) */
    for (int i = 0; i < n - 1; ++i) {
        int k = i; double p = d[i];
        for (int j = i + 1; j < n; ++j) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
            for (int j = 0; j < n; ++j) { double t = VV(j, i); VV(j, i) = VV(j, k); VV(j, k) = t; }
        }
    }
#undef VV
}

/* A symmetric n x n (row-major, ld = n).  evecs: column k is the k-th eigenvector. */
void oracle_eigh(const double* A, int n, double* evals, double* evecs) {
    double* e = dalloc((size_t)n);
    memcpy(evecs, A, sizeof(double) * (size_t)n * n);
    if (n == 1) { evals[0] = A[0]; evecs[0] = 1.0; free(e); return; }
    tred2(n, evecs, evals, e);
    tql2(n, evecs, evals, e);
    free(e);
}

/* ------------------------------------------------------------------ create / init */
oracle_filter* oracle_create(int kind, const oracle_params* p, int max_landmarks) {
    if (kind != ORACLE_EKF_SLAM && kind != ORACLE_UKF_SLAM && kind != ORACLE_UKF_LOC && kind != ORACLE_NAIVE) return NULL;
    if (kind == ORACLE_UKF_LOC || kind == ORACLE_NAIVE) max_landmarks = 1;   /* the state never holds landmarks */
    oracle_filter* f = (oracle_filter*)calloc(1, sizeof(*f));
    f->kind = kind;
    f->base = (kind == ORACLE_EKF_SLAM || kind == ORACLE_NAIVE) ? 3 : 4;
    f->p = *p;
    /* Filter::readCommonParams, filter.h:105-121 */
    f->V[0][0] = 1; f->V[1][1] = 1; f->V[0][1] = f->V[1][0] = 0;     /* :107 */
    f->V[0][0] = p->V_00; f->V[1][1] = p->V_11;                       /* :110-111 */
    f->W[0][0] = 1; f->W[1][1] = 1; f->W[0][1] = f->W[1][0] = 0;     /* :113 */
    if (p->compat_noise_bug) { f->V[0][0] = p->W_00; f->V[1][1] = p->W_11; }  /* :116-117 (bug) */
    else { f->W[0][0] = p->W_00; f->W[1][1] = p->W_11; }
    f->cap_lm = max_landmarks;
    f->ncap = f->base + 2 * max_landmarks + 2; /* +2: room for an insertion's (n+2) temporaries */
    size_t nn = (size_t)f->ncap * f->ncap;
    f->x_t = dalloc(f->ncap); f->x_pred = dalloc(f->ncap);
    f->P_t = dalloc(nn); f->P_pred = dalloc(nn);
    f->w1 = dalloc(nn); f->w2 = dalloc(nn); f->w3 = dalloc(nn); f->w4 = dalloc(nn);
    f->vec1 = dalloc(8 * f->ncap + 16); f->vec2 = dalloc(8 * f->ncap + 16);
    f->lm_ids = (int*)calloc(max_landmarks + 1, sizeof(int));
    f->assoc_cap = 4096; f->assoc = (int*)calloc(f->assoc_cap, sizeof(int));
    int nb = f->base, ld = f->ncap;
    /* EKF::EKF ekf.cpp:8-20 / UKF::UKF ukf.cpp:7-22 */
    for (int i = 0; i < nb; ++i) { f->P_t[i * ld + i] = 1.0; f->P_pred[i * ld + i] = 1.0; }
    f->P_t[0] = f->P_pred[0] = 0.01 * 0.01;
    f->P_t[ld + 1] = f->P_pred[ld + 1] = 0.01 * 0.01;
    f->P_t[2 * ld + 2] = f->P_pred[2 * ld + 2] = 0.005 * 0.005;
    if (kind == ORACLE_UKF_SLAM || kind == ORACLE_UKF_LOC) {
        f->P_t[3 * ld + 3] = f->P_pred[3 * ld + 3] = 0.005 * 0.005;
        size_t ns = (size_t)f->ncap * (2 * f->ncap + 1);
        f->X = dalloc(ns); f->X_pred = dalloc(ns);
        f->Wts = dalloc(2 * f->ncap + 1);
        f->Qd = dalloc(f->ncap);
        f->X_rows = 4; /* ukf.cpp:20 */
    }
    return f;
}

void oracle_destroy(oracle_filter* f) {
    if (!f) return;
    free(f->x_t); free(f->x_pred); free(f->P_t); free(f->P_pred);
    free(f->w1); free(f->w2); free(f->w3); free(f->w4); free(f->vec1); free(f->vec2);
    free(f->lm_ids); free(f->assoc); free(f->X); free(f->X_pred); free(f->Wts); free(f->Qd); free(f->map);
    free(f);
}

/* the /truth/landmarks message the localisation-only UKF keeps as `map` (filter.h:68, localization_node.cpp:66-75) */
void oracle_set_map(oracle_filter* f, const float* map_id_x_y, int n_landmarks) {
    free(f->map);
    f->map = (float*)malloc(sizeof(float) * 3 * (size_t)(n_landmarks > 0 ? n_landmarks : 1));
    memcpy(f->map, map_id_x_y, sizeof(float) * 3 * (size_t)n_landmarks);
    f->n_map = n_landmarks;
}

void oracle_init(oracle_filter* f, float x0, float y0, float yaw0) {
    if (f->kind == ORACLE_EKF_SLAM || f->kind == ORACLE_NAIVE) {
        f->x_t[0] = x0; f->x_t[1] = y0; f->x_t[2] = yaw0;                /* ekf.cpp:31 */
    } else {
        f->x_t[0] = x0; f->x_t[1] = y0; f->x_t[2] = cos_f(yaw0); f->x_t[3] = sin_f(yaw0); /* ukf.cpp:33 */
        float w0 = 0.2f;                                                  /* filter.h:207 */
        float w = (1 - w0) / 8;                                           /* ukf.cpp:35 (float) */
        for (int i = 0; i < 9; ++i) f->Wts[i] = (double)w;
        f->Wts[0] = (double)w0;                                           /* ukf.cpp:36 */
        float yaw = (float)remainder(atan2(f->x_t[3], f->x_t[2]), 2 * PI_REF); /* ukf.cpp:38 */
        f->Qd[0] = f->V[0][0] * (double)cos_f(yaw);                       /* ukf.cpp:39-42 */
        f->Qd[1] = f->V[0][0] * (double)sin_f(yaw);
        f->Qd[2] = f->V[1][1] * (double)cos_f(yaw);
        f->Qd[3] = f->V[1][1] * (double)sin_f(yaw);
    }
}

void oracle_set_state(oracle_filter* f, const double* x, const double* P, const int* ids, int M, int timestep) {
    int n = f->base + 2 * M, ld = f->ncap;
    f->M = M; f->M_commit = M; f->timestep = timestep; f->in_step = 0;
    for (int i = 0; i < n; ++i) { f->x_t[i] = x[i]; for (int j = 0; j < n; ++j) f->P_t[i * ld + j] = P[i * n + j]; }
    for (int i = 0; i < M; ++i) f->lm_ids[i] = ids[i];
}

int oracle_state_dim(const oracle_filter* f) { return f->base + 2 * f->M_commit; }
int oracle_num_landmarks(const oracle_filter* f) { return f->M_commit; }
int oracle_timestep(const oracle_filter* f) { return f->timestep; }
int oracle_status(const oracle_filter* f) { return f->status; }
void oracle_get_state(const oracle_filter* f, double* x) { memcpy(x, f->x_t, sizeof(double) * oracle_state_dim(f)); }
void oracle_get_cov(const oracle_filter* f, double* P) {
    int n = oracle_state_dim(f), ld = f->ncap;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) P[i * n + j] = f->P_t[i * ld + j];
}
void oracle_get_landmark_ids(const oracle_filter* f, int* ids) { memcpy(ids, f->lm_ids, sizeof(int) * f->M_commit); }
int oracle_get_assoc_log(const oracle_filter* f, int* idx, int cap) {
    int k = f->n_assoc < cap ? f->n_assoc : cap;
    memcpy(idx, f->assoc, sizeof(int) * k);
    return f->n_assoc;
}
int oracle_sigma_rows(const oracle_filter* f) { return f->X ? f->X_rows : 0; }
void oracle_get_sigma_points(const oracle_filter* f, double* Xo) {
    if (!f->X) return;
    int n = f->X_rows;
    for (int j = 0; j < 2 * n + 1; ++j) for (int i = 0; i < n; ++i) Xo[j * n + i] = f->X[(size_t)j * f->ncap + i];
}

static void log_assoc(oracle_filter* f, int idx) {
    if (f->n_assoc < f->assoc_cap) f->assoc[f->n_assoc] = idx;
    f->n_assoc++;
}

static int has_nan(const oracle_filter* f) {
    int n = f->base + 2 * f->M, ld = f->ncap;
    for (int i = 0; i < n; ++i) {
        if (!isfinite(f->x_t[i])) return 1;
        for (int j = 0; j < n; ++j) if (!isfinite(f->P_t[i * ld + j])) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ EKF, ekf.cpp:37-179 */
static void ekf_predict(oracle_filter* f, float d_d, float d_th, int mode) {
    const int n = 3 + 2 * f->M, ld = f->ncap;
    f->timestep += 1;                                                   /* :39 */
    const double th = f->x_t[2];
    const double fx02 = (double)(-1 * d_d) * sin(th);                   /* :48 */
    const double fx12 = (double)d_d * cos(th);                          /* :49 */
    const double c = cos(th), s = sin(th);                              /* :52-53 */
    for (int i = 0; i < n; ++i) f->x_pred[i] = f->x_t[i];               /* :56 */
    const float dv = d_d + f->p.v_d;                                    /* float add, :57-58 */
    f->x_pred[0] = f->x_t[0] + (double)dv * cos(th);
    f->x_pred[1] = f->x_t[1] + (double)dv * sin(th);
    f->x_pred[2] = remainder(f->x_t[2] + (double)d_th + (double)f->p.v_th, 2 * PI_REF); /* :59 */
    double* P = f->P_t; double* Pp = f->P_pred;
    if (mode == ORACLE_DENSE) {
        double *Fx = f->w1, *T = f->w2, *Fv = f->w3, *FV = f->w4;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Fx[i * ld + j] = (i == j) ? 1.0 : 0.0; /* :47 */
        Fx[0 * ld + 2] = fx02; Fx[1 * ld + 2] = fx12;
        for (int i = 0; i < n; ++i) { Fv[i * 2] = 0.0; Fv[i * 2 + 1] = 0.0; }                     /* :51 */
        Fv[0] = c; Fv[2] = s; Fv[5] = 1.0;                                                       /* :52-54 */
        mm(T, ld, Fx, ld, P, ld, n, n, n);                 /* F_x * P_t */
        mmT(Pp, ld, T, ld, Fx, ld, n, n, n);               /* (..) * F_x^T */
        double Vm[4] = {f->V[0][0], f->V[0][1], f->V[1][0], f->V[1][1]};
        mm(FV, 2, Fv, 2, Vm, 2, n, 2, 2);                  /* F_v * V */
        mmT(T, ld, FV, 2, Fv, 2, n, 2, n);                 /* (..) * F_v^T */
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Pp[i * ld + j] += T[i * ld + j];  /* :61 */
    } else {
        /* T = F_x P : rows 0,1 pick up row 2 */
        double* T = f->w2;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) T[i * ld + j] = P[i * ld + j];
        for (int j = 0; j < n; ++j) {
            T[0 * ld + j] = P[0 * ld + j] + fx02 * P[2 * ld + j];
            T[1 * ld + j] = P[1 * ld + j] + fx12 * P[2 * ld + j];
        }
        /* P' = T F_x^T : cols 0,1 pick up col 2 */
        for (int i = 0; i < n; ++i) {
            for (int j = 2; j < n; ++j) Pp[i * ld + j] = T[i * ld + j];
            Pp[i * ld + 0] = T[i * ld + 0] + T[i * ld + 2] * fx02;
            Pp[i * ld + 1] = T[i * ld + 1] + T[i * ld + 2] * fx12;
        }
        /* + (F_v V) F_v^T, non-zero only in the 3x3 vehicle block */
        const double cV = c * f->V[0][0], sV = s * f->V[0][0];
        Pp[0 * ld + 0] += cV * c; Pp[0 * ld + 1] += cV * s;
        Pp[1 * ld + 0] += sV * c; Pp[1 * ld + 1] += sV * s;
        Pp[2 * ld + 2] += f->V[1][1];
    }
}

static void ekf_commit(oracle_filter* f) {
    const int n = 3 + 2 * f->M, ld = f->ncap;
    for (int i = 0; i < n; ++i) { f->x_t[i] = f->x_pred[i]; for (int j = 0; j < n; ++j) f->P_t[i * ld + j] = f->P_pred[i * ld + j]; }
    f->M_commit = f->M;
}

static int ekf_measure(oracle_filter* f, const float* lm_meas, int num_landmarks, int mode) {
    const int ld = f->ncap;
    const int M_start = f->M;        /* x_t (stale) still has 3+2*M_start entries, ekf.cpp:115 */
    double* Pp = f->P_pred; double* xp = f->x_pred;
    f->n_assoc = 0;
    if (num_landmarks < 1) { ekf_commit(f); return f->status; }           /* :67-71 */
    for (int l = 0; l < num_landmarks; ++l) {                            /* :73 */
        const float r = lm_meas[l * 3 + 1], b = lm_meas[l * 3 + 2];      /* :75-76 */
        int i = -1, id;
        if (!f->p.landmark_id_is_known) {                                /* :82-98 */
            id = f->M;
            const float x_detected = (float)(xp[0] + (double)r * cos(xp[2] + (double)b));
            const float y_detected = (float)(xp[1] + (double)r * sin(xp[2] + (double)b));
            for (int j = 0; j < f->M; ++j) {
                const float x_diff = (float)fabs((double)x_detected - xp[3 + 2 * j]);
                const float y_diff = (float)fabs((double)y_detected - xp[3 + 2 * j + 1]);
                if (x_diff < f->p.min_landmark_separation && y_diff < f->p.min_landmark_separation) { i = j; id = j; break; }
            }
        } else {                                                         /* :99-109 */
            id = (int)lm_meas[l * 3];
            for (int j = 0; j < f->M; ++j) if (f->lm_ids[j] == id) { i = j; break; }
        }
        log_assoc(f, i);
        const int n = 3 + 2 * f->M;
        if (i != -1) {
            /* ---------------- landmark update, :110-140 */
            if (i >= M_start) {           /* reference would index x_t out of range -> eigen_assert throws (SURVEY B-9) */
                f->status |= ORACLE_ERR_SAME_STEP_REMATCH;
                f->M = f->M_commit;       /* freeze at the last committed state */
                return f->status;
            }
            i = i * 2 + 3;                                               /* :113 */
            const double dx = f->x_t[i] - xp[0], dy = f->x_t[i + 1] - xp[1];
            const float dist = (float)sqrt(dx * dx + dy * dy);           /* :115 */
            const double dd = (double)dist;
            const double d2 = (double)(float)(dist * dist);              /* float product, :120 */
            double H[2][5]; const int hc[5] = {0, 1, 2, i, i + 1};
            H[0][0] = -(dx) / dd;  H[0][1] = -(dy) / dd;  H[0][2] = 0.0;  H[0][3] = (dx) / dd;  H[0][4] = (dy) / dd;   /* :118-119,123-124 */
            H[1][0] = (dy) / d2;   H[1][1] = -(dx) / d2;  H[1][2] = -1.0; H[1][3] = -(dy) / d2; H[1][4] = (dx) / d2;   /* :120-122,125-126 */
            const float ang = (float)remainder(atan2(dy, dx) - xp[2], 2 * PI_REF);  /* :129 */
            const double nu0 = (double)(float)(r - dist - f->p.w_r);    /* all-float, :130 */
            const double nu1 = (double)(float)(b - ang - f->p.w_b);     /* :131 */
            double S[2][2], Si[2][2];
            double* K = f->vec1;      /* n x 2 */
            if (mode == ORACLE_DENSE) {
                double *Hx = f->w1, *HP = f->w2, *PHt = f->w3, *A = f->w4;
                for (int j = 0; j < n; ++j) { Hx[j] = 0.0; Hx[ld + j] = 0.0; }              /* :117 */
                for (int q = 0; q < 5; ++q) { Hx[hc[q]] = H[0][q]; Hx[ld + hc[q]] = H[1][q]; }
                mm(HP, ld, Hx, ld, Pp, ld, 2, n, n);               /* H_x * P_pred */
                double HPHt[4]; mmT(HPHt, 2, HP, ld, Hx, ld, 2, n, 2);
                /* H_w * W * H_w^T with H_w = I (ekf.cpp:20) reproduces W exactly */
                S[0][0] = HPHt[0] + f->W[0][0]; S[0][1] = HPHt[1] + f->W[0][1];
                S[1][0] = HPHt[2] + f->W[1][0]; S[1][1] = HPHt[3] + f->W[1][1];            /* :133 */
                inv2_lu(S, Si);
                mmT(PHt, 2, Pp, ld, Hx, ld, n, n, 2);              /* P_pred * H_x^T  (n x 2, ld 2) */
                double Sim[4] = {Si[0][0], Si[0][1], Si[1][0], Si[1][1]};
                mm(K, 2, PHt, 2, Sim, 2, n, 2, 2);                 /* :135 */
                for (int q = 0; q < n; ++q) xp[q] = xp[q] + (K[q * 2] * nu0 + K[q * 2 + 1] * nu1);   /* :138 */
                xp[2] = remainder(xp[2], 2 * PI_REF);              /* :139 */
                mm(A, ld, K, 2, Hx, ld, n, 2, n);                  /* K * H_x   (n x n) */
                mm(HP, ld, A, ld, Pp, ld, n, n, n);                /* (..) * P_pred */
                for (int q = 0; q < n; ++q) for (int j = 0; j < n; ++j) Pp[q * ld + j] = Pp[q * ld + j] - HP[q * ld + j]; /* :140 */
            } else {
                double* HP = f->w2;     /* 2 x n */
                for (int j = 0; j < n; ++j)
                    for (int rr = 0; rr < 2; ++rr) {
                        double s = 0.0;
                        for (int q = 0; q < 5; ++q) s += H[rr][q] * Pp[hc[q] * ld + j];
                        HP[rr * ld + j] = s;
                    }
                for (int rr = 0; rr < 2; ++rr) for (int cc = 0; cc < 2; ++cc) {
                    double s = 0.0;
                    for (int q = 0; q < 5; ++q) s += HP[rr * ld + hc[q]] * H[cc][q];
                    S[rr][cc] = s + f->W[rr][cc];
                }
                inv2_lu(S, Si);
                double* A5 = f->vec2;   /* n x 5 : non-zero columns of K*H_x */
                for (int q = 0; q < n; ++q) {
                    double ph0 = 0.0, ph1 = 0.0;
                    for (int t = 0; t < 5; ++t) { ph0 += Pp[q * ld + hc[t]] * H[0][t]; ph1 += Pp[q * ld + hc[t]] * H[1][t]; }
                    K[q * 2] = ph0 * Si[0][0] + ph1 * Si[1][0];
                    K[q * 2 + 1] = ph0 * Si[0][1] + ph1 * Si[1][1];
                }
                for (int q = 0; q < n; ++q) {
                    xp[q] = xp[q] + (K[q * 2] * nu0 + K[q * 2 + 1] * nu1);
                    for (int t = 0; t < 5; ++t) A5[q * 5 + t] = K[q * 2] * H[0][t] + K[q * 2 + 1] * H[1][t];
                }
                xp[2] = remainder(xp[2], 2 * PI_REF);
                /* rows hc[] of P_pred are read by every output row: snapshot them first */
                double* R5 = f->w3;
                for (int t = 0; t < 5; ++t) for (int j = 0; j < n; ++j) R5[t * ld + j] = Pp[hc[t] * ld + j];
                for (int q = 0; q < n; ++q) for (int j = 0; j < n; ++j) {
                    double s = 0.0;
                    for (int t = 0; t < 5; ++t) s += A5[q * 5 + t] * R5[t * ld + j];
                    Pp[q * ld + j] = Pp[q * ld + j] - s;
                }
            }
        } else {
            /* ---------------- landmark insertion, :141-173 */
            if (f->M >= f->cap_lm) { f->status |= ORACLE_ERR_CAPACITY; continue; }
            f->M += 1;                                                           /* :144 */
            const int nn = 3 + 2 * f->M, r0 = nn - 2, r1 = nn - 1;
            const double cb = cos(xp[2] + (double)b), sb = sin(xp[2] + (double)b);
            xp[r0] = xp[0] + (double)r * cb;                                     /* :147 */
            xp[r1] = xp[1] + (double)r * sb;                                     /* :148 */
            f->lm_ids[f->M - 1] = id;                                            /* :150 */
            const double Gx[2][3] = {{1.0, 0.0, -(double)r * sb}, {0.0, 1.0, (double)r * cb}};      /* :160-165 */
            const double Gz[2][2] = {{cb, -(double)r * sb}, {sb, (double)r * cb}};                  /* :155-158 */
            if (mode == ORACLE_DENSE) {
                double *Y = f->w1, *pt = f->w2, *T = f->w3;
                for (int a = 0; a < nn; ++a) for (int c2 = 0; c2 < nn; ++c2) { Y[a * ld + c2] = (a == c2) ? 1.0 : 0.0; pt[a * ld + c2] = 0.0; }
                Y[r0 * ld + r0] = Gz[0][0]; Y[r0 * ld + r1] = Gz[0][1]; Y[r1 * ld + r0] = Gz[1][0]; Y[r1 * ld + r1] = Gz[1][1];
                for (int c2 = 0; c2 < 3; ++c2) { Y[r0 * ld + c2] = Gx[0][c2]; Y[r1 * ld + c2] = Gx[1][c2]; }
                for (int a = 0; a < n; ++a) for (int c2 = 0; c2 < n; ++c2) pt[a * ld + c2] = Pp[a * ld + c2];   /* :169 */
                pt[r0 * ld + r0] = f->W[0][0]; pt[r0 * ld + r1] = f->W[0][1]; pt[r1 * ld + r0] = f->W[1][0]; pt[r1 * ld + r1] = f->W[1][1]; /* :170 */
                mm(T, ld, Y, ld, pt, ld, nn, nn, nn);
                mmT(Pp, ld, T, ld, Y, ld, nn, nn, nn);                           /* :172 */
            } else {
                double T[2][3];  /* rows r0,r1 of Y*p_temp restricted to columns 0..2 (needed for the new 2x2 block) */
                double* Tr = f->w2; /* 2 x nn */
                for (int rr = 0; rr < 2; ++rr) {
                    for (int j = 0; j < n; ++j) {
                        double s = 0.0;
                        for (int k = 0; k < 3; ++k) s += Gx[rr][k] * Pp[k * ld + j];
                        Tr[rr * ld + j] = s;
                    }
                    for (int c2 = 0; c2 < 2; ++c2) Tr[rr * ld + n + c2] = Gz[rr][0] * f->W[0][c2] + Gz[rr][1] * f->W[1][c2];
                    for (int k = 0; k < 3; ++k) T[rr][k] = Tr[rr * ld + k];
                }
                /* new columns for old rows: sum_k P[a][k] Gx[c][k] */
                for (int a = 0; a < n; ++a) for (int c2 = 0; c2 < 2; ++c2) {
                    double s = 0.0;
                    for (int k = 0; k < 3; ++k) s += Pp[a * ld + k] * Gx[c2][k];
                    Pp[a * ld + n + c2] = s;
                }
                for (int rr = 0; rr < 2; ++rr) {
                    for (int j = 0; j < n; ++j) Pp[(n + rr) * ld + j] = Tr[rr * ld + j];
                    for (int c2 = 0; c2 < 2; ++c2) {
                        double s = 0.0;
                        for (int k = 0; k < 3; ++k) s += T[rr][k] * Gx[c2][k];
                        s += Tr[rr * ld + n] * Gz[c2][0];
                        s += Tr[rr * ld + n + 1] * Gz[c2][1];
                        Pp[(n + rr) * ld + n + c2] = s;
                    }
                }
            }
        }
    }
    ekf_commit(f);                                                              /* :176-177 */
    return f->status;
}

/* ------------------------------------------------------------------ UKF, ukf.cpp:106-371 */
static float ukf_yaw(const double* x) { return (float)remainder(atan2(x[3], x[2]), 2 * PI_REF); }

/* ukf.cpp:125-135 */
static void ukf_motion_model(const oracle_filter* f, const double* x, int n, float u_d, float u_th, double* out) {
    for (int i = 0; i < n; ++i) out[i] = x[i];
    const float yaw = ukf_yaw(x);
    const float ud = u_d + f->p.v_d;
    out[0] = x[0] + (double)(float)(ud * cos_f(yaw));
    out[1] = x[1] + (double)(float)(ud * sin_f(yaw));
    const float fsum = yaw + u_th + f->p.v_th;
    const float new_yaw = (float)remainder((double)fsum, 2 * PI_REF);
    out[2] = (double)cos_f(new_yaw);
    out[3] = (double)sin_f(new_yaw);
}

/* ukf.cpp:137-159 (both branches: SLAM reads the landmark from the sigma point, localisation from the true map) */
static void ukf_sensing_model(const oracle_filter* f, const double* x, int lm_i, double z[2]) {
    const float yaw = ukf_yaw(f->x_t);                 /* prior x_t, same for every sigma point (:139) */
    double dx, dy;
    if (f->kind == ORACLE_UKF_SLAM) { dx = x[lm_i] - x[0]; dy = x[lm_i + 1] - x[1]; }                         /* :144-145 */
    else { dx = (double)f->map[lm_i * 3 + 1] - x[0]; dy = (double)f->map[lm_i * 3 + 2] - x[1]; }              /* :152-153 */
    z[0] = sqrt(dx * dx + dy * dy) + (double)f->p.w_r;
    z[1] = atan2(dy, dx) - (double)yaw + (double)f->p.w_b;
    z[1] = remainder(z[1], 2 * PI_REF);
}

static void ukf_prediction(oracle_filter* f, float u_d, float u_th, int mode) {
    const int n = 4 + 2 * f->M, ld = f->ncap, ns = 2 * n + 1;
    /* nearestSPD, :106-123 */
    double *Y = f->w1, *Qv = f->w2, *Yp = f->w3, *sq = f->w4;
    double* D = f->vec1; double* tmp = f->vec2;
    const float scale = (2 * f->M + 4) / (1 - 0.2f);                   /* :114 (float) */
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j)
        Y[i * n + j] = (0.5 * (f->P_t[i * ld + j] + f->P_t[j * ld + i])) * (double)scale;   /* :112,114 */
    oracle_eigh(Y, n, D, Qv);                                            /* :116-118 */
    for (int i = 0; i < n; ++i) if (D[i] < 0.00000001) D[i] = 0.00000001; /* :120 */
    if (mode == ORACLE_DENSE) {
        /* Qv * Dplus.asDiagonal() * Qv^T, :122 */
        double* QD = sq;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) QD[i * n + j] = Qv[i * n + j] * D[j];
        mmT(Yp, n, QD, n, Qv, n, n, n, n);
        /* .sqrt(), :208 -- restated as a symmetric eigen sqrt of the reconstructed matrix (D-3) */
        double* U = Y; double* lam = tmp;
        for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) { /* eigh reads a symmetric matrix */
            double a = 0.5 * (Yp[i * n + j] + Yp[j * n + i]); Yp[i * n + j] = a; Yp[j * n + i] = a; }
        oracle_eigh(Yp, n, lam, U);
        for (int i = 0; i < n; ++i) lam[i] = lam[i] > 0 ? sqrt(lam[i]) : 0.0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) QD[i * n + j] = U[i * n + j] * lam[j];
        mmT(Yp, n, QD, n, U, n, n, n, n);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) sq[i * n + j] = Yp[i * n + j];
    } else {
        /* single decomposition: Qv * sqrt(Dplus) * Qv^T */
        for (int i = 0; i < n; ++i) tmp[i] = sqrt(D[i]);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Yp[i * n + j] = Qv[i * n + j] * tmp[j];
        mmT(sq, n, Yp, n, Qv, n, n, n, n);
    }
    /* sigma points, :214-220 */
    double *X = f->X, *Xp = f->X_pred;
    for (int i = 0; i < n; ++i) X[i] = f->x_t[i];
    for (int c = 1; c <= n; ++c) for (int i = 0; i < n; ++i) X[(size_t)c * ld + i] = f->x_t[i] + sq[i * n + (c - 1)];
    for (int c = 1; c <= n; ++c) for (int i = 0; i < n; ++i) X[(size_t)(c + n) * ld + i] = f->x_t[i] - sq[i * n + (c - 1)];
    /* propagate, :223-226 */
    for (int c = 0; c < ns; ++c) ukf_motion_model(f, X + (size_t)c * ld, n, u_d, u_th, Xp + (size_t)c * ld);
    /* mean, :228-232 */
    for (int i = 0; i < n; ++i) f->x_pred[i] = 0.0;
    for (int c = 0; c < ns; ++c) for (int i = 0; i < n; ++i) f->x_pred[i] += f->Wts[c] * Xp[(size_t)c * ld + i];
    /* covariance, :235-240 */
    double* Pp = f->P_pred;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Pp[i * ld + j] = 0.0;
    for (int c = 0; c < ns; ++c) {
        for (int i = 0; i < n; ++i) tmp[i] = Xp[(size_t)c * ld + i] - f->x_pred[i];
        const double w = f->Wts[c];
        for (int i = 0; i < n; ++i) { const double wi = w * tmp[i]; for (int j = 0; j < n; ++j) Pp[i * ld + j] += wi * tmp[j]; }
    }
    for (int i = 0; i < 4; ++i) Pp[i * ld + i] += f->Qd[i];           /* :240 (Q is zero elsewhere) */
}

/* ukf.cpp:293-349 */
static void ukf_landmark_update(oracle_filter* f, int lm_slot, float r, float b) {
    const int n = 4 + 2 * f->M, ld = f->ncap, ns = 2 * n + 1;
    const int lm_i = (f->kind == ORACLE_UKF_SLAM) ? lm_slot * 2 + 4 : lm_slot;   /* :296-302 (localisation: lm_i = id) */
    double* Z = f->w1; /* 2 x ns, column c at Z[2c] */
    for (int c = 0; c < ns; ++c) ukf_sensing_model(f, f->X_pred + (size_t)c * ld, lm_i, Z + 2 * c);   /* :305-308 */
    double z_est[2] = {0.0, 0.0};
    for (int c = 0; c < ns; ++c) z_est[0] += f->Wts[c] * Z[2 * c];    /* :312-314, bearing mean never accumulated */
    double S[2][2] = {{0, 0}, {0, 0}};
    for (int c = 0; c < ns; ++c) {                                     /* :318-324 */
        double d0 = Z[2 * c] - z_est[0];
        double d1 = remainder(Z[2 * c + 1] - z_est[1], 2 * PI_REF);
        const double w = f->Wts[c];
        S[0][0] += (w * d0) * d0; S[0][1] += (w * d0) * d1; S[1][0] += (w * d1) * d0; S[1][1] += (w * d1) * d1;
    }
    for (int a = 0; a < 2; ++a) for (int c = 0; c < 2; ++c) S[a][c] += f->W[a][c];     /* :326 */
    double* C = f->vec1; /* n x 2 */
    for (int i = 0; i < 2 * n; ++i) C[i] = 0.0;
    for (int c = 0; c < ns; ++c) {                                     /* :329-337 */
        double d20 = Z[2 * c] - z_est[0];
        double d21 = remainder(Z[2 * c + 1] - z_est[1], 2 * PI_REF);
        const double w = f->Wts[c];
        const double* xc = f->X_pred + (size_t)c * ld;
        for (int i = 0; i < n; ++i) {
            const double wd = w * (xc[i] - f->x_pred[i]);
            C[i * 2] += wd * d20; C[i * 2 + 1] += wd * d21;
        }
    }
    double Si[2][2]; inv2_lu(S, Si);
    double* K = f->vec2; /* n x 2 */
    for (int i = 0; i < n; ++i) {                                      /* :339 */
        K[i * 2] = C[i * 2] * Si[0][0] + C[i * 2 + 1] * Si[1][0];
        K[i * 2 + 1] = C[i * 2] * Si[0][1] + C[i * 2 + 1] * Si[1][1];
    }
    const double in0 = (double)r - z_est[0];                           /* :342-344 */
    const double in1 = remainder((double)b - z_est[1], 2 * PI_REF);
    for (int i = 0; i < n; ++i) f->x_pred[i] = f->x_pred[i] + (K[i * 2] * in0 + K[i * 2 + 1] * in1);   /* :345 */
    /* P_pred - (K*S)*K^T, :348 */
    double* Pp = f->P_pred;
    for (int i = 0; i < n; ++i) {
        const double ks0 = K[i * 2] * S[0][0] + K[i * 2 + 1] * S[1][0];
        const double ks1 = K[i * 2] * S[0][1] + K[i * 2 + 1] * S[1][1];
        for (int j = 0; j < n; ++j) Pp[i * ld + j] = Pp[i * ld + j] - (ks0 * K[j * 2] + ks1 * K[j * 2 + 1]);
    }
}

/* ukf.cpp:351-371 */
static void ukf_landmark_insertion(oracle_filter* f, int id, float r, float b) {
    if (f->M >= f->cap_lm) { f->status |= ORACLE_ERR_CAPACITY; return; }
    const int n = 4 + 2 * f->M, ld = f->ncap;
    const float yaw = ukf_yaw(f->x_pred);                              /* :356 */
    const float yb = yaw + b;
    f->x_pred[n] = f->x_pred[0] + (double)(float)(r * cos_f(yb));       /* :358 */
    f->x_pred[n + 1] = f->x_pred[1] + (double)(float)(r * sin_f(yb));   /* :359 */
    f->lm_ids[f->M] = id;                                              /* :361 */
    double* Pp = f->P_pred;
    for (int i = 0; i < n + 2; ++i) for (int j = (i < n ? n : 0); j < n + 2; ++j) Pp[i * ld + j] = 0.0;   /* :365-366 */
    Pp[n * ld + n] = f->W[0][0]; Pp[n * ld + n + 1] = f->W[0][1];
    Pp[(n + 1) * ld + n] = f->W[1][0]; Pp[(n + 1) * ld + n + 1] = f->W[1][1];                             /* :367 */
    f->M += 1;                                                         /* :371 */
}

static int ukf_update(oracle_filter* f, float u_d, float u_th, const float* lm_meas, int num_landmarks, int mode) {
    f->timestep += 1;                                                  /* :164 */
    const int n = f->M * 2 + 4, ld = f->ncap;
    if (n != f->X_rows) {                                              /* :169-180 */
        f->X_rows = n;
        const float w = (1 - 0.2f) / (2 * n);                          /* float, :175 */
        for (int i = 0; i < 2 * n + 1; ++i) f->Wts[i] = 1.0 * (double)w;
        f->Wts[0] = (double)0.2f;
        for (int i = 0; i < n; ++i) f->Qd[i] = 0.0;
    }
    const float yaw = ukf_yaw(f->x_t);                                 /* :182 */
    f->Qd[0] = f->V[0][0] * (double)cos_f(yaw);                        /* :183-186 */
    f->Qd[1] = f->V[0][0] * (double)sin_f(yaw);
    f->Qd[2] = f->V[1][1] * (double)cos_f(yaw);
    f->Qd[3] = f->V[1][1] * (double)sin_f(yaw);
    ukf_prediction(f, u_d, u_th, mode);                                /* :189 */
    /* updateStage, :243-291 */
    f->n_assoc = 0;
    int* new_idx = (int*)malloc(sizeof(int) * (size_t)(num_landmarks + 1));
    int n_new = 0;
    for (int l = 0; l < num_landmarks; ++l) {
        const int id = (int)lm_meas[l * 3];
        const float r = lm_meas[l * 3 + 1], b = lm_meas[l * 3 + 2];
        int lm_i = -1;
        if (f->kind == ORACLE_UKF_LOC) {                               /* :262,272: every detection is an update, by map id */
            if (id < 0 || id >= f->n_map) { f->status |= ORACLE_ERR_BAD_ID; log_assoc(f, -1); continue; }   /* the reference reads map[] out of range */
            log_assoc(f, id);
            ukf_landmark_update(f, id, r, b);
            continue;
        }
        for (int j = 0; j < f->M; ++j) if (f->lm_ids[j] == id) { lm_i = j; break; }
        log_assoc(f, lm_i);
        if (lm_i == -1) new_idx[n_new++] = l; else ukf_landmark_update(f, lm_i, r, b);
    }
    for (int q = 0; q < n_new; ++q) {
        const int l = new_idx[q];
        ukf_landmark_insertion(f, (int)lm_meas[l * 3], lm_meas[l * 3 + 1], lm_meas[l * 3 + 2]);
    }
    free(new_idx);
    const int n2 = f->M * 2 + 4;
    for (int i = 0; i < n2; ++i) { f->x_t[i] = f->x_pred[i]; for (int j = 0; j < n2; ++j) f->P_t[i * ld + j] = f->P_pred[i * ld + j]; }  /* :289-290 */
    f->M_commit = f->M;
    return f->status;
}

/* ------------------------------------------------------------------ public update entry points */
int oracle_update(oracle_filter* f, float fwd, float ang, const float* meas, int n_meas, int mode) {
    if (f->status & ORACLE_ERR_SAME_STEP_REMATCH) return f->status;
    if (f->kind == ORACLE_EKF_SLAM) { ekf_predict(f, fwd, ang, mode); ekf_measure(f, meas, n_meas, mode); }
    else if (f->kind == ORACLE_NAIVE) {
        /* NaiveFilter::update, filter.h:342-348: measurements ignored, pose propagated by the command */
        f->timestep += 1;
        const double th = f->x_t[2];
        f->x_t[0] = f->x_t[0] + (double)fwd * cos(th);
        f->x_t[1] = f->x_t[1] + (double)fwd * sin(th);
        f->x_t[2] = remainder(th + (double)ang, 2 * PI_REF);
    }
    else ukf_update(f, fwd, ang, meas, n_meas, mode);
    if (has_nan(f)) f->status |= ORACLE_ERR_NAN;
    return f->status;
}
int oracle_predict(oracle_filter* f, float fwd, float ang, int mode) {
    if (f->kind != ORACLE_EKF_SLAM) return -1;
    if (f->status & ORACLE_ERR_SAME_STEP_REMATCH) return f->status;
    ekf_predict(f, fwd, ang, mode);
    ekf_commit(f);      /* a predict with no measurement batch is the reference's early return, ekf.cpp:67-71 */
    f->in_step = 1;
    return f->status;
}
int oracle_measure(oracle_filter* f, const float* meas, int n_meas, int mode) {
    if (f->kind != ORACLE_EKF_SLAM) return -1;
    if (f->status & ORACLE_ERR_SAME_STEP_REMATCH) return f->status;
    const int n = 3 + 2 * f->M, ld = f->ncap;
    /* x_t already equals x_pred (committed by oracle_predict); landmark means are unchanged by predict,
       so the stale-x_t reads of ekf.cpp:115-129 see the same values as in the fused form. */
    for (int i = 0; i < n; ++i) { f->x_pred[i] = f->x_t[i]; for (int j = 0; j < n; ++j) f->P_pred[i * ld + j] = f->P_t[i * ld + j]; }
    ekf_measure(f, meas, n_meas, mode);
    f->in_step = 0;
    if (has_nan(f)) f->status |= ORACLE_ERR_NAN;
    return f->status;
}

/* ------------------------------------------------------------------ Philox4x32-10 + simulator */
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 53-bit uniform in [0,1), the construction of CPython's random.random() (sim_node.py:16) */
double oracle_uniform(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

int oracle_sim_step(const oracle_params* p, double truth[3], float fwd, float ang,
                    const double* lm_xy, int n_lm, uint64_t seed, uint32_t instance, uint32_t step,
                    float* meas_out, int cap) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t rn[4];
    oracle_philox(instance, step, 0u, 0u, k0, k1, rn);
    /* sim_node.py:216-217 (msg.fwd/ang are the float32 wire values) */
    double d = (double)fwd + 2 * p->V_00 * oracle_uniform(rn[0], rn[1]) - p->V_00;
    double hdg = (double)ang + 2 * p->V_11 * oracle_uniform(rn[2], rn[3]) - p->V_11;
    d = fmax(0.0, fmin(d, p->d_max));                                             /* :219 */
    hdg = fmax(-p->th_max, fmin(hdg, p->th_max));                                 /* :220 */
    const double nx = truth[0] + d * cos(truth[2]), ny = truth[1] + d * sin(truth[2]), nyaw = truth[2] + hdg; /* :222 */
    truth[0] = nx; truth[1] = ny; truth[2] = nyaw;
    int k = 0;
    for (int id = 0; id < n_lm; ++id) {                                           /* :231-243 */
        const double dx = lm_xy[2 * id] - truth[0], dy = lm_xy[2 * id + 1] - truth[1];
        const double r = sqrt(dx * dx + dy * dy);
        const double gb = atan2(dy, dx);
        const double beta = remainder(gb - truth[2], 2 * PI_REF);
        if (r > p->range_max) continue;
        if (beta > p->fov_min && beta < p->fov_max) {
            if (k >= cap) continue;
            oracle_philox(instance, step, 1u + (uint32_t)id, 0u, k0, k1, rn);
            meas_out[3 * k] = (float)id;                                          /* :246-249, float32 wire */
            meas_out[3 * k + 1] = (float)(r + 2 * p->W_00 * oracle_uniform(rn[0], rn[1]) - p->W_00);
            meas_out[3 * k + 2] = (float)(beta + 2 * p->W_11 * oracle_uniform(rn[2], rn[3]) - p->W_11);
            ++k;
        }
    }
    return k;
}

/* ------------------------------------------------------------------ generate_trajectory, sim_node.py:63-152
 * noisy copy of the map (:83-87), nearest-neighbour tour from the start pose (:89-112), one clamped command per step,
 * rotating the tour when within the visitation threshold (:118-152).  Map noise: Philox keyed (seed; instance, id, 0, 1)
 * (deviation D-5; the reference uses the unseeded random.random()). */
static double tsp_norm(double ax, double ay, double bx, double by) { const double dx = ax - bx, dy = ay - by; return sqrt(dx * dx + dy * dy); }
/* generate_landmarks, sim_node.py:155-206: map_type 0 "grid" (:165-176), 1 "random" (:177-188, blank occupancy map).  Returns the
 * number of landmarks written to lm_xy ([cap][2]), or -1 (invalid map_type, :196-198) / -2 (capacity or attempts exhausted).
 * Random draws: Philox (seed; instance, attempt, 0, 2), two uniforms per attempt in the order of :180 (deviation D-5). */
int oracle_make_map(int map_type, int n_landmarks, double bound, double grid_step, double min_sep, uint64_t seed, uint32_t instance,
                    double* lm_xy, int cap) {
    if (map_type == 0) {
        const double shift = grid_step / 2;                                            /* :166 */
        int id = 0;
        const double start = -bound + shift;
        const int cnt = (int)ceil((bound - start) / grid_step);                       /* len(np.arange(start, bound, step)) */
        /* np.arange fills start + i * delta with delta = (start + step) - start as rounded (numpy's arange fill) */
        const volatile double next = start + grid_step;
        const double delta = next - start;
        for (int r = 0; r < cnt; ++r)
            for (int c = 0; c < cnt; ++c) {                                            /* :168-171 */
                if (id >= cap) return -2;
                lm_xy[2 * id] = start + r * delta; lm_xy[2 * id + 1] = start + c * delta;
                ++id;
            }
        return id;
    }
    if (map_type != 1) return -1;
    if (n_landmarks > cap) return -2;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    int have = 0;
    for (uint32_t a = 0; have < n_landmarks && a < 64u * (uint32_t)n_landmarks; ++a) { /* :179 */
        uint32_t rn[4];
        oracle_philox(instance, a, 0u, 2u, k0, k1, rn);
        const double px = 2 * bound * oracle_uniform(rn[0], rn[1]) - bound, py = 2 * bound * oracle_uniform(rn[2], rn[3]) - bound;   /* :180 */
        int close = 0;
        for (int q = 0; q < have && !close; ++q) close = tsp_norm(lm_xy[2 * q], lm_xy[2 * q + 1], px, py) < min_sep;                  /* :184 */
        if (close) continue;
        lm_xy[2 * have] = px; lm_xy[2 * have + 1] = py;                                /* :186-187 */
        ++have;
    }
    return have < n_landmarks ? -2 : have;
}

int oracle_tsp_trajectory(const oracle_params* p, const double* lm_xy, int n_lm, double landmark_noise,
                          double visitation_threshold, double bound, double x0, double y0, double yaw0, int T,
                          uint64_t seed, uint32_t instance, float* fwd_out, float* ang_out) {
    if (n_lm < 1 || T < 0) return -1;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    double* nx = dalloc((size_t)n_lm); double* ny = dalloc((size_t)n_lm);
    int* path = (int*)calloc((size_t)n_lm, sizeof(int));
    char* seen = (char*)calloc((size_t)n_lm, 1);
    const double lo = -bound + 1, hi = bound - 1;                                  /* :86-87 */
    for (int i = 0; i < n_lm; ++i) {
        uint32_t rn[4];
        oracle_philox(instance, (uint32_t)i, 0u, 1u, k0, k1, rn);
        const double ax = lm_xy[2 * i] + 2 * landmark_noise * oracle_uniform(rn[0], rn[1]) - landmark_noise;      /* :84 */
        const double ay = lm_xy[2 * i + 1] + 2 * landmark_noise * oracle_uniform(rn[2], rn[3]) - landmark_noise;  /* :85 */
        nx[i] = fmax(lo, fmin(ax, hi)); ny[i] = fmax(lo, fmin(ay, hi));
    }
    double x = x0, y = y0, th = yaw0;
    int cur = 0;
    double best = tsp_norm(nx[0], ny[0], x, y);                                    /* :90-96 */
    for (int i = 0; i < n_lm; ++i) { const double d = tsp_norm(nx[i], ny[i], x, y); if (d < best) { cur = i; best = d; } }
    path[0] = cur; seen[cur] = 1;
    for (int k = 1; k < n_lm; ++k) {                                               /* :100-112 */
        int g = -1; double bd = -1.0;
        for (int i = 0; i < n_lm; ++i) {
            if (seen[i]) continue;
            const double d = tsp_norm(nx[i], ny[i], nx[cur], ny[cur]);
            if (bd < 0 || d < bd) { g = i; bd = d; }
        }
        path[k] = g; seen[g] = 1; cur = g;
    }
    int head = 0;
    for (int t = 0; t < T; ++t) {                                                  /* :118-152 */
        if (tsp_norm(x, y, nx[path[head]], ny[path[head]]) < visitation_threshold) head = (head + 1 == n_lm) ? 0 : head + 1;
        const double gx = nx[path[head]], gy = ny[path[head]];
        double d = tsp_norm(gx, gy, x, y);
        const double gb = atan2(gy - y, gx - x);
        double hdg = remainder(gb - th, 2 * PI_REF);
        d = fmin(d, p->d_max);
        if (fabs(hdg) > p->th_max) hdg = (hdg > 0) ? p->th_max : -p->th_max;
        x = x + d * cos(th); y = y + d * sin(th); th = th + hdg;
        fwd_out[t] = (float)d; ang_out[t] = (float)hdg;
    }
    free(nx); free(ny); free(path); free(seen);
    return 0;
}

/* ------------------------------------------------------------------ Monte-Carlo instance + CPU timing */
int oracle_run_instance(int kind, const oracle_params* p, const double* lm_xy, int n_lm,
                        const float* cmd_fwd, const float* cmd_ang, int T,
                        uint64_t seed, uint32_t instance, int max_landmarks, int mode,
                        double* pose_trace, double* truth_trace, oracle_filter** keep) {
    oracle_filter* f = oracle_create(kind, p, max_landmarks);
    if (!f) return -1;
    oracle_init(f, 0.0f, 0.0f, 0.0f);                       /* params.yaml:19-22 */
    if (kind == ORACLE_UKF_LOC) {                           /* /truth/landmarks: [id, x, y]* float32, ids ascending (sim_node.py:190-196) */
        float* mp = (float*)malloc(sizeof(float) * 3 * (size_t)(n_lm > 0 ? n_lm : 1));
        for (int j = 0; j < n_lm; ++j) { mp[3 * j] = (float)j; mp[3 * j + 1] = (float)lm_xy[2 * j]; mp[3 * j + 2] = (float)lm_xy[2 * j + 1]; }
        oracle_set_map(f, mp, n_lm);
        free(mp);
    }
    double truth[3] = {0.0, 0.0, 0.0};                      /* sim_node.py:32 */
    const int cap = n_lm > 0 ? n_lm : 1;
    float* meas = (float*)malloc(sizeof(float) * 3 * (size_t)cap);
    for (int t = 0; t < T; ++t) {
        const int k = oracle_sim_step(p, truth, cmd_fwd[t], cmd_ang[t], lm_xy, n_lm, seed, instance, (uint32_t)t, meas, cap);
        oracle_update(f, cmd_fwd[t], cmd_ang[t], meas, k, mode);
        if (pose_trace) {
            pose_trace[3 * t] = f->x_t[0]; pose_trace[3 * t + 1] = f->x_t[1];
            pose_trace[3 * t + 2] = (f->base == 3) ? f->x_t[2] : remainder(atan2(f->x_t[3], f->x_t[2]), 2 * PI_REF);
        }
        if (truth_trace) { truth_trace[3 * t] = truth[0]; truth_trace[3 * t + 1] = truth[1]; truth_trace[3 * t + 2] = truth[2]; }
    }
    free(meas);
    const int st = f->status;
    if (keep) *keep = f; else oracle_destroy(f);
    return st;
}

typedef struct {
    int kind; const oracle_params* p; const double* lm; int n_lm; const float* fwd; const float* ang; int T;
    uint64_t seed; uint32_t first; int count; int max_lm; int mode;
} bench_arg;

static void* bench_thread(void* a_) {
    bench_arg* a = (bench_arg*)a_;
    for (int i = 0; i < a->count; ++i)
        oracle_run_instance(a->kind, a->p, a->lm, a->n_lm, a->fwd, a->ang, a->T, a->seed, a->first + (uint32_t)i, a->max_lm, a->mode, NULL, NULL, NULL);
    return NULL;
}

double oracle_bench(int kind, const oracle_params* p, const double* lm_xy, int n_lm,
                    const float* cmd_fwd, const float* cmd_ang, int T, uint64_t seed,
                    int n_threads, int per_thread, int max_landmarks, int mode, long long* updates) {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    bench_arg* args = (bench_arg*)malloc(sizeof(bench_arg) * (size_t)n_threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < n_threads; ++i) {
        bench_arg a = {kind, p, lm_xy, n_lm, cmd_fwd, cmd_ang, T, seed, (uint32_t)(i * per_thread), per_thread, max_landmarks, mode};
        args[i] = a;
        pthread_create(&th[i], NULL, bench_thread, &args[i]);
    }
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(args);
    if (updates) *updates = (long long)n_threads * per_thread * T;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
