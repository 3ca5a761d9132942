// Host harness for csrc/eig3.cuh: runs the per-thread building blocks of the generation-3 tridiagonal eigensolver with the
// kernel's orchestration (ukf_eig3_kernel: splits, bisection, clusters, twisted vectors, Gram-Schmidt rounds, refinement,
// residual check) emulated thread by thread on the CPU.  stdin: n maxc, then d[0..n), e[0..n-1).  stdout: status line
// ("ok" | "declined <why>"), then lambda[0..n), then V row by row (V[i][k] = component i of eigenvector k).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#ifdef TRACEJ
#define EIG3_TRACE_ROOT(j) ((j) == TRACEJ)
#endif
#include "../live_ekf_slam_b200/csrc/eig3.cuh"

using namespace slam::eig3;

int main(int argc, char** argv) {
    const bool variants = argc > 1 && std::string(argv[1]) == "variants";
    int n = 0, maxc = 16;
    if (scanf("%d %d", &n, &maxc) != 2) return 2;
    std::vector<double> d(n), e(n, 0.0), e2(n), lam(n), xs(n);
    for (int i = 0; i < n; ++i) if (scanf("%lf", &d[i]) != 1) return 2;
    for (int i = 0; i + 1 < n; ++i) if (scanf("%lf", &e[i]) != 1) return 2;
    double tn = 0.0;
    for (int t = 0; t < n; ++t) tn = std::max(tn, std::fabs(d[t]) + (t > 0 ? std::fabs(e[t - 1]) : 0.0) + std::fabs(e[t]));
    std::vector<double> e0 = e;
    for (int t = 0; t + 1 < n; ++t) if (std::fabs(e0[t]) <= EPS * (std::fabs(d[t]) + std::fabs(d[t + 1]))) e[t] = 0.0;
    double e2max = 0.0;
    for (int t = 0; t < n; ++t) { e2[t] = e[t] * e[t]; e2max = std::max(e2max, e2[t]); }
    const double pivmin = std::max(2.2250738585072014e-308 * std::max(1.0, e2max), 1e-300), pivf = EPS * tn;
    std::vector<int> b0(n), b1(n), crank(n), cfirst(n), tw(n);
    std::vector<De> de(n);
    for (int t = 0; t < n; ++t) de[t] = De{d[t], t > 0 ? e2[t - 1] : 0.0};
    std::vector<double> px(n, 0.0), glo(n, 0.0), ghi(n, 0.0);
    std::vector<int> pc(n, 0);
    std::vector<Fval> pf(n, Fval{0.0, 0});
    long evals = 0; int evmax = 0;
    for (int t = 0; t < n; ++t) {
        int a = t; while (a > 0 && e[a - 1] != 0.0) --a;
        int b = t + 1; while (b < n && e[b - 1] != 0.0) ++b;
        b0[t] = a; b1[t] = b;
        if (b - a > 1) {                           // multisection start, as in the kernel
            block_bounds(d.data(), e.data(), a, b, pivmin, glo[t], ghi[t]);
            px[t] = grid_point(glo[t], ghi[t], t - a, b - a);
            pc[t] = sturm_eval(de.data(), a, b, px[t], &pf[t]);
        }
    }
    for (int t = 0; t < n; ++t) {
        const int a = b0[t], b = b1[t];
        if (b - a == 1) { lam[t] = d[a]; continue; }
        double lo = glo[t], hi = ghi[t];
        int qlo, qhi;
        bracket_from_grid2(px.data(), pc.data(), a, b, t - a, lo, hi, qlo, qhi);
        const End el{lo, qlo >= 0 ? pc[qlo] : 0, qlo >= 0 ? pf[qlo] : Fval{0.0, 0}, qlo >= 0};
        const End eh{hi, qhi >= 0 ? pc[qhi] : b - a, qhi >= 0 ? pf[qhi] : Fval{0.0, 0}, qhi >= 0};
        int ne = 0;
        lam[t] = root_bracket(de.data(), a, b, t - a, el, eh, 2.0 * EPS * tn + 2.0 * pivmin, &ne);
        evals += ne; evmax = std::max(evmax, ne);
        if (getenv("EIG3_TRACE")) fprintf(stderr, "  t %d ne %d lam %.17g lo %.17g hi %.17g clo %d chi %d klo %d khi %d\n", t, ne, lam[t], lo, hi, el.c, eh.c, (int)el.known, (int)eh.known);
    }
    fprintf(stderr, "evals mean %.1f max %d\n", (double)evals / n, evmax);
    int maxrank = 0;
    for (int t = 0; t < n; ++t) {
        const int a = b0[t];
        const double ortol = 1.0e-3 * tn;
        int cr = 0;
        while (t - cr > a && lam[t - cr] - lam[t - cr - 1] < ortol && cr <= maxc) ++cr;
        if (cr >= maxc && cr > 0) { printf("declined cluster\n"); return 0; }
        crank[t] = cr; cfirst[t] = t - cr; maxrank = std::max(maxrank, cr);
        double x = lam[cfirst[t]];
        for (int q = cfirst[t] + 1; q <= t; ++q) { const double lq = lam[q], pert = 10.0 * EPS * std::fabs(lq); x = (lq - x < pert) ? x + pert : lq; }
        xs[t] = x;
    }
    if (variants) {
        // the four formulations of the twisted-factorisation eigenvector (two work vectors with a second backward sweep; the same
        // with batched read-backs and parked backward pivots; one work vector; one work vector + parked pivots) must agree BITWISE
        std::vector<double> z0(n), w0(n), z1(n), w1(n), z2(n), z3(n), g3(n);
        long bad_bits = 0;
        for (int t = 0; t < n; ++t) {
            const int a = b0[t], b = b1[t];
            if (b - a == 1) continue;
            double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
            std::fill(z0.begin(), z0.end(), 0.0); std::fill(z1.begin(), z1.end(), 0.0); std::fill(z2.begin(), z2.end(), 0.0); std::fill(z3.begin(), z3.end(), 0.0);
            std::fill(w0.begin(), w0.end(), 0.0); std::fill(w1.begin(), w1.end(), 0.0);
            const int k0 = twisted_vector(d.data(), e.data(), e2.data(), a, b, xs[t], pivf, Slot{z0.data(), 1}, Slot{w0.data(), 1}, &n0);
            const int k1 = twisted_vector_pf<8>(d.data(), e.data(), e2.data(), a, b, xs[t], pivf, Slot{z1.data(), 1}, Slot{w1.data(), 1}, &n1);
            const int k2 = twisted_vector1(d.data(), e.data(), e2.data(), a, b, xs[t], pivf, Slot{z2.data(), 1}, &n2);
            const int k3 = twisted_vector1g(d.data(), e.data(), e2.data(), a, b, xs[t], pivf, Slot{z3.data(), 1}, Slot{g3.data(), 1}, &n3);
            if (k0 != k1 || k0 != k2 || k0 != k3) ++bad_bits;
            if (memcmp(&n0, &n1, 8) || memcmp(&n0, &n2, 8) || memcmp(&n0, &n3, 8)) ++bad_bits;
            for (int i = a; i < b; ++i) {
                if (memcmp(&z0[i], &z1[i], 8) || memcmp(&z0[i], &z2[i], 8) || memcmp(&z0[i], &z3[i], 8)) ++bad_bits;
                if (memcmp(&w0[i], &w1[i], 8)) ++bad_bits;
            }
        }
        printf(bad_bits ? "variants differ %ld\n" : "variants ok\n", bad_bits);
        return 0;
    }
    std::vector<double> V((size_t)n * n, 0.0), W((size_t)n * n, 0.0);
    bool bad = false;
    for (int t = 0; t < n; ++t) {
        const Slot z{V.data() + t, n}, w{W.data() + t, n};
        const int a = b0[t], b = b1[t];
        tw[t] = a;
        if (b - a == 1) { z.set(a, 1.0); continue; }
        double n2 = 1.0;
        tw[t] = twisted_vector_pf<8>(d.data(), e.data(), e2.data(), a, b, xs[t], pivf, z, w, &n2);
        if (!(n2 > 0.0) || !std::isfinite(n2)) bad = true;
        const double sc = 1.0 / std::sqrt(n2);
        for (int i = a; i < b; ++i) z.set(i, z.get(i) * sc);
    }
    bool refine = false;        // set by the first Gram-Schmidt pass: some vector lost more than a quarter of its norm^2
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            if (maxrank == 0 || !refine) break;
            for (int t = 0; t < n; ++t) {
                const int a = b0[t], b = b1[t];
                const bool incl = b - a > 1 && (crank[t] > 0 || (t + 1 < b && crank[t + 1] > 0));
                if (!incl) continue;
                const Slot z{V.data() + t, n}, w{W.data() + t, n};
                const double n2 = twisted_solve(e.data(), a, b, tw[t], z, w);
                if (!(n2 > 0.0) || !std::isfinite(n2)) bad = true;
                const double sc = 1.0 / std::sqrt(n2);
                for (int i = a; i < b; ++i) z.set(i, z.get(i) * sc);
            }
        }
        for (int r = 1; r <= maxrank; ++r)
            for (int t = 0; t < n; ++t) if (crank[t] == r) {
                const Slot z{V.data() + t, n};
                const int a = b0[t], b = b1[t];
                for (int q = cfirst[t]; q < t; ++q) {
                    double dot = 0.0;
                    for (int i = a; i < b; ++i) dot += V[(size_t)i * n + q] * z.get(i);
                    for (int i = a; i < b; ++i) z.set(i, z.get(i) - dot * V[(size_t)i * n + q]);
                }
                double n2 = 0.0;
                for (int i = a; i < b; ++i) n2 += z.get(i) * z.get(i);
                if (pass == 0 && !(n2 > REFINE_BELOW)) refine = true;
                if (pass == 1 && !(n2 > 1.0e-6)) bad = true;    // (first pass: twisted vectors of a pathologically close pair may coincide; the refinement separates them)
                if (getenv("EIG3_DEBUG")) fprintf(stderr, "mgs pass %d t=%d n2=%g\n", pass, t, n2);
                const double sc = 1.0 / std::sqrt(n2);
                for (int i = a; i < b; ++i) z.set(i, z.get(i) * sc);
            }
    }
    for (int t = 0; t < n; ++t) {
        const int a = b0[t], b = b1[t];
        if (b - a == 1) continue;
        const double r = residual_inf(d.data(), e.data(), a, b, lam[t], Slot{V.data() + t, n}, 1.0);
        if (!(r <= 1.0e-12 * tn)) bad = true;
    }
    fprintf(stderr, "refine %d maxrank %d\n", (int)refine, maxrank);
    if (bad) { printf("declined residual\n"); return 0; }
    printf("ok\n");
    for (int t = 0; t < n; ++t) printf("%.17g\n", lam[t]);
    for (int i = 0; i < n; ++i) { for (int k = 0; k < n; ++k) printf("%.17g ", V[(size_t)i * n + k]); printf("\n"); }
    return 0;
}
