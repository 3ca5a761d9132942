// eig3.cuh -- per-thread building blocks of the parallel symmetric-tridiagonal eigensolver used by generation 3 of the UKF
// step (csrc/ukf_batch.cu: ukf_eig3_kernel): thread t of a CTA owns eigenpair t of the instance's tridiagonal matrix
//     T = tridiag(e, d, e)      (d[0..n), e[i] couples i and i+1, e[n-1] = 0; T = Q^T Y Q from the Householder reduction)
// and computes it with O(n) storage and no communication except inside clusters of close eigenvalues:
//   1. T is split where |e_i| <= eps (|d_i| + |d_i+1|) (as every QL / QR implementation deflates); thread t belongs to the
//      block [b0, b1) that contains index t and takes the (t - b0)-th smallest eigenvalue of that block;
//   2. eigenvalue: bisection on the Sturm count of the block (division-free three-term recurrence with rescaling);
//   3. eigenvector: twisted factorisation of T - lambda I (Fernando / Parlett-Dhillon: forward pivots D+, backward pivots D-,
//      twist at argmin |gamma_k|), z_k = 1 and the two recurrences away from k; one inverse-iteration refinement through the
//      same factors for members of a cluster, with modified Gram-Schmidt inside the cluster (LAPACK dstein's criterion:
//      neighbours closer than 1e-3 |T|) -- done by the kernel, which owns the barriers.
// Everything here is plain scalar code (__host__ __device__), so tests/test_eig3_host.py compiles this header with g++ and
// checks it against LAPACK on the CPU; the kernel adds only the thread mapping, the barriers and the memory layout.
//
// This replaces, for the UKF's nearestSPD + sqrt (ukf.cpp:106-123,208), the serial QL iteration (one thread per instance,
// ~0.85 n^2 dependent plane rotations) and the replay of its rotation log on every vector: with explicit eigenvectors V of T,
// S v = Q V sqrt(D+) V^T Q^T v is two dense n x n products per pass.  Instances the solver declines (a cluster larger than
// MAXC, a residual above tolerance) are flagged and take the QL path of generation 2 in the same step.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define EIG3_HD __host__ __device__ __forceinline__
#else
#define EIG3_HD inline
#endif

namespace slam {
namespace eig3 {

constexpr double EPS = 2.220446049250313e-16;
constexpr int MAXC_DEFAULT = 16;          // largest cluster re-orthogonalised in the kernel; larger ones fall back to QL
constexpr int MAX_BISECT = 64;
// Gram-Schmidt inside a cluster: a (unit) vector that keeps less than this much of its norm^2 after the projections was nearly
// parallel to an earlier member, so the subtraction amplified its error -- only then are the cluster members refined by one
// more inverse-iteration step and orthogonalised again.  (Twisted-factorisation vectors of eigenvalues 1e-3 |T| ... 1e-9 |T|
// apart are orthogonal to ~eps |T| / gap already: the projections remove ~1e-13 ... 1e-7 and nothing needs repair.)
constexpr double REFINE_BELOW = 0.75;

// strided view of a per-thread work vector: element i of thread t lives at base[i * stride] (stride = threads' pitch, so the
// accesses of a warp coalesce)
struct Slot {
    double* p;
    int stride;
    EIG3_HD double get(int i) const { return p[(size_t)i * stride]; }
    EIG3_HD void set(int i, double v) const { p[(size_t)i * stride] = v; }
};

EIG3_HD uint32_t hi_word(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32);
#endif
}
EIG3_HD uint32_t lo_word(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}

// {d[i], e[i-1]^2}: what one term of the Sturm recurrence reads, as ONE 16-byte shared-memory load
struct alignas(16) De { double d, e2; };

EIG3_HD uint32_t shift_in_sign(const uint32_t bits, const double p) {      // bits << 1 | sign(p): one funnel shift on the device
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)__double2hiint(p), bits, 1);
#else
    return (bits << 1) | (hi_word(p) >> 31);
#endif
}
EIG3_HD int popcount32(const uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// Number of eigenvalues of the block [a, b) of T that are < x: sign changes of the Sturm sequence p_0 = 1, p_1 = d_a - x,
// p_i+1 = (d_i - x) p_i - e_i-1^2 p_i-1 (division free).  The signs are shifted into a word, one instruction per term, and
// counted by popc every 28 terms.  An exactly vanishing p_i needs no special case: with e_i != 0 inside a block its
// neighbours have opposite signs, so (s, +0, -s) and (s, -0, -s) both show exactly one change.  The pair (p_i-1, p_i) is
// rescaled by 2^-+400 when it leaves [2^-400, 2^400] (checked every fourth term: one term grows by at most ~(2 |T|)^2).
// WANT_F: also return the last term, p_n(x) = det(T_block - x I) = fm * 2^fs -- the function whose root the eigenvalue is.
struct Fval { double m; int s; };
template <bool WANT_F>
EIG3_HD int sturm_core(const De* de, const int a, const int b, const double x, Fval* f) {
    double pm1 = 1.0, p = de[a].d - x;
    uint32_t bits = shift_in_sign(0u, p);            // bit 1: sign of p_0 (+), bit 0: sign of p_1
    int cnt = (int)(bits & 1u);
    int sexp = 0;
    bits &= 1u;
#ifdef EIG3_PREFETCH
    De nx0 = de[(a + 1 < b) ? a + 1 : a], nx1 = de[(a + 2 < b) ? a + 2 : a];     // the next two terms' operands, fetched ahead of the chain
#endif
    for (int i0 = a + 1; i0 < b; i0 += 28) {
        const int kmax = (b - i0 < 28) ? b - i0 : 28;
        for (int k = 0; k < kmax; ++k) {
#ifdef EIG3_PREFETCH
            const De v = nx0;
            nx0 = nx1;
            nx1 = de[(i0 + k + 2 < b) ? i0 + k + 2 : a];
#else
            const De v = de[i0 + k];
#endif
            const double pn = fma(v.d - x, p, -(v.e2 * pm1));
            pm1 = p; p = pn;
            bits = shift_in_sign(bits, p);
            if ((k & 3) == 3) {
                const uint32_t ex = hi_word(p) & 0x7ff00000u;
                if (ex - 0x26f00000u > 0x58f00000u - 0x26f00000u) {                 // one unsigned compare: outside [2^-400, 2^400] (rare)
                    const bool big = ex > 0x58f00000u;
                    const double sc = big ? 3.87259191484932e-121 : 2.5822498780869086e+120;    // 2^-400 : 2^400
                    p *= sc; pm1 *= sc;
                    if (WANT_F) sexp += big ? 400 : -400;
                }
            }
        }
        cnt += popcount32((bits ^ (bits >> 1)) & ((1u << kmax) - 1u));   // changes among the kmax + 1 most recent signs
        bits &= 1u;
    }
    if (WANT_F) { f->m = p; f->s = sexp; }
    return cnt;
}
EIG3_HD int sturm_count(const De* de, const int a, const int b, const double x) { return sturm_core<false>(de, a, b, x, nullptr); }
EIG3_HD int sturm_eval(const De* de, const int a, const int b, const double x, Fval* f) { return sturm_core<true>(de, a, b, x, f); }

// Gershgorin interval of the block [a, b), widened like dstebz
EIG3_HD void block_bounds(const double* d, const double* e, const int a, const int b, const double pivmin, double& lo, double& hi) {
    double g0 = 1e300, g1 = -1e300;
    for (int i = a; i < b; ++i) {
        const double r = (i > a ? fabs(e[i - 1]) : 0.0) + (i < b - 1 ? fabs(e[i]) : 0.0);
        g0 = fmin(g0, d[i] - r); g1 = fmax(g1, d[i] + r);
    }
    const double bn = fmax(fabs(g0), fabs(g1));
    lo = g0 - 2.0 * EPS * bn * (double)(b - a) - 2.0 * pivmin;
    hi = g1 + 2.0 * EPS * bn * (double)(b - a) + 2.0 * pivmin;
}

// bisection on [lo, hi] (count(lo) <= j < count(hi)) down to hi - lo <= tol
EIG3_HD double bisect_bracket(const De* de, const int a, const int b, const int j, double lo, double hi, const double tol) {
    for (int it = 0; it < MAX_BISECT; ++it) {
        if (hi - lo <= tol) break;
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;                     // adjacent doubles
        if (sturm_count(de, a, b, mid) > j) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

// j-th smallest eigenvalue (0-based) of the block [a, b), to an ABSOLUTE accuracy of eps |T| (what a QL / QR iteration delivers as
// well: the Sturm count itself is only that accurate).  With the same tolerance for every eigenvalue all threads of a CTA run
// the same number of bisection steps, so nobody waits at the barrier behind a thread that is polishing a small eigenvalue.
EIG3_HD double bisect(const double* d, const double* e, const De* de, const int a, const int b, const int j, const double pivmin,
                      const double tnorm) {
    if (b - a == 1) return d[a];
    double lo, hi;
    block_bounds(d, e, a, b, pivmin, lo, hi);
    return bisect_bracket(de, a, b, j, lo, hi, 2.0 * EPS * tnorm + 2.0 * pivmin);
}

// Multisection start (the kernel runs it once per CTA before the root search): the threads of a block evaluate the Sturm count
// on a uniform grid of nb points over the block's Gershgorin interval, one point each, and share (point, count, p_n); every
// thread then starts from the grid cell that brackets its eigenvalue -- log2(nb) steps saved for one extra Sturm sweep.
EIG3_HD double grid_point(const double glo, const double ghi, const int j, const int nb) {
    return glo + ((double)j + 0.5) * ((ghi - glo) / (double)nb);
}
EIG3_HD void bracket_from_grid(const double* px, const int* pc, const int a, const int b, const int j, double& lo, double& hi) {
    for (int q = a; q < b; ++q) {
        const double x = px[q];
        if (pc[q] > j) hi = fmin(hi, x); else lo = fmax(lo, x);
    }
}

// State of one end of a bracket: the point, the Sturm count there and (when `known`) p_n there
struct End { double x; int c; Fval f; bool known; };
// the same search as bracket_from_grid, also returning which grid point (or -1: the Gershgorin end) each side came from
EIG3_HD void bracket_from_grid2(const double* px, const int* pc, const int a, const int b, const int j, double& lo, double& hi,
                                int& qlo, int& qhi) {
    qlo = -1; qhi = -1;
    for (int q = a; q < b; ++q) {
        const double x = px[q];
        if (pc[q] > j) { if (x < hi) { hi = x; qhi = q; } }
        else if (x > lo) { lo = x; qlo = q; }
    }
}

// j-th smallest eigenvalue of the block from a bracket lo.c <= j < hi.c, to hi - lo <= tol.  While the bracket holds more than
// one eigenvalue (or p_n is unknown at an end) the step is a bisection; once the eigenvalue is ISOLATED, p_n changes sign across
// the bracket exactly once and the step is a secant step on p_n with the Illinois modification (an end kept twice has its
// function value halved, which pulls that end in: superlinear, ~1.44 per evaluation, against one bit per evaluation).  Every
// step is decided by the Sturm count at the new point, so the bracket invariant is the bisection's.  Two safeguards:
//   * the proposal is kept 0.75 tol inside the bracket (Brent's minimal step: a root next to an end closes the bracket in one
//     move, so the search ends when the ESTIMATE has converged, not only when both ends have);
//   * a secant step that did not reduce |p_n| at its end by at least 4x was taken where p_n is far from linear (a wide bracket of
//     a degree-n polynomial): the next step is a bisection -- at worst two evaluations per halving.
EIG3_HD bool fval_reduced(const Fval& fnew, const Fval& fold) {      // |fnew| < |fold| / 4
    if (fnew.s != fold.s) return fnew.s < fold.s;
    return fabs(fnew.m) * 4.0 < fabs(fold.m);
}
EIG3_HD double root_bracket(const De* de, const int a, const int b, const int j, End lo, End hi, const double tol, int* nevals) {
    int side = 0, ne = 0;
    bool bisect_next = false;
    for (int it = 0; it < 2 * MAX_BISECT; ++it) {
        const double w = hi.x - lo.x;
        if (w <= tol) break;
        const int ds = hi.f.s - lo.f.s;
        bool secant = hi.c - lo.c == 1 && lo.known && hi.known && !bisect_next && w > 1.5 * tol && ds <= 400 && ds >= -400;
        double x = 0.5 * (lo.x + hi.x);
        if (secant) {
            double r = hi.f.m / lo.f.m;                               // p_n(hi) / p_n(lo) < 0
            if (ds != 0) r *= (ds > 0) ? 2.5822498780869086e+120 : 3.87259191484932e-121;
            if (r < 0.0) {
                x = lo.x + w / (1.0 - r);
                x = fmin(fmax(x, lo.x + 0.75 * tol), hi.x - 0.75 * tol);
            } else secant = false;
        }
        if (x <= lo.x || x >= hi.x) break;                           // adjacent doubles
        Fval f;
        const int c = sturm_eval(de, a, b, x, &f);
        ++ne;
        bisect_next = false;
        if (c > j) {
            if (secant) bisect_next = !fval_reduced(f, hi.f);
            hi.x = x; hi.c = c; hi.f = f; hi.known = true;
            if (side == 1) lo.f.m *= 0.5;
            side = 1;
        } else {
            if (secant) bisect_next = !fval_reduced(f, lo.f);
            lo.x = x; lo.c = c; lo.f = f; lo.known = true;
            if (side == -1) hi.f.m *= 0.5;
            side = -1;
        }
#ifdef EIG3_TRACE_ROOT
        if (EIG3_TRACE_ROOT(j)) fprintf(stderr, "    it %d secant %d x %.17g c %d f %.3e s %d  lo %.17g hi %.17g\n", it, (int)secant, x, c, f.m, f.s, lo.x, hi.x);
#endif
    }
    if (nevals) *nevals = ne;
    return 0.5 * (lo.x + hi.x);
}

// Twisted factorisation of T - x I on the block [a, b) and the eigenvector it yields.  z: the thread's vector (on exit the
// UNNORMALISED eigenvector, zero outside the block is the caller's business); w: the factors kept for a refinement step
// (D+_i for i < k, D-_i for i > k, gamma_k at k).  Returns the twist index k; *nrm2 = sum z_i^2.
EIG3_HD int twisted_vector(const double* d, const double* e, const double* e2, const int a, const int b, const double x,
                           const double pivf, const Slot z, const Slot w, double* nrm2) {
    // forward pivots D+ into z[a..b)
    double dp = d[a] - x;
    for (int i = a; i < b - 1; ++i) {
        if (fabs(dp) < pivf) dp = -pivf;
        z.set(i, dp);
        dp = (d[i + 1] - x) - e2[i] / dp;
    }
    if (fabs(dp) < pivf) dp = -pivf;
    z.set(b - 1, dp);
    // backward pivots D- on the fly; gamma_i = D+_i + D-_i - (d_i - x); twist at the smallest |gamma| (lowest index on ties)
    double dm = d[b - 1] - x;
    if (fabs(dm) < pivf) dm = -pivf;
    int k = b - 1;
    double gk = dp + dm - (d[b - 1] - x), best = fabs(gk);
    double dpn = z.get(b - 2);                                 // D+_i-1, fetched one step ahead of the division chain
    for (int i = b - 1; i > a; --i) {
        const double dpi = dpn;
        if (i - 2 >= a) dpn = z.get(i - 2);
        double dn = (d[i - 1] - x) - e2[i - 1] / dm;
        if (fabs(dn) < pivf) dn = -pivf;
        const double g = dpi + dn - (d[i - 1] - x);
        if (fabs(g) <= best) { best = fabs(g); gk = g; k = i - 1; }
        dm = dn;
    }
    // second backward sweep down to k + 1, this time keeping D-_i in z[i] (i > k)
    dm = d[b - 1] - x;
    if (fabs(dm) < pivf) dm = -pivf;
    for (int i = b - 1; i > k; --i) {
        z.set(i, dm);
        double dn = (d[i - 1] - x) - e2[i - 1] / dm;
        if (fabs(dn) < pivf) dn = -pivf;
        dm = dn;
    }
    if (fabs(gk) < pivf) gk = pivf;
    // the vector: z_k = 1, z_i = -(e_i / D+_i) z_i+1 below k, z_i+1 = -(e_i / D-_i+1) z_i above; factors move to w
    // (the ratios e_i / pivot do not depend on the running z: the pivots are fetched two ahead so the divisions pipeline)
    double s2 = 1.0, zi = 1.0;
    {
        double p0 = (k - 1 >= a) ? z.get(k - 1) : 1.0, p1 = (k - 2 >= a) ? z.get(k - 2) : 1.0;
        for (int i = k - 1; i >= a; --i) {
            const double piv = p0;
            p0 = p1;
            p1 = (i - 2 >= a) ? z.get(i - 2) : 1.0;
            zi = -(e[i] / piv) * zi;
            w.set(i, piv); z.set(i, zi);
            s2 += zi * zi;
        }
    }
    zi = 1.0;
    {
        double p0 = (k + 1 < b) ? z.get(k + 1) : 1.0, p1 = (k + 2 < b) ? z.get(k + 2) : 1.0;
        for (int i = k; i < b - 1; ++i) {
            const double piv = p0;
            p0 = p1;
            p1 = (i + 3 < b) ? z.get(i + 3) : 1.0;
            zi = -(e[i] / piv) * zi;
            w.set(i + 1, piv); z.set(i + 1, zi);
            s2 += zi * zi;
        }
    }
    z.set(k, 1.0); w.set(k, gk);
    *nrm2 = s2;
    return k;
}

// twisted_vector with every read-back BATCHED: the recurrences read pivots they (or an earlier sweep) stored in the strided work
// vectors -- global memory in the kernel, an L2 round trip each -- but no pivot depends on the running value of its chain, so
// PB of them are fetched one batch ahead and the chain of dependent divisions never waits for memory.  The backward pivots are
// parked in w during the first backward sweep (no second sweep).  Same arithmetic and same results as twisted_vector.
template <int PB>
EIG3_HD int twisted_vector_pf(const double* d, const double* e, const double* e2, const int a, const int b, const double x,
                              const double pivf, const Slot z, const Slot w, double* nrm2) {
    double dp = d[a] - x;
    for (int i = a; i < b - 1; ++i) {
        if (fabs(dp) < pivf) dp = -pivf;
        z.set(i, dp);
        dp = (d[i + 1] - x) - e2[i] / dp;
    }
    if (fabs(dp) < pivf) dp = -pivf;
    z.set(b - 1, dp);
    double dm = d[b - 1] - x;
    if (fabs(dm) < pivf) dm = -pivf;
    int k = b - 1;
    double gk = dp + dm - (d[b - 1] - x), best = fabs(gk);
    double cur[PB], nxt[PB];
    {   // backward sweep, i = b-1 .. a+1: gamma_i-1 needs D+_i-1 = z[i-1]
        int i0 = b - 1;
#pragma unroll
        for (int q = 0; q < PB; ++q) cur[q] = (i0 - q - 1 >= a) ? z.get(i0 - q - 1) : 0.0;
        while (i0 > a) {
#pragma unroll
            for (int q = 0; q < PB; ++q) nxt[q] = (i0 - PB - q - 1 >= a) ? z.get(i0 - PB - q - 1) : 0.0;
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int i = i0 - q;
                if (i > a) {
                    w.set(i, dm);                              // D-_i
                    double dn = (d[i - 1] - x) - e2[i - 1] / dm;
                    if (fabs(dn) < pivf) dn = -pivf;
                    const double g = cur[q] + dn - (d[i - 1] - x);
                    if (fabs(g) <= best) { best = fabs(g); gk = g; k = i - 1; }
                    dm = dn;
                }
            }
#pragma unroll
            for (int q = 0; q < PB; ++q) cur[q] = nxt[q];
            i0 -= PB;
        }
    }
    if (fabs(gk) < pivf) gk = pivf;
    double s2 = 1.0, zi = 1.0;
    {   // downward, i = k-1 .. a: pivot D+_i = z[i]; the factor moves to w[i], the component takes its place
        int i0 = k - 1;
#pragma unroll
        for (int q = 0; q < PB; ++q) cur[q] = (i0 - q >= a) ? z.get(i0 - q) : 1.0;
        while (i0 >= a) {
#pragma unroll
            for (int q = 0; q < PB; ++q) nxt[q] = (i0 - PB - q >= a) ? z.get(i0 - PB - q) : 1.0;
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int i = i0 - q;
                if (i >= a) {
                    const double piv = cur[q];
                    zi = -(e[i] / piv) * zi;
                    w.set(i, piv); z.set(i, zi);
                    s2 += zi * zi;
                }
            }
#pragma unroll
            for (int q = 0; q < PB; ++q) cur[q] = nxt[q];
            i0 -= PB;
        }
    }
    zi = 1.0;
    {   // upward, i = k .. b-2: pivot D-_i+1 = w[i+1] (parked above, stays there)
        int i0 = k;
#pragma unroll
        for (int q = 0; q < PB; ++q) cur[q] = (i0 + q + 1 < b) ? w.get(i0 + q + 1) : 1.0;
        while (i0 < b - 1) {
#pragma unroll
            for (int q = 0; q < PB; ++q) nxt[q] = (i0 + PB + q + 1 < b) ? w.get(i0 + PB + q + 1) : 1.0;
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int i = i0 + q;
                if (i < b - 1) {
                    zi = -(e[i] / cur[q]) * zi;
                    z.set(i + 1, zi);
                    s2 += zi * zi;
                }
            }
#pragma unroll
            for (int q = 0; q < PB; ++q) cur[q] = nxt[q];
            i0 += PB;
        }
    }
    z.set(k, 1.0); w.set(k, gk);
    *nrm2 = s2;
    return k;
}

// The same eigenvector in ONE work vector (the shared-memory variant of the kernel: a column of an n x n tile per thread, no
// second array): the pivots are overwritten by the vector's components as the recurrences consume them.  Nothing is kept for a
// refinement step -- the caller hands such (rare) instances to the two-array routine above.
EIG3_HD int twisted_vector1(const double* d, const double* e, const double* e2, const int a, const int b, const double x,
                            const double pivf, const Slot z, double* nrm2) {
    double dp = d[a] - x;
    for (int i = a; i < b - 1; ++i) {
        if (fabs(dp) < pivf) dp = -pivf;
        z.set(i, dp);
        dp = (d[i + 1] - x) - e2[i] / dp;
    }
    if (fabs(dp) < pivf) dp = -pivf;
    z.set(b - 1, dp);
    double dm = d[b - 1] - x;
    if (fabs(dm) < pivf) dm = -pivf;
    int k = b - 1;
    double gk = dp + dm - (d[b - 1] - x), best = fabs(gk);
    for (int i = b - 1; i > a; --i) {
        double dn = (d[i - 1] - x) - e2[i - 1] / dm;
        if (fabs(dn) < pivf) dn = -pivf;
        const double g = z.get(i - 1) + dn - (d[i - 1] - x);
        if (fabs(g) <= best) { best = fabs(g); gk = g; k = i - 1; }
        dm = dn;
    }
    dm = d[b - 1] - x;
    if (fabs(dm) < pivf) dm = -pivf;
    for (int i = b - 1; i > k; --i) {
        z.set(i, dm);
        double dn = (d[i - 1] - x) - e2[i - 1] / dm;
        if (fabs(dn) < pivf) dn = -pivf;
        dm = dn;
    }
    double s2 = 1.0, zi = 1.0;
    for (int i = k - 1; i >= a; --i) {
        zi = -(e[i] / z.get(i)) * zi;
        z.set(i, zi);
        s2 += zi * zi;
    }
    zi = 1.0;
    for (int i = k; i < b - 1; ++i) {
        zi = -(e[i] / z.get(i + 1)) * zi;
        z.set(i + 1, zi);
        s2 += zi * zi;
    }
    z.set(k, 1.0);
    *nrm2 = s2;
    return k;
}

// twisted_vector1 with the backward pivots D-_i PARKED in a second (global-memory) slot during the first backward sweep instead of
// being recomputed by a second sweep down to the twist: the stores are fire-and-forget, and the upward recurrence -- whose
// pivots do not depend on the running component -- fetches them back eight at a time, one batch ahead, so the round trip stays off
// the chain.  Removes up to n dependent divisions (n/2 on average, but the CTA waits for its slowest thread) from 3.5 n.
// Same arithmetic, same values as twisted_vector1.
EIG3_HD int twisted_vector1g(const double* d, const double* e, const double* e2, const int a, const int b, const double x,
                             const double pivf, const Slot z, const Slot g, double* nrm2) {
    double dp = d[a] - x;
    for (int i = a; i < b - 1; ++i) {
        if (fabs(dp) < pivf) dp = -pivf;
        z.set(i, dp);
        dp = (d[i + 1] - x) - e2[i] / dp;
    }
    if (fabs(dp) < pivf) dp = -pivf;
    z.set(b - 1, dp);
    double dm = d[b - 1] - x;
    if (fabs(dm) < pivf) dm = -pivf;
    int k = b - 1;
    double gk = dp + dm - (d[b - 1] - x), best = fabs(gk);
    for (int i = b - 1; i > a; --i) {
        g.set(i, dm);                                          // D-_i
        double dn = (d[i - 1] - x) - e2[i - 1] / dm;
        if (fabs(dn) < pivf) dn = -pivf;
        const double gg = z.get(i - 1) + dn - (d[i - 1] - x);
        if (fabs(gg) <= best) { best = fabs(gg); gk = gg; k = i - 1; }
        dm = dn;
    }
    double s2 = 1.0, zi = 1.0;
    for (int i = k - 1; i >= a; --i) {
        zi = -(e[i] / z.get(i)) * zi;
        z.set(i, zi);
        s2 += zi * zi;
    }
    zi = 1.0;
    {
        constexpr int PB = 8;
        double cur[PB], nxt[PB];
        int i0 = k;
#pragma unroll
        for (int q = 0; q < PB; ++q) cur[q] = (i0 + q + 1 < b) ? g.get(i0 + q + 1) : 1.0;
        while (i0 < b - 1) {
#pragma unroll
            for (int q = 0; q < PB; ++q) nxt[q] = (i0 + PB + q + 1 < b) ? g.get(i0 + PB + q + 1) : 1.0;
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int i = i0 + q;
                if (i < b - 1) {
                    zi = -(e[i] / cur[q]) * zi;
                    z.set(i + 1, zi);
                    s2 += zi * zi;
                }
            }
#pragma unroll
            for (int q = 0; q < PB; ++q) cur[q] = nxt[q];
            i0 += PB;
        }
    }
    z.set(k, 1.0);
    *nrm2 = s2;
    return k;
}

// One inverse-iteration step through the kept factors: solves N_k D_k N_k^T y = z in place (z <- y, unnormalised) on the
// block [a, b) with twist index k.  Returns sum y_i^2.
EIG3_HD double twisted_solve(const double* e, const int a, const int b, const int k, const Slot z, const Slot w) {
    // eliminate towards k from both ends:  y_i+1 -= (e_i / D+_i) y_i  (i < k),   y_i-1 -= (e_i-1 / D-_i) y_i  (i > k)
    double y = z.get(a);
    for (int i = a; i < k; ++i) {
        const double nxt = z.get(i + 1) - (e[i] / w.get(i)) * y;
        z.set(i + 1, nxt);
        y = nxt;
    }
    y = z.get(b - 1);
    for (int i = b - 1; i > k; --i) {
        const double nxt = z.get(i - 1) - (e[i - 1] / w.get(i)) * y;
        z.set(i - 1, nxt);
        y = nxt;
    }
    // diagonal, then back-substitute away from k
    double yk = z.get(k) / w.get(k), s2 = yk * yk;
    z.set(k, yk);
    y = yk;
    for (int i = k - 1; i >= a; --i) {
        const double piv = w.get(i);
        y = z.get(i) / piv - (e[i] / piv) * y;
        z.set(i, y);
        s2 += y * y;
    }
    y = yk;
    for (int i = k; i < b - 1; ++i) {
        const double piv = w.get(i + 1);
        y = z.get(i + 1) / piv - (e[i] / piv) * y;
        z.set(i + 1, y);
        s2 += y * y;
    }
    return s2;
}

// max_i |((T - x I) z)_i| over the block, for z as stored (scaled by `scale`)
EIG3_HD double residual_inf(const double* d, const double* e, const int a, const int b, const double x, const Slot z, const double scale) {
    double r = 0.0, zm = 0.0, zc = z.get(a);
    for (int i = a; i < b; ++i) {
        const double zn = (i + 1 < b) ? z.get(i + 1) : 0.0;
        const double v = (i > a ? e[i - 1] * zm : 0.0) + (d[i] - x) * zc + (i + 1 < b ? e[i] * zn : 0.0);
        r = fmax(r, fabs(v));
        zm = zc; zc = zn;
    }
    return r * fabs(scale);
}

// residual_inf with the vector fetched PB components at a time (global work vectors: one round trip per batch, not per component)
template <int PB>
EIG3_HD double residual_inf_pf(const double* d, const double* e, const int a, const int b, const double x, const Slot z) {
    double r = 0.0, zm = 0.0, zc = z.get(a);
    double buf[PB];
    for (int i0 = a; i0 < b; i0 += PB) {
#pragma unroll
        for (int q = 0; q < PB; ++q) buf[q] = (i0 + q + 1 < b) ? z.get(i0 + q + 1) : 0.0;
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            const int i = i0 + q;
            if (i < b) {
                const double zn = buf[q];
                const double v = (i > a ? e[i - 1] * zm : 0.0) + (d[i] - x) * zc + (i + 1 < b ? e[i] * zn : 0.0);
                r = fmax(r, fabs(v));
                zm = zc; zc = zn;
            }
        }
    }
    return r;
}

}  // namespace eig3
}  // namespace slam
