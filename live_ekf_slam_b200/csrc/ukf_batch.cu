// ukf_batch.cu -- batched UKF-SLAM step for sm_100a: one CTA per filter instance.
//
// Restates UKF::update, ekf_ws/src/localization_pkg/src/ukf.cpp:161-371 (nearestSPD :106-123, motionModel :125-135,
// sensingModel :137-159, predictionStage :197-241, updateStage :243-291, landmarkUpdate :293-349,
// landmarkInsertion :351-371) with the reference's float/double roundings.
//
// What the reference does with dense n x (2n+1) sigma-point matrices is evaluated here from the symmetric
// eigendecomposition  Y = scale * sym(P) = Z D Z^T  alone (Z kept transposed in shared memory):
//   * sqrt(nearestSPD) = S = Z sqrt(D+) Z^T is never formed: only rows 0..3 of S (vehicle rows of the sigma
//     points), the two rows of the landmark being updated, and products S*v are needed             -> O(n^2)
//   * the motion model only changes rows 0..3, so for landmark rows a,b >= 4
//       P_pred[a][b] = 2w (S S^T)[a][b] + (sum w) e_a e_b,   S S^T = Y + sum_{clipped k} (1e-8 - d_k) z_k z_k^T
//     and the vehicle/landmark cross block is  w * S (Xp_i - Xp_{i+n}) + e_b * sum_i w_i dv_i        -> O(n^2)
//   * cross covariance of a landmark update: C[a] = f_a * sum_i w_i dz_i + w * S (dz_i - dz_{i+n})   -> O(n^2)
// so the step is dominated by the eigendecomposition (Householder tridiagonalisation + implicit QL, 9 n^3 nominal):
// the UKF batch is FP64-compute bound, not HBM bound.  P stays in global memory (L2) with a fixed
// leading dimension; only Z lives in shared memory, which lets two CTAs share an SM.
//
// Two generations of the step live here.  Generation 2 (default; second half of the file) never forms the eigenvector
// matrix: front2 (tridiagonalisation only) -> ukf_ql_kernel -> back2 (warp per instance; S*v through the reflectors and
// the rotation log), with the generation-1 kernels below kept as its rescue pass and as the fall-back for handle
// capacities the generation-2 tile cannot hold.  Generation 1 is THREE launches, split where the parallelism changes shape:
//   ukf_front_kernel  CTA per instance: Y, Householder tridiagonalisation, explicit Q^T  -> HBM scratch (d, e, Q^T)
//   ukf_ql_kernel     THREAD per instance: implicit QL on (d, e) -- a serial chain of ~0.85 n^2 plane rotations, each a
//                     dependent rsqrt; one lane per filter runs thousands of these chains side by side instead of
//                     stalling a whole CTA on one thread -- eigenvalues + rotation log -> HBM scratch
//   ukf_back_kernel   CTA per instance: replays the rotation log on Q^T (thread per eigenvector component), then the
//                     sigma-point algebra, landmark updates and insertions.
#include "common.cuh"
#include "eig3.cuh"

#include <climits>

namespace slam {

constexpr int UKF_THREADS = 256;
constexpr int UKF_WARPS = UKF_THREADS / 32;

struct UkfSmem {
    double* A;      // n_max x lds : Y, then Q, then Z^T
    double* x;      // prior x_t
    double* xp;     // running x_pred
    double* d;      // eigenvalues -> sqrt(max(d,1e-8))
    double* e;      // off-diagonal / scratch
    double* Xp;     // [4][ns_max] propagated vehicle rows of the sigma points
    double* S4;     // [4][nmp] rows 0..3 of S ; later z0|z1 [ns_max] each
    double* K;      // [nmp][2]
    double* pool;   // 8*nmp doubles of phase-local scratch
    double* red;    // 160 doubles reduction scratch
    double* corr;   // clipped-eigenvalue corrections (1e-8 - d_k), aligned with clip[]
    int* clip;      // indices of clipped eigenvalues
    int* ids;
    float* meas;
    int* assoc;
    int* iscr;
};

__host__ __device__ inline size_t ukf_smem_carve(const BatchState& b, unsigned char* base, UkfSmem* s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const int nmp = ldg_of(b.n_max), nsm = 2 * b.n_max + 2;
    size_t oA = take(sizeof(double) * (size_t)b.n_max * b.lds);
    size_t ox = take(sizeof(double) * nmp), oxp = take(sizeof(double) * nmp);
    size_t od = take(sizeof(double) * nmp), oe = take(sizeof(double) * nmp);
    size_t oXp = take(sizeof(double) * 4 * nsm);
    size_t oS4 = take(sizeof(double) * (4 * nmp > 2 * nsm ? 4 * nmp : 2 * nsm));
    size_t oK = take(sizeof(double) * 2 * nmp);
    size_t opool = take(sizeof(double) * 8 * nmp);
    size_t ored = take(sizeof(double) * 160);   // block_sum<14> needs 14*8; QL control words live at +140
    size_t ocorr = take(sizeof(double) * nmp);
    size_t oclip = take(sizeof(int) * nmp);
    size_t oids = take(sizeof(int) * (b.max_lm + 1));
    size_t omeas = take(sizeof(float) * 3 * b.max_meas);
    size_t oassoc = take(sizeof(int) * b.max_meas);
    size_t oi = take(sizeof(int) * 8);
    if (s) {
        s->A = (double*)(base + oA); s->x = (double*)(base + ox); s->xp = (double*)(base + oxp);
        s->d = (double*)(base + od); s->e = (double*)(base + oe); s->Xp = (double*)(base + oXp);
        s->S4 = (double*)(base + oS4); s->K = (double*)(base + oK); s->pool = (double*)(base + opool);
        s->red = (double*)(base + ored); s->corr = (double*)(base + ocorr); s->clip = (int*)(base + oclip);
        s->ids = (int*)(base + oids); s->meas = (float*)(base + omeas); s->assoc = (int*)(base + oassoc);
        s->iscr = (int*)(base + oi);
    }
    return off;
}

// sum NV per-thread values over the CTA; every thread receives the totals.  Two barriers.
template <int NV, int NWARPS = UKF_WARPS>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[k * NWARPS + warp] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) t += red[k * NWARPS + w];
        v[k] = t;
    }
    __syncthreads();
}

__device__ __forceinline__ float yaw_of(double c, double s) { return (float)remainder(atan2(s, c), TWO_PI_REF); }

// ---------------------------------------------------------------------------------------------------------
// Symmetric eigendecomposition of the n x n matrix in A (full storage, leading dimension lds), in two parts.
//
// tridiag_q: Householder tridiagonalisation (LAPACK dsytd2 'L' organisation) and the explicit orthogonal factor
// (dorg2r organisation), then an in-place transpose: on exit d / e hold the tridiagonal matrix (e[k] couples k and
// k+1) and A holds Q^T.  Threads are mapped 2-D on the trailing block: consecutive lanes own consecutive columns
// (conflict-free shared-memory rows), G = 256 / columns row groups split the long inner loops, partial sums meet in
// `part` ([8][n]).  Scratch: v, p, w, tau (n each).
//
// ql_serial: implicit QL on (d, e) (EISPACK tql2 organisation) with each sweep's rotations generated by one thread
// and applied to A = Z^T by one thread per column.  Only the fall-back of the back kernel (rotation log overflow).
// ---------------------------------------------------------------------------------------------------------
// One Householder step's scalars for column k (rows k+1..n-1 hold the column below the diagonal): warp 0 only.
// Writes tau[k], e[k], d[k] and (tk, scal) into slot[0..1].
__device__ __forceinline__ void householder_scalars(const double* A, const int lds, const int n, const int k, const int lane,
                                                    double* d, double* e, double* tau, double* slot) {
    double ss = 0.0;
    for (int i = k + 2 + lane; i < n; i += 32) { const double a = A[(size_t)i * lds + k]; ss += a * a; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
        const double alpha = A[(size_t)(k + 1) * lds + k];
        double tk = 0.0, scal = 0.0, beta = alpha;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            tk = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        tau[k] = tk; e[k] = beta; d[k] = A[(size_t)k * lds + k];
        slot[0] = tk; slot[1] = scal;
    }
}

template <int NT>
__device__ void tridiag(double* A, const int lds, const int n, double* d, double* e, double* scratch, double* part, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* v = scratch;            // [n]
    double* p = scratch + n;        // [n]
    double* w = scratch + 2 * n;    // [n]
    double* tau = scratch + 3 * n;  // [n]

    // ---- reduction to tridiagonal form: A = Q T Q^T.  The scalar chain of a step (column norm, sqrt, two divisions: ~20 % of the
    // kernel when every warp waited for it at a barrier) is OVERLAPPED with the previous step's rank-2 update: warp 0 updates
    // column k+1 of the trailing block first and goes on to the next step's scalars while the other warps update the rest.
    // (tk, scal) of step k live in red[2 (k & 1) ..]: double buffered, so the readers of step k never race the writer of k+1.
    if (warp == 0 && n > 1) householder_scalars(A, lds, n, 0, lane, d, e, tau, red);
    __syncthreads();
    for (int k = 0; k < n - 1; ++k) {
        const int m0 = k + 1;                      // first row/col of the trailing block
        const int m = n - m0;                      // its size
        const double tk = red[2 * (k & 1)], scal = red[2 * (k & 1) + 1];
        double* next_slot = red + 2 * ((k + 1) & 1);
        if (tk != 0.0) {                            // uniform
            for (int i = m0 + tid; i < n; i += NT) {
                const double vi = (i == m0) ? 1.0 : A[(size_t)i * lds + k] * scal;
                v[i] = vi;
                A[(size_t)i * lds + k] = vi;        // keep the reflector in column k
            }
            __syncthreads();
            // 2-D mapping: thread -> (column PAIR c, c+1 with c even, row group g); rows of group g: m0 + g*rb .. (exclusive end
            // clipped).  One 16-byte load feeds two accumulators and v[i] is fetched once per pair: 4 instructions per two
            // elements instead of 6, and twice as many row groups for the same threads.  When m0 is odd the first pair starts
            // in column k (the finished reflector): that half of the result is simply not stored.
            const int c0 = m0 & ~1, np = (n - c0) >> 1;
            int G = NT / np; if (G > 8) G = 8; if (G < 1) G = 1;
            const int g = tid / np, pl = tid - g * np;
            const int rb = (m + G - 1) / G;
            const bool act = g < G;
            const int c = c0 + 2 * pl;
            const int r_lo = m0 + g * rb, r_hi = (r_lo + rb < n) ? r_lo + rb : n;
            // p = tau * A22 * v  (columns c, c+1 of the symmetric block, conflict-free), and p^T v
            if (np > NT) {                 // (never for n_max <= 256) one group, strided columns
                for (int cc = m0 + tid; cc < n; cc += NT) {
                    double acc = 0.0;
                    for (int i = m0; i < n; ++i) acc += A[(size_t)i * lds + cc] * v[i];
                    part[cc] = acc;
                }
            } else if (act) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
                for (int i = r_lo; i < r_hi; ++i) {
                    const double2 a = *reinterpret_cast<const double2*>(A + (size_t)i * lds + c);
                    const double vi = v[i];
                    a0 = fma(a.x, vi, a0); a1 = fma(a.y, vi, a1);
                }
                if (c >= m0) part[g * n + c] = a0;
                part[g * n + c + 1] = a1;
            }
            __syncthreads();
            double pv[1] = {0.0};
            const int Gs = (np > NT) ? 1 : G;
            for (int cc = m0 + tid; cc < n; cc += NT) {
                double acc = 0.0;
                for (int q = 0; q < Gs; ++q) acc += part[q * n + cc];
                acc *= tk;
                p[cc] = acc;
                pv[0] += acc * v[cc];
            }
            block_sum<1, NT / 32>(pv, red + 8);
            const double a2 = -0.5 * tk * pv[0];
            for (int cc = m0 + tid; cc < n; cc += NT) w[cc] = p[cc] + a2 * v[cc];
            __syncthreads();
            // A22 -= v w^T + w v^T   (both triangles kept; the two products are added symmetrically)
            if (NT > 32 && warp == 0) {
                // column m0 (rows m0..n-1), then the NEXT step's scalars from it; row m0 beyond the diagonal is dead from here on
                const double vc = v[m0], wc = w[m0];
                for (int i = m0 + lane; i < n; i += 32) {
                    const double t1 = __dmul_rn(v[i], wc), t2 = __dmul_rn(w[i], vc);
                    A[(size_t)i * lds + m0] -= __dadd_rn(t1, t2);
                }
                __syncwarp();
                if (m0 < n - 1) householder_scalars(A, lds, n, m0, lane, d, e, tau, next_slot);
            } else {
                // the rest of the block: rows and columns m0+1..n-1, on the remaining NT - 32 threads
                const int c_lo = (NT > 32) ? m0 + 1 : m0, mm = n - c_lo;
                const int t2 = (NT > 32) ? tid - 32 : tid, T2 = (NT > 32) ? NT - 32 : NT;
                if (mm > 0) {
                    if (mm > T2) {
                        for (int cc = c_lo + t2; cc < n; cc += T2) {
                            const double vc = v[cc], wc = w[cc];
                            for (int i = c_lo; i < n; ++i) {
                                const double t1 = __dmul_rn(v[i], wc), t2b = __dmul_rn(w[i], vc);
                                A[(size_t)i * lds + cc] -= __dadd_rn(t1, t2b);
                            }
                        }
                    } else {
                        // column pairs again (first pair may start in column m0, which warp 0 owns: that half is not stored)
                        const int e0 = c_lo & ~1, np2 = (n - e0) >> 1;
                        int G2 = T2 / np2; if (G2 > 8) G2 = 8; if (G2 < 1) G2 = 1;
                        const int g2 = t2 / np2, pl2 = t2 - g2 * np2;
                        const int rb2 = (mm + G2 - 1) / G2;
                        if (g2 < G2) {
                            const int c2 = e0 + 2 * pl2;
                            const int lo2 = c_lo + g2 * rb2, hi2 = (lo2 + rb2 < n) ? lo2 + rb2 : n;
                            const bool both = c2 >= c_lo;
                            const double vc0 = v[c2], wc0 = w[c2], vc1 = v[c2 + 1], wc1 = w[c2 + 1];
#pragma unroll 2
                            for (int i = lo2; i < hi2; ++i) {
                                double2 a = *reinterpret_cast<const double2*>(A + (size_t)i * lds + c2);
                                const double vi = v[i], wi = w[i];
                                a.y -= __dadd_rn(__dmul_rn(vi, wc1), __dmul_rn(wi, vc1));
                                if (both) {
                                    a.x -= __dadd_rn(__dmul_rn(vi, wc0), __dmul_rn(wi, vc0));
                                    *reinterpret_cast<double2*>(A + (size_t)i * lds + c2) = a;
                                } else A[(size_t)i * lds + c2 + 1] = a.y;
                            }
                        }
                    }
                }
                if (NT == 32) { __syncwarp(); if (m0 < n - 1) householder_scalars(A, lds, n, m0, lane, d, e, tau, next_slot); }
            }
        } else if (warp == 0 && m0 < n - 1) householder_scalars(A, lds, n, m0, lane, d, e, tau, next_slot);   // nothing to update
        __syncthreads();
    }
    if (tid == 0) { d[n - 1] = A[(size_t)(n - 1) * lds + (n - 1)]; e[n - 1] = 0.0; }
    __syncthreads();
}

// ---- The same reduction on the PACKED lower triangle (row i holds columns 0..i at tri_off(i)): half the shared memory, so the
// front kernel of generations 2 / 3 keeps FOUR CTAs per SM up to n = 109 where the full square allowed two from n = 90 (the
// Householder steps are latency bound: residency is the lever).  The symmetric product reads A(i, c) through
// the symmetric address (column part: consecutive lanes, consecutive words; row part: a thread walks its own row, the rows of a
// warp start at triangular offsets = two-way bank conflicts), the rank-2 update touches each stored element once (no mirror,
// exact symmetry by construction, two FMAs per element).  Same reflector storage on exit: v_k in column k below the diagonal.
__device__ __forceinline__ int tri_off(const int i) { return (i * (i + 1)) >> 1; }

__device__ __forceinline__ void householder_scalars_packed(const double* A, const int n, const int k, const int lane,
                                                           double* d, double* e, double* tau, double* slot) {
    double ss = 0.0;
    for (int i = k + 2 + lane; i < n; i += 32) { const double a = A[tri_off(i) + k]; ss += a * a; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
        const double alpha = A[tri_off(k + 1) + k];
        double tk = 0.0, scal = 0.0, beta = alpha;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            tk = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        tau[k] = tk; e[k] = beta; d[k] = A[tri_off(k) + k];
        slot[0] = tk; slot[1] = scal;
    }
}

constexpr int TRIP_GMAX = 4;             // row groups of the symmetric product (partial sums in part[TRIP_GMAX][n])
template <int NT>
__device__ void tridiag_packed(double* A, const int n, double* d, double* e, double* scratch, double* part, double* red) {
    static_assert(NT > 32, "warp 0 runs one step ahead of the other warps");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* v = scratch;            // [n]
    double* p = scratch + n;        // [n]
    double* w = scratch + 2 * n;    // [n]
    double* tau = scratch + 3 * n;  // [n]
    if (warp == 0 && n > 1) householder_scalars_packed(A, n, 0, lane, d, e, tau, red);
    __syncthreads();
    for (int k = 0; k < n - 1; ++k) {
        const int m0 = k + 1;                      // first row/col of the trailing block
        const int m = n - m0;                      // its size
        const double tk = red[2 * (k & 1)], scal = red[2 * (k & 1) + 1];
        double* next_slot = red + 2 * ((k + 1) & 1);
        if (tk != 0.0) {                            // uniform
            for (int i = m0 + tid; i < n; i += NT) {
                const int o = tri_off(i) + k;
                const double vi = (i == m0) ? 1.0 : A[o] * scal;
                v[i] = vi;
                A[o] = vi;                          // keep the reflector in column k
            }
            __syncthreads();
            // thread -> (index c, row group g): p[c] = sum_i A(i, c) v[i] over the rows of the group; A(i, c) = stored(c, i) for i < c
            // (the thread's own row, contiguous), stored(i, c) for i >= c (column c, the lanes of a warp side by side)
            int G = NT / m; if (G > TRIP_GMAX) G = TRIP_GMAX; if (G < 1) G = 1;
            const int g = tid / m, cl = tid - g * m;
            const int rb = (m + G - 1) / G;
            const int c = m0 + cl;
            if (g < G) {                            // (m <= NT: n_max <= 256)
                const int r_lo = m0 + g * rb, r_hi = (r_lo + rb < n) ? r_lo + rb : n;
                int r_mid = c > r_lo ? c : r_lo; if (r_mid > r_hi) r_mid = r_hi;
                double a0 = 0.0, a1 = 0.0;
                const double* row = A + tri_off(c);
                int i = r_lo;
                for (; i + 1 < r_mid; i += 2) { a0 = fma(row[i], v[i], a0); a1 = fma(row[i + 1], v[i + 1], a1); }
                if (i < r_mid) { a0 = fma(row[i], v[i], a0); ++i; }
                int o = tri_off(i) + c;
                for (; i + 1 < r_hi; i += 2) {
                    a0 = fma(A[o], v[i], a0);
                    a1 = fma(A[o + i + 1], v[i + 1], a1);
                    o += 2 * i + 3;                 // tri_off(i + 2) - tri_off(i)
                }
                if (i < r_hi) a0 = fma(A[o], v[i], a0);
                part[g * n + c] = a0 + a1;
            }
            __syncthreads();
            double pv[1] = {0.0};
            for (int cc = m0 + tid; cc < n; cc += NT) {
                double acc = 0.0;
                for (int q = 0; q < G; ++q) acc += part[q * n + cc];
                acc *= tk;
                p[cc] = acc;
                pv[0] += acc * v[cc];
            }
            block_sum<1, NT / 32>(pv, red + 8);
            const double a2 = -0.5 * tk * pv[0];
            for (int cc = m0 + tid; cc < n; cc += NT) w[cc] = p[cc] + a2 * v[cc];
            __syncthreads();
            // stored part of A22 -= v w^T + w v^T
            if (warp == 0) {
                // column m0 (rows m0..n-1), then the NEXT step's scalars from it
                const double vc = v[m0], wc = w[m0];
                for (int i = m0 + lane; i < n; i += 32) {
                    const int o = tri_off(i) + m0;
                    A[o] = fma(-v[i], wc, fma(-w[i], vc, A[o]));
                }
                __syncwarp();
                if (m0 < n - 1) householder_scalars_packed(A, n, m0, lane, d, e, tau, next_slot);
            } else {
                // columns m0+1..n-1 on the remaining NT - 32 threads: thread -> (column c2, row group), rows >= c2 only
                const int c_lo = m0 + 1, mm = n - c_lo;
                const int t2 = tid - 32, T2 = NT - 32;
                if (mm > 0) {
                    int G2 = T2 / mm; if (G2 > 8) G2 = 8; if (G2 < 1) G2 = 1;
                    const int g2 = t2 / mm, cl2 = t2 - g2 * mm;
                    const int rb2 = (mm + G2 - 1) / G2;
                    if (g2 < G2) {
                        const int c2 = c_lo + cl2;
                        const int lo2 = c_lo + g2 * rb2, hi2 = (lo2 + rb2 < n) ? lo2 + rb2 : n;
                        const double vc = v[c2], wc = w[c2];
                        int i = lo2 > c2 ? lo2 : c2;
                        int o = tri_off(i) + c2;
#pragma unroll 2
                        for (; i < hi2; ++i) {
                            A[o] = fma(-v[i], wc, fma(-w[i], vc, A[o]));
                            o += i + 1;
                        }
                    }
                }
            }
        } else if (warp == 0 && m0 < n - 1) householder_scalars_packed(A, n, m0, lane, d, e, tau, next_slot);   // nothing to update
        __syncthreads();
    }
    if (tid == 0) { d[n - 1] = A[tri_off(n - 1) + (n - 1)]; e[n - 1] = 0.0; }
    __syncthreads();
}

// explicit orthogonal factor from the reflectors left in A by tridiag (tau in scratch[3n..4n)), transposed in place
__device__ void form_qt(double* A, const int lds, const int n, double* scratch, double* part) {
    const int tid = threadIdx.x;
    double* tau = scratch + 3 * n;  // [n]
    // ---- explicit Q in place.  Shift the reflectors one column to the right (reflector k -> column k+1),
    //      first row/column of Q = unit vector, then dorg2r on the (n-1) x (n-1) trailing block.
    for (int i = tid; i < n; i += UKF_THREADS) {          // one thread per row, high column to low: no hazard
        for (int k = i - 2; k >= 0; --k) A[(size_t)i * lds + k + 1] = A[(size_t)i * lds + k];
        A[(size_t)i * lds + 0] = (i == 0) ? 1.0 : 0.0;
        if (i > 0) A[i] = 0.0;                            // row 0
    }
    __syncthreads();
    // B = A[1:,1:] ; reflector j (tau[j]) sits in B column j below the diagonal; last column = unit vector
    if (n >= 2) {
        for (int i = tid; i < n - 1; i += UKF_THREADS)
            A[(size_t)(1 + i) * lds + (n - 1)] = (i == n - 2) ? 1.0 : 0.0;
        __syncthreads();
    }
    for (int j = n - 3; j >= 0; --j) {
        const double tj = tau[j];
        const int r0 = 1 + j;                              // row/col of B[j][j] in A
        const int m = n - (r0 + 1);                        // columns r0+1 .. n-1 (and rows)
        // apply H_j to B[j:, j+1:] from the left: t[c] = tau * v^T B[:,c] (v[j] = 1, B[j][c] = 0 on entry)
        int G = UKF_THREADS / m; if (G > 8) G = 8; if (G < 1) G = 1;
        const int g = tid / m, cl = tid - g * m;
        const int rb = (m + G - 1) / G;
        const bool act = g < G;
        const int c = r0 + 1 + cl;
        const int r_lo = r0 + 1 + g * rb, r_hi = (r_lo + rb < n) ? r_lo + rb : n;
        if (act) {
            double a0 = 0.0, a1 = 0.0;
            int i = r_lo;
            for (; i + 1 < r_hi; i += 2) {
                a0 += A[(size_t)i * lds + r0] * A[(size_t)i * lds + c];
                a1 += A[(size_t)(i + 1) * lds + r0] * A[(size_t)(i + 1) * lds + c];
            }
            if (i < r_hi) a0 += A[(size_t)i * lds + r0] * A[(size_t)i * lds + c];
            part[g * n + c] = a0 + a1;
        }
        __syncthreads();
        if (act) {
            double t = 0.0;
            for (int q = 0; q < G; ++q) t += part[q * n + c];
            t *= tj;
            if (g == 0) A[(size_t)r0 * lds + c] = -t;
            for (int i = r_lo; i < r_hi; ++i) A[(size_t)i * lds + c] -= A[(size_t)i * lds + r0] * t;
        }
        __syncthreads();
        for (int i = r0 + tid; i < n; i += UKF_THREADS)
            A[(size_t)i * lds + r0] = (i == r0) ? 1.0 - tj : -tj * A[(size_t)i * lds + r0];
        for (int i = 1 + tid; i < r0; i += UKF_THREADS) A[(size_t)i * lds + r0] = 0.0;
        __syncthreads();
    }

    // ---- transpose in place: A <- Q^T so that a QL rotation touches two ROWS (conflict-free per-row threads)
    for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
        const int i = idx / n, j = idx - i * n;
        if (j > i) {
            const double a = A[(size_t)i * lds + j], bb = A[(size_t)j * lds + i];
            A[(size_t)i * lds + j] = bb; A[(size_t)j * lds + i] = a;
        }
    }
    __syncthreads();
}

__device__ void ql_serial(double* A, const int lds, const int n, double* d, double* e, double* scratch, double* red) {
    const int tid = threadIdx.x;
    double* csc = scratch;          // [n]
    double* css = scratch + n;      // [n]
    int* ctl = reinterpret_cast<int*>(red + 140);          // [0]=l-range start, [1]=m, [2]=done flag
    double f = 0.0, tst1 = 0.0;                            // live in thread 0 only
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            if (tid == 0) {
                if (iter == 0) { const double t = fabs(d[l]) + fabs(e[l]); if (t > tst1) tst1 = t; }
                int m = l;
                while (m < n - 1) { if (fabs(e[m]) <= eps * tst1) break; ++m; }
                if (m == l || iter >= 60) { d[l] += f; e[l] = 0.0; ctl[2] = 1; }
                else {
                    ctl[2] = 0; ctl[1] = m;
                    double g = d[l];
                    double pp = (d[l + 1] - g) / (2.0 * e[l]);
                    double r = sqrt(pp * pp + 1.0);
                    if (pp < 0) r = -r;
                    d[l] = e[l] / (pp + r);
                    d[l + 1] = e[l] * (pp + r);
                    const double dl1 = d[l + 1];
                    double h = g - d[l];
                    for (int i = l + 2; i < n; ++i) d[i] -= h;
                    f += h;
                    pp = d[m];
                    double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
                    const double el1 = e[l + 1];
                    for (int i = m - 1; i >= l; --i) {
                        c3 = c2; c2 = c; s2 = s;
                        g = c * e[i];
                        h = c * pp;
                        const double rinv = rsqrt(pp * pp + e[i] * e[i]);
                        r = (pp * pp + e[i] * e[i]) * rinv;
                        e[i + 1] = s * r;
                        s = e[i] * rinv;
                        c = pp * rinv;
                        pp = c * d[i] - s * g;
                        d[i + 1] = h + s * (c * g + s * d[i]);
                        csc[i] = c; css[i] = s;
                    }
                    pp = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * pp;
                    d[l] = c * pp;
                }
            }
            __syncthreads();
            if (ctl[2]) break;
            const int m = ctl[1];
            // apply the sweep's rotations to Z^T: thread r owns component r of every eigenvector
            for (int r = tid; r < n; r += UKF_THREADS) {
                double fz = A[(size_t)m * lds + r];
                for (int i = m - 1; i >= l; --i) {
                    const double zi = A[(size_t)i * lds + r];
                    const double c = csc[i], s = css[i];
                    A[(size_t)(i + 1) * lds + r] = s * zi + c * fz;
                    fz = c * zi - s * fz;
                }
                A[(size_t)l * lds + r] = fz;
            }
            ++iter;
            __syncthreads();
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// ukf_ql_kernel: implicit QL (EISPACK tql2 organisation, the same recurrences as ql_serial) on the tridiagonal
// matrices of ALL instances at once, one THREAD per instance.  (d, e) live in shared memory as [k][lane]; every
// plane rotation (c, s) is appended to the instance's log in HBM together with the (l, m) range of its sweep, for the
// back kernel to replay on Q^T.  On success the eigenvalues overwrite dg and nswp = number of sweeps; if a log would
// overflow, nswp = -1 and dg / eg stay untouched (the back kernel then runs ql_serial itself).
// ---------------------------------------------------------------------------------------------------------
constexpr int QL_LANES = 32;

// 1/sqrt(x) for normal positive x to ~1 ulp: rsqrt.approx.ftz.f64 (relative error < 2^-22.9) refined by one third-order step
// y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2: four dependent FP64 operations instead of the library routine's two Newton steps
// plus range fix-ups (the arguments here are sums of squares of matrix entries, never subnormal).
__device__ __forceinline__ double ql_rsqrt(const double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// SMEM = false: (d, e) are worked on in place in the HBM scratch ([k][instance], coalesced over the lanes, L1/L2 cached; the
// operands of a rotation are fetched one rotation ahead), so the kernel needs no shared memory and its 32-thread CTAs
// co-reside with the other slices' front / back kernels.  On overflow dg / eg are then clobbered: only for callers that
// redo the tridiagonalisation (the generation-2 rescue pass).
template <bool SMEM>
__global__ void __launch_bounds__(QL_LANES)
ukf_ql_kernel(UkfScratch u, const int4* __restrict__ meta, const int batch, const int i0, const int i1, const int only_flagged) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int inst = i0 + blockIdx.x * QL_LANES + lane;
    if (inst >= i1) return;
    if (meta[inst].y & SLAM_STATUS_SAME_STEP_REMATCH) return;
    if (only_flagged && u.nswp[inst] != -2) return;     // generation 3: only the instances the parallel eigensolver declined
    const int n = 4 + 2 * meta[inst].x;
    double* sd = reinterpret_cast<double*>(smem_raw);           // [n_max][32]
    double* se = sd + (size_t)u.n_max * QL_LANES;               // [n_max][32]
    double* const gd = u.dg + inst;
    double* const ge = u.eg + inst;
#define D_(k) (*(SMEM ? sd + (k) * QL_LANES + lane : gd + (size_t)(k) * batch))
#define E_(k) (*(SMEM ? se + (k) * QL_LANES + lane : ge + (size_t)(k) * batch))
    if (SMEM) for (int k = 0; k < n; ++k) { D_(k) = u.dg[(size_t)k * batch + inst]; E_(k) = u.eg[(size_t)k * batch + inst]; }
    double2* rot = u.rot + (size_t)inst * u.rot_cap;
    int2* swp = u.swp + (size_t)inst * u.swp_cap;
    long long nrot = 0;
    int nsw = 0;
    bool over = false;
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n && !over; ++l) {
        int iter = 0;
        while (true) {
            if (iter == 0) { const double t = fabs(D_(l)) + fabs(E_(l)); if (t > tst1) tst1 = t; }
            int m = l;
            while (m < n - 1) { if (fabs(E_(m)) <= eps * tst1) break; ++m; }
            if (m == l || iter >= 60) { D_(l) += f; E_(l) = 0.0; break; }
            if (nsw >= u.swp_cap || nrot + (m - l) > u.rot_cap) { over = true; break; }
            swp[nsw++] = make_int2(l, m);
            double g = D_(l);
            const double el = E_(l);
            double pp = (D_(l + 1) - g) / (2.0 * el);
            double r = sqrt(pp * pp + 1.0);
            if (pp < 0) r = -r;
            D_(l) = el / (pp + r);
            D_(l + 1) = el * (pp + r);
            const double dl1 = D_(l + 1);
            double h = g - D_(l);
            {
                int i = l + 2;
                for (; i + 7 < n; i += 8) {                       // loads first, stores after: eight in flight (HBM variant)
                    double t8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t8[j] = D_(i + j);
#pragma unroll
                    for (int j = 0; j < 8; ++j) D_(i + j) = t8[j] - h;
                }
                for (; i < n; ++i) D_(i) -= h;
            }
            f += h;
            pp = D_(m);
            double c = 1.0, c2 = c, c3 = c, s = 0.0, s2 = 0.0;
            const double el1 = E_(l + 1);
            double ei = E_(m - 1), di = D_(m - 1);           // operands of the next rotation, fetched one ahead
            for (int i = m - 1; i >= l; --i) {
                const double e_i = ei, d_i = di;
                if (i > l) { ei = E_(i - 1); di = D_(i - 1); }
                c3 = c2; c2 = c; s2 = s;
                g = c * e_i;
                h = c * pp;
                // The rotation chain is the critical path of the whole kernel (one thread per instance, nothing to overlap
                // with): pp -> rr -> 1/sqrt(rr) -> pp'.  u is off the chain (it only needs the previous c), so the new pp is ONE
                // multiply after the reciprocal square root, and that is a 23-bit hardware seed + one cubically convergent step.
                const double rr = fma(pp, pp, e_i * e_i);
                const double u = fma(pp, d_i, -(e_i * g));          // = r * (c' d_i - s' g)
                const double rinv = ql_rsqrt(rr);
                E_(i + 1) = s * (rr * rinv);
                s = e_i * rinv;
                c = pp * rinv;
                pp = u * rinv;
                D_(i + 1) = h + s * (c * g + s * d_i);
                rot[nrot++] = make_double2(c, s);
            }
            pp = -s * s2 * c3 * el1 * E_(l) / dl1;
            E_(l) = s * pp;
            D_(l) = c * pp;
            ++iter;
        }
    }
    if (over) { u.nswp[inst] = -1; return; }
    if (SMEM) for (int k = 0; k < n; ++k) u.dg[(size_t)k * batch + inst] = D_(k);
    u.nswp[inst] = nsw;
#undef D_
#undef E_
}

// replay the rotation log of the QL kernel on A = Q^T -> Z^T: thread r owns component r of every eigenvector; the log
// is staged through shared memory in pieces (stage: 2 * STAGE doubles).
constexpr int ROT_STAGE = 256;
__device__ void ql_replay(double* A, const int lds, const int n, const double2* __restrict__ rot, const int2* __restrict__ swp,
                          const int nsw, double2* stage) {
    const int tid = threadIdx.x;
    long long base = 0;
    for (int sidx = 0; sidx < nsw; ++sidx) {
        const int2 lm = swp[sidx];
        const int l = lm.x, m = lm.y, cnt = m - l;
        double fz = (tid < n) ? A[(size_t)m * lds + tid] : 0.0;
        for (int done = 0; done < cnt; done += ROT_STAGE) {
            const int piece = (cnt - done < ROT_STAGE) ? cnt - done : ROT_STAGE;
            __syncthreads();                               // previous piece fully consumed
            for (int q = tid; q < piece; q += UKF_THREADS) stage[q] = rot[base + done + q];
            __syncthreads();
            if (tid < n) {
                int i = m - 1 - done;
                double zi = A[(size_t)i * lds + tid];
                for (int q = 0; q < piece; ++q, --i) {
                    const double2 cs = stage[q];
                    const double zn = (q + 1 < piece) ? A[(size_t)(i - 1) * lds + tid] : 0.0;   // next row, fetched ahead
                    A[(size_t)(i + 1) * lds + tid] = cs.y * zi + cs.x * fz;
                    fz = cs.x * zi - cs.y * fz;
                    zi = zn;
                }
            }
        }
        if (tid < n) A[(size_t)l * lds + tid] = fz;
        base += cnt;
    }
    __syncthreads();
}

// rows r0.. of S = Z sqrt(D+) Z^T:  out[q][c] = sum_k Zt[k][rows[q]] * sq[k] * Zt[k][c]   (thread per column c)
template <int NR>
__device__ __forceinline__ void s_rows(const double* Zt, const int lds, const int n, const double* sq,
                                       const int (&rows)[NR], double* out, const int ldo) {
    for (int c = threadIdx.x; c < n; c += UKF_THREADS) {
        double acc[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) acc[q] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double* zr = Zt + (size_t)k * lds;
            const double zc = zr[c] * sq[k];
#pragma unroll
            for (int q = 0; q < NR; ++q) acc[q] += zr[rows[q]] * zc;
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) out[q * ldo + c] = acc[q];
    }
}

// out[q][b] = (S vin[q])_b for NV vectors:  t[q][k] = sq[k] * sum_i Zt[k][i] vin[q][i]  (warp per k), then
// out[q][b] = sum_k Zt[k][b] t[q][k]  (thread per b).  t is scratch [NV][ldv].  One barrier inside, one after.
template <int NV>
__device__ __forceinline__ void s_times(const double* Zt, const int lds, const int n, const double* sq,
                                        const double* vin, double* t, double* out, const int ldv) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < n; k += UKF_WARPS) {
        const double* zr = Zt + (size_t)k * lds;
        double acc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double z = zr[i];
#pragma unroll
            for (int q = 0; q < NV; ++q) acc[q] += z * vin[q * ldv + i];
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double a = acc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) t[q * ldv + k] = a * sq[k];
        }
    }
    __syncthreads();
    for (int bcol = threadIdx.x; bcol < n; bcol += UKF_THREADS) {
        double acc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double z = Zt[(size_t)k * lds + bcol];
#pragma unroll
            for (int q = 0; q < NV; ++q) acc[q] += z * t[q * ldv + k];
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) out[q * ldv + bcol] = acc[q];
    }
    __syncthreads();
}

// ---- launch 1 of 3: Y = scale * sym(P) (ukf.cpp:112-114), landmark block of P_pred seeded with 2 w Y, Householder
//      tridiagonalisation + explicit Q^T -> HBM scratch
__global__ void __launch_bounds__(UKF_THREADS, 2)
ukf_front_kernel(BatchState b, FilterConst fc, StepInputs in, UkfScratch u, const int rescue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UkfSmem s;
    ukf_smem_carve(b, smem_raw, &s);
    const int tid = threadIdx.x;
    const int inst = blockIdx.x;
    const int lds = b.lds;
    const int ldp = b.fixed_ld;                    // global leading dimension of P (fixed for the UKF)
    const int4 meta_in = b.meta[inst];
    if (meta_in.y & SLAM_STATUS_SAME_STEP_REMATCH) return;
    if (rescue && u.nswp[inst] != -1) return;      // rescue pass: only instances whose rotation log overflowed
    if (fc.loc) return;                            // (n = 4 never overflows the log; generation 1 has no map branch)
    const int M = meta_in.x;
    const int n = 4 + 2 * M;                       // ukf.cpp:167 (state size of this step's sigma points)
    double* gP = b.P + (size_t)inst * b.p_stride;
    // P -> A (row loads are coalesced)
    for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
        const int i = idx / n, j = idx - i * n;
        s.A[(size_t)i * lds + j] = gP[(size_t)i * ldp + j];
    }
    __syncthreads();
    const float W0f = 0.2f;                                                // filter.h:207
    const double wgt = (double)((1 - W0f) / (2 * n));                      // :175
    const double scale = (double)((2 * M + 4) / (1 - W0f));                // :114
    // ---- nearestSPD input: Y = 0.5 (P + P^T) * scale (:112-114); and P_LL <- 2 w Y (landmark block of P_pred
    //      before the clipped-eigenvalue correction)
    for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
        const int i = idx / n, j = idx - i * n;
        if (j >= i) {
            const double y = (0.5 * (s.A[(size_t)i * lds + j] + s.A[(size_t)j * lds + i])) * scale;
            s.A[(size_t)i * lds + j] = y; s.A[(size_t)j * lds + i] = y;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
        const int i = idx / n, j = idx - i * n;
        if (i >= 4 && j >= 4) gP[(size_t)i * ldp + j] = (2.0 * wgt) * s.A[(size_t)i * lds + j];
    }
    __syncthreads();
    // ---- first half of the eigendecomposition (:116-118)
    tridiag<UKF_THREADS>(s.A, lds, n, s.d, s.e, s.pool, s.Xp, s.red);       // Xp ([4][2 n_max + 2]) is free in this launch
    form_qt(s.A, lds, n, s.pool, s.Xp);
    double* Zg = u.Zg + (size_t)inst * u.n_max * u.n_max;
    for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
        const int i = idx / n, j = idx - i * n;
        Zg[idx] = s.A[(size_t)i * lds + j];
    }
    for (int k = tid; k < n; k += UKF_THREADS) { u.dg[(size_t)k * b.batch + inst] = s.d[k]; u.eg[(size_t)k * b.batch + inst] = s.e[k]; }
}

// ---- launch 3 of 3: rotation replay -> Z^T, then the sigma-point algebra, updates, insertions, commit
__global__ void __launch_bounds__(UKF_THREADS, 2)
ukf_back_kernel(BatchState b, FilterConst fc, StepInputs in, UkfScratch u, const int rescue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UkfSmem s;
    ukf_smem_carve(b, smem_raw, &s);
    const int tid = threadIdx.x, lane = tid & 31;
    const int inst = blockIdx.x;
    const int lds = b.lds;
    const int ldp = b.fixed_ld;                    // global leading dimension of P (fixed for the UKF)
    const int nmp = ldg_of(b.n_max), nsm = 2 * b.n_max + 2;

    const int4 meta_in = b.meta[inst];
    int nm = in.n_meas[inst];
    int status = meta_in.y;
    if (status & SLAM_STATUS_SAME_STEP_REMATCH) return;
    if (rescue && u.nswp[inst] != -1) return;      // rescue pass: only instances whose rotation log overflowed
    if (fc.loc) return;
    int M = meta_in.x;
    const int M_start = M;
    const int n = 4 + 2 * M;                       // ukf.cpp:167 (state size of this step's sigma points)
    const int ns = 2 * n + 1;
    if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
    double* gP = b.P + (size_t)inst * b.p_stride;
    double* gx = b.x + (size_t)inst * b.x_stride;

    for (int i = tid; i < n; i += UKF_THREADS) { const double v = gx[i]; s.x[i] = v; }
    for (int i = tid; i < M; i += UKF_THREADS) s.ids[i] = b.ids[(size_t)inst * b.max_lm + i];
    for (int i = tid; i < 3 * nm; i += UKF_THREADS) s.meas[i] = in.meas[(size_t)inst * b.max_meas * 3 + i];
    if (tid == 0) { s.iscr[0] = INT_MAX; s.iscr[1] = INT_MAX; s.iscr[2] = 0; }
    {
        const double* Zg = u.Zg + (size_t)inst * u.n_max * u.n_max;
        for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
            const int i = idx / n, j = idx - i * n;
            s.A[(size_t)i * lds + j] = Zg[idx];
        }
        for (int k = tid; k < n; k += UKF_THREADS) { s.d[k] = u.dg[(size_t)k * b.batch + inst]; s.e[k] = u.eg[(size_t)k * b.batch + inst]; }
    }
    __syncthreads();

    // weights and scale are float-valued (ukf.cpp:35,114,175)
    const float W0f = 0.2f;                                                // filter.h:207
    const double W0 = (double)W0f;
    const double wgt = (double)((1 - W0f) / (2 * n));                      // :175
    const double sw = W0 + (double)(2 * n) * wgt;                          // sum of the 2n+1 weights (not 1)
    const float u_d = in.fwd[in.cmd_stride ? inst : 0], u_th = in.ang[in.cmd_stride ? inst : 0];
    const float yaw_prior = yaw_of(s.x[2], s.x[3]);                        // :182 and :139 (prior x_t)
    const double cy = (double)cos_f(yaw_prior), sy = (double)sin_f(yaw_prior);
    const double Qd[4] = {fc.V00 * cy, fc.V00 * sy, fc.V11 * cy, fc.V11 * sy};   // :183-186

    // ---- second half of the eigendecomposition (:116-118) -> s.d eigenvalues, s.A = Z^T
    {
        const int nsw = u.nswp[inst];
        if (nsw >= 0) ql_replay(s.A, lds, n, u.rot + (size_t)inst * u.rot_cap, u.swp + (size_t)inst * u.swp_cap, nsw,
                                reinterpret_cast<double2*>(s.pool));
        else ql_serial(s.A, lds, n, s.d, s.e, s.pool, s.red);
    }

    // ---- clip (:120) and sqrt; list of clipped eigenpairs
    if (tid == 0) {
        int nc = 0;
        for (int k = 0; k < n; ++k) {
            double dk = s.d[k];
            if (dk < 0.00000001) { s.clip[nc] = k; s.corr[nc] = 0.00000001 - dk; ++nc; dk = 0.00000001; }
            s.d[k] = sqrt(dk);
        }
        s.iscr[3] = nc;
    }
    __syncthreads();
    const double* Zt = s.A;
    const double* sq = s.d;
    const int nclip = s.iscr[3];

    // ---- sigma points, vehicle rows (:214-226): rows 0..3 of S, motion model per sigma point
    {
        const int rows4[4] = {0, 1, 2, 3};
        s_rows<4>(Zt, lds, n, sq, rows4, s.S4, nmp);
    }
    __syncthreads();
    for (int i = tid; i < ns; i += UKF_THREADS) {
        double X[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double xr = s.x[r];
            X[r] = (i == 0) ? xr : (i <= n ? xr + s.S4[r * nmp + (i - 1)] : xr - s.S4[r * nmp + (i - 1 - n)]);
        }
        const float yaw = yaw_of(X[2], X[3]);                              // :128
        const float ud = u_d + fc.v_d;
        s.Xp[0 * nsm + i] = X[0] + (double)(ud * cos_f(yaw));              // :129 float product
        s.Xp[1 * nsm + i] = X[1] + (double)(ud * sin_f(yaw));              // :130
        const float fsum = yaw + u_th + fc.v_th;
        const float new_yaw = (float)remainder((double)fsum, TWO_PI_REF);  // :131
        s.Xp[2 * nsm + i] = (double)cos_f(new_yaw);                        // :132
        s.Xp[3 * nsm + i] = (double)sin_f(new_yaw);                        // :133
    }
    __syncthreads();
    // ---- mean (:228-232): vehicle rows by reduction, landmark rows analytically (sum w) * x
    {
        double acc[4] = {0, 0, 0, 0};
        for (int i = tid; i < ns; i += UKF_THREADS) {
            const double wi = (i == 0) ? W0 : wgt;
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] += wi * s.Xp[r * nsm + i];
        }
        block_sum<4>(acc, s.red);
        if (tid < 4) s.xp[tid] = acc[tid];
        for (int r = 4 + tid; r < n; r += UKF_THREADS) s.xp[r] = sw * s.x[r];
    }
    __syncthreads();
    // ---- covariance (:235-240)
    //  vehicle block: sum_i w_i dv_i dv_i^T + Q ; mv[a] = sum_i w_i dv_i[a]
    double mv[4];
    {
        double acc[14] = {0};
        for (int i = tid; i < ns; i += UKF_THREADS) {
            const double wi = (i == 0) ? W0 : wgt;
            double dv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) dv[r] = s.Xp[r * nsm + i] - s.xp[r];
            int q = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = a; c < 4; ++c) acc[q++] += (wi * dv[a]) * dv[c];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[10 + a] += wi * dv[a];
        }
        block_sum<14>(acc, s.red);
        if (tid == 0) {
            int q = 0;
            for (int a = 0; a < 4; ++a)
                for (int c = a; c < 4; ++c) {
                    const double val = acc[q++] + ((a == c) ? Qd[a] : 0.0);
                    gP[(size_t)a * ldp + c] = val; gP[(size_t)c * ldp + a] = val;
                }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) mv[a] = acc[10 + a];
    }
    //  cross block: P[a][b] = w (S g_a)_b + e_b mv[a],  g_a[i] = Xp[a][1+i] - Xp[a][1+n+i],  e_b = x[b] - xp[b]
    {
        double* g = s.pool;                 // [4][nmp]
        double* t = s.pool + 4 * nmp;       // [4][nmp]
        for (int i = tid; i < n; i += UKF_THREADS)
#pragma unroll
            for (int a = 0; a < 4; ++a) g[a * nmp + i] = s.Xp[a * nsm + 1 + i] - s.Xp[a * nsm + 1 + n + i];
        __syncthreads();
        s_times<4>(Zt, lds, n, sq, g, t, g, nmp);       // g <- S g (out may alias vin: vin is dead after phase 1)
        for (int bcol = 4 + tid; bcol < n; bcol += UKF_THREADS) {
            const double eb = s.x[bcol] - s.xp[bcol];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double val = wgt * g[a * nmp + bcol] + eb * mv[a];
                gP[(size_t)a * ldp + bcol] = val;
                gP[(size_t)bcol * ldp + a] = val;
            }
        }
    }
    //  landmark block: += 2w sum_{clipped k} (1e-8 - d_k) z_k z_k^T + (sum w) e_a e_b
    for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
        const int a = idx / n, c = idx - a * n;
        if (a >= 4 && c >= 4) {
            double add = sw * (s.x[a] - s.xp[a]) * (s.x[c] - s.xp[c]);
            for (int q = 0; q < nclip; ++q) {
                const double* zr = Zt + (size_t)s.clip[q] * lds;
                add += (2.0 * wgt) * s.corr[q] * zr[a] * zr[c];
            }
            gP[(size_t)a * ldp + c] += add;
        }
    }
    __syncthreads();

    // ---- update stage (:243-291): updates first (in message order), insertions afterwards
    int n_upd = 0;
    double* z0 = s.S4;                       // [nsm]  (S4 is dead)
    double* z1 = s.S4 + nsm;                 // [nsm]
    for (int l = 0; l < nm; ++l) {
        const int id = (int)s.meas[3 * l];                                  // :258
        const float r = s.meas[3 * l + 1], bb = s.meas[3 * l + 2];
        int* match = &s.iscr[l & 1];
        int cand = INT_MAX;
        for (int j = tid; j < M_start; j += UKF_THREADS) if (s.ids[j] == id) { cand = j; break; }   // :264-269
        cand = __reduce_min_sync(0xffffffffu, cand);
        if (lane == 0 && cand != INT_MAX) atomicMin(match, cand);
        __syncthreads();
        const int slot = *match;
        if (tid == 0) { s.iscr[(l + 1) & 1] = INT_MAX; s.assoc[l] = (slot == INT_MAX) ? -1 : slot; }
        if (slot == INT_MAX) { __syncthreads(); continue; }                 // new landmark: handled after the loop (:272-274)
        ++n_upd;
        // -------- landmarkUpdate (:293-349)
        const int li = slot * 2 + 4;                                        // :298
        double* SL = s.pool;                 // [2][nmp] rows li, li+1 of S
        {
            const int rows2[2] = {li, li + 1};
            s_rows<2>(Zt, lds, n, sq, rows2, SL, nmp);
        }
        __syncthreads();
        for (int i = tid; i < ns; i += UKF_THREADS) {                       // sensingModel per sigma point (:305-308)
            double lx = s.x[li], ly = s.x[li + 1];
            if (i >= 1 && i <= n) { lx += SL[i - 1]; ly += SL[nmp + i - 1]; }
            else if (i > n) { lx -= SL[i - 1 - n]; ly -= SL[nmp + i - 1 - n]; }
            const double dx = lx - s.Xp[0 * nsm + i], dy = ly - s.Xp[1 * nsm + i];
            z0[i] = sqrt(dx * dx + dy * dy) + (double)fc.w_r;               // :144
            z1[i] = remainder(atan2(dy, dx) - (double)yaw_prior + (double)fc.w_b, TWO_PI_REF);   // :145,156
        }
        __syncthreads();
        double zest0;
        {
            double acc[1] = {0.0};
            for (int i = tid; i < ns; i += UKF_THREADS) acc[0] += ((i == 0) ? W0 : wgt) * z0[i];   // :312-314
            block_sum<1>(acc, s.red);
            zest0 = acc[0];
        }
        // dz in place; S2, sum_i w_i dz_i, and the vehicle rows of C
        double S2[2][2], sdz[2], Cv[4][2];
        {
            double acc[13] = {0};
            for (int i = tid; i < ns; i += UKF_THREADS) {
                const double wi = (i == 0) ? W0 : wgt;
                const double d0 = z0[i] - zest0;
                const double d1 = remainder(z1[i] - 0.0, TWO_PI_REF);       // z_est(1) is never accumulated (:310-314,321)
                z0[i] = d0; z1[i] = d1;
                acc[0] += (wi * d0) * d0; acc[1] += (wi * d0) * d1; acc[2] += (wi * d1) * d1;
                acc[3] += wi * d0; acc[4] += wi * d1;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const double wd = wi * (s.Xp[a * nsm + i] - s.xp[a]);
                    acc[5 + 2 * a] += wd * d0; acc[6 + 2 * a] += wd * d1;
                }
            }
            block_sum<13>(acc, s.red);
            S2[0][0] = acc[0] + fc.W00; S2[0][1] = acc[1]; S2[1][0] = acc[1]; S2[1][1] = acc[2] + fc.W11;   // :326
            sdz[0] = acc[3]; sdz[1] = acc[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { Cv[a][0] = acc[5 + 2 * a]; Cv[a][1] = acc[6 + 2 * a]; }
        }
        // landmark rows of C: f_a * sdz + w * S (dz_i - dz_{i+n})
        double* hv = s.pool + 2 * nmp;       // [2][nmp]
        double* tv = s.pool + 4 * nmp;       // [2][nmp]
        for (int i = tid; i < n; i += UKF_THREADS) {
            hv[i] = z0[1 + i] - z0[1 + n + i];
            hv[nmp + i] = z1[1 + i] - z1[1 + n + i];
        }
        __syncthreads();
        s_times<2>(Zt, lds, n, sq, hv, tv, hv, nmp);
        // K = C S2^-1 (:339, partial-pivot LU like Eigen's dynamic inverse), x_pred += K innovation (:342-345)
        double i00, i01, i10, i11;
        {
            const bool swp = fabs(S2[1][0]) > fabs(S2[0][0]);
            const double a00 = swp ? S2[1][0] : S2[0][0], a01 = swp ? S2[1][1] : S2[0][1];
            const double a10 = swp ? S2[0][0] : S2[1][0], a11 = swp ? S2[0][1] : S2[1][1];
            const double l10 = a10 / a00, u11 = a11 - l10 * a01;
            const double b0c0 = swp ? 0.0 : 1.0, b1c0 = swp ? 1.0 : 0.0, b0c1 = swp ? 1.0 : 0.0, b1c1 = swp ? 0.0 : 1.0;
            double y1 = b1c0 - l10 * b0c0; i10 = y1 / u11; i00 = (b0c0 - a01 * i10) / a00;
            y1 = b1c1 - l10 * b0c1; i11 = y1 / u11; i01 = (b0c1 - a01 * i11) / a00;
        }
        const double in0 = (double)r - zest0;
        const double in1 = remainder((double)bb - 0.0, TWO_PI_REF);         // :344
        for (int a = tid; a < n; a += UKF_THREADS) {
            double c0, c1;
            if (a < 4) { c0 = Cv[a][0]; c1 = Cv[a][1]; }
            else {
                const double fa = s.x[a] - s.xp[a];
                c0 = fa * sdz[0] + wgt * hv[a];
                c1 = fa * sdz[1] + wgt * hv[nmp + a];
            }
            const double k0 = c0 * i00 + c1 * i10, k1 = c0 * i01 + c1 * i11;
            s.K[2 * a] = k0; s.K[2 * a + 1] = k1;
            s.xp[a] = s.xp[a] + (k0 * in0 + k1 * in1);
        }
        __syncthreads();
        // P_pred -= (K S2) K^T (:348) on the L2-resident covariance
        for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
            const int i = idx / n, j = idx - i * n;
            const double ki0 = s.K[2 * i], ki1 = s.K[2 * i + 1];
            const double ks0 = ki0 * S2[0][0] + ki1 * S2[1][0], ks1 = ki0 * S2[0][1] + ki1 * S2[1][1];
            gP[(size_t)i * ldp + j] -= ks0 * s.K[2 * j] + ks1 * s.K[2 * j + 1];
        }
        __syncthreads();
    }
    // -------- landmarkInsertion for the unmatched measurements, in message order (:278-287,351-371)
    for (int l = 0; l < nm; ++l) {
        if (s.assoc[l] != -1) continue;
        if (M >= b.max_lm) { status |= SLAM_STATUS_CAPACITY; continue; }
        const int nn = 4 + 2 * M;
        const float r = s.meas[3 * l + 1], bb = s.meas[3 * l + 2];
        if (tid == 0) {
            const float yaw = yaw_of(s.xp[2], s.xp[3]);                     // :356
            const float yb = yaw + bb;
            s.xp[nn] = s.xp[0] + (double)(r * cos_f(yb));                   // :358
            s.xp[nn + 1] = s.xp[1] + (double)(r * sin_f(yb));               // :359
            s.ids[M] = (int)s.meas[3 * l];                                  // :361
        }
        for (int i = tid; i < nn + 2; i += UKF_THREADS) {                   // :365-368, W block and zero cross terms
            gP[(size_t)i * ldp + nn] = (i == nn) ? fc.W00 : 0.0;
            gP[(size_t)i * ldp + nn + 1] = (i == nn + 1) ? fc.W11 : 0.0;
            gP[(size_t)nn * ldp + i] = (i == nn) ? fc.W00 : 0.0;
            gP[(size_t)(nn + 1) * ldp + i] = (i == nn + 1) ? fc.W11 : 0.0;
        }
        M += 1;
        __syncthreads();
    }

    // ---- commit (:289-290)
    const int n_out = 4 + 2 * M;
    for (int i = tid; i < n_out; i += UKF_THREADS) {
        const double v = s.xp[i];
        gx[i] = v;
        if (!isfinite(v)) s.iscr[2] = 1;
    }
    for (int i = tid + M_start; i < M; i += UKF_THREADS) b.ids[(size_t)inst * b.max_lm + i] = s.ids[i];
    for (int i = tid; i < nm; i += UKF_THREADS) b.assoc[(size_t)inst * b.max_meas + i] = s.assoc[i];
    __syncthreads();
    // what slam_get_sigma_points needs of this step (ukf.cpp:214-220): the prior x_t and the factor of S = Z sqrt(D+) Z^T
    {
        double* Zg = u.Zg + (size_t)inst * u.n_max * u.n_max;
        for (int idx = tid; idx < n * n; idx += UKF_THREADS) {
            const int i = idx / n, j = idx - i * n;
            Zg[idx] = s.A[(size_t)i * lds + j];
        }
        for (int k = tid; k < n; k += UKF_THREADS) {
            u.dg[(size_t)k * b.batch + inst] = s.d[k];
            u.xprior[(size_t)inst * u.n_max + k] = s.x[k];
        }
    }
    if (tid == 0) {
        if (s.iscr[2]) status |= SLAM_STATUS_NAN;
        b.meta[inst] = make_int4(M, status, meta_in.z + 1, nm);             // timestep, :164
        u.sigfmt[inst] = make_int2(1, n);
        atomicAdd(u.routes + 2, 1ull);
        if (M > M_start) atomicMax(b.max_M, M);
        double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
        const double nd = (double)n;
        st[8] += 16.0 * nd * nd + 16.0 * nd + 12.0 * nm + 8.0;
        st[9] += 9.0 * nd * nd * nd + 2.0 * nd * nd * nd + 2.0 * nd * nd * (2.0 * nd + 1.0) + 12.0 * n_upd * nd * nd;
        st[10] += nd;
        st[11] += (double)nm;
        // generation 1 moves P (front: read + seed write; back: corrections, updates), Q^T out and in; executes the explicit
        // eigenvector route: 4/3 n^3 (tridiagonal) + 4/3 n^3 (Q) + ~5 n^3 (rotations on n components) + the S-products
        st[12] += 8.0 * nd * nd * (6.0 + 2.0 * n_upd);
        st[13] += (8.0 / 3.0 + 5.0) * nd * nd * nd + 4.0 * nd * nd * (8.0 + 4.0 * n_upd);
    }
}


// =========================================================================================================
// Generation 2 of the step (default): the eigenvector matrix Z = Q V is never formed.
//
// Everything the sigma-point algebra needs from S = Z sqrt(D+) Z^T is "S times a handful of vectors": the unit vectors
// e_0..e_3 and e_li, e_li+1 of the landmarks being updated (rows of S), then g_0..g_3 and the (dz_i - dz_{i+n}) vectors
// of the updates, plus the few eigenvectors z_k = Q V e_k whose eigenvalue was clipped (<= 4 in practice: the
// negative Q entries of ukf.cpp:183-186 only make vehicle directions indefinite).  With Y = Q T Q^T (Householder
// reflectors) and T = V D V^T (the QL kernel's plane-rotation log)
//        S v = Q V sqrt(D+) V^T Q^T v
// is evaluated by pushing the vectors through the reflectors, the rotation log, the scaling, and back: O(n^2) per
// vector instead of the 4/3 n^3 (explicit Q) + ~4.7 n^3 (rotations applied to n eigenvector components) of the first
// generation.  One WARP owns an instance; lane j carries vector j in a padded shared-memory tile W[n][33] (the odd
// pitch makes both the lane = vector walks of the transforms and the lane = component walks of the sigma-point
// algebra conflict free).  The k landmark updates of a step only couple through the running x_pred (K and S2 come
// from the step's sigma points, ukf.cpp:305-339), so their S-products share the second pass, and P_pred is written
// exactly once:  P = [vehicle | cross | 2w Y + (sum w) e e^T + clipped-eigenpair terms] - sum_q (K_q S2_q) K_q^T.
//   ukf_front2_kernel  CTA per instance: Y, tridiagonalisation -> HBM scratch (reflectors + tau, 2wY, d, e); P untouched
//   ukf_ql_kernel      thread per instance (shared with generation 1)
//   ukf_back2_kernel   warp per instance: the two S-passes and the algebra
// An instance whose rotation log overflowed (nswp = -1; never observed: the log holds 2 n^2 rotations, 2.4x the
// typical count) is left untouched by ukf_back2_kernel and redone by the generation-1 kernels in rescue mode.
// =========================================================================================================
// pitch of the vector tile (doubles): odd, >= the lanes a step can need (4 + 2 max_meas, + 4 clipped eigenvectors)
__host__ __device__ inline int ukf_wcols(const BatchState& b) { const int c = 4 + 2 * b.max_meas + 4; return c > 32 ? 32 : c; }
__host__ __device__ inline int ukf_wld(const BatchState& b) { const int c = ukf_wcols(b); return c <= 12 ? 13 : (c <= 24 ? 25 : 33); }   // compile-time variants
// The typical step needs far fewer lanes (4 + 2 * 1.5 updates + <= 4 clipped), so the batch first goes through a NARROW
// tile (pitch 13: up to 4 updates; 27 KB of shared memory per instance -> 8 instances in flight per SM instead of 6); an
// instance with more updates in this step flags itself and is taken by a second launch with the full-width tile.
constexpr int UKF_NARROW_WLD = 13;
constexpr int UPD_LD = 24;               // per-update scalars
constexpr int UKF2_MAX_UPD = 14;         // 4 + 2 * updates <= 32 lanes

struct UkfWarpSmem {
    double* W;      // [2 + n_max + 2][wld]   lane j = vector j; two spare rows on either side (prefetch overrun)
    double* x;      // prior x_t
    double* xp;     // running x_pred
    double* sq;     // sqrt(max(d, 1e-8))
    double* Xp;     // [2][nsm] propagated vehicle rows x, y of the sigma points
    float* Xcs;     // [2][nsm] rows cos, sin: float VALUES in the reference (ukf.cpp:132-133), stored as such (lossless)
    double* upd;    // [max_meas][UPD_LD]
    double* veh;    // [24] multi-warp back kernel: xp0v[4], mv[4], vv[10] (vehicle mean / covariance terms) shared by the warps
    double2* stage; // [2 + 64 + 2] rotation-log ring (spare entries for the prefetch overrun)
    double* corr;   // [32] 1e-8 - d_k of the clipped eigenvalues
    int* clip;      // [32] their indices
    int* ids;
    float* meas;
    int* assoc;
    int* uq;        // measurement index of update q
    int* ctl;       // [8] multi-warp back kernel: nu, nclip, status bits shared by warp 0
};

__host__ __device__ inline size_t ukf_warp_carve(const BatchState& b, const int wld, unsigned char* base, UkfWarpSmem* s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const int nmp = ldg_of(b.n_max), nsm = 2 * b.n_max + 2;
    size_t oW = take(sizeof(double) * (size_t)(b.n_max + 4) * wld);
    size_t ox = take(sizeof(double) * nmp), oxp = take(sizeof(double) * nmp), osq = take(sizeof(double) * nmp);
    size_t oXp = take(sizeof(double) * 2 * nsm), oXcs = take(sizeof(float) * 2 * nsm);
    int nupd = (wld - 5) / 2; if (nupd > b.max_meas) nupd = b.max_meas; if (nupd < 1) nupd = 1;     // updates a tile of this pitch can hold
    size_t oupd = take(sizeof(double) * UPD_LD * nupd);
    size_t oveh = take(sizeof(double) * 24);
    size_t ostage = take(sizeof(double2) * 68);
    size_t ocorr = take(sizeof(double) * 32), oclip = take(sizeof(int) * 32);
    size_t oids = take(sizeof(int) * (b.max_lm + 1));
    size_t omeas = take(sizeof(float) * 3 * (b.max_meas > 0 ? b.max_meas : 1));
    size_t oassoc = take(sizeof(int) * (b.max_meas > 0 ? b.max_meas : 1));
    size_t ouq = take(sizeof(int) * (b.max_meas > 0 ? b.max_meas : 1));
    size_t octl = take(sizeof(int) * 8);
    if (s) {
        s->W = (double*)(base + oW) + 2 * wld; s->x = (double*)(base + ox); s->xp = (double*)(base + oxp); s->sq = (double*)(base + osq);
        s->Xp = (double*)(base + oXp); s->Xcs = (float*)(base + oXcs); s->upd = (double*)(base + oupd); s->veh = (double*)(base + oveh);
        s->stage = (double2*)(base + ostage) + 2; s->corr = (double*)(base + ocorr); s->clip = (int*)(base + oclip);
        s->ids = (int*)(base + oids); s->meas = (float*)(base + omeas); s->assoc = (int*)(base + oassoc); s->uq = (int*)(base + ouq); s->ctl = (int*)(base + octl);
    }
    return off;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// w <- H_k w over all reflectors, k ascending (Q^T w) or descending (Q w), for the vectors in columns [0, nv).
// Row k of R: tau_k at column k and the reflector below it (v_k[k+1] = 1 stored).  lane = COMPONENT here: a lane owns
// the rows i = lane (mod 32) of the tile for the whole phase (no synchronisation between reflectors), the reflector is
// one coalesced row load fetched one reflector ahead, and the nv dot products are reduced by shuffles four at a time.
template <bool ASC, int NT, int WLD>
__device__ __forceinline__ void apply_reflectors(double* W, const int n, const int lane, const int nv,
                                                 const double* __restrict__ R, const int j_lo = 0, const int j_step = 4) {
    double va[NT], taua = 0.0;
    auto loadk = [&](const int k, double (&dst)[NT], double& tau) {
        if (k < 0 || k > n - 2) { tau = 0.0; return; }
        const double* __restrict__ row = R + (size_t)k * n;
        tau = __ldg(row + k);
#pragma unroll
        for (int t = 0; t < NT; ++t) { const int i = lane + 32 * t; dst[t] = (i > k && i < n) ? __ldg(row + i) : 0.0; }
    };
    const int step = ASC ? 1 : -1;
    int k = ASC ? 0 : n - 2;
    loadk(k, va, taua);
    if (j_lo < nv && j_lo + j_step >= nv) {
        // ONE group of four vectors for this warp (always so in the multi-warp kernel): the lane's rows of the group stay in
        // registers over all n - 1 reflectors -- per reflector 16 + 16 FMA and the shuffles, no shared-memory traffic at all
        // (the tile was read and written back for every reflector: half of this phase's instructions)
        double w[NT][4];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int i = lane + 32 * t;
            const double* wr = W + i * WLD + j_lo;
#pragma unroll
            for (int c = 0; c < 4; ++c) w[t][c] = (i < n) ? wr[c] : 0.0;
        }
        for (int kk = 0; kk < n - 1; ++kk, k += step) {
            double v[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) v[t] = va[t];
            const double tau = taua;
            loadk(k + step, va, taua);
            if (tau == 0.0) continue;
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) { p0 += v[t] * w[t][0]; p1 += v[t] * w[t][1]; p2 += v[t] * w[t][2]; p3 += v[t] * w[t][3]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                p2 += __shfl_xor_sync(0xffffffffu, p2, o); p3 += __shfl_xor_sync(0xffffffffu, p3, o);
            }
            p0 *= tau; p1 *= tau; p2 *= tau; p3 *= tau;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (v[t] != 0.0) { w[t][0] -= p0 * v[t]; w[t][1] -= p1 * v[t]; w[t][2] -= p2 * v[t]; w[t][3] -= p3 * v[t]; }
            }
        }
        const int left = nv - j_lo;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int i = lane + 32 * t;
            if (i < n) {
                double* wr = W + i * WLD + j_lo;
                wr[0] = w[t][0];
                if (left > 1) wr[1] = w[t][1];
                if (left > 2) wr[2] = w[t][2];
                if (left > 3) wr[3] = w[t][3];
            }
        }
        __syncwarp();
        return;
    }
    for (int kk = 0; kk < n - 1; ++kk, k += step) {
        double v[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) v[t] = va[t];
        const double tau = taua;
        loadk(k + step, va, taua);
        if (tau == 0.0) continue;
        for (int j0 = j_lo; j0 < nv; j0 += j_step) {         // (a warp of the multi-warp kernel owns the groups j_lo, j_lo + j_step, ..)
            double w[NT][4];
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int i = lane + 32 * t;
                if (i < n) {
                    const double* wr = W + i * WLD + j0;
                    w[t][0] = wr[0]; w[t][1] = wr[1]; w[t][2] = wr[2]; w[t][3] = wr[3];
                    p0 += v[t] * w[t][0]; p1 += v[t] * w[t][1]; p2 += v[t] * w[t][2]; p3 += v[t] * w[t][3];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                p2 += __shfl_xor_sync(0xffffffffu, p2, o); p3 += __shfl_xor_sync(0xffffffffu, p3, o);
            }
            p0 *= tau; p1 *= tau; p2 *= tau; p3 *= tau;
            const int left = nv - j0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int i = lane + 32 * t;
                if (i < n && v[t] != 0.0) {
                    double* wr = W + i * WLD + j0;
                    wr[0] = w[t][0] - p0 * v[t];
                    if (left > 1) wr[1] = w[t][1] - p1 * v[t];
                    if (left > 2) wr[2] = w[t][2] - p2 * v[t];
                    if (left > 3) wr[3] = w[t][3] - p3 * v[t];
                }
            }
        }
    }
    __syncwarp();
}

// chunk c (32 rotations) of an instance's log, one entry per lane; identity beyond either end
__device__ __forceinline__ double2 rot_chunk(const double2* __restrict__ rot, const int c, const int lane, const int nrot) {
    const int ix = c * 32 + lane;
    return (c >= 0 && ix < nrot) ? rot[ix] : make_double2(1.0, 0.0);
}
__device__ __forceinline__ int2 swp_at(const int2* __restrict__ swp, const int i, const int nsw) {
    return (i >= 0 && i < nsw) ? swp[i] : make_int2(0, 0);
}

// u <- V^T u: the QL kernel's rotation log replayed in generation order on the vector components (lane = VECTOR; the
// (c, s) stream is staged through a 2 x 32 shared-memory ring; the next chunk and the next sweep range are
// fetched while the current ones are consumed)
template <int WLD>
__device__ __forceinline__ void apply_rot_fwd(double* W, const int lane, const bool act, const double2* __restrict__ rot,
                                              const int2* __restrict__ swp, const int nsw, const int nrot, double2* stage) {
    int g = 0;
    double2 q0 = rot_chunk(rot, 0, lane, nrot);
    int2 lm0 = swp_at(swp, 0, nsw);
    double* w = W + lane;
    for (int sidx = 0; sidx < nsw; ++sidx) {
        const int2 lm = lm0;
        lm0 = swp_at(swp, sidx + 1, nsw);
        const int l = lm.x, m = lm.y;
        double fz = 0.0, zi = 0.0;
        if (act) { fz = w[m * WLD]; zi = w[(m - 1) * WLD]; }
        int i = m - 1;
        while (i >= l) {
            if ((g & 31) == 0) {
                const int c = g >> 5;
                __syncwarp();
                stage[(c & 1) * 32 + lane] = q0;
                q0 = rot_chunk(rot, c + 1, lane, nrot);
                __syncwarp();
            }
            int cnt = 32 - (g & 31);
            if (cnt > i - l + 1) cnt = i - l + 1;
            if (act) {
                const double2* __restrict__ st = stage + (g & 63);
                double* wi = w + i * WLD;
                double2 c0 = st[0], c1 = st[1];
                double z1 = wi[-WLD];                               // row i - 1 (a spare row when i = 0)
#pragma unroll 4
                for (int q = 0; q < cnt; ++q) {
                    const double2 c2 = st[q + 2];                   // operands two rotations ahead
                    const double z2 = wi[-2 * WLD];
                    const double t = c0.x * zi;
                    wi[WLD] = fma(c0.x, fz, c0.y * zi);
                    fz = fma(-c0.y, fz, t);
                    zi = z1; z1 = z2; c0 = c1; c1 = c2;
                    wi -= WLD;
                }
            }
            i -= cnt; g += cnt;
        }
        if (act) w[l * WLD] = fz;
    }
    __syncwarp();
}

// u <- V u: the log replayed backwards, every plane rotation inverted (each 2x2 step [[s,c],[c,-s]] is its own inverse)
template <int WLD>
__device__ __forceinline__ void apply_rot_bwd(double* W, const int lane, const bool act, const double2* __restrict__ rot,
                                              const int2* __restrict__ swp, const int nsw, const int nrot, double2* stage) {
    if (nrot <= 0) return;
    int g = nrot - 1;
    const int cl = g >> 5;
    {
        const double2 cur = rot_chunk(rot, cl, lane, nrot);
        __syncwarp();
        stage[(cl & 1) * 32 + lane] = cur;
        __syncwarp();
    }
    double2 q0 = rot_chunk(rot, cl - 1, lane, nrot);
    int2 lm0 = swp_at(swp, nsw - 1, nsw);
    double* w = W + lane;
    for (int sidx = nsw - 1; sidx >= 0; --sidx) {
        const int2 lm = lm0;
        lm0 = swp_at(swp, sidx - 1, nsw);
        const int l = lm.x, m = lm.y;
        double fz = 0.0, a = 0.0;
        if (act) { fz = w[l * WLD]; a = w[(l + 1) * WLD]; }
        int i = l;
        while (i < m) {
            int cnt = (g & 31) + 1;                              // rotations left in the staged chunk (going down)
            if (cnt > m - i) cnt = m - i;
            if (act) {
                const double2* __restrict__ st = stage + (g & 63);
                double* wi = w + i * WLD;
                double2 c0 = st[0], c1 = st[-1];
                double a1 = wi[2 * WLD];                            // row i + 2 (a spare row beyond n - 1)
#pragma unroll 4
                for (int q = 0; q < cnt; ++q) {
                    const double2 c2 = st[-(q + 2)];                // operands two rotations ahead
                    const double a2 = wi[3 * WLD];
                    const double t = c0.x * a;
                    wi[0] = fma(c0.x, fz, c0.y * a);
                    fz = fma(-c0.y, fz, t);
                    a = a1; a1 = a2; c0 = c1; c1 = c2;
                    wi += WLD;
                }
            }
            i += cnt; g -= cnt;
            if ((g & 31) == 31 && g > 0) {                       // the next rotation opens chunk g >> 5
                const int c = g >> 5;
                __syncwarp();
                stage[(c & 1) * 32 + lane] = q0;
                q0 = rot_chunk(rot, c - 1, lane, nrot);
                __syncwarp();
            }
        }
        if (act) w[m * WLD] = fz;
    }
    __syncwarp();
}

// ---- launch 1 of 3 (generation 2): Y = scale * sym(P) (ukf.cpp:112-114), tridiagonalisation; P is NOT modified
constexpr int FRONT2_THREADS = 256;      // (512 threads measured slower: 3.79 -> 4.11 ms; the Householder steps are barrier bound)
// Shared memory of the kernel, sized by a CLASS of instance sizes (ncap = largest n of the class), not by the batch's capacity:
// the Householder steps are barrier and latency bound, so the instances with few landmarks run at 4 CTAs per SM instead of 2.
struct Front2Smem { double *A, *d, *e, *pool, *part, *red; };
__host__ __device__ inline size_t front2_carve(const int ncap, unsigned char* base, Front2Smem* s, const bool packed = true) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const size_t oA = take(sizeof(double) * (packed ? (size_t)ncap * (ncap + 1) / 2 : (size_t)ncap * lds_of(ncap)));
    const size_t od = take(sizeof(double) * ncap), oe = take(sizeof(double) * ncap);
    const size_t opool = take(sizeof(double) * 4 * ncap), opart = take(sizeof(double) * ((packed ? TRIP_GMAX : 8) * ncap + 8)), ored = take(sizeof(double) * 32);
    if (s) {
        s->A = (double*)(base + oA); s->d = (double*)(base + od); s->e = (double*)(base + oe);
        s->pool = (double*)(base + opool); s->part = (double*)(base + opart); s->red = (double*)(base + ored);
    }
    return off;
}
// classes for MINB = 4, 3, 2 CTAs per SM (register budget 64 / 80 / 128 per thread); caps ascending, returns the count
inline int front2_classes(const BatchState& b, int caps[3], int minb[3], const bool packed) {
    static const int per_sm[3] = {4, 3, 2};
    int nc = 0, prev = 0;
    for (int k = 0; k < 3; ++k) {
        const size_t budget = (size_t)(227 * 1024) / per_sm[k] - 1024;
        int cap = prev;
        while (cap < b.n_max && front2_carve(cap + 1, nullptr, nullptr, packed) <= budget) ++cap;
        if (k == 2) cap = b.n_max;                 // the last class takes whatever is left (one CTA per SM beyond 2 x 113 KB)
        if (cap > prev) { caps[nc] = cap; minb[nc] = per_sm[k]; ++nc; prev = cap; }
        if (cap >= b.n_max) break;
    }
    return nc;
}

template <int MINB, bool PACKED>
__global__ void __launch_bounds__(FRONT2_THREADS, MINB)
ukf_front2_kernel(BatchState b, UkfScratch u, const int i0, const int nlo, const int ncap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Front2Smem s;
    front2_carve(ncap, smem_raw, &s, PACKED);
    const int tid = threadIdx.x;
    const int inst = i0 + blockIdx.x;
    const int lds = lds_of(ncap);
    const int ldp = b.fixed_ld;
    const int4 meta_in = b.meta[inst];
    if (meta_in.y & SLAM_STATUS_SAME_STEP_REMATCH) return;
    const int M = meta_in.x;
    const int n = 4 + 2 * M;                       // ukf.cpp:167
    if (n <= nlo || n > ncap) return;              // another class's launch takes this instance
    const double* gP = b.P + (size_t)inst * b.p_stride;
    const float W0f = 0.2f;                                                // filter.h:207
    const double wgt = (double)((1 - W0f) / (2 * n));                      // :175
    const double scale = (double)((2 * M + 4) / (1 - W0f));                // :114
    double* Yg = u.Yg + (size_t)inst * u.n_max * u.n_max;
    double* Rg = u.Zg + (size_t)inst * u.n_max * u.n_max;
    const double* tau = s.pool + 3 * n;
    if constexpr (PACKED) {
        // Y = scale * sym(P), lower triangle only: (i, j <= i) from P[j][i] + P[i][j] (the mirror element straight from global memory)
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int i = idx / n, j = idx - i * n;
            if (j <= i) s.A[tri_off(i) + j] = (0.5 * (gP[(size_t)j * ldp + i] + gP[(size_t)i * ldp + j])) * scale;
        }
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int i = idx / n, j = idx - i * n;
            if (i >= 4 && j >= 4) Yg[idx] = (2.0 * wgt) * s.A[i >= j ? tri_off(i) + j : tri_off(j) + i];   // landmark block of P_pred before corrections
        }
        __syncthreads();
        tridiag_packed<FRONT2_THREADS>(s.A, n, s.d, s.e, s.pool, s.part, s.red);
        // reflector k (column k of A below the diagonal) -> row k of the scratch matrix, tau_k on its diagonal
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int k = idx / n, i = idx - k * n;
            if (i > k) Rg[idx] = s.A[tri_off(i) + k];
            else if (i == k) Rg[idx] = (k < n - 1) ? tau[k] : 0.0;
        }
    } else {
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int i = idx / n, j = idx - i * n;
            s.A[(size_t)i * lds + j] = gP[(size_t)i * ldp + j];
        }
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int i = idx / n, j = idx - i * n;
            if (j >= i) {
                const double y = (0.5 * (s.A[(size_t)i * lds + j] + s.A[(size_t)j * lds + i])) * scale;
                s.A[(size_t)i * lds + j] = y; s.A[(size_t)j * lds + i] = y;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int i = idx / n, j = idx - i * n;
            if (i >= 4 && j >= 4) Yg[idx] = (2.0 * wgt) * s.A[(size_t)i * lds + j];   // landmark block of P_pred before corrections
        }
        __syncthreads();
        tridiag<FRONT2_THREADS>(s.A, lds, n, s.d, s.e, s.pool, s.part, s.red);
        // reflector k (column k of A below the diagonal) -> row k of the scratch matrix, tau_k on its diagonal
        for (int idx = tid; idx < n * n; idx += FRONT2_THREADS) {
            const int k = idx / n, i = idx - k * n;
            if (i > k) Rg[idx] = s.A[(size_t)i * lds + k];
            else if (i == k) Rg[idx] = (k < n - 1) ? tau[k] : 0.0;
        }
    }
    for (int k = tid; k < n; k += FRONT2_THREADS) { u.dg[(size_t)k * b.batch + inst] = s.d[k]; u.eg[(size_t)k * b.batch + inst] = s.e[k]; }
}

// =========================================================================================================
// Generation 3 (default): the tridiagonal eigenproblem T = V D V^T is solved in parallel INSIDE an instance (eig3.cuh:
// thread t <-> eigenpair t; bisection on Sturm counts, twisted-factorisation inverse iteration, Gram-Schmidt only inside
// clusters of close eigenvalues) and V is written out explicitly, so that the back kernel evaluates
//      S v = Q V sqrt(D+) V^T Q^T v
// with two dense n x n products per S-pass instead of replaying ~0.85 n^2 plane rotations serially on every vector, and the
// one-thread-per-instance QL chain disappears from the step.  An instance the solver declines (cluster larger than `maxc`,
// residual above tolerance, non-finite pivots) is flagged nswp = -2 and takes the generation-2 route (QL kernel + rotation
// replay) in the same step; nswp = -3 marks "explicit V in the scratch".
//   scratch:  Vg  [batch][n^2]  V[i][k]  = component i of eigenvector k of T (rows contiguous over k)
//             VTg [batch][n^2]  VT[k][i] = the transpose (rows contiguous over i); doubles as the solver's per-thread work space
// =========================================================================================================
struct Eig3Smem {
    double *d, *e, *e2, *lam, *lt, *red, *gm;
    eig3::De* de;       // {d[i], e[i-1]^2} packed for the Sturm recurrence
    int *b0, *b1, *crank, *cfirst, *tw, *flag, *gs;     // (lt, tw, gm, gs): the multisection grid -- point, count, p_n = gm 2^gs
    double* tile;       // shared-memory variant: n x (n_max | 1) work tile, column t = thread t's vector
};
__host__ __device__ inline size_t eig3_carve(const int n_max, const int nt, unsigned char* base, Eig3Smem* s, const bool tile = false) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const size_t od = take(8 * n_max), oe = take(8 * n_max), oe2 = take(8 * n_max), ol = take(8 * n_max), olt = take(8 * n_max), ored = take(8 * 40);
    const size_t ode = take(16 * n_max), ogm = take(8 * n_max), ogs = take(4 * n_max);
    const size_t ob0 = take(4 * n_max), ob1 = take(4 * n_max), ocr = take(4 * n_max), ocf = take(4 * n_max), otw = take(4 * n_max), ofl = take(4 * 8);
    const size_t oti = tile ? take((size_t)8 * n_max * (n_max | 1)) : 0;
    if (s) {
        s->tile = (double*)(base + oti);
        s->d = (double*)(base + od); s->e = (double*)(base + oe); s->e2 = (double*)(base + oe2); s->lam = (double*)(base + ol);
        s->lt = (double*)(base + olt); s->red = (double*)(base + ored); s->de = (eig3::De*)(base + ode);
        s->b0 = (int*)(base + ob0); s->b1 = (int*)(base + ob1); s->crank = (int*)(base + ocr); s->cfirst = (int*)(base + ocf);
        s->tw = (int*)(base + otw); s->flag = (int*)(base + ofl); s->gm = (double*)(base + ogm); s->gs = (int*)(base + ogs);
    }
    return off;
}

// Size classes of the shared-memory variant: the tile (and with it the CTAs per SM) is sized by the class, not by the batch's
// capacity, so an instance with few landmarks runs at 4 CTAs per SM and only the largest ones at 2.  One launch per class; a CTA
// whose instance belongs to another class exits at once.  caps[k] = largest n of class k (ascending), returns the class count;
// 0 = the variant is not used (n_max beyond what fits twice per SM, or 128 threads).
constexpr int EIG3_MAX_CLASSES = 6;
inline int eig3_classes(const BatchState& b, int caps[EIG3_MAX_CLASSES], int* per = nullptr, int* nts = nullptr) {
    if (b.n_max > 128) return 0;
    // up to n = 64 a CTA of 64 threads (two warps, both live) at 8 / 6 / 5 CTAs per SM; beyond, 128 threads at 4 / 3 / 2
    // (<= 128 registers per thread)
    static const int per_sm[6] = {8, 6, 5, 4, 3, 2}, nt_of[6] = {64, 64, 64, 128, 128, 128};
    int nc = 0, prev = 0;
    for (int k = 0; k < 6; ++k) {
        const size_t budget = (size_t)(227 * 1024) / per_sm[k] - 1024;
        const int lim = (nt_of[k] == 64) ? (b.n_max < 64 ? b.n_max : 64) : b.n_max;
        int cap = prev;
        while (cap < lim && eig3_carve(cap + 1, nt_of[k], nullptr, nullptr, true) <= budget) ++cap;
        if (cap > prev) { if (per) per[nc] = per_sm[k]; if (nts) nts[nc] = nt_of[k]; caps[nc++] = cap; prev = cap; }
        if (cap >= b.n_max) return nc;
    }
    return 0;                                      // the largest instances would not fit twice per SM
}

template <int NT>
__device__ __forceinline__ double cta_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) m = fmax(m, red[w]);
    return m;
}

// ---- launch 2 of 3 (generation 3): CTA per instance, thread t <-> eigenpair t of the tridiagonal matrix
// modified Gram-Schmidt inside the clusters of an instance, vectors = columns of `base` (leading dimension ld; global scratch or
// the shared-memory tile).  Every WARP takes whole clusters, round robin, and walks the members of a cluster in order (lanes
// stride over the components: dot products and updates are 32 wide; __syncwarp orders a member's stores before the next
// member's loads).  flag[2]: a projection removed a sizeable part of a vector (first pass); flag[1]: still dependent (second).
__device__ __forceinline__ void eig3_gram_schmidt(const Eig3Smem& s, double* base, const int ld, const int n, const int warp, const int lane,
                                                  const int nwarps, const int pass) {
    int ord = 0;
    for (int m0 = 0; m0 + 1 < n; ++m0) {
        if (s.crank[m0] != 0 || s.crank[m0 + 1] == 0) continue;          // m0: first index of a cluster
        if ((ord++ % nwarps) != warp) continue;
        const int ma = s.b0[m0], mb = s.b1[m0];
        for (int m = m0 + 1; m < mb && s.crank[m] > 0; ++m) {
            double* zm = base + m;
            for (int q = m0; q < m; ++q) {
                const double* zq = base + q;
                double dot = 0.0;
                for (int i = ma + lane; i < mb; i += 32) dot += zq[(size_t)i * ld] * zm[(size_t)i * ld];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                for (int i = ma + lane; i < mb; i += 32) zm[(size_t)i * ld] -= dot * zq[(size_t)i * ld];
            }
            double n2 = 0.0;
            for (int i = ma + lane; i < mb; i += 32) { const double v = zm[(size_t)i * ld]; n2 += v * v; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
            // unit vector before: a small remainder means nearly dependent vectors.  In the first pass the twisted vectors of a
            // pathologically close pair may coincide (the refinement step separates them); after it, decline the instance
            if (pass == 0 && !(n2 > eig3::REFINE_BELOW) && lane == 0) s.flag[2] = 1;
            if (pass == 1 && !(n2 > 1.0e-6) && lane == 0) s.flag[1] = 1;
            const double sc = rsqrt(n2);
            for (int i = ma + lane; i < mb; i += 32) zm[(size_t)i * ld] *= sc;
            __syncwarp();
        }
    }
}

// TILE = true (default when the tile fits twice per SM): the eigenvectors are built in a shared-memory tile, column t = thread t.
// The pivot and vector recurrences are chains of dependent divisions that read back what they wrote a moment ago; in the global
// scratch (TILE = false) every such read-back is an L2 round trip inside the chain (stores do not allocate in L1), which made
// the kernel wait on memory, not on arithmetic.  From the tile, V and V^T go out once, coalesced.  An instance that needs the
// refinement step (nearly parallel twisted vectors; the tile keeps no factors) is flagged nswp = -4 and re-done by the
// TILE = false kernel, launched behind this one with only_flagged = 1.
template <int NT, bool TILE>
__global__ void __launch_bounds__(NT, TILE ? 512 / NT : 512 / NT)       // 128 registers: room for the batches of pivots fetched ahead of the chains
ukf_eig3_kernel(BatchState b, UkfScratch u, const int i0, const int maxc, const int only_flagged, const int nlo, const int ncap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Eig3Smem s;
    eig3_carve(ncap, NT, smem_raw, &s, TILE);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int inst = i0 + blockIdx.x;
    const int4 meta_in = b.meta[inst];
    if (meta_in.y & SLAM_STATUS_SAME_STEP_REMATCH) return;
    const int n = 4 + 2 * meta_in.x;
    if (TILE && (n <= nlo || n > ncap)) return;    // size classes: this launch's tile holds nlo < n <= ncap (see eig3_classes)
    if (!TILE && only_flagged && u.nswp[inst] != -4 && n <= nlo) return;     // behind the tile launches: the instances they handed over
                                                                             // and the sizes (n > nlo) no tile class was launched for
    double* const Vg = u.Vg + (size_t)inst * u.n_max * u.n_max;       // [i][k], compact leading dimension n
    double* const Wg = u.VTg + (size_t)inst * u.n_max * u.n_max;      // work space now, V^T at the end
    const bool live = t < n;
    if (live) { s.d[t] = u.dg[(size_t)t * b.batch + inst]; s.e[t] = (t < n - 1) ? u.eg[(size_t)t * b.batch + inst] : 0.0; }
    if (t == 0) { s.flag[0] = 0; s.flag[1] = 0; s.flag[2] = 0; }
    __syncthreads();
    // |T| (largest absolute row sum), then the splits (eig3.cuh, step 1)
    double rs = 0.0;
    if (live) rs = fabs(s.d[t]) + (t > 0 ? fabs(s.e[t - 1]) : 0.0) + fabs(s.e[t]);
    const double tn = cta_max<NT>(rs, s.red);
    if (live && t < n - 1 && fabs(s.e[t]) <= eig3::EPS * (fabs(s.d[t]) + fabs(s.d[t + 1]))) s.e[t] = 0.0;
    __syncthreads();
    double e2v = 0.0;
    if (live) { e2v = s.e[t] * s.e[t]; s.e2[t] = e2v; s.de[t].d = s.d[t]; if (t + 1 < n) s.de[t + 1].e2 = e2v; if (t == 0) s.de[0].e2 = 0.0; }
    const double e2max = cta_max<NT>(e2v, s.red);
    const double pivmin = fmax(2.2250738585072014e-308 * fmax(1.0, e2max), 1e-300);
    const double pivf = eig3::EPS * tn;
    int a = 0, bb = 0;
    double lam = 0.0, glo = 0.0, ghi = 0.0;
    if (live) {
        a = t; while (a > 0 && s.e[a - 1] != 0.0) --a;
        bb = t + 1; while (bb < n && s.e[bb - 1] != 0.0) ++bb;
        s.b0[t] = a; s.b1[t] = bb;
        if (bb - a > 1) {
            // multisection start: one Sturm count per thread on a uniform grid over the block's Gershgorin interval, shared below
            eig3::block_bounds(s.d, s.e, a, bb, pivmin, glo, ghi);
            const double x0 = eig3::grid_point(glo, ghi, t - a, bb - a);
            eig3::Fval f0;
            s.lt[t] = x0;
            s.tw[t] = eig3::sturm_eval(s.de, a, bb, x0, &f0);
            s.gm[t] = f0.m; s.gs[t] = f0.s;
        }
    }
    __syncthreads();
    if (live) {
        if (bb - a == 1) lam = s.d[a];
        else {
            double lo = glo, hi = ghi;
            int qlo, qhi;
            eig3::bracket_from_grid2(s.lt, s.tw, a, bb, t - a, lo, hi, qlo, qhi);
#ifdef EIG3_PURE_BISECTION
            lam = eig3::bisect_bracket(s.de, a, bb, t - a, lo, hi, 2.0 * eig3::EPS * tn + 2.0 * pivmin);
#else
            // bisection until the eigenvalue is alone in its bracket, then secant steps on p_n (eig3.cuh: root_bracket)
            eig3::End el{lo, 0, eig3::Fval{0.0, 0}, qlo >= 0}, eh{hi, bb - a, eig3::Fval{0.0, 0}, qhi >= 0};
            if (qlo >= 0) { el.c = s.tw[qlo]; el.f.m = s.gm[qlo]; el.f.s = s.gs[qlo]; }
            if (qhi >= 0) { eh.c = s.tw[qhi]; eh.f.m = s.gm[qhi]; eh.f.s = s.gs[qhi]; }
            lam = eig3::root_bracket(s.de, a, bb, t - a, el, eh, 2.0 * eig3::EPS * tn + 2.0 * pivmin, nullptr);
#endif
        }
        s.lam[t] = lam;
    }
    __syncthreads();
    // clusters inside a block (neighbours closer than 1e-3 |T|, LAPACK dstein) and dstein's perturbation of coincident shifts
    int crank = 0, cfirst = t;
    double x = lam;
    if (live) {
        const double ortol = 1.0e-3 * tn;
        while (t - crank > a && s.lam[t - crank] - s.lam[t - crank - 1] < ortol && crank <= maxc) ++crank;
        cfirst = t - crank;
        if (crank >= maxc && crank > 0) s.flag[0] = 1;               // cluster too large for the in-kernel Gram-Schmidt
        x = s.lam[cfirst];
        for (int q = cfirst + 1; q <= t; ++q) {
            const double lq = s.lam[q], pert = 10.0 * eig3::EPS * fabs(lq);
            x = (lq - x < pert) ? x + pert : lq;
        }
        s.crank[t] = crank; s.cfirst[t] = cfirst;
    }
    const int maxrank = (int)cta_max<NT>((double)crank, s.red);
    if (s.flag[0]) { if (t == 0) u.nswp[inst] = -2; return; }
    if constexpr (TILE) {
        const int ldw = ncap | 1;                  // odd: rows (consecutive t) and columns (consecutive i) are both conflict free
        const eig3::Slot zt{s.tile + t, ldw};
        bool badt = false;
        if (live) {
            for (int i = 0; i < a; ++i) zt.set(i, 0.0);
            for (int i = bb; i < n; ++i) zt.set(i, 0.0);
            if (bb - a == 1) zt.set(a, 1.0);
            else {
                double n2 = 1.0;
#ifdef EIG3_SECOND_SWEEP
                eig3::twisted_vector1(s.d, s.e, s.e2, a, bb, x, pivf, zt, &n2);
#else
                eig3::twisted_vector1g(s.d, s.e, s.e2, a, bb, x, pivf, zt, eig3::Slot{Wg + t, n}, &n2);   // D- parked in the (still unused) V^T region
#endif
                const double sc = rsqrt(n2);
                if (!(n2 > 0.0) || !isfinite(n2)) badt = true;
#pragma unroll 4
                for (int i = a; i < bb; ++i) zt.set(i, zt.get(i) * sc);
            }
        }
        if (maxrank > 0) {
            __syncthreads();
            eig3_gram_schmidt(s, s.tile, ldw, n, warp, lane, NT / 32, 0);
        }
        __syncthreads();
        if (s.flag[2] || (u.refine_all && maxrank > 0)) { if (t == 0) u.nswp[inst] = -4; return; }     // needs the refinement step: the two-array kernel re-does it
        if (live && bb - a > 1) {
            const double r = eig3::residual_inf(s.d, s.e, a, bb, lam, zt, 1.0);
            if (!(r <= 1.0e-12 * tn)) badt = true;
        }
        if (badt) s.flag[1] = 1;
        __syncthreads();
        if (s.flag[1]) { if (t == 0) u.nswp[inst] = -2; return; }
        if (live) {
#pragma unroll 4
            for (int i = 0; i < n; ++i) Vg[(size_t)i * n + t] = s.tile[i * ldw + t];          // V[i][k = t]
#pragma unroll 4
            for (int k = 0; k < n; ++k) Wg[(size_t)k * n + t] = s.tile[t * ldw + k];          // V^T[k][i = t] = V[i = t][k]
            u.dg[(size_t)t * b.batch + inst] = lam;
        }
        if (t == 0) u.nswp[inst] = -3;
        return;
    }
    // eigenvectors: twisted factorisation, z in column t of V, factors in column t of the work space
    const eig3::Slot z{Vg + t, n}, w{Wg + t, n};
    int tw = a;
    bool bad = false;
    if (live) {
        for (int i = 0; i < a; ++i) z.set(i, 0.0);
        for (int i = bb; i < n; ++i) z.set(i, 0.0);
        if (bb - a == 1) z.set(a, 1.0);
        else {
            double n2 = 1.0;
            tw = eig3::twisted_vector_pf<8>(s.d, s.e, s.e2, a, bb, x, pivf, z, w, &n2);   // every read-back one batch ahead of its chain
            const double sc = rsqrt(n2);
            if (!(n2 > 0.0) || !isfinite(n2)) bad = true;
#pragma unroll 8
            for (int i = a; i < bb; ++i) z.set(i, z.get(i) * sc);
        }
    }
    // modified Gram-Schmidt inside clusters; then, ONLY IF a projection removed a sizeable part of some vector (nearly parallel
    // twisted vectors: eig3.cuh REFINE_BELOW), one refinement step for the cluster members and Gram-Schmidt again
    for (int pass = 0; pass < 2; ++pass) {
        if (maxrank == 0) break;
        if (pass == 1) {
            if (!s.flag[2]) break;                 // CTA-uniform: written before the barrier that ended pass 0
            const bool incl = live && bb - a > 1 && (crank > 0 || (t + 1 < bb && s.crank[t + 1] > 0));
            if (incl) {
                const double n2 = eig3::twisted_solve(s.e, a, bb, tw, z, w);
                const double sc = rsqrt(n2);
                if (!(n2 > 0.0) || !isfinite(n2)) bad = true;
                for (int i = a; i < bb; ++i) z.set(i, z.get(i) * sc);
            }
        }
        __syncthreads();                           // (global memory: the vectors are visible to the CTA after the barrier)
        eig3_gram_schmidt(s, Vg, n, n, warp, lane, NT / 32, pass);
        __syncthreads();
    }
    // residual |(T - lambda I) z| against |T|
    if (live && bb - a > 1) {
        const double r = eig3::residual_inf_pf<8>(s.d, s.e, a, bb, lam, z);
        if (!(r <= 1.0e-12 * tn)) bad = true;
    }
    if (bad) s.flag[1] = 1;
    __syncthreads();
    if (s.flag[1]) { if (t == 0) u.nswp[inst] = -2; return; }        // dg / eg untouched: the QL route takes the instance
    // V^T: every thread reads its column of V (coalesced over the threads) and writes it as row t of the transpose -- four
    // loads in flight, consecutive addresses per thread (the sectors merge in L2); no shared-memory tile, which would cap the
    // CTAs per SM of this latency-bound kernel.  The work space is dead by now (barrier above).  Eigenvalues over diag(T).
    if (live) {
        double* row = Wg + (size_t)t * n;
        int i = 0;
        for (; i + 7 < n; i += 8) {
            double v8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v8[q] = z.get(i + q);
#pragma unroll
            for (int q = 0; q < 8; ++q) row[i + q] = v8[q];
        }
        for (; i < n; ++i) row[i] = z.get(i);
        u.dg[(size_t)t * b.batch + inst] = lam;
    }
    if (t == 0) u.nswp[inst] = -3;
}

// W[:, j] <- rs .* (Mat^T W[:, j]) for the columns j < ncols of the warp's vector tile (written back only for j < jw):
// out[c][j] = rs[c] * sum_r Mat[r][c] W[r][j], Mat row-major with leading dimension n.  Lane <-> output rows c = lane + 32 t
// (coalesced row loads of Mat), the inputs W[r][*] are shared-memory broadcasts, JB columns accumulate in registers.
template <int NTT, int WLD>
__device__ __forceinline__ void dense_apply(double* W, const double* __restrict__ Mat, const int n, const int lane, const int ncols,
                                            const int jw, const double* rs) {
    constexpr int JB = (NTT <= 4) ? 12 : 6;
    for (int j0 = 0; j0 < ncols; j0 += JB) {
        double acc[NTT][JB];
#pragma unroll
        for (int q = 0; q < NTT; ++q)
#pragma unroll
            for (int j = 0; j < JB; ++j) acc[q][j] = 0.0;
#pragma unroll 2
        for (int r = 0; r < n; ++r) {
            double m[NTT];
#pragma unroll
            for (int q = 0; q < NTT; ++q) { const int c = lane + 32 * q; m[q] = (c < n) ? __ldg(Mat + (size_t)r * n + c) : 0.0; }
            const double* wr = W + r * WLD + j0;
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const double wv = wr[j];
#pragma unroll
                for (int q = 0; q < NTT; ++q) acc[q][j] = fma(m[q], wv, acc[q][j]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NTT; ++q) {
            const int c = lane + 32 * q;
            if (c < n) {
                const double sc = rs ? rs[c] : 1.0;
#pragma unroll
                for (int j = 0; j < JB; ++j) if (j0 + j < jw) W[c * WLD + j0 + j] = acc[q][j] * sc;
            }
        }
        __syncwarp();
    }
}

// ---- launch 3 of 3 (generations 2 and 3): warp per instance.  DENSE = false: S-products through the QL rotation log
// (generation 2, and the fall-back of generation 3); DENSE = true: through the explicit eigenvectors of T (generation 3).
template <int wld, bool DENSE>
__global__ void __launch_bounds__(32, 10)        // ten one-warp CTAs per SM: <= 200 registers per thread
ukf_back2_kernel(BatchState b, FilterConst fc, StepInputs in, UkfScratch u, const int i0, const int pass) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UkfWarpSmem s;
    ukf_warp_carve(b, wld, smem_raw, &s);
    const int lane = threadIdx.x;
    const int inst = i0 + blockIdx.x;
    const int ldp = b.fixed_ld;
    const int nsm = 2 * b.n_max + 2;
    const unsigned FULL = 0xffffffffu;
    const int wcols = wld - 1 < ukf_wcols(b) ? wld - 1 : ukf_wcols(b);
    double* const W_ = s.W;
    const bool small_n = b.n_max <= 128;

    const int4 meta_in = b.meta[inst];
    int nm = in.n_meas[inst];
    int status = meta_in.y;
    if (status & SLAM_STATUS_SAME_STEP_REMATCH) return;
    const int nsw = u.nswp[inst];
    if (DENSE ? (nsw != -3) : (nsw < 0)) return;   // not this route: rotation log overflow (-1, rescue pass), declined by the
                                                   // parallel eigensolver (-2, QL route) or solved by it (-3, dense route)
    if (pass == 1 && !u.defer[inst]) return;       // full-width pass: only the instances the narrow pass handed over
    int M = meta_in.x;
    const int M_start = M;
    const int n = 4 + 2 * M;                       // ukf.cpp:167
    const int ns = 2 * n + 1;
    if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
    double* gP = b.P + (size_t)inst * b.p_stride;
    double* gx = b.x + (size_t)inst * b.x_stride;
    const double* __restrict__ R = u.Zg + (size_t)inst * u.n_max * u.n_max;
    double* Yg = u.Yg + (size_t)inst * u.n_max * u.n_max;
    const double2* __restrict__ rot = u.rot + (size_t)inst * u.rot_cap;
    const int2* __restrict__ swp = u.swp + (size_t)inst * u.swp_cap;
    const double* __restrict__ Vm = u.Vg + (size_t)inst * u.n_max * u.n_max;      // generation 3: V[i][k] and its transpose
    const double* __restrict__ VTm = u.VTg + (size_t)inst * u.n_max * u.n_max;

    for (int i = lane; i < n; i += 32) { const double v = gx[i]; s.x[i] = v; u.xprior[(size_t)inst * u.n_max + i] = v; }
    for (int i = lane; i < M; i += 32) s.ids[i] = b.ids[(size_t)inst * b.max_lm + i];
    for (int i = lane; i < 3 * nm; i += 32) s.meas[i] = in.meas[(size_t)inst * b.max_meas * 3 + i];
    // eigenvalues: clip (:120), sqrt, ordered list of the clipped ones; total length of the rotation log
    int nclip = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        const double dk = (k < n) ? u.dg[(size_t)k * b.batch + inst] : 1.0;
        const bool cl = (k < n) && (dk < 0.00000001);
        const unsigned mask = __ballot_sync(FULL, cl);
        const int pos = nclip + __popc(mask & ((1u << lane) - 1u));
        if (cl && pos < 32) { s.clip[pos] = k; s.corr[pos] = 0.00000001 - dk; }
        if (k < n) s.sq[k] = sqrt(cl ? 0.00000001 : dk);
        nclip += __popc(mask);
    }
    if (nclip > 32) {
        // more clipped directions than the 32 correction slots of this kernel (the reference clips any number, ukf.cpp:120):
        // handled like a rotation-log overflow -- the instance is left untouched and redone by the generation-1 rescue pass
        if (lane == 0) u.nswp[inst] = -1;
        return;
    }
    int nrot = 0;
    if (!DENSE) for (int q = lane; q < nsw; q += 32) { const int2 lm = swp[q]; nrot += lm.y - lm.x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrot += __shfl_xor_sync(FULL, nrot, o);
    __syncwarp();

    // weights and scale are float-valued (ukf.cpp:35,114,175)
    const float W0f = 0.2f;                                                // filter.h:207
    const double W0 = (double)W0f;
    const double wgt = (double)((1 - W0f) / (2 * n));                      // :175
    const double sw = W0 + (double)(2 * n) * wgt;                          // sum of the 2n+1 weights (not 1)
    const float u_d = in.fwd[in.cmd_stride ? inst : 0], u_th = in.ang[in.cmd_stride ? inst : 0];
    const float yaw_prior = yaw_of(s.x[2], s.x[3]);                        // :182 and :139 (prior x_t)
    const double cy = (double)cos_f(yaw_prior), sy = (double)sin_f(yaw_prior);
    const double Qd[4] = {fc.V00 * cy, fc.V00 * sy, fc.V11 * cy, fc.V11 * sy};   // :183-186

    // ---- association of every measurement (:258-274): known-ID lookup against the landmarks of the step start
    int nu = 0;
    for (int l = 0; l < nm; ++l) {
        const int id = (int)s.meas[3 * l];                                  // :258
        int cand = INT_MAX;
        if (fc.loc) {                                                        // :262,272,300-302: every detection updates, by map id
            if (id >= 0 && id < fc.n_map) cand = id; else status |= SLAM_STATUS_BAD_ID;   // the reference reads map[] out of range
        } else {
            for (int j = lane; j < M_start; j += 32) if (s.ids[j] == id) { cand = j; break; }   // :264-269
            cand = __reduce_min_sync(FULL, cand);
        }
        if (lane == 0) {
            s.assoc[l] = (cand == INT_MAX) ? -1 : cand;
            if (cand != INT_MAX) s.uq[nu] = l;
        }
        if (cand != INT_MAX) ++nu;
    }
    __syncwarp();
    if (pass == 0) {                                // narrow tile: hand the instance over if its updates do not fit
        const bool over = 4 + 2 * nu > wcols;
        if (lane == 0) u.defer[inst] = over ? 1 : 0;
        if (over) return;
    }
    const int nvec = 4 + 2 * nu;                    // lanes of the two S-passes
    int ncf = wcols - nvec; if (ncf > nclip) ncf = nclip;   // clipped eigenvectors riding in pass A
    if (u.clip_lanes > 0 && ncf > u.clip_lanes) ncf = u.clip_lanes;        // test knob

    // ---- clipped eigenpairs that do not fit beside pass A (never in practice): z_k = Q V e_k, 32 at a time,
    //      folded into the landmark-block seed  Yg += 2w (1e-8 - d_k) z_k z_k^T
    for (int c0 = ncf; c0 < nclip && c0 < 32; c0 += wcols) {
        int cnt = (nclip < 32 ? nclip : 32) - c0; if (cnt > wcols) cnt = wcols;
        const bool act = lane < cnt;
        if (lane < wcols) for (int i = 0; i < n; ++i) W_[i * wld + lane] = (act && s.clip[c0 + lane] == i) ? 1.0 : 0.0;
        __syncwarp();
        if (DENSE) { if (small_n) dense_apply<4, wld>(W_, VTm, n, lane, cnt, cnt, nullptr); else dense_apply<8, wld>(W_, VTm, n, lane, cnt, cnt, nullptr); }
        else apply_rot_bwd<wld>(W_, lane, act, rot, swp, nsw, nrot, s.stage);
        if (small_n) apply_reflectors<false, 4, wld>(W_, n, lane, cnt, R); else apply_reflectors<false, 8, wld>(W_, n, lane, cnt, R);
        for (int a = 4; a < n; ++a)
            for (int c = 4 + lane; c < n; c += 32) {
                double add = 0.0;
                for (int q = 0; q < cnt; ++q) add += (2.0 * wgt) * s.corr[c0 + q] * W_[a * wld + q] * W_[c * wld + q];
                Yg[(size_t)a * n + c] += add;
            }
        __syncwarp();
    }

    // ---- pass A: S e_r for the vehicle rows and the rows of the landmarks being updated; clipped eigenvectors beside
    {
        int row = -1;
        if (lane < 4) row = lane;
        else if (lane < nvec && !fc.loc) { const int q = (lane - 4) >> 1; row = s.assoc[s.uq[q]] * 2 + 4 + ((lane - 4) & 1); }   // :298
        if (lane < wcols) for (int i = 0; i < n; ++i) W_[i * wld + lane] = (i == row) ? 1.0 : 0.0;
        const bool act = lane < nvec;
        __syncwarp();
        if (small_n) apply_reflectors<true, 4, wld>(W_, n, lane, nvec, R); else apply_reflectors<true, 8, wld>(W_, n, lane, nvec, R);
        const bool isclip = lane >= nvec && lane < nvec + ncf;
        const int ck = isclip ? s.clip[lane - nvec] : -1;
        if (DENSE) {
            // sqrt(D+) V^T (.) for the active columns, unit vectors e_ck in the clipped columns, then V (.) for all of them
            if (small_n) dense_apply<4, wld>(W_, Vm, n, lane, nvec, nvec, s.sq); else dense_apply<8, wld>(W_, Vm, n, lane, nvec, nvec, s.sq);
            if (lane < wcols && !act) for (int i = 0; i < n; ++i) W_[i * wld + lane] = (i == ck) ? 1.0 : 0.0;
            __syncwarp();
            if (small_n) dense_apply<4, wld>(W_, VTm, n, lane, nvec + ncf, nvec + ncf, nullptr); else dense_apply<8, wld>(W_, VTm, n, lane, nvec + ncf, nvec + ncf, nullptr);
        } else {
            apply_rot_fwd<wld>(W_, lane, act, rot, swp, nsw, nrot, s.stage);
            if (lane < wcols) for (int i = 0; i < n; ++i) {
                if (act) W_[i * wld + lane] *= s.sq[i];
                else W_[i * wld + lane] = (i == ck) ? 1.0 : 0.0;
            }
            const bool act2 = act || isclip;
            apply_rot_bwd<wld>(W_, lane, act2, rot, swp, nsw, nrot, s.stage);
        }
        if (small_n) apply_reflectors<false, 4, wld>(W_, n, lane, nvec + ncf, R); else apply_reflectors<false, 8, wld>(W_, n, lane, nvec + ncf, R);
    }

    // ---- sigma points, vehicle rows (:214-226): X = x +- column of S, motion model per sigma point
    for (int i = lane; i < ns; i += 32) {
        double X[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double xr = s.x[r];
            X[r] = (i == 0) ? xr : (i <= n ? xr + W_[(i - 1) * wld + r] : xr - W_[(i - 1 - n) * wld + r]);
        }
        const float yaw = yaw_of(X[2], X[3]);                              // :128
        const float ud = u_d + fc.v_d;
        s.Xp[0 * nsm + i] = X[0] + (double)(ud * cos_f(yaw));              // :129 float product
        s.Xp[1 * nsm + i] = X[1] + (double)(ud * sin_f(yaw));              // :130
        const float fsum = yaw + u_th + fc.v_th;
        const float new_yaw = (float)remainder((double)fsum, TWO_PI_REF);  // :131
        s.Xcs[0 * nsm + i] = cos_f(new_yaw);                               // :132
        s.Xcs[1 * nsm + i] = sin_f(new_yaw);                               // :133
    }
    __syncwarp();
    auto XP = [&](const int r, const int i) -> double { return r < 2 ? s.Xp[r * nsm + i] : (double)s.Xcs[(r - 2) * nsm + i]; };
    // ---- mean (:228-232): vehicle rows by reduction, landmark rows analytically (sum w) * x
    double xp0v[4];
    {
        double acc[4] = {0, 0, 0, 0};
        for (int i = lane; i < ns; i += 32) {
            const double wi = (i == 0) ? W0 : wgt;
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] += wi * XP(r, i);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) xp0v[r] = warp_sum(acc[r]);
        if (lane < 4) s.xp[lane] = xp0v[lane];
        for (int r = 4 + lane; r < n; r += 32) s.xp[r] = sw * s.x[r];
    }
    __syncwarp();
    // ---- covariance (:235-240), vehicle block: sum_i w_i dv_i dv_i^T + Q ; mv[a] = sum_i w_i dv_i[a]
    double mv[4], vv[10];
    {
        double acc[14] = {0};
        for (int i = lane; i < ns; i += 32) {
            const double wi = (i == 0) ? W0 : wgt;
            double dv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) dv[r] = XP(r, i) - xp0v[r];
            int q = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = a; c < 4; ++c) acc[q++] += (wi * dv[a]) * dv[c];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[10 + a] += wi * dv[a];
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) acc[k] = warp_sum(acc[k]);
        {
            int q = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = a; c < 4; ++c) { vv[q] = acc[q] + ((a == c) ? Qd[a] : 0.0); ++q; }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) mv[a] = acc[10 + a];
    }
    // ---- the updates' sigma-point statistics (:293-336), in message order; every update q leaves
    //      hv = dz_i - dz_{i+n} in its two lanes of W (the S rows it read from them are dead by then)
    for (int q = 0; q < nu; ++q) {
        const int l = s.uq[q];
        const int li = fc.loc ? 0 : s.assoc[l] * 2 + 4;                     // :298
        const int c0 = 4 + 2 * q;
        // sensingModel of sigma point i (:305-308); sgn / row select the column of S it was built from (0: the mean point)
        const double mlx = fc.loc ? (double)fc.map[3 * s.assoc[l] + 1] : s.x[li];             // :152-153 (true map, float) / :144
        const double mly = fc.loc ? (double)fc.map[3 * s.assoc[l] + 2] : s.x[li + 1];
        auto sense = [&](const int i, const double sgn, const int row, double& a0, double& a1) {
            double lx = mlx, ly = mly;
            if (!fc.loc && sgn != 0.0) { lx += sgn * W_[row * wld + c0]; ly += sgn * W_[row * wld + c0 + 1]; }
            const double dx = lx - s.Xp[0 * nsm + i], dy = ly - s.Xp[1 * nsm + i];
            a0 = sqrt(dx * dx + dy * dy) + (double)fc.w_r;                  // :144
            a1 = remainder(atan2(dy, dx) - (double)yaw_prior + (double)fc.w_b, TWO_PI_REF);   // :145,156
        };
        // The measurement predictions are not stored (their 3.4 KB per instance is what separated 9 from 10 instances per
        // SM, i.e. four from three rounds of CTAs): a first pass gives the range mean (:312-314), a second one recomputes
        // them pair by pair -- lane i owns sigma points 1 + i and 1 + n + i, so dz_i - dz_{i+n} never leaves the lane.
        double zest0 = 0.0;
        for (int i = lane; i < ns; i += 32) {
            double a0, a1;
            sense(i, i == 0 ? 0.0 : (i <= n ? 1.0 : -1.0), i <= n ? i - 1 : i - 1 - n, a0, a1);
            zest0 += ((i == 0) ? W0 : wgt) * a0;
        }
        zest0 = warp_sum(zest0);
        double acc[13] = {0};
        auto accum = [&](const int i, const double wi, const double a0, const double a1, double& d0, double& d1) {
            d0 = a0 - zest0;
            d1 = remainder(a1 - 0.0, TWO_PI_REF);                           // z_est(1) is never accumulated (:310-314,321)
            acc[0] += (wi * d0) * d0; acc[1] += (wi * d0) * d1; acc[2] += (wi * d1) * d1;
            acc[3] += wi * d0; acc[4] += wi * d1;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double wd = wi * (XP(a, i) - xp0v[a]);                // about the predicted mean; the running
                acc[5 + 2 * a] += wd * d0; acc[6 + 2 * a] += wd * d1;       // x_pred enters below as a rank-1 shift
            }
        };
        if (lane == 0) { double a0, a1, d0, d1; sense(0, 0.0, 0, a0, a1); accum(0, W0, a0, a1, d0, d1); }
        for (int i = lane; i < n; i += 32) {
            double a0, a1, b0, b1, da0, da1, db0, db1;
            sense(1 + i, 1.0, i, a0, a1);
            sense(1 + n + i, -1.0, i, b0, b1);
            accum(1 + i, wgt, a0, a1, da0, da1);
            accum(1 + n + i, wgt, b0, b1, db0, db1);
            if (!fc.loc) {                                                  // hv overwrites the S rows this lane just read
                W_[i * wld + c0] = da0 - db0;
                W_[i * wld + c0 + 1] = da1 - db1;
            }
        }
#pragma unroll
        for (int k = 0; k < 13; ++k) acc[k] = warp_sum(acc[k]);
        __syncwarp();
        if (lane == 0) {
            double* ud = s.upd + q * UPD_LD;
            const double S00 = acc[0] + fc.W00, S01 = acc[1], S11 = acc[2] + fc.W11;     // :326
            ud[0] = S00; ud[1] = S01; ud[2] = S11; ud[3] = acc[3]; ud[4] = acc[4];
            for (int a = 0; a < 8; ++a) ud[5 + a] = acc[5 + a];
            // S2^-1 (:339, partial-pivot LU like Eigen's dynamic inverse)
            const bool swpv = fabs(S01) > fabs(S00);
            const double a00 = swpv ? S01 : S00, a01 = swpv ? S11 : S01;
            const double a10 = swpv ? S00 : S01, a11 = swpv ? S01 : S11;
            const double l10 = a10 / a00, u11 = a11 - l10 * a01;
            const double b0c0 = swpv ? 0.0 : 1.0, b1c0 = swpv ? 1.0 : 0.0, b0c1 = swpv ? 1.0 : 0.0, b1c1 = swpv ? 0.0 : 1.0;
            double y1 = b1c0 - l10 * b0c0;
            const double i10 = y1 / u11, i00 = (b0c0 - a01 * i10) / a00;
            y1 = b1c1 - l10 * b0c1;
            const double i11 = y1 / u11, i01 = (b0c1 - a01 * i11) / a00;
            ud[13] = i00; ud[14] = i01; ud[15] = i10; ud[16] = i11;
            ud[17] = (double)s.meas[3 * l + 1] - zest0;                                   // innovation (:342-344)
            ud[18] = remainder((double)s.meas[3 * l + 2] - 0.0, TWO_PI_REF);
        }
        __syncwarp();
    }
    // g_a[i] = Xp[a][1+i] - Xp[a][1+n+i] -> lanes 0..3
    for (int i = lane; i < n; i += 32)
#pragma unroll
        for (int a = 0; a < 4; ++a) W_[i * wld + a] = XP(a, 1 + i) - XP(a, 1 + n + i);
    __syncwarp();

    // ---- pass B: S g_a and S hv for every update; the clipped eigenvectors stay put in their lanes
    {
        const bool act = lane < nvec;
        if (small_n) apply_reflectors<true, 4, wld>(W_, n, lane, nvec, R); else apply_reflectors<true, 8, wld>(W_, n, lane, nvec, R);
        if (DENSE) {
            if (small_n) { dense_apply<4, wld>(W_, Vm, n, lane, nvec, nvec, s.sq); dense_apply<4, wld>(W_, VTm, n, lane, nvec, nvec, nullptr); }
            else { dense_apply<8, wld>(W_, Vm, n, lane, nvec, nvec, s.sq); dense_apply<8, wld>(W_, VTm, n, lane, nvec, nvec, nullptr); }
        } else {
            apply_rot_fwd<wld>(W_, lane, act, rot, swp, nsw, nrot, s.stage);
            if (act) for (int i = 0; i < n; ++i) W_[i * wld + lane] *= s.sq[i];
            apply_rot_bwd<wld>(W_, lane, act, rot, swp, nsw, nrot, s.stage);
        }
        if (small_n) apply_reflectors<false, 4, wld>(W_, n, lane, nvec, R); else apply_reflectors<false, 8, wld>(W_, n, lane, nvec, R);
    }

    // ---- gains (:336-345) in message order: C uses the RUNNING x_pred; K_q overwrites the update's lanes of W
    for (int q = 0; q < nu; ++q) {
        const double* ud = s.upd + q * UPD_LD;
        const int c0 = 4 + 2 * q;
        const double sdz0 = ud[3], sdz1 = ud[4];
        const double i00 = ud[13], i01 = ud[14], i10 = ud[15], i11 = ud[16], in0 = ud[17], in1 = ud[18];
        for (int a = lane; a < n; a += 32) {
            double c0v, c1v;
            if (a < 4) {
                const double sh = xp0v[a] - s.xp[a];
                c0v = ud[5 + 2 * a] + sh * sdz0; c1v = ud[6 + 2 * a] + sh * sdz1;
            } else {
                const double fa = s.x[a] - s.xp[a];
                c0v = fa * sdz0 + wgt * W_[a * wld + c0];
                c1v = fa * sdz1 + wgt * W_[a * wld + c0 + 1];
            }
            const double k0 = c0v * i00 + c1v * i10, k1 = c0v * i01 + c1v * i11;
            W_[a * wld + c0] = k0; W_[a * wld + c0 + 1] = k1;
            s.xp[a] = s.xp[a] + (k0 * in0 + k1 * in1);
        }
    }
    __syncwarp();

    // ---- P_pred, written once (:235-240 then :348 per update, in message order)
    for (int i = 0; i < n; ++i) {
        const double ei = (i >= 4) ? s.x[i] - sw * s.x[i] : 0.0;
        for (int jb = 0; jb < n; jb += 128) {
            double yv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {                       // the row's seed loads go out together
                const int j = jb + lane + 32 * t;
                yv[t] = (i >= 4 && j >= 4 && j < n) ? Yg[(size_t)i * n + j] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = jb + lane + 32 * t;
                if (j >= n) continue;
                double val;
                if (i < 4 && j < 4) {
                    const int a = i < j ? i : j, c = i < j ? j : i;
                    val = vv[a * 4 - (a * (a - 1)) / 2 + (c - a)];
                } else if (i < 4) {
                    const double eb = s.x[j] - sw * s.x[j];
                    val = wgt * W_[j * wld + i] + eb * mv[i];
                } else if (j < 4) {
                    val = wgt * W_[i * wld + j] + ei * mv[j];
                } else {
                    const double ej = s.x[j] - sw * s.x[j];
                    double add = sw * ei * ej;
                    for (int q = 0; q < ncf; ++q) add += (2.0 * wgt) * s.corr[q] * W_[i * wld + nvec + q] * W_[j * wld + nvec + q];
                    val = yv[t] + add;
                }
                for (int q = 0; q < nu; ++q) {
                    const double* ud = s.upd + q * UPD_LD;
                    const int c0 = 4 + 2 * q;
                    const double ki0 = W_[i * wld + c0], ki1 = W_[i * wld + c0 + 1];
                    const double ks0 = ki0 * ud[0] + ki1 * ud[1], ks1 = ki0 * ud[1] + ki1 * ud[2];
                    val -= ks0 * W_[j * wld + c0] + ks1 * W_[j * wld + c0 + 1];
                }
                gP[(size_t)i * ldp + j] = val;
            }
        }
    }
    __syncwarp();

    // -------- landmarkInsertion for the unmatched measurements, in message order (:278-287,351-371)
    for (int l = 0; l < nm && !fc.loc; ++l) {
        if (s.assoc[l] != -1) continue;
        if (M >= b.max_lm) { status |= SLAM_STATUS_CAPACITY; continue; }
        const int nn = 4 + 2 * M;
        const float r = s.meas[3 * l + 1], bb = s.meas[3 * l + 2];
        if (lane == 0) {
            const float yaw = yaw_of(s.xp[2], s.xp[3]);                     // :356
            const float yb = yaw + bb;
            s.xp[nn] = s.xp[0] + (double)(r * cos_f(yb));                   // :358
            s.xp[nn + 1] = s.xp[1] + (double)(r * sin_f(yb));               // :359
            s.ids[M] = (int)s.meas[3 * l];                                  // :361
        }
        for (int i = lane; i < nn + 2; i += 32) {                           // :365-368, W block and zero cross terms
            gP[(size_t)i * ldp + nn] = (i == nn) ? fc.W00 : 0.0;
            gP[(size_t)i * ldp + nn + 1] = (i == nn + 1) ? fc.W11 : 0.0;
            gP[(size_t)nn * ldp + i] = (i == nn) ? fc.W00 : 0.0;
            gP[(size_t)(nn + 1) * ldp + i] = (i == nn + 1) ? fc.W11 : 0.0;
        }
        M += 1;
        __syncwarp();
    }

    // ---- commit (:289-290)
    const int n_out = 4 + 2 * M;
    bool bad = false;
    for (int i = lane; i < n_out; i += 32) {
        const double v = s.xp[i];
        gx[i] = v;
        if (!isfinite(v)) bad = true;
    }
    bad = __any_sync(FULL, bad);
    for (int i = lane + M_start; i < M; i += 32) b.ids[(size_t)inst * b.max_lm + i] = s.ids[i];
    for (int i = lane; i < nm; i += 32) b.assoc[(size_t)inst * b.max_meas + i] = s.assoc[i];
    if (lane == 0) {
        if (bad) status |= SLAM_STATUS_NAN;
        b.meta[inst] = make_int4(M, status, meta_in.z + 1, nm);             // timestep, :164
        atomicAdd(u.routes + (DENSE ? 0 : 1), 1ull);
        u.sigfmt[inst] = make_int2(DENSE ? 3 : 2, n);                       // reflectors + (rotation log | explicit V) + eigenvalues stay in the scratch
        if (M > M_start) atomicMax(b.max_M, M);
        double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
        const double nd = (double)n;
        st[8] += 16.0 * nd * nd + 16.0 * nd + 12.0 * nm + 8.0;
        st[9] += 9.0 * nd * nd * nd + 2.0 * nd * nd * nd + 2.0 * nd * nd * (2.0 * nd + 1.0) + 12.0 * nu * nd * nd;
        st[10] += nd;
        st[11] += (double)nm;
        // bytes the three launches really move for the instance: P in (front) and out (back), reflectors + seed out and in,
        // the rotation log out (QL) and in four times (two S-passes, forward and backward)
        st[12] += 8.0 * nd * nd * 6.0 + 16.0 * (double)nrot * 5.0;
        // flops they execute: tridiagonalisation 4/3 n^3, QL ~25 per rotation, per vector of the two S-passes 12 per
        // rotation (forward + backward replay) and 4 n^2 (Q^T and Q through the reflectors), P_pred assembly
        st[13] += 4.0 / 3.0 * nd * nd * nd + 25.0 * (double)nrot + (double)(2 * nvec + ncf) * (12.0 * (double)nrot + 4.0 * nd * nd)
                  + nd * nd * (2.0 + 8.0 * nu + 2.0 * ncf);
    }
}

// One group of JB columns (j0 .. j0 + JB) of dense_apply, for the multi-warp back kernel: each warp owns its columns.
template <int NTT, int WLD, int JB>
__device__ __forceinline__ void dense_cols(double* W, const double* __restrict__ Mat, const int n, const int lane, const int j0,
                                           const int jw, const double* rs) {
    double acc[NTT][JB];
#pragma unroll
    for (int q = 0; q < NTT; ++q)
#pragma unroll
        for (int j = 0; j < JB; ++j) acc[q][j] = 0.0;
#pragma unroll 4
    for (int r = 0; r < n; ++r) {
        double m[NTT];
#pragma unroll
        for (int q = 0; q < NTT; ++q) { const int c = lane + 32 * q; m[q] = (c < n) ? __ldg(Mat + (size_t)r * n + c) : 0.0; }
        const double* wr = W + r * WLD + j0;
#pragma unroll
        for (int j = 0; j < JB; ++j) {
            const double wv = wr[j];
#pragma unroll
            for (int q = 0; q < NTT; ++q) acc[q][j] = fma(m[q], wv, acc[q][j]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NTT; ++q) {
        const int c = lane + 32 * q;
        if (c < n) {
            const double sc = rs ? rs[c] : 1.0;
#pragma unroll
            for (int j = 0; j < JB; ++j) if (j0 + j < jw) W[c * WLD + j0 + j] = acc[q][j] * sc;
        }
    }
    __syncwarp();
}

// ---- launch 3 of 3 (generation 3, default): the back kernel with NW = (wld - 1) / 4 WARPS per instance.  Same algebra and the
// same shared-memory tile as ukf_back2_kernel<wld, true>; the column groups of the vector tile (4 vectors each) are owned by
// different warps, so the reflector products -- a chain of n - 1 dependent dot / update steps per vector group, half of the
// single-warp kernel's time -- and the dense V / V^T products run NW wide; the updates' sigma-point statistics are spread over
// the warps by update, the rows of the P_pred assembly by row.  Small scalar phases are evaluated redundantly by every warp.
template <int wld, int NTT>      // NTT = row slots per lane: 4 (n_max <= 128) or 8 -- a template parameter so that the small case does not
                                 // carry the register allocation of the large one
__global__ void __launch_bounds__(32 * ((wld - 1) / 4), (wld == 9) ? (NTT == 4 ? 9 : 6) : (wld == 13) ? (NTT == 4 ? 6 : 4) : 2)
ukf_back3_kernel(BatchState b, FilterConst fc, StepInputs in, UkfScratch u, const int i0, const int stage, const int last) {
    constexpr int NW = (wld - 1) / 4, NTHR = 32 * NW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UkfWarpSmem s;
    ukf_warp_carve(b, wld, smem_raw, &s);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int inst = i0 + blockIdx.x;
    const int ldp = b.fixed_ld;
    const int nsm = 2 * b.n_max + 2;
    const unsigned FULL = 0xffffffffu;
    const int wcols = wld - 1 < ukf_wcols(b) ? wld - 1 : ukf_wcols(b);
    double* const W_ = s.W;

    const int4 meta_in = b.meta[inst];
    int nm = in.n_meas[inst];
    int status = meta_in.y;
    if (status & SLAM_STATUS_SAME_STEP_REMATCH) return;
    if (u.nswp[inst] != -3) return;                // only instances the parallel eigensolver solved (explicit V in the scratch)
    if (stage > 0 && u.defer[inst] != stage) return;       // a wider pass: only the instances the pass before handed over
    int M = meta_in.x;
    const int M_start = M;
    const int n = 4 + 2 * M;                       // ukf.cpp:167
    const int ns = 2 * n + 1;
    if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
    double* gP = b.P + (size_t)inst * b.p_stride;
    double* gx = b.x + (size_t)inst * b.x_stride;
    const double* __restrict__ R = u.Zg + (size_t)inst * u.n_max * u.n_max;
    double* Yg = u.Yg + (size_t)inst * u.n_max * u.n_max;
    const double* __restrict__ Vm = u.Vg + (size_t)inst * u.n_max * u.n_max;
    const double* __restrict__ VTm = u.VTg + (size_t)inst * u.n_max * u.n_max;

    for (int i = tid; i < n; i += NTHR) { const double v = gx[i]; s.x[i] = v; u.xprior[(size_t)inst * u.n_max + i] = v; }
    for (int i = tid; i < M; i += NTHR) s.ids[i] = b.ids[(size_t)inst * b.max_lm + i];
    for (int i = tid; i < 3 * nm; i += NTHR) s.meas[i] = in.meas[(size_t)inst * b.max_meas * 3 + i];
    __syncthreads();
    if (warp == 0) {
        // eigenvalues: clip (:120), sqrt, ordered list of the clipped ones
        int nclip = 0;
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            const double dk = (k < n) ? u.dg[(size_t)k * b.batch + inst] : 1.0;
            const bool cl = (k < n) && (dk < 0.00000001);
            const unsigned mask = __ballot_sync(FULL, cl);
            const int pos = nclip + __popc(mask & ((1u << lane) - 1u));
            if (cl && pos < 32) { s.clip[pos] = k; s.corr[pos] = 0.00000001 - dk; }
            if (k < n) s.sq[k] = sqrt(cl ? 0.00000001 : dk);
            nclip += __popc(mask);
        }
        // association of every measurement (:258-274): known-ID lookup against the landmarks of the step start
        int nu = 0, st = 0;
        for (int l = 0; l < nm; ++l) {
            const int id = (int)s.meas[3 * l];                                  // :258
            int cand = INT_MAX;
            if (fc.loc) {                                                        // :262,272,300-302: every detection updates, by map id
                if (id >= 0 && id < fc.n_map) cand = id; else st |= SLAM_STATUS_BAD_ID;
            } else {
                for (int j = lane; j < M_start; j += 32) if (s.ids[j] == id) { cand = j; break; }   // :264-269
                cand = __reduce_min_sync(FULL, cand);
            }
            if (lane == 0) {
                s.assoc[l] = (cand == INT_MAX) ? -1 : cand;
                if (cand != INT_MAX) s.uq[nu] = l;
            }
            if (cand != INT_MAX) ++nu;
        }
        if (lane == 0) { s.ctl[0] = nu; s.ctl[1] = nclip; s.ctl[2] = st; }
    }
    __syncthreads();
    const int nu = s.ctl[0], nclip = s.ctl[1];
    status |= s.ctl[2];
    if (nclip > 32) { if (tid == 0) u.nswp[inst] = -1; return; }       // (see ukf_back2_kernel: the rescue pass takes the instance)

    // weights and scale are float-valued (ukf.cpp:35,114,175)
    const float W0f = 0.2f;                                                // filter.h:207
    const double W0 = (double)W0f;
    const double wgt = (double)((1 - W0f) / (2 * n));                      // :175
    const double sw = W0 + (double)(2 * n) * wgt;                          // sum of the 2n+1 weights (not 1)
    const float u_d = in.fwd[in.cmd_stride ? inst : 0], u_th = in.ang[in.cmd_stride ? inst : 0];
    const float yaw_prior = yaw_of(s.x[2], s.x[3]);                        // :182 and :139 (prior x_t)
    const double cy = (double)cos_f(yaw_prior), sy = (double)sin_f(yaw_prior);

    if (!last) {                                    // narrow tile: hand the instance over if its updates do not fit (the narrowest
                                                    // tile also when the clipped eigenvectors would not ride beside them)
        const bool over = 4 + 2 * nu + (stage == 0 && wld < 13 ? nclip : 0) > wcols;
        if (tid == 0) u.defer[inst] = over ? stage + 1 : 0;
        if (over) return;
    }
    const int nvec = 4 + 2 * nu;                    // columns of the two S-passes
    int ncf = wcols - nvec; if (ncf > nclip) ncf = nclip;   // clipped eigenvectors riding in pass A
    if (u.clip_lanes > 0 && ncf > u.clip_lanes) ncf = u.clip_lanes;        // test knob
    const int jown = 4 * warp;                      // this warp's column group of the vector tile

    // ---- clipped eigenpairs that do not fit beside pass A (never in practice): z_k = Q V e_k, folded into the landmark-block seed
    for (int c0 = ncf; c0 < nclip && c0 < 32; c0 += wcols) {
        int cnt = (nclip < 32 ? nclip : 32) - c0; if (cnt > wcols) cnt = wcols;
        for (int i = warp; i < n; i += NW) if (lane < wcols) W_[i * wld + lane] = (lane < cnt && s.clip[c0 + lane] == i) ? 1.0 : 0.0;
        __syncthreads();
        if (jown < cnt) {
            dense_cols<NTT, wld, 4>(W_, VTm, n, lane, jown, cnt, nullptr);
            apply_reflectors<false, NTT, wld>(W_, n, lane, cnt, R, jown, 4 * NW);
        }
        __syncthreads();
        for (int a = 4 + warp; a < n; a += NW)
            for (int c = 4 + lane; c < n; c += 32) {
                double add = 0.0;
                for (int q = 0; q < cnt; ++q) add += (2.0 * wgt) * s.corr[c0 + q] * W_[a * wld + q] * W_[c * wld + q];
                Yg[(size_t)a * n + c] += add;
            }
        __syncthreads();
    }

    // ---- pass A: S e_r for the vehicle rows and the rows of the landmarks being updated; clipped eigenvectors beside
    {
        int row = -1;
        if (lane < 4) row = lane;
        else if (lane < nvec && !fc.loc) { const int q = (lane - 4) >> 1; row = s.assoc[s.uq[q]] * 2 + 4 + ((lane - 4) & 1); }   // :298
        for (int i = warp; i < n; i += NW) if (lane < wcols) W_[i * wld + lane] = (i == row) ? 1.0 : 0.0;
        __syncthreads();
        if (jown < nvec + ncf) {
            if (jown < nvec) {
                apply_reflectors<true, NTT, wld>(W_, n, lane, nvec, R, jown, 4 * NW);
                dense_cols<NTT, wld, 4>(W_, Vm, n, lane, jown, nvec, s.sq);
            }
            // unit vectors e_ck (in the eigenbasis) in the clipped columns of this group
            for (int j = (jown > nvec ? jown : nvec); j < jown + 4 && j < nvec + ncf; ++j) {
                const int ck = s.clip[j - nvec];
                for (int i = lane; i < n; i += 32) W_[i * wld + j] = (i == ck) ? 1.0 : 0.0;
            }
            __syncwarp();
            dense_cols<NTT, wld, 4>(W_, VTm, n, lane, jown, nvec + ncf, nullptr);
            apply_reflectors<false, NTT, wld>(W_, n, lane, nvec + ncf, R, jown, 4 * NW);
        }
        __syncthreads();
    }

    // ---- sigma points, vehicle rows (:214-226): X = x +- column of S, motion model per sigma point
    for (int i = tid; i < ns; i += NTHR) {
        double X[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double xr = s.x[r];
            X[r] = (i == 0) ? xr : (i <= n ? xr + W_[(i - 1) * wld + r] : xr - W_[(i - 1 - n) * wld + r]);
        }
        const float yaw = yaw_of(X[2], X[3]);                              // :128
        const float ud = u_d + fc.v_d;
        s.Xp[0 * nsm + i] = X[0] + (double)(ud * cos_f(yaw));              // :129 float product
        s.Xp[1 * nsm + i] = X[1] + (double)(ud * sin_f(yaw));              // :130
        const float fsum = yaw + u_th + fc.v_th;
        const float new_yaw = (float)remainder((double)fsum, TWO_PI_REF);  // :131
        s.Xcs[0 * nsm + i] = cos_f(new_yaw);                               // :132
        s.Xcs[1 * nsm + i] = sin_f(new_yaw);                               // :133
    }
    __syncthreads();
    auto XP = [&](const int r, const int i) -> double { return r < 2 ? s.Xp[r * nsm + i] : (double)s.Xcs[(r - 2) * nsm + i]; };
    // ---- mean (:228-232) and vehicle block of the covariance (:235-240): warp 0, results in shared memory for everybody
    //      (s.veh: xp0v[0..4), mv[4..8), vv[8..18)) -- they are read late, and registers are what caps the CTAs per SM here
    if (warp == 0) {
        double xp0v[4];
        {
            double acc[4] = {0, 0, 0, 0};
            for (int i = lane; i < ns; i += 32) {
                const double wi = (i == 0) ? W0 : wgt;
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] += wi * XP(r, i);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) xp0v[r] = warp_sum(acc[r]);
            if (lane < 4) { s.xp[lane] = xp0v[lane]; s.veh[lane] = xp0v[lane]; }
        }
        double acc[14] = {0};
        for (int i = lane; i < ns; i += 32) {
            const double wi = (i == 0) ? W0 : wgt;
            double dv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) dv[r] = XP(r, i) - xp0v[r];
            int q = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = a; c < 4; ++c) acc[q++] += (wi * dv[a]) * dv[c];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[10 + a] += wi * dv[a];
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
            const double Qd[4] = {fc.V00 * cy, fc.V00 * sy, fc.V11 * cy, fc.V11 * sy};   // :183-186
            int q = 0;
            for (int a = 0; a < 4; ++a)
                for (int c = a; c < 4; ++c) { s.veh[8 + q] = acc[q] + ((a == c) ? Qd[a] : 0.0); ++q; }
            for (int a = 0; a < 4; ++a) s.veh[4 + a] = acc[10 + a];
        }
    }
    for (int r = 4 + tid; r < n; r += NTHR) s.xp[r] = sw * s.x[r];
    __syncthreads();
    const double* const xp0v = s.veh;
    const double* const mv = s.veh + 4;
    const double* const vv = s.veh + 8;
    // ---- the updates' sigma-point statistics (:293-336): update q on warp q mod NW; it leaves hv = dz_i - dz_{i+n} in its two
    //      columns of W (the S rows it read from them are dead by then)
    for (int q = warp; q < nu; q += NW) {
        const int l = s.uq[q];
        const int li = fc.loc ? 0 : s.assoc[l] * 2 + 4;                     // :298
        const int c0 = 4 + 2 * q;
        const double mlx = fc.loc ? (double)fc.map[3 * s.assoc[l] + 1] : s.x[li];             // :152-153 (true map, float) / :144
        const double mly = fc.loc ? (double)fc.map[3 * s.assoc[l] + 2] : s.x[li + 1];
        auto sense = [&](const int i, const double sgn, const int row, double& a0, double& a1) {
            double lx = mlx, ly = mly;
            if (!fc.loc && sgn != 0.0) { lx += sgn * W_[row * wld + c0]; ly += sgn * W_[row * wld + c0 + 1]; }
            const double dx = lx - s.Xp[0 * nsm + i], dy = ly - s.Xp[1 * nsm + i];
            a0 = sqrt(dx * dx + dy * dy) + (double)fc.w_r;                  // :144
            a1 = remainder(atan2(dy, dx) - (double)yaw_prior + (double)fc.w_b, TWO_PI_REF);   // :145,156
        };
        double zest0 = 0.0;
        for (int i = lane; i < ns; i += 32) {
            double a0, a1;
            sense(i, i == 0 ? 0.0 : (i <= n ? 1.0 : -1.0), i <= n ? i - 1 : i - 1 - n, a0, a1);
            zest0 += ((i == 0) ? W0 : wgt) * a0;
        }
        zest0 = warp_sum(zest0);
        double acc[13] = {0};
        auto accum = [&](const int i, const double wi, const double a0, const double a1, double& d0, double& d1) {
            d0 = a0 - zest0;
            d1 = remainder(a1 - 0.0, TWO_PI_REF);                           // z_est(1) is never accumulated (:310-314,321)
            acc[0] += (wi * d0) * d0; acc[1] += (wi * d0) * d1; acc[2] += (wi * d1) * d1;
            acc[3] += wi * d0; acc[4] += wi * d1;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double wd = wi * (XP(a, i) - xp0v[a]);
                acc[5 + 2 * a] += wd * d0; acc[6 + 2 * a] += wd * d1;
            }
        };
        if (lane == 0) { double a0, a1, d0, d1; sense(0, 0.0, 0, a0, a1); accum(0, W0, a0, a1, d0, d1); }
        for (int i = lane; i < n; i += 32) {
            double a0, a1, b0, b1, da0, da1, db0, db1;
            sense(1 + i, 1.0, i, a0, a1);
            sense(1 + n + i, -1.0, i, b0, b1);
            accum(1 + i, wgt, a0, a1, da0, da1);
            accum(1 + n + i, wgt, b0, b1, db0, db1);
            if (!fc.loc) {
                W_[i * wld + c0] = da0 - db0;
                W_[i * wld + c0 + 1] = da1 - db1;
            }
        }
#pragma unroll
        for (int k = 0; k < 13; ++k) acc[k] = warp_sum(acc[k]);
        __syncwarp();
        if (lane == 0) {
            double* ud = s.upd + q * UPD_LD;
            const double S00 = acc[0] + fc.W00, S01 = acc[1], S11 = acc[2] + fc.W11;     // :326
            ud[0] = S00; ud[1] = S01; ud[2] = S11; ud[3] = acc[3]; ud[4] = acc[4];
            for (int a = 0; a < 8; ++a) ud[5 + a] = acc[5 + a];
            const bool swpv = fabs(S01) > fabs(S00);                        // S2^-1 (:339, partial-pivot LU like Eigen's dynamic inverse)
            const double a00 = swpv ? S01 : S00, a01 = swpv ? S11 : S01;
            const double a10 = swpv ? S00 : S01, a11 = swpv ? S01 : S11;
            const double l10 = a10 / a00, u11 = a11 - l10 * a01;
            const double b0c0 = swpv ? 0.0 : 1.0, b1c0 = swpv ? 1.0 : 0.0, b0c1 = swpv ? 1.0 : 0.0, b1c1 = swpv ? 0.0 : 1.0;
            double y1 = b1c0 - l10 * b0c0;
            const double i10 = y1 / u11, i00 = (b0c0 - a01 * i10) / a00;
            y1 = b1c1 - l10 * b0c1;
            const double i11 = y1 / u11, i01 = (b0c1 - a01 * i11) / a00;
            ud[13] = i00; ud[14] = i01; ud[15] = i10; ud[16] = i11;
            ud[17] = (double)s.meas[3 * l + 1] - zest0;                                   // innovation (:342-344)
            ud[18] = remainder((double)s.meas[3 * l + 2] - 0.0, TWO_PI_REF);
        }
    }
    __syncthreads();
    // g_a[i] = Xp[a][1+i] - Xp[a][1+n+i] -> columns 0..3
    for (int i = tid; i < n; i += NTHR)
#pragma unroll
        for (int a = 0; a < 4; ++a) W_[i * wld + a] = XP(a, 1 + i) - XP(a, 1 + n + i);
    __syncthreads();

    // ---- pass B: S g_a and S hv for every update; the clipped eigenvectors stay put in their columns
    if (jown < nvec) {
        apply_reflectors<true, NTT, wld>(W_, n, lane, nvec, R, jown, 4 * NW);
        dense_cols<NTT, wld, 4>(W_, Vm, n, lane, jown, nvec, s.sq);
        dense_cols<NTT, wld, 4>(W_, VTm, n, lane, jown, nvec, nullptr);
        apply_reflectors<false, NTT, wld>(W_, n, lane, nvec, R, jown, 4 * NW);
    }
    __syncthreads();

    // ---- gains (:336-345): C uses the RUNNING x_pred, which couples the updates only through the thread's own component a:
    //      every thread walks the updates in message order for its components; K_q overwrites the update's columns of W
    for (int a = tid; a < n; a += NTHR) {
        double xpa = s.xp[a];
        for (int q = 0; q < nu; ++q) {
            const double* ud = s.upd + q * UPD_LD;
            const int c0 = 4 + 2 * q;
            const double sdz0 = ud[3], sdz1 = ud[4];
            double c0v, c1v;
            if (a < 4) {
                const double sh = xp0v[a] - xpa;
                c0v = ud[5 + 2 * a] + sh * sdz0; c1v = ud[6 + 2 * a] + sh * sdz1;
            } else {
                const double fa = s.x[a] - xpa;
                c0v = fa * sdz0 + wgt * W_[a * wld + c0];
                c1v = fa * sdz1 + wgt * W_[a * wld + c0 + 1];
            }
            const double k0 = c0v * ud[13] + c1v * ud[15], k1 = c0v * ud[14] + c1v * ud[16];
            W_[a * wld + c0] = k0; W_[a * wld + c0 + 1] = k1;
            xpa = xpa + (k0 * ud[17] + k1 * ud[18]);
        }
        s.xp[a] = xpa;
    }
    __syncthreads();

    // ---- P_pred, written once (:235-240 then :348 per update, in message order); rows spread over the warps
    for (int i = warp; i < n; i += NW) {
        const double ei = (i >= 4) ? s.x[i] - sw * s.x[i] : 0.0;
        for (int jb = 0; jb < n; jb += 128) {
            double yv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = jb + lane + 32 * t;
                yv[t] = (i >= 4 && j >= 4 && j < n) ? Yg[(size_t)i * n + j] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = jb + lane + 32 * t;
                if (j >= n) continue;
                double val;
                if (i < 4 && j < 4) {
                    const int a = i < j ? i : j, c = i < j ? j : i;
                    val = vv[a * 4 - (a * (a - 1)) / 2 + (c - a)];
                } else if (i < 4) {
                    const double eb = s.x[j] - sw * s.x[j];
                    val = wgt * W_[j * wld + i] + eb * mv[i];
                } else if (j < 4) {
                    val = wgt * W_[i * wld + j] + ei * mv[j];
                } else {
                    const double ej = s.x[j] - sw * s.x[j];
                    double add = sw * ei * ej;
                    for (int q = 0; q < ncf; ++q) add += (2.0 * wgt) * s.corr[q] * W_[i * wld + nvec + q] * W_[j * wld + nvec + q];
                    val = yv[t] + add;
                }
                for (int q = 0; q < nu; ++q) {
                    const double* ud = s.upd + q * UPD_LD;
                    const int c0 = 4 + 2 * q;
                    const double ki0 = W_[i * wld + c0], ki1 = W_[i * wld + c0 + 1];
                    const double ks0 = ki0 * ud[0] + ki1 * ud[1], ks1 = ki0 * ud[1] + ki1 * ud[2];
                    val -= ks0 * W_[j * wld + c0] + ks1 * W_[j * wld + c0 + 1];
                }
                gP[(size_t)i * ldp + j] = val;
            }
        }
    }
    __syncthreads();

    // -------- landmarkInsertion for the unmatched measurements, in message order (:278-287,351-371); warp 0 writes
    for (int l = 0; l < nm && !fc.loc; ++l) {
        if (s.assoc[l] != -1) continue;
        if (M >= b.max_lm) { status |= SLAM_STATUS_CAPACITY; continue; }
        const int nn = 4 + 2 * M;
        if (warp == 0) {
            const float r = s.meas[3 * l + 1], bb = s.meas[3 * l + 2];
            if (lane == 0) {
                const float yaw = yaw_of(s.xp[2], s.xp[3]);                     // :356
                const float yb = yaw + bb;
                s.xp[nn] = s.xp[0] + (double)(r * cos_f(yb));                   // :358
                s.xp[nn + 1] = s.xp[1] + (double)(r * sin_f(yb));               // :359
                s.ids[M] = (int)s.meas[3 * l];                                  // :361
            }
            for (int i = lane; i < nn + 2; i += 32) {                           // :365-368, W block and zero cross terms
                gP[(size_t)i * ldp + nn] = (i == nn) ? fc.W00 : 0.0;
                gP[(size_t)i * ldp + nn + 1] = (i == nn + 1) ? fc.W11 : 0.0;
                gP[(size_t)nn * ldp + i] = (i == nn) ? fc.W00 : 0.0;
                gP[(size_t)(nn + 1) * ldp + i] = (i == nn + 1) ? fc.W11 : 0.0;
            }
        }
        M += 1;
    }
    __syncthreads();

    // ---- commit (:289-290)
    const int n_out = 4 + 2 * M;
    int bad = 0;
    for (int i = tid; i < n_out; i += NTHR) {
        const double v = s.xp[i];
        gx[i] = v;
        if (!isfinite(v)) bad = 1;
    }
    bad = __syncthreads_or(bad);
    for (int i = tid + M_start; i < M; i += NTHR) b.ids[(size_t)inst * b.max_lm + i] = s.ids[i];
    for (int i = tid; i < nm; i += NTHR) b.assoc[(size_t)inst * b.max_meas + i] = s.assoc[i];
    if (tid == 0) {
        if (bad) status |= SLAM_STATUS_NAN;
        b.meta[inst] = make_int4(M, status, meta_in.z + 1, nm);             // timestep, :164
        atomicAdd(u.routes + 0, 1ull);
        u.sigfmt[inst] = make_int2(3, n);
        if (M > M_start) atomicMax(b.max_M, M);
        double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
        const double nd = (double)n;
        st[8] += 16.0 * nd * nd + 16.0 * nd + 12.0 * nm + 8.0;
        st[9] += 9.0 * nd * nd * nd + 2.0 * nd * nd * nd + 2.0 * nd * nd * (2.0 * nd + 1.0) + 12.0 * nu * nd * nd;
        st[10] += nd;
        st[11] += (double)nm;
        // bytes the three launches really move for the instance: P in (front) and out (back), reflectors + seed out and in, V and
        // V^T out (eigensolver; its work space is the shared-memory tile) and in twice each (two S-passes)
        st[12] += 8.0 * nd * nd * (6.0 + 2.0 + 2.0 + 4.0);
        // flops they execute: tridiagonalisation 4/3 n^3, eigensolver ~15 Sturm sweeps of 3 n per eigenvalue (grid + bisection until
        // isolated + secant steps) + ~40 n per vector,
        // per vector of the two S-passes 4 n^2 (V^T, V) + 4 n^2 (Q^T, Q through the reflectors), P_pred assembly
        st[13] += 4.0 / 3.0 * nd * nd * nd + nd * nd * (45.0 + 40.0) + (double)(2 * nvec + ncf) * 8.0 * nd * nd
                  + nd * nd * (2.0 + 8.0 * nu + 2.0 * ncf);
    }
}

// ---------------------------------------------------------------------------------------------------------
// slam_get_sigma_points: the matrix X of the last step (ukf.cpp:214-220; published point-major by ukf.cpp:91-99),
//   X[:,0] = x_t,  X[:,1+c] = x_t + S e_c,  X[:,1+n+c] = x_t - S e_c,   S = sqrt(nearestSPD) of the step's prior.
// The step kernels never form S; it is materialised here, on demand, from what the step left in the scratch.
// Generation 2: S e_c = Q V sqrt(D+) V^T Q^T e_c -- one warp per block of 32 columns pushes unit vectors through the
// reflectors and the rotation log exactly like pass A of ukf_back2_kernel.  Generation 1 / rescue: explicit Z^T.
// ---------------------------------------------------------------------------------------------------------
template <bool DENSE>
__global__ void __launch_bounds__(32)
ukf_sigma2_kernel(BatchState b, UkfScratch u, const int inst, const int n, double* __restrict__ X) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int wld = 33;
    UkfWarpSmem s;
    ukf_warp_carve(b, wld, smem_raw, &s);
    const int lane = threadIdx.x;
    const int c0 = blockIdx.x * 32;
    const int cnt = (n - c0 < 32) ? n - c0 : 32;
    const unsigned FULL = 0xffffffffu;
    double* const W_ = s.W;
    const double* __restrict__ R = u.Zg + (size_t)inst * u.n_max * u.n_max;
    const double2* __restrict__ rot = u.rot + (size_t)inst * u.rot_cap;
    const int2* __restrict__ swp = u.swp + (size_t)inst * u.swp_cap;
    const int nsw = DENSE ? 0 : u.nswp[inst];
    const double* __restrict__ Vm = u.Vg + (size_t)inst * u.n_max * u.n_max;
    const double* __restrict__ VTm = u.VTg + (size_t)inst * u.n_max * u.n_max;
    for (int k = lane; k < n; k += 32) {
        const double dk = u.dg[(size_t)k * b.batch + inst];
        s.sq[k] = sqrt(dk < 0.00000001 ? 0.00000001 : dk);                  // ukf.cpp:120 and the eigen-sqrt (D-3)
        s.x[k] = u.xprior[(size_t)inst * u.n_max + k];
    }
    int nrot = 0;
    for (int q = lane; q < nsw; q += 32) { const int2 lm = swp[q]; nrot += lm.y - lm.x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrot += __shfl_xor_sync(FULL, nrot, o);
    const bool act = lane < cnt;
    for (int i = 0; i < n; ++i) W_[i * wld + lane] = (act && i == c0 + lane) ? 1.0 : 0.0;
    __syncwarp();
    if (b.n_max <= 128) apply_reflectors<true, 4, wld>(W_, n, lane, cnt, R); else apply_reflectors<true, 8, wld>(W_, n, lane, cnt, R);
    if (DENSE) {
        if (b.n_max <= 128) { dense_apply<4, wld>(W_, Vm, n, lane, cnt, cnt, s.sq); dense_apply<4, wld>(W_, VTm, n, lane, cnt, cnt, nullptr); }
        else { dense_apply<8, wld>(W_, Vm, n, lane, cnt, cnt, s.sq); dense_apply<8, wld>(W_, VTm, n, lane, cnt, cnt, nullptr); }
    } else {
        apply_rot_fwd<wld>(W_, lane, act, rot, swp, nsw, nrot, s.stage);
        if (act) for (int i = 0; i < n; ++i) W_[i * wld + lane] *= s.sq[i];
        __syncwarp();
        apply_rot_bwd<wld>(W_, lane, act, rot, swp, nsw, nrot, s.stage);
    }
    if (b.n_max <= 128) apply_reflectors<false, 4, wld>(W_, n, lane, cnt, R); else apply_reflectors<false, 8, wld>(W_, n, lane, cnt, R);
    // lanes walk the components: consecutive addresses of one sigma point
    for (int c = 0; c < cnt; ++c)
        for (int i = lane; i < n; i += 32) {
            const double sv = W_[i * wld + c], xi = s.x[i];
            X[(size_t)(1 + c0 + c) * n + i] = xi + sv;
            X[(size_t)(1 + n + c0 + c) * n + i] = xi - sv;
        }
    if (blockIdx.x == 0) for (int i = lane; i < n; i += 32) X[i] = s.x[i];
}

__global__ void ukf_sigma1_kernel(BatchState b, UkfScratch u, const int inst, const int n, double* __restrict__ X) {
    const double* __restrict__ Zt = u.Zg + (size_t)inst * u.n_max * u.n_max;   // [k][i], compact
    const double* __restrict__ xp = u.xprior + (size_t)inst * u.n_max;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * n; idx += gridDim.x * blockDim.x) {
        const int c = idx / n, i = idx - c * n;
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += Zt[(size_t)k * n + i] * (u.dg[(size_t)k * b.batch + inst] * Zt[(size_t)k * n + c]);
        X[(size_t)(1 + c) * n + i] = xp[i] + acc;
        X[(size_t)(1 + n + c) * n + i] = xp[i] - acc;
        if (c == 0) X[i] = xp[i];
    }
}

size_t ukf_step_smem_bytes(const BatchState& b) { return ukf_smem_carve(b, nullptr, nullptr); }

static size_t ql_smem_bytes(const BatchState& b) { return sizeof(double) * 2 * (size_t)b.n_max * QL_LANES; }

static size_t ukf_warp_smem_bytes(const BatchState& b, int wld = 0) { return ukf_warp_carve(b, wld ? wld : ukf_wld(b), nullptr, nullptr); }

cudaError_t ukf_step_configure(const BatchState& b) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(ukf_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_step_smem_bytes(b))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_step_smem_bytes(b))) != cudaSuccess) return e;
    {
        size_t fp = front2_carve(b.n_max, nullptr, nullptr, true), ff = front2_carve(b.n_max, nullptr, nullptr, false);
        if (fp > 227 * 1024) fp = 227 * 1024;
        if (ff > 227 * 1024) ff = 227 * 1024;
        if ((e = cudaFuncSetAttribute(ukf_front2_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_front2_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_front2_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_front2_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_front2_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_front2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024)) != cudaSuccess) return e;
    }
    if ((e = cudaFuncSetAttribute(ukf_back2_kernel<13, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 13))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back2_kernel<25, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 25))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back2_kernel<33, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 33))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back2_kernel<13, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 13))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back2_kernel<25, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 25))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back2_kernel<33, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 33))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<13, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 13))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<25, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 25))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<33, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 33))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<9, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 9))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<9, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 9))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<13, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 13))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<25, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 25))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ukf_back3_kernel<33, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 33))) != cudaSuccess) return e;
    if (ukf_warp_smem_bytes(b, 33) <= 227 * 1024) {
        if ((e = cudaFuncSetAttribute(ukf_sigma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 33))) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_sigma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ukf_warp_smem_bytes(b, 33))) != cudaSuccess) return e;
    }
    if (b.n_max <= 256) {
        if ((e = cudaFuncSetAttribute(ukf_eig3_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig3_carve(b.n_max, 128, nullptr, nullptr))) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ukf_eig3_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig3_carve(b.n_max, 256, nullptr, nullptr))) != cudaSuccess) return e;
        int caps[EIG3_MAX_CLASSES];
        if (eig3_classes(b, caps) > 0) {
            if ((e = cudaFuncSetAttribute(ukf_eig3_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig3_carve(b.n_max, 128, nullptr, nullptr, true))) != cudaSuccess) return e;
            if ((e = cudaFuncSetAttribute(ukf_eig3_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig3_carve(b.n_max < 64 ? b.n_max : 64, 64, nullptr, nullptr, true))) != cudaSuccess) return e;
        }
    }
    return cudaFuncSetAttribute(ukf_ql_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ql_smem_bytes(b));
}

cudaError_t launch_ukf_sigma_points(const BatchState& b, const UkfScratch& u, int inst, int fmt, int n, double* d_X, cudaStream_t st) {
    if (fmt == 3) ukf_sigma2_kernel<true><<<(n + 31) / 32, 32, ukf_warp_smem_bytes(b, 33), st>>>(b, u, inst, n, d_X);
    else if (fmt == 2) ukf_sigma2_kernel<false><<<(n + 31) / 32, 32, ukf_warp_smem_bytes(b, 33), st>>>(b, u, inst, n, d_X);
    else ukf_sigma1_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(b, u, inst, n, d_X);
    return cudaGetLastError();
}

// NaiveFilter::update, filter.h:342-348: measurements ignored, pose propagated by the command (one thread per instance)
__global__ void naive_step_kernel(BatchState b, StepInputs in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.batch) return;
    double* x = b.x + (size_t)i * b.x_stride;
    const float fwd = in.fwd[in.cmd_stride ? i : 0], ang = in.ang[in.cmd_stride ? i : 0];
    const double th = x[2];
    x[0] = x[0] + (double)fwd * cos(th);
    x[1] = x[1] + (double)fwd * sin(th);
    x[2] = remainder(th + (double)ang, TWO_PI_REF);
    int4 m = b.meta[i];
    m.z += 1; m.w = 0;
    b.meta[i] = m;
}
cudaError_t launch_naive_step(const BatchState& b, const StepInputs& in, cudaStream_t st) {
    naive_step_kernel<<<(b.batch + 127) / 128, 128, 0, st>>>(b, in);
    return cudaGetLastError();
}

bool ukf_gen2_supported(const BatchState& b) { return b.max_meas <= UKF2_MAX_UPD && b.n_max <= 256 && ukf_warp_smem_bytes(b) <= 227 * 1024; }

cudaError_t launch_ukf_step(const BatchState& b, const FilterConst& fc, const StepInputs& in, const UkfScratch& u, cudaStream_t st,
                            const UkfStreams& xs, int* launched) {
    const int qblocks = (b.batch + QL_LANES - 1) / QL_LANES;
    if (u.gen >= 2 && ukf_gen2_supported(b)) {
        const bool gen3 = u.gen == 3;
        // The batch is cut into nsub contiguous slices (slam_tune key 10; 0 = automatic), each with its own chain of launches on its
        // own stream, so that one slice's kernels run under the tails of the other's (every launch ends with a partial wave, and
        // the size-class launches have several).  Generation 3, 4096 instances: 2 slices 2.6 % faster than 1 (3 slices 1.8 %, 4
        // slices 1.1 %): automatic = 2 from 2048 instances.  Generation 2 keeps 1: its QL kernel is a pure latency chain, one
        // thread per instance, that stretches when it shares schedulers (measured 30 % SLOWER in slices with the shared-memory-free
        // variant that can co-reside; 2 % faster with the shared-memory one, which cannot).
        int nsub = xs.nsub > 0 ? xs.nsub : ((gen3 && b.batch >= 2048) ? 2 : 1);
        if (nsub > UKF_MAX_SUB) nsub = UKF_MAX_SUB;
        if (b.batch < 64 * nsub) nsub = 1;
        cudaError_t e;
        if (nsub > 1) {
            if ((e = cudaEventRecord(xs.fork, st)) != cudaSuccess) return e;
            for (int k = 1; k < nsub; ++k) if ((e = cudaStreamWaitEvent(xs.aux[k - 1], xs.fork, 0)) != cudaSuccess) return e;
        }
        const size_t fsm = ukf_step_smem_bytes(b), wsm = ukf_warp_smem_bytes(b);
        int nback = 0;
        for (int k = 0; k < nsub; ++k) {
            const int i0 = (int)((long long)b.batch * k / nsub), i1 = (int)((long long)b.batch * (k + 1) / nsub);
            if (i1 <= i0) continue;
            cudaStream_t sk = (k == 0) ? st : xs.aux[k - 1];
            {
                // classes: (largest n, CTAs per SM, packed?) ascending.  front_packed = 2 (default, hybrid): the full square while it
                // fits four times per SM (column pairs, no bank conflicts), the packed triangle beyond; 1 = packed only; 0 = full only
                int capsf[3], minbf[3], capsp[3], minbp[3], cap_[6], minb_[6], pk_[6], ncls = 0;
                const int nf = u.front_packed != 1 ? front2_classes(b, capsf, minbf, false) : 0;
                const int np_ = u.front_packed != 0 ? front2_classes(b, capsp, minbp, true) : 0;
                int covered = 0;
                for (int c = 0; c < nf; ++c)
                    if (u.front_packed == 0 || minbf[c] == 4) { cap_[ncls] = capsf[c]; minb_[ncls] = minbf[c]; pk_[ncls] = 0; covered = capsf[c]; ++ncls; }
                for (int c = 0; c < np_; ++c)
                    if (capsp[c] > covered) { cap_[ncls] = capsp[c]; minb_[ncls] = minbp[c]; pk_[ncls] = 1; covered = capsp[c]; ++ncls; }
                for (int c = 0; c < ncls; ++c) {
                    const int lo = c ? cap_[c - 1] : 0;
                    const size_t sm = front2_carve(cap_[c], nullptr, nullptr, pk_[c] != 0);
                    if (pk_[c]) {
                        if (minb_[c] == 4) ukf_front2_kernel<4, true><<<i1 - i0, FRONT2_THREADS, sm, sk>>>(b, u, i0, lo, cap_[c]);
                        else if (minb_[c] == 3) ukf_front2_kernel<3, true><<<i1 - i0, FRONT2_THREADS, sm, sk>>>(b, u, i0, lo, cap_[c]);
                        else ukf_front2_kernel<2, true><<<i1 - i0, FRONT2_THREADS, sm, sk>>>(b, u, i0, lo, cap_[c]);
                    } else {
                        if (minb_[c] == 4) ukf_front2_kernel<4, false><<<i1 - i0, FRONT2_THREADS, sm, sk>>>(b, u, i0, lo, cap_[c]);
                        else if (minb_[c] == 3) ukf_front2_kernel<3, false><<<i1 - i0, FRONT2_THREADS, sm, sk>>>(b, u, i0, lo, cap_[c]);
                        else ukf_front2_kernel<2, false><<<i1 - i0, FRONT2_THREADS, sm, sk>>>(b, u, i0, lo, cap_[c]);
                    }
                }
                nback += ncls - 1;
            }
            const int full = ukf_wld(b);
            const bool two_pass = u.narrow && full > UKF_NARROW_WLD;
            if (gen3) {
                // parallel eigensolver + dense S-products; the instances it declines (nswp = -2) go on to the QL route below
                // tile classes (eig3_tile = 1: all of them; 2: only those that fit at least three times per SM -- the sizes beyond go to the
                // global-scratch kernel, whose chains prefetch their pivots and which keeps 4 CTAs per SM at any size)
                int caps[EIG3_MAX_CLASSES], per[EIG3_MAX_CLASSES], nts[EIG3_MAX_CLASSES];
                int ncls = u.eig3_tile ? eig3_classes(b, caps, per, nts) : 0;
                if (u.eig3_tile == 2) while (ncls > 0 && per[ncls - 1] < 3) --ncls;
                for (int c = 0; c < ncls; ++c) {
                    const size_t sm = eig3_carve(caps[c], nts[c], nullptr, nullptr, true);
                    if (nts[c] == 64) ukf_eig3_kernel<64, true><<<i1 - i0, 64, sm, sk>>>(b, u, i0, u.maxc, 0, c ? caps[c - 1] : 0, caps[c]);
                    else ukf_eig3_kernel<128, true><<<i1 - i0, 128, sm, sk>>>(b, u, i0, u.maxc, 0, c ? caps[c - 1] : 0, caps[c]);
                }
                if (b.n_max <= 128) ukf_eig3_kernel<128, false><<<i1 - i0, 128, eig3_carve(b.n_max, 128, nullptr, nullptr), sk>>>(b, u, i0, u.maxc, ncls ? 1 : 0, ncls ? caps[ncls - 1] : 0, b.n_max);
                else ukf_eig3_kernel<256, false><<<i1 - i0, 256, eig3_carve(b.n_max, 256, nullptr, nullptr), sk>>>(b, u, i0, u.maxc, 0, 0, b.n_max);
                nback += ncls;
                if (u.multiwarp) {
                    // tiles of 8, 12 and `full` - 1 vector columns (2, 3, (full - 1) / 4 warps): an instance runs in the narrowest
                    // one that holds its 4 + 2 k columns, the others see it only in the hand-over word
                    const size_t sm9 = ukf_warp_smem_bytes(b, 9), sm13 = ukf_warp_smem_bytes(b, 13);
                    if (b.n_max <= 128) {
                        if (u.narrow) {
                            const int s13 = u.narrow >= 2 ? 1 : 0;
                            if (s13) ukf_back3_kernel<9, 4><<<i1 - i0, 64, sm9, sk>>>(b, fc, in, u, i0, 0, 0);
                            ukf_back3_kernel<13, 4><<<i1 - i0, 96, sm13, sk>>>(b, fc, in, u, i0, s13, full == 13 ? 1 : 0);
                            if (full == 25) ukf_back3_kernel<25, 4><<<i1 - i0, 192, wsm, sk>>>(b, fc, in, u, i0, s13 + 1, 1);
                            else if (full == 33) ukf_back3_kernel<33, 4><<<i1 - i0, 256, wsm, sk>>>(b, fc, in, u, i0, s13 + 1, 1);
                        } else {
                            if (full == 13) ukf_back3_kernel<13, 4><<<i1 - i0, 96, sm13, sk>>>(b, fc, in, u, i0, 0, 1);
                            else if (full == 25) ukf_back3_kernel<25, 4><<<i1 - i0, 192, wsm, sk>>>(b, fc, in, u, i0, 0, 1);
                            else ukf_back3_kernel<33, 4><<<i1 - i0, 256, wsm, sk>>>(b, fc, in, u, i0, 0, 1);
                        }
                    } else {
                        if (u.narrow) {
                            const int s13 = u.narrow >= 2 ? 1 : 0;
                            if (s13) ukf_back3_kernel<9, 8><<<i1 - i0, 64, sm9, sk>>>(b, fc, in, u, i0, 0, 0);
                            ukf_back3_kernel<13, 8><<<i1 - i0, 96, sm13, sk>>>(b, fc, in, u, i0, s13, full == 13 ? 1 : 0);
                            if (full == 25) ukf_back3_kernel<25, 8><<<i1 - i0, 192, wsm, sk>>>(b, fc, in, u, i0, s13 + 1, 1);
                            else if (full == 33) ukf_back3_kernel<33, 8><<<i1 - i0, 256, wsm, sk>>>(b, fc, in, u, i0, s13 + 1, 1);
                        } else {
                            if (full == 13) ukf_back3_kernel<13, 8><<<i1 - i0, 96, sm13, sk>>>(b, fc, in, u, i0, 0, 1);
                            else if (full == 25) ukf_back3_kernel<25, 8><<<i1 - i0, 192, wsm, sk>>>(b, fc, in, u, i0, 0, 1);
                            else ukf_back3_kernel<33, 8><<<i1 - i0, 256, wsm, sk>>>(b, fc, in, u, i0, 0, 1);
                        }
                    }
                    nback += u.narrow >= 2 ? 1 : 0;
                } else {
                    if (two_pass || full == 13) ukf_back2_kernel<13, true><<<i1 - i0, 32, ukf_warp_smem_bytes(b, 13), sk>>>(b, fc, in, u, i0, two_pass ? 0 : 2);
                    if (full == 25) ukf_back2_kernel<25, true><<<i1 - i0, 32, wsm, sk>>>(b, fc, in, u, i0, two_pass ? 1 : 2);
                    else if (full == 33) ukf_back2_kernel<33, true><<<i1 - i0, 32, wsm, sk>>>(b, fc, in, u, i0, two_pass ? 1 : 2);
                }
                nback += 1 + (two_pass ? 2 : 1);
            }
            if (nsub == 1) ukf_ql_kernel<true><<<(i1 - i0 + QL_LANES - 1) / QL_LANES, QL_LANES, ql_smem_bytes(b), sk>>>(u, b.meta, b.batch, i0, i1, gen3 ? 1 : 0);
            else ukf_ql_kernel<false><<<(i1 - i0 + QL_LANES - 1) / QL_LANES, QL_LANES, 0, sk>>>(u, b.meta, b.batch, i0, i1, gen3 ? 1 : 0);
            if (two_pass || full == 13) ukf_back2_kernel<13, false><<<i1 - i0, 32, ukf_warp_smem_bytes(b, 13), sk>>>(b, fc, in, u, i0, two_pass ? 0 : 2);
            if (full == 25) ukf_back2_kernel<25, false><<<i1 - i0, 32, wsm, sk>>>(b, fc, in, u, i0, two_pass ? 1 : 2);
            else if (full == 33) ukf_back2_kernel<33, false><<<i1 - i0, 32, wsm, sk>>>(b, fc, in, u, i0, two_pass ? 1 : 2);
            nback += (two_pass ? 2 : 1);
        }
        for (int k = 1; k < nsub; ++k) {
            if ((e = cudaEventRecord(xs.join[k - 1], xs.aux[k - 1])) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(st, xs.join[k - 1], 0)) != cudaSuccess) return e;
        }
        // rescue pass (generation-1 kernels, in-CTA QL): instances whose rotation log overflowed; everyone else exits at once
        ukf_front_kernel<<<b.batch, UKF_THREADS, fsm, st>>>(b, fc, in, u, 1);
        ukf_back_kernel<<<b.batch, UKF_THREADS, fsm, st>>>(b, fc, in, u, 1);
        if (launched) *launched = 2 * nsub + nback + 2;
    } else {
        ukf_front_kernel<<<b.batch, UKF_THREADS, ukf_step_smem_bytes(b), st>>>(b, fc, in, u, 0);
        ukf_ql_kernel<true><<<qblocks, QL_LANES, ql_smem_bytes(b), st>>>(u, b.meta, b.batch, 0, b.batch, 0);
        ukf_back_kernel<<<b.batch, UKF_THREADS, ukf_step_smem_bytes(b), st>>>(b, fc, in, u, 0);
        if (launched) *launched = 3;
    }
    return cudaGetLastError();
}

}  // namespace slam
