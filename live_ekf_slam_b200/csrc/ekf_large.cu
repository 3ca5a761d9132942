// ekf_large.cu -- single large-map EKF-SLAM instance (thousands of landmarks): P lives in HBM and the k landmark
// updates of a step are applied as ONE rank-2k FP64 tensor-core contraction (DMMA, mma.sync.m8n8k4.f64).
//
// Restates EKF::update, ekf_ws/src/localization_pkg/src/ekf.cpp:37-179, for state sizes where P (n^2 * 8 B, 128 MB at
// n = 4003) cannot be staged on chip.  The reference applies  P <- P - (K_q H_q) P  once per re-observed landmark
// (:140, a dense n^3 GEMM each).  Here the step's updates are deferred:
//     P_m = P_0 - sum_{q<=m} K_q G_q,      G_q = H_q P_{q-1}  (2 x n),   K_q = P_{q-1} H_q^T S_q^-1  (n x 2)
// Every quantity update m+1 needs -- the five rows/columns of P_m that H touches -- is reconstructed from P_0 plus
// O(n m) low-rank corrections, so P_0 is read only along 5 rows / 5 columns per measurement, and the whole step ends
// with  P <- P_0 + U G,  U = [-K_1 .. -K_k] (n x 2k),  G = [G_1; ..; G_k] (2k x n): 4 k n^2 flops over 16 n^2 bytes,
// tensor-bound for k >~ 24.  Each element still receives its k corrections in measurement order, so
// the result equals the sequential evaluation up to the rounding of the accumulation.
// Insertions (:141-173) are materialised into P_0 immediately (rows/cols built from the reconstructed P_m) with the
// earlier K_q / G_q zero-extended, which keeps the single closing contraction valid.
#include "common.cuh"

#include <climits>

namespace slam {

// ---- predict (ekf.cpp:43-61): x_pred, then T = F_x P (rows 0,1 pick up row 2)
__global__ void lm_predict_rows(LargeState L, FilterConst fc, const float* fwd, const float* ang) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int M = L.meta[0].x, n = 3 + 2 * M;
    const float d_d = fwd[0], d_th = ang[0];
    const double th = L.x[2];
    double sn, c;
    sincos(th, &sn, &c);
    const double fa = (double)(-1 * d_d) * sn, fb = (double)d_d * c;
    if (j < n) {
        const double p2 = L.P[(size_t)2 * L.ld + j];
        L.P[j] = L.P[j] + fa * p2;
        L.P[(size_t)L.ld + j] = L.P[(size_t)L.ld + j] + fb * p2;
        double xv = L.x[j];
        if (j == 0) xv = L.x[0] + (double)(d_d + fc.v_d) * c;
        if (j == 1) xv = L.x[1] + (double)(d_d + fc.v_d) * sn;
        if (j == 2) xv = remainder(L.x[2] + (double)d_th + (double)fc.v_th, TWO_PI_REF);
        L.xp[j] = xv;
    }
    if (j == 0) { L.cur[0] = 0; L.cur[1] = M; L.cur[2] = L.meta[0].y; L.cur[3] = M; }   // also valid when the step has no detections
}
// P' = T F_x^T (cols 0,1 pick up col 2) + (F_v V) F_v^T
__global__ void lm_predict_cols(LargeState L, FilterConst fc, const float* fwd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 3 + 2 * L.cur[3];
    if (i >= n) return;
    const float d_d = fwd[0];
    double sn, c;
    sincos(L.x[2], &sn, &c);
    const double fa = (double)(-1 * d_d) * sn, fb = (double)d_d * c;
    double* row = L.P + (size_t)i * L.ld;
    const double t2 = row[2];
    double p0 = row[0] + t2 * fa, p1 = row[1] + t2 * fb;
    if (i == 0) { const double cV = c * fc.V00; p0 += cV * c; p1 += cV * sn; }
    if (i == 1) { const double sV = sn * fc.V00; p0 += sV * c; p1 += sV * sn; }
    row[0] = p0; row[1] = p1;
    if (i == 2) row[2] = t2 + fc.V11;
}

// ---- the sequential part of the step: ONE persistent kernel walks the measurements (ekf.cpp:73-174); the CTAs own
// slices of the state and meet at a software grid barrier once per measurement (after x_pred / U / G of the
// measurement are published; the association vote is evaluated redundantly by every CTA).  All CTAs are co-resident (cooperative launch).
constexpr int LM_THREADS = 256;
constexpr int LM_WARPS = LM_THREADS / 32;
constexpr int LM_SPAN = 32;      // state indices owned by a CTA per pass (one per lane); the warps split the q range
constexpr int LM_QMAX = 256;     // max updates per step held in the coefficient tables
constexpr int LM_VPF = 8;        // landmark estimates per thread fetched ahead of the vote's sincos (covers 2048 landmarks)
constexpr int LM_PF = 10;        // q-iterations per warp whose operands are fetched ahead of the tables (7 warps: covers 70 updates per step)

__device__ __forceinline__ void grid_sync(unsigned* counter, const unsigned nblocks, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // arrive: a release reduction (no return value, so the poll below starts at once; cumulative with the bar.sync above,
        // which ordered the CTA's stores before it)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
        const unsigned target = (++gen) * nblocks;
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    } else ++gen;
    __syncthreads();
}

__global__ void __launch_bounds__(LM_THREADS) lm_front(LargeState L, FilterConst fc, const float* meas, const int* __restrict__ d_nmeas) {
    const int n_meas = (*d_nmeas < L.max_meas) ? *d_nmeas : L.max_meas;      // the detection count never visits the host
    __shared__ double s_cq[LM_QMAX][4];   // (H K_q)   [r][s]
    __shared__ double s_eq[LM_QMAX][4];   // (G_q H^T) [s][r]
    __shared__ double s_sc[16];           // H[10], nu[2], cb, sb / x_detected, y_detected
    __shared__ double s_Sinv[4];
    __shared__ double s_part[4][LM_WARPS][LM_SPAN];       // per-warp partial sums of the low-rank corrections
    __shared__ int s_min[2];          // vote result, double buffered: measurement l + 1's word is reset while l's is in use
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = L.ld;
    const size_t ustride = (size_t)L.n_max * 2, gstride = (size_t)2 * ld;
    unsigned* bar = reinterpret_cast<unsigned*>(L.cur + 8);
    unsigned gen = 0;
    // replicated bookkeeping: every thread of every CTA tracks the same (m, M, status)
    const int M_start = L.cur[3];
    int m = 0, M = M_start, status = L.cur[2];
    if (tid == 0) s_min[0] = INT_MAX;
    __syncthreads();

    for (int l = 0; l < n_meas; ++l) {
        const float r = meas[3 * l + 1], bb = meas[3 * l + 2];
        // -------- association vote (:79-109): first match in ascending slot order = global min.
        // Every CTA scans all M landmarks itself (16 KB of L2-resident state): the vote needs no grid barrier.  Every THREAD
        // derives the detected position from its own loads of the running vehicle estimate, so those loads and the landmark
        // loads are one L2 round trip; a thread that sees a match fetches that landmark's stale x_t (what H is built from) at
        // once, so the winner can publish H without another round trip.
        const double xv0 = __ldcg(L.xp + 0), xv1 = __ldcg(L.xp + 1), xv2 = __ldcg(L.xp + 2);
        if (tid == 0) s_min[(l + 1) & 1] = INT_MAX;
        int cand = INT_MAX;
        double lxs = 0.0, lys = 0.0;
        if (!fc.id_known) {
            // the landmark estimates first (LM_VPF per thread in registers: M <= 256 LM_VPF), THEN the sincos that waits for xv2:
            // one round trip for both
            double plx[LM_VPF], ply[LM_VPF];
#pragma unroll
            for (int t = 0; t < LM_VPF; ++t) {
                const int j = tid + t * LM_THREADS;
                if (j < M) { plx[t] = __ldcg(L.xp + 3 + 2 * j); ply[t] = __ldcg(L.xp + 4 + 2 * j); }
            }
            double sa, ca;
            sincos(xv2 + (double)bb, &sa, &ca);
            const double xd = (double)(float)(xv0 + (double)r * ca);   // float x_detected, :87
            const double yd = (double)(float)(xv1 + (double)r * sa);   // float y_detected, :88
#pragma unroll
            for (int t = 0; t < LM_VPF; ++t) {                      // ascending j: the first match stays
                const int j = tid + t * LM_THREADS;
                if (j < M) {
                    const float x_diff = (float)fabs(xd - plx[t]);                             // :91
                    const float y_diff = (float)fabs(yd - ply[t]);                             // :92
                    if (x_diff < fc.min_sep && y_diff < fc.min_sep && j < cand) {
                        cand = j;
                        if (j < M_start) { lxs = L.x[2 * j + 3]; lys = L.x[2 * j + 4]; }
                    }
                }
            }
#pragma unroll 4
            for (int j = tid + LM_VPF * LM_THREADS; j < M; j += LM_THREADS) {   // (larger maps) no early exit: the loads stay in flight
                const double lx = __ldcg(L.xp + 3 + 2 * j), ly = __ldcg(L.xp + 4 + 2 * j);
                const float x_diff = (float)fabs(xd - lx);                                 // :91
                const float y_diff = (float)fabs(yd - ly);                                 // :92
                if (x_diff < fc.min_sep && y_diff < fc.min_sep && j < cand) {
                    cand = j;
                    if (j < M_start) { lxs = L.x[2 * j + 3]; lys = L.x[2 * j + 4]; }
                }
            }
        } else {
            const int want = (int)meas[3 * l];
#pragma unroll 4
            for (int j = tid; j < M; j += LM_THREADS)
                if (__ldcg(L.ids + j) == want && j < cand) {                               // :101-108
                    cand = j;
                    if (j < M_start) { lxs = L.x[2 * j + 3]; lys = L.x[2 * j + 4]; }
                }
        }
        const int mine = cand;
        cand = __reduce_min_sync(0xffffffffu, cand);
        if ((tid & 31) == 0 && cand != INT_MAX) atomicMin(&s_min[l & 1], cand);
        __syncthreads();
        const int slot = s_min[l & 1];
        int kind = 0;
        if (status & SLAM_STATUS_SAME_STEP_REMATCH) kind = 0;                  // dead for the rest of the step
        else if (slot != INT_MAX) {
            if (slot >= M_start) status |= SLAM_STATUS_SAME_STEP_REMATCH;       // ekf.cpp:115 would throw here
            else kind = 1;
        } else {
            if (M >= L.max_lm) status |= SLAM_STATUS_CAPACITY;
            else kind = 2;
        }
        if (blockIdx.x == 0 && tid == 0) L.assoc[l] = (slot == INT_MAX) ? -1 : slot;
        const int n = 3 + 2 * M;

        if (kind == 1) {
            // -------- deferred landmark update (:110-140)
            const int hc[5] = {0, 1, 2, slot * 2 + 3, slot * 2 + 4};
            if (mine == slot) {
                // the winner of the vote: landmark from the stale x_t, vehicle from the running x_pred (:115-131)
                const double dx = lxs - xv0, dy = lys - xv1;
                const float dist = (float)sqrt(dx * dx + dy * dy);
                const double dd = (double)dist, d2 = (double)(dist * dist);
                s_sc[0] = -(dx) / dd; s_sc[1] = -(dy) / dd; s_sc[2] = 0.0; s_sc[3] = dx / dd; s_sc[4] = dy / dd;
                s_sc[5] = dy / d2; s_sc[6] = -(dx) / d2; s_sc[7] = -1.0; s_sc[8] = -(dy) / d2; s_sc[9] = dx / d2;
                s_sc[14] = lxs; s_sc[15] = lys;
            }
            // Everything the rest of this measurement reads from memory depends on the slot only -- not on H, not on S -- and is
            // fetched NOW, into one register array with two layouts, so that the loads fly under the winner's H arithmetic:
            //   warps 1..7: operands of this CTA's slice of the rank-2 pass, LM_PF q-iterations x (G_q[0][idx], G_q[1][idx], U_q[idx]);
            //   warp 0:     operands of the coefficient tables and of S, two q-iterations x (K_q[hc], G_q[.][hc]).
            // A longer update list (m > 7 LM_PF resp. 64) loads the rest in place.
            const int idx0 = blockIdx.x * LM_SPAN + lane;
            constexpr int LM_RW = LM_WARPS - 1;                                 // warps that share the q range of the rank-2 pass
            double pf[4 * LM_PF];
            double pf_p[5], pv[5], xp0 = 0.0;
#pragma unroll
            for (int c = 0; c < 5; ++c) { pf_p[c] = 0.0; pv[c] = 0.0; }
            if (warp == 0) {
                if (idx0 < n) {
                    xp0 = __ldcg(L.xp + idx0);
#pragma unroll
                    for (int c = 0; c < 5; ++c) pf_p[c] = __ldcg(L.P + (size_t)hc[c] * ld + idx0);        // column idx of H P_0
                }
                if (lane < 5) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) pv[c] = __ldcg(L.P + (size_t)hc[c] * ld + hc[lane]);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int q = lane + 32 * t;
                    if (q < m) {
                        const double* Uq = L.U + q * ustride;          // holds -K_q
                        const double* Gq = L.G + q * gstride;
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            pf[20 * t + c] = -__ldcg(Uq + 2 * hc[c]); pf[20 * t + 5 + c] = -__ldcg(Uq + 2 * hc[c] + 1);
                            pf[20 * t + 10 + c] = __ldcg(Gq + hc[c]); pf[20 * t + 15 + c] = __ldcg(Gq + ld + hc[c]);
                        }
                    }
                }
            } else if (idx0 < n) {
                if (warp == 1) {
                    const double* row = L.P + (size_t)idx0 * ld;                                          // row idx of P_0 H^T
#pragma unroll
                    for (int c = 0; c < 5; ++c) pf_p[c] = __ldcg(row + hc[c]);
                }
#pragma unroll
                for (int t = 0; t < LM_PF; ++t) {
                    const int q = (warp - 1) + t * LM_RW;
                    if (q < m) {
                        const double* Gq = L.G + q * gstride;
                        pf[4 * t] = __ldcg(Gq + idx0); pf[4 * t + 1] = __ldcg(Gq + ld + idx0);
                        const double2 u = __ldcg(reinterpret_cast<const double2*>(L.U + q * ustride + 2 * (size_t)idx0));   // -K_q[i]
                        pf[4 * t + 2] = u.x; pf[4 * t + 3] = u.y;
                    }
                }
            }
            __syncthreads();                                                   // H visible
            double H[10];
#pragma unroll
            for (int q = 0; q < 10; ++q) H[q] = s_sc[q];
            if (warp == 0) {
                // coefficient tables for the earlier updates of this step and, from the same operands, the columns hc[0..4] of
                // H P_m: lanes split the q range, lane cc < 5 adds the P_0 part of column hc[cc]
                double g0[5], g1[5];
#pragma unroll
                for (int cc = 0; cc < 5; ++cc) { g0[cc] = 0.0; g1[cc] = 0.0; }
                auto table_row = [&](const int q, const double* k0, const double* k1, const double* ga, const double* gb) {
                    double c00 = 0, c01 = 0, c10 = 0, c11 = 0, e00 = 0, e01 = 0, e10 = 0, e11 = 0;
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        c00 += H[c] * k0[c]; c01 += H[c] * k1[c]; c10 += H[5 + c] * k0[c]; c11 += H[5 + c] * k1[c];
                        e00 += ga[c] * H[c]; e01 += ga[c] * H[5 + c]; e10 += gb[c] * H[c]; e11 += gb[c] * H[5 + c];
                    }
                    s_cq[q][0] = c00; s_cq[q][1] = c01; s_cq[q][2] = c10; s_cq[q][3] = c11;
                    s_eq[q][0] = e00; s_eq[q][1] = e01; s_eq[q][2] = e10; s_eq[q][3] = e11;
#pragma unroll
                    for (int cc = 0; cc < 5; ++cc) {
                        g0[cc] -= c00 * ga[cc] + c01 * gb[cc];
                        g1[cc] -= c10 * ga[cc] + c11 * gb[cc];
                    }
                };
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int q = lane + 32 * t;
                    if (q < m) table_row(q, pf + 20 * t, pf + 20 * t + 5, pf + 20 * t + 10, pf + 20 * t + 15);
                }
                for (int q = lane + 64; q < m; q += 32) {       // (beyond 64 updates in a step: operands loaded in place)
                    const double* Uq = L.U + q * ustride;
                    const double* Gq = L.G + q * gstride;
                    double k0[5], k1[5], ga[5], gb[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        k0[c] = -__ldcg(Uq + 2 * hc[c]); k1[c] = -__ldcg(Uq + 2 * hc[c] + 1);
                        ga[c] = __ldcg(Gq + hc[c]); gb[c] = __ldcg(Gq + ld + hc[c]);
                    }
                    table_row(q, k0, k1, ga, gb);
                }
#pragma unroll
                for (int cc = 0; cc < 5; ++cc) {
                    if (lane == cc) {
#pragma unroll
                        for (int c = 0; c < 5; ++c) { g0[cc] += H[c] * pv[c]; g1[cc] += H[5 + c] * pv[c]; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        g0[cc] += __shfl_xor_sync(0xffffffffu, g0[cc], o);
                        g1[cc] += __shfl_xor_sync(0xffffffffu, g1[cc], o);
                    }
                }
                if (lane == 0) {
                    double S00 = 0, S01 = 0, S10 = 0, S11 = 0;
#pragma unroll
                    for (int cc = 0; cc < 5; ++cc) {
                        S00 += g0[cc] * H[cc]; S01 += g0[cc] * H[5 + cc]; S10 += g1[cc] * H[cc]; S11 += g1[cc] * H[5 + cc];
                    }
                    S00 += fc.W00; S11 += fc.W11;
                    // S^-1 (:135) by the adjugate: one division on the dependent chain instead of the three of the LU form
                    // (same value to a few ulp; S = H P H^T + W is well conditioned, W = I under the reference's noise bug)
                    const double rdet = 1.0 / (S00 * S11 - S01 * S10);
                    const double i00 = S11 * rdet, i01 = -S01 * rdet, i10 = -S10 * rdet, i11 = S00 * rdet;
                    s_Sinv[0] = i00; s_Sinv[1] = i01; s_Sinv[2] = i10; s_Sinv[3] = i11;
                }
            } else if (tid == 32) {
                // innovation (:129-131), off the critical path: the atan2 / remainder chain runs while warp 0 builds S
                const double dx = s_sc[14] - xv0, dy = s_sc[15] - xv1;
                const float dist = (float)sqrt(dx * dx + dy * dy);
                const float ang = (float)remainder(atan2(dy, dx) - xv2, TWO_PI_REF);
                s_sc[10] = (double)(r - dist - fc.w_r);
                s_sc[11] = (double)(bb - ang - fc.w_b);
            }
            __syncthreads();
            for (int base = blockIdx.x * LM_SPAN; base < n; base += gridDim.x * LM_SPAN) {
                const int idx = base + lane;                   // this lane's row i and column j
                const bool first = base == (int)(blockIdx.x * LM_SPAN);
                double g0 = 0, g1 = 0, a0 = 0, a1 = 0;
                if (idx < n) {
                    if (warp == 0) {
#pragma unroll
                        for (int c = 0; c < 5; ++c) {          // column idx of H P_0
                            const double pvv = first ? pf_p[c] : __ldcg(L.P + (size_t)hc[c] * ld + idx);
                            g0 += H[c] * pvv; g1 += H[5 + c] * pvv;
                        }
                        if (!first) xp0 = __ldcg(L.xp + idx);
                    } else {
                        if (warp == 1) {
                            const double* row = L.P + (size_t)idx * ld;    // row idx of P_0 H^T
#pragma unroll
                            for (int c = 0; c < 5; ++c) { const double pvv = first ? pf_p[c] : __ldcg(row + hc[c]); a0 += pvv * H[c]; a1 += pvv * H[5 + c]; }
                        }
                        if (first) {
#pragma unroll
                            for (int t = 0; t < LM_PF; ++t) {       // low-rank corrections from the prefetched operands
                                const int q = (warp - 1) + t * LM_RW;
                                if (q < m) {
                                    g0 -= s_cq[q][0] * pf[4 * t] + s_cq[q][1] * pf[4 * t + 1];
                                    g1 -= s_cq[q][2] * pf[4 * t] + s_cq[q][3] * pf[4 * t + 1];
                                    a0 += pf[4 * t + 2] * s_eq[q][0] + pf[4 * t + 3] * s_eq[q][2];
                                    a1 += pf[4 * t + 2] * s_eq[q][1] + pf[4 * t + 3] * s_eq[q][3];
                                }
                            }
                        }
#pragma unroll 4
                        for (int q = (warp - 1) + (first ? LM_PF * LM_RW : 0); q < m; q += LM_RW) {  // (the rest,) q range split over warps 1..7
                            const double* Gq = L.G + q * gstride;
                            const double ga = __ldcg(Gq + idx), gb = __ldcg(Gq + ld + idx);
                            g0 -= s_cq[q][0] * ga + s_cq[q][1] * gb;
                            g1 -= s_cq[q][2] * ga + s_cq[q][3] * gb;
                            const double2 u = __ldcg(reinterpret_cast<const double2*>(L.U + q * ustride + 2 * (size_t)idx));   // -K_q[i]
                            a0 += u.x * s_eq[q][0] + u.y * s_eq[q][2];
                            a1 += u.x * s_eq[q][1] + u.y * s_eq[q][3];
                        }
                    }
                }
                s_part[0][warp][lane] = g0; s_part[1][warp][lane] = g1; s_part[2][warp][lane] = a0; s_part[3][warp][lane] = a1;
                __syncthreads();
                if (warp == 0 && idx < n) {
                    g0 = 0; g1 = 0; a0 = 0; a1 = 0;
#pragma unroll
                    for (int w = 0; w < LM_WARPS; ++w) {
                        g0 += s_part[0][w][lane]; g1 += s_part[1][w][lane]; a0 += s_part[2][w][lane]; a1 += s_part[3][w][lane];
                    }
                    double* Gm = L.G + m * gstride;
                    Gm[idx] = g0; Gm[ld + idx] = g1;
                    const double k0 = a0 * s_Sinv[0] + a1 * s_Sinv[2];
                    const double k1 = a0 * s_Sinv[1] + a1 * s_Sinv[3];
                    *reinterpret_cast<double2*>(L.U + m * ustride + 2 * (size_t)idx) = make_double2(-k0, -k1);
                    double xv = xp0 + (k0 * s_sc[10] + k1 * s_sc[11]);                  // :138
                    if (idx == 2) xv = remainder(xv, TWO_PI_REF);                       // :139
                    L.xp[idx] = xv;
                }
                if (base + (int)(gridDim.x * LM_SPAN) < n) __syncthreads();            // (s_part is reused by the next pass)
            }
            m += 1;
        } else if (kind == 2) {
            // -------- insertion (:141-173): rows/cols nn, nn+1 materialised into P_0 from the reconstructed P_m
            const int nn = n;
            double sb, cb;
            sincos(__ldcg(L.xp + 2) + (double)bb, &sb, &cb);
            const double rr_ = (double)r;
            const double g02 = -rr_ * sb, g12 = rr_ * cb;
            if (blockIdx.x == 0 && tid == 0) {
                double Pv[3][3];
                for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) {
                    double pv = __ldcg(L.P + (size_t)a * ld + c);
                    for (int q = 0; q < m; ++q) {
                        const double* Uq = L.U + q * ustride; const double* Gq = L.G + q * gstride;
                        pv += __ldcg(Uq + 2 * a) * __ldcg(Gq + c) + __ldcg(Uq + 2 * a + 1) * __ldcg(Gq + ld + c);
                    }
                    Pv[a][c] = pv;
                }
                const double gx[2][3] = {{1.0, 0.0, g02}, {0.0, 1.0, g12}};
                const double gz[2][2] = {{cb, -rr_ * sb}, {sb, rr_ * cb}};
                const double Wm[2][2] = {{fc.W00, 0.0}, {0.0, fc.W11}};
                double T3[2][3];
                for (int rr = 0; rr < 2; ++rr) for (int k = 0; k < 3; ++k) {
                    double t = gx[rr][0] * Pv[0][k]; t += gx[rr][1] * Pv[1][k]; t += gx[rr][2] * Pv[2][k];
                    T3[rr][k] = t;
                }
                for (int rr = 0; rr < 2; ++rr) for (int c2 = 0; c2 < 2; ++c2) {
                    const double t20 = gz[rr][0] * Wm[0][0] + gz[rr][1] * Wm[1][0], t21 = gz[rr][0] * Wm[0][1] + gz[rr][1] * Wm[1][1];
                    double t = T3[rr][0] * gx[c2][0]; t += T3[rr][1] * gx[c2][1]; t += T3[rr][2] * gx[c2][2];
                    t += t20 * gz[c2][0]; t += t21 * gz[c2][1];
                    L.P[(size_t)(nn + rr) * ld + nn + c2] = t;
                }
                L.xp[nn] = __ldcg(L.xp + 0) + rr_ * cb;        // :147
                L.xp[nn + 1] = __ldcg(L.xp + 1) + rr_ * sb;    // :148
                L.ids[M] = fc.id_known ? (int)meas[3 * l] : M; // :84,150
            }
            for (int idx = blockIdx.x * LM_THREADS + tid; idx < nn; idx += gridDim.x * LM_THREADS) {
                double pr[3], pc[3];                           // P_m rows 0..2 at column idx / columns 0..2 at row idx
#pragma unroll
                for (int k = 0; k < 3; ++k) { pr[k] = __ldcg(L.P + (size_t)k * ld + idx); pc[k] = __ldcg(L.P + (size_t)idx * ld + k); }
                for (int q = 0; q < m; ++q) {
                    const double* Uq = L.U + q * ustride; const double* Gq = L.G + q * gstride;
                    const double ga = __ldcg(Gq + idx), gb = __ldcg(Gq + ld + idx);
                    const double ui0 = __ldcg(Uq + 2 * (size_t)idx), ui1 = __ldcg(Uq + 2 * (size_t)idx + 1);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        pr[k] += __ldcg(Uq + 2 * k) * ga + __ldcg(Uq + 2 * k + 1) * gb;
                        pc[k] += ui0 * __ldcg(Gq + k) + ui1 * __ldcg(Gq + ld + k);
                    }
                }
                double t0 = 1.0 * pr[0]; t0 += 0.0 * pr[1]; t0 += g02 * pr[2];
                double t1 = 0.0 * pr[0]; t1 += 1.0 * pr[1]; t1 += g12 * pr[2];
                L.P[(size_t)nn * ld + idx] = t0;
                L.P[(size_t)(nn + 1) * ld + idx] = t1;
                double c0 = pc[0] * 1.0; c0 += pc[1] * 0.0; c0 += pc[2] * g02;
                double c1 = pc[0] * 0.0; c1 += pc[1] * 1.0; c1 += pc[2] * g12;
                L.P[(size_t)idx * ld + nn] = c0;
                L.P[(size_t)idx * ld + nn + 1] = c1;
            }
            M += 1;
        }
        grid_sync(bar, gridDim.x, gen);
    }
    if (blockIdx.x == 0 && tid == 0) { L.cur[0] = m; L.cur[1] = M; L.cur[2] = status; }
}

// ---- closing contraction  P <- P + U G  on the FP64 tensor cores.  U: [q][i][2] = -K_q[i], G: [q][2][ld].
// CTA tile 64 x 64, 4 warps (2 x 2), warp tile 32 x 32 = 4 x 4 DMMA m8n8k4 tiles; K chunk = 16 updates (32 deep).
constexpr int GT = 64, GK = 32, GLDA = GK + 4, GLDB = GT + 8;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// 16-byte global -> shared copy without a register round trip (LDGSTS); bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, const int bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sa), "l"(gsrc), "r"(bytes) : "memory");
}
constexpr size_t LM_GEMM_SMEM = sizeof(double) * 2 * (GT * GLDA + GK * GLDB);   // two stages of (A tile, B tile)

// P <- P + U G on the FP64 tensor cores (DMMA m8n8k4), 64 x 64 tile of P per CTA, K = 2 (updates of the step) in chunks of
// 32 staged through a two-deep cp.async pipeline: the next chunk's U / G tiles stream into shared memory while the current
// one feeds the tensor cores (ncu of the synchronous version: long-scoreboard stalls 40 % of the samples at 12 warps per SM).
// Only the tiles ON OR BELOW the diagonal are computed: sum_q K_q G_q = sum_q P H^T S^-1 H P is symmetric (the batched kernels
// already keep P packed-symmetric, ekf_batch.cu), so a tile (ty, tx), ty > tx, is written to (ty, tx) and, transposed, to
// (tx, ty); a diagonal tile keeps its lower half and mirrors it.  Half the flops and half the reads of the full square.
__global__ void __launch_bounds__(128) lm_gemm(LargeState L) {
    extern __shared__ __align__(16) double lm_gemm_smem[];
    const int m = L.cur[0];
    const int n = 3 + 2 * L.cur[1];
    if (m == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // linear index over the lower-triangular tiles, row by row: tile t = ty (ty + 1) / 2 + tx, 0 <= tx <= ty
    int ty = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((ty + 1) * (ty + 2) / 2 <= (int)blockIdx.x) ++ty;
    while (ty * (ty + 1) / 2 > (int)blockIdx.x) --ty;
    const int tx = (int)blockIdx.x - ty * (ty + 1) / 2;
    const int row0 = ty * GT, col0 = tx * GT;
    const bool diag = ty == tx;
    if (row0 >= n || col0 >= n) return;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
    const int ld = L.ld;
    const size_t ustride = (size_t)L.n_max * 2, gstride = (size_t)2 * ld;
    const int g = lane >> 2, t4 = lane & 3;
    const int K2 = 2 * m;
    const int nchunk = (K2 + GK - 1) / GK;
    // stage chunk c into buffer b: A = U[rows row0..+64][k0..k0+32) as 16-byte (k even, k odd) pairs, contiguous over the rows
    // of one update; B = G[k0..k0+32)[cols col0..+64) as 16-byte column pairs
    auto stage = [&](const int c, const int bsel) {
        double* sA = lm_gemm_smem + (size_t)bsel * (GT * GLDA + GK * GLDB);
        double* sB = sA + GT * GLDA;
        const int k0 = c * GK;
#pragma unroll
        for (int j = 0; j < (GT * GK / 2) / 128; ++j) {
            const int e = tid + 128 * j;
            const int kp = e >> 6, r = e & 63;                 // update pair index within the chunk, row
            const int kk = k0 + 2 * kp, gr = row0 + r;
            const bool ok = kk < K2 && gr < n;
            const double* src = ok ? L.U + (size_t)(kk >> 1) * ustride + 2 * (size_t)gr : L.U;
            cp_async16(sA + r * GLDA + 2 * kp, src, ok ? 16 : 0);
        }
#pragma unroll
        for (int j = 0; j < (GK * GT / 2) / 128; ++j) {
            const int e = tid + 128 * j;
            const int k = e >> 5, cp = e & 31;                 // k within the chunk, column pair
            const int kk = k0 + k, gc = col0 + 2 * cp;
            const bool ok = kk < K2 && gc < n;
            const double* src = ok ? L.G + (size_t)(kk >> 1) * gstride + (size_t)(kk & 1) * ld + gc : L.G;
            cp_async16(sB + k * GLDB + 2 * cp, src, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);
    double acc[4][4][2];
    // C fragments: row = g, cols = 2*t4, 2*t4+1 within each 8x8 tile
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
            const int r = row0 + wr + a * 8 + g, c = col0 + wc + bq * 8 + 2 * t4;
            acc[a][bq][0] = (r < n && c < n) ? L.P[(size_t)r * ld + c] : 0.0;
            acc[a][bq][1] = (r < n && c + 1 < n) ? L.P[(size_t)r * ld + c + 1] : 0.0;
        }
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            stage(c + 1, (c + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* sA = lm_gemm_smem + (size_t)(c & 1) * (GT * GLDA + GK * GLDB);
        const double* sB = sA + GT * GLDA;
#pragma unroll
        for (int ks = 0; ks < GK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = sA[(wr + a * 8 + g) * GLDA + ks + t4];
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) bf[bq] = sB[(ks + t4) * GLDB + wc + bq * 8 + g];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int bq = 0; bq < 4; ++bq) dmma_m8n8k4(acc[a][bq][0], acc[a][bq][1], af[a], bf[bq]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
            const int r = row0 + wr + a * 8 + g, c = col0 + wc + bq * 8 + 2 * t4;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int cc = c + h2;
                if (r >= n || cc >= n) continue;
                if (diag && cc > r) continue;                       // upper half of a diagonal tile: the mirror image writes it
                const double v = acc[a][bq][h2];
                L.P[(size_t)r * ld + cc] = v;
                if (cc != r) L.P[(size_t)cc * ld + r] = v;          // mirror (8 consecutive doubles per group of lanes with equal t4)
            }
        }
}

// ---- commit (ekf.cpp:176-177)
__global__ void lm_commit(LargeState L, const int* __restrict__ d_nmeas) {
    const int n_meas = (*d_nmeas < L.max_meas) ? *d_nmeas : L.max_meas;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int status = L.cur[2];
    const bool dead = (status & SLAM_STATUS_SAME_STEP_REMATCH) != 0;
    const int M = dead ? L.cur[3] : L.cur[1];
    const int n = 3 + 2 * M;
    if (!dead && i < n) L.x[i] = L.xp[i];
    if (i == 0) {
        const int4 mi = L.meta[0];
        L.meta[0] = make_int4(M, status, mi.z + 1, n_meas);
        const double nd = (double)n;
        L.stats[8] += 16.0 * nd * nd + 16.0 * nd + 12.0 * n_meas + 8.0;
        L.stats[9] += 4.0 * (double)L.cur[0] * nd * nd;
        L.stats[10] += nd;
        L.stats[11] += (double)n_meas;
        // really moved: P once each way through lm_gemm + the 5 rows / 5 columns per measurement; U and G written and re-read
        L.stats[12] += 12.0 * nd * nd + (double)L.cur[0] * (80.0 * nd + 64.0 * nd);   // lower triangle read, full square written
        L.stats[13] += 2.0 * (double)L.cur[0] * nd * nd;                               // the contraction runs on the lower triangle
    }
}

// 72 KB of dynamic shared memory (two pipeline stages) needs the opt-in; per device, so it is done for every handle
// (slam_create, after cudaSetDevice) rather than once per process
cudaError_t ekf_large_configure() {
    return cudaFuncSetAttribute(lm_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_GEMM_SMEM);
}

// U and G are zero-extended for landmarks inserted later in the step: clear the rows of this step's measurements (the count is
// read on the device, so the cost follows the message, not max_meas)
__global__ void lm_clear(LargeState L, const int* __restrict__ d_nmeas) {
    const int nm = (*d_nmeas < L.max_meas) ? *d_nmeas : L.max_meas;
    const size_t nu = (size_t)nm * L.n_max * 2, ng = (size_t)nm * 2 * L.ld;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nu + ng; i += (size_t)gridDim.x * blockDim.x) {
        if (i < nu) L.U[i] = 0.0; else L.G[i - nu] = 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) L.cur[8] = 0;        // grid-barrier counter of lm_front
}

// One reference EKF::update for the large-map instance.  meas: DEVICE pointer to [max_meas][3] float32, d_nmeas: DEVICE pointer to
// the detection count -- nothing of the step depends on the host knowing it, so a step is five asynchronous launches.
cudaError_t launch_ekf_large_step(const LargeState& L, const FilterConst& fc, const float* d_fwd, const float* d_ang,
                                  const float* d_meas, const int* d_nmeas, int n_upper, cudaStream_t st, long long* launches,
                                  cudaEvent_t gemm_ev0, cudaEvent_t gemm_ev1) {
    const int tb = 256;
    const int gb = (n_upper + tb - 1) / tb;
    cudaError_t e;
    if (L.max_meas > LM_QMAX) return cudaErrorInvalidValue;
    static int sms = 0;
    if (!sms) sms = device_sm_count();
    lm_clear<<<2 * sms, 256, 0, st>>>(L, d_nmeas);
    lm_predict_rows<<<gb, tb, 0, st>>>(L, fc, d_fwd, d_ang);
    lm_predict_cols<<<gb, tb, 0, st>>>(L, fc, d_fwd);
    int gm = (n_upper + LM_SPAN - 1) / LM_SPAN;                    // one CTA per 32 state indices, all co-resident
    if (gm > sms) gm = sms;
    LargeState Lc = L; FilterConst fcc = fc; const float* mp = d_meas; const int* np = d_nmeas;
    void* args[] = {&Lc, &fcc, &mp, &np};
    e = cudaLaunchCooperativeKernel((const void*)lm_front, dim3(gm), dim3(LM_THREADS), args, 0, st);
    if (e != cudaSuccess) return e;
    const int gt = (n_upper + GT - 1) / GT;
    if (gemm_ev0) cudaEventRecord(gemm_ev0, st);
    lm_gemm<<<gt * (gt + 1) / 2, 128, LM_GEMM_SMEM, st>>>(L);
    if (gemm_ev1) cudaEventRecord(gemm_ev1, st);
    lm_commit<<<gb, tb, 0, st>>>(L, d_nmeas);
    *launches += 6;
    return cudaGetLastError();
}

}  // namespace slam
