// ekf_batch.cu -- batched EKF-SLAM step for sm_100a: one CTA per filter instance, the covariance staged
// in shared memory by 1-D bulk async copies (cp.async.bulk / mbarrier), one HBM round trip per step.
//
// Restates EKF::update, ekf_ws/src/localization_pkg/src/ekf.cpp:37-179, with the reference's
// float/double roundings and evaluates its dense products structurally:
//   :61   F_x P F_x^T + F_v V F_v^T   -> rows/cols 0..1 pick up row/col 2, + 3x3 block      O(n)
//   :133  H P H^T + W                 -> 5 rows of P                                        O(n)
//   :135  P H^T S^-1                  -> 5 columns of P                                     O(n)
//   :140  P - (K H) P                 -> rank-2 update P -= K (H P)                         O(n^2)
//   :172  Y blkdiag(P,W) Y^T          -> two new rows/cols                                  O(n)
// so a step moves 16 n^2 bytes through HBM and is bandwidth bound.
#include "common.cuh"

#include <climits>

namespace slam {

// Launch geometry.  Shared memory is sized for the landmark capacity of THIS launch (cap_lm <= max_lm), chosen
// by the host from a slightly stale device-side max(M) plus headroom, so that early in a run -- while the map
// is still small -- many more CTAs fit per SM.  An instance whose M + detections could exceed cap_lm is
// deferred untouched to a retry list that a second, full-capacity launch drains (correctness never depends
// on the hint).
struct EkfLaunch {
    int cap_lm;        // landmark capacity of this launch's shared-memory tile
    int n_cap;         // 3 + 2*cap_lm
    int lds;           // shared-memory leading dimension for n_cap
    int from_list;     // 0: instance = blockIdx.x ; 1: instances come from b.retry_list (persistent loop)
};

struct EkfSmem {
    double* P;      // n_max x lds
    double* x;      // running x_pred
    double* xs;     // x_t at step start (stale landmark means, ekf.cpp:115)
    double* HP;     // 2 x lds   (H_x * P_pred)
    double* PH;     // n_max x 2 (P_pred * H_x^T)
    double* K;      // n_max x 2
    double* sc;     // scalars
    int* ids;
    float* meas;
    int* assoc;
    int* iscr;      // [0..1] match slots, [2] nan flag
    uint64_t* bar;
};

__host__ __device__ inline size_t ekf_smem_carve(const BatchState& b, const EkfLaunch& L, unsigned char* base, EkfSmem* s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const int nmp = ldg_of(L.n_cap);
    size_t oP = take(sizeof(double) * (size_t)L.n_cap * L.lds);
    size_t ox = take(sizeof(double) * nmp);
    size_t oxs = take(sizeof(double) * nmp);
    size_t oHP = take(sizeof(double) * 2 * L.lds);
    size_t oPH = take(sizeof(double) * 2 * nmp);
    size_t oK = take(sizeof(double) * 2 * nmp);
    size_t osc = take(sizeof(double) * 32);
    size_t oids = take(sizeof(int) * (L.cap_lm + 1));
    size_t omeas = take(sizeof(float) * 3 * b.max_meas);
    size_t oassoc = take(sizeof(int) * b.max_meas);
    size_t oi = take(sizeof(int) * 8);
    size_t obar = take(sizeof(uint64_t));
    if (s) {
        s->P = (double*)(base + oP); s->x = (double*)(base + ox); s->xs = (double*)(base + oxs);
        s->HP = (double*)(base + oHP); s->PH = (double*)(base + oPH); s->K = (double*)(base + oK);
        s->sc = (double*)(base + osc); s->ids = (int*)(base + oids); s->meas = (float*)(base + omeas);
        s->assoc = (int*)(base + oassoc); s->iscr = (int*)(base + oi); s->bar = (uint64_t*)(base + obar);
    }
    return off;
}

// scalar slots in sc[]
enum { SC_H = 0 /*10*/, SC_NU = 10 /*2*/, SC_XD = 12, SC_YD = 13, SC_CB = 14, SC_SB = 15 };

// One reference EKF::update for instance `inst`.  Returns true when the mbarrier phase `parity` was consumed.
template <int THREADS>
__device__ __forceinline__ bool ekf_instance(const BatchState& b, const FilterConst& fc, const StepInputs& in,
                                             const int phases, const EkfLaunch& L, const EkfSmem& s, const int inst,
                                             const uint32_t parity) {
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lds = L.lds;

    const int4 meta_in = b.meta[inst];
    int nm = 0;
    if (phases & STEP_UPDATE) nm = in.n_meas[inst];
    const int status_in = meta_in.y;
    if (status_in & SLAM_STATUS_SAME_STEP_REMATCH) return false;   // the reference process is dead past this point
    int M = meta_in.x;
    const int M_start = M;
    int n = 3 + 2 * M;
    int status = status_in;
    if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
    if (M + nm > L.cap_lm && L.cap_lm < b.max_lm) {
        // this launch's tile may be too small for the insertions of this step: defer, untouched
        if (tid == 0) b.retry_list[atomicAdd(b.retry_count, 1)] = inst;
        return false;
    }
    double* gP = b.P + (size_t)inst * b.p_stride;
    double* gx = b.x + (size_t)inst * b.x_stride;

    // ---- stage P: one bulk copy per live row, all completing on one mbarrier
    if (tid == 0) { s.iscr[0] = INT_MAX; s.iscr[1] = INT_MAX; s.iscr[2] = 0; }
    {
        const int ldg = ldg_of(n);
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(s.bar, (uint32_t)(n * ldg * sizeof(double)));
            __syncwarp();
            for (int row = lane; row < n; row += 32)
                bulk_g2s(s.P + (size_t)row * lds, gP + (size_t)row * ldg, (uint32_t)(ldg * sizeof(double)), s.bar);
        }
    }
    // ---- meanwhile: state, ids, messages
    for (int i = tid; i < n; i += THREADS) { const double v = gx[i]; s.x[i] = v; s.xs[i] = v; }
    for (int i = tid; i < M; i += THREADS) s.ids[i] = b.ids[(size_t)inst * b.max_lm + i];
    for (int i = tid; i < 3 * nm; i += THREADS) s.meas[i] = in.meas[(size_t)inst * b.max_meas * 3 + i];
    __syncthreads();

    // ---- PREDICT, ekf.cpp:43-61
    double fa = 0.0, fb = 0.0, c = 1.0, sn = 0.0;
    if (phases & STEP_PREDICT) {
        const float d_d = in.fwd[in.cmd_stride ? inst : 0];
        const float d_th = in.ang[in.cmd_stride ? inst : 0];
        const double th = s.xs[2];
        sincos(th, &sn, &c);
        fa = (double)(-1 * d_d) * sn;            // F_x(0,2), :48
        fb = (double)d_d * c;                    // F_x(1,2), :49
        if (tid == 0) {
            const float dv = d_d + fc.v_d;       // float add, :57-58
            s.x[0] = s.xs[0] + (double)dv * c;
            s.x[1] = s.xs[1] + (double)dv * sn;
            s.x[2] = remainder(s.xs[2] + (double)d_th + (double)fc.v_th, TWO_PI_REF);   // :59
        }
    }
    mbar_wait(s.bar, parity);
    if (phases & STEP_PREDICT) {
        // T = F_x P : rows 0,1 pick up row 2
        for (int j = tid; j < n; j += THREADS) {
            const double p2 = s.P[2 * lds + j];
            s.P[j] = s.P[j] + fa * p2;
            s.P[lds + j] = s.P[lds + j] + fb * p2;
        }
        __syncthreads();
        // P' = T F_x^T : cols 0,1 pick up col 2 ; + (F_v V) F_v^T on the vehicle block
        for (int i = tid; i < n; i += THREADS) {
            const double t2 = s.P[i * lds + 2];
            double p0 = s.P[i * lds + 0] + t2 * fa;
            double p1 = s.P[i * lds + 1] + t2 * fb;
            if (i == 0) { const double cV = c * fc.V00; p0 += cV * c; p1 += cV * sn; }
            if (i == 1) { const double sV = sn * fc.V00; p0 += sV * c; p1 += sV * sn; }
            s.P[i * lds + 0] = p0;
            s.P[i * lds + 1] = p1;
            if (i == 2) s.P[2 * lds + 2] = t2 + fc.V11;
        }
    }
    __syncthreads();

    // ---- UPDATE, ekf.cpp:63-174
    bool dead = false;
    int n_upd = 0;
    for (int l = 0; l < nm; ++l) {
        const float r = s.meas[3 * l + 1], bb = s.meas[3 * l + 2];
        int* match = &s.iscr[l & 1];
        int id;
        // -- association, :79-109
        int cand = INT_MAX;
        if (!fc.id_known) {
            id = M;
            if (tid == 0) {
                double sa, ca; sincos(s.x[2] + (double)bb, &sa, &ca);
                s.sc[SC_XD] = (double)(float)(s.x[0] + (double)r * ca);   // float x_detected, :87
                s.sc[SC_YD] = (double)(float)(s.x[1] + (double)r * sa);   // float y_detected, :88
            }
            __syncthreads();
            const double xd = s.sc[SC_XD], yd = s.sc[SC_YD];
            for (int j = tid; j < M; j += THREADS) {
                const float x_diff = (float)fabs(xd - s.x[3 + 2 * j]);     // :91
                const float y_diff = (float)fabs(yd - s.x[4 + 2 * j]);     // :92
                if (x_diff < fc.min_sep && y_diff < fc.min_sep) { cand = j; break; }   // first match per thread
            }
        } else {
            id = (int)s.meas[3 * l];                                       // :101
            for (int j = tid; j < M; j += THREADS)
                if (s.ids[j] == id) { cand = j; break; }
        }
        cand = __reduce_min_sync(0xffffffffu, cand);
        if (lane == 0 && cand != INT_MAX) atomicMin(match, cand);
        __syncthreads();
        const int slot = *match;                                           // first j in ascending order, :93-97
        if (tid == 0) { s.iscr[(l + 1) & 1] = INT_MAX; s.assoc[l] = (slot == INT_MAX) ? -1 : slot; }
        if (!fc.id_known && slot != INT_MAX) id = slot;

        if (slot != INT_MAX) {
            // -------- landmark update, :110-140
            if (slot >= M_start) { status |= SLAM_STATUS_SAME_STEP_REMATCH; dead = true; break; }
            const int i = slot * 2 + 3;
            ++n_upd;
            if (tid == 0) {
                const double dx = s.xs[i] - s.x[0], dy = s.xs[i + 1] - s.x[1];
                const float dist = (float)sqrt(dx * dx + dy * dy);         // :115
                const double dd = (double)dist;
                const double d2 = (double)(dist * dist);                   // float product, :120
                double* H = s.sc + SC_H;
                H[0] = -(dx) / dd; H[1] = -(dy) / dd; H[2] = 0.0; H[3] = dx / dd; H[4] = dy / dd;
                H[5] = dy / d2; H[6] = -(dx) / d2; H[7] = -1.0; H[8] = -(dy) / d2; H[9] = dx / d2;
                // innovation is all-float arithmetic, :129-131 (evaluated on the second warp below)
            }
            if (tid == 32) {
                const double dx = s.xs[i] - s.x[0], dy = s.xs[i + 1] - s.x[1];
                const float dist = (float)sqrt(dx * dx + dy * dy);
                const float ang = (float)remainder(atan2(dy, dx) - s.x[2], TWO_PI_REF);   // :129
                s.sc[SC_NU] = (double)(r - dist - fc.w_r);                 // :130
                s.sc[SC_NU + 1] = (double)(bb - ang - fc.w_b);             // :131
            }
            __syncthreads();
            double H[10];
#pragma unroll
            for (int q = 0; q < 10; ++q) H[q] = s.sc[SC_H + q];
            const int hc3 = i, hc4 = i + 1;
            // -- phase A: H P (2 x n) on the low half of the CTA, P H^T (n x 2) on the high half
            if (tid < THREADS / 2) {
                for (int j = tid; j < ldg_of(n); j += THREADS / 2) {
                    double h0 = 0.0, h1 = 0.0;
                    if (j < n) {
                        const double p0 = s.P[j], p1 = s.P[lds + j], p2 = s.P[2 * lds + j];
                        const double p3 = s.P[hc3 * lds + j], p4 = s.P[hc4 * lds + j];
                        h0 = H[0] * p0; h0 += H[1] * p1; h0 += H[2] * p2; h0 += H[3] * p3; h0 += H[4] * p4;
                        h1 = H[5] * p0; h1 += H[6] * p1; h1 += H[7] * p2; h1 += H[8] * p3; h1 += H[9] * p4;
                    }
                    s.HP[j] = h0; s.HP[lds + j] = h1;
                }
            } else {
                for (int q = tid - THREADS / 2; q < n; q += THREADS / 2) {
                    const double* row = s.P + (size_t)q * lds;
                    const double p0 = row[0], p1 = row[1], p2 = row[2], p3 = row[hc3], p4 = row[hc4];
                    double a0 = p0 * H[0]; a0 += p1 * H[1]; a0 += p2 * H[2]; a0 += p3 * H[3]; a0 += p4 * H[4];
                    double a1 = p0 * H[5]; a1 += p1 * H[6]; a1 += p2 * H[7]; a1 += p3 * H[8]; a1 += p4 * H[9];
                    s.PH[2 * q] = a0; s.PH[2 * q + 1] = a1;
                }
            }
            __syncthreads();
            // -- phase B: S = (H P) H^T + W, S^-1 (partial-pivot LU like Eigen's dynamic inverse(), :133-135), K, x
            {
                double S00, S01, S10, S11;
                {
                    const double a0 = s.HP[0], a1 = s.HP[1], a2 = s.HP[2], a3 = s.HP[hc3], a4 = s.HP[hc4];
                    const double b0 = s.HP[lds], b1 = s.HP[lds + 1], b2 = s.HP[lds + 2], b3 = s.HP[lds + hc3], b4 = s.HP[lds + hc4];
                    S00 = a0 * H[0]; S00 += a1 * H[1]; S00 += a2 * H[2]; S00 += a3 * H[3]; S00 += a4 * H[4];
                    S01 = a0 * H[5]; S01 += a1 * H[6]; S01 += a2 * H[7]; S01 += a3 * H[8]; S01 += a4 * H[9];
                    S10 = b0 * H[0]; S10 += b1 * H[1]; S10 += b2 * H[2]; S10 += b3 * H[3]; S10 += b4 * H[4];
                    S11 = b0 * H[5]; S11 += b1 * H[6]; S11 += b2 * H[7]; S11 += b3 * H[8]; S11 += b4 * H[9];
                    S00 += fc.W00; S11 += fc.W11;
                }
                double i00, i01, i10, i11;
                {
                    const bool sw = fabs(S10) > fabs(S00);
                    const double a00 = sw ? S10 : S00, a01 = sw ? S11 : S01, a10 = sw ? S00 : S10, a11 = sw ? S01 : S11;
                    const double l10 = a10 / a00, u11 = a11 - l10 * a01;
                    // columns of the permuted identity
                    const double b0c0 = sw ? 0.0 : 1.0, b1c0 = sw ? 1.0 : 0.0;
                    const double b0c1 = sw ? 1.0 : 0.0, b1c1 = sw ? 0.0 : 1.0;
                    double y1 = b1c0 - l10 * b0c0; i10 = y1 / u11; i00 = (b0c0 - a01 * i10) / a00;
                    y1 = b1c1 - l10 * b0c1; i11 = y1 / u11; i01 = (b0c1 - a01 * i11) / a00;
                }
                const double nu0 = s.sc[SC_NU], nu1 = s.sc[SC_NU + 1];
                for (int q = tid; q < n; q += THREADS) {
                    const double ph0 = s.PH[2 * q], ph1 = s.PH[2 * q + 1];
                    const double k0 = ph0 * i00 + ph1 * i10;
                    const double k1 = ph0 * i01 + ph1 * i11;
                    s.K[2 * q] = k0; s.K[2 * q + 1] = k1;
                    double xv = s.x[q] + (k0 * nu0 + k1 * nu1);            // :138
                    if (q == 2) xv = remainder(xv, TWO_PI_REF);            // :139
                    s.x[q] = xv;
                }
            }
            __syncthreads();
            // -- phase C: P -= K (H P), :140 as a rank-2 update over the packed row width
            {
                // warp w owns rows w, w+8, ...; a lane owns one double2 column pair per 32-pair chunk and keeps
                // its two (H P) pairs in registers for the whole sweep; 4 rows in flight per iteration.
                const int hp = ldg_of(n) >> 1;               // double2 per row
                for (int c0 = 0; c0 < hp; c0 += 32) {
                    const int jp = c0 + lane;
                    if (jp < hp) {
                        const double2 h0 = *reinterpret_cast<const double2*>(s.HP + 2 * jp);
                        const double2 h1 = *reinterpret_cast<const double2*>(s.HP + lds + 2 * jp);
                        double* col = s.P + 2 * jp;
                        int row = warp;
                        for (; row + 3 * WARPS < n; row += 4 * WARPS) {
                            double2 k[4], p[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                k[u] = *reinterpret_cast<const double2*>(s.K + 2 * (row + u * WARPS));
                                p[u] = *reinterpret_cast<const double2*>(col + (size_t)(row + u * WARPS) * lds);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                p[u].x = p[u].x - (k[u].x * h0.x + k[u].y * h1.x);
                                p[u].y = p[u].y - (k[u].x * h0.y + k[u].y * h1.y);
                                *reinterpret_cast<double2*>(col + (size_t)(row + u * WARPS) * lds) = p[u];
                            }
                        }
                        for (; row < n; row += WARPS) {
                            const double2 k = *reinterpret_cast<const double2*>(s.K + 2 * row);
                            double2 p = *reinterpret_cast<const double2*>(col + (size_t)row * lds);
                            p.x = p.x - (k.x * h0.x + k.y * h1.x);
                            p.y = p.y - (k.x * h0.y + k.y * h1.y);
                            *reinterpret_cast<double2*>(col + (size_t)row * lds) = p;
                        }
                    }
                }
            }
            __syncthreads();
        } else {
            // -------- landmark insertion, :141-173
            if (M >= b.max_lm) { status |= SLAM_STATUS_CAPACITY; if (tid == 0) s.assoc[l] = -1; __syncthreads(); continue; }
            if (tid == 0) {
                double sb, cb; sincos(s.x[2] + (double)bb, &sb, &cb);
                s.sc[SC_CB] = cb; s.sc[SC_SB] = sb;
            }
            __syncthreads();
            const double cb = s.sc[SC_CB], sb = s.sc[SC_SB];
            const double g02 = -(double)r * sb, g12 = (double)r * cb;      // G_x(0,2), G_x(1,2), :162,165
            // rows n, n+1 over old columns; columns n, n+1 over old rows
            for (int j = tid; j < n; j += THREADS) {
                const double p0 = s.P[j], p1 = s.P[lds + j], p2 = s.P[2 * lds + j];
                double t0 = 1.0 * p0; t0 += 0.0 * p1; t0 += g02 * p2;
                double t1 = 0.0 * p0; t1 += 1.0 * p1; t1 += g12 * p2;
                s.P[(size_t)n * lds + j] = t0;
                s.P[(size_t)(n + 1) * lds + j] = t1;
                const double* row = s.P + (size_t)j * lds;
                const double q0 = row[0], q1 = row[1], q2 = row[2];
                double c0 = q0 * 1.0; c0 += q1 * 0.0; c0 += q2 * g02;
                double c1 = q0 * 0.0; c1 += q1 * 1.0; c1 += q2 * g12;
                s.P[(size_t)j * lds + n] = c0;
                s.P[(size_t)j * lds + n + 1] = c1;
            }
            if (tid == 32) {
                // new 2x2 block: G_x P_vv G_x^T + G_z W G_z^T, :155-172
                const double gx[2][3] = {{1.0, 0.0, g02}, {0.0, 1.0, g12}};
                const double gz[2][2] = {{cb, -(double)r * sb}, {sb, (double)r * cb}};
                const double Wm[2][2] = {{fc.W00, 0.0}, {0.0, fc.W11}};
                double T3[2][3], T2[2][2];
                for (int rr = 0; rr < 2; ++rr) {
                    for (int k = 0; k < 3; ++k) {
                        double t = gx[rr][0] * s.P[k]; t += gx[rr][1] * s.P[lds + k]; t += gx[rr][2] * s.P[2 * lds + k];
                        T3[rr][k] = t;
                    }
                    for (int c2 = 0; c2 < 2; ++c2) T2[rr][c2] = gz[rr][0] * Wm[0][c2] + gz[rr][1] * Wm[1][c2];
                }
                for (int rr = 0; rr < 2; ++rr)
                    for (int c2 = 0; c2 < 2; ++c2) {
                        double t = T3[rr][0] * gx[c2][0]; t += T3[rr][1] * gx[c2][1]; t += T3[rr][2] * gx[c2][2];
                        t += T2[rr][0] * gz[c2][0]; t += T2[rr][1] * gz[c2][1];
                        s.P[(size_t)(n + rr) * lds + n + c2] = t;
                    }
                s.x[n] = s.x[0] + (double)r * cb;                          // :147
                s.x[n + 1] = s.x[1] + (double)r * sb;                      // :148
                s.ids[M] = id;                                             // :150
            }
            M += 1; n += 2;
            __syncthreads();
        }
    }

    // ---- commit, :176-177
    if (dead) {
        // frozen at the last committed state: nothing but the status word changes
        if (tid == 0) b.meta[inst] = make_int4(meta_in.x, status, meta_in.z, 0);
        return true;
    }
    for (int i = tid; i < n; i += THREADS) {
        const double v = s.x[i];
        gx[i] = v;
        if (!isfinite(v) || !isfinite(s.P[(size_t)i * lds + i])) s.iscr[2] = 1;
    }
    for (int i = tid + M_start; i < M; i += THREADS) b.ids[(size_t)inst * b.max_lm + i] = s.ids[i];
    for (int i = tid; i < nm; i += THREADS) b.assoc[(size_t)inst * b.max_meas + i] = s.assoc[i];
    fence_proxy_async();     // generic-proxy writes of P must be visible to the bulk-copy engine
    __syncthreads();
    if (tid == 0) {
        if (s.iscr[2]) status |= SLAM_STATUS_NAN;
        b.meta[inst] = make_int4(M, status, meta_in.z + ((phases & STEP_PREDICT) ? 1 : 0),   // timestep, :39
                                 (phases & STEP_UPDATE) ? nm : meta_in.w);
        if (M > M_start) atomicMax(b.max_M, M);
        // algorithmic work of this update, using the live n at the end of the step
        double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
        const double nd = (double)n;
        st[8] += 16.0 * nd * nd + 16.0 * nd + 12.0 * nm + 8.0;
        st[9] += 4.0 * (double)n_upd * nd * nd;
        st[10] += nd;
        st[11] += (double)nm;
    }
    if (warp == 0) {
        const int ldg = ldg_of(n);
        for (int row = lane; row < n; row += 32)
            bulk_s2g(gP + (size_t)row * ldg, s.P + (size_t)row * lds, (uint32_t)(ldg * sizeof(double)));
        bulk_commit();
        bulk_wait_all();
    }
    return true;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS >= 256) ? 2 : 6)
ekf_step_kernel(BatchState b, FilterConst fc, StepInputs in, int phases, EkfLaunch L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EkfSmem s;
    ekf_smem_carve(b, L, smem_raw, &s);
    if (threadIdx.x == 0) { mbar_init(s.bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (!L.from_list) {
        ekf_instance<THREADS>(b, fc, in, phases, L, s, (int)blockIdx.x, 0u);
        return;
    }
    // retry pass: a persistent grid drains the list of instances the capacity-limited launch deferred
    const int count = *b.retry_count;
    uint32_t parity = 0;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        if (ekf_instance<THREADS>(b, fc, in, phases, L, s, b.retry_list[q], parity)) parity ^= 1u;
        __syncthreads();
    }
}

// Filter::init on the device (ekf.cpp:8-18,29-34 / ukf.cpp:7-18,31-45): one thread per instance writes x_0 and the
// base x base P_0 block; nothing else of P is live while M == 0.
__global__ void reset_kernel(BatchState b, double x0, double y0, double a2, double a3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.batch) return;
    double* x = b.x + (size_t)i * b.x_stride;
    double* P = b.P + (size_t)i * b.p_stride;
    const int nb = b.base, ld = ldp_of(b.fixed_ld, nb);
    x[0] = x0; x[1] = y0; x[2] = a2; if (nb == 4) x[3] = a3;
    for (int r = 0; r < nb; ++r) for (int c = 0; c < ld; ++c) P[r * ld + c] = 0.0;
    P[0] = 0.01 * 0.01; P[ld + 1] = 0.01 * 0.01; P[2 * ld + 2] = 0.005 * 0.005;
    if (nb == 4) P[3 * ld + 3] = 0.005 * 0.005;
    b.meta[i] = make_int4(0, 0, 0, 0);
    double* st = b.stats + (size_t)i * SLAM_NUM_STATS;
    for (int k = 0; k < SLAM_NUM_STATS; ++k) st[k] = 0.0;
}

cudaError_t launch_reset(const BatchState& b, double x0, double y0, double a2, double a3, cudaStream_t st) {
    reset_kernel<<<(b.batch + 127) / 128, 128, 0, st>>>(b, x0, y0, a2, a3);
    return cudaGetLastError();
}

static EkfLaunch make_launch(const BatchState& b, int cap_lm, int from_list) {
    EkfLaunch L;
    L.cap_lm = cap_lm < b.max_lm ? cap_lm : b.max_lm;
    if (L.cap_lm < 1) L.cap_lm = 1;
    L.n_cap = 3 + 2 * L.cap_lm;
    L.lds = lds_of(L.n_cap);
    L.from_list = from_list;
    return L;
}

size_t ekf_step_smem_bytes(const BatchState& b) { return ekf_smem_carve(b, make_launch(b, b.max_lm, 0), nullptr, nullptr); }

cudaError_t ekf_step_configure(const BatchState& b) {
    const int bytes = (int)ekf_step_smem_bytes(b);
    cudaError_t e = cudaFuncSetAttribute(ekf_step_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ekf_step_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// cap_hint: landmark capacity to size this launch for (<= 0 or >= max_lm: full capacity, no retry pass).
cudaError_t launch_ekf_step(const BatchState& b, const FilterConst& fc, const StepInputs& in, int phases, int cap_hint,
                            cudaStream_t st) {
    const bool limited = cap_hint > 0 && cap_hint < b.max_lm;
    const EkfLaunch L = make_launch(b, limited ? cap_hint : b.max_lm, 0);
    const size_t smem = ekf_smem_carve(b, L, nullptr, nullptr);
    if (limited) {
        cudaError_t e = cudaMemsetAsync(b.retry_count, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
    }
    // small tiles: 128-thread CTAs (the O(n^2) sweep is short and more CTAs fit per SM)
    if (L.n_cap <= 67) ekf_step_kernel<128><<<b.batch, 128, smem, st>>>(b, fc, in, phases, L);
    else ekf_step_kernel<256><<<b.batch, 256, smem, st>>>(b, fc, in, phases, L);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !limited) return e;
    const EkfLaunch R = make_launch(b, b.max_lm, 1);
    const int grid = b.batch < 296 ? b.batch : 296;
    ekf_step_kernel<256><<<grid, 256, ekf_smem_carve(b, R, nullptr, nullptr), st>>>(b, fc, in, phases, R);
    return cudaGetLastError();
}

}  // namespace slam
