// ekf_batch.cu -- batched EKF-SLAM for sm_100a: one CTA per filter instance, the covariance staged in shared memory
// by 1-D bulk async copies (cp.async.bulk / mbarrier).  Two kernels share one core:
//   ekf_step_kernel   one reference EKF::update per launch: P crosses HBM once each way per step (HBM bound);
//   ekf_sweep_kernel  a chunk of a Monte-Carlo sweep (simulator -> filter -> error terms) or of a trajectory replay
//                     (uploaded messages -> filter -> poses), T steps per launch with P RESIDENT in shared memory.
//
// Restates EKF::update, ekf_ws/src/localization_pkg/src/ekf.cpp:37-179, with the reference's float/double roundings
//  and evaluates its dense products structurally on the PACKED SYMMETRIC covariance
// (common.cuh: bpl_idx; P = P^T up to rounding in the reference -- its asymmetry stays at 1e-15 |P|, SURVEY App. E --
// so only the lower triangle is kept: half the shared memory, half the HBM bytes, half the rank-2 flops):
//   :61   F_x P F_x^T + F_v V F_v^T   -> column entries (j,0),(j,1) pick up (j,2), + 3x3 block   O(n)
//   :133  H P H^T + W                 -> the 5x5 sub-block of P that H touches                   O(1)
//   :135  P H^T S^-1                  -> (H P)^T S^-1: the same five rows of P as H P            O(n)
//   :140  P - (K H) P                 -> rank-2 update of the lower triangle, 2x2 blocks         O(n^2 / 2)
//   :172  Y blkdiag(P,W) Y^T          -> one new block row                                      O(n)
//
// Thread organisation (NT = 32 * WARPS filter threads, WARPS in {1,2,4,8,16} chosen per launch from the tile size so
// that as many instances as possible are resident per SM: the step is a chain of scalar latencies -- sqrt, divisions,
// atan2, sincos -- and instance-level parallelism is what hides it):
//   * warp 0 evaluates the H / S / S^-1 chain of a landmark update ONCE (lane-parallel divisions) while lane 0 of the
//     last warp evaluates the innovation (atan2 / remainder) and, at step start, the predict trigonometry;
//   * known-ID association (integer compares on lm_IDs) is resolved for the whole message before P is touched;
//   * in the sweep kernel a dedicated producer warp runs the simulator (or fetches the uploaded messages) one step
//     ahead of the filter warps and folds the error terms (or writes the poses) one step behind them.
#include "sim_device.cuh"

#include <climits>
#include <cstdio>
#include <cstdlib>

namespace slam {

// Launch geometry.  Shared memory is sized for the landmark capacity of THIS launch (cap_lm <= max_lm), chosen
// by the host from a slightly stale device-side max(M) plus headroom, so that while the map is still small many
// more CTAs fit per SM.  An instance that could outgrow the tile is left untouched for a full-capacity launch
// (step kernel: retry list; sweep kernel: progress gate), so results never depend on the hint.
struct EkfLaunch {
    int cap_lm;        // landmark capacity of this launch's shared-memory tile
    int n_cap;         // 3 + 2*cap_lm
    int ps2;           // shared-memory plane stride (doubles) of the packed covariance
    int from_list;     // 0: instance = blockIdx.x ; 1: instances come from b.retry_list (persistent loop)
    int off[16];       // shared-memory byte offsets (laid out on the host: the kernels add constant-bank offsets)
    int smem_bytes;    // step kernel's dynamic shared memory
    int sweep_bytes;   // sweep kernel's (step layout + message double buffer + pose snapshots)
};

struct EkfSmem {
    double* P;      // packed symmetric covariance, two planes of ps2 doubles
    double* x;      // running x_pred (unpadded)
    double* xs;     // x_t at step start (stale landmark means, ekf.cpp:115)
    double* H0;     // row 0 of H_x P_pred, padded index (H0[0] = 0)
    double* H1;     // row 1
    double* K;      // K as double2 per padded row (K[0] = 0)
    double* sc;     // scalars
    int* ids;
    float* meas;
    int* assoc;
    int* iscr;      // [0] dead flag of the association pre-pass, [1] nan flag, [2] work item, [3] tile overflow flag
    uint64_t* bar;
};

enum { EO_P = 0, EO_X, EO_XS, EO_H0, EO_H1, EO_K, EO_SC, EO_IDS, EO_MEAS, EO_ASSOC, EO_ISCR, EO_BAR, EO_WM0, EO_WM1, EO_WNM, EO_WSNAP };

static void ekf_smem_layout(const int max_meas, EkfLaunch& L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
    const int np = L.n_cap + 1;                 // padded state size (even)
    L.off[EO_P] = take(sizeof(double) * 2 * (size_t)L.ps2);
    L.off[EO_X] = take(sizeof(double) * np);
    L.off[EO_XS] = take(sizeof(double) * np);
    L.off[EO_H0] = take(sizeof(double) * np);
    L.off[EO_H1] = take(sizeof(double) * np);
    L.off[EO_K] = take(sizeof(double) * 2 * np);
    L.off[EO_SC] = take(sizeof(double) * 24);
    L.off[EO_IDS] = take(sizeof(int) * (L.cap_lm + 1));
    L.off[EO_MEAS] = take(sizeof(float) * 3 * max_meas);
    L.off[EO_ASSOC] = take(sizeof(int) * max_meas);
    L.off[EO_ISCR] = take(sizeof(int) * 8);
    L.off[EO_BAR] = take(sizeof(uint64_t));
    L.smem_bytes = (int)off;
    L.off[EO_WM0] = take(sizeof(float) * 3 * max_meas);
    L.off[EO_WM1] = take(sizeof(float) * 3 * max_meas);
    L.off[EO_WNM] = take(sizeof(int) * 4);
    L.off[EO_WSNAP] = take(sizeof(double) * 24);
    L.sweep_bytes = (int)off;
}

__device__ __forceinline__ void ekf_smem_bind(const EkfLaunch& L, unsigned char* base, EkfSmem& s) {
    s.P = (double*)(base + L.off[EO_P]); s.x = (double*)(base + L.off[EO_X]); s.xs = (double*)(base + L.off[EO_XS]);
    s.H0 = (double*)(base + L.off[EO_H0]); s.H1 = (double*)(base + L.off[EO_H1]); s.K = (double*)(base + L.off[EO_K]);
    s.sc = (double*)(base + L.off[EO_SC]);
    s.ids = (int*)(base + L.off[EO_IDS]); s.meas = (float*)(base + L.off[EO_MEAS]); s.assoc = (int*)(base + L.off[EO_ASSOC]);
    s.iscr = (int*)(base + L.off[EO_ISCR]); s.bar = (uint64_t*)(base + L.off[EO_BAR]);
}

// scalar slots in sc[]
enum { SC_FA = 0, SC_FB, SC_C, SC_S, SC_NX0, SC_NX1, SC_NX2, SC_NU0, SC_NU1, SC_XD, SC_YD, SC_CB, SC_SB,
       SC_Q0 = 13 /* 4 quotients of H */, SC_I00 = 17 /* S^-1, 4 */, SC_END = 21 };
enum { IS_DEAD = 0, IS_NAN = 1, IS_WORK = 2, IS_OVER = 3 };
enum { ASSOC_NEW = -1, ASSOC_DROPPED = -2 };   // internal codes of the pre-pass; both read back as -1 (new landmark)

// barrier among the NT threads that run the filter core
template <int NT>
struct CtaSync {            // the whole CTA runs the core (ekf_step_kernel)
    static __device__ __forceinline__ void sync() { if constexpr (NT == 32) __syncwarp(); else __syncthreads(); }
};
// Named barriers carry their id and thread count as IMMEDIATES: with a register operand ptxas cannot tell which of the 16
// hardware barriers a CTA uses and reserves all of them, and the SM's barrier file then caps the residency at 4 CTAs per SM
// whatever the tile size (ncu: launch__barrier_count 16, launch__occupancy_limit_barriers 4 -- what held the sweep kernel at
// 4 instances per SM while the map was still small).
template <int NT, int ID>
struct NamedSync {          // a subset of the CTA's warps runs the core (ekf_sweep_kernel consumers)
    static __device__ __forceinline__ void sync() {
        if constexpr (NT == 32) __syncwarp(); else asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NT) : "memory");
    }
};
template <int ID, int COUNT>
__device__ __forceinline__ void named_sync_i() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT>
__device__ __forceinline__ void named_arrive_i() {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
// barrier ID0 + p, p in {0, 1} known at run time only
template <int ID0, int COUNT>
__device__ __forceinline__ void named_sync(const int p) { if (p) named_sync_i<ID0 + 1, COUNT>(); else named_sync_i<ID0, COUNT>(); }
template <int ID0, int COUNT>
__device__ __forceinline__ void named_arrive(const int p) { if (p) named_arrive_i<ID0 + 1, COUNT>(); else named_arrive_i<ID0, COUNT>(); }

// ---- predict scalars (ekf.cpp:43-59), one thread: F_x(0,2), F_x(1,2), cos, sin and the new vehicle pose
__device__ __forceinline__ void ekf_predict_scalars(const FilterConst& fc, const EkfSmem& s, const float d_d, const float d_th) {
    double sn, c;
    const double th = s.x[2];
    sincos(th, &sn, &c);
    s.sc[SC_FA] = (double)(-1 * d_d) * sn;        // F_x(0,2), :48
    s.sc[SC_FB] = (double)d_d * c;                // F_x(1,2), :49
    s.sc[SC_C] = c; s.sc[SC_S] = sn;
    const float dv = d_d + fc.v_d;                // float add, :57-58
    s.sc[SC_NX0] = s.x[0] + (double)dv * c;
    s.sc[SC_NX1] = s.x[1] + (double)dv * sn;
    s.sc[SC_NX2] = wrap_2pi(th + (double)d_th + (double)fc.v_th);   // :59
}

// ---- known-ID association for the whole message (ekf.cpp:99-109), one warp, before anything is modified.
// s.assoc[l] <- slot in the committed lm_IDs, ASSOC_NEW (will be inserted) or ASSOC_DROPPED (max_lm reached).
// A measurement that repeats the id of a landmark inserted earlier in the same step would make the reference index
// x_t out of range (:115): flagged dead (SLAM_STATUS_SAME_STEP_REMATCH), nothing of the step is applied.
// s.iscr[IS_OVER] <- 1 when the insertions of this message would not fit a tile of tile_lm landmarks.
__device__ __forceinline__ void ekf_assoc_prepass(const EkfSmem& s, const int lane, const int M, const int nm, const int max_lm,
                                                  const int tile_lm) {
    int M_run = M;
    bool dead = false;
#pragma unroll 1
    for (int l = 0; l < nm; ++l) {
        const int id = (int)s.meas[3 * l];                                 // :101
        int cand = INT_MAX;
#pragma unroll 1
        for (int j = lane; j < M; j += 32)
            if (s.ids[j] == id) { cand = j; break; }
        cand = __reduce_min_sync(0xffffffffu, cand);                       // first match in ascending slot order
        int code = cand;
        if (cand == INT_MAX) {
            bool dup = false;
#pragma unroll 1
            for (int q = lane; q < l; q += 32) dup |= (s.assoc[q] == ASSOC_NEW) && ((int)s.meas[3 * q] == id);
            if (__any_sync(0xffffffffu, dup)) { dead = true; break; }
            if (M_run < max_lm) { code = ASSOC_NEW; ++M_run; } else code = ASSOC_DROPPED;
        }
        if (lane == 0) s.assoc[l] = code;
        __syncwarp();
    }
    if (lane == 0) { s.iscr[IS_DEAD] = dead ? 1 : 0; s.iscr[IS_OVER] = (M_run > tile_lm) ? 1 : 0; }
}

// ---- P -= K (H P) (ekf.cpp:140) on the packed lower triangle.  A plane is a flat array of 16-byte words: block (a, bc),
// bc <= a, of plane 0 / 1 (upper / lower row of the 2x2 block) is word a (a + 1) / 2 + bc.  Every block is updated by the one
// expression below, whatever the walk: the results do not depend on the walk or on the CTA width.
__device__ __forceinline__ void ekf_rank2_block(double2& p0, double2& p1, const double2 klo, const double2 khi, const double2 h0,
                                                const double2 h1) {
    p0.x = p0.x - (klo.x * h0.x + klo.y * h1.x);
    p0.y = p0.y - (klo.x * h0.y + klo.y * h1.y);
    p1.x = p1.x - (khi.x * h0.x + khi.y * h1.x);
    p1.y = p1.y - (khi.x * h0.y + khi.y * h1.y);
}
// one row of K = (H P)^T S^-1: (h0 i00 + h1 i10, h0 i01 + h1 i11), the roundings fixed by explicit intrinsics
__device__ __forceinline__ double2 ekf_gain_row(const double h0, const double h1, const double (&si)[4]) {
    return make_double2(__fma_rn(h0, si[0], __dmul_rn(h1, si[2])), __fma_rn(h0, si[1], __dmul_rn(h1, si[3])));
}
// Flat walk: thread t takes the words t, t + NT, ... of the triangle in STORAGE order, so the lanes of a warp always touch 32
// consecutive 16-byte words of each plane (no bank conflicts, also where a warp straddles block rows -- the earlier walk over
// paired block rows (A-1-q, q) split a warp into runs at unrelated addresses: 17 % excess shared-memory wavefronts at the
// 50-landmark tile, where that pipe is 74 % busy).  The block row of word t is a = floor((sqrt(8 t + 1) - 1) / 2); with 1.5 in
// place of 1 under an APPROXIMATE square root (MUFU.SQRT, relative error ~2^-22) the truncation is exact for every t the
// batched tiles can hold: a row's first word has sqrt(8 t + 1) = 2 a + 1 exactly and the half pushes it ~1e-3 above the
// integer, a row's last word stays ~0.017 below the next one (checked exhaustively on the host for A <= 120).  Every block
// fetches K of its row (a broadcast within the row) and H P of its column: 128 bytes of shared-memory traffic for 12 FP64
// instructions (80 % of the kernel's shared-memory wavefronts at the full map).
// Measured and rejected (B200, configs[1], same launches otherwise): 2 x 2 super-blocks per thread (half the
// operand loads and index arithmetic, but 32-byte lane strides = two-way bank conflicts: -18 %), column strips with H P in
// registers and one K broadcast per block row (-13 %; with the next row's loads issued ahead of the stores -30 %: a warp
// per block row serialises A / warps row trips), two blocks in flight per trip (-5 %; as vertical pairs (a, bc), (a + 1, bc) that
// share the H P operand on top of the storage-order walk: -3 %).  All three also slowed the
// launches whose rank-2 pass is negligible (8-landmark tiles): the kernel's hot path is ~60 KB of SASS walked by 4-12 CTAs per
// SM in different phases, and what grows it pays in instruction fetch (stall_no_instruction 0.6 warps per issue).
__device__ __forceinline__ float sqrt_approx(const float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int NT>
__device__ __forceinline__ void ekf_rank2_flat(const EkfSmem& s, const int ps2, const int A /* live block rows */) {
    const int words = (A * (A + 1)) >> 1;           // 16-byte words per plane
    const double2* K2 = reinterpret_cast<const double2*>(s.K);
    const double2* H02 = reinterpret_cast<const double2*>(s.H0);
    const double2* H12 = reinterpret_cast<const double2*>(s.H1);
    double2* P0 = reinterpret_cast<double2*>(s.P);
    double2* P1 = reinterpret_cast<double2*>(s.P + ps2);
    for (int t = threadIdx.x; t < words; t += NT) {
        const int a = __float2int_rz((sqrt_approx((float)(8 * t) + 1.5f) - 1.0f) * 0.5f);
        const int bc = t - ((a * (a + 1)) >> 1);
        const double2 klo = K2[2 * a], khi = K2[2 * a + 1];
        const double2 h0 = H02[bc], h1 = H12[bc];
        double2 p0 = P0[t], p1 = P1[t];
        ekf_rank2_block(p0, p1, klo, khi, h0, h1);
        P0[t] = p0; P1[t] = p1;
    }
}

// ---- one reference EKF::update on the shared-memory-resident filter, executed by NT threads (threadIdx.x < NT)
// that synchronise through Sync.  On entry (all visible): s.P / s.x / s.xs (= s.x) / s.ids hold the committed filter
// with M landmarks, s.meas the message, s.sc the predict scalars, and -- known-ID mode -- s.assoc / s.iscr[IS_DEAD]
// the pre-pass.  Returns true when the step was applied; false when the instance died (same-step re-match): in
// known-ID mode nothing has been modified then, in unknown-ID mode the shared-memory state must be discarded.
template <int NT, class Sync, bool KNOWN_ONLY = false>
__device__ __forceinline__ bool ekf_core(const FilterConst& fc, const EkfSmem& s, const int ps2, const int max_lm,
                                         const int phases, int& M, const int nm, int& status, int& n_upd) {
    constexpr int WARPS = NT / 32;
    constexpr int NUW = WARPS - 1;                              // lane 0 of this warp owns the atan2 / sincos chains
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool nu_thread = (warp == NUW) && (lane == 0);
    const int M_start = M;
    n_upd = 0;
    // (the sweep kernel runs known-ID batches only: KNOWN_ONLY compiles the box gate and its trigonometry out of its hot path)
    const bool id_known = KNOWN_ONLY || fc.id_known;

    if ((phases & STEP_UPDATE) && id_known && s.iscr[IS_DEAD]) { status |= SLAM_STATUS_SAME_STEP_REMATCH; return false; }

    // ---- PREDICT, ekf.cpp:43-61.  T = F_x P (rows 0,1 pick up row 2), P' = T F_x^T (cols 0,1 pick up col 2)
    //      + (F_v V) F_v^T on the vehicle block.  In the lower triangle only the column form exists for j >= 3:
    //      (j,0) += (j,2) fa, (j,1) += (j,2) fb; the 3x3 vehicle block is done by one thread in the reference's order.
    if (phases & STEP_PREDICT) {
        const double fa = s.sc[SC_FA], fb = s.sc[SC_FB];
        const int np = 4 + 2 * M;
        for (int j = tid; j < np; j += NT) {
            if (j >= 4) {
                double* rb = s.P + bpl_idx(j, 0, ps2);          // padded row j: [0, (j,x), (j,y), (j,yaw), ...]
                double2 b0 = *reinterpret_cast<double2*>(rb), b1 = *reinterpret_cast<double2*>(rb + 2);
                const double t2 = b1.y;
                b0.y = b0.y + t2 * fa;
                b1.x = b1.x + t2 * fb;
                *reinterpret_cast<double2*>(rb) = b0; *reinterpret_cast<double2*>(rb + 2) = b1;
            } else if (j == 0) {
                const double c = s.sc[SC_C], sn = s.sc[SC_S];
                double Pv[3][3], T[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int q = 0; q <= a; ++q) { const double v = s.P[bpl_idx(a + 1, q + 1, ps2)]; Pv[a][q] = v; Pv[q][a] = v; }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double p2 = Pv[2][q];
                    T[0][q] = Pv[0][q] + fa * p2;
                    T[1][q] = Pv[1][q] + fb * p2;
                    T[2][q] = p2;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double t2 = T[i][2];
                    double p0 = T[i][0] + t2 * fa;
                    double p1 = T[i][1] + t2 * fb;
                    if (i == 0) { const double cV = c * fc.V00; p0 += cV * c; p1 += cV * sn; }
                    if (i == 1) { const double sV = sn * fc.V00; p0 += sV * c; p1 += sV * sn; }
                    s.P[bpl_idx(i + 1, 1, ps2)] = p0;
                    if (i >= 1) s.P[bpl_idx(i + 1, 2, ps2)] = p1;
                    if (i == 2) s.P[bpl_idx(3, 3, ps2)] = t2 + fc.V11;
                }
                s.x[0] = s.sc[SC_NX0]; s.x[1] = s.sc[SC_NX1]; s.x[2] = s.sc[SC_NX2];
            }
        }
        Sync::sync();
    }

    // ---- UPDATE, ekf.cpp:63-174
    for (int l = 0; l < nm; ++l) {
        const float r = s.meas[3 * l + 1], bb = s.meas[3 * l + 2];
        const int n = 3 + 2 * M;
        int slot, id;
        if (id_known) {
            slot = s.assoc[l];
            id = (int)s.meas[3 * l];
            if (slot == ASSOC_DROPPED) { status |= SLAM_STATUS_CAPACITY; continue; }
        } else {
            // -- unknown IDs, :82-98: first landmark inside the float box gate around the detection
            id = M;
            if (nu_thread) {
                double sa, ca; sincos(s.x[2] + (double)bb, &sa, &ca);
                s.sc[SC_XD] = (double)(float)(s.x[0] + (double)r * ca);    // float x_detected, :87
                s.sc[SC_YD] = (double)(float)(s.x[1] + (double)r * sa);    // float y_detected, :88
                s.sc[SC_CB] = ca; s.sc[SC_SB] = sa;                        // reused by the insertion below
            }
            Sync::sync();
            const double xd = s.sc[SC_XD], yd = s.sc[SC_YD];
            int cand = INT_MAX;
            for (int j = lane; j < M; j += 32) {                           // every warp scans the whole map: no barrier
                const float x_diff = (float)fabs(xd - s.x[3 + 2 * j]);     // :91
                const float y_diff = (float)fabs(yd - s.x[4 + 2 * j]);     // :92
                if (x_diff < fc.min_sep && y_diff < fc.min_sep) { cand = j; break; }
            }
            cand = __reduce_min_sync(0xffffffffu, cand);                   // first j in ascending order, :93-97
            slot = (cand == INT_MAX) ? ASSOC_NEW : cand;
            if (slot != ASSOC_NEW) id = slot;
            if (slot == ASSOC_NEW && M >= max_lm) {
                status |= SLAM_STATUS_CAPACITY;
                if (tid == 0) s.assoc[l] = ASSOC_NEW;
                Sync::sync();             // sc[] is rewritten by the next measurement
                continue;
            }
            if (tid == 0) s.assoc[l] = slot;
            if (slot >= M_start) { status |= SLAM_STATUS_SAME_STEP_REMATCH; return false; }   // :115 reads x_t out of range
        }

        if (slot >= 0) {
            // -------- landmark update, :110-140
            const int i = slot * 2 + 3;          // unpadded state index of the landmark
            const int ip = i + 1;                // padded
            ++n_upd;
            // -- scalar phase: warp 0 -> H (4 distinct quotients), S, S^-1 ; nu thread -> innovation
            if (warp == 0 || nu_thread) {
                // landmark from the stale x_t, vehicle from the running x_pred (:115)
                const double dx = s.xs[i] - s.x[0], dy = s.xs[i + 1] - s.x[1];
                const float dist = (float)sqrt(dx * dx + dy * dy);         // :115
                if (nu_thread) {
                    const float ang = (float)wrap_2pi(atan2(dy, dx) - s.x[2]);   // :129
                    s.sc[SC_NU0] = (double)(r - dist - fc.w_r);            // all-float arithmetic, :130
                    s.sc[SC_NU1] = (double)(bb - ang - fc.w_b);            // :131
                }
                if (warp == 0) {
                    // the four distinct quotients of H_x (:118-126), one per lane, shuffled to the warp
                    const double dd = (double)dist;
                    const double d2 = (double)(dist * dist);               // float product, :120
                    const int ql = lane & 3;
                    const double qv = ((ql == 0 || ql == 3) ? dx : dy) / ((ql < 2) ? dd : d2);
                    const double q0 = __shfl_sync(0xffffffffu, qv, 0);     // dx / dist
                    const double q1 = __shfl_sync(0xffffffffu, qv, 1);     // dy / dist
                    const double q2 = __shfl_sync(0xffffffffu, qv, 2);     // dy / dist^2
                    const double q3 = __shfl_sync(0xffffffffu, qv, 3);     // dx / dist^2
                    const double H[10] = {-q0, -q1, 0.0, q0, q1, q2, -q3, -1.0, -q2, q3};
                    // S = H P H^T + W (:133) from the 5x5 sub-block of P that H touches: lane c < 5 forms column
                    // hc[c] of H P, the 2x2 sums are reduced over the 5 lanes in ascending c order
                    const int hcl = (lane < 3) ? lane + 1 : ip + (lane - 3);    // padded hc[lane] for lane < 5
                    double g0 = 0.0, g1 = 0.0;
                    if (lane < 5) {
                        const double p0 = s.P[bpl_sym(1, hcl, ps2)], p1 = s.P[bpl_sym(2, hcl, ps2)], p2 = s.P[bpl_sym(3, hcl, ps2)];
                        const double p3 = s.P[bpl_sym(ip, hcl, ps2)], p4 = s.P[bpl_sym(ip + 1, hcl, ps2)];
                        g0 = H[0] * p0; g0 += H[1] * p1; g0 += H[2] * p2; g0 += H[3] * p3; g0 += H[4] * p4;
                        g1 = H[5] * p0; g1 += H[6] * p1; g1 += H[7] * p2; g1 += H[8] * p3; g1 += H[9] * p4;
                    }
                    double S00 = 0, S01 = 0, S10 = 0, S11 = 0;
#pragma unroll
                    for (int c = 0; c < 5; ++c) {
                        const double a0 = __shfl_sync(0xffffffffu, g0, c), a1 = __shfl_sync(0xffffffffu, g1, c);
                        S00 += a0 * H[c]; S01 += a0 * H[5 + c]; S10 += a1 * H[c]; S11 += a1 * H[5 + c];
                    }
                    S00 += fc.W00; S11 += fc.W11;
                    // S^-1 (:135).  Eigen's dynamic inverse() is a partial-pivot LU (three dependent divisions); S = H P H^T
                    // + W is a well-conditioned 2x2, so the adjugate form -- ONE division on the critical path of every
                    // landmark update -- agrees with it to a few ulp (the oracle keeps the LU)
                    const double idet = 1.0 / (S00 * S11 - S01 * S10);
                    if (lane == 0) {
                        s.sc[SC_I00] = S11 * idet; s.sc[SC_I00 + 1] = -S01 * idet;
                        s.sc[SC_I00 + 2] = -S10 * idet; s.sc[SC_I00 + 3] = S00 * idet;
                    }
                    if (lane == 0) { s.sc[SC_Q0] = q0; s.sc[SC_Q0 + 1] = q1; s.sc[SC_Q0 + 2] = q2; s.sc[SC_Q0 + 3] = q3; }
                }
            }
            Sync::sync();
            const double si[4] = {s.sc[SC_I00], s.sc[SC_I00 + 1], s.sc[SC_I00 + 2], s.sc[SC_I00 + 3]};     // S^-1
            // -- O(n) phase: H P (2 x n, kept for the sweep) and K = P H^T S^-1 = (H P)^T S^-1 (n x 2)
            {
                const double q0 = s.sc[SC_Q0], q1 = s.sc[SC_Q0 + 1], q2 = s.sc[SC_Q0 + 2], q3 = s.sc[SC_Q0 + 3];
                const double H[10] = {-q0, -q1, 0.0, q0, q1, q2, -q3, -1.0, -q2, q3};
                const int np = n + 1;
                for (int j = tid; j < np; j += NT) {
                    double h0 = 0.0, h1 = 0.0;
                    if (j >= 1) {
                        const double p0 = s.P[bpl_sym(1, j, ps2)], p1 = s.P[bpl_sym(2, j, ps2)], p2 = s.P[bpl_sym(3, j, ps2)];
                        const double p3 = s.P[bpl_sym(ip, j, ps2)], p4 = s.P[bpl_sym(ip + 1, j, ps2)];
                        h0 = H[0] * p0; h0 += H[1] * p1; h0 += H[2] * p2; h0 += H[3] * p3; h0 += H[4] * p4;
                        h1 = H[5] * p0; h1 += H[6] * p1; h1 += H[7] * p2; h1 += H[8] * p3; h1 += H[9] * p4;
                    }
                    s.H0[j] = h0; s.H1[j] = h1;
                    const double2 kj = ekf_gain_row(h0, h1, si);
                    s.K[2 * j] = kj.x; s.K[2 * j + 1] = kj.y;
                }
            }
            Sync::sync();
            // -- x_pred += K nu (:138), yaw wrapped (:139)
            {
                const double nu0 = s.sc[SC_NU0], nu1 = s.sc[SC_NU1];
                for (int q = tid; q < n; q += NT) {
                    double xv = s.x[q] + (s.K[2 * q + 2] * nu0 + s.K[2 * q + 3] * nu1);
                    if (q == 2) xv = wrap_2pi(xv);
                    s.x[q] = xv;
                }
            }
            // -- P -= K (H P), :140, on the lower triangle
            ekf_rank2_flat<NT>(s, ps2, 2 + M);
            Sync::sync();
        } else {
            // -------- landmark insertion, :141-173: one new block row (padded rows np, np+1)
            if (id_known) {
                if (nu_thread) {
                    double sb, cb; sincos(s.x[2] + (double)bb, &sb, &cb);
                    s.sc[SC_CB] = cb; s.sc[SC_SB] = sb;
                }
                Sync::sync();
            }
            const double cb = s.sc[SC_CB], sb = s.sc[SC_SB];
            const double g02 = -(double)r * sb, g12 = (double)r * cb;      // G_x(0,2), G_x(1,2), :162,165
            const int np = n + 1;
            double* r0 = s.P + bpl_idx(np, 0, ps2);
            double* r1 = s.P + bpl_idx(np + 1, 0, ps2);
            for (int j = tid; j < np; j += NT) {
                double t0 = 0.0, t1 = 0.0;                                 // phantom column 0 stays zero
                if (j >= 1) {
                    const double p0 = s.P[bpl_sym(1, j, ps2)], p1 = s.P[bpl_sym(2, j, ps2)], p2 = s.P[bpl_sym(3, j, ps2)];
                    t0 = 1.0 * p0; t0 += 0.0 * p1; t0 += g02 * p2;
                    t1 = 0.0 * p0; t1 += 1.0 * p1; t1 += g12 * p2;
                }
                r0[j] = t0; r1[j] = t1;
            }
            if (nu_thread) {
                // new 2x2 block: G_x P_vv G_x^T + G_z W G_z^T, :155-172
                const double gx[2][3] = {{1.0, 0.0, g02}, {0.0, 1.0, g12}};
                const double gz[2][2] = {{cb, -(double)r * sb}, {sb, (double)r * cb}};
                const double Wm[2][2] = {{fc.W00, 0.0}, {0.0, fc.W11}};
                double T3[2][3], T2[2][2];
                for (int rr = 0; rr < 2; ++rr) {
                    for (int k = 0; k < 3; ++k) {
                        double t = gx[rr][0] * s.P[bpl_sym(1, k + 1, ps2)]; t += gx[rr][1] * s.P[bpl_sym(2, k + 1, ps2)];
                        t += gx[rr][2] * s.P[bpl_sym(3, k + 1, ps2)];
                        T3[rr][k] = t;
                    }
                    for (int c2 = 0; c2 < 2; ++c2) T2[rr][c2] = gz[rr][0] * Wm[0][c2] + gz[rr][1] * Wm[1][c2];
                }
                for (int rr = 0; rr < 2; ++rr)
                    for (int c2 = 0; c2 < 2; ++c2) {
                        double t = T3[rr][0] * gx[c2][0]; t += T3[rr][1] * gx[c2][1]; t += T3[rr][2] * gx[c2][2];
                        t += T2[rr][0] * gz[c2][0]; t += T2[rr][1] * gz[c2][1];
                        (rr ? r1 : r0)[np + c2] = t;
                    }
                s.x[n] = s.x[0] + (double)r * cb;                          // :147
                s.x[n + 1] = s.x[1] + (double)r * sb;                      // :148
                s.ids[M] = id;                                             // :150
            }
            M += 1;
            Sync::sync();
        }
    }
    return true;
}

// algorithmic work of one update, using the live n at the end of the step
// w[4]: flops the kernel executes for the step's rank-2 updates: the lower triangle only, 16 per 2x2 block = ~2 k n^2
__device__ __forceinline__ void ekf_work_terms(const int n, const int nm, const int n_upd, double (&w)[4]) {
    const double nd = (double)n;
    w[0] += 16.0 * nd * nd + 16.0 * nd + 12.0 * nm + 8.0;
    w[1] += 4.0 * (double)n_upd * nd * nd;
    w[2] += nd;
    w[3] += (double)nm;
}
// bytes one pass of the packed covariance over HBM moves (one direction): two planes of A (A + 1) doubles, A = 2 + M
__device__ __forceinline__ double ekf_plane_bytes(const int M) { return 16.0 * (double)bpl_plane_doubles(2 + M); }

// stage the packed covariance of an instance with M landmarks: one bulk copy per plane, both on one mbarrier
__device__ __forceinline__ void ekf_load_P(const BatchState& b, const EkfSmem& s, const int ps2, const double* gP, const int M) {
    const uint32_t bytes = (uint32_t)(bpl_plane_doubles(2 + M) * sizeof(double));
    mbar_expect_tx(s.bar, 2 * bytes);
    bulk_g2s(s.P, gP, bytes, s.bar);
    bulk_g2s(s.P + ps2, gP + b.ps2g, bytes, s.bar);
}
__device__ __forceinline__ void ekf_store_P(const BatchState& b, const EkfSmem& s, const int ps2, double* gP, const int M) {
    const uint32_t bytes = (uint32_t)(bpl_plane_doubles(2 + M) * sizeof(double));
    bulk_s2g(gP, s.P, bytes);
    bulk_s2g(gP + b.ps2g, s.P + ps2, bytes);
    bulk_commit();
}

// ------------------------------------------------------------------------------------------------------------
// ekf_step_kernel: one reference EKF::update per instance per launch.  Returns true when the mbarrier phase
// `parity` was consumed.
// ------------------------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ bool ekf_instance(const BatchState& b, const FilterConst& fc, const StepInputs& in,
                                             const int phases, const EkfLaunch& L, const EkfSmem& s, const int inst,
                                             const uint32_t parity) {
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ps2 = L.ps2;

    const int4 meta_in = b.meta[inst];
    int nm = (phases & STEP_UPDATE) ? in.n_meas[inst] : 0;
    int status = meta_in.y;
    if (status & SLAM_STATUS_SAME_STEP_REMATCH) {               // the reference process is dead past this point
        if (in.poses && tid < 3) in.poses[3 * (size_t)inst + tid] = b.x[(size_t)inst * b.x_stride + tid];   // frozen pose
        return false;
    }
    int M = meta_in.x;
    const int M_start = M;
    const int n0 = 3 + 2 * M;
    if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
    if (M + nm > L.cap_lm && L.cap_lm < b.max_lm) {
        // this launch's tile may be too small for the insertions of this step: defer, untouched
        if (tid == 0) b.retry_list[atomicAdd(b.retry_count, 1)] = inst;
        return false;
    }
    double* gP = b.P + (size_t)inst * b.p_stride;
    double* gx = b.x + (size_t)inst * b.x_stride;

    // ---- stage P: one bulk copy per plane
    if (tid == 0) { s.iscr[IS_NAN] = 0; s.iscr[IS_DEAD] = 0; ekf_load_P(b, s, ps2, gP, M); }
    // ---- meanwhile: state, ids, messages
    for (int i = tid; i < n0; i += THREADS) { const double v = gx[i]; s.x[i] = v; s.xs[i] = v; }
    for (int i = tid; i < M; i += THREADS) s.ids[i] = b.ids[(size_t)inst * b.max_lm + i];
    for (int i = tid; i < 3 * nm; i += THREADS) s.meas[i] = in.meas[(size_t)inst * b.max_meas * 3 + i];
    CtaSync<THREADS>::sync();
    // ---- scalar pre-work while P is in flight
    if ((phases & STEP_PREDICT) && warp == WARPS - 1 && lane == 0)
        ekf_predict_scalars(fc, s, in.fwd[in.cmd_stride ? inst : 0], in.ang[in.cmd_stride ? inst : 0]);
    if ((phases & STEP_UPDATE) && fc.id_known && warp == 0) ekf_assoc_prepass(s, lane, M, nm, b.max_lm, b.max_lm);
    if (warp == 0) mbar_wait(s.bar, parity);     // one warp polls the mbarrier; the others park at the CTA barrier
    CtaSync<THREADS>::sync();

    int n_upd = 0;
    const bool alive = ekf_core<THREADS, CtaSync<THREADS>>(fc, s, ps2, b.max_lm, phases, M, nm, status, n_upd);

    // ---- commit, :176-177
    if (!alive) {
        // frozen at the last committed state (global memory still holds it): only the status word changes
        if (tid == 0) b.meta[inst] = make_int4(meta_in.x, status, meta_in.z, 0);
        if (in.poses && tid < 3) in.poses[3 * (size_t)inst + tid] = gx[tid];
        return true;
    }
    const int n = 3 + 2 * M;
    for (int i = tid; i < n; i += THREADS) {
        const double v = s.x[i];
        gx[i] = v;
        if (!isfinite(v) || !isfinite(s.P[bpl_idx(i + 1, i + 1, ps2)])) s.iscr[IS_NAN] = 1;
    }
    for (int i = tid + M_start; i < M; i += THREADS) b.ids[(size_t)inst * b.max_lm + i] = s.ids[i];
    for (int i = tid; i < nm; i += THREADS) { const int a = s.assoc[i]; b.assoc[(size_t)inst * b.max_meas + i] = a < 0 ? -1 : a; }
    if (in.poses && tid < 3) in.poses[3 * (size_t)inst + tid] = s.x[tid];      // publishState's pose, fused (ekf.cpp:196-198)
    fence_proxy_async();     // generic-proxy writes of P must be visible to the bulk-copy engine
    CtaSync<THREADS>::sync();
    if (tid == 0) {
        ekf_store_P(b, s, ps2, gP, M);
        if (s.iscr[IS_NAN]) status |= SLAM_STATUS_NAN;
        b.meta[inst] = make_int4(M, status, meta_in.z + ((phases & STEP_PREDICT) ? 1 : 0),   // timestep, :39
                                 (phases & STEP_UPDATE) ? nm : meta_in.w);
        if (M > M_start) atomicMax(b.max_M, M);
        double w[4] = {0, 0, 0, 0};
        ekf_work_terms(n, nm, n_upd, w);
        double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
        st[8] += w[0]; st[9] += w[1]; st[10] += w[2]; st[11] += w[3]; st[13] += 0.5 * w[1];
        // bytes this launch really moved for the instance: packed P in and out, x in and out, ids, message, meta
        st[12] += ekf_plane_bytes(M_start) + ekf_plane_bytes(M) + 8.0 * (n0 + n) + 4.0 * M + 12.0 * nm + 4.0 * nm + 32.0;
        // shared memory may be reused / released once the bulk engine has READ it; global visibility of the writes
        // is guaranteed at kernel completion
        bulk_wait_read();
    }
    return true;
}

// resident CTAs the register allocation must allow: about 768 threads per SM (<= 80 registers per thread)
constexpr int step_min_blocks(int threads) { return threads >= 512 ? 1 : 768 / threads; }
constexpr int sweep_min_blocks(int cw) { return cw == 1 ? 12 : cw == 2 ? 8 : cw == 3 ? 6 : cw == 4 ? 4 : 2; }   // (cw == 4 at 6 blocks / 64 registers measured slower: 113.8 -> 101 M updates/s)

template <int THREADS>
__global__ void __launch_bounds__(THREADS, step_min_blocks(THREADS))
ekf_step_kernel(BatchState b, FilterConst fc, StepInputs in, int phases, EkfLaunch L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EkfSmem s;
    ekf_smem_bind(L, smem_raw, s);
    if (threadIdx.x == 0) { mbar_init(s.bar, 1); fence_mbar_init(); }
    CtaSync<THREADS>::sync();
    if (!L.from_list) {
        ekf_instance<THREADS>(b, fc, in, phases, L, s, (int)blockIdx.x, 0u);
        return;
    }
    // retry pass: a persistent grid drains the list of instances the capacity-limited launch deferred
    const int count = *b.retry_count;
    uint32_t parity = 0;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        if (ekf_instance<THREADS>(b, fc, in, phases, L, s, b.retry_list[q], parity)) parity ^= 1u;
        CtaSync<THREADS>::sync();
    }
    // the last CTA to finish re-arms the list for the next step (no memset on the stream) and posts the capacity hint into
    // mapped host memory (no copy, no event: max(M) only grows between resets, so a stale value is merely conservative)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(b.retry_count + 1, 1) == (int)gridDim.x - 1) {
            b.retry_count[0] = 0; b.retry_count[1] = 0;
            if (b.hint_host) { *(volatile int*)b.hint_host = *(volatile int*)b.max_M; __threadfence_system(); }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// ekf_sweep_kernel: a chunk of T consecutive reference steps of every instance in ONE launch with the filter
// resident in shared memory.  A persistent grid pulls instances from a work counter.  CW consumer warps run the
// filter; one producer warp runs AHEAD of them (messages double-buffered) and trails them by one step with the
// per-step outputs, so neither is on the filter's critical path:
//   REPLAY = false  Monte-Carlo sweep: simulator (sim_node.py:209-250) -> EKF::update -> error terms;
//   REPLAY = true   trajectory replay: uploaded [id,r,b] messages -> EKF::update -> pose estimates (publishState).
// Named barriers: FULL[p] producer -> consumers "message of a step with parity p is ready", DONE[p] consumers ->
// producer "step finished, snapshot[p] valid, message buffer p free".
// Chunking: only instances whose progress counter equals a.t0 run; an instance whose insertions would outgrow this
// launch's tile ABORTS the chunk untouched (nothing committed, progress unchanged) and is picked up by the next
// launch of the same chunk, which uses a larger tile.
// Known-ID mode only (the host falls back to per-step launches otherwise): the pre-pass association detects a
// same-step re-match before the step touches anything, so a dead instance stays at its committed state.
// ------------------------------------------------------------------------------------------------------------
// (BAR_CONS is the highest id: the one-warp variant synchronises its filter threads with __syncwarp and never names it, so it
// holds 5 hardware barriers instead of 6 and 12 instead of 10 of its CTAs fit the SM's barrier file)
enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_DONE0 = 3, BAR_DONE1 = 4, BAR_CONS = 5 };

struct SweepSmem {
    float* meas[2];     // [max_meas][3] message of step parity p
    int* nm;            // [0..1] detections of step parity p, [2] abort flag
    double* snap;       // [2][12] pose (3) + pose covariance (9) after the step of parity p
};

__device__ __forceinline__ void sweep_smem_bind(const EkfLaunch& L, unsigned char* base, SweepSmem& w) {
    w.meas[0] = (float*)(base + L.off[EO_WM0]); w.meas[1] = (float*)(base + L.off[EO_WM1]);
    w.nm = (int*)(base + L.off[EO_WNM]); w.snap = (double*)(base + L.off[EO_WSNAP]);
}

template <int CW, bool REPLAY>      // consumer warps; the CTA has CW + 1 warps
__global__ void __launch_bounds__(32 * (CW + 1), sweep_min_blocks(CW))
ekf_sweep_kernel(BatchState b, FilterConst fc, SimState sim, SimConst sc, SweepArgs a, EkfLaunch L) {
    constexpr int NT = 32 * CW;                 // filter threads
    constexpr int THREADS = NT + 32;
    using Sync = NamedSync<NT, BAR_CONS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EkfSmem s;
    SweepSmem w;
    ekf_smem_bind(L, smem_raw, s);
    sweep_smem_bind(L, smem_raw, w);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ps2 = L.ps2;
    const int T = a.T;
    const bool producer = warp == CW;
    if (tid == 0) { mbar_init(s.bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t parity = 0;

    while (true) {
        if (tid == 0) { s.iscr[IS_WORK] = atomicAdd(a.work_counter, 1); w.nm[2] = 0; }
        __syncthreads();
        const int inst = s.iscr[IS_WORK];
        if (inst >= b.batch) break;
        const int4 meta_in = b.meta[inst];
        if (a.progress[inst] != a.t0 || meta_in.x > L.cap_lm) { __syncthreads(); continue; }   // not this chunk / cannot fit
        const size_t cstep = a.cmd_stride ? (size_t)b.batch : 1, coff = a.cmd_stride ? (size_t)inst : 0;
        bool aborted = false;

        if (producer) {
            // ================= producer warp: messages ahead, per-step outputs behind =================
            double tr[3] = {0, 0, 0};
            const double* const lm_mine = REPLAY ? nullptr : sim.lm_xy + (size_t)inst * sim.lm_stride;   // this vehicle's map
            if (!REPLAY) { const double* t = sim.truth + 3 * (size_t)inst; tr[0] = t[0]; tr[1] = t[1]; tr[2] = t[2]; }
            double trp[2][3] = {{0, 0, 0}, {0, 0, 0}};   // truth of the step with parity p
            double eacc[6] = {0, 0, 0, 0, 0, 0};
            int overflow = 0, nm_last = 0;
            for (int t = 0; t <= T; ++t) {
                if (t < T) {
                    const int p = t & 1;
                    if (REPLAY) {
                        const size_t row = (size_t)t * b.batch + inst;
                        const int count = a.r_nmeas[row];
                        nm_last = count < b.max_meas ? count : b.max_meas;
                        const float* src = a.r_meas + row * (size_t)b.max_meas * 3;
                        for (int i = lane; i < 3 * nm_last; i += 32) w.meas[p][i] = src[i];
                        if (lane == 0) w.nm[p] = count;         // raw count: the consumers flag the overflow
                    } else {
                        const int count = sim_get_cmd_warp(lane, sc, lm_mine, sim.n_lm, b.max_meas, sim.k0, sim.k1,
                                                           sim.instance_offset + (uint32_t)inst, a.first_step + (uint32_t)t,
                                                           a.cmd_fwd[(size_t)t * cstep + coff], a.cmd_ang[(size_t)t * cstep + coff],
                                                           tr, w.meas[p]);
                        if (count > b.max_meas) overflow = 1;
                        nm_last = count < b.max_meas ? count : b.max_meas;
                        if (lane == 0) w.nm[p] = nm_last;
                        trp[p][0] = tr[0]; trp[p][1] = tr[1]; trp[p][2] = tr[2];
                    }
                    named_arrive<BAR_FULL0, THREADS>(p);
                }
                if (t >= 1) {
                    const int q = (t - 1) & 1;
                    named_sync<BAR_DONE0, THREADS>(q);          // step t-1 finished: snapshot[q] valid
                    if (w.nm[2]) { aborted = true; break; }
                    const double* sn = w.snap + 12 * q;
                    if (REPLAY) {
                        if (a.r_poses != nullptr && lane < 3) a.r_poses[((size_t)(t - 1) * b.batch + inst) * 3 + lane] = sn[lane];
                    } else if (lane == 0) {
                        double C[3][3];
                        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[r][c] = sn[3 + 3 * r + c];
                        pose_error_terms(sn[0] - trp[q][0], sn[1] - trp[q][1], wrap_2pi(sn[2] - trp[q][2]), C, eacc);
                    }
                }
            }
            if (!REPLAY && T > 0 && !aborted) {
                if (lane == 0) {
                    double* g = b.stats + (size_t)inst * SLAM_NUM_STATS;
                    for (int k = 0; k < 6; ++k) g[k] += eacc[k];
                    double* tg = sim.truth + 3 * (size_t)inst;
                    tg[0] = tr[0]; tg[1] = tr[1]; tg[2] = tr[2];
                    sim.n_meas[inst] = nm_last;
                    if (overflow) sim.overflow[inst] = 1;
                }
                const float* last = w.meas[(T - 1) & 1];
                for (int i = lane; i < 3 * nm_last; i += 32) sim.meas[(size_t)inst * b.max_meas * 3 + i] = last[i];
            }
        } else {
            // ================= consumer warps: the filter =================
            int status = meta_in.y;
            int M = meta_in.x;
            const int M_first = M;
            double* gP = b.P + (size_t)inst * b.p_stride;
            double* gx = b.x + (size_t)inst * b.x_stride;
            const int n0 = 3 + 2 * M;
            if (tid == 0) { s.iscr[IS_NAN] = 0; s.iscr[IS_DEAD] = 0; s.iscr[IS_OVER] = 0; ekf_load_P(b, s, ps2, gP, M); }
            for (int i = tid; i < n0; i += NT) s.x[i] = gx[i];
            for (int i = tid; i < M; i += NT) s.ids[i] = b.ids[(size_t)inst * b.max_lm + i];
            double wacc[4] = {0, 0, 0, 0};               // work counters, carried by thread 0
            int timestep = meta_in.z, nm = 0;
            if (warp == 0) mbar_wait(s.bar, parity);
            Sync::sync();

            for (int t = 0; t < T; ++t) {
                const int p = t & 1;
                const bool frozen = (status & SLAM_STATUS_SAME_STEP_REMATCH) != 0;
                named_sync<BAR_FULL0, THREADS>(p);              // message of step t is in w.meas[p]
                nm = w.nm[p];
                if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
                if (!frozen) {
                    s.meas = w.meas[p];
                    // ---- phase 0: association (warp 0) | predict trigonometry (last warp) | x_t snapshot
                    if (warp == 0) ekf_assoc_prepass(s, lane, M, nm, b.max_lm, L.cap_lm);
                    if (warp == CW - 1 && lane == 0)
                        ekf_predict_scalars(fc, s, a.cmd_fwd[(size_t)t * cstep + coff], a.cmd_ang[(size_t)t * cstep + coff]);
                    for (int i = tid; i < 3 + 2 * M; i += NT) s.xs[i] = s.x[i];
                    Sync::sync();
                    if (s.iscr[IS_OVER]) {
                        // the tile of this launch cannot take the insertions of this step: abort the chunk untouched
                        aborted = true;
                        if (tid == 0) w.nm[2] = 1;
                        named_arrive<BAR_DONE0, THREADS>(p);
                        if (t + 1 < T) named_sync<BAR_FULL0, THREADS>((t + 1) & 1);   // drain the message already under way
                        break;
                    }
                    int n_upd = 0;
                    if (ekf_core<NT, Sync, true>(fc, s, ps2, b.max_lm, STEP_PREDICT | STEP_UPDATE, M, nm, status, n_upd)) {
                        ++timestep;
                        if (tid == 0) ekf_work_terms(3 + 2 * M, nm, n_upd, wacc);
                    }
                }
                // pose snapshot for the per-step outputs (the producer consumes it while the next step runs)
                if (tid < 12) w.snap[12 * p + tid] = (tid < 3) ? s.x[tid] : s.P[bpl_sym((tid - 3) / 3 + 1, (tid - 3) % 3 + 1, ps2)];
                named_arrive<BAR_DONE0, THREADS>(p);
            }
            if (!aborted) {
                // ---- commit the instance
                const bool frozen = (status & SLAM_STATUS_SAME_STEP_REMATCH) != 0;
                const int n = 3 + 2 * M;
                for (int i = tid; i < n; i += NT) {
                    const double v = s.x[i];
                    gx[i] = v;
                    if (!isfinite(v) || !isfinite(s.P[bpl_idx(i + 1, i + 1, ps2)])) s.iscr[IS_NAN] = 1;
                }
                for (int i = tid + M_first; i < M; i += NT) b.ids[(size_t)inst * b.max_lm + i] = s.ids[i];
                if (T > 0 && !frozen)
                    for (int i = tid; i < nm; i += NT) { const int c = s.assoc[i]; b.assoc[(size_t)inst * b.max_meas + i] = c < 0 ? -1 : c; }
                fence_proxy_async();
                Sync::sync();
                if (tid == 0) {
                    ekf_store_P(b, s, ps2, gP, M);
                    if (s.iscr[IS_NAN]) status |= SLAM_STATUS_NAN;
                    b.meta[inst] = make_int4(M, status, timestep, (T > 0 && !frozen) ? nm : (frozen ? 0 : meta_in.w));
                    if (M > M_first) atomicMax(b.max_M, M);
                    double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
                    st[8] += wacc[0]; st[9] += wacc[1]; st[10] += wacc[2]; st[11] += wacc[3];
                    st[13] += 0.5 * wacc[1];     // executed: the lower triangle only = half of 4 k n^2 (+ ~4/n for the diagonal blocks; kept out of the step loop)
                    // HBM bytes this chunk really moves for the instance: P and x in and out once per chunk; per step only the
                    // message (replay: 12 nm + 4 in, 24 out) crosses HBM
                    st[12] += ekf_plane_bytes(M_first) + ekf_plane_bytes(M) + 8.0 * (n0 + n) + 4.0 * M + 16.0
                              + (REPLAY ? 12.0 * wacc[3] + 28.0 * T : 8.0 * T);
                    a.progress[inst] = a.t0 + T;
                    bulk_wait_read();    // the tile is re-filled by the next instance
                }
            }
        }
        parity ^= 1u;
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
    const int nb = b.base;
    x[0] = x0; x[1] = y0; x[2] = a2; if (nb == 4) x[3] = a3;
    if (b.ps2g) {
        // packed symmetric layout: block rows 0..1 of both planes, diagonal (1,1), (2,2), (3,3) (padded indices)
        const int live = bpl_plane_doubles(2);
        for (int c = 0; c < live; ++c) { P[c] = 0.0; P[b.ps2g + c] = 0.0; }
        P[bpl_idx(1, 1, b.ps2g)] = 0.01 * 0.01; P[bpl_idx(2, 2, b.ps2g)] = 0.01 * 0.01; P[bpl_idx(3, 3, b.ps2g)] = 0.005 * 0.005;
    } else {
        const int ld = ldp_of(b.fixed_ld, nb);
        for (int r = 0; r < nb; ++r) for (int c = 0; c < ld; ++c) P[r * ld + c] = 0.0;
        P[0] = 0.01 * 0.01; P[ld + 1] = 0.01 * 0.01; P[2 * ld + 2] = 0.005 * 0.005;
        if (nb == 4) P[3 * ld + 3] = 0.005 * 0.005;
    }
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
    L.ps2 = bpl_plane_doubles(2 + L.cap_lm);
    L.from_list = from_list;
    ekf_smem_layout(b.max_meas, L);
    return L;
}

size_t ekf_step_smem_bytes(const BatchState& b) { return (size_t)make_launch(b, b.max_lm, 0).sweep_bytes; }

static constexpr size_t SMEM_PER_SM = 227 * 1024, SMEM_CTA_RESERVED = 1024;

template <int THREADS>
static int step_occupancy(size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ekf_step_kernel<THREADS>, THREADS, smem) != cudaSuccess) nb = 0;
    return nb;
}

// CTA width for a tile: the width that keeps the most instances resident per SM (shared memory, registers and the
// 2048-thread limit decide); among equals the wider CTA (more lanes on each instance's O(n^2) sweep).
static int pick_threads_uncached(size_t smem, int* per_sm_out);
// (five occupancy queries per call are far too slow for a per-tick launch path: remembered per tile size)
static int pick_threads(size_t smem, int* per_sm_out) {
    struct Entry { size_t smem; int threads, per_sm; };
    static Entry cache[128];
    static int n_cache = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].smem == smem) { if (per_sm_out) *per_sm_out = cache[i].per_sm; return cache[i].threads; }
    int per_sm = 1;
    const int th = pick_threads_uncached(smem, &per_sm);
    if (n_cache < 128) cache[n_cache++] = Entry{smem, th, per_sm};
    if (per_sm_out) *per_sm_out = per_sm;
    return th;
}
static int pick_threads_uncached(size_t smem, int* per_sm_out) {
    const int widths[5] = {32, 64, 128, 256, 512};
    const int occ[5] = {step_occupancy<32>(smem), step_occupancy<64>(smem), step_occupancy<128>(smem),
                        step_occupancy<256>(smem), step_occupancy<512>(smem)};
    int best = 2;
    for (int k = 0; k < 5; ++k) {
        // at least ~24 warps per SM are needed before more instances stop paying; then prefer wider CTAs
        const long cur = (long)occ[k] * 1000 + widths[k], bst = (long)occ[best] * 1000 + widths[best];
        const int warps_k = occ[k] * widths[k] / 32;
        if (occ[k] > 0 && (occ[best] == 0 || (warps_k >= 8 && cur > bst))) best = k;
    }
    if (per_sm_out) *per_sm_out = occ[best] > 0 ? occ[best] : 1;
    return widths[best];
}

template <int THREADS>
static cudaError_t set_smem_step(int bytes) {
    return cudaFuncSetAttribute(ekf_step_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
template <int CW>
static cudaError_t set_smem_sweep(int bytes) {
    cudaError_t e = cudaFuncSetAttribute(ekf_sweep_kernel<CW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ekf_sweep_kernel<CW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

cudaError_t ekf_step_configure(const BatchState& b) {
    const EkfLaunch L = make_launch(b, b.max_lm, 0);
    const int bytes = L.smem_bytes;
    cudaError_t e;
    if ((e = set_smem_step<32>(bytes)) != cudaSuccess) return e;
    if ((e = set_smem_step<64>(bytes)) != cudaSuccess) return e;
    if ((e = set_smem_step<128>(bytes)) != cudaSuccess) return e;
    if ((e = set_smem_step<256>(bytes)) != cudaSuccess) return e;
    if ((e = set_smem_step<512>(bytes)) != cudaSuccess) return e;
    const int sbytes = L.sweep_bytes;
    if ((e = set_smem_sweep<1>(sbytes)) != cudaSuccess) return e;
    if ((e = set_smem_sweep<2>(sbytes)) != cudaSuccess) return e;
    if ((e = set_smem_sweep<3>(sbytes)) != cudaSuccess) return e;
    if ((e = set_smem_sweep<4>(sbytes)) != cudaSuccess) return e;
    return set_smem_sweep<8>(sbytes);
}

static cudaError_t launch_step_threads(int threads, int grid, size_t smem, cudaStream_t st, const BatchState& b,
                                       const FilterConst& fc, const StepInputs& in, int phases, const EkfLaunch& L) {
    if (threads == 96) threads = 128;          // 96 (three filter warps) exists for the sweep kernel only
    switch (threads) {
        case 32: ekf_step_kernel<32><<<grid, 32, smem, st>>>(b, fc, in, phases, L); break;
        case 64: ekf_step_kernel<64><<<grid, 64, smem, st>>>(b, fc, in, phases, L); break;
        case 128: ekf_step_kernel<128><<<grid, 128, smem, st>>>(b, fc, in, phases, L); break;
        case 256: ekf_step_kernel<256><<<grid, 256, smem, st>>>(b, fc, in, phases, L); break;
        default: ekf_step_kernel<512><<<grid, 512, smem, st>>>(b, fc, in, phases, L); break;
    }
    return cudaGetLastError();
}

// cap_hint: landmark capacity to size this launch for (<= 0 or >= max_lm: full capacity, no retry pass).
// force_threads: 0 = automatic CTA width, else 32..512 (tuning / tests).
cudaError_t launch_ekf_step(const BatchState& b, const FilterConst& fc, const StepInputs& in, int phases, int cap_hint,
                            int force_threads, cudaStream_t st) {
    const bool limited = cap_hint > 0 && cap_hint < b.max_lm;
    const EkfLaunch L = make_launch(b, limited ? cap_hint : b.max_lm, 0);
    const size_t smem = (size_t)L.smem_bytes;
    // (the deferred-instance counter is re-armed by the retry pass itself: no memset on the stream)
    cudaError_t e = launch_step_threads(force_threads ? force_threads : pick_threads(smem, nullptr), b.batch, smem, st, b, fc, in, phases, L);
    // A full-capacity launch defers nothing, and the hint the retry pass would post is not needed any more either: max(M) only
    // grows between resets, so once hint + headroom reaches max_lm every later launch is a full-capacity one (the posted word keeps
    // its last value; slam_reset zeroes it).  Skipping the empty pass takes ~4 us off every such tick.
    if (e != cudaSuccess || !limited) return e;
    const EkfLaunch R = make_launch(b, b.max_lm, 1);
    const size_t rsmem = (size_t)R.smem_bytes;
    int per_sm = 1;
    int rthreads = pick_threads(rsmem, &per_sm);
    if (force_threads) { rthreads = force_threads; per_sm = (int)(SMEM_PER_SM / (rsmem + SMEM_CTA_RESERVED)) > 0 ? (int)(SMEM_PER_SM / (rsmem + SMEM_CTA_RESERVED)) : 1; }
    // deferrals are rare (the tile has headroom): one CTA per SM drains them, and an empty pass costs a launch of 148 CTAs
    static int sms = 0;
    if (!sms) sms = device_sm_count();
    const int grid = b.batch < sms ? b.batch : sms;
    (void)per_sm;
    return launch_step_threads(rthreads, grid, rsmem, st, b, fc, in, phases, R);
}

template <int CW, bool REPLAY>
static int sweep_occupancy(size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ekf_sweep_kernel<CW, REPLAY>, 32 * (CW + 1), smem) != cudaSuccess) nb = 0;
    return nb;
}

template <int CW, bool REPLAY>
static void sweep_go(int grid, size_t smem, cudaStream_t st, const BatchState& b, const FilterConst& fc, const SimState& sim,
                     const SimConst& sc, const SweepArgs& a, const EkfLaunch& L) {
    ekf_sweep_kernel<CW, REPLAY><<<grid, 32 * (CW + 1), smem, st>>>(b, fc, sim, sc, a, L);
}

template <bool REPLAY>
static cudaError_t launch_sweep_t(const BatchState& b, const FilterConst& fc, const SimState& sim, const SimConst& sc,
                                  const SweepArgs& a, int cap_lm, int force_threads, cudaStream_t st) {
    const EkfLaunch L = make_launch(b, cap_lm, 0);
    const size_t smem = (size_t)L.sweep_bytes;
    cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    // filter warps per instance: the count that keeps the most instances resident per SM; among equals the widest
    const int cws[5] = {1, 2, 3, 4, 8};
    const int occ[5] = {sweep_occupancy<1, REPLAY>(smem), sweep_occupancy<2, REPLAY>(smem), sweep_occupancy<3, REPLAY>(smem),
                        sweep_occupancy<4, REPLAY>(smem), sweep_occupancy<8, REPLAY>(smem)};
    // Filter warps per instance by tile size.  Measured per chunk on B200 (4096 instances, profiles/r02p_sweep_chunks.txt): up
    // to ~20 landmarks one warp with 10 instances per SM (the barrier file allows 64 / 6) wins, to ~38 two warps x 8 (7), to ~46
    // three warps x 6 (5), beyond it four warps x 4 (the rank-2 pass grows with n^2 and wants lanes; residency stops paying once the
    // shared-memory pipe is ~3/4 busy).  A batch that fits the SMs at a wider CTA takes the wider CTA.
    int best = L.cap_lm <= 20 ? 0 : L.cap_lm <= 38 ? 1 : occ[2] >= 5 ? 2 : 3;       // (three warps while five such CTAs fit, ~46 landmarks)
    const int sms_ = device_sm_count();
    while (best < 3 && occ[best + 1] > 0 && b.batch <= sms_ * occ[best + 1]) ++best;
    if (occ[best] <= 0) { best = 0; for (int k = 1; k < 5; ++k) if (occ[k] >= occ[best]) best = k; }
    if (force_threads == 32) best = 0; else if (force_threads == 64) best = 1; else if (force_threads == 96) best = 2;
    else if (force_threads == 128) best = 3; else if (force_threads >= 256) best = 4;
    const int per_sm = occ[best] > 0 ? occ[best] : 1;
    const int sms = device_sm_count();
    static const bool dbg = getenv("SLAM_DEBUG_SWEEP") != nullptr;
    if (dbg) fprintf(stderr, "sweep launch: cap_lm %d smem %zu occ {%d,%d,%d,%d,%d} -> CW %d x %d per SM (t0 %d T %d)\n", L.cap_lm, smem,
                     occ[0], occ[1], occ[2], occ[3], occ[4], cws[best], per_sm, a.t0, a.T);
    // (launching only the CTAs that ceil(batch / resident CTAs) rounds need -- 1366 instead of 1480 at 10 per SM -- measured
    // slower, 130.1 -> 126.9 M updates/s: the work counter already evens the rounds out)
    const int grid = b.batch < sms * per_sm ? b.batch : sms * per_sm;
    switch (cws[best]) {
        case 1: sweep_go<1, REPLAY>(grid, smem, st, b, fc, sim, sc, a, L); break;
        case 2: sweep_go<2, REPLAY>(grid, smem, st, b, fc, sim, sc, a, L); break;
        case 3: sweep_go<3, REPLAY>(grid, smem, st, b, fc, sim, sc, a, L); break;
        case 4: sweep_go<4, REPLAY>(grid, smem, st, b, fc, sim, sc, a, L); break;
        default: sweep_go<8, REPLAY>(grid, smem, st, b, fc, sim, sc, a, L); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_ekf_sweep(const BatchState& b, const FilterConst& fc, const SimState& sim, const SimConst& sc,
                             const SweepArgs& a, bool replay, int cap_lm, int force_threads, cudaStream_t st) {
    return replay ? launch_sweep_t<true>(b, fc, sim, sc, a, cap_lm, force_threads, st)
                  : launch_sweep_t<false>(b, fc, sim, sc, a, cap_lm, force_threads, st);
}

}  // namespace slam
