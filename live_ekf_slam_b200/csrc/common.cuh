// common.cuh -- shared device helpers and the batch memory layout for the sm_100a filter kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slam_filter.h"

#define PI_REF 3.14159265358979323846  // filter.h:42
#define TWO_PI_REF (2 * PI_REF)

namespace slam {

// ---------------------------------------------------------------------------------------------
// HBM layout of a batch of filter instances (all arrays are struct-of-arrays over instances).
//
//   P      [batch][p_stride]   covariance.  UKF / large-map EKF: row-major, fixed leading dimension ldg(n_max).
//                              Batched EKF: PACKED SYMMETRIC (see bpl_idx below): only the lower triangle exists,
//                              so ~8 n^2 bytes cross HBM per step and direction-pair instead of 16 n^2.
//   x      [batch][x_stride]   committed state x_t (EKF: x,y,yaw,lm.. ; UKF: x,y,cos,sin,lm..)
//   ids    [batch][max_lm]     lm_IDs (filter.h:70)
//   meta   [batch] int4 {M, status, timestep, n_assoc}
//   assoc  [batch][max_meas]   slot index (or -1) chosen for each measurement of the last step
// ---------------------------------------------------------------------------------------------
struct BatchState {
    double* P;
    double* x;
    int* ids;
    int4* meta;        // [batch] {M, status, timestep, n_assoc}: one 16-byte load per instance at kernel entry
    int* assoc;
    int* retry_list;   // [batch] instances deferred by a capacity-limited launch
    int* retry_count;  // [2]: [0] deferred instances of this step, [1] CTAs of the retry pass that have finished
    int* hint_host;    // mapped pinned host word (or null): the retry pass posts max_M there -- the capacity hint without a copy
    int* max_M;        // [1] running max of M over the batch (capacity hint for the next launches)
    double* stats;     // [batch][SLAM_NUM_STATS] per-instance accumulators
    double* sigma;     // UKF only: [batch][sigma_stride] sigma points X (point-major), optional
    long long p_stride;
    long long sigma_stride;
    int x_stride;
    int batch;
    int max_lm;
    int max_meas;
    int base;          // 3 (EKF) or 4 (UKF)
    int n_max;         // base + 2*max_lm
    int lds;           // shared-memory leading dimension of P
    int fixed_ld;      // global leading dimension of P (ldg(n_max)); a live row is its first ldg(n) doubles
    int ps2g;          // EKF batch kernels: plane stride (doubles) of the packed symmetric layout in HBM, 0 = row-major
};

// Effective filter constants after readCommonParams (filter.h:105-121).
struct FilterConst {
    double V00, V11;   // process noise used by the filter
    double W00, W11;   // sensing noise used by the filter
    float v_d, v_th, w_r, w_b;
    float min_sep;
    int id_known;
    int loc;             // UKF localisation-only mode (FilterChoice::UKF_LOC): landmarks come from the true map
    int n_map;           // landmarks in `map`
    const float* map;    // device copy of /truth/landmarks, float32 [id, x, y]*  (filter.h:68)
};

struct SimConst {
    double V_00, V_11, W_00, W_11;   // uniform half-widths, sim_node.py:216-217,247-248
    double d_max, th_max, range_max, fov_min, fov_max;
};

__host__ __device__ inline int ldg_of(int n) { return (n + 1) & ~1; }
__host__ __device__ inline int ldp_of(const int fixed_ld, int n) { return fixed_ld ? fixed_ld : ((n + 1) & ~1); }

// smallest even leading dimension >= n_max+1 with lds % 4 == 2: rows stay 16-byte aligned for bulk copies
// and a column walk hits 8 distinct 8-byte bank slots per half-warp (2-way conflict at worst).
__host__ __device__ inline int lds_of(int n_max) {
    int l = ldg_of(n_max);
    while ((l & 3) != 2) l += 2;
    return l;
}

// ---------------------------------------------------------------------------------------------
// Packed symmetric covariance of the batched EKF kernels (shared memory AND HBM).
// Indices are PADDED by one (r' = r + 1): row/column 0 is a permanent zero phantom, so that the vehicle is rows
// 1..3 and landmark s is rows 4+2s, 5+2s = block row 2+s -- inserting a landmark appends exactly one block row.
// Only the lower triangle (r' >= c') is stored, as rows of even length in two planes: plane (r' & 1) holds row r' at
// offset a(a+1), a = r' >> 1, with 2(a+1) entries (columns 0 .. 2a+1; the entry (2a, 2a+1) is never read).  A 2x2
// block (a, b) is therefore one double2 in each plane at the same offset a(a+1) + 2b: the rank-2 update works on
// whole blocks with conflict-free 16-byte accesses, a row is contiguous for the O(n) gathers, and a live plane is
// ONE contiguous run of A(A+1) doubles (A = 2 + M block rows) = one bulk copy.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int bpl_plane_doubles(int A) { return A * (A + 1); }
__host__ __device__ inline int bpl_idx(int r, int c, int ps2) { return (r & 1) * ps2 + (r >> 1) * ((r >> 1) + 1) + c; }   // r >= c
__host__ __device__ inline int bpl_sym(int r, int c, int ps2) { return r >= c ? bpl_idx(r, c, ps2) : bpl_idx(c, r, ps2); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, identical to oracle_philox (counter = instance, step, channel, 0; key = seed)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ inline double uniform53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

#ifdef __CUDACC__
// D-1: cos/sin of a float argument pinned as (float)cos((double)x) (oracle/slam_oracle.c cos_f/sin_f)
__device__ __forceinline__ float cos_f(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float sin_f(float x) { return (float)sin((double)x); }

// remainder(a, 2*pi) (filter.h:42 pi), exact.  |a| <= pi: the IEEE result is a itself (n = 0, ties to even) -- the
// common case on the filter path.  Otherwise n' = rint(a / 2pi) and r = fma(-n', 2pi, a): whenever |r| is clearly
// below pi, n' is the integer nearest to a / 2pi, so a - n' 2pi IS the IEEE remainder, which is representable, and the
// fma (exact product, one rounding) returns it exactly.  Anything near the +-pi boundary goes to the library routine.
// (the library routine stays out of line: its loop inlined at every call site grows the kernels' hot paths, which the
// persistent EKF kernel pays for in instruction fetch)
static __device__ __noinline__ double wrap_2pi_library(double a) { return remainder(a, TWO_PI_REF); }
__device__ __forceinline__ double wrap_2pi(double a) {
    if (fabs(a) <= PI_REF) return a;
    if (fabs(a) < 1.0e6) {
        const double n = rint(a * (1.0 / TWO_PI_REF));
        const double r = fma(-n, TWO_PI_REF, a);
        if (fabs(r) < 3.1415) return r;
    }
    return wrap_2pi_library(a);
}

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA without a tensor map: cp.async.bulk, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on an mbarrier (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, bulk-group completion
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif

// SMs of the current device (grids of the persistent / cooperative kernels are sized from it, never from a constant)
inline int device_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 1;
    return sms;
}

// host-side launchers (defined in the .cu files, called from capi.cu)
struct StepInputs {
    const float* fwd;      // device
    const float* ang;      // device
    int cmd_stride;        // 0 shared / 1 per instance
    const float* meas;     // device [batch][max_meas][3]
    const int* n_meas;     // device [batch]
    double* poses;         // optional [batch][3]: the batched EKF step kernel writes the committed vehicle pose itself (the
                           // pose read-back of publishState fused into the step); device or MAPPED PINNED HOST memory
};
// (all StepInputs pointers may be mapped pinned host memory: the kernels then read the caller's buffers over PCIe directly)

enum { STEP_PREDICT = 1, STEP_UPDATE = 2 };

cudaError_t launch_ekf_step(const BatchState& b, const FilterConst& fc, const StepInputs& in, int phases, int cap_hint,
                            int force_threads, cudaStream_t st);
size_t ekf_step_smem_bytes(const BatchState& b);
cudaError_t ekf_step_configure(const BatchState& b);

// HBM scratch between the three launches of a UKF step (csrc/ukf_batch.cu)
struct UkfScratch {
    double* Zg;         // [batch][n_max * n_max]  generation 2: Householder reflectors by rows, tau on the diagonal;
                        //                         generation 1 / rescue: Q^T of the tridiagonalisation (compact, ld n)
    double* Yg;         // [batch][n_max * n_max]  generation 2: 2w * Y, the landmark-block seed of P_pred (compact, ld n)
    double* dg;         // [n_max][batch]  diagonal of T -> eigenvalues
    double* eg;         // [n_max][batch]  off-diagonal of T
    double2* rot;       // [batch][rot_cap]  (c, s) of every QL plane rotation, in generation order
    int2* swp;          // [batch][swp_cap]  (l, m) range of every QL sweep
    int* defer;         // [batch]  generation 2: 1 = this step's updates did not fit the narrow tile (full-width pass takes it)
    int narrow;         // 1 (default): narrow-tile first pass + full-width deferred pass; 0: full-width tile only
    int* nswp;          // [batch]  number of sweeps logged; -1: log overflow (the back kernel redoes the QL itself)
    long long rot_cap;
    int swp_cap;
    int n_max;
    int gen;            // 3 (default): parallel tridiagonal eigensolver + dense products with its eigenvectors; 2: QL rotation
                        // log replayed on the vectors, warp per instance; 1: explicit eigenvector matrix of Y, CTA per instance
    double* Vg;         // [batch][n_max * n_max]  generation 3: V[i][k], eigenvectors of the tridiagonal matrix (compact, ld n)
    double* VTg;        // [batch][n_max * n_max]  generation 3: its transpose (work space of the eigensolver before)
    unsigned long long* routes;   // [4] instance-steps taken by: dense route (gen 3), QL route (gen 2), gen-1 / rescue, (spare)
    int multiwarp;      // generation 3: 1 (default) = back kernel with one warp per group of four vectors, 0 = one warp per instance
    int maxc;           // generation 3: largest cluster of close eigenvalues handled in the kernel (larger: QL route)
    int front_packed;   // front kernel's matrix in shared memory: 2 (default) = full square while it fits 4x per SM, packed lower triangle beyond; 1 = packed; 0 = full
    int refine_all;     // test knob: the tile kernel hands EVERY instance with a cluster to the two-array kernel (as if it needed refinement)
    int eig3_tile;      // generation 3: 1 (default) = eigenvectors built in a shared-memory tile when it fits twice per SM, 0 = in the global scratch
    double* xprior;     // [batch][n_max]  x_t at the start of the last step: column 0 of the sigma-point matrix X (ukf.cpp:214)
    int2* sigfmt;       // [batch]  what the scratch holds of the last step's sqrt factor, for slam_get_sigma_points:
                        //          .x = 0 nothing yet (X is the constructor's 4 x 9 zero matrix, ukf.cpp:20), 2 = reflectors (Zg) +
                        //          rotation log + eigenvalues (dg), 1 = explicit Z^T (Zg, compact) + sqrt(max(d, 1e-8)) (dg); .y = n
    int clip_lanes;     // test knob: max clipped eigenvectors riding beside pass A (0 = as many as fit)
};
// streams / events of the sliced generation-2 step (owned by the handle)
constexpr int UKF_MAX_SUB = 8;
struct UkfStreams {
    int nsub = 1;                               // slices of the batch (1 = everything on the handle's stream)
    cudaStream_t aux[UKF_MAX_SUB - 1] = {};
    cudaEvent_t fork = nullptr;
    cudaEvent_t join[UKF_MAX_SUB - 1] = {};
};
// launches one UKF step; *launched receives the number of kernels launched
cudaError_t launch_ukf_step(const BatchState& b, const FilterConst& fc, const StepInputs& in, const UkfScratch& u, cudaStream_t st,
                            const UkfStreams& xs, int* launched);
size_t ukf_step_smem_bytes(const BatchState& b);
// X of the last UKF step of one instance (ukf.cpp:214-220), point-major n x (2n+1), into the DEVICE buffer d_X; n, fmt from sigfmt
cudaError_t launch_ukf_sigma_points(const BatchState& b, const UkfScratch& u, int inst, int fmt, int n, double* d_X, cudaStream_t st);
bool ukf_gen2_supported(const BatchState& b);
cudaError_t launch_naive_step(const BatchState& b, const StepInputs& in, cudaStream_t st);
cudaError_t ukf_step_configure(const BatchState& b);

// single large-map EKF instance (csrc/ekf_large.cu): P stays in HBM with a fixed leading dimension
struct LargeState {
    double* P;        // [n_max][ld]
    double* x;        // committed x_t
    double* xp;       // running x_pred
    double* U;        // [max_meas][n_max][2]   -K_q of the step's updates
    double* G;        // [max_meas][2][ld]      G_q = H_q P_{q-1}
    int* ids;         // lm_IDs
    int4* meta;       // {M, status, timestep, n_assoc}
    int* assoc;       // [max_meas]
    int* ctl;         // per measurement: [4l+0]=slot, [4l+1]=updates before, [4l+2]=M before, [4l+3]=kind (0 skip, 1 update, 2 insert)
    int* cur;         // [0]=updates so far, [1]=M so far, [2]=status, [3]=M at step start
    double* sc;       // scalars of the current measurement: H[10], nu[2], cb, sb, r, b
    double* stats;
    int ld, n_max, max_lm, max_meas;
};
cudaError_t ekf_large_configure();      // per handle / device: opt-in shared memory of the contraction kernel
cudaError_t launch_ekf_large_step(const LargeState& L, const FilterConst& fc, const float* d_fwd, const float* d_ang,
                                  const float* d_meas, const int* d_nmeas, int n_upper, cudaStream_t st, long long* launches,
                                  cudaEvent_t gemm_ev0 = nullptr, cudaEvent_t gemm_ev1 = nullptr);

struct SimState {
    double* truth;         // [batch][3]
    const double* lm_xy;   // [n_lm][2] shared by the batch (lm_stride 0), or [batch][n_lm][2] (lm_stride = 2 n_lm) after slam_sim_make_maps
    long long lm_stride;   // doubles between the maps of consecutive vehicles (0: one shared map)
    float* meas;           // [batch][max_meas][3]
    int* n_meas;           // [batch]
    int* overflow;         // [batch] detections dropped because max_meas was reached
    int n_lm;
    int batch;
    int max_meas;
    uint32_t instance_offset;
    uint32_t k0, k1;
};
// sim_node.py:63-152 trajectory generation parameters (params.yaml:70-73,90-91 + the start pose)
struct TspParams {
    double landmark_noise, visitation_threshold, bound;
    double x0, y0, yaw0;
    int T;
};
// generate_landmarks (sim_node.py:155-206) for every vehicle on the device; map_type 0 = grid, 1 = random
cudaError_t launch_make_maps(const SimState& s, int map_type, int n_lm, double bound, double grid_step, double min_sep, double* d_maps,
                             int* d_fail, cudaStream_t st);
cudaError_t launch_tsp_trajectories(const SimState& s, const SimConst& sc, const TspParams& tp, float* d_fwd, float* d_ang, cudaStream_t st);
// One chunk of a Monte-Carlo sweep / trajectory replay (csrc/ekf_batch.cu: ekf_sweep_kernel).
struct SweepArgs {
    const float* cmd_fwd;  // device, [T] (cmd_stride 0) or [T][batch], already offset to the chunk
    const float* cmd_ang;
    int cmd_stride;
    int T;                 // steps of this chunk
    uint32_t first_step;   // Philox step counter of the chunk's first step
    int t0;                // run-relative index of the chunk's first step: only instances with progress == t0 run
    int* progress;         // [batch] run-relative steps completed
    int* work_counter;     // zeroed by the launcher
    // replay mode: the messages come from HBM (uploaded host buffers) instead of the simulator, poses go out
    const float* r_meas;   // [T][batch][max_meas][3]
    const int* r_nmeas;    // [T][batch]
    double* r_poses;       // [T][batch][3] (may be null)
};
// cap_lm: landmark capacity of the shared-memory tile of this launch; an instance that would outgrow it is left
// untouched (progress unchanged) for a later launch with a larger tile.  Returns the number of kernels launched.
cudaError_t launch_ekf_sweep(const BatchState& b, const FilterConst& fc, const SimState& sim, const SimConst& sc,
                             const SweepArgs& a, bool replay, int cap_lm, int force_threads, cudaStream_t st);
cudaError_t launch_sim_step(const SimState& s, const SimConst& sc, const float* d_fwd, const float* d_ang,
                            int cmd_stride, uint32_t step, cudaStream_t st);
cudaError_t launch_accumulate_error(const BatchState& b, const SimState& s, cudaStream_t st);
cudaError_t launch_error_histogram(const BatchState& b, double* d_avg, unsigned long long* d_counts, double lo, double hi, int nbins,
                                   cudaStream_t st);
cudaError_t launch_reduce_stats(const BatchState& b, double* d_out, cudaStream_t st);
cudaError_t launch_poses(const BatchState& b, double* d_out, cudaStream_t st);
cudaError_t launch_gather_inputs(const float* h_fwd, const float* h_ang, int ncmd, const int* h_n, const float* h_meas, int batch,
                                 int meas_floats, float* d_fwd, float* d_ang, int* d_n, float* d_meas, cudaStream_t st);
cudaError_t launch_reset(const BatchState& b, double x0, double y0, double a2, double a3, cudaStream_t st);

}  // namespace slam
