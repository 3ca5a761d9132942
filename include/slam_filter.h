/*
 * slam_filter.h -- C-ABI of the B200-native EKF-SLAM / UKF-SLAM filter hot path.
 *
 * Drop-in boundary for the abstract `Filter` plugin of kevin-robb/live_ekf_slam
 * (ekf_ws/src/localization_pkg/include/localization_pkg/filter.h:54-145) and for the simulator's
 * measurement generator (ekf_ws/src/base_pkg/src/sim_node.py:209-250).  Plain pointers and sizes only;
 * no exceptions cross this boundary (every call returns 0 on success, non-zero on error, and
 * slam_last_error() explains).  Every entry point launches hand-written sm_100a CUDA kernels;
 * there is no CPU fallback.
 *
 * A handle owns `batch` independent filter instances (Monte-Carlo runs).  batch == 1 is the
 * reference's single filter.  A handle is NOT thread-safe (the reference's filter is driven by a
 * single-threaded ros::spin(), localization_node.cpp:197); each handle owns one CUDA stream, calls are
 * asynchronous on it and getters synchronise.
 *
 * Reference paths cited below are relative to ekf_ws/src/.
 */
#ifndef SLAM_FILTER_H
#define SLAM_FILTER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FilterChoice, localization_pkg/include/localization_pkg/filter.h:44-51 (same numeric values) */
#define SLAM_EKF_SLAM 1
#define SLAM_UKF_LOC  2   /* UKF localisation on the true map (ukf.cpp:146-154,262,272,296-302); state stays (x,y,cos,sin) */
#define SLAM_UKF_SLAM 3
#define SLAM_NAIVE    5   /* NaiveFilter, filter.h:325-370: command propagation only */

/* per-instance status bits (replace the reference's uncaught std::runtime_error / eigen_assert, filter.h:5) */
#define SLAM_STATUS_NAN               1  /* state or covariance became non-finite                               */
#define SLAM_STATUS_SAME_STEP_REMATCH 2  /* unknown-ID EKF matched a landmark inserted in the same step; the     */
                                         /* reference indexes x_t out of range there (ekf.cpp:115) and dies      */
#define SLAM_STATUS_CAPACITY          4  /* an insertion was dropped: max_landmarks reached                      */
#define SLAM_STATUS_MEAS_OVERFLOW     8  /* a step delivered more than max_meas measurements (extra ones dropped)*/
#define SLAM_STATUS_BAD_ID           16  /* UKF_LOC: a detection's id is outside the true map (the reference reads */
                                         /* map[] out of range there, ukf.cpp:152); the detection is skipped      */

/* What Filter::readCommonParams reads from params.yaml (filter.h:105-121) plus the simulator's
 * constraints (base_pkg/config/params.yaml:27-32).  yaml values go in unchanged; with
 * compat_noise_bug != 0 the library reproduces filter.h:116-117 (V <- sensing covs, W = I). */
typedef struct slam_params {
    float  v_d, v_th;               /* process_noise.mean.{v_d,v_th}          filter.h:108-109 */
    float  w_r, w_b;                /* sensing_noise.mean.{w_r,w_b}           filter.h:114-115 */
    double V_00, V_11;              /* process_noise.cov.{V_00,V_11}          filter.h:110-111 */
    double W_00, W_11;              /* sensing_noise.cov.{W_00,W_11}          filter.h:116-117 */
    int    landmark_id_is_known;    /* constraints.measurements.*             filter.h:119     */
    float  min_landmark_separation; /*                                        filter.h:120     */
    int    compat_noise_bug;        /* 1 (default of the shims) = reference behaviour          */
    double d_max, th_max;           /* constraints.commands   (simulator)     sim_node.py:219-220 */
    double range_max, fov_min, fov_max; /* constraints.vision (simulator)     sim_node.py:239-241 */
} slam_params;

typedef struct slam_filter* slam_handle_t;
typedef struct slam_sim*    slam_sim_t;

/* ---- construction: replaces `filter = std::make_unique<EKF|UKF>(); filter->readParams(config)`
 *      (localization_pkg/src/localization_node.cpp:33-47, ekf.cpp:4-27, ukf.cpp:3-29).
 *      kind: SLAM_EKF_SLAM | SLAM_UKF_SLAM | SLAM_UKF_LOC | SLAM_NAIVE.  max_landmarks bounds M (UKF_LOC: the
 *      size of the true map); max_meas bounds detections/step. */
int  slam_create(int kind, const slam_params* params, int batch, int max_landmarks, int max_meas,
                 int device, slam_handle_t* out);
int  slam_destroy(slam_handle_t h);
/* UKF_LOC only: the true map, as the /truth/landmarks message the reference stores in Filter::map (filter.h:68,
 * localization_node.cpp:66-75): float32 [id, x, y] * n_landmarks with id == index.  HOST pointer; copied. */
int  slam_set_map(slam_handle_t h, const float* map_id_x_y, int n_landmarks);
const char* slam_last_error(slam_handle_t h);          /* h may be NULL: error of the last failed create */
void* slam_stream(slam_handle_t h);                    /* the handle's cudaStream_t */
int  slam_synchronize(slam_handle_t h);
int  slam_batch(slam_handle_t h);
int  slam_kind(slam_handle_t h);

/* ---- Filter::init(float x_0, float y_0, float yaw_0)  (filter.h:60, ekf.cpp:29-34, ukf.cpp:31-45).
 *      Every instance of the batch gets the same start pose and is reset to timestep 0, M = 0. */
int  slam_init(slam_handle_t h, float x_0, float y_0, float yaw_0);

/* ---- Filter::update(Command cmdMsg, Float32MultiArray lmMeasMsg)  (filter.h:61, ekf.cpp:37-179,
 *      ukf.cpp:161-195): one fused predict+update per instance.
 *      fwd/ang: Command.msg:3-5 float32; cmd_stride 0 = one command shared by the batch, 1 = per instance.
 *      meas: float32 [batch][max_meas][3] = [id, range, bearing]* exactly as on /landmark
 *      (sim_node.py:245-250); n_meas[batch] = detections per instance (lm_meas.size()/3, ekf.cpp:65).
 *      slam_step takes HOST pointers (copied on the handle's stream); *_device takes DEVICE pointers. */
int  slam_step(slam_handle_t h, const float* fwd, const float* ang, int cmd_stride,
               const float* meas, const int* n_meas);
int  slam_step_device(slam_handle_t h, const float* d_fwd, const float* d_ang, int cmd_stride,
                      const float* d_meas, const int* d_n_meas);

/* ---- split form asked for by the north star: predict from the commanded motion, then update from the
 *      measurements.  predict == ekf.cpp:43-61 followed by the early-return commit of :67-71;
 *      update == ekf.cpp:73-177 with the landmark means snapshotted at entry (SURVEY B-3).  EKF only:
 *      the UKF's update stage consumes the propagated sigma points of the same call (ukf.cpp:305-337). */
int  slam_predict(slam_handle_t h, const float* fwd, const float* ang, int cmd_stride);
int  slam_update(slam_handle_t h, const float* meas, const int* n_meas);
int  slam_predict_device(slam_handle_t h, const float* d_fwd, const float* d_ang, int cmd_stride);
int  slam_update_device(slam_handle_t h, const float* d_meas, const int* d_n_meas);

/* ---- getters: replace getStateVector() (filter.h:76, ekf.cpp:182-185, ukf.cpp:47-53) and the fields
 *      publishState() serialises (ekf.cpp:192-220, ukf.cpp:60-104).  FP64 out; the shim down-casts to the
 *      float32 wire types.  All synchronise the handle's stream. */
int  slam_get_timestep(slam_handle_t h, int inst, int* timestep);
int  slam_get_num_landmarks(slam_handle_t h, int inst, int* M);
int  slam_get_status(slam_handle_t h, int inst, int* status);
int  slam_get_state(slam_handle_t h, int inst, double* x, int* n);          /* raw x_t: EKF 3+2M, UKF 4+2M */
int  slam_get_state_vector(slam_handle_t h, int inst, double* xv, int* n);  /* (x,y,yaw,lm...) = 3+2M      */
int  slam_get_cov(slam_handle_t h, int inst, double* P_rowmajor, int* n);   /* n*n row-major (ekf.cpp:211-217) */
int  slam_get_landmark_ids(slam_handle_t h, int inst, int* ids, int* M);    /* lm_IDs, filter.h:70 */
int  slam_get_assoc(slam_handle_t h, int inst, int* slot, int* k);          /* last step: slot index or -1 (new) per measurement */
/* UKF sigma points of the last update() (ukf.cpp:214-220), flattened point-major exactly as UKFState.X is filled
 * (ukf.cpp:91-99): X[j*n + i] = component i of sigma point j, j = 0..2n, with n = 4 + 2M of that update's PRIOR (the
 * matrix keeps the size it had when the step started, ukf.cpp:167-171).  X must hold (4 + 2 max_landmarks) *
 * (2 (4 + 2 max_landmarks) + 1) doubles.  Before the first update: the constructor's 4 x 9 zeros (ukf.cpp:20).
 * Materialised on demand from the factors the step left on the device; getter-only cost. */
int  slam_get_sigma_points(slam_handle_t h, int inst, double* X, int* n);
int  slam_get_poses(slam_handle_t h, double* xyyaw);                        /* [batch][3] vehicle pose estimates */
int  slam_get_all_status(slam_handle_t h, int* status);                     /* [batch] */
int  slam_get_all_num_landmarks(slam_handle_t h, int* M);                   /* [batch] */
/* teacher forcing (tests): overwrite the committed state of one instance */
int  slam_set_state(slam_handle_t h, int inst, const double* x, const double* P_rowmajor,
                    const int* ids, int M, int timestep);

/* ---- workload source: the simulator's measurement generator, base_pkg/src/sim_node.py:209-250
 *      (get_cmd: noisy clamped command -> truth propagation -> range/FOV visibility -> noisy float32
 *      [id,r,b]).  One simulated vehicle per filter instance, all on the shared landmark map lm_xy
 *      [n_lm][2].  Noise: Philox4x32-10 keyed (seed; instance_offset+i, step, channel), uniform +-cov
 *      exactly like sim_node.py:216-217,247-248.  Bound to the filter handle's stream and batch. */
int  slam_sim_create(slam_handle_t h, const double* lm_xy, int n_lm, uint64_t seed,
                     uint32_t instance_offset, slam_sim_t* out);
int  slam_sim_destroy(slam_sim_t s);
/* Precomputed command trajectories for the whole batch on the device (replaces the host pre-pass generate_trajectory,
 * base_pkg/src/sim_node.py:63-152, run once per Monte-Carlo instance): noisy copy of the map, nearest-neighbour tour,
 * one clamped command per step.  Map noise: Philox keyed (seed; global instance, landmark id, 0, 1).  d_fwd / d_ang:
 * DEVICE buffers [T][batch] float32 (the Command.msg wire values), usable as slam_run_device(..., cmd_stride = 1, ...).
 * landmark_noise / visitation_threshold / bound: params.yaml:90,91,70. */
int  slam_sim_make_trajectories(slam_sim_t s, double landmark_noise, double visitation_threshold, double bound,
                                double x_0, double y_0, double yaw_0, int T, float* d_fwd, float* d_ang);
/* generate_landmarks (base_pkg/src/sim_node.py:155-206) on the device, ONE MAP PER simulated vehicle (Monte-Carlo sweeps over
 * random maps, like the reference's recorded runs): map_type 0 = "grid" (:165-176; the lattice np.arange(-bound + step/2, bound,
 * step)^2, ids row-major, n_landmarks ignored), 1 = "random" / "rand" (:177-188 on the blank occupancy map: uniform positions in
 * [-bound, bound)^2 at least min_sep apart, rejection sampled; Philox keyed (seed; global instance, attempt, 0, 2)).  The "demo"
 * and "igvc1" choices are fixed coordinate tables: pass them to slam_sim_create.  Replaces the simulator's shared map; *n_out =
 * landmarks per map.  The filter handle's max_landmarks must cover it.  Other map_type values fail like sim_node.py:196-198. */
int  slam_sim_make_maps(slam_sim_t s, int map_type, int n_landmarks, double bound, double grid_step, double min_sep, int* n_out);
int  slam_sim_get_map(slam_sim_t s, int inst, double* lm_xy /* [n_lm][2] */, int* n_lm);   /* a vehicle's map to HOST memory */
int  slam_sim_reset(slam_sim_t s, double x_0, double y_0, double yaw_0);
/* one get_cmd() for every vehicle; commands are HOST (slam_sim_step) or DEVICE (…_device) float32.
 * Results stay on the device: slam_sim_meas()/slam_sim_n_meas() return the DEVICE buffers
 * ([batch][max_meas][3] float32, [batch] int32) that slam_step_device consumes. */
int  slam_sim_step(slam_sim_t s, const float* fwd, const float* ang, int cmd_stride, uint32_t step);
int  slam_sim_step_device(slam_sim_t s, const float* d_fwd, const float* d_ang, int cmd_stride, uint32_t step);
const float* slam_sim_meas(slam_sim_t s);
const int*   slam_sim_n_meas(slam_sim_t s);
int  slam_sim_get_truth(slam_sim_t s, double* xyyaw);                       /* [batch][3] to host */
int  slam_sim_get_meas(slam_sim_t s, float* meas, int* n_meas);             /* copy last step's messages to host */

/* ---- Monte-Carlo sweep: T reference steps for every instance in ONE launch sequence, the simulator
 *      feeding the filter on the device (sim_node.py:143-152 publishes cmd t, get_cmd produces meas t,
 *      localization_node.cpp:108-131 consumes both).  cmd_fwd/cmd_ang: HOST float32 [T] shared trajectory
 *      (cmd_stride 0) or [T][batch] (cmd_stride 1).  first_step numbers the Philox step counter.
 *      Error statistics against the simulator's truth are accumulated per instance (see slam_get_stats). */
int  slam_run(slam_handle_t h, slam_sim_t s, const float* cmd_fwd, const float* cmd_ang, int cmd_stride,
              int T, uint32_t first_step);
/* same with the command trajectory already resident in HBM (DEVICE pointers) */
int  slam_run_device(slam_handle_t h, slam_sim_t s, const float* d_cmd_fwd, const float* d_cmd_ang, int cmd_stride,
                     int T, uint32_t first_step);
/* Filter::init executed on the device (no host staging, asynchronous): used between Monte-Carlo sweeps */
int  slam_reset(slam_handle_t h, float x_0, float y_0, float yaw_0);
/* Filter::update + the pose read-back of publishState in one asynchronous call: HOST buffers in, HOST poses
 * [batch][3] out; the caller synchronises (slam_synchronize) before reading poses_out.  With PINNED host buffers
 * (cudaHostAlloc / cudaHostRegister / torch pin_memory) the kernels read the inputs and write the poses in place over
 * PCIe -- no staging copies, one kernel launch per tick for the batched EKF; pageable buffers are staged through copies. */
int  slam_step_io(slam_handle_t h, const float* fwd, const float* ang, int cmd_stride,
                  const float* meas, const int* n_meas, double* poses_out);

/* T x slam_step_io for a whole recorded run in ONE asynchronous call (the reference's results-only mode replays a
 * precomputed trajectory the same way: base_pkg/launch/filter_demo_results_only.launch, sim_node.py:143-152).
 * HOST buffers (pinned memory makes the copies overlap the kernels): cmd_fwd/cmd_ang [T] (cmd_stride 0) or
 * [T][batch]; meas [T][batch][max_meas][3]; n_meas [T][batch]; poses_out [T][batch][3] (may be NULL).  Chunks of the
 * run are uploaded, filtered and downloaded in a three-stage pipeline; known-ID EKF batches keep each filter resident
 * in shared memory for a whole chunk.  Available on every path (the HBM-resident large map runs per-step launches).  The caller
 * synchronises (slam_synchronize) before reading poses_out. */
/* (Available on every path, the HBM-resident large map included.) */
int  slam_run_io(slam_handle_t h, const float* cmd_fwd, const float* cmd_ang, int cmd_stride,
                 const float* meas, const int* n_meas, double* poses_out, int T);

/* ---- accuracy statistics accumulated by slam_run / slam_accumulate_error, summed over the batch:
 *      out[0]=count, [1]=sum ex^2, [2]=sum ey^2, [3]=sum eyaw^2 (wrapped), [4]=sum sqrt(ex^2+ey^2)
 *      (the reference's only metric, plotting_node.py:212-214), [5]=sum 3-dof pose NEES (extension),
 *      [6]=instances with non-zero status, [7]=sum of M; work counters kept by the step kernels:
 *      [8]=sum of algorithmic HBM bytes (16 n^2 + 16 n + 12 (k+j) + 8 per update), [9]=sum of
 *      algorithmic flops (EKF 4 k n^2; UKF 9n^3+2n^3+2n^2(2n+1)+12kn^2), [10]=sum of n, [11]=sum of k+j;
 *      and what the kernels really do, as a model kept beside the algorithmic figures: [12]=sum of the HBM bytes the
 *      launches move for an instance (batched EKF: the PACKED lower triangle each way -- once per step on the per-step
 *      kernel, once per chunk on the sweep kernel), [13]=sum of the flops they execute (batched EKF and large map:
 *      the lower triangle only, half of 4 k n^2; UKF: tridiagonalisation + eigensolver + the S-products, a model).
 *      Ranks all-reduce this vector (SUM). */
#define SLAM_NUM_STATS 14
int  slam_accumulate_error(slam_handle_t h, slam_sim_t s);
int  slam_get_stats(slam_handle_t h, double* out /* SLAM_NUM_STATS */);
int  slam_reset_stats(slam_handle_t h);
/* Per-run accuracy analytics at Monte-Carlo scale.  The reference reduces every run to ONE number, the average position
 * error of the vehicle history (compute_average_error, base_pkg/src/plotting_node.py:195-218), and tabulates it over the
 * recorded runs of a setting (base_pkg/src/make_bar_graphs.py:11-18,55).  Here that number is produced on the device for
 * every instance of the batch (avg_err[batch], may be NULL) together with its histogram: counts[0] = runs below lo,
 * counts[1 + k] = runs in [lo + k (hi-lo)/nbins, lo + (k+1) (hi-lo)/nbins), counts[nbins + 1] = runs at or above hi
 * (or non-finite).  HOST pointers; synchronises.  Ranks all-reduce `counts` (SUM, exact integers). */
int  slam_get_error_histogram(slam_handle_t h, double lo, double hi, int nbins, long long* counts /* nbins + 2 */,
                              double* avg_err /* batch, or NULL */);

/* ---- introspection for the benchmark harness */
long long slam_kernel_launches(slam_handle_t h);      /* kernels launched by this handle so far */
/* per-launch device timing (CUDA events on the handle's stream).  on = 1: the filter-step kernel(s) of every Filter::update
 * (forces per-step launches), 2: the persistent sweep kernel of slam_run*, 3: large-map path, the closing DMMA contraction
 * (lm_gemm) only, 0: off */
int  slam_set_profiling(slam_handle_t h, int on);
int  slam_get_profile(slam_handle_t h, double* total_ms, long long* launches);  /* synchronises; resets the pool */
int  slam_build_info(char* buf, int cap);             /* arch, compile flags */
/* UKF: instance-steps taken so far by each route of the step (see slam_tune key 7): out[0] = parallel eigensolver + dense
 * products (generation 3), out[1] = QL rotation log (generation 2, or an instance generation 3 declined), out[2] = explicit
 * eigenvector matrix (generation 1, or the rescue pass).  Synchronises. */
int  slam_get_ukf_routes(slam_handle_t h, long long* out /* 3 */);
/* tuning / test knobs.  key 0: force the shared-memory landmark capacity of the first pass (0 = automatic;
 * instances that do not fit are drained by the full-capacity retry pass); key 1: headroom (landmarks) added to
 * the stale max(M) hint of the per-step launches; key 2: CTA width of the EKF kernels (0 = automatic: by tile size; 32 / 64 / 128 / 256 / 512, and 96 =
 * three filter warps, sweep kernel only); key 3: 1 = slam_run* uses per-step launches instead of the persistent sweep kernel;
 * key 5: steps per sweep-kernel launch (default 48);
 * key 6: headroom of the sweep kernel's tile; key 7: UKF step generation (3 = parallel tridiagonal eigensolver + dense products
 * with its eigenvectors, default; 2 = QL rotation log replayed on the vectors, warp per instance; 1 = explicit eigenvector matrix,
 * CTA per instance); key 12: largest cluster of close eigenvalues the generation-3 eigensolver re-orthogonalises itself (larger
 * clusters send the instance down the generation-2 route in the same step; 1 forces that route for any cluster); key 13: 0 = the
 * one-warp-per-instance back kernel of generation 3 instead of the multi-warp one; key 14: 1 = slam_step_io always stages its
 * buffers through device copies, even when they are pinned host memory the kernels could read in place; key 15: where the generation-3
 * eigensolver builds its eigenvectors: 1 (default) = a shared-memory tile for every size class, 0 = the global scratch (pivots
 * prefetched in batches ahead of the division chains), 2 = the tile for the classes that fit three times per SM, the global
 * scratch beyond (measured equal to 1 on BASELINE configs[2]); key 16: 1 = (test knob) the tile
 * kernel hands every instance with a cluster of close eigenvalues to the global-scratch kernel, as if it needed the refinement step;
 * key 17: storage of the UKF tridiagonalisation in shared memory: 2 (default) = the full square while it fits four times per
 * SM, the packed lower triangle beyond; 1 = packed for every size; 0 = full square for every size; key 8: capacity of the UKF rotation log
 * (shrinking it forces the rescue pass); key 9: max clipped eigenvectors riding beside the first S-pass; key 10: slices of the UKF batch that run their
 * chains of launches on separate streams (1..8; 0 = automatic, the default: 2 for generation 3 from 2048 instances, else 1); key 11: narrow-tile passes of the UKF back kernel before the full-width
 * one: 1 (default) = a 12-column pass, 2 = an 8-column pass before it (measured 1 % slower on BASELINE configs[2]), 0 = none.  The environment variable
 * SLAM_TUNE="key=value,..." applies the same settings to every handle the process creates (a measuring aid).
 * Results never depend on any of them (keys 7-9: up to rounding, inside the parity tolerance). */
int  slam_tune(slam_handle_t h, int key, int value);

#ifdef __cplusplus
}
#endif
#endif /* SLAM_FILTER_H */
