/*
 * slam_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the filter predict/update hot path of kevin-robb/live_ekf_slam
 * (EKF-SLAM, UKF-SLAM, and the simulator's measurement generator).  It exists so the
 * CUDA path can be checked against something; it is NOT part of the shipped library.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
 * path, and its own filters cannot be compiled here (Eigen, ROS, yaml-cpp, GTSAM and
 * SE-Sync are absent).  Truth is therefore this restatement, written line by line from
 * the cited reference sources, cross-checked by an independent NumPy restatement
 * (oracle/oracle_np.py) and by hand-derived known-answer tests (tests/test_oracle_kat.py).
 *
 * Reference files restated (paths under ekf_ws/src/):
 *   localization_pkg/include/localization_pkg/filter.h:79-121   fields, readCommonParams
 *   localization_pkg/src/ekf.cpp:4-34,37-179,182-220            EKF
 *   localization_pkg/src/ukf.cpp:3-45,106-371                   UKF (SLAM and localisation-only branches)
 *   localization_pkg/include/localization_pkg/filter.h:325-370  NaiveFilter
 *   base_pkg/src/sim_node.py:209-250                            measurement generator
 *   base_pkg/src/sim_node.py:63-152                             command trajectory generator
 */
#ifndef SLAM_ORACLE_H
#define SLAM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* yaml-level parameters (BP/config/params.yaml) -- same field order as slam_params in
 * include/slam_filter.h so the Python side can fill both from one dict. */
typedef struct oracle_params {
    float  v_d, v_th;              /* process_noise.mean   filter.h:108-109 */
    float  w_r, w_b;               /* sensing_noise.mean   filter.h:114-115 */
    double V_00, V_11;             /* process_noise.cov    filter.h:110-111 */
    double W_00, W_11;             /* sensing_noise.cov    filter.h:116-117 */
    int    landmark_id_is_known;   /* filter.h:119 */
    float  min_landmark_separation;/* filter.h:120 */
    int    compat_noise_bug;       /* 1 = reproduce filter.h:116-117 (V<-W covs, W=I) */
    /* simulator constraints (params.yaml:27-32) */
    double d_max, th_max, range_max, fov_min, fov_max;
} oracle_params;

enum { ORACLE_EKF_SLAM = 1, ORACLE_UKF_LOC = 2, ORACLE_UKF_SLAM = 3, ORACLE_NAIVE = 5 };   /* FilterChoice filter.h:44-51 */
enum { ORACLE_OK = 0, ORACLE_ERR_NAN = 1, ORACLE_ERR_SAME_STEP_REMATCH = 2, ORACLE_ERR_CAPACITY = 4, ORACLE_ERR_BAD_ID = 16 };
/* evaluation mode: 0 = dense-faithful (the reference's literal O(n^3) products; CPU baseline)
 *                  1 = structured (identical arithmetic with exact-zero terms skipped) */
enum { ORACLE_DENSE = 0, ORACLE_STRUCTURED = 1 };

typedef struct oracle_filter oracle_filter;

void  oracle_set_trig_mode(int mode);   /* 0 = pinned (float)cos((double)x) [D-1], 1 = libm cosf/sinf */
oracle_filter* oracle_create(int kind, const oracle_params* p, int max_landmarks);
void  oracle_destroy(oracle_filter* f);
void  oracle_init(oracle_filter* f, float x0, float y0, float yaw0);
/* UKF_LOC: the true map as the /truth/landmarks wire format, float32 [id, x, y]* with id == index (filter.h:68) */
void  oracle_set_map(oracle_filter* f, const float* map_id_x_y, int n_landmarks);
/* one reference Filter::update(): predict + all landmark updates/insertions */
int   oracle_update(oracle_filter* f, float fwd, float ang, const float* meas, int n_meas, int mode);
/* split form (EKF only): predict alone, then update alone */
int   oracle_predict(oracle_filter* f, float fwd, float ang, int mode);
int   oracle_measure(oracle_filter* f, const float* meas, int n_meas, int mode);
/* teacher forcing: overwrite the committed state */
void  oracle_set_state(oracle_filter* f, const double* x, const double* P_rowmajor, const int* ids, int M, int timestep);

int   oracle_state_dim(const oracle_filter* f);     /* 3+2M (EKF) or 4+2M (UKF) */
int   oracle_num_landmarks(const oracle_filter* f);
int   oracle_timestep(const oracle_filter* f);
int   oracle_status(const oracle_filter* f);
void  oracle_get_state(const oracle_filter* f, double* x);          /* raw x_t, state_dim entries */
void  oracle_get_cov(const oracle_filter* f, double* P_rowmajor);   /* n*n, row-major */
void  oracle_get_landmark_ids(const oracle_filter* f, int* ids);
int   oracle_get_assoc_log(const oracle_filter* f, int* idx, int cap); /* per measurement of the last step: slot index or -1 (new) */
int   oracle_sigma_rows(const oracle_filter* f);                              /* rows of X (ukf.cpp:169): n of the last step's prior */
void  oracle_get_sigma_points(const oracle_filter* f, double* X_colmajor); /* UKF: n*(2n+1), ukf.cpp:91-99 order */

/* Philox4x32-10 counter RNG shared (by definition) with the GPU workload source */
void   oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]);
double oracle_uniform(uint32_t hi, uint32_t lo);

/* simulator measurement generator, sim_node.py:209-250.  truth[3] is advanced in place.
 * meas_out receives float32 [id, r, b]* ; returns the number of detections (<= cap). */
int   oracle_sim_step(const oracle_params* p, double truth[3], float fwd, float ang,
                      const double* lm_xy, int n_lm, uint64_t seed, uint32_t instance, uint32_t step,
                      float* meas_out, int cap);

/* generate_trajectory, sim_node.py:63-152: the precomputed command trajectory of one Monte-Carlo instance
 * (noisy map copy keyed (seed; instance, id, 0, 1), nearest-neighbour tour, clamped commands); float32 wire values out. */
int   oracle_make_map(int map_type, int n_landmarks, double bound, double grid_step, double min_sep, uint64_t seed, uint32_t instance,
                      double* lm_xy, int cap);   /* sim_node.py:155-206: 0 grid, 1 random; returns n or < 0 */
int   oracle_tsp_trajectory(const oracle_params* p, const double* lm_xy, int n_lm, double landmark_noise,
                            double visitation_threshold, double bound, double x0, double y0, double yaw0, int T,
                            uint64_t seed, uint32_t instance, float* fwd_out, float* ang_out);

/* whole Monte-Carlo instance: simulator + filter for T steps (used for traces and CPU timing).
 * pose_trace (optional) receives T*3 doubles (x,y,yaw estimate); truth_trace (optional) T*3. */
int   oracle_run_instance(int kind, const oracle_params* p, const double* lm_xy, int n_lm,
                          const float* cmd_fwd, const float* cmd_ang, int T,
                          uint64_t seed, uint32_t instance, int max_landmarks, int mode,
                          double* pose_trace, double* truth_trace, oracle_filter** keep);
/* CPU baseline: n_threads pthreads, each running `per_thread` instances back to back.
 * Returns wall seconds; *updates receives instances*T. */
double oracle_bench(int kind, const oracle_params* p, const double* lm_xy, int n_lm,
                    const float* cmd_fwd, const float* cmd_ang, int T, uint64_t seed,
                    int n_threads, int per_thread, int max_landmarks, int mode, long long* updates);

/* symmetric eigen-decomposition used by the UKF restatement (exposed for tests) */
void  oracle_eigh(const double* A_rowmajor, int n, double* evals, double* evecs_rowmajor);

#ifdef __cplusplus
}
#endif
#endif
