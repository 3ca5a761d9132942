// capi.cu -- the C-ABI of include/slam_filter.h: handle lifetime, host<->device staging and kernel launches.
// No compute happens on the host; every entry point that advances a filter launches sm_100a kernels.
#include "common.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace slam;

static thread_local std::string g_create_error;

struct slam_filter {
    int kind = 0;
    int device = 0;
    slam_params params{};
    FilterConst fc{};
    SimConst sc{};
    BatchState b{};
    cudaStream_t stream = nullptr;
    // staging for host-pointer entry points
    float* d_fwd = nullptr; float* d_ang = nullptr; float* d_meas = nullptr; int* d_nmeas = nullptr;
    float* d_traj_fwd = nullptr; float* d_traj_ang = nullptr; size_t traj_cap = 0;
    double* d_out = nullptr;      // scratch for stats / poses
    size_t d_out_cap = 0;
    long long launches = 0;
    // single large-map instance (P in HBM, deferred rank-2k DMMA update): csrc/ekf_large.cu
    bool large = false;
    LargeState lg{};
    UkfScratch uk{};                  // HBM scratch between the three launches of a UKF step
    float* d_map = nullptr;           // UKF_LOC: device copy of the true map
    int map_cap = 0;
    UkfStreams uks{};                 // slices of the batch run front -> QL -> back on their own streams
    int* h_nmeas_pin = nullptr;       // pinned scratch for the host-side measurement count
    // per-launch timing of the filter-step kernel
    // capacity hint: device max(M) read back with a lag of HINT_LAG launches (never waited on in steady state)
    static constexpr int HINT_RING = 16, HINT_LAG = 8;
    int* h_hint = nullptr;            // pinned [HINT_RING]
    cudaEvent_t hint_ev[HINT_RING] = {};
    long long step_seq = 0;           // filter-step launches since the last init/reset
    int hint_base = 0;                // max(M) known on the host at step_seq == 0
    int cap_headroom = 4;             // landmarks of slack on top of the stale max(M)
    int cap_force = 0;                // > 0: force this capacity for the first pass (tests of the retry path)
    int force_threads = 0;            // > 0: CTA width of the EKF kernels (tuning)
    int sweep_off = 0;                // 1: slam_run* always uses per-step launches (tests compare both paths)
    int sweep_chunk = 48;             // steps per launch of the persistent sweep kernel (B200, configs[1]: 16 / 24 / 32 / 48 / 64 steps -> 123 / 128 / 130 / 133 / 129 M updates/s)
    int sweep_headroom = 8;           // landmarks of slack on top of the stale max(M) when sizing a chunk's tile
    int* d_work = nullptr;            // [2] work counters of the sweep kernel (tile-sized launch, full-capacity launch)
    int* d_progress = nullptr;        // [batch] run-relative steps completed (chunk gate of the sweep kernel)
    int* h_run_hint = nullptr;        // pinned [HINT_RING]: max(M) after each chunk of the current run
    cudaEvent_t run_ev[HINT_RING] = {};
    // trajectory replay through host buffers (slam_run_io): two copy streams, double-buffered device chunks
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    float* d_rmeas[2] = {nullptr, nullptr}; int* d_rn[2] = {nullptr, nullptr}; double* d_rposes[2] = {nullptr, nullptr};
    int r_chunk_cap = 0;              // steps the replay buffers hold
    cudaEvent_t ev_h2d[2] = {}, ev_comp[2] = {}, ev_d2h[2] = {};
    static constexpr int MAX_MAPPED = 8;          // verified pinned-host ranges (slam_step_io zero-copy path)
    uintptr_t mapped_lo[MAX_MAPPED] = {}, mapped_hi[MAX_MAPPED] = {};
    int n_mapped = 0, mapped_next = 0;
    int no_zero_copy = 0;                          // slam_tune key 14: force the staged path (tests compare both)
    unsigned long long* d_hist = nullptr; int hist_cap = 0;   // scratch of slam_get_error_histogram (grown on demand, kept)
    double* d_avg = nullptr;
    bool profiling = false;           // per-launch events around the per-step filter kernel
    bool profiling_sweep = false;     // events around the persistent sweep kernel
    bool profiling_gemm = false;      // large map: events around lm_gemm only
    std::vector<cudaEvent_t> ev;      // pairs
    size_t ev_used = 0;
    std::string err;
};

struct slam_sim {
    slam_filter* owner = nullptr;
    SimState s{};
    double* d_lm = nullptr;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) h->err = buf_; else g_create_error = buf_;                                      \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

static int fail(slam_filter* h, const char* msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return 1;
}

extern "C" {

const char* slam_last_error(slam_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int slam_create(int kind, const slam_params* params, int batch, int max_landmarks, int max_meas, int device,
                slam_handle_t* out) {
    slam_filter* h = nullptr;
    if (!out) return fail(h, "slam_create: out is NULL");
    *out = nullptr;
    if (kind != SLAM_EKF_SLAM && kind != SLAM_UKF_SLAM && kind != SLAM_UKF_LOC && kind != SLAM_NAIVE)
        return fail(h, "Invalid filter choice (expected SLAM_EKF_SLAM, SLAM_UKF_SLAM, SLAM_UKF_LOC or SLAM_NAIVE).");   // localization_node.cpp:44
    const int map_size = max_landmarks;
    if (kind == SLAM_UKF_LOC || kind == SLAM_NAIVE) max_landmarks = 1;      // the state never holds landmarks
    if (!params) return fail(h, "slam_create: params is NULL");
    if (batch < 1 || max_landmarks < 1 || max_meas < 1) return fail(h, "slam_create: batch, max_landmarks and max_meas must be >= 1");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(h, "slam_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(h, "slam_create: bad device ordinal");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(h, "slam_create: kernels are built for sm_100a (B200) only");

    h = new slam_filter();
    h->kind = kind; h->device = device; h->params = *params;
    // Filter::readCommonParams, filter.h:105-121 (the V/W mix-up is reproduced when compat_noise_bug != 0)
    FilterConst& fc = h->fc;
    fc.V00 = params->V_00; fc.V11 = params->V_11; fc.W00 = 1.0; fc.W11 = 1.0;
    if (params->compat_noise_bug) { fc.V00 = params->W_00; fc.V11 = params->W_11; }
    else { fc.W00 = params->W_00; fc.W11 = params->W_11; }
    fc.v_d = params->v_d; fc.v_th = params->v_th; fc.w_r = params->w_r; fc.w_b = params->w_b;
    fc.min_sep = params->min_landmark_separation; fc.id_known = params->landmark_id_is_known;
    fc.loc = (kind == SLAM_UKF_LOC); fc.n_map = 0; fc.map = nullptr;
    h->map_cap = map_size;
    SimConst& sc = h->sc;
    sc.V_00 = params->V_00; sc.V_11 = params->V_11; sc.W_00 = params->W_00; sc.W_11 = params->W_11;
    sc.d_max = params->d_max; sc.th_max = params->th_max; sc.range_max = params->range_max;
    sc.fov_min = params->fov_min; sc.fov_max = params->fov_max;

    BatchState& b = h->b;
    b.batch = batch; b.max_lm = max_landmarks; b.max_meas = max_meas;
    b.base = (kind == SLAM_EKF_SLAM || kind == SLAM_NAIVE) ? 3 : 4;
    b.n_max = b.base + 2 * max_landmarks;
    b.lds = lds_of(b.n_max);
    b.x_stride = ldg_of(b.n_max);
    b.p_stride = (long long)b.n_max * ldg_of(b.n_max);
    b.sigma_stride = 0;
    b.fixed_ld = ldg_of(b.n_max);      // row-major layouts (UKF, large map): rows of P keep a fixed stride in HBM
    b.ps2g = 0;

    const size_t smem = (kind == SLAM_EKF_SLAM || kind == SLAM_NAIVE) ? ekf_step_smem_bytes(b) : ukf_step_smem_bytes(b);
    if (kind == SLAM_UKF_LOC && !ukf_gen2_supported(b)) { delete h; h = nullptr; return fail(h, "slam_create: UKF_LOC needs max_meas <= 14"); }
    if (smem > (size_t)prop.sharedMemPerBlockOptin) {
        if (kind == SLAM_EKF_SLAM && batch == 1 && max_meas <= 256) {
            h->large = true;                       // P stays in HBM: large-map path
            b.fixed_ld = ldg_of(b.n_max);
        } else {
            delete h; h = nullptr;
            return fail(h, "slam_create: max_landmarks too large for the shared-memory-resident batch kernels "
                           "(the HBM-resident large-map path needs kind = EKF_SLAM, batch = 1, max_meas <= 256)");
        }
    }
    if (kind == SLAM_EKF_SLAM && !h->large) {
        // batched EKF: packed symmetric covariance, two planes of A_max (A_max + 1) doubles (common.cuh: bpl_idx)
        b.ps2g = bpl_plane_doubles(2 + b.max_lm);
        b.p_stride = 2LL * b.ps2g;
    }
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
    CK(cudaMalloc(&b.P, sizeof(double) * (size_t)batch * b.p_stride));
    CK(cudaMalloc(&b.x, sizeof(double) * (size_t)batch * b.x_stride));
    CK(cudaMalloc(&b.ids, sizeof(int) * (size_t)batch * b.max_lm));
    CK(cudaMalloc(&b.meta, sizeof(int4) * batch));
    CK(cudaMalloc(&b.assoc, sizeof(int) * (size_t)batch * b.max_meas));
    CK(cudaMalloc(&b.retry_list, sizeof(int) * batch));
    CK(cudaMalloc(&b.retry_count, 2 * sizeof(int)));
    CK(cudaMalloc(&b.max_M, sizeof(int)));
    CK(cudaMalloc(&h->d_work, 2 * sizeof(int)));
    CK(cudaMalloc(&h->d_progress, sizeof(int) * batch));
    CK(cudaMemset(h->d_progress, 0, sizeof(int) * batch));
    CK(cudaMallocHost(&h->h_run_hint, sizeof(int) * slam_filter::HINT_RING));
    for (int i = 0; i < slam_filter::HINT_RING; ++i) CK(cudaEventCreateWithFlags(&h->run_ev[i], cudaEventDisableTiming));
    CK(cudaMemset(b.retry_count, 0, 2 * sizeof(int)));
    CK(cudaMemset(b.max_M, 0, sizeof(int)));
    CK(cudaHostAlloc(&h->h_hint, sizeof(int) * slam_filter::HINT_RING, cudaHostAllocMapped));
    b.hint_host = nullptr;
    if (kind == SLAM_EKF_SLAM) {      // the batched EKF's retry pass posts max(M) into h_hint[0] itself
        int* dptr = nullptr;
        if (cudaHostGetDevicePointer(&dptr, h->h_hint, 0) == cudaSuccess) b.hint_host = dptr; else cudaGetLastError();
    }
    for (int i = 0; i < slam_filter::HINT_RING; ++i) { h->h_hint[i] = 0; CK(cudaEventCreateWithFlags(&h->hint_ev[i], cudaEventDisableTiming)); }
    CK(cudaMalloc(&b.stats, sizeof(double) * (size_t)batch * SLAM_NUM_STATS));
    if (kind == SLAM_UKF_SLAM || kind == SLAM_UKF_LOC) {
        b.sigma_stride = (long long)b.n_max * (2 * b.n_max + 1);
        b.sigma = nullptr;   // allocated lazily by slam_get_sigma_points
        UkfScratch& u = h->uk;
        u.n_max = b.n_max;
        u.gen = 3; u.clip_lanes = 0; u.maxc = 16; u.multiwarp = 1; u.eig3_tile = 1; u.refine_all = 0; u.front_packed = 2;
        h->uks.nsub = 0;                            // automatic (slam_tune key 10)
        CK(cudaEventCreateWithFlags(&h->uks.fork, cudaEventDisableTiming));
        for (int k = 0; k < UKF_MAX_SUB - 1; ++k) {
            CK(cudaStreamCreateWithFlags(&h->uks.aux[k], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&h->uks.join[k], cudaEventDisableTiming));
        }
        u.rot_cap = 2LL * b.n_max * b.n_max;        // ~0.85 n^2 rotations are typical
        if (u.rot_cap < 256) u.rot_cap = 256;
        u.swp_cap = 6 * b.n_max;                    // ~1.7 n sweeps are typical
        CK(cudaMalloc(&u.Zg, sizeof(double) * (size_t)batch * b.n_max * b.n_max));
        CK(cudaMalloc(&u.Yg, sizeof(double) * (size_t)batch * b.n_max * b.n_max));
        CK(cudaMalloc(&u.Vg, sizeof(double) * (size_t)batch * b.n_max * b.n_max));
        CK(cudaMalloc(&u.VTg, sizeof(double) * (size_t)batch * b.n_max * b.n_max));
        CK(cudaMalloc(&u.dg, sizeof(double) * (size_t)batch * b.n_max));
        CK(cudaMalloc(&u.eg, sizeof(double) * (size_t)batch * b.n_max));
        CK(cudaMalloc(&u.rot, sizeof(double2) * (size_t)batch * u.rot_cap));
        CK(cudaMalloc(&u.swp, sizeof(int2) * (size_t)batch * u.swp_cap));
        CK(cudaMalloc(&u.nswp, sizeof(int) * batch));
        CK(cudaMalloc(&u.defer, sizeof(int) * batch));
        CK(cudaMemset(u.defer, 0, sizeof(int) * batch));
        CK(cudaMalloc(&u.xprior, sizeof(double) * (size_t)batch * b.n_max));
        CK(cudaMalloc(&u.sigfmt, sizeof(int2) * batch));
        CK(cudaMalloc(&u.routes, sizeof(unsigned long long) * 4));
        CK(cudaMemset(u.routes, 0, sizeof(unsigned long long) * 4));
        CK(cudaMemset(u.sigfmt, 0, sizeof(int2) * batch));
        u.narrow = 1;
    }
    CK(cudaMalloc(&h->d_fwd, sizeof(float) * batch));
    CK(cudaMalloc(&h->d_ang, sizeof(float) * batch));
    CK(cudaMalloc(&h->d_meas, sizeof(float) * 3 * (size_t)batch * max_meas));
    CK(cudaMalloc(&h->d_nmeas, sizeof(int) * batch));
    h->d_out_cap = sizeof(double) * (size_t)(3 * batch + SLAM_NUM_STATS);
    CK(cudaMalloc(&h->d_out, h->d_out_cap));
    if (h->large) {
        LargeState& g = h->lg;
        g.P = b.P; g.x = b.x; g.ids = b.ids; g.meta = b.meta; g.assoc = b.assoc; g.stats = b.stats;
        g.ld = b.fixed_ld; g.n_max = b.n_max; g.max_lm = b.max_lm; g.max_meas = b.max_meas;
        CK(cudaMalloc(&g.xp, sizeof(double) * b.x_stride));
        CK(cudaMalloc(&g.U, sizeof(double) * (size_t)b.max_meas * b.n_max * 2));
        CK(cudaMalloc(&g.G, sizeof(double) * (size_t)b.max_meas * 2 * g.ld));
        CK(cudaMalloc(&g.ctl, sizeof(int) * 4 * b.max_meas));
        CK(cudaMalloc(&g.cur, sizeof(int) * 16));      // [0..3] step bookkeeping, [8] grid-barrier counter
        CK(cudaMalloc(&g.sc, sizeof(double) * 16));
        CK(cudaMemset(g.cur, 0, sizeof(int) * 16));
        CK(cudaMallocHost(&h->h_nmeas_pin, sizeof(int)));
        CK(ekf_large_configure());
    } else if (kind == SLAM_EKF_SLAM) {
        CK(ekf_step_configure(b));
    } else if (kind != SLAM_NAIVE) CK(ukf_step_configure(b));
    *out = h;
    // development aid: SLAM_TUNE="key=value,key=value" applies slam_tune() settings to every handle of the process, so that one
    // binary can be A/B-measured through an unmodified caller (bench.py).  Results never depend on the keys (see slam_tune).
    if (const char* env = getenv("SLAM_TUNE")) {
        const char* p = env;
        while (*p) {
            char* end = nullptr;
            const long key = strtol(p, &end, 10);
            if (end == p || *end != '=') break;
            p = end + 1;
            const long value = strtol(p, &end, 10);
            if (end == p) break;
            if (slam_tune(h, (int)key, (int)value) != 0) return 1;
            p = (*end == ',') ? end + 1 : end;
            if (*end != ',' ) break;
        }
    }
    return slam_init(h, 0.f, 0.f, 0.f);
}

int slam_set_map(slam_handle_t h, const float* map_id_x_y, int n_landmarks) {
    if (!h) return 1;
    if (h->kind != SLAM_UKF_LOC) return fail(h, "slam_set_map: only the localisation-only UKF keeps the true map (filter.h:68)");
    if (!map_id_x_y || n_landmarks < 0) return fail(h, "slam_set_map: bad arguments");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_map); h->d_map = nullptr;
    CK(cudaMalloc(&h->d_map, sizeof(float) * 3 * (size_t)(n_landmarks > 0 ? n_landmarks : 1)));
    CK(cudaMemcpy(h->d_map, map_id_x_y, sizeof(float) * 3 * (size_t)n_landmarks, cudaMemcpyHostToDevice));
    h->fc.map = h->d_map; h->fc.n_map = n_landmarks;
    return 0;
}

int slam_destroy(slam_handle_t h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    BatchState& b = h->b;
    cudaFree(b.P); cudaFree(b.x); cudaFree(b.ids); cudaFree(b.meta);
    cudaFree(b.assoc); cudaFree(b.stats); cudaFree(b.sigma);
    cudaFree(h->d_map); cudaFree(h->d_hist); cudaFree(h->d_avg);
    cudaFree(h->uk.Zg); cudaFree(h->uk.Yg); cudaFree(h->uk.Vg); cudaFree(h->uk.VTg); cudaFree(h->uk.dg); cudaFree(h->uk.eg); cudaFree(h->uk.rot); cudaFree(h->uk.swp); cudaFree(h->uk.nswp); cudaFree(h->uk.defer); cudaFree(h->uk.xprior); cudaFree(h->uk.sigfmt); cudaFree(h->uk.routes);
    cudaFree(b.retry_list); cudaFree(b.retry_count); cudaFree(b.max_M); cudaFree(h->d_work); cudaFree(h->d_progress);
    if (h->h_run_hint) cudaFreeHost(h->h_run_hint);
    for (int i = 0; i < slam_filter::HINT_RING; ++i) if (h->run_ev[i]) cudaEventDestroy(h->run_ev[i]);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->d_rmeas[i]); cudaFree(h->d_rn[i]); cudaFree(h->d_rposes[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
    }
    if (h->s_h2d) { cudaStreamSynchronize(h->s_h2d); cudaStreamDestroy(h->s_h2d); }
    if (h->s_d2h) { cudaStreamSynchronize(h->s_d2h); cudaStreamDestroy(h->s_d2h); }
    if (h->large) { cudaFree(h->lg.xp); cudaFree(h->lg.U); cudaFree(h->lg.G); cudaFree(h->lg.ctl); cudaFree(h->lg.cur); cudaFree(h->lg.sc); }
    if (h->h_nmeas_pin) cudaFreeHost(h->h_nmeas_pin);
    if (h->h_hint) cudaFreeHost(h->h_hint);
    for (int i = 0; i < slam_filter::HINT_RING; ++i) if (h->hint_ev[i]) cudaEventDestroy(h->hint_ev[i]);
    cudaFree(h->d_fwd); cudaFree(h->d_ang); cudaFree(h->d_meas); cudaFree(h->d_nmeas);
    cudaFree(h->d_traj_fwd); cudaFree(h->d_traj_ang); cudaFree(h->d_out);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->uks.fork) cudaEventDestroy(h->uks.fork);
    for (int k = 0; k < UKF_MAX_SUB - 1; ++k) {
        if (h->uks.aux[k]) { cudaStreamSynchronize(h->uks.aux[k]); cudaStreamDestroy(h->uks.aux[k]); }
        if (h->uks.join[k]) cudaEventDestroy(h->uks.join[k]);
    }
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

void* slam_stream(slam_handle_t h) { return h ? (void*)h->stream : nullptr; }
int slam_synchronize(slam_handle_t h) { if (!h) return 1; CK(cudaStreamSynchronize(h->stream)); return 0; }
int slam_batch(slam_handle_t h) { return h ? h->b.batch : 0; }
int slam_kind(slam_handle_t h) { return h ? h->kind : 0; }
long long slam_kernel_launches(slam_handle_t h) { return h ? h->launches : 0; }

int slam_tune(slam_handle_t h, int key, int value) {
    if (!h) return 1;
    if (key == 0) h->cap_force = value;
    else if (key == 1) h->cap_headroom = value;
    else if (key == 2) {
        if (value != 0 && value != 32 && value != 64 && value != 96 && value != 128 && value != 256 && value != 512)
            return fail(h, "slam_tune: CTA width must be 0 (automatic), 32, 64, 96 (sweep kernel only), 128, 256 or 512");
        h->force_threads = value;
    } else if (key == 3) h->sweep_off = value;
    else if (key == 5) { if (value < 1) return fail(h, "slam_tune: the sweep chunk must be >= 1 step"); h->sweep_chunk = value; }
    else if (key == 6) h->sweep_headroom = value < 0 ? 0 : value;
    else if (key == 7) { if (value < 1 || value > 3) return fail(h, "slam_tune: UKF generation must be 1, 2 or 3"); h->uk.gen = value; }
    else if (key == 12) h->uk.maxc = value < 0 ? 0 : value;
    else if (key == 13) h->uk.multiwarp = value ? 1 : 0;
    else if (key == 15) h->uk.eig3_tile = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (key == 16) h->uk.refine_all = value ? 1 : 0;
    else if (key == 17) h->uk.front_packed = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (key == 14) h->no_zero_copy = value ? 1 : 0;
    else if (key == 8) {     // shrink the rotation log (test knob: forces the rescue pass); never beyond the allocation
        long long full = 2LL * h->b.n_max * h->b.n_max;
        if (full < 256) full = 256;
        h->uk.rot_cap = (value <= 0 || value > full) ? full : value;
    } else if (key == 9) h->uk.clip_lanes = value < 0 ? 0 : value;
    else if (key == 11) h->uk.narrow = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (key == 10) { if (value < 0 || value > UKF_MAX_SUB) return fail(h, "slam_tune: UKF slices must be 0 (automatic) or 1..8"); h->uks.nsub = value; }
    else return fail(h, "slam_tune: unknown key");
    return 0;
}

int slam_get_ukf_routes(slam_handle_t h, long long* out) {
    if (!h || !out) return 1;
    if (!h->uk.routes) return fail(h, "slam_get_ukf_routes: UKF only");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, h->uk.routes, sizeof(long long) * 3, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int slam_build_info(char* buf, int cap) {
    return snprintf(buf, cap, "sm_100a; nvcc %d.%d; kernels: ekf_step_kernel ukf_step_kernel sim_step_kernel", __CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__);
}

// Filter::init, ekf.cpp:29-34 / ukf.cpp:31-45 together with the constructors' P_0 (ekf.cpp:8-18, ukf.cpp:7-18)
int slam_init(slam_handle_t h, float x_0, float y_0, float yaw_0) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    BatchState& b = h->b;
    double a2 = yaw_0, a3 = 0.0;
    if (b.base == 4) { a2 = (double)(float)std::cos((double)yaw_0); a3 = (double)(float)std::sin((double)yaw_0); }   // ukf.cpp:33 (D-1)
    CK(cudaMemsetAsync(b.P, 0, sizeof(double) * (size_t)b.batch * b.p_stride, h->stream));
    CK(cudaMemsetAsync(b.x, 0, sizeof(double) * (size_t)b.batch * b.x_stride, h->stream));
    CK(launch_reset(b, (double)x_0, (double)y_0, a2, a3, h->stream));      // x_0, P_0, meta, stats on the device
    h->launches += 1;
    CK(cudaMemsetAsync(b.ids, 0, sizeof(int) * (size_t)b.batch * b.max_lm, h->stream));
    CK(cudaMemsetAsync(b.max_M, 0, sizeof(int), h->stream));
    h->step_seq = 0; h->hint_base = 0;
    if (h->h_hint) h->h_hint[0] = 0;
    if (h->uk.sigfmt) CK(cudaMemsetAsync(h->uk.sigfmt, 0, sizeof(int2) * b.batch, h->stream));
    CK(cudaMemsetAsync(b.assoc, 0xff, sizeof(int) * (size_t)b.batch * b.max_meas, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

static int do_step(slam_filter* h, const float* d_fwd, const float* d_ang, int cmd_stride, const float* d_meas,
                   const int* d_nmeas, int phases, double* fused_poses = nullptr) {
    CK(cudaSetDevice(h->device));
    if (h->large) {
        if (phases != (STEP_PREDICT | STEP_UPDATE)) return fail(h, "split predict/update is not available on the large-map path");
        // (the detection count stays on the device: the step is asynchronous like every other path)
        if (h->profiling || h->profiling_gemm) {
            if (h->ev_used + 2 > h->ev.size()) {
                const size_t old = h->ev.size();
                h->ev.resize(old + 4096);
                for (size_t i = old; i < h->ev.size(); ++i) CK(cudaEventCreate(&h->ev[i]));
            }
            if (h->profiling) CK(cudaEventRecord(h->ev[h->ev_used], h->stream));
        }
        const bool pg = h->profiling_gemm;                // events around the closing DMMA contraction only
        CK(launch_ekf_large_step(h->lg, h->fc, d_fwd, d_ang, d_meas, d_nmeas, h->b.n_max, h->stream, &h->launches,
                                 pg ? h->ev[h->ev_used] : nullptr, pg ? h->ev[h->ev_used + 1] : nullptr));
        if (h->profiling) CK(cudaEventRecord(h->ev[h->ev_used + 1], h->stream));
        if (h->profiling || pg) h->ev_used += 2;
        return 0;
    }
    StepInputs in{d_fwd, d_ang, cmd_stride, d_meas, d_nmeas, fused_poses};
    if (h->kind != SLAM_EKF_SLAM && phases != (STEP_PREDICT | STEP_UPDATE))
        return fail(h, "split predict/update is defined for EKF_SLAM only: the UKF update stage consumes the sigma points of the same call (ukf.cpp:305-337)");
    if (h->profiling) {
        if (h->ev_used + 2 > h->ev.size()) {
            const size_t old = h->ev.size();
            h->ev.resize(old + 4096);
            for (size_t i = old; i < h->ev.size(); ++i) CK(cudaEventCreate(&h->ev[i]));
        }
        CK(cudaEventRecord(h->ev[h->ev_used], h->stream));
    }
    // capacity for this launch: max(M) as it was HINT_LAG launches ago (its copy has long completed, so the wait
    // below never blocks in steady state but bounds how far the host can run ahead) plus headroom
    const bool posted_hint = h->kind == SLAM_EKF_SLAM && !h->large && h->b.hint_host != nullptr;
    int cap = h->b.max_lm;
    if (h->cap_force > 0) cap = h->cap_force;
    else if (h->step_seq >= slam_filter::HINT_LAG) {
        if (posted_hint) {
            // max(M) as the retry pass of an earlier launch posted it into mapped host memory (stale values are conservative:
            // too small a tile only sends instances to the retry pass).  Every 8th launch leaves an event behind and waits for
            // the one of 16 launches ago, which bounds how far the host can run ahead of the device.
            if ((h->step_seq & 7) == 0) {
                const int slot = (int)((h->step_seq >> 3) % slam_filter::HINT_RING);
                if (h->step_seq >= 24) CK(cudaEventSynchronize(h->hint_ev[(slot + slam_filter::HINT_RING - 2) % slam_filter::HINT_RING]));
            }
            cap = *(volatile int*)&h->h_hint[0] + h->cap_headroom;
        } else {
            const int slot = (int)((h->step_seq - slam_filter::HINT_LAG) % slam_filter::HINT_RING);
            CK(cudaEventSynchronize(h->hint_ev[slot]));
            cap = h->h_hint[slot] + h->cap_headroom;
        }
    } else cap = h->hint_base + (int)(h->step_seq + 1) * h->b.max_meas;      // M can grow by at most max_meas per step
    if (h->kind == SLAM_EKF_SLAM) {
        CK(launch_ekf_step(h->b, h->fc, in, phases, cap, h->force_threads, h->stream));
    }
    else if (h->kind == SLAM_NAIVE) { CK(launch_naive_step(h->b, in, h->stream)); h->launches += 1; }
    else { int nl = 0; CK(launch_ukf_step(h->b, h->fc, in, h->uk, h->stream, h->uks, &nl)); h->launches += nl; }
    if (h->profiling) { CK(cudaEventRecord(h->ev[h->ev_used + 1], h->stream)); h->ev_used += 2; }
    if (posted_hint) {
        if ((h->step_seq & 7) == 0) CK(cudaEventRecord(h->hint_ev[(int)((h->step_seq >> 3) % slam_filter::HINT_RING)], h->stream));
        h->step_seq += 1;
    } else {
        const int slot = (int)(h->step_seq % slam_filter::HINT_RING);
        CK(cudaMemcpyAsync(&h->h_hint[slot], h->b.max_M, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->hint_ev[slot], h->stream));
        h->step_seq += 1;
    }
    if (h->kind == SLAM_EKF_SLAM) h->launches += (cap > 0 && cap < h->b.max_lm) ? 2 : 1;      // (a capacity-limited launch is followed by its retry pass)
    return 0;
}

static int stage_cmd(slam_filter* h, const float* fwd, const float* ang, int cmd_stride) {
    const size_t nc = cmd_stride ? (size_t)h->b.batch : 1;
    CK(cudaMemcpyAsync(h->d_fwd, fwd, sizeof(float) * nc, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_ang, ang, sizeof(float) * nc, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
static int stage_meas(slam_filter* h, const float* meas, const int* n_meas) {
    CK(cudaMemcpyAsync(h->d_nmeas, n_meas, sizeof(int) * h->b.batch, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_meas, meas, sizeof(float) * 3 * (size_t)h->b.batch * h->b.max_meas, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int slam_step(slam_handle_t h, const float* fwd, const float* ang, int cmd_stride, const float* meas, const int* n_meas) {
    if (!h) return 1;
    if (!fwd || !ang || !meas || !n_meas) return fail(h, "slam_step: NULL argument");
    CK(cudaSetDevice(h->device));
    if (stage_cmd(h, fwd, ang, cmd_stride) || stage_meas(h, meas, n_meas)) return 1;
    return do_step(h, h->d_fwd, h->d_ang, cmd_stride, h->d_meas, h->d_nmeas, STEP_PREDICT | STEP_UPDATE);
}
int slam_step_device(slam_handle_t h, const float* d_fwd, const float* d_ang, int cmd_stride, const float* d_meas, const int* d_n_meas) {
    if (!h) return 1;
    if (!d_fwd || !d_ang || !d_meas || !d_n_meas) return fail(h, "slam_step_device: NULL argument");
    return do_step(h, d_fwd, d_ang, cmd_stride, d_meas, d_n_meas, STEP_PREDICT | STEP_UPDATE);
}
int slam_predict(slam_handle_t h, const float* fwd, const float* ang, int cmd_stride) {
    if (!h) return 1;
    if (!fwd || !ang) return fail(h, "slam_predict: NULL argument");
    CK(cudaSetDevice(h->device));
    if (stage_cmd(h, fwd, ang, cmd_stride)) return 1;
    return do_step(h, h->d_fwd, h->d_ang, cmd_stride, h->d_meas, h->d_nmeas, STEP_PREDICT);
}
int slam_update(slam_handle_t h, const float* meas, const int* n_meas) {
    if (!h) return 1;
    if (!meas || !n_meas) return fail(h, "slam_update: NULL argument");
    CK(cudaSetDevice(h->device));
    if (stage_meas(h, meas, n_meas)) return 1;
    return do_step(h, h->d_fwd, h->d_ang, 0, h->d_meas, h->d_nmeas, STEP_UPDATE);
}
int slam_predict_device(slam_handle_t h, const float* d_fwd, const float* d_ang, int cmd_stride) {
    if (!h) return 1;
    return do_step(h, d_fwd, d_ang, cmd_stride, h->d_meas, h->d_nmeas, STEP_PREDICT);
}
int slam_update_device(slam_handle_t h, const float* d_meas, const int* d_n_meas) {
    if (!h) return 1;
    return do_step(h, h->d_fwd, h->d_ang, 0, d_meas, d_n_meas, STEP_UPDATE);
}

// ------------------------------------------------------------------------------------------ getters
static int check_inst(slam_filter* h, int inst) {
    if (!h) return 1;
    if (inst < 0 || inst >= h->b.batch) return fail(h, "instance index out of range");
    return 0;
}
enum { META_M = 0, META_STATUS = 1, META_TIMESTEP = 2, META_NASSOC = 3 };
static int get_meta(slam_filter* h, int field, int inst, int* out) {
    if (check_inst(h, inst)) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out, reinterpret_cast<const int*>(h->b.meta + inst) + field, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
static int get_meta_all(slam_filter* h, int field, int* out) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy2DAsync(out, sizeof(int), reinterpret_cast<const int*>(h->b.meta) + field, sizeof(int4), sizeof(int),
                         h->b.batch, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int slam_get_timestep(slam_handle_t h, int inst, int* t) { return get_meta(h, META_TIMESTEP, inst, t); }
int slam_get_num_landmarks(slam_handle_t h, int inst, int* M) { return get_meta(h, META_M, inst, M); }
int slam_get_status(slam_handle_t h, int inst, int* s) { return get_meta(h, META_STATUS, inst, s); }

int slam_get_state(slam_handle_t h, int inst, double* x, int* n) {
    int M = 0;
    if (slam_get_num_landmarks(h, inst, &M)) return 1;
    const int nn = h->b.base + 2 * M;
    CK(cudaMemcpyAsync(x, h->b.x + (size_t)inst * h->b.x_stride, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n) *n = nn;
    return 0;
}

int slam_get_state_vector(slam_handle_t h, int inst, double* xv, int* n) {
    if (check_inst(h, inst)) return 1;
    if (h->kind == SLAM_EKF_SLAM || h->kind == SLAM_NAIVE) return slam_get_state(h, inst, xv, n);   // ekf.cpp:182-185, filter.h:349-352
    std::vector<double> raw(h->b.x_stride);
    int nr = 0;
    if (slam_get_state(h, inst, raw.data(), &nr)) return 1;
    // the (x, y, yaw, landmarks...) vector UKF::getStateVector means to build (ukf.cpp:47-53; the reference's
    // own version resizes a fixed-size Vector3d and is broken for M > 0, SURVEY B-8)
    xv[0] = raw[0]; xv[1] = raw[1];
    xv[2] = std::remainder(std::atan2(raw[3], raw[2]), TWO_PI_REF);
    for (int i = 4; i < nr; ++i) xv[i - 1] = raw[i];
    if (n) *n = nr - 1;
    return 0;
}

int slam_get_cov(slam_handle_t h, int inst, double* P, int* n) {
    int M = 0;
    if (slam_get_num_landmarks(h, inst, &M)) return 1;
    const int nn = h->b.base + 2 * M, ld = ldp_of(h->b.fixed_ld, nn);
    if (h->b.ps2g) {
        // packed symmetric storage -> the full row-major matrix publishState serialises (ekf.cpp:211-217)
        const int live = bpl_plane_doubles(2 + M), ps2 = h->b.ps2g;
        std::vector<double> pk(2 * (size_t)live);
        const double* g = h->b.P + (size_t)inst * h->b.p_stride;
        CK(cudaMemcpyAsync(pk.data(), g, sizeof(double) * live, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(pk.data() + live, g + ps2, sizeof(double) * live, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < nn; ++r)
            for (int c = 0; c < nn; ++c) P[(size_t)r * nn + c] = pk[bpl_sym(r + 1, c + 1, live)];
        if (n) *n = nn;
        return 0;
    }
    CK(cudaMemcpy2DAsync(P, sizeof(double) * nn, h->b.P + (size_t)inst * h->b.p_stride, sizeof(double) * ld,
                         sizeof(double) * nn, nn, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n) *n = nn;
    return 0;
}

int slam_get_landmark_ids(slam_handle_t h, int inst, int* ids, int* M) {
    int m = 0;
    if (slam_get_num_landmarks(h, inst, &m)) return 1;
    if (m > 0) {
        CK(cudaMemcpyAsync(ids, h->b.ids + (size_t)inst * h->b.max_lm, sizeof(int) * m, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (M) *M = m;
    return 0;
}

int slam_get_assoc(slam_handle_t h, int inst, int* slot, int* k) {
    int kk = 0;
    if (get_meta(h, META_NASSOC, inst, &kk)) return 1;
    if (kk > 0) {
        CK(cudaMemcpyAsync(slot, h->b.assoc + (size_t)inst * h->b.max_meas, sizeof(int) * kk, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (k) *k = kk;
    return 0;
}

int slam_get_sigma_points(slam_handle_t h, int inst, double* X, int* n) {
    if (check_inst(h, inst)) return 1;
    if (h->kind != SLAM_UKF_SLAM && h->kind != SLAM_UKF_LOC) return fail(h, "slam_get_sigma_points: UKF only");
    if (!X) return fail(h, "slam_get_sigma_points: X is NULL");
    CK(cudaSetDevice(h->device));
    int2 sf = make_int2(0, 0);
    CK(cudaMemcpyAsync(&sf, h->uk.sigfmt + inst, sizeof(int2), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (sf.x == 0) {
        // no update() yet: X is the constructor's 4 x 9 zero matrix (ukf.cpp:20)
        for (int i = 0; i < 4 * 9; ++i) X[i] = 0.0;
        if (n) *n = 4;
        return 0;
    }
    const int nn = sf.y;
    if (!h->b.sigma) CK(cudaMalloc(&h->b.sigma, sizeof(double) * (size_t)h->b.sigma_stride));   // one instance's X
    CK(launch_ukf_sigma_points(h->b, h->uk, inst, sf.x, nn, h->b.sigma, h->stream));
    h->launches += 1;
    CK(cudaMemcpyAsync(X, h->b.sigma, sizeof(double) * (size_t)nn * (2 * nn + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n) *n = nn;
    return 0;
}

int slam_get_poses(slam_handle_t h, double* xyyaw) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    CK(launch_poses(h->b, h->d_out, h->stream));
    h->launches += 1;
    CK(cudaMemcpyAsync(xyyaw, h->d_out, sizeof(double) * 3 * h->b.batch, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int slam_get_all_status(slam_handle_t h, int* status) { return get_meta_all(h, META_STATUS, status); }
int slam_get_all_num_landmarks(slam_handle_t h, int* M) { return get_meta_all(h, META_M, M); }

int slam_set_state(slam_handle_t h, int inst, const double* x, const double* P, const int* ids, int M, int timestep) {
    if (check_inst(h, inst)) return 1;
    if (M < 0 || M > h->b.max_lm) return fail(h, "slam_set_state: M out of range");
    CK(cudaSetDevice(h->device));
    const int nn = h->b.base + 2 * M, ld = ldp_of(h->b.fixed_ld, nn);
    const int4 meta = make_int4(M, 0, timestep, 0);
    CK(cudaMemcpyAsync(h->b.x + (size_t)inst * h->b.x_stride, x, sizeof(double) * nn, cudaMemcpyHostToDevice, h->stream));
    std::vector<double> pk;
    if (h->b.ps2g) {
        // pack the lower triangle (the kernels keep P symmetric; the upper triangle of the input is ignored)
        const int live = bpl_plane_doubles(2 + M), ps2 = h->b.ps2g;
        pk.assign(2 * (size_t)live, 0.0);
        for (int r = 0; r < nn; ++r)
            for (int c = 0; c <= r; ++c) pk[bpl_idx(r + 1, c + 1, live)] = P[(size_t)r * nn + c];
        double* g = h->b.P + (size_t)inst * h->b.p_stride;
        CK(cudaMemcpyAsync(g, pk.data(), sizeof(double) * live, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(g + ps2, pk.data() + live, sizeof(double) * live, cudaMemcpyHostToDevice, h->stream));
    } else
    CK(cudaMemcpy2DAsync(h->b.P + (size_t)inst * h->b.p_stride, sizeof(double) * ld, P, sizeof(double) * nn,
                         sizeof(double) * nn, nn, cudaMemcpyHostToDevice, h->stream));
    if (M > 0) CK(cudaMemcpyAsync(h->b.ids + (size_t)inst * h->b.max_lm, ids, sizeof(int) * M, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->b.meta + inst, &meta, sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    {   // the capacity hint must not be below the state just loaded
        int cur = 0;
        CK(cudaMemcpyAsync(&cur, h->b.max_M, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (M > cur) CK(cudaMemcpyAsync(h->b.max_M, &M, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        h->step_seq = 0; h->hint_base = M > cur ? M : cur;   // next launches size conservatively again
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------ simulator
int slam_sim_create(slam_handle_t h, const double* lm_xy, int n_lm, uint64_t seed, uint32_t instance_offset, slam_sim_t* out) {
    if (!h) return 1;
    if (!out || !lm_xy || n_lm < 1) return fail(h, "slam_sim_create: bad argument");
    CK(cudaSetDevice(h->device));
    slam_sim* s = new slam_sim();
    s->owner = h;
    SimState& st = s->s;
    st.batch = h->b.batch; st.max_meas = h->b.max_meas; st.n_lm = n_lm;
    st.instance_offset = instance_offset;
    st.k0 = (uint32_t)seed; st.k1 = (uint32_t)(seed >> 32);
    CK(cudaMalloc(&s->d_lm, sizeof(double) * 2 * (size_t)n_lm));
    CK(cudaMemcpy(s->d_lm, lm_xy, sizeof(double) * 2 * (size_t)n_lm, cudaMemcpyHostToDevice));
    st.lm_xy = s->d_lm; st.lm_stride = 0;
    CK(cudaMalloc(&st.truth, sizeof(double) * 3 * (size_t)st.batch));
    CK(cudaMalloc(&st.meas, sizeof(float) * 3 * (size_t)st.batch * st.max_meas));
    CK(cudaMalloc(&st.n_meas, sizeof(int) * st.batch));
    CK(cudaMalloc(&st.overflow, sizeof(int) * st.batch));
    CK(cudaMemset(st.meas, 0, sizeof(float) * 3 * (size_t)st.batch * st.max_meas));
    *out = s;
    return slam_sim_reset(s, 0.0, 0.0, 0.0);
}
int slam_sim_make_trajectories(slam_sim_t s, double landmark_noise, double visitation_threshold, double bound,
                               double x_0, double y_0, double yaw_0, int T, float* d_fwd, float* d_ang) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    if (!d_fwd || !d_ang || T < 1) return fail(h, "slam_sim_make_trajectories: bad argument");
    if (s->s.n_lm > 256) return fail(h, "slam_sim_make_trajectories: at most 256 landmarks (per-thread tour state)");
    CK(cudaSetDevice(h->device));
    TspParams tp{landmark_noise, visitation_threshold, bound, x_0, y_0, yaw_0, T};
    CK(launch_tsp_trajectories(s->s, h->sc, tp, d_fwd, d_ang, h->stream));
    h->launches += 1;
    return 0;
}
int slam_sim_make_maps(slam_sim_t s, int map_type, int n_landmarks, double bound, double grid_step, double min_sep, int* n_out) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    if (map_type != 0 && map_type != 1) return fail(h, "Invalid map_type provided.");            // sim_node.py:196-198
    if (!(bound > 0)) return fail(h, "slam_sim_make_maps: bound must be positive");
    int n = n_landmarks;
    if (map_type == 0) {
        if (!(grid_step > 0)) return fail(h, "slam_sim_make_maps: grid_step must be positive");
        const double start = -bound + grid_step / 2;
        const int cnt = (int)std::ceil((bound - start) / grid_step);                              // len(np.arange(start, bound, step))
        n = cnt * cnt;
    }
    if (n < 1) return fail(h, "slam_sim_make_maps: no landmarks");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    double* maps = nullptr; int* d_fail = nullptr;
    CK(cudaMalloc(&maps, sizeof(double) * 2 * (size_t)n * s->s.batch));
    if (cudaMalloc(&d_fail, sizeof(int) * s->s.batch) != cudaSuccess) { cudaFree(maps); return fail(h, "slam_sim_make_maps: out of memory"); }
    std::vector<int> hf(s->s.batch, 0);
    bool ok = launch_make_maps(s->s, map_type, n, bound, grid_step, min_sep, maps, d_fail, h->stream) == cudaSuccess &&
              cudaMemcpyAsync(hf.data(), d_fail, sizeof(int) * s->s.batch, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess &&
              cudaStreamSynchronize(h->stream) == cudaSuccess;
    cudaFree(d_fail);
    h->launches += 1;
    if (ok) for (int f : hf) if (f) ok = false;
    if (!ok) { cudaFree(maps); return fail(h, "slam_sim_make_maps: a map could not be completed (min_landmark_separation too large for the bounds?)"); }
    cudaFree(s->d_lm);
    s->d_lm = maps;
    s->s.lm_xy = maps; s->s.n_lm = n; s->s.lm_stride = 2LL * n;
    if (n_out) *n_out = n;
    return 0;
}
int slam_sim_get_map(slam_sim_t s, int inst, double* lm_xy, int* n_lm) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    if (inst < 0 || inst >= s->s.batch || !lm_xy) return fail(h, "slam_sim_get_map: bad argument");
    CK(cudaMemcpyAsync(lm_xy, s->s.lm_xy + (size_t)inst * s->s.lm_stride, sizeof(double) * 2 * (size_t)s->s.n_lm, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n_lm) *n_lm = s->s.n_lm;
    return 0;
}
int slam_sim_destroy(slam_sim_t s) {
    if (!s) return 0;
    cudaSetDevice(s->owner->device);
    cudaStreamSynchronize(s->owner->stream);
    cudaFree(s->d_lm); cudaFree(s->s.truth); cudaFree(s->s.meas); cudaFree(s->s.n_meas); cudaFree(s->s.overflow);
    delete s;
    return 0;
}
int slam_sim_reset(slam_sim_t s, double x_0, double y_0, double yaw_0) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    CK(cudaSetDevice(h->device));
    std::vector<double> t(3 * (size_t)s->s.batch);
    for (int i = 0; i < s->s.batch; ++i) { t[3 * i] = x_0; t[3 * i + 1] = y_0; t[3 * i + 2] = yaw_0; }
    CK(cudaMemcpyAsync(s->s.truth, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(s->s.n_meas, 0, sizeof(int) * s->s.batch, h->stream));
    CK(cudaMemsetAsync(s->s.overflow, 0, sizeof(int) * s->s.batch, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int slam_sim_step_device(slam_sim_t s, const float* d_fwd, const float* d_ang, int cmd_stride, uint32_t step) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    CK(cudaSetDevice(h->device));
    CK(launch_sim_step(s->s, h->sc, d_fwd, d_ang, cmd_stride, step, h->stream));
    h->launches += 1;
    return 0;
}
int slam_sim_step(slam_sim_t s, const float* fwd, const float* ang, int cmd_stride, uint32_t step) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    CK(cudaSetDevice(h->device));
    if (stage_cmd(h, fwd, ang, cmd_stride)) return 1;
    return slam_sim_step_device(s, h->d_fwd, h->d_ang, cmd_stride, step);
}
const float* slam_sim_meas(slam_sim_t s) { return s ? s->s.meas : nullptr; }
const int* slam_sim_n_meas(slam_sim_t s) { return s ? s->s.n_meas : nullptr; }
int slam_sim_get_truth(slam_sim_t s, double* xyyaw) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    CK(cudaMemcpyAsync(xyyaw, s->s.truth, sizeof(double) * 3 * s->s.batch, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int slam_sim_get_meas(slam_sim_t s, float* meas, int* n_meas) {
    if (!s) return 1;
    slam_filter* h = s->owner;
    CK(cudaMemcpyAsync(meas, s->s.meas, sizeof(float) * 3 * (size_t)s->s.batch * s->s.max_meas, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(n_meas, s->s.n_meas, sizeof(int) * s->s.batch, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------ sweep + stats
}  // extern "C"

// ---- chunked persistent sweep (known-ID EKF batches): the T steps of a run are cut into chunks of sweep_chunk
// steps; each chunk is ONE launch of ekf_sweep_kernel with the shared-memory tile sized for the landmarks the batch
// holds now (device max(M), read back with a lag of two chunks, plus headroom), followed by a full-capacity launch
// that picks up the instances the first one had to leave untouched (normally none: it exits at once).
static bool sweep_capable(const slam_filter* h) {
    return h->kind == SLAM_EKF_SLAM && !h->large && h->fc.id_known && !h->sweep_off && !h->profiling;
}
struct SweepRun {            // host bookkeeping of one run's capacity hints
    int base_hint = 0;       // max(M) at run start (synchronous read)
    int chunk = 0;           // chunks launched so far
};
static int sweep_begin(slam_filter* h, SweepRun& run) {
    CK(cudaMemsetAsync(h->d_progress, 0, sizeof(int) * h->b.batch, h->stream));
    CK(cudaMemcpyAsync(&h->h_run_hint[0], h->b.max_M, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    run.base_hint = h->h_run_hint[0];
    run.chunk = 0;
    return 0;
}
static int sweep_chunk_launch(slam_filter* h, SweepRun& run, const SimState& sim, SweepArgs a, bool replay) {
    constexpr int LAG = 2;
    int hint = run.base_hint;
    if (run.chunk >= LAG) {
        const int slot = (run.chunk - LAG) % slam_filter::HINT_RING;
        CK(cudaEventSynchronize(h->run_ev[slot]));      // completed long ago in steady state; bounds the host's lead
        hint = h->h_run_hint[slot];
    }
    int cap = hint + h->sweep_headroom;
    if (h->cap_force > 0) cap = h->cap_force;
    if (cap > h->b.max_lm) cap = h->b.max_lm;
    if (h->profiling_sweep) {
        if (h->ev_used + 2 > h->ev.size()) {
            const size_t old = h->ev.size();
            h->ev.resize(old + 256);
            for (size_t i = old; i < h->ev.size(); ++i) CK(cudaEventCreate(&h->ev[i]));
        }
        CK(cudaEventRecord(h->ev[h->ev_used], h->stream));
    }
    a.progress = h->d_progress;
    a.work_counter = h->d_work;
    CK(launch_ekf_sweep(h->b, h->fc, sim, h->sc, a, replay, cap, h->force_threads, h->stream));
    h->launches += 1;
    if (cap < h->b.max_lm) {
        a.work_counter = h->d_work + 1;
        CK(launch_ekf_sweep(h->b, h->fc, sim, h->sc, a, replay, h->b.max_lm, h->force_threads, h->stream));
        h->launches += 1;
    }
    if (h->profiling_sweep) { CK(cudaEventRecord(h->ev[h->ev_used + 1], h->stream)); h->ev_used += 2; }
    const int slot = run.chunk % slam_filter::HINT_RING;
    CK(cudaMemcpyAsync(&h->h_run_hint[slot], h->b.max_M, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->run_ev[slot], h->stream));
    run.chunk += 1;
    return 0;
}

// T fused steps.  Known-ID EKF batches run on the chunked persistent sweep kernel with the filters resident in
// shared memory; everything else loops over per-step launches.
static int run_steps(slam_filter* h, slam_sim* s, const float* d_fwd, const float* d_ang, int cmd_stride, int T, uint32_t first_step) {
    const size_t per = cmd_stride ? (size_t)h->b.batch : 1;
    if (sweep_capable(h) && T > 0) {
        SweepRun run;
        if (sweep_begin(h, run)) return 1;
        for (int t0 = 0; t0 < T; t0 += h->sweep_chunk) {
            SweepArgs a{};
            a.cmd_fwd = d_fwd + per * (size_t)t0; a.cmd_ang = d_ang + per * (size_t)t0; a.cmd_stride = cmd_stride;
            a.T = (T - t0 < h->sweep_chunk) ? T - t0 : h->sweep_chunk;
            a.first_step = first_step + (uint32_t)t0; a.t0 = t0;
            if (sweep_chunk_launch(h, run, s->s, a, false)) return 1;
        }
        // the per-step capacity hint is stale now: size the next per-step launches conservatively
        h->step_seq = 0; h->hint_base = h->b.max_lm;
        return 0;
    }
    for (int t = 0; t < T; ++t) {
        const float* f = d_fwd + per * (size_t)t;
        const float* a = d_ang + per * (size_t)t;
        CK(launch_sim_step(s->s, h->sc, f, a, cmd_stride, first_step + (uint32_t)t, h->stream));
        if (do_step(h, f, a, cmd_stride, s->s.meas, s->s.n_meas, STEP_PREDICT | STEP_UPDATE)) return 1;
        CK(launch_accumulate_error(h->b, s->s, h->stream));
        h->launches += 2;
    }
    return 0;
}

static int ensure_traj(slam_filter* h, size_t need) {
    if (need > h->traj_cap) {
        cudaFree(h->d_traj_fwd); cudaFree(h->d_traj_ang);
        h->d_traj_fwd = h->d_traj_ang = nullptr; h->traj_cap = 0;
        CK(cudaMalloc(&h->d_traj_fwd, sizeof(float) * need));
        CK(cudaMalloc(&h->d_traj_ang, sizeof(float) * need));
        h->traj_cap = need;
    }
    return 0;
}

extern "C" {
int slam_run(slam_handle_t h, slam_sim_t s, const float* cmd_fwd, const float* cmd_ang, int cmd_stride, int T, uint32_t first_step) {
    if (!h || !s || s->owner != h) return fail(h, "slam_run: simulator is not bound to this handle");
    if (!cmd_fwd || !cmd_ang || T < 0) return fail(h, "slam_run: bad argument");
    CK(cudaSetDevice(h->device));
    const size_t per = cmd_stride ? (size_t)h->b.batch : 1;
    const size_t need = per * (size_t)T;
    if (ensure_traj(h, need)) return 1;
    CK(cudaMemcpyAsync(h->d_traj_fwd, cmd_fwd, sizeof(float) * need, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_traj_ang, cmd_ang, sizeof(float) * need, cudaMemcpyHostToDevice, h->stream));
    return run_steps(h, s, h->d_traj_fwd, h->d_traj_ang, cmd_stride, T, first_step);
}

int slam_run_device(slam_handle_t h, slam_sim_t s, const float* d_cmd_fwd, const float* d_cmd_ang, int cmd_stride, int T, uint32_t first_step) {
    if (!h || !s || s->owner != h) return fail(h, "slam_run_device: simulator is not bound to this handle");
    if (!d_cmd_fwd || !d_cmd_ang || T < 0) return fail(h, "slam_run_device: bad argument");
    CK(cudaSetDevice(h->device));
    return run_steps(h, s, d_cmd_fwd, d_cmd_ang, cmd_stride, T, first_step);
}

int slam_reset(slam_handle_t h, float x_0, float y_0, float yaw_0) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    double a2 = yaw_0, a3 = 0.0;
    if (h->b.base == 4) { a2 = (double)(float)std::cos((double)yaw_0); a3 = (double)(float)std::sin((double)yaw_0); }   // ukf.cpp:33 (D-1)
    CK(launch_reset(h->b, (double)x_0, (double)y_0, a2, a3, h->stream));
    CK(cudaMemsetAsync(h->b.max_M, 0, sizeof(int), h->stream));
    if (h->uk.sigfmt) CK(cudaMemsetAsync(h->uk.sigfmt, 0, sizeof(int2) * h->b.batch, h->stream));
    h->step_seq = 0; h->hint_base = 0;
    if (h->h_hint) h->h_hint[0] = 0;
    h->launches += 1;
    return 0;
}

// Is [p, p + bytes) pinned host memory the device can address at the same pointer value (cudaHostAlloc / cudaHostRegister /
// torch pin_memory under unified addressing)?  Verified ranges are remembered per handle, so a caller that walks through one
// big pinned buffer tick after tick pays the driver query once.
static bool host_mapped(slam_filter* h, const void* p, size_t bytes) {
    const uintptr_t a = (uintptr_t)p;
    for (int i = 0; i < h->n_mapped; ++i)
        if (a >= h->mapped_lo[i] && a + bytes <= h->mapped_hi[i]) return true;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || at.devicePointer != p) return false;
    // extent of the allocation (driver API through the runtime's entry-point lookup: no link-time dependency on libcuda)
    typedef int (*attr_fn)(void*, int, unsigned long long);
    static attr_fn fn = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (attr_fn)f;
        else cudaGetLastError();
    }
    unsigned long long start = 0;
    size_t size = 0;
    uintptr_t lo = a, hi = a + bytes;
    if (fn && fn(&start, 11 /* CU_POINTER_ATTRIBUTE_RANGE_START_ADDR */, (unsigned long long)a) == 0 &&
        fn(&size, 12 /* CU_POINTER_ATTRIBUTE_RANGE_SIZE */, (unsigned long long)a) == 0 && start && size) {
        lo = (uintptr_t)start; hi = lo + size;
        if (a + bytes > hi) return false;
    }
    const int slot = h->n_mapped < slam_filter::MAX_MAPPED ? h->n_mapped++ : (h->mapped_next++ % slam_filter::MAX_MAPPED);
    h->mapped_lo[slot] = lo; h->mapped_hi[slot] = hi;
    return true;
}

// Filter::update + the pose read-back of publishState in one asynchronous call.  When the caller's buffers are pinned (mapped)
// host memory the kernels read them and write the poses IN PLACE over PCIe: no staging copies, and for the batched EKF no
// separate pose kernel either (one launch per tick).  Pageable buffers take the staged path (copies on the handle's stream).
int slam_step_io(slam_handle_t h, const float* fwd, const float* ang, int cmd_stride, const float* meas, const int* n_meas, double* poses_out) {
    if (!h) return 1;
    if (!fwd || !ang || !meas || !n_meas) return fail(h, "slam_step_io: NULL argument");
    CK(cudaSetDevice(h->device));
    const BatchState& b = h->b;
    const size_t nc = cmd_stride ? (size_t)b.batch : 1;
    const bool zero_copy = !h->large && !h->no_zero_copy && host_mapped(h, fwd, sizeof(float) * nc) && host_mapped(h, ang, sizeof(float) * nc) &&
                           host_mapped(h, meas, sizeof(float) * 3 * (size_t)b.batch * b.max_meas) && host_mapped(h, n_meas, sizeof(int) * b.batch);
    const bool poses_mapped = poses_out && !h->no_zero_copy && host_mapped(h, poses_out, sizeof(double) * 3 * b.batch);
    const bool fuse = poses_mapped && h->kind == SLAM_EKF_SLAM && !h->large;
    if (zero_copy) {
        // the float4 path of the gather kernel needs a 16-byte aligned message block; a small batch reads in place
        const size_t mf = 3 * (size_t)b.batch * b.max_meas;
        if (b.batch >= 64 && ((uintptr_t)meas & 15) == 0) {
            CK(launch_gather_inputs(fwd, ang, (int)nc, n_meas, meas, b.batch, (int)mf, h->d_fwd, h->d_ang, h->d_nmeas, h->d_meas, h->stream));
            h->launches += 1;
            if (do_step(h, h->d_fwd, h->d_ang, cmd_stride, h->d_meas, h->d_nmeas, STEP_PREDICT | STEP_UPDATE, fuse ? poses_out : nullptr)) return 1;
        } else if (do_step(h, fwd, ang, cmd_stride, meas, n_meas, STEP_PREDICT | STEP_UPDATE, fuse ? poses_out : nullptr)) return 1;
    } else {
        if (stage_cmd(h, fwd, ang, cmd_stride) || stage_meas(h, meas, n_meas)) return 1;
        if (do_step(h, h->d_fwd, h->d_ang, cmd_stride, h->d_meas, h->d_nmeas, STEP_PREDICT | STEP_UPDATE, fuse ? poses_out : nullptr)) return 1;
    }
    if (poses_out && !fuse) {
        double* dst = poses_mapped ? poses_out : h->d_out;
        CK(launch_poses(h->b, dst, h->stream));
        h->launches += 1;
        if (!poses_mapped) CK(cudaMemcpyAsync(poses_out, h->d_out, sizeof(double) * 3 * h->b.batch, cudaMemcpyDeviceToHost, h->stream));
    }
    return 0;
}

// T x (Filter::update + the pose read-back of publishState) for a whole recorded run, HOST buffers in and out.
// The trajectory is cut into chunks; chunk c+1 is uploaded (copy stream) while chunk c runs (known-ID EKF batches:
// ekf_sweep_kernel in replay mode with the filters resident in shared memory; otherwise per-step launches) and the
// poses of chunk c-1 are downloaded (second copy stream).  Asynchronous: slam_synchronize() waits for everything,
// including the last download.
int slam_run_io(slam_handle_t h, const float* cmd_fwd, const float* cmd_ang, int cmd_stride, const float* meas,
                const int* n_meas, double* poses_out, int T) {
    if (!h) return 1;
    if (!cmd_fwd || !cmd_ang || !meas || !n_meas || T < 0) return fail(h, "slam_run_io: bad argument");
    CK(cudaSetDevice(h->device));
    const BatchState& b = h->b;
    const size_t per = cmd_stride ? (size_t)b.batch : 1;
    const size_t m_step = (size_t)b.batch * b.max_meas * 3, n_step = (size_t)b.batch, p_step = (size_t)b.batch * 3;
    const int C = h->sweep_chunk;
    if (h->r_chunk_cap < C) {
        for (int i = 0; i < 2; ++i) {
            cudaFree(h->d_rmeas[i]); cudaFree(h->d_rn[i]); cudaFree(h->d_rposes[i]);
            h->d_rmeas[i] = nullptr; h->d_rn[i] = nullptr; h->d_rposes[i] = nullptr;
        }
        h->r_chunk_cap = 0;
        for (int i = 0; i < 2; ++i) {
            CK(cudaMalloc(&h->d_rmeas[i], sizeof(float) * m_step * C));
            CK(cudaMalloc(&h->d_rn[i], sizeof(int) * n_step * C));
            CK(cudaMalloc(&h->d_rposes[i], sizeof(double) * p_step * C));
        }
        h->r_chunk_cap = C;
    }
    if (ensure_traj(h, per * (size_t)(T > 0 ? T : 1))) return 1;
    CK(cudaMemcpyAsync(h->d_traj_fwd, cmd_fwd, sizeof(float) * per * T, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_traj_ang, cmd_ang, sizeof(float) * per * T, cudaMemcpyHostToDevice, h->stream));
    const bool sweep = sweep_capable(h);
    SweepRun run;
    if (sweep && T > 0 && sweep_begin(h, run)) return 1;
    SimState nosim{};
    int c = 0;
    for (int t0 = 0; t0 < T; t0 += C, ++c) {
        const int slot = c & 1, Tc = (T - t0 < C) ? T - t0 : C;
        // upload chunk c once the kernels that read this slot two chunks ago are done
        CK(cudaStreamWaitEvent(h->s_h2d, h->ev_comp[slot], 0));
        CK(cudaMemcpyAsync(h->d_rmeas[slot], meas + m_step * t0, sizeof(float) * m_step * Tc, cudaMemcpyHostToDevice, h->s_h2d));
        CK(cudaMemcpyAsync(h->d_rn[slot], n_meas + n_step * t0, sizeof(int) * n_step * Tc, cudaMemcpyHostToDevice, h->s_h2d));
        CK(cudaEventRecord(h->ev_h2d[slot], h->s_h2d));
        CK(cudaStreamWaitEvent(h->stream, h->ev_h2d[slot], 0));
        CK(cudaStreamWaitEvent(h->stream, h->ev_d2h[slot], 0));        // the pose buffer of this slot has been drained
        if (sweep) {
            SweepArgs a{};
            a.cmd_fwd = h->d_traj_fwd + per * (size_t)t0; a.cmd_ang = h->d_traj_ang + per * (size_t)t0; a.cmd_stride = cmd_stride;
            a.T = Tc; a.first_step = 0; a.t0 = t0;
            a.r_meas = h->d_rmeas[slot]; a.r_nmeas = h->d_rn[slot]; a.r_poses = poses_out ? h->d_rposes[slot] : nullptr;
            if (sweep_chunk_launch(h, run, nosim, a, true)) return 1;
        } else {
            for (int t = 0; t < Tc; ++t) {
                if (do_step(h, h->d_traj_fwd + per * (size_t)(t0 + t), h->d_traj_ang + per * (size_t)(t0 + t), cmd_stride,
                            h->d_rmeas[slot] + m_step * t, h->d_rn[slot] + n_step * t, STEP_PREDICT | STEP_UPDATE)) return 1;
                if (poses_out) { CK(launch_poses(h->b, h->d_rposes[slot] + p_step * t, h->stream)); h->launches += 1; }
            }
        }
        CK(cudaEventRecord(h->ev_comp[slot], h->stream));
        if (poses_out) {
            CK(cudaStreamWaitEvent(h->s_d2h, h->ev_comp[slot], 0));
            CK(cudaMemcpyAsync(poses_out + p_step * t0, h->d_rposes[slot], sizeof(double) * p_step * Tc, cudaMemcpyDeviceToHost, h->s_d2h));
            CK(cudaEventRecord(h->ev_d2h[slot], h->s_d2h));
        }
    }
    if (poses_out) for (int i = 0; i < 2; ++i) CK(cudaStreamWaitEvent(h->stream, h->ev_d2h[i], 0));
    if (sweep) { h->step_seq = 0; h->hint_base = h->b.max_lm; }
    return 0;
}

int slam_set_profiling(slam_handle_t h, int on) {
    if (!h) return 1;
    h->profiling = on == 1;           // 1: per-step kernel (forces the per-step path), 2: sweep kernel, 3: lm_gemm (large map)
    h->profiling_sweep = on == 2;
    h->profiling_gemm = on == 3 && h->large;
    h->ev_used = 0;
    return 0;
}
int slam_get_profile(slam_handle_t h, double* total_ms, long long* launches) {
    if (!h) return 1;
    CK(cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
        tot += ms;
        if (getenv("SLAM_DEBUG_SWEEP")) fprintf(stderr, "profiled launch %zu: %.4f ms\n", i / 2, ms);
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = (long long)(h->ev_used / 2);
    h->ev_used = 0;
    return 0;
}

int slam_accumulate_error(slam_handle_t h, slam_sim_t s) {
    if (!h || !s || s->owner != h) return fail(h, "slam_accumulate_error: simulator is not bound to this handle");
    CK(cudaSetDevice(h->device));
    CK(launch_accumulate_error(h->b, s->s, h->stream));
    h->launches += 1;
    return 0;
}
int slam_get_stats(slam_handle_t h, double* out) {
    if (!h) return 1;
    CK(cudaSetDevice(h->device));
    double* d = h->d_out + 3 * (size_t)h->b.batch;
    CK(launch_reduce_stats(h->b, d, h->stream));
    h->launches += 1;
    CK(cudaMemcpyAsync(out, d, sizeof(double) * SLAM_NUM_STATS, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
int slam_get_error_histogram(slam_handle_t h, double lo, double hi, int nbins, long long* counts, double* avg_err) {
    if (!h) return 1;
    if (!(hi > lo) || nbins < 1 || nbins > 8190 || !counts) return fail(h, "slam_get_error_histogram: need lo < hi, 1 <= nbins <= 8190, counts != NULL");
    CK(cudaSetDevice(h->device));
    if (h->hist_cap < nbins + 2) {
        cudaFree(h->d_hist); h->d_hist = nullptr; h->hist_cap = 0;
        CK(cudaMalloc(&h->d_hist, sizeof(unsigned long long) * (nbins + 2)));
        h->hist_cap = nbins + 2;
    }
    if (!h->d_avg) CK(cudaMalloc(&h->d_avg, sizeof(double) * h->b.batch));
    unsigned long long* d_cnt = h->d_hist;
    double* d_avg = h->d_avg;
    int rc = 0;
    do {
        if (cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * (nbins + 2), h->stream) != cudaSuccess) { rc = 1; break; }
        if (launch_error_histogram(h->b, d_avg, d_cnt, lo, hi, nbins, h->stream) != cudaSuccess) { rc = 1; break; }
        h->launches += 1;
        if (cudaMemcpyAsync(counts, d_cnt, sizeof(long long) * (nbins + 2), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { rc = 1; break; }
        if (avg_err && cudaMemcpyAsync(avg_err, d_avg, sizeof(double) * h->b.batch, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { rc = 1; break; }
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = 1; break; }
    } while (false);
    if (rc) return fail(h, "slam_get_error_histogram: CUDA error");
    return 0;
}
int slam_reset_stats(slam_handle_t h) {
    if (!h) return 1;
    CK(cudaMemsetAsync(h->b.stats, 0, sizeof(double) * (size_t)h->b.batch * SLAM_NUM_STATS, h->stream));
    return 0;
}

}  // extern "C"
