// sim.cu -- on-GPU workload source: the simulator's measurement generator (get_cmd,
// ekf_ws/src/base_pkg/src/sim_node.py:209-250), one warp per simulated vehicle, plus the small
// per-instance error accumulators that feed the RMSE / NEES reduce.
#include "sim_device.cuh"

namespace slam {

constexpr int SIM_THREADS = 128;

// sim_node.py:209-250: one warp per simulated vehicle (body in sim_device.cuh).
__global__ void __launch_bounds__(SIM_THREADS)
sim_step_kernel(SimState s, SimConst sc, const float* __restrict__ fwd, const float* __restrict__ ang,
                int cmd_stride, uint32_t step) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= s.batch) return;
    double* trg = s.truth + 3 * (size_t)w;
    double tr[3] = {trg[0], trg[1], trg[2]};
    __syncwarp();
    const int count = sim_get_cmd_warp(lane, sc, s.lm_xy + (size_t)w * s.lm_stride, s.n_lm, s.max_meas, s.k0, s.k1, s.instance_offset + (uint32_t)w,
                                       step, fwd[cmd_stride ? w : 0], ang[cmd_stride ? w : 0], tr,
                                       s.meas + (size_t)w * s.max_meas * 3);
    if (lane == 0) {
        trg[0] = tr[0]; trg[1] = tr[1]; trg[2] = tr[2];
        s.n_meas[w] = count < s.max_meas ? count : s.max_meas;
        if (count > s.max_meas) s.overflow[w] = 1;
    }
}

// The same generator for a few vehicles on a LARGE map (the single large-map instance, BASELINE configs[3]): one CTA of
// 32 warps per vehicle.  Warp w takes the id chunks w, w + 32, ...; the per-chunk detection counts meet in shared memory,
// an exclusive scan over the chunks gives every chunk its position in the ascending-id message (:231-249), and the
// entries are written in a second phase -- bit-identical to the one-warp kernel (same expressions, same Philox keys).
constexpr int SIMW_THREADS = 1024, SIMW_MAXCH = 8;       // up to 8 chunks per warp: n_lm <= 8192
__global__ void __launch_bounds__(SIMW_THREADS)
sim_step_wide_kernel(SimState s, SimConst sc, const float* __restrict__ fwd, const float* __restrict__ ang,
                     int cmd_stride, uint32_t step) {
    __shared__ int s_cnt[32 * SIMW_MAXCH + 1];
    const int w = blockIdx.x;                              // vehicle
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* trg = s.truth + 3 * (size_t)w;
    const uint32_t inst = s.instance_offset + (uint32_t)w;
    uint32_t rn[4];
    philox4x32_10(inst, step, 0u, 0u, s.k0, s.k1, rn);
    double d = (double)fwd[cmd_stride ? w : 0] + 2 * sc.V_00 * uniform53(rn[0], rn[1]) - sc.V_00;       // :216-220
    double hdg = (double)ang[cmd_stride ? w : 0] + 2 * sc.V_11 * uniform53(rn[2], rn[3]) - sc.V_11;
    d = fmax(0.0, fmin(d, sc.d_max));
    hdg = fmax(-sc.th_max, fmin(hdg, sc.th_max));
    double sy, cy;
    sincos(trg[2], &sy, &cy);
    const double tx = trg[0] + d * cy, ty = trg[1] + d * sy, tyaw = trg[2] + hdg;   // :222 (yaw never wrapped)
    const int nch = (s.n_lm + 31) / 32;
    double rr[SIMW_MAXCH], bt[SIMW_MAXCH];
    unsigned bal[SIMW_MAXCH];
#pragma unroll
    for (int c = 0; c < SIMW_MAXCH; ++c) {
        const int chunk = warp + 32 * c;
        const int id = chunk * 32 + lane;
        bool vis = false;
        double r = 0.0, beta = 0.0;
        if (chunk < nch && id < s.n_lm) {
            const double* lmw = s.lm_xy + (size_t)w * s.lm_stride;
            const double dx = lmw[2 * id] - tx, dy = lmw[2 * id + 1] - ty;
            r = sqrt(dx * dx + dy * dy);                                    // :235
            if (!(r > sc.range_max)) {                                      // :239
                beta = wrap_2pi(atan2(dy, dx) - tyaw);                      // :236-237
                vis = beta > sc.fov_min && beta < sc.fov_max;               // :240-241
            }
        }
        rr[c] = r; bt[c] = beta;
        bal[c] = __ballot_sync(0xffffffffu, vis);
        if (lane == 0 && chunk < nch) s_cnt[chunk] = __popc(bal[c]);
    }
    __syncthreads();                                       // (also orders every thread's read of trg before the write below)
    if (threadIdx.x == 0) {                                // exclusive scan over <= 256 chunk counts
        int run = 0;
        for (int c = 0; c < nch; ++c) { const int t = s_cnt[c]; s_cnt[c] = run; run += t; }
        s_cnt[nch] = run;
        trg[0] = tx; trg[1] = ty; trg[2] = tyaw;
        s.n_meas[w] = run < s.max_meas ? run : s.max_meas;
        if (run > s.max_meas) s.overflow[w] = 1;
    }
    __syncthreads();
    float* out = s.meas + (size_t)w * s.max_meas * 3;
#pragma unroll
    for (int c = 0; c < SIMW_MAXCH; ++c) {
        const int chunk = warp + 32 * c;
        if (chunk >= nch) continue;
        const int id = chunk * 32 + lane;
        const bool vis = (bal[c] >> lane) & 1u;
        const int pos = s_cnt[chunk] + __popc(bal[c] & ((1u << lane) - 1u));
        if (vis && pos < s.max_meas) {
            philox4x32_10(inst, step, 1u + (uint32_t)id, 0u, s.k0, s.k1, rn);
            out[3 * pos] = (float)id;                                       // float32 wire, :245-249
            out[3 * pos + 1] = (float)(rr[c] + 2 * sc.W_00 * uniform53(rn[0], rn[1]) - sc.W_00);
            out[3 * pos + 2] = (float)(bt[c] + 2 * sc.W_11 * uniform53(rn[2], rn[3]) - sc.W_11);
        }
    }
}

cudaError_t launch_sim_step(const SimState& s, const SimConst& sc, const float* d_fwd, const float* d_ang,
                            int cmd_stride, uint32_t step, cudaStream_t st) {
    if (s.batch <= 16 && s.n_lm > 256 && s.n_lm <= 32 * 32 * SIMW_MAXCH) {
        sim_step_wide_kernel<<<s.batch, SIMW_THREADS, 0, st>>>(s, sc, d_fwd, d_ang, cmd_stride, step);
        return cudaGetLastError();
    }
    const int warps_per_block = SIM_THREADS / 32;
    const int blocks = (s.batch + warps_per_block - 1) / warps_per_block;
    sim_step_kernel<<<blocks, SIM_THREADS, 0, st>>>(s, sc, d_fwd, d_ang, cmd_stride, step);
    return cudaGetLastError();
}

// Per-instance accuracy accumulators against the simulator's truth.  Position error is the reference's
// metric (plotting_node.py:212-214); RMSE terms and the 3-dof pose NEES are the extension BASELINE.json asks for.
__global__ void accumulate_error_kernel(BatchState b, SimState s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.batch) return;
    const double* x = b.x + (size_t)i * b.x_stride;
    const double* P = b.P + (size_t)i * b.p_stride;
    const int n = b.base + 2 * b.meta[i].x;
    const int ld = ldp_of(b.fixed_ld, n);
    const double* tr = s.truth + 3 * (size_t)i;
    double yaw, C[3][3];
    if (b.base == 3) {
        yaw = x[2];
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) C[a][c] = b.ps2g ? P[bpl_sym(a + 1, c + 1, b.ps2g)] : P[a * ld + c];
    } else {
        // UKF keeps (cos, sin): yaw = atan2(s, c) with Jacobian [-s, c]/(c^2+s^2)
        const double cc = x[2], ss = x[3], q = cc * cc + ss * ss;
        yaw = remainder(atan2(ss, cc), TWO_PI_REF);
        const double J[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, -ss / q, cc / q}};
        double T[3][4];
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 4; ++c) {
            double t = 0; for (int k = 0; k < 4; ++k) t += J[a][k] * P[k * ld + c];
            T[a][c] = t;
        }
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) {
            double t = 0; for (int k = 0; k < 4; ++k) t += T[a][k] * J[c][k];
            C[a][c] = t;
        }
    }
    const double ex = x[0] - tr[0], ey = x[1] - tr[1], eyaw = wrap_2pi(yaw - tr[2]);
    double acc[6] = {0, 0, 0, 0, 0, 0};
    pose_error_terms(ex, ey, eyaw, C, acc);
    double* st = b.stats + (size_t)i * SLAM_NUM_STATS;
    for (int k = 0; k < 6; ++k) st[k] += acc[k];
}

cudaError_t launch_accumulate_error(const BatchState& b, const SimState& s, cudaStream_t st) {
    accumulate_error_kernel<<<(b.batch + 127) / 128, 128, 0, st>>>(b, s);
    return cudaGetLastError();
}

__global__ void reduce_stats_kernel(BatchState b, double* out) {
    __shared__ double sh[SLAM_NUM_STATS][8];
    double acc[SLAM_NUM_STATS];
    for (int k = 0; k < SLAM_NUM_STATS; ++k) acc[k] = 0.0;
    for (int i = threadIdx.x; i < b.batch; i += blockDim.x) {
        const double* st = b.stats + (size_t)i * SLAM_NUM_STATS;
        for (int k = 0; k < 6; ++k) acc[k] += st[k];
        for (int k = 8; k < SLAM_NUM_STATS; ++k) acc[k] += st[k];
        const int4 m = b.meta[i];
        acc[6] += (m.y != 0) ? 1.0 : 0.0;
        acc[7] += (double)m.x;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < SLAM_NUM_STATS; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < SLAM_NUM_STATS) {
        double v = 0.0;
        for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) v += sh[threadIdx.x][w2];
        out[threadIdx.x] = v;
    }
}

cudaError_t launch_reduce_stats(const BatchState& b, double* d_out, cudaStream_t st) {
    reduce_stats_kernel<<<1, 256, 0, st>>>(b, d_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Precomputed command trajectories on the device, one Monte-Carlo instance per thread (sim_node.py:63-152,
// generate_trajectory): a noisy copy of the map (uniform +-landmark_noise, clamped 1 m inside the region, :83-87), the
// nearest-neighbour tour from the start pose (:89-112, first minimum wins), then one command per step that drives at most
// d_max / th_max towards the current goal and rotates the tour when within visitation_threshold (:118-152).  The
// reference draws the map noise from an unseeded random.random(); here Philox keyed (seed; instance, landmark id, 0, 1).
// Output: the float32 wire values of Command.msg:3-5, [T][batch] (cmd_stride 1 of slam_run_device).
// ---------------------------------------------------------------------------------------------------------
constexpr int TSP_MAX_LM = 256;
__global__ void tsp_trajectory_kernel(SimState s, SimConst sc, TspParams tp, float* __restrict__ fwd_out, float* __restrict__ ang_out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= s.batch) return;
    const int N = s.n_lm;
    const uint32_t inst = s.instance_offset + (uint32_t)w;
    double nx[TSP_MAX_LM], ny[TSP_MAX_LM];
    short path[TSP_MAX_LM];
    unsigned char seen[TSP_MAX_LM];
    const double lo = -tp.bound + 1, hi = tp.bound - 1;
    for (int i = 0; i < N; ++i) {
        uint32_t rn[4];
        philox4x32_10(inst, (uint32_t)i, 0u, 1u, s.k0, s.k1, rn);
        const double* lmw = s.lm_xy + (size_t)w * s.lm_stride;
        const double ax = lmw[2 * i] + 2 * tp.landmark_noise * uniform53(rn[0], rn[1]) - tp.landmark_noise;
        const double ay = lmw[2 * i + 1] + 2 * tp.landmark_noise * uniform53(rn[2], rn[3]) - tp.landmark_noise;
        nx[i] = fmax(lo, fmin(ax, hi)); ny[i] = fmax(lo, fmin(ay, hi));
        seen[i] = 0;
    }
    double x = tp.x0, y = tp.y0, th = tp.yaw0;
    auto dist = [](double ax, double ay, double bx, double by) { const double dx = ax - bx, dy = ay - by; return sqrt(dx * dx + dy * dy); };
    int cur = 0;
    double best = dist(nx[0], ny[0], x, y);
    for (int i = 0; i < N; ++i) { const double d = dist(nx[i], ny[i], x, y); if (d < best) { cur = i; best = d; } }
    path[0] = (short)cur; seen[cur] = 1;
    for (int k = 1; k < N; ++k) {
        int g = -1; double bd = -1.0;
        for (int i = 0; i < N; ++i) {
            if (seen[i]) continue;
            const double d = dist(nx[i], ny[i], nx[cur], ny[cur]);
            if (bd < 0 || d < bd) { g = i; bd = d; }
        }
        path[k] = (short)g; seen[g] = 1; cur = g;
    }
    int head = 0;
    for (int t = 0; t < tp.T; ++t) {
        if (dist(x, y, nx[path[head]], ny[path[head]]) < tp.visitation_threshold) head = (head + 1 == N) ? 0 : head + 1;
        const double gx = nx[path[head]], gy = ny[path[head]];
        double d = dist(gx, gy, x, y);
        const double gb = atan2(gy - y, gx - x);
        double hdg = remainder(gb - th, TWO_PI_REF);
        d = fmin(d, sc.d_max);
        if (fabs(hdg) > sc.th_max) hdg = (hdg > 0) ? sc.th_max : -sc.th_max;
        double sn, cs;
        sincos(th, &sn, &cs);
        x = x + d * cs; y = y + d * sn; th = th + hdg;
        fwd_out[(size_t)t * s.batch + w] = (float)d;
        ang_out[(size_t)t * s.batch + w] = (float)hdg;
    }
}

cudaError_t launch_tsp_trajectories(const SimState& s, const SimConst& sc, const TspParams& tp, float* d_fwd, float* d_ang, cudaStream_t st) {
    if (s.n_lm > TSP_MAX_LM) return cudaErrorInvalidValue;
    tsp_trajectory_kernel<<<(s.batch + 63) / 64, 64, 0, st>>>(s, sc, tp, d_fwd, d_ang);
    return cudaGetLastError();
}

// generate_landmarks, ekf_ws/src/base_pkg/src/sim_node.py:155-206, one map per simulated vehicle, one thread per vehicle.
//   map_type 0 "grid"   (:165-176): the lattice np.arange(-bound + step / 2, bound, step) squared, ids row-major -- the same for
//                       every vehicle; n_lm = count^2 is computed by the host with numpy's arange length rule.
//   map_type 1 "random" (:177-188): positions uniform in [-bound, bound)^2, rejected when closer than min_sep to an accepted
//                       landmark (the occupancy test of :182 always passes on the blank map the benchmarks use).  The reference
//                       draws from random.random(); here attempt a of vehicle i draws Philox (seed; i, a, 0, 2) (deviation D-5).
// fail[w] = 1 when a vehicle's map could not be completed in 64 n_lm attempts (min_sep too large for the area).
__global__ void make_maps_kernel(SimState s, const int map_type, const int n_lm, const double bound, const double grid_step,
                                 const double min_sep, double* __restrict__ maps, int* __restrict__ fail) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= s.batch) return;
    double* lm = maps + (size_t)w * 2 * n_lm;
    if (map_type == 0) {
        const double start = -bound + grid_step / 2;
        int cnt = 0;
        while (cnt * cnt < n_lm) ++cnt;
        for (int r = 0; r < cnt; ++r)
            for (int c = 0; c < cnt; ++c) {      // np.arange's fill: start + i * delta, delta = (start + step) - start, no FMA contraction
                const double delta = __dadd_rn(__dadd_rn(start, grid_step), -start);
                lm[2 * (r * cnt + c)] = __dadd_rn(start, __dmul_rn((double)r, delta));
                lm[2 * (r * cnt + c) + 1] = __dadd_rn(start, __dmul_rn((double)c, delta));
            }
        fail[w] = 0;
        return;
    }
    const uint32_t inst = s.instance_offset + (uint32_t)w;
    int have = 0;
    for (uint32_t a = 0; have < n_lm && a < 64u * (uint32_t)n_lm; ++a) {
        uint32_t rn[4];
        philox4x32_10(inst, a, 0u, 2u, s.k0, s.k1, rn);
        // (explicit roundings: the map must be bit-identical to the CPU restatement, so no FMA contraction here)
        const double px = __dadd_rn(__dmul_rn(2 * bound, uniform53(rn[0], rn[1])), -bound);                               // :180
        const double py = __dadd_rn(__dmul_rn(2 * bound, uniform53(rn[2], rn[3])), -bound);
        bool close = false;
        for (int q = 0; q < have && !close; ++q) {
            const double dx = lm[2 * q] - px, dy = lm[2 * q + 1] - py;
            close = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) < min_sep;                                        // :184
        }
        if (close) continue;
        lm[2 * have] = px; lm[2 * have + 1] = py;                                                                          // :186-187
        ++have;
    }
    fail[w] = (have < n_lm) ? 1 : 0;
}
cudaError_t launch_make_maps(const SimState& s, int map_type, int n_lm, double bound, double grid_step, double min_sep, double* d_maps,
                             int* d_fail, cudaStream_t st) {
    make_maps_kernel<<<(s.batch + 63) / 64, 64, 0, st>>>(s, map_type, n_lm, bound, grid_step, min_sep, d_maps, d_fail);
    return cudaGetLastError();
}

// Per-run accuracy analytics (plotting_node.py:195-218 computes ONE number per run, the average position error, which
// make_bar_graphs.py then tabulates over the ten recorded runs of a setting).  At Monte-Carlo scale the same number is
// produced for every instance on the device together with its histogram: avg[i] = sum |e_pos| / steps; bin 0 counts runs
// below `lo`, bin nbins + 1 runs at or above `hi` (and non-finite ones).  Counts are exact integers, so ranks all-reduce
// them (SUM) and percentiles are read off the merged histogram.
__global__ void error_histogram_kernel(BatchState b, double* avg_out, unsigned long long* counts, double lo, double hi, int nbins) {
    extern __shared__ unsigned int sh_cnt[];
    for (int k = threadIdx.x; k < nbins + 2; k += blockDim.x) sh_cnt[k] = 0u;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.batch) {
        const double* st = b.stats + (size_t)i * SLAM_NUM_STATS;
        const double e = st[0] > 0.0 ? st[4] / st[0] : 0.0;          // avg_err = sum(errors) / num_iters, :214
        if (avg_out) avg_out[i] = e;
        int bin;
        if (!(e >= lo)) bin = isfinite(e) ? 0 : nbins + 1;
        else if (!(e < hi)) bin = nbins + 1;
        else { bin = 1 + (int)((e - lo) / (hi - lo) * nbins); if (bin > nbins) bin = nbins; }
        atomicAdd(&sh_cnt[bin], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nbins + 2; k += blockDim.x) if (sh_cnt[k]) atomicAdd(&counts[k], (unsigned long long)sh_cnt[k]);
}

cudaError_t launch_error_histogram(const BatchState& b, double* d_avg, unsigned long long* d_counts, double lo, double hi, int nbins,
                                   cudaStream_t st) {
    error_histogram_kernel<<<(b.batch + 255) / 256, 256, sizeof(unsigned int) * (nbins + 2), st>>>(b, d_avg, d_counts, lo, hi, nbins);
    return cudaGetLastError();
}

// vehicle pose (x, y, yaw) of every instance -> [batch][3]
__global__ void poses_kernel(BatchState b, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.batch) return;
    const double* x = b.x + (size_t)i * b.x_stride;
    out[3 * i] = x[0];
    out[3 * i + 1] = x[1];
    out[3 * i + 2] = (b.base == 3) ? x[2] : remainder(atan2(x[3], x[2]), TWO_PI_REF);   // ukf.cpp:71
}

// One tick's inputs from MAPPED PINNED HOST memory into the handle's HBM staging buffers, read by a wide grid with 16-byte
// loads: the whole message block crosses PCIe once at bandwidth, instead of every step CTA stalling on a few dependent
// round trips (and instead of four separate small DMA copies on the stream).
__global__ void gather_inputs_kernel(const float* __restrict__ h_fwd, const float* __restrict__ h_ang, const int ncmd,
                                     const int* __restrict__ h_n, const float* __restrict__ h_meas, const int batch, const int meas_floats,
                                     float* d_fwd, float* d_ang, int* d_n, float* d_meas) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int n4 = meas_floats >> 2;
    const float4* src = reinterpret_cast<const float4*>(h_meas);
    float4* dst = reinterpret_cast<float4*>(d_meas);
    for (int i = tid; i < n4; i += nth) dst[i] = src[i];
    for (int i = (n4 << 2) + tid; i < meas_floats; i += nth) d_meas[i] = h_meas[i];
    for (int i = tid; i < batch; i += nth) d_n[i] = h_n[i];
    for (int i = tid; i < ncmd; i += nth) { d_fwd[i] = h_fwd[i]; d_ang[i] = h_ang[i]; }
}
cudaError_t launch_gather_inputs(const float* h_fwd, const float* h_ang, int ncmd, const int* h_n, const float* h_meas, int batch,
                                 int meas_floats, float* d_fwd, float* d_ang, int* d_n, float* d_meas, cudaStream_t st) {
    int blocks = (meas_floats / 4 + 255) / 256;
    if (blocks > 296) blocks = 296;
    if (blocks < 1) blocks = 1;
    gather_inputs_kernel<<<blocks, 256, 0, st>>>(h_fwd, h_ang, ncmd, h_n, h_meas, batch, meas_floats, d_fwd, d_ang, d_n, d_meas);
    return cudaGetLastError();
}

cudaError_t launch_poses(const BatchState& b, double* d_out, cudaStream_t st) {
    poses_kernel<<<(b.batch + 127) / 128, 128, 0, st>>>(b, d_out);
    return cudaGetLastError();
}

}  // namespace slam
