// sim.cu -- on-GPU workload source: the simulator's measurement generator (get_cmd,
// ekf_ws/src/base_pkg/src/sim_node.py:209-250), one warp per simulated vehicle, plus the small
// per-instance error accumulators that feed the RMSE / NEES reduce.
#include "common.cuh"

namespace slam {

constexpr int SIM_THREADS = 128;

// sim_node.py:209-250.  Every lane of the warp carries the (identical) truth state; lanes stride over
// landmark ids and an ordered ballot compaction keeps the message in ascending-id order (:231-249).
__global__ void __launch_bounds__(SIM_THREADS)
sim_step_kernel(SimState s, SimConst sc, const float* __restrict__ fwd, const float* __restrict__ ang,
                int cmd_stride, uint32_t step) {
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= s.batch) return;
    const uint32_t inst = s.instance_offset + (uint32_t)w;
    uint32_t rn[4];
    philox4x32_10(inst, step, 0u, 0u, s.k0, s.k1, rn);
    // add noise to the command, clamp (:216-220); msg.fwd / msg.ang are the float32 wire values
    double d = (double)fwd[cmd_stride ? w : 0] + 2 * sc.V_00 * uniform53(rn[0], rn[1]) - sc.V_00;
    double hdg = (double)ang[cmd_stride ? w : 0] + 2 * sc.V_11 * uniform53(rn[2], rn[3]) - sc.V_11;
    d = fmax(0.0, fmin(d, sc.d_max));
    hdg = fmax(-sc.th_max, fmin(hdg, sc.th_max));
    double* tr = s.truth + 3 * (size_t)w;
    double sy, cy;
    sincos(tr[2], &sy, &cy);
    const double tx = tr[0] + d * cy, ty = tr[1] + d * sy, tyaw = tr[2] + hdg;   // :222 (yaw never wrapped)
    __syncwarp();
    if (lane == 0) { tr[0] = tx; tr[1] = ty; tr[2] = tyaw; }
    float* out = s.meas + (size_t)w * s.max_meas * 3;
    int count = 0;
    for (int base = 0; base < s.n_lm; base += 32) {
        const int id = base + lane;
        bool vis = false;
        double r = 0.0, beta = 0.0;
        if (id < s.n_lm) {
            const double dx = s.lm_xy[2 * id] - tx, dy = s.lm_xy[2 * id + 1] - ty;
            r = sqrt(dx * dx + dy * dy);                                    // :235
            beta = remainder(atan2(dy, dx) - tyaw, TWO_PI_REF);             // :236-237
            vis = !(r > sc.range_max) && (beta > sc.fov_min && beta < sc.fov_max);   // :239-241
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, vis);
        const int pos = count + __popc(ballot & ((1u << lane) - 1u));
        if (vis && pos < s.max_meas) {
            philox4x32_10(inst, step, 1u + (uint32_t)id, 0u, s.k0, s.k1, rn);
            out[3 * pos] = (float)id;                                       // float32 wire, :245-249
            out[3 * pos + 1] = (float)(r + 2 * sc.W_00 * uniform53(rn[0], rn[1]) - sc.W_00);
            out[3 * pos + 2] = (float)(beta + 2 * sc.W_11 * uniform53(rn[2], rn[3]) - sc.W_11);
        }
        count += __popc(ballot);
    }
    if (lane == 0) {
        s.n_meas[w] = count < s.max_meas ? count : s.max_meas;
        if (count > s.max_meas) s.overflow[w] = 1;
    }
}

cudaError_t launch_sim_step(const SimState& s, const SimConst& sc, const float* d_fwd, const float* d_ang,
                            int cmd_stride, uint32_t step, cudaStream_t st) {
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
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) C[a][c] = P[a * ld + c];
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
    const double ex = x[0] - tr[0], ey = x[1] - tr[1], eyaw = remainder(yaw - tr[2], TWO_PI_REF);
    // NEES = e^T C^-1 e via the adjugate of the (symmetrised) 3x3 block
    for (int a = 0; a < 3; ++a) for (int c = a + 1; c < 3; ++c) { const double m = 0.5 * (C[a][c] + C[c][a]); C[a][c] = m; C[c][a] = m; }
    const double c00 = C[1][1] * C[2][2] - C[1][2] * C[2][1];
    const double c01 = C[1][2] * C[2][0] - C[1][0] * C[2][2];
    const double c02 = C[1][0] * C[2][1] - C[1][1] * C[2][0];
    const double det = C[0][0] * c00 + C[0][1] * c01 + C[0][2] * c02;
    const double c11 = C[0][0] * C[2][2] - C[0][2] * C[2][0];
    const double c12 = C[0][1] * C[2][0] - C[0][0] * C[2][1];
    const double c22 = C[0][0] * C[1][1] - C[0][1] * C[1][0];
    const double quad = ex * (c00 * ex + c01 * ey + c02 * eyaw) + ey * (c01 * ex + c11 * ey + c12 * eyaw) +
                        eyaw * (c02 * ex + c12 * ey + c22 * eyaw);
    double* st = b.stats + (size_t)i * SLAM_NUM_STATS;
    st[0] += 1.0;
    st[1] += ex * ex;
    st[2] += ey * ey;
    st[3] += eyaw * eyaw;
    st[4] += sqrt(ex * ex + ey * ey);
    st[5] += quad / det;
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

// vehicle pose (x, y, yaw) of every instance -> [batch][3]
__global__ void poses_kernel(BatchState b, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.batch) return;
    const double* x = b.x + (size_t)i * b.x_stride;
    out[3 * i] = x[0];
    out[3 * i + 1] = x[1];
    out[3 * i + 2] = (b.base == 3) ? x[2] : remainder(atan2(x[3], x[2]), TWO_PI_REF);   // ukf.cpp:71
}

cudaError_t launch_poses(const BatchState& b, double* d_out, cudaStream_t st) {
    poses_kernel<<<(b.batch + 127) / 128, 128, 0, st>>>(b, d_out);
    return cudaGetLastError();
}

}  // namespace slam
