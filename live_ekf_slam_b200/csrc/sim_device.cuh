// sim_device.cuh -- warp-level body of the simulator's measurement generator and the per-instance error terms,
// shared by the stand-alone kernels (sim.cu) and the fused Monte-Carlo sweep kernel (ekf_batch.cu).
#pragma once

#include "common.cuh"

namespace slam {

// get_cmd, ekf_ws/src/base_pkg/src/sim_node.py:209-250, for ONE simulated vehicle executed by one warp.
// Every lane carries the (identical) truth state tr[3] (updated in place); lanes stride over landmark ids and an
// ordered ballot compaction keeps the message in ascending-id order (:231-249).  `out` ([max_meas][3] float32, any
// address space) receives the message; returns the number of detections (may exceed max_meas: extra ones dropped).
__device__ __forceinline__ int sim_get_cmd_warp(const int lane, const SimConst& sc, const double* __restrict__ lm_xy,
                                                const int n_lm, const int max_meas, const uint32_t k0, const uint32_t k1,
                                                const uint32_t inst, const uint32_t step, const float fwd, const float ang,
                                                double (&tr)[3], float* out) {
    uint32_t rn[4];
    philox4x32_10(inst, step, 0u, 0u, k0, k1, rn);
    // add noise to the command, clamp (:216-220); msg.fwd / msg.ang are the float32 wire values
    double d = (double)fwd + 2 * sc.V_00 * uniform53(rn[0], rn[1]) - sc.V_00;
    double hdg = (double)ang + 2 * sc.V_11 * uniform53(rn[2], rn[3]) - sc.V_11;
    d = fmax(0.0, fmin(d, sc.d_max));
    hdg = fmax(-sc.th_max, fmin(hdg, sc.th_max));
    double sy, cy;
    sincos(tr[2], &sy, &cy);
    const double tx = tr[0] + d * cy, ty = tr[1] + d * sy, tyaw = tr[2] + hdg;   // :222 (yaw never wrapped)
    tr[0] = tx; tr[1] = ty; tr[2] = tyaw;
    int count = 0;
    for (int base = 0; base < n_lm; base += 32) {
        const int id = base + lane;
        bool vis = false;
        double r = 0.0, beta = 0.0;
        if (id < n_lm) {
            const double dx = lm_xy[2 * id] - tx, dy = lm_xy[2 * id + 1] - ty;
            r = sqrt(dx * dx + dy * dy);                                    // :235
            if (!(r > sc.range_max)) {                                      // :239 (the bearing only matters in range)
                beta = wrap_2pi(atan2(dy, dx) - tyaw);                   // :236-237
                vis = beta > sc.fov_min && beta < sc.fov_max;               // :240-241
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, vis);
        const int pos = count + __popc(ballot & ((1u << lane) - 1u));
        if (vis && pos < max_meas) {
            philox4x32_10(inst, step, 1u + (uint32_t)id, 0u, k0, k1, rn);
            out[3 * pos] = (float)id;                                       // float32 wire, :245-249
            out[3 * pos + 1] = (float)(r + 2 * sc.W_00 * uniform53(rn[0], rn[1]) - sc.W_00);
            out[3 * pos + 2] = (float)(beta + 2 * sc.W_11 * uniform53(rn[2], rn[3]) - sc.W_11);
        }
        count += __popc(ballot);
    }
    return count;
}

// Accuracy terms of one instance against the simulator's truth: position error is the reference's metric
// (plotting_node.py:212-214); RMSE terms and the 3-dof pose NEES are the extension BASELINE.json asks for.
// C = 3x3 pose covariance (row-major), acc[0..5] += {1, ex^2, ey^2, eyaw^2, |e_pos|, NEES}.
__device__ __forceinline__ void pose_error_terms(const double ex, const double ey, const double eyaw, double (&C)[3][3],
                                                 double (&acc)[6]) {
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
    acc[0] += 1.0;
    acc[1] += ex * ex;
    acc[2] += ey * ey;
    acc[3] += eyaw * eyaw;
    acc[4] += sqrt(ex * ex + ey * ey);
    acc[5] += quad / det;
}

}  // namespace slam
