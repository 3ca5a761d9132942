// ekf_stream.cu -- HBM-streaming EKF-SLAM step for sm_100a (known landmark IDs): one CTA per filter instance, but the
// covariance is NEVER staged as a whole.  Every operation of EKF::update (ekf_ws/src/localization_pkg/src/ekf.cpp:37-179)
// is row-local once a handful of "special" rows is known:
//   predict  :61   row i:  P[i][0] += P[i][2] fa,  P[i][1] += P[i][2] fb          (+ rows 0,1 += f * row 2: special rows)
//   update   :140  row i:  K[i] = (P[i][{0,1,2,l,l+1}] H^T) S^-1,  P[i][:] -= K[i] (H P),  x[i] += K[i] nu
//                  with H, S, S^-1, nu and (H P) (2 x n) built from rows {0,1,2,l,l+1} alone
//   insert   :172  row i:  P[i][n], P[i][n+1] = P[i][0..2] G_x^T                    (+ new rows n, n+1 = G_x * rows 0..2)
// so a step is a short sequential "program" (one op per measurement, message order) that every row executes on its
// own registers.  Phase 1 (per CTA): associate the message (ekf.cpp:99-109), load the special rows {0,1,2} + the rows
// of every observed / inserted landmark (<= 3 + 2 k rows) into shared memory and run the program on them, which
// yields each op's scalars and (H P).  Phase 2: the warps stream all other rows HBM -> registers -> HBM in groups of
// four rows with the next group prefetched into registers.  For a group, lane u first runs the program on the five
// scalars of row u it depends on (P[u][0..2] and the two landmark columns of each update, corrected for the earlier
// ops of the step), which gives K_l[u] / the new column values for all ops at a cost of a few instructions per
// row; the row registers then only see  p -= K_l[u] (H P)_l  per op.  P crosses HBM exactly once each way, fully
// coalesced (a row is one contiguous ldg(n)*8-byte run), with no whole-matrix residency and ~35 KB of shared memory
// per CTA whatever n is: 5+ CTAs per SM keep the memory system busy while other CTAs sit in their scalar chains.
// Algorithmic bytes per update: 16 n^2 + 16 n + 12 (k+j) + 8.
//
// Unknown-ID association needs the running x of ALL landmarks after every update (ekf.cpp:91-92), i.e. K for every
// row before the next measurement can be associated; that mode stays on the shared-memory-resident kernel
// (ekf_batch.cu).  Float/double roundings as in SURVEY.md Appendix A; expressions match ekf_batch.cu's core.
#include "common.cuh"

#include <climits>

namespace slam {

constexpr int ST_THREADS = 128;
constexpr int ST_WARPS = ST_THREADS / 32;

struct StreamLaunch {
    int cap_lm;        // landmark capacity this launch is sized for (instances that may outgrow it are deferred)
    int n_cap;         // 3 + 2*cap_lm
    int ldc;           // leading dimension of the special-row tile and of (H P)
    int rmax;          // special rows: 3 + 2*max_meas
    int from_list;
    int off[16];       // shared-memory byte offsets of the StreamSmem arrays (carved on the host)
    int smem_bytes;
};

struct __align__(16) StreamOp {        // one measurement of the step, message order
    double q[4];       // the four distinct quotients of H_x (update)
    double inv[4];     // S^-1: i00 i01 i10 i11           (update)
    double nu[2];      // innovation                      (update)
    double g02, g12;   // G_x(0,2), G_x(1,2)              (insert)
    double xs0, xs1;   // stale landmark mean x_t(i), x_t(i+1) (update, ekf.cpp:115)
    float r, b;
    int kind;          // 0 skip, 1 update, 2 insert
    int i;             // update: state index of the landmark (3 + 2 slot); insert: n before the insertion
    int pos;           // position of row i in the special-row tile (row i+1 follows)
    int hp;            // double2 per live row while this op runs (after it, for an insert)
    int id;
    int pad;
};

struct StreamSmem {
    StreamOp* ops;          // [max_meas]
    double* R;              // [rmax][ldc] special rows
    double* HP;             // [max_meas][2][ldc]
    double* xsr;            // [rmax] running x of the special rows
    double* Ksr;            // [rmax][2]
    double* sc;             // fa fb c s nx0 nx1 nx2 cb sb
    int* sr_row;            // [rmax] state index of special row r
    int* sr_born;           // [rmax] op that creates the row (-1: it exists at step start)
    int* assoc;             // [max_meas]
    int* iscr;              // [0] dead [1] nan [2] sr_count
    float* meas;            // [max_meas][3]
    unsigned char* pos;     // [n_cap + 1] state row -> special-row position, 0xFF = streamed
    unsigned char* srow;    // [n_cap] compact list of the streamed rows
    double2* kscr;          // [ST_WARPS][max_meas + 1][4] per warp: K_l[u] (update) / new column values (insert) of the
                            // group's rows; slot max_meas holds the predicted (P[u][0], P[u][1])
};

enum { QS_FA = 0, QS_FB, QS_C, QS_S, QS_NX0, QS_NX1, QS_NX2, QS_CB, QS_SB };
enum { QI_DEAD = 0, QI_NAN = 1, QI_SR = 2, QI_NSTR = 3 };
enum { Q_NEW = -1, Q_DROPPED = -2 };

enum { O_OPS = 0, O_R, O_HP, O_XSR, O_KSR, O_SC, O_ROW, O_BORN, O_ASSOC, O_ISCR, O_MEAS, O_POS, O_SROW, O_KSCR };

// host: lay the arrays out once per launch; the kernel only adds constant-bank offsets to the shared-memory base
static void stream_smem_layout(const int max_meas, StreamLaunch& L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return (int)o; };
    L.off[O_OPS] = take(sizeof(StreamOp) * max_meas);
    L.off[O_R] = take(sizeof(double) * (size_t)L.rmax * L.ldc);
    L.off[O_HP] = take(sizeof(double) * (size_t)max_meas * 2 * L.ldc);
    L.off[O_XSR] = take(sizeof(double) * L.rmax);
    L.off[O_KSR] = take(sizeof(double) * 2 * L.rmax);
    L.off[O_SC] = take(sizeof(double) * 12);
    L.off[O_ROW] = take(sizeof(int) * L.rmax);
    L.off[O_BORN] = take(sizeof(int) * L.rmax);
    L.off[O_ASSOC] = take(sizeof(int) * max_meas);
    L.off[O_ISCR] = take(sizeof(int) * 8);
    L.off[O_MEAS] = take(sizeof(float) * 3 * max_meas);
    L.off[O_POS] = take((size_t)L.n_cap + 4);
    L.off[O_SROW] = take((size_t)L.n_cap + 4);
    L.off[O_KSCR] = take(sizeof(double2) * ST_WARPS * (size_t)(max_meas + 1) * 4);
    L.smem_bytes = (int)off;
}

__device__ __forceinline__ void stream_smem_bind(const StreamLaunch& L, unsigned char* base, StreamSmem& s) {
    s.ops = (StreamOp*)(base + L.off[O_OPS]); s.R = (double*)(base + L.off[O_R]); s.HP = (double*)(base + L.off[O_HP]);
    s.xsr = (double*)(base + L.off[O_XSR]); s.Ksr = (double*)(base + L.off[O_KSR]); s.sc = (double*)(base + L.off[O_SC]);
    s.sr_row = (int*)(base + L.off[O_ROW]); s.sr_born = (int*)(base + L.off[O_BORN]); s.assoc = (int*)(base + L.off[O_ASSOC]);
    s.iscr = (int*)(base + L.off[O_ISCR]); s.meas = (float*)(base + L.off[O_MEAS]); s.pos = (unsigned char*)(base + L.off[O_POS]);
    s.srow = (unsigned char*)(base + L.off[O_SROW]); s.kscr = (double2*)(base + L.off[O_KSCR]);
}

// ---- streamed rows: a group of GR rows lives in registers, NS double2 per lane per row (double2 index lane + 32 s);
//      lane u < cnt additionally carries the scalars of row u that the step's program depends on.
constexpr int GR = 4;

template <int NS>
struct Group {
    double2 p[GR][NS];
    double c0, c1, c2, xi;      // lane u: P[row_u][0..2], x[row_u]
    int row;                    // lane u: state index of row u
    int cnt;                    // rows in the group (warp-uniform)
};

template <int NS>
__device__ __forceinline__ void group_load(Group<NS>& g, const StreamSmem& s, const double2* __restrict__ gP2_lane,
                                           const double* __restrict__ gP, const double* __restrict__ gx, const int ld,
                                           const int first, const int n_str, const int hp0, const int lane) {
    // gP2_lane = (double2*)gP + lane: a row is double2 index row * (ld / 2) + lane (+ 32 for the second slot)
    g.cnt = n_str - first < GR ? n_str - first : GR;
    if (g.cnt <= 0) { g.cnt = 0; return; }
    const int ldh = ld >> 1;
    const uchar4 r4 = *reinterpret_cast<const uchar4*>(s.srow + first);      // first is a multiple of GR = 4
    const int rows[GR] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int u = 0; u < GR; ++u) {
        if (u < g.cnt) {
            const double2* src = gP2_lane + rows[u] * ldh;
#pragma unroll
            for (int q = 0; q < NS; ++q) g.p[u][q] = (lane + 32 * q < hp0) ? src[32 * q] : make_double2(0.0, 0.0);
        }
    }
    g.row = 0; g.c0 = g.c1 = g.c2 = g.xi = 0.0;
    if (lane < g.cnt) {
        g.row = s.srow[first + lane];
        const double* prow = gP + g.row * ld;
        g.c0 = prow[0]; g.c1 = prow[1]; g.c2 = prow[2];
        g.xi = gx[g.row];
    }
}

// run the step's program on the group and write it back
template <int NS>
__device__ __forceinline__ void group_run_store(Group<NS>& g, const StreamSmem& s, double2* __restrict__ gP2_lane,
                                                double* __restrict__ gP, double* __restrict__ gx, const int ld, const int ldc, const int nm, const int max_meas, const bool predict,
                                                const int hp1, const int lane, const int warp) {
    double2* ks = s.kscr + (size_t)warp * (max_meas + 1) * GR;
    // ---- lane u: the program on the scalars of row u
    if (lane < g.cnt) {
        const double* prow = gP + g.row * ld;
        double c0 = g.c0, c1 = g.c1, c2 = g.c2, xi = g.xi;
        if (predict) {                                            // T F_x^T: cols 0,1 pick up col 2 (ekf.cpp:61)
            c0 = c0 + c2 * s.sc[QS_FA];
            c1 = c1 + c2 * s.sc[QS_FB];
        }
        ks[max_meas * GR + lane] = make_double2(c0, c1);
        for (int l = 0; l < nm; ++l) {
            const StreamOp& op = s.ops[l];
            const int kind = op.kind;
            if (kind == 1) {
                // K[u] = (P[u][{0,1,2,i,i+1}] H^T) S^-1 (:135) with the landmark columns brought up to date
                const int i = op.i;
                double c3 = prow[i], c4 = prow[i + 1];
                for (int m = 0; m < l; ++m) {
                    if (s.ops[m].kind == 1) {
                        const double2 km = ks[m * GR + lane];
                        const double* hm = s.HP + (size_t)m * 2 * ldc;
                        c3 = c3 - (km.x * hm[i] + km.y * hm[ldc + i]);
                        c4 = c4 - (km.x * hm[i + 1] + km.y * hm[ldc + i + 1]);
                    }
                }
                const double q0 = op.q[0], q1 = op.q[1], q2 = op.q[2], q3 = op.q[3];
                const double H[10] = {-q0, -q1, 0.0, q0, q1, q2, -q3, -1.0, -q2, q3};
                double a0 = c0 * H[0]; a0 += c1 * H[1]; a0 += c2 * H[2]; a0 += c3 * H[3]; a0 += c4 * H[4];
                double a1 = c0 * H[5]; a1 += c1 * H[6]; a1 += c2 * H[7]; a1 += c3 * H[8]; a1 += c4 * H[9];
                const double k0 = a0 * op.inv[0] + a1 * op.inv[2];
                const double k1 = a0 * op.inv[1] + a1 * op.inv[3];
                ks[l * GR + lane] = make_double2(k0, k1);
                const double* hl = s.HP + (size_t)l * 2 * ldc;
                c0 = c0 - (k0 * hl[0] + k1 * hl[ldc + 0]);
                c1 = c1 - (k0 * hl[1] + k1 * hl[ldc + 1]);
                c2 = c2 - (k0 * hl[2] + k1 * hl[ldc + 2]);
                xi = xi + (k0 * op.nu[0] + k1 * op.nu[1]);          // :138
            } else if (kind == 2) {
                // new columns n, n+1 of this row: P[u][0..2] G_x^T (:153-172)
                double v0 = c0 * 1.0; v0 += c1 * 0.0; v0 += c2 * op.g02;
                double v1 = c0 * 0.0; v1 += c1 * 1.0; v1 += c2 * op.g12;
                ks[l * GR + lane] = make_double2(v0, v1);
            }
        }
        gx[g.row] = xi;
        if (!isfinite(xi)) s.iscr[QI_NAN] = 1;
    }
    __syncwarp();
    // ---- the rows: predicted cols 0,1, then p -= K_l[u] (H P)_l per update / the new columns per insert
    if (predict && lane == 0) {
#pragma unroll
        for (int u = 0; u < GR; ++u) if (u < g.cnt) g.p[u][0] = ks[max_meas * GR + u];
    }
    for (int l = 0; l < nm; ++l) {
        const StreamOp& op = s.ops[l];
        const int kind = op.kind;
        if (kind == 1) {
            const double* hl = s.HP + (size_t)l * 2 * ldc;
            const int hp = op.hp;
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                const int jp = lane + 32 * q;
                if (jp < hp) {
                    const double2 h0 = *reinterpret_cast<const double2*>(hl + 2 * jp);
                    const double2 h1 = *reinterpret_cast<const double2*>(hl + ldc + 2 * jp);
#pragma unroll
                    for (int u = 0; u < GR; ++u) {
                        if (u < g.cnt) {
                            const double2 k = ks[l * GR + u];
                            g.p[u][q].x = g.p[u][q].x - (k.x * h0.x + k.y * h1.x);      // :140
                            g.p[u][q].y = g.p[u][q].y - (k.x * h0.y + k.y * h1.y);
                        }
                    }
                }
            }
        } else if (kind == 2) {
            const int n = op.i;                                  // odd
            const int ja = (n - 1) >> 1, jb = (n + 1) >> 1;      // double2 holding col n (.y) / cols n+1 (.x), n+2 (.y, pad)
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                const int jp = lane + 32 * q;
#pragma unroll
                for (int u = 0; u < GR; ++u) {
                    if (u < g.cnt) {
                        const double2 v = ks[l * GR + u];
                        if (jp == ja) g.p[u][q].y = v.x;
                        if (jp == jb) g.p[u][q] = make_double2(v.y, 0.0);
                    }
                }
            }
        }
    }
    // ---- store; the diagonal entry feeds the non-finite check like the resident kernel's commit
    const int ldh = ld >> 1;
    bool bad = false;
#pragma unroll
    for (int u = 0; u < GR; ++u) {
        if (u < g.cnt) {
            const int row = __shfl_sync(0xffffffffu, g.row, u);
            double2* dst = gP2_lane + row * ldh;
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                const int jp = lane + 32 * q;
                if (jp < hp1) dst[32 * q] = g.p[u][q];
                if (jp == (row >> 1)) bad |= !isfinite((row & 1) ? g.p[u][q].y : g.p[u][q].x);
            }
        }
    }
    if (bad) s.iscr[QI_NAN] = 1;
    __syncwarp();
}

// One reference EKF::update for instance `inst`.
template <int NS>
__device__ __forceinline__ void stream_instance(const BatchState& b, const FilterConst& fc, const StepInputs& in, const int phases,
                                                const StreamLaunch& L, const StreamSmem& s, const int inst) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldc = L.ldc, ld = b.fixed_ld;
    const bool predict = (phases & STEP_PREDICT) != 0;

    const int4 meta_in = b.meta[inst];
    int nm = (phases & STEP_UPDATE) ? in.n_meas[inst] : 0;
    int status = meta_in.y;
    if (status & SLAM_STATUS_SAME_STEP_REMATCH) return;          // the reference process is dead past this point
    const int M0 = meta_in.x;
    const int n0 = 3 + 2 * M0;
    if (nm > b.max_meas) { nm = b.max_meas; status |= SLAM_STATUS_MEAS_OVERFLOW; }
    if (M0 + nm > L.cap_lm && L.cap_lm < b.max_lm) {
        if (tid == 0) b.retry_list[atomicAdd(b.retry_count, 1)] = inst;     // tile too small: defer, untouched
        return;
    }
    double* gP = b.P + (size_t)inst * b.p_stride;
    double* gx = b.x + (size_t)inst * b.x_stride;
    const int* gids = b.ids + (size_t)inst * b.max_lm;

    // ---- A: message, row map, predict trigonometry
    for (int i = tid; i < 3 * nm; i += ST_THREADS) s.meas[i] = in.meas[(size_t)inst * b.max_meas * 3 + i];
    for (int i = tid; 4 * i <= L.n_cap; i += ST_THREADS) reinterpret_cast<unsigned*>(s.pos)[i] = 0xFFFFFFFFu;
    if (tid == 0) { s.iscr[QI_DEAD] = 0; s.iscr[QI_NAN] = 0; }
    if (predict && tid == ST_THREADS - 32) {
        const float d_d = in.fwd[in.cmd_stride ? inst : 0], d_th = in.ang[in.cmd_stride ? inst : 0];
        const double x0 = gx[0], x1 = gx[1], th = gx[2];
        double sn, c;
        sincos(th, &sn, &c);
        s.sc[QS_FA] = (double)(-1 * d_d) * sn;        // F_x(0,2), :48
        s.sc[QS_FB] = (double)d_d * c;                // F_x(1,2), :49
        s.sc[QS_C] = c; s.sc[QS_S] = sn;
        const float dv = d_d + fc.v_d;                // float add, :57-58
        s.sc[QS_NX0] = x0 + (double)dv * c;
        s.sc[QS_NX1] = x1 + (double)dv * sn;
        s.sc[QS_NX2] = wrap_2pi(th + (double)d_th + (double)fc.v_th);   // :59
    }
    __syncthreads();
    // ---- B: association of the whole message (:99-109) and the special-row list, warp 0
    if (warp == 0) {
        int M_run = M0;
        bool dead = false;
        for (int l = 0; l < nm; ++l) {
            const int id = (int)s.meas[3 * l];                                 // :101
            int cand = INT_MAX;
            for (int j = lane; j < M0; j += 32)
                if (gids[j] == id) { cand = j; break; }
            cand = __reduce_min_sync(0xffffffffu, cand);                       // first match in ascending slot order
            int code = cand;
            if (cand == INT_MAX) {
                bool dup = false;
                for (int q = lane; q < l; q += 32) dup |= (s.assoc[q] == Q_NEW) && ((int)s.meas[3 * q] == id);
                if (__any_sync(0xffffffffu, dup)) { dead = true; break; }      // :115 would index x_t out of range
                if (M_run < b.max_lm) { code = Q_NEW; ++M_run; } else code = Q_DROPPED;
            }
            if (lane == 0) s.assoc[l] = code;
            __syncwarp();
        }
        if (lane == 0) {
            s.iscr[QI_DEAD] = dead ? 1 : 0;
            int cnt = 3, n_run = n0;
            for (int r = 0; r < 3; ++r) { s.sr_row[r] = r; s.sr_born[r] = -1; s.pos[r] = (unsigned char)r; }
            for (int l = 0; l < nm && !dead; ++l) {
                StreamOp& op = s.ops[l];
                const int code = s.assoc[l];
                op.r = s.meas[3 * l + 1]; op.b = s.meas[3 * l + 2]; op.id = (int)s.meas[3 * l];
                if (code >= 0) {
                    const int i = 3 + 2 * code;
                    if (s.pos[i] == 0xFF) {
                        s.sr_row[cnt] = i; s.sr_row[cnt + 1] = i + 1; s.sr_born[cnt] = -1; s.sr_born[cnt + 1] = -1;
                        s.pos[i] = (unsigned char)cnt; s.pos[i + 1] = (unsigned char)(cnt + 1);
                        cnt += 2;
                    }
                    op.kind = 1; op.i = i; op.pos = s.pos[i]; op.hp = ldg_of(n_run) >> 1;
                } else if (code == Q_NEW) {
                    s.sr_row[cnt] = n_run; s.sr_row[cnt + 1] = n_run + 1; s.sr_born[cnt] = l; s.sr_born[cnt + 1] = l;
                    s.pos[n_run] = (unsigned char)cnt; s.pos[n_run + 1] = (unsigned char)(cnt + 1);
                    op.kind = 2; op.i = n_run; op.pos = cnt;
                    cnt += 2; n_run += 2;
                    op.hp = ldg_of(n_run) >> 1;
                } else { op.kind = 0; op.i = 0; op.pos = 0; op.hp = ldg_of(n_run) >> 1; }
            }
            s.iscr[QI_SR] = cnt;
        }
    }
    __syncthreads();
    if (s.iscr[QI_DEAD]) {
        // frozen at the last committed state: only the status word changes
        if (tid == 0) b.meta[inst] = make_int4(meta_in.x, status | SLAM_STATUS_SAME_STEP_REMATCH, meta_in.z, 0);
        return;
    }
    const int sr_count = s.iscr[QI_SR];
    const int hp0 = ldg_of(n0) >> 1;

    // ---- compact list of the streamed rows (warp 1), needed by every warp's row pipeline
    if (warp == 1) {
        int cnt = 0;
        for (int base = 0; base < n0; base += 32) {
            const int row = base + lane;
            const bool str = row >= 3 && row < n0 && s.pos[row] == 0xFF;
            const unsigned bal = __ballot_sync(0xffffffffu, str);
            if (str) s.srow[cnt + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)row;
            cnt += __popc(bal);
        }
        if (lane == 0) s.iscr[QI_NSTR] = cnt;
    }
    // ---- C: load the special rows that exist (warp per row) and the stale landmark means
    for (int r = warp; r < sr_count; r += ST_WARPS) {
        if (s.sr_born[r] < 0) {
            const int row = s.sr_row[r];
            const double2* src = reinterpret_cast<const double2*>(gP + (size_t)row * ld);
            double2* dst = reinterpret_cast<double2*>(s.R + (size_t)r * ldc);
            for (int jp = lane; jp < hp0; jp += 32) dst[jp] = src[jp];
            if (lane == 0) s.xsr[r] = gx[row];
        }
    }
    for (int l = tid; l < nm; l += ST_THREADS)
        if (s.ops[l].kind == 1) { s.ops[l].xs0 = gx[s.ops[l].i]; s.ops[l].xs1 = gx[s.ops[l].i + 1]; }
    __syncthreads();

    // ---- D: predict on the special rows (:43-61)
    if (predict) {
        const double fa = s.sc[QS_FA], fb = s.sc[QS_FB];
        for (int j = tid; j < n0; j += ST_THREADS) {
            if (j >= 3) {
                const double p2 = s.R[2 * ldc + j];
                s.R[j] = s.R[j] + fa * p2;
                s.R[ldc + j] = s.R[ldc + j] + fb * p2;
            } else if (j == 0) {
                const double c = s.sc[QS_C], sn = s.sc[QS_S];
                double T[3][3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double p2 = s.R[2 * ldc + q];
                    T[0][q] = s.R[q] + fa * p2;
                    T[1][q] = s.R[ldc + q] + fb * p2;
                    T[2][q] = p2;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double t2 = T[i][2];
                    double p0 = T[i][0] + t2 * fa;
                    double p1 = T[i][1] + t2 * fb;
                    if (i == 0) { const double cV = c * fc.V00; p0 += cV * c; p1 += cV * sn; }
                    if (i == 1) { const double sV = sn * fc.V00; p0 += sV * c; p1 += sV * sn; }
                    s.R[i * ldc + 0] = p0;
                    s.R[i * ldc + 1] = p1;
                    s.R[i * ldc + 2] = (i == 2) ? t2 + fc.V11 : t2;
                }
                s.xsr[0] = s.sc[QS_NX0]; s.xsr[1] = s.sc[QS_NX1]; s.xsr[2] = s.sc[QS_NX2];
            }
        }
        for (int r = 3 + tid; r < sr_count; r += ST_THREADS) {
            if (s.sr_born[r] < 0) {
                double* row = s.R + (size_t)r * ldc;
                const double t2 = row[2];
                row[0] = row[0] + t2 * fa;
                row[1] = row[1] + t2 * fb;
            }
        }
        __syncthreads();
    }

    // ---- E: the program on the special rows, op by op (:73-174)
    int n_run = n0, M_run = M0, n_upd = 0;
    for (int l = 0; l < nm; ++l) {
        StreamOp& op = s.ops[l];
        const int kind = op.kind;
        if (kind == 0) { status |= SLAM_STATUS_CAPACITY; continue; }
        if (kind == 1) {
            ++n_upd;
            const int i = op.i, pi = op.pos;
            const bool nu_thread = tid == ST_THREADS - 32;
            if (warp == 0 || nu_thread) {
                // landmark from the stale x_t, vehicle from the running x_pred (:115)
                const double dx = op.xs0 - s.xsr[0], dy = op.xs1 - s.xsr[1];
                const float dist = (float)sqrt(dx * dx + dy * dy);         // :115
                if (nu_thread) {
                    const float ang = (float)wrap_2pi(atan2(dy, dx) - s.xsr[2]);   // :129
                    op.nu[0] = (double)(op.r - dist - fc.w_r);             // all-float arithmetic, :130
                    op.nu[1] = (double)(op.b - ang - fc.w_b);              // :131
                }
                if (warp == 0) {
                    const double dd = (double)dist;
                    const double d2 = (double)(dist * dist);               // float product, :120
                    const int ql = lane & 3;
                    const double qv = ((ql == 0 || ql == 3) ? dx : dy) / ((ql < 2) ? dd : d2);
                    const double q0 = __shfl_sync(0xffffffffu, qv, 0);     // dx / dist
                    const double q1 = __shfl_sync(0xffffffffu, qv, 1);     // dy / dist
                    const double q2 = __shfl_sync(0xffffffffu, qv, 2);     // dy / dist^2
                    const double q3 = __shfl_sync(0xffffffffu, qv, 3);     // dx / dist^2
                    const double H[10] = {-q0, -q1, 0.0, q0, q1, q2, -q3, -1.0, -q2, q3};
                    // S = H P H^T + W (:133) from the 5x5 sub-block: lane c < 5 forms column hc[c] of H P
                    const int hcl = (lane < 3) ? lane : i + (lane - 3);
                    double g0 = 0.0, g1 = 0.0;
                    if (lane < 5) {
                        const double p0 = s.R[hcl], p1 = s.R[ldc + hcl], p2 = s.R[2 * ldc + hcl];
                        const double p3 = s.R[(size_t)pi * ldc + hcl], p4 = s.R[(size_t)(pi + 1) * ldc + hcl];
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
                    // S^-1 by partial-pivot LU like Eigen's dynamic inverse() (:135), one column per lane
                    const bool sw = fabs(S10) > fabs(S00);
                    const double a00 = sw ? S10 : S00, a01 = sw ? S11 : S01, a10 = sw ? S00 : S10, a11 = sw ? S01 : S11;
                    const double l10 = a10 / a00, u11 = a11 - l10 * a01;
                    const bool col1 = (lane & 1) != 0;
                    const double b0 = (sw != col1) ? 0.0 : 1.0, b1 = (sw != col1) ? 1.0 : 0.0;
                    const double y1 = b1 - l10 * b0;
                    const double i1c = y1 / u11;
                    const double i0c = (b0 - a01 * i1c) / a00;
                    if (lane < 2) { op.inv[lane] = i0c; op.inv[2 + lane] = i1c; }
                    if (lane == 0) { op.q[0] = q0; op.q[1] = q1; op.q[2] = q2; op.q[3] = q3; }
                }
            }
            __syncthreads();
            // H P (2 x n) for every row of the step, K for the special rows
            {
                const double q0 = op.q[0], q1 = op.q[1], q2 = op.q[2], q3 = op.q[3];
                const double H[10] = {-q0, -q1, 0.0, q0, q1, q2, -q3, -1.0, -q2, q3};
                double* hpl = s.HP + (size_t)l * 2 * ldc;
                const int ldg = ldg_of(n_run);
                for (int idx = tid; idx < ldg + sr_count; idx += ST_THREADS) {
                    if (idx < ldg) {
                        const int j = idx;
                        double h0 = 0.0, h1 = 0.0;
                        if (j < n_run) {
                            const double p0 = s.R[j], p1 = s.R[ldc + j], p2 = s.R[2 * ldc + j];
                            const double p3 = s.R[(size_t)pi * ldc + j], p4 = s.R[(size_t)(pi + 1) * ldc + j];
                            h0 = H[0] * p0; h0 += H[1] * p1; h0 += H[2] * p2; h0 += H[3] * p3; h0 += H[4] * p4;
                            h1 = H[5] * p0; h1 += H[6] * p1; h1 += H[7] * p2; h1 += H[8] * p3; h1 += H[9] * p4;
                        }
                        hpl[j] = h0; hpl[ldc + j] = h1;
                    } else {
                        const int r = idx - ldg;
                        if (s.sr_born[r] < l) {
                            const double* row = s.R + (size_t)r * ldc;
                            const double p0 = row[0], p1 = row[1], p2 = row[2], p3 = row[i], p4 = row[i + 1];
                            double a0 = p0 * H[0]; a0 += p1 * H[1]; a0 += p2 * H[2]; a0 += p3 * H[3]; a0 += p4 * H[4];
                            double a1 = p0 * H[5]; a1 += p1 * H[6]; a1 += p2 * H[7]; a1 += p3 * H[8]; a1 += p4 * H[9];
                            s.Ksr[2 * r] = a0 * op.inv[0] + a1 * op.inv[2];
                            s.Ksr[2 * r + 1] = a0 * op.inv[1] + a1 * op.inv[3];
                        }
                    }
                }
            }
            __syncthreads();
            // rank-2 update and x of the special rows
            {
                const double* hpl = s.HP + (size_t)l * 2 * ldc;
                const int hp = ldg_of(n_run) >> 1;
                for (int e = tid; e < sr_count * hp; e += ST_THREADS) {
                    const int r = e / hp, jp = e - r * hp;
                    if (s.sr_born[r] < l) {
                        const double k0 = s.Ksr[2 * r], k1 = s.Ksr[2 * r + 1];
                        const double2 h0 = *reinterpret_cast<const double2*>(hpl + 2 * jp);
                        const double2 h1 = *reinterpret_cast<const double2*>(hpl + ldc + 2 * jp);
                        double2* pp = reinterpret_cast<double2*>(s.R + (size_t)r * ldc + 2 * jp);
                        double2 p = *pp;
                        p.x = p.x - (k0 * h0.x + k1 * h1.x);
                        p.y = p.y - (k0 * h0.y + k1 * h1.y);
                        *pp = p;
                    }
                }
                for (int r = tid; r < sr_count; r += ST_THREADS) {
                    if (s.sr_born[r] < l) {
                        double xv = s.xsr[r] + (s.Ksr[2 * r] * op.nu[0] + s.Ksr[2 * r + 1] * op.nu[1]);   // :138
                        if (r == 2) xv = wrap_2pi(xv);                                                    // :139
                        s.xsr[r] = xv;
                    }
                }
            }
            __syncthreads();
        } else {
            // -------- landmark insertion, :141-173 (the new rows join the special rows at position op.pos)
            const int n = op.i, pn = op.pos;
            if (tid == ST_THREADS - 32) {
                double sb, cb; sincos(s.xsr[2] + (double)op.b, &sb, &cb);
                s.sc[QS_CB] = cb; s.sc[QS_SB] = sb;
                op.g02 = -(double)op.r * sb; op.g12 = (double)op.r * cb;       // G_x(0,2), G_x(1,2), :162,165
            }
            __syncthreads();
            const double g02 = op.g02, g12 = op.g12;
            for (int idx = tid; idx < n + sr_count; idx += ST_THREADS) {
                if (idx < n) {
                    const int j = idx;                                         // new rows over the old columns
                    const double p0 = s.R[j], p1 = s.R[ldc + j], p2 = s.R[2 * ldc + j];
                    double t0 = 1.0 * p0; t0 += 0.0 * p1; t0 += g02 * p2;
                    double t1 = 0.0 * p0; t1 += 1.0 * p1; t1 += g12 * p2;
                    s.R[(size_t)pn * ldc + j] = t0;
                    s.R[(size_t)(pn + 1) * ldc + j] = t1;
                } else {
                    const int r = idx - n;                                     // new columns of the special rows
                    if (s.sr_born[r] < l) {
                        double* row = s.R + (size_t)r * ldc;
                        const double q0 = row[0], q1 = row[1], q2 = row[2];
                        double c0 = q0 * 1.0; c0 += q1 * 0.0; c0 += q2 * g02;
                        double c1 = q0 * 0.0; c1 += q1 * 1.0; c1 += q2 * g12;
                        row[n] = c0; row[n + 1] = c1; row[n + 2] = 0.0;
                    }
                }
            }
            if (tid == ST_THREADS - 32) {
                // new 2x2 block: G_x P_vv G_x^T + G_z W G_z^T, :155-172
                const double cb = s.sc[QS_CB], sb = s.sc[QS_SB], r = (double)op.r;
                const double gxm[2][3] = {{1.0, 0.0, g02}, {0.0, 1.0, g12}};
                const double gz[2][2] = {{cb, -r * sb}, {sb, r * cb}};
                const double Wm[2][2] = {{fc.W00, 0.0}, {0.0, fc.W11}};
                double T3[2][3], T2[2][2];
                for (int rr = 0; rr < 2; ++rr) {
                    for (int k = 0; k < 3; ++k) {
                        double t = gxm[rr][0] * s.R[k]; t += gxm[rr][1] * s.R[ldc + k]; t += gxm[rr][2] * s.R[2 * ldc + k];
                        T3[rr][k] = t;
                    }
                    for (int c2 = 0; c2 < 2; ++c2) T2[rr][c2] = gz[rr][0] * Wm[0][c2] + gz[rr][1] * Wm[1][c2];
                }
                for (int rr = 0; rr < 2; ++rr) {
                    for (int c2 = 0; c2 < 2; ++c2) {
                        double t = T3[rr][0] * gxm[c2][0]; t += T3[rr][1] * gxm[c2][1]; t += T3[rr][2] * gxm[c2][2];
                        t += T2[rr][0] * gz[c2][0]; t += T2[rr][1] * gz[c2][1];
                        s.R[(size_t)(pn + rr) * ldc + n + c2] = t;
                    }
                    s.R[(size_t)(pn + rr) * ldc + n + 2] = 0.0;
                }
                s.xsr[pn] = s.xsr[0] + r * cb;                                 // :147
                s.xsr[pn + 1] = s.xsr[1] + r * sb;                             // :148
                b.ids[(size_t)inst * b.max_lm + M_run] = op.id;               // :150
            }
            M_run += 1; n_run += 2;
            __syncthreads();
        }
    }
    const int n1 = n_run, hp1 = ldg_of(n1) >> 1;

    // ---- F: phase 2, stream every other row through registers: groups of GR rows, the next group prefetched
    {
        const int n_str = s.iscr[QI_NSTR];
        Group<NS> A;
        double2* gP2_lane = reinterpret_cast<double2*>(gP) + lane;
        for (int first = warp * GR; first < n_str; first += ST_WARPS * GR) {
            group_load<NS>(A, s, gP2_lane, gP, gx, ld, first, n_str, hp0, lane);
            group_run_store<NS>(A, s, gP2_lane, gP, gx, ld, ldc, nm, b.max_meas, predict, hp1, lane, warp);
        }
    }

    // ---- G: write the special rows back and commit (:176-177)
    for (int r = warp; r < sr_count; r += ST_WARPS) {
        const int row = s.sr_row[r];
        const double2* src = reinterpret_cast<const double2*>(s.R + (size_t)r * ldc);
        double2* dst = reinterpret_cast<double2*>(gP + (size_t)row * ld);
        for (int jp = lane; jp < hp1; jp += 32) dst[jp] = src[jp];
        if (lane == 0) {
            const double xv = s.xsr[r];
            gx[row] = xv;
            if (!isfinite(xv) || !isfinite(s.R[(size_t)r * ldc + row])) s.iscr[QI_NAN] = 1;
        }
    }
    for (int i = tid; i < nm; i += ST_THREADS) { const int a = s.assoc[i]; b.assoc[(size_t)inst * b.max_meas + i] = a < 0 ? -1 : a; }
    __syncthreads();
    if (tid == 0) {
        if (s.iscr[QI_NAN]) status |= SLAM_STATUS_NAN;
        b.meta[inst] = make_int4(M_run, status, meta_in.z + (predict ? 1 : 0),     // timestep, :39
                                 (phases & STEP_UPDATE) ? nm : meta_in.w);
        if (M_run > M0) atomicMax(b.max_M, M_run);
        const double nd = (double)n1;
        double* st = b.stats + (size_t)inst * SLAM_NUM_STATS;
        st[8] += 16.0 * nd * nd + 16.0 * nd + 12.0 * nm + 8.0;
        st[9] += 4.0 * (double)n_upd * nd * nd;
        st[10] += nd;
        st[11] += (double)nm;
    }
}

template <int NS>
__global__ void __launch_bounds__(ST_THREADS, 6)
ekf_stream_kernel(BatchState b, FilterConst fc, StepInputs in, int phases, StreamLaunch L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StreamSmem s;
    stream_smem_bind(L, smem_raw, s);
    if (!L.from_list) {
        stream_instance<NS>(b, fc, in, phases, L, s, (int)blockIdx.x);
        return;
    }
    const int count = *b.retry_count;
    for (int q = blockIdx.x; q < count; q += gridDim.x) {
        stream_instance<NS>(b, fc, in, phases, L, s, b.retry_list[q]);
        __syncthreads();
    }
}

static StreamLaunch make_stream_launch(const BatchState& b, int cap_lm, int from_list) {
    StreamLaunch L;
    L.cap_lm = cap_lm < b.max_lm ? cap_lm : b.max_lm;
    if (L.cap_lm < 1) L.cap_lm = 1;
    L.n_cap = 3 + 2 * L.cap_lm;
    L.ldc = ldg_of(L.n_cap) + 2;
    L.rmax = 3 + 2 * b.max_meas;
    L.from_list = from_list;
    stream_smem_layout(b.max_meas, L);
    return L;
}

bool ekf_stream_supported(const BatchState& b) {
    if (3 + 2 * b.max_lm > 126) return false;                       // two double2 per lane per row
    if (3 + 2 * b.max_meas > 250) return false;                     // special-row positions are bytes
    return make_stream_launch(b, b.max_lm, 0).smem_bytes <= 200 * 1024;
}

cudaError_t ekf_stream_configure(const BatchState& b) {
    const int bytes = make_stream_launch(b, b.max_lm, 0).smem_bytes;
    cudaError_t e = cudaFuncSetAttribute(ekf_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ekf_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

static cudaError_t launch_stream(const BatchState& b, const FilterConst& fc, const StepInputs& in, int phases,
                                 const StreamLaunch& L, int grid, cudaStream_t st) {
    const size_t smem = (size_t)L.smem_bytes;
    if (L.n_cap <= 62) ekf_stream_kernel<1><<<grid, ST_THREADS, smem, st>>>(b, fc, in, phases, L);
    else ekf_stream_kernel<2><<<grid, ST_THREADS, smem, st>>>(b, fc, in, phases, L);
    return cudaGetLastError();
}

// cap_hint: landmark capacity to size this launch for (<= 0 or >= max_lm: full capacity, no retry pass).
cudaError_t launch_ekf_stream_step(const BatchState& b, const FilterConst& fc, const StepInputs& in, int phases, int cap_hint,
                                   cudaStream_t st) {
    const bool limited = cap_hint > 0 && cap_hint < b.max_lm;
    const StreamLaunch L = make_stream_launch(b, limited ? cap_hint : b.max_lm, 0);
    if (limited) {
        cudaError_t e = cudaMemsetAsync(b.retry_count, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = launch_stream(b, fc, in, phases, L, b.batch, st);
    if (e != cudaSuccess || !limited) return e;
    const StreamLaunch R = make_stream_launch(b, b.max_lm, 1);
    return launch_stream(b, fc, in, phases, R, b.batch < 148 * 4 ? b.batch : 148 * 4, st);
}

}  // namespace slam
