// Batched dop_sim quadrotor plant, float64 like the reference.
// Restates /root/reference/dop_sim/scripts/quadrotor/ : AtpRate (b_autopilot/atp_rate.py:60-112) with its
// three PID rate loops (pid_control.py:36-65), QdDynamics (a_dynamics/qd_dynamics.py:75-248: motor lag,
// rotor model, drag, pairwise downwash, RK4 of the 13-state rigid body rigid_body_use_vw.py:32-108 /
// tools/ode.py:22-31, quaternion renormalisation, derived angles).  One thread per quadrotor; the CTA
// moves its [128][35] state rows through shared memory with coalesced accesses; the pairwise downwash
// streams the position snapshot of the agent's group through shared memory tiles.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ndp {

constexpr int PL_NS = 35;      // state row (params/def_mul_states.py:10-31)
constexpr int PL_THREADS = 128;

struct PlantCfg {
    long long n, group;  // downwash couples agents of the same contiguous group (group = n: the reference's all-pairs)
    double dt, motor_alpha;
    int has_downwash, has_motor;
    // physical_param.py
    double mass, gravity, iyy, g1, g2, g3, g4, g5, g6, g7, g8;
    double o_min, o_max, o_min_sat, o_max_sat, k_t, kd_x, kd_y, kd_z, k_h;
    double dw_h, dw_v, rp, k_d1, k_d2, k_d3, half_pi_f32;
    double G1[16];
};

struct AutopilotCfg {
    long long n;
    double ts_ctl, voltage_cf, k_th, b_th, k_t, a1, a2, u_limit;
    double kp[3], ki[3], kd[3];
    double G1inv[16];
};

// AtpRate.forward: thrust from throttle, three PID loops, mixer, sqrt -> rotor speed commands (kRPM)
__global__ void plant_autopilot_kernel(const AutopilotCfg c, const double* __restrict__ state, const double* __restrict__ cmd,
                                       double* __restrict__ pid, double* __restrict__ delta) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    double thrust = 4 * (cmd[i * 4 + 3] * c.k_th + c.b_th) * c.voltage_cf;
    if (thrust < 0) thrust = 0;
    double tq[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double* ps = pid + i * 9 + a * 3;  // integrator, error_delay_1, error_dot_delay_1
        const double err = cmd[i * 4 + a] - state[i * PL_NS + 19 + a];
        double integ = ps[0] + (c.ts_ctl / 2) * (err + ps[1]);
        const double err_dot = c.a1 * ps[2] + c.a2 * (err - ps[1]);
        const double u = c.kp[a] * err + c.ki[a] * integ + c.kd[a] * err_dot;
        const double u_sat = (u <= -c.u_limit) ? -c.u_limit : ((u >= c.u_limit) ? c.u_limit : u);
        if (fabs(c.ki[a]) > 0.0001) integ = integ + (c.ts_ctl / c.ki[a]) * (u_sat - u);
        ps[0] = integ; ps[1] = err; ps[2] = err_dot;
        tq[a] = u_sat;
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        double t = c.G1inv[m * 4] * thrust + c.G1inv[m * 4 + 1] * tq[0] + c.G1inv[m * 4 + 2] * tq[1] + c.G1inv[m * 4 + 3] * tq[2];
        if (t < 0) t = 0;
        delta[i * 4 + m] = sqrt(t / c.k_t);
    }
}

__global__ void plant_snapshot_kernel(long long n, const double* __restrict__ state, double* __restrict__ pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos[i * 3 + 0] = state[i * PL_NS + 3];
    pos[i * 3 + 1] = state[i * PL_NS + 4];
    pos[i * 3 + 2] = state[i * PL_NS + 5];
}

__device__ __forceinline__ void plant_rigid_body(const PlantCfg& c, const double (&x)[13], const double (&u)[6], double (&d)[13]) {
    const double ew = x[6], ex = x[7], ey = x[8], ez = x[9], p = x[10], q = x[11], r = x[12];
    d[0] = x[3]; d[1] = x[4]; d[2] = x[5];
    d[3] = u[0] / c.mass; d[4] = u[1] / c.mass; d[5] = u[2] / c.mass - c.gravity;
    d[6] = (0 - p * ex - q * ey - r * ez) / 2;
    d[7] = (p * ew + 0 + r * ey - q * ez) / 2;
    d[8] = (q * ew - r * ex + 0 + p * ez) / 2;
    d[9] = (r * ew + q * ex - p * ey + 0) / 2;
    d[10] = c.g1 * p * q - c.g2 * q * r + c.g3 * u[3] + c.g4 * u[5];
    d[11] = c.g5 * p * r - c.g6 * (p * p - r * r) + u[4] / c.iyy;
    d[12] = c.g7 * p * q - c.g1 * q * r + c.g4 * u[3] + c.g8 * u[5];
}

// QdDynamics.forward, in place on state [n][35]; delta_cmd [n][4]; pos = pre-step positions [n][3]
__global__ void __launch_bounds__(PL_THREADS) plant_step_kernel(const PlantCfg c, double* __restrict__ state, const double* __restrict__ delta_cmd,
                                                                const double* __restrict__ pos) {
    __shared__ double sS[PL_THREADS * PL_NS];
    __shared__ double sPos[PL_THREADS * 3];
    const long long base = (long long)blockIdx.x * PL_THREADS;
    const int t = threadIdx.x;
    const long long rows = (c.n - base < PL_THREADS) ? c.n - base : PL_THREADS;
    for (long long e = t; e < rows * PL_NS; e += PL_THREADS) sS[e] = state[base * PL_NS + e];
    __syncthreads();
    const long long i = base + t;
    const bool act = i < c.n;
    double* s = sS + t * PL_NS;
    double delta[4], f_i[3] = {0, 0, 0}, tt[4] = {0, 0, 0, 0}, R[9];
    if (act) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const double dc = delta_cmd[i * 4 + m];
            delta[m] = c.has_motor ? c.motor_alpha * s[31 + m] + (1 - c.motor_alpha) * dc : dc;
        }
        // rotor model: saturate (limits are float32 values, tools/saturate.py:24-30), thrust = k_t o^2, G_1 @ thrust
        double th[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const double d = (delta[m] <= c.o_min) ? c.o_min_sat : ((delta[m] >= c.o_max) ? c.o_max_sat : delta[m]);
            th[m] = c.k_t * (d * d);
        }
#pragma unroll
        for (int a = 0; a < 4; a++) tt[a] = c.G1[a * 4] * th[0] + c.G1[a * 4 + 1] * th[1] + c.G1[a * 4 + 2] * th[2] + c.G1[a * 4 + 3] * th[3];
        const double e0 = s[9], e1 = s[10], e2 = s[11], e3 = s[12];
        R[0] = e1 * e1 + e0 * e0 - e2 * e2 - e3 * e3; R[1] = 2.0 * (e1 * e2 - e3 * e0); R[2] = 2.0 * (e1 * e3 + e2 * e0);
        R[3] = 2.0 * (e1 * e2 + e3 * e0); R[4] = e2 * e2 + e0 * e0 - e1 * e1 - e3 * e3; R[5] = 2.0 * (e2 * e3 - e1 * e0);
        R[6] = 2.0 * (e1 * e3 - e2 * e0); R[7] = 2.0 * (e2 * e3 + e1 * e0); R[8] = e3 * e3 + e0 * e0 - e1 * e1 - e2 * e2;
        const double wn = s[28], we = s[29], wd = s[30];
        const double ur = s[16] - (R[0] * wn + R[3] * we + R[6] * wd);
        const double vr = s[17] - (R[1] * wn + R[4] * we + R[7] * wd);
        const double wr = s[18] - (R[2] * wn + R[5] * we + R[8] * wd);
        const double Va = sqrt(ur * ur + vr * vr + wr * wr);
        s[22] = Va;
        s[24] = ((ur == 0) ? c.half_pi_f32 : 0.0) + ((ur != 0) ? atan2(wr, ur) : 0.0);
        s[25] = ((Va != 0) ? 1.0 : 0.0) * asin(vr / Va);  // NaN when Va == 0, as in the reference
        const double fbx = -c.kd_x * ur, fby = -c.kd_y * vr, fbz = tt[0] + (-c.kd_z * wr + c.k_h * (ur * ur + vr * vr));
        f_i[0] = R[0] * fbx + R[1] * fby + R[2] * fbz;
        f_i[1] = R[3] * fbx + R[4] * fby + R[5] * fbz;
        f_i[2] = R[6] * fbx + R[7] * fby + R[8] * fbz;
    }
    if (c.has_downwash) {
        // sum over the other agents of the group of -k_d1 (rp / (4 dz))^2 exp(-0.5 (d_h / (k_d2 dz + k_d3))^2)
        // for 0 < dz < 4 m and d_h < 1.5 m (qd_dynamics.py:161-198); positions are the pre-step snapshot
        const long long g0 = (base / c.group) * c.group;                           // first group touched by this CTA
        long long g1 = ((base + rows - 1) / c.group + 1) * c.group;                 // end of the last one
        if (g1 > c.n) g1 = c.n;
        const long long my_g0 = act ? (i / c.group) * c.group : 0, my_g1 = my_g0 + c.group;
        const double px = act ? s[3] : 0, py = act ? s[4] : 0, pz = act ? s[5] : 0;
        double fd = 0;
        for (long long j0 = g0; j0 < g1; j0 += PL_THREADS) {
            __syncthreads();
            const long long cnt = (g1 - j0 < PL_THREADS) ? g1 - j0 : PL_THREADS;
            for (long long e = t; e < cnt * 3; e += PL_THREADS) sPos[e] = pos[j0 * 3 + e];
            __syncthreads();
            if (act) {
                for (int jj = 0; jj < cnt; jj++) {
                    const long long j = j0 + jj;
                    if (j < my_g0 || j >= my_g1) continue;
                    double dx = sPos[jj * 3] - px, dy = sPos[jj * 3 + 1] - py, dz = sPos[jj * 3 + 2] - pz;
                    if (dx > c.dw_h) dx = c.dw_h;
                    if (dy > c.dw_h) dy = c.dw_h;
                    const double dh = sqrt(dx * dx + dy * dy);
                    if (dh < c.dw_h && dz > 0 && dz < c.dw_v) {
                        const double a = c.rp / 4 / dz, b = dh / (c.k_d2 * dz + c.k_d3);
                        fd += -c.k_d1 * (a * a) * exp(-0.5 * (b * b));
                    }
                }
            }
        }
        f_i[2] += fd;
    }
    if (act) {
        const double u[6] = {f_i[0], f_i[1], f_i[2], tt[1], tt[2], tt[3]};
        double x[13] = {s[3], s[4], s[5], s[13], s[14], s[15], s[9], s[10], s[11], s[12], s[19], s[20], s[21]};
        double k1[13], k2[13], k3[13], k4[13], xt[13];
        const double dt = c.dt;
        plant_rigid_body(c, x, u, k1);
#pragma unroll
        for (int e = 0; e < 13; e++) xt[e] = x[e] + dt / 2.0 * k1[e];
        plant_rigid_body(c, xt, u, k2);
#pragma unroll
        for (int e = 0; e < 13; e++) xt[e] = x[e] + dt / 2.0 * k2[e];
        plant_rigid_body(c, xt, u, k3);
#pragma unroll
        for (int e = 0; e < 13; e++) xt[e] = x[e] + dt * k3[e];
        plant_rigid_body(c, xt, u, k4);
#pragma unroll
        for (int e = 0; e < 13; e++) x[e] += dt / 6 * (k1[e] + 2 * k2[e] + 2 * k3[e] + k4[e]);
        const double nq = sqrt(x[6] * x[6] + x[7] * x[7] + x[8] * x[8] + x[9] * x[9]);
        x[6] /= nq; x[7] /= nq; x[8] /= nq; x[9] /= nq;
        s[3] = x[0]; s[4] = x[1]; s[5] = x[2];
        s[13] = x[3]; s[14] = x[4]; s[15] = x[5];
        s[9] = x[6]; s[10] = x[7]; s[11] = x[8]; s[12] = x[9];
        s[19] = x[10]; s[20] = x[11]; s[21] = x[12];
        // _update_other_states: euler angles of the new quaternion, ground-speed angles with the OLD rotation
        const double e0 = x[6], e1 = x[7], e2 = x[8], e3 = x[9];
        s[6] = atan2(2.0 * (e0 * e1 + e2 * e3), e0 * e0 + e3 * e3 - e1 * e1 - e2 * e2);
        s[7] = asin(2.0 * (e0 * e2 - e1 * e3));
        s[8] = atan2(2.0 * (e0 * e3 + e1 * e2), e0 * e0 + e1 * e1 - e2 * e2 - e3 * e3);
        const double pdx = R[0] * s[16] + R[1] * s[17] + R[2] * s[18];
        const double pdy = R[3] * s[16] + R[4] * s[17] + R[5] * s[18];
        const double pdz = R[6] * s[16] + R[7] * s[17] + R[8] * s[18];
        const double vg = sqrt(pdx * pdx + pdy * pdy + pdz * pdz);
        s[23] = vg;
        s[26] = asin(pdz / vg);
        s[27] = atan2(pdy, pdx);
#pragma unroll
        for (int m = 0; m < 4; m++) s[31 + m] = delta[m];
    }
    __syncthreads();
    for (long long e = t; e < rows * PL_NS; e += PL_THREADS) state[base * PL_NS + e] = sS[e];
}

// pt_publisher.py:106-122 odom_2_nmpc_x: plant state -> NMPC x0 = (p, v, qw, qx, qy, qz), in the engine precision
template <typename T>
__global__ void plant_nmpc_x0_kernel(long long n, const double* __restrict__ state, T* __restrict__ x0) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* s = state + i * PL_NS;
    T* x = x0 + i * 10;
    x[0] = (T)s[3]; x[1] = (T)s[4]; x[2] = (T)s[5];
    x[3] = (T)s[13]; x[4] = (T)s[14]; x[5] = (T)s[15];
    x[6] = (T)s[9]; x[7] = (T)s[10]; x[8] = (T)s[11]; x[9] = (T)s[12];
}

// nmpc_node.py:273-283 nmpc_u_2_att_tgt: body rates = u0[0:3], thrust = c * mass / k_throttle (0 if k_throttle == 0)
template <typename T>
__global__ void plant_cmd_from_u0_kernel(long long n, const T* __restrict__ u0, double mass, double k_throttle, double* __restrict__ cmd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cmd[i * 4 + 0] = (double)u0[i * 4 + 0];
    cmd[i * 4 + 1] = (double)u0[i * 4 + 1];
    cmd[i * 4 + 2] = (double)u0[i * 4 + 2];
    cmd[i * 4 + 3] = (k_throttle != 0.0) ? (double)u0[i * 4 + 3] * mass / k_throttle : 0.0;
}

}  // namespace ndp
