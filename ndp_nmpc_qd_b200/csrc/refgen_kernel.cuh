// Batched NMPC reference generation: xr[B][N+1][10], ur[B][N][4] straight on the
// device, from the TrajCoefficients of the reference's planner.
// Restates ndp_nmpc/scripts/pt_pub/: piecewise-polynomial evaluation in normalised segment time with the
// hover branch after the end (base_pt_publisher.py:81-148), differential flatness (pt_publisher.py:188-248,
// quaternion by the ROS tf quaternion_from_matrix algorithm, w >= 0 for the attitudes of interest) and the
// packing x = (p, v, qw, qx, qy, qz), u = (wx, wy, wz, collective force / mass) (pt_publisher.py:124-147).
// One thread per (problem, node); node k is evaluated at t_b + k th_pred (the 101-point sliding list of
// pt_publisher.py:57-103 sampled every 5th point, without its wall-clock jitter).  Arithmetic in float64.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace ndp {

struct RefGenTable {
    int n_traj;
    const int* seg_off;      // [n_traj + 1] first segment of each trajectory
    const double* t_cum;     // [total_seg + n_traj]: per trajectory n_seg + 1 knot times (offset seg_off[j] + j)
    const double* cxyz;      // [total_seg][3][8]
    const double* cyaw;      // [total_seg][4]
    const double* final_pt;  // [n_traj][3]
};

template <int kOrd>
__device__ __forceinline__ void poly_derivs(const double* __restrict__ c, double s, double inv_ts, double (&d)[4], int n_deriv) {
    // d[k] = k-th real-time derivative of sum_j c_j s^j, k < n_deriv
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < n_deriv) {
            double acc = 0.0;
#pragma unroll
            for (int j = kOrd; j >= k; j--) {
                double f = 1.0;
#pragma unroll
                for (int q = 0; q < k; q++) f *= (double)(j - q);
                acc = acc * s + c[j] * f;
            }
            double sc = 1.0;
            for (int q = 0; q < k; q++) sc *= inv_ts;
            d[k] = acc * sc;
        }
    }
}

template <typename T>
__global__ void refgen_horizon_kernel(const RefGenTable tb, long long B, const int* __restrict__ traj_id, const double* __restrict__ t0, int N,
                                      double th_pred, const double* __restrict__ offset, double mass, double gravity, T* __restrict__ xr,
                                      T* __restrict__ ur) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * (N + 1)) return;
    const long long b = idx / (N + 1);
    const int k = (int)(idx - b * (N + 1));
    // an out-of-range trajectory id is clamped (the host wrapper validates ids it can see; a bad id produced on the
    // device must not read outside the tables)
    int tj = traj_id ? traj_id[b] : 0;
    tj = tj < 0 ? 0 : (tj >= tb.n_traj ? tb.n_traj - 1 : tj);
    const int s0 = tb.seg_off[tj], n_seg = tb.seg_off[tj + 1] - s0;
    const double* tc = tb.t_cum + s0 + tj;
    // the reference never evaluates before its start time (t = ros_t - start_ros_t >= 0, base_pt_publisher.py:91):
    // clamp instead of extrapolating segment 0 backwards
    const double t = fmax(t0[b] + k * th_pred, tc[0]);
    double pos[3], vel[3] = {0, 0, 0}, acc[3] = {0, 0, 0}, jerk[3] = {0, 0, 0}, yaw = 0.0, yaw_dot = 0.0;
    if (t >= tc[n_seg]) {  // finished: hover at the final point (base_pt_publisher.py:93-96)
#pragma unroll
        for (int a = 0; a < 3; a++) pos[a] = tb.final_pt[tj * 3 + a];
    } else {
        int i = 0;
        while (i + 1 < n_seg && !(tc[i + 1] > t)) i++;  // first knot strictly greater than t, minus one
        const double ts = tc[i + 1] - tc[i], s = (t - tc[i]) / ts, inv_ts = 1.0 / ts;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double d[4];
            poly_derivs<7>(tb.cxyz + ((long long)(s0 + i) * 3 + a) * 8, s, inv_ts, d, 4);
            pos[a] = d[0]; vel[a] = d[1]; acc[a] = d[2]; jerk[a] = d[3];
        }
        double d[4];
        poly_derivs<3>(tb.cyaw + (long long)(s0 + i) * 4, s, inv_ts, d, 2);
        yaw = d[0]; yaw_dot = d[1];
    }
    // differential flatness
    const double tx = acc[0] + 0.0, ty = acc[1] + 0.0, tz = acc[2] + gravity;
    // The reference raises ValueError for a free-fall point (|t_des| = 0) and for z_b parallel to x_c
    // (pt_publisher.py:203-204,214-215).  A kernel cannot raise: those points fall back to the hover attitude
    // (z_b = e_z) resp. to y_b = e_y x-rotated by yaw, and the thrust direction stays finite instead of NaN.
    double tn = sqrt(tx * tx + ty * ty + tz * tz);
    const bool free_fall = !(tn > 1e-9);
    const double zb[3] = {free_fall ? 0.0 : tx / tn, free_fall ? 0.0 : ty / tn, free_fall ? 1.0 : tz / tn};
    if (free_fall) tn = 0.0;
    const double u1 = tn * mass;
    const double xc[3] = {cos(yaw), sin(yaw), 0.0};
    double yb[3] = {zb[1] * xc[2] - zb[2] * xc[1], zb[2] * xc[0] - zb[0] * xc[2], zb[0] * xc[1] - zb[1] * xc[0]};
    double yn = sqrt(yb[0] * yb[0] + yb[1] * yb[1] + yb[2] * yb[2]);
    if (!(yn > 1e-9)) { yb[0] = -sin(yaw); yb[1] = cos(yaw); yb[2] = 0.0; yn = 1.0; }
    yb[0] /= yn; yb[1] /= yn; yb[2] /= yn;
    const double xb[3] = {yb[1] * zb[2] - yb[2] * zb[1], yb[2] * zb[0] - yb[0] * zb[2], yb[0] * zb[1] - yb[1] * zb[0]};
    const double zj = zb[0] * jerk[0] + zb[1] * jerk[1] + zb[2] * jerk[2], mu = free_fall ? 0.0 : mass / u1;
    const double ho[3] = {mu * (jerk[0] - zj * zb[0]), mu * (jerk[1] - zj * zb[1]), mu * (jerk[2] - zj * zb[2])};
    const double wp = -(ho[0] * yb[0] + ho[1] * yb[1] + ho[2] * yb[2]);
    const double wq = ho[0] * xb[0] + ho[1] * xb[1] + ho[2] * xb[2];
    const double wr = yaw_dot * zb[2];
    // quaternion_from_matrix on R = [x_b y_b z_b] (columns); M(r, c): c = 0 -> x_b, 1 -> y_b, 2 -> z_b
    const double M[3][3] = {{xb[0], yb[0], zb[0]}, {xb[1], yb[1], zb[1]}, {xb[2], yb[2], zb[2]}};
    double q[4];
    double tr = M[0][0] + M[1][1] + M[2][2] + 1.0;
    if (tr > 1.0) {
        q[3] = tr; q[2] = M[1][0] - M[0][1]; q[1] = M[0][2] - M[2][0]; q[0] = M[2][1] - M[1][2];
    } else {
        int i = 0, j = 1, kk = 2;
        if (M[1][1] > M[0][0]) { i = 1; j = 2; kk = 0; }
        if (M[2][2] > M[i][i]) { i = 2; j = 0; kk = 1; }
        tr = M[i][i] - (M[j][j] + M[kk][kk]) + 1.0;
        q[i] = tr; q[j] = M[i][j] + M[j][i]; q[kk] = M[kk][i] + M[i][kk]; q[3] = M[kk][j] - M[j][kk];
    }
    const double qs = 0.5 / sqrt(tr);
    T* x = xr + idx * 10;
    const double ox = offset ? offset[b * 3] : 0.0, oy = offset ? offset[b * 3 + 1] : 0.0, oz = offset ? offset[b * 3 + 2] : 0.0;
    x[0] = (T)(pos[0] + ox); x[1] = (T)(pos[1] + oy); x[2] = (T)(pos[2] + oz);
    x[3] = (T)vel[0]; x[4] = (T)vel[1]; x[5] = (T)vel[2];
    x[6] = (T)(q[3] * qs); x[7] = (T)(q[0] * qs); x[8] = (T)(q[1] * qs); x[9] = (T)(q[2] * qs);
    if (k < N) {
        T* u = ur + (b * N + k) * 4;
        u[0] = (T)wp; u[1] = (T)wq; u[2] = (T)wr; u[3] = (T)(u1 / mass);
    }
}

}  // namespace ndp
