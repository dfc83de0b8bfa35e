// Batched forms of the per-tick glue around the controller, so that a
// device-resident swarm / closed loop never leaves the GPU:
//   * PredXU (ndp_nmpc/msg/PredXU.msg:1-4): Float64MultiArray[] x (N+1 rows of 10), Float64MultiArray[] u (N rows of 4)
//     -- what do_pub_ref publishes (nmpc_node.py:116-133) and the follower / NDP leader consume
//     (nmpc_follower_node.py:57-74 adds the filtered formation offset to x[0:3]; ndp_nmpc_leader_node.py:60-76);
//   * hover-throttle Kalman filter (hv_throttle_est/hover_throttle_estimator.py:15-53 with the Tustin differentiator
//     differentiator.py:3-23), one thread per quadrotor, fp64 like the reference;
//   * nmpc_u_2_att_tgt with a per-quadrotor k_throttle (nmpc_node.py:273-283).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ndp {

// message payload of one quadrotor, flat float64: x rows first, then u rows -- (N+1)*10 + N*4 doubles (2 320 B at N=20)
__host__ __device__ inline long long predxu_len(int N) { return (long long)(N + 1) * 10 + (long long)N * 4; }

template <typename T>
__global__ void predxu_pack_kernel(long long B, int N, const T* __restrict__ xr, const T* __restrict__ ur, double* __restrict__ msg) {
    const long long per = predxu_len(N);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * per) return;
    const long long b = idx / per, r = idx - b * per;
    const long long nx = (long long)(N + 1) * 10;
    msg[idx] = (r < nx) ? (double)xr[b * nx + r] : (double)ur[b * (long long)N * 4 + (r - nx)];
}

// offset: [B][3] float64 formation offset added to the position of every x row (may be null)
template <typename T>
__global__ void predxu_unpack_kernel(long long B, int N, const double* __restrict__ msg, const double* __restrict__ offset, T* __restrict__ xr,
                                     T* __restrict__ ur) {
    const long long per = predxu_len(N);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * per) return;
    const long long b = idx / per, r = idx - b * per;
    const long long nx = (long long)(N + 1) * 10;
    double v = msg[idx];
    if (r < nx) {
        const int e = (int)(r % 10);
        if (offset && e < 3) v += offset[b * 3 + e];
        xr[b * nx + r] = (T)v;
    } else {
        ur[b * (long long)N * 4 + (r - nx)] = (T)v;
    }
}

// estimator state per quadrotor, 8 doubles: vz[k-1], az[k-1] (differentiator), x = (f_collect, k_throttle), P (row-major 2x2)
constexpr int HTE_NS = 8;

__global__ void hover_throttle_init_kernel(long long n, double k_init, double* __restrict__ est, double* __restrict__ k_throttle) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* e = est + i * HTE_NS;
    e[0] = 0.0; e[1] = 0.0; e[2] = 0.0; e[3] = k_init;
    e[4] = 1.0; e[5] = 0.0; e[6] = 0.0; e[7] = 1.0;
    if (k_throttle) k_throttle[i] = k_init;
}

struct HoverThrottleCfg {
    double a1, a2;       // Tustin differentiator, tau = 0.05
    double mass, gravity, q0, q1, r;
};

__global__ void hover_throttle_update_kernel(long long n, HoverThrottleCfg c, const double* __restrict__ vz, long long vz_ld,
                                             const double* __restrict__ thr, long long thr_ld, double* __restrict__ est,
                                             double* __restrict__ k_throttle) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* e = est + i * HTE_NS;
    const double v = vz[i * vz_ld], th = thr[i * thr_ld];
    const double az = c.a1 * e[1] + c.a2 * (v - e[0]);
    e[0] = v; e[1] = az;
    if (th > 0.1 && th < 1.0) {
        const double z = az + c.gravity, im = 1.0 / c.mass;
        const double p11 = e[7];
        // P^- = Phi P Phi' + Q, Phi = [[0, thr], [0, 1]]
        const double m00 = th * th * p11 + c.q0, m01 = th * p11, m11 = p11 + c.q1;
        const double s = m00 * im * im + c.r;
        const double k0 = m00 * im / s, k1 = m01 * im / s;
        const double xp0 = th * e[3], xp1 = e[3];
        const double innov = z - xp0 * im;
        e[2] = xp0 + k0 * innov;
        e[3] = xp1 + k1 * innov;
        e[4] = (1.0 - k0 * im) * m00;
        e[5] = (1.0 - k0 * im) * m01;
        e[6] = m01 - k1 * im * m00;
        e[7] = m11 - k1 * im * m01;
    }
    if (k_throttle) k_throttle[i] = e[3];
}

template <typename T>
__global__ void cmd_from_u0_dev_kernel(long long n, const T* __restrict__ u0, double mass, const double* __restrict__ k_throttle,
                                       double* __restrict__ cmd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double k = k_throttle[i];
    cmd[i * 4 + 0] = (double)u0[i * 4 + 0];
    cmd[i * 4 + 1] = (double)u0[i * 4 + 1];
    cmd[i * 4 + 2] = (double)u0[i * 4 + 2];
    cmd[i * 4 + 3] = (k != 0.0) ? (double)u0[i * 4 + 3] * mass / k : 0.0;
}

}  // namespace ndp

// ---- device-resident sliding reference list (NMPCRefPublisher, pt_pub/pt_publisher.py:57-103) ----
// The reference keeps `long_list_size` = 101 points (x, u) at ts_nmpc = 0.02 s per quadrotor and, every control tick,
// pops the front, appends ONE new point (the trajectory at t + T_horizon) and hands the controller every 5th point
// (xr = list[0::5], 21 nodes; ur = the same without its last entry; params/nmpc_params.py:40-43).  Here the lists live
// on the device as rings -- list[i] = ring[(head + i) % len] -- so a tick uploads one point per quadrotor instead of the
// whole horizon; `other` is the neighbour's list (the 6 position / velocity columns DownwashNN reads), fed by the
// neighbour's own new point.  One thread per output element; node N takes the new point straight from the input.
namespace ndp {

template <typename T>
__global__ void longlist_push_kernel(long long B, int N, int stride, int len, int head, const T* __restrict__ new_x, const T* __restrict__ new_u,
                                     const T* __restrict__ new_o, T* __restrict__ ring_x, T* __restrict__ ring_u, T* __restrict__ ring_o,
                                     T* __restrict__ xr, T* __restrict__ ur, T* __restrict__ other) {
    const int nx = (N + 1) * 10, nu = N * 4, no = ring_o ? (N + 1) * 6 : 0;
    const int per = nx + nu + no;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * per) return;
    const long long b = idx / per;
    int r = (int)(idx - b * per);
    const int tail = (head + len - 1) % len;  // where the appended point goes (head is the list's front AFTER the pop)
    if (r < nx) {
        const int k = r / 10, e = r - k * 10;
        const int slot = (head + stride * k) % len;
        // a node that falls on the list's last entry reads the point appended by this very tick
        xr[b * nx + r] = (stride * k == len - 1) ? new_x[b * 10 + e] : ring_x[(b * len + slot) * 10 + e];
        if (k == 0) ring_x[(b * len + tail) * 10 + e] = new_x[b * 10 + e];  // tail = the slot the pop just freed: nobody reads it
    } else if (r < nx + nu) {
        r -= nx;
        const int k = r / 4, e = r - k * 4;
        ur[b * nu + r] = ring_u[(b * len + (head + stride * k) % len) * 4 + e];
        if (k == 0) ring_u[(b * len + tail) * 4 + e] = new_u[b * 4 + e];  // the new u point is not part of this tick's ur
    } else {
        r -= nx + nu;
        const int k = r / 6, e = r - k * 6;
        const int slot = (head + stride * k) % len;
        other[b * no + r] = (stride * k == len - 1) ? new_o[b * 6 + e] : ring_o[(b * len + slot) * 6 + e];
        if (k == 0) ring_o[(b * len + tail) * 6 + e] = new_o[b * 6 + e];
    }
}

}  // namespace ndp
