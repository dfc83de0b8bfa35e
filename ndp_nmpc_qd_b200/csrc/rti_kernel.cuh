// Fused SQP-RTI step kernel for the quadrotor body-rate OCP (sm_100a).
//
// One 16-lane group (half a warp) owns one problem.  Lane j < 14 owns COLUMN j of every
// 14-wide stage matrix ([A_k B_k], P+[A B], the 14x14 stage Hessian), lane 14 owns the
// affine "column" (b_k -> stage gradient), lane 15 idles.  With that mapping
//   * the RK4 forward-sensitivity column j of interval k is integrated by lane j and lands
//     in the registers where the Riccati step needs it (reference: acados ERK + sens_forw,
//     integrator_type="ERK" nmpc_body_rate_ctl.py:76; dynamics ndp_nmpc_body_rate_ctl.py:151-162),
//   * P+ [A B], [A B]' P+ [A B], the 4x4 Cholesky solve and the Schur complement are all
//     column-local; cross-lane traffic is broadcast reads of small shared-memory tiles plus
//     ten shuffles for the 4x4 input block,
//   * the gradient recursion rides along as lane 14 with the same instruction stream.
// Two launches per solve.  rti_step_kernel (nominal, 128 registers, 8 CTAs of 4 problems per SM) linearises and runs the
// unconstrained Riccati sweep; if the step satisfies all box bounds it IS the QP solution and is accepted on the fly in
// the forward sweep.  A step that leaves its box is handed to rti_constrained_kernel (programmatic dependent launch, a
// device queue of problem indices + the violated bounds as the first active set), which solves the QP exactly with
// primal-dual active-set rounds on the same Riccati recursion: pinned inputs are eliminated from the 4x4 input block,
// pinned velocity components of x_{k+1} are resolved as a stage-k mixed constraint in the range space of the inputs
// (backward_stage, kBar == 2), multipliers decide releases, a hash history detects cycling (damped single releases
// then), rounds restart the backward sweep at the highest stage whose pins changed.  A Mehrotra predictor-corrector
// IPM on the same kernel (HPIPM's algorithm, qp_solver="PARTIAL_CONDENSING_HPIPM" with cond_N = N,
// nmpc_body_rate_ctl.py:71-79) is kept as the fallback that proposes an active set when the rounds do not settle.
// Cost: NONLINEAR_LS Gauss-Newton blocks in closed form (nmpc_body_rate_ctl.py:48-53,163-180), bounds :56-61, iterate
// protocol :86-112 (global memory keeps the old iterate until a step is accepted).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace ndp {

constexpr int NX = 10, NU = 4, NZ = 14, GL = 16;
constexpr int NYS = 14;  // yref stride per stage (global)
constexpr int SYS = 16;  // shared-memory stride of the per-stage cost record (residuals, see cost_records)
constexpr int NPS = 8;   // parameter stride per stage
constexpr int TLD = 12;  // leading dimension of the [A B b] tiles and of the forward-sweep records
constexpr int PTRI = 56;  // 55 elements of the packed lower triangle of a 10 x 10 symmetric matrix, padded to a multiple of 4
constexpr int PSAVE = PTRI + 12;  // (P | p) as saved for partial sweeps
constexpr int RTI_THREADS = 128;
constexpr int RTI_PPC = RTI_THREADS / GL;  // problems per CTA (upper bound)

template <typename T>
struct RtiCfg {
    int N, ipm_max_iter, polish_max, as_first_max;
    T h, inv_mass, g;
    T Q[10], R[4], hQ[10], hR[4], umin[4], umax[4], vmin[3], vmax[3];  // hQ = h Q, hR = h R (stage cost scaled by the interval)
    T tol_mu, tol_res, mu0, t_floor, t_min;
};

template <typename T>
struct RtiArgs {
    const T* x0;    // [B][10]
    const T* yref;  // [B][N+1][14]
    const T* par;   // [B][N+1][8]   q_r(4) f(3) pad
    T* X;           // [B][N+1][10]  iterate, in/out
    T* U;           // [B][N][4]
    T* u0;          // [B][4] or null
    int32_t* status;  // [B]
    int32_t* status2; // optional second destination of the status (a caller's output record), or null
    int32_t* stats;   // [B][4]
    // optional fused controller.update(): when xr != null the kernel builds yref / p from
    // (xr[B][N+1][10], ur[B][N][4], f[B][N+1][3] or null) itself and stores them in yref_w / par_w
    const T* xr;
    const T* ur;
    const T* f;
    T* yref_w;
    T* par_w;
    // [B][AS_OWNERS][4]: active set a problem starts its constrained solve from (lo.w0 lo.w1 hi.w0 hi.w1 per owning lane:
    // the three velocity components, then the four inputs) -- written by the nominal kernel for the problems it hands
    // over (the bounds the unconstrained step violates) and, with as_warm, left by the constrained kernel as the first
    // guess of the problem's next solve
    unsigned long long* as_store;
    int as_warm;
    T* ws;            // nominal kernel: [slots][ws_stride] forward-sweep records; constrained kernel: its full workspace
    long long ws_stride;
    // hand-over from the nominal to the constrained kernel: queue [B] of problem indices (bit 30: the unconstrained
    // sweep has run) filled from both ends (front: the problems expected to take longest), qctl = {count of the back
    // part, head, done, count of the front part}
    int* queue;
    int* qctl;
    int B;
    // constrained kernel only: the nominal kernel's record workspace (its slot s holds the stage tiles of problem s when
    // no slot was reused, i.e. ws_n_slots >= B)
    T* ws_n;
    long long ws_n_stride;
    int ws_n_slots;
};

constexpr int AS_OWNERS = 7;
__device__ __forceinline__ int as_owner(int lane) { return (lane >= 10) ? lane - 7 : lane - 3; }  // lanes 3..5 -> 0..2, 10..13 -> 3..6

__host__ __device__ constexpr int al4(int o) { return (o + 3) & ~3; }

// ---- per-problem shared memory layout (elements of T; every region 4-element aligned) ----
struct SmemLayout {
    int oX, oU, oPar, oY, oDz, oP, op, oT0, oT1, oHux, oIpm, total;  // oY and oDz are adjacent: see forward_sweep
    // nominal: the layout of rti_step_kernel, which never forms a QP step array (sDz, (N+1) x 16 elements -- 5 KB of the
    // 19 KB per problem at N = 80): only the constrained kernel's layout carries it
    __host__ __device__ constexpr explicit SmemLayout(int N, bool nominal = false)
        : oX(0), oU(al4((N + 1) * NX)), oPar(oU + N * NU), oY(oPar + (N + 1) * NPS), oDz(oY + (N + 1) * SYS),
          // QP step [k][lane]; the FW_RING x (14 x TLD) ring of the accepted forward sweep runs from oY over sDz, P+,
          // p+ and the tiles: oY .. oHux must span at least FW_RING * 14 * TLD elements (static_assert below)
          oP(oY + (((N + 1) * (nominal ? SYS : SYS + 16) > 700) ? (N + 1) * (nominal ? SYS : SYS + 16) : 700)),
          op(oP + PTRI),           // P+ as its packed lower triangle (P is symmetric): element (a, b), a >= b, at a (a + 1) / 2 + b; then p+
          oT0(op + 12),            // tile of stage k:   rows r = 0..9 of [A B](:,6..13) | b | pad3, stride TLD
          oT1(oT0 + 10 * TLD),     // tile of stage k-1 (the integrator fills two stages per pass)
          oHux(oT1 + 10 * TLD),    // Hux transposed [i][m]
          // constrained kernel only: the interior-point vectors TL TU LL LU CL CU, each [k][8] (the seven box-owning lanes),
          // then the barrier diagonal and gradient of the sweeps, [k <= N][8] each:
          // their element-wise updates are dependent load - compute - store chains over the stages, ~50 us per
          // iteration out of the L2-resident workspace, a few us from here
          oIpm(oHux + 10 * 4),
          // problem stride = 16 (mod 32) words: the two problems of a warp then sit on disjoint bank halves
          total(((oIpm + (nominal ? 0 : 6 * N * 8 + 2 * (N + 1) * 8) + 15) & ~31) + 16) {}
};

// ---- per-slot global workspace layout (elements of T) ----
struct WsLayout {
    long long oRec, oBarD, oBarG, oIpm, oZc, oHrow, oTv, oPs, total;
    __host__ __device__ constexpr explicit WsLayout(int N)
        : oRec(0),                                   // [k][14][TLD]: rows 0..9 = tile rows (x-lane records of the
                                                     // forward sweep), rows 10..13 = [K(m, 0..9) kappa_m pad]
          oBarD(oRec + (long long)N * 14 * TLD), oBarG(oBarD + (long long)(N + 1) * 16), oIpm(oBarG + (long long)(N + 1) * 16),
          oZc(oIpm + (long long)7 * N * 16),  // LL LU TL TU CL CU ACT, each [k][lane]
          oHrow(oZc + (long long)(N + 1) * 16),  // [k][m][16]: row m of [Hux Guu], [14] = gradient
          oTv(oHrow + (long long)N * 4 * 16),    // [k][a][12]: multiplier row of pinned velocity component a of stage k+1:
                                                 // nu = -(T(a, 0..9) . dx_k + T(a, 10))
          oPs(oTv + (long long)N * 3 * 12),      // [k][PSAVE]: (P_k | p_k) as left in shared memory by backward stage k -- lets an
                                                 // active-set round restart its backward sweep at the highest stage that changed
          total(oPs + (long long)N * PSAVE) {}
};

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void ld(const float* p, float& a, float& b, float& c, float& d) {
        float4 v = *reinterpret_cast<const float4*>(p);
        a = v.x; b = v.y; c = v.z; d = v.w;
    }
    static __device__ __forceinline__ void st(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
};
template <> struct Vec4<double> {
    static __device__ __forceinline__ void ld(const double* p, double& a, double& b, double& c, double& d) {
        double2 v = *reinterpret_cast<const double2*>(p);
        double2 w = *reinterpret_cast<const double2*>(p + 2);
        a = v.x; b = v.y; c = w.x; d = w.y;
    }
    static __device__ __forceinline__ void st(double* p, double a, double b, double c, double d) {
        *reinterpret_cast<double2*>(p) = make_double2(a, b);
        *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
    }
};

// reciprocal square root: MUFU.RSQ + one Newton step (fp32, ~1 ulp) / IEEE (fp64)
template <typename T> __device__ __forceinline__ T trsqrt(T x);
template <> __device__ __forceinline__ float trsqrt<float>(float x) {
    // the bare MUFU.RSQ: rsqrtf() wraps it in a denormal-range fix-up (two compares / selects / scalings per call) that the
    // pivots of a Riccati stage (O(1) .. O(1e4), tested for > 0 separately) never need
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * (1.5f - 0.5f * x * y * y);
}
template <> __device__ __forceinline__ double trsqrt<double>(double x) { return 1.0 / sqrt(x); }

// Division inside the interior-point vector updates: fp32 takes the two-instruction approximate form (2 ulp; the IEEE
// sequence is ~10 instructions and these loops are a lone problem's serial tail); the iteration only has to deliver an
// active-set estimate to the exact rounds.  fp64 divides exactly.
template <typename T> __device__ __forceinline__ T tdiv(T a, T b) { return a / b; }
template <> __device__ __forceinline__ float tdiv<float>(float a, float b) { return __fdividef(a, b); }

// Packed fp32 pairs (sm_100 FFMA2): (d0, d1) += (a0, a1) * (b0, b1) in ONE issue slot.  The nominal kernel is bound by
// issue slots (~70 % of them used on every scheduler, FMA pipe under 40 %), so its long dot products run on pairs with
// two partial sums each; fp64 keeps scalar FMAs.
template <typename T>
__device__ __forceinline__ void fma2(T& d0, T& d1, T a0, T a1, T b0, T b1) {
    d0 += a0 * b0;
    d1 += a1 * b1;
}
template <>
__device__ __forceinline__ void fma2<float>(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// (d0, d1) = (a0, a1) * (b0, b1) + (c0, c1), the accumulator input separate from the result (no copy of c needed)
template <typename T>
__device__ __forceinline__ void fma2o(T& d0, T& d1, T a0, T a1, T b0, T b1, T c0, T c1) {
    d0 = a0 * b0 + c0;
    d1 = a1 * b1 + c1;
}
template <>
__device__ __forceinline__ void fma2o<float>(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(d0), "=f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}

template <typename T>
__device__ __forceinline__ T grp_sum(T v, unsigned mask) {
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(mask, v, o, GL);
    return v;
}
// Group votes.  The nominal kernel runs its two problems per warp in lockstep and passes the constant full mask (no
// MATCH / REDUX / BRA.DIV convergence check in front of every shuffle and __syncwarp); the vote is then taken from the
// ballot bits of this thread's own half.  The constrained kernel's halves diverge and pass their half mask.
__device__ __forceinline__ bool grp_any(unsigned mask, bool p) {
    if (mask == 0xffffffffu) return ((__ballot_sync(0xffffffffu, p) >> (threadIdx.x & 16)) & 0xffffu) != 0u;
    return __any_sync(mask, p) != 0;
}
__device__ __forceinline__ bool grp_all(unsigned mask, bool p) {
    if (mask == 0xffffffffu) return ((__ballot_sync(0xffffffffu, p) >> (threadIdx.x & 16)) & 0xffffu) == 0xffffu;
    return __all_sync(mask, p) != 0;
}
template <typename T>
__device__ __forceinline__ T grp_min(T v, unsigned mask) {
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        T w = __shfl_xor_sync(mask, v, o, GL);
        v = w < v ? w : v;
    }
    return v;
}

// One RK4 step of the state (every lane, redundantly) and of this lane's sensitivity column.
// x[10], u[4] in shared memory; fm = f / mass.  Column j: initial e_j for j < 10, forcing
// B_c[:, j-10] for j = 10..13, nothing for j >= 14.
// The state and its sensitivity column are carried as PAIRS (x_i, s_i): wherever both see the same coefficient -- the
// stage values x0 + a k, the quaternion rows (linear in (q, s) with the body rates as coefficients) and the weighted sum
// of the slopes -- one packed FFMA2 does both (the nominal kernel is bound by issue slots: 29 of ~90 instructions per
// RK stage saved).
template <typename T>
__device__ __forceinline__ void rk4_column(const RtiCfg<T>& c, int j, const T* __restrict__ x, const T* __restrict__ u,
                                           T fm0, T fm1, T fm2, T (&xa)[10], T (&sa)[10]) {
    const T h = c.h;
    const T hwx = T(0.5) * u[0], hwy = T(0.5) * u[1], hwz = T(0.5) * u[2], cc = u[3];
    const T nwx = -hwx, nwy = -hwy, nwz = -hwz;
    const T ew0 = (j == 10) ? T(0.5) : T(0), ew1 = (j == 11) ? T(0.5) : T(0), ew2 = (j == 12) ? T(0.5) : T(0);
    const T ec = (j == 13) ? T(1) : T(0);
    const T c2 = T(2) * cc, c4n = T(-4) * cc, fg2 = fm2 - c.g;
    T x0[10], s0[10], kx[10], ks[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        x0[i] = x[i];
        s0[i] = (j == i) ? T(1) : T(0);
        xa[i] = x0[i];
        sa[i] = s0[i];
        kx[i] = T(0);
        ks[i] = T(0);
    }
#pragma unroll
    for (int st = 0; st < 4; st++) {
        const T a = (st == 3) ? h : h * T(0.5);
        // stage values of (v, q) and of their sensitivities
        T zx[10], zs[10];
#pragma unroll
        for (int i = 3; i < 10; i++) {
            if (st == 0) { zx[i] = x0[i]; zs[i] = s0[i]; }
            else fma2o<T>(zx[i], zs[i], a, a, kx[i], ks[i], x0[i], s0[i]);
        }
        const T qw = zx[6], qx = zx[7], qy = zx[8], qz = zx[9];
        const T s6 = zs[6], s7 = zs[7], s8 = zs[8], s9 = zs[9];
        const T r13 = T(2) * (qx * qz + qw * qy), r23 = T(2) * (qy * qz - qw * qx);
        const T r33 = T(1) - T(2) * qx * qx - T(2) * qy * qy;
        kx[0] = zx[3]; kx[1] = zx[4]; kx[2] = zx[5];
        ks[0] = zs[3]; ks[1] = zs[4]; ks[2] = zs[5];
        kx[3] = r13 * cc + fm0;
        kx[4] = r23 * cc + fm1;
        kx[5] = r33 * cc + fg2;
        ks[3] = c2 * (qy * s6 + qz * s7 + qw * s8 + qx * s9) + ec * r13;
        ks[4] = c2 * (-qx * s6 - qw * s7 + qz * s8 + qy * s9) + ec * r23;
        ks[5] = c4n * (qx * s7 + qy * s8) + ec * r33;
        // quaternion rows: (kx, ks) = M(w / 2) (q, s) + (0, M(e / 2) q)
        const T e6 = -qx * ew0 - qy * ew1 - qz * ew2;
        const T e7 = qw * ew0 - qz * ew1 + qy * ew2;
        const T e8 = qz * ew0 + qw * ew1 - qx * ew2;
        const T e9 = -qy * ew0 + qx * ew1 + qw * ew2;
        fma2o<T>(kx[6], ks[6], nwx, nwx, qx, s7, T(0), e6);
        fma2<T>(kx[6], ks[6], nwy, nwy, qy, s8);
        fma2<T>(kx[6], ks[6], nwz, nwz, qz, s9);
        fma2o<T>(kx[7], ks[7], hwx, hwx, qw, s6, T(0), e7);
        fma2<T>(kx[7], ks[7], hwz, hwz, qy, s8);
        fma2<T>(kx[7], ks[7], nwy, nwy, qz, s9);
        fma2o<T>(kx[8], ks[8], hwy, hwy, qw, s6, T(0), e8);
        fma2<T>(kx[8], ks[8], nwz, nwz, qx, s7);
        fma2<T>(kx[8], ks[8], hwx, hwx, qz, s9);
        fma2o<T>(kx[9], ks[9], hwz, hwz, qw, s6, T(0), e9);
        fma2<T>(kx[9], ks[9], hwy, hwy, qx, s7);
        fma2<T>(kx[9], ks[9], nwx, nwx, qy, s8);
        const T bw = (st == 0 || st == 3) ? h * T(1.0 / 6.0) : h * T(1.0 / 3.0);
#pragma unroll
        for (int i = 0; i < 10; i++) fma2<T>(xa[i], sa[i], bw, bw, kx[i], ks[i]);
    }
}

// Per-stage cost record (shared memory, stride SYS), built once per problem from the iterate and yref/p:
//   [0..5] x - yref (p, v)   [6] 0   [7..9] q_r(xyz) - yref_q(xyz)   [10..13] u - yref_u   [14..15] 0
// so that the gradient lane reads its stage residual with four 128-bit loads.
template <typename T>
__device__ __forceinline__ void cost_records(int N, int lane, T* __restrict__ sY, const T* __restrict__ sX, const T* __restrict__ sU,
                                             const T* __restrict__ sPar) {
    // SYS == GL: lane e owns element e of every stage record -- its source and stride are fixed, no branches
    static_assert(SYS == GL, "cost_records: one lane per record element");
    const T* src = (lane < 6) ? sX + lane : ((lane < 10) ? sPar + (lane - 6) : sU + ((lane - 10) & 3));
    const int str = (lane < 6) ? NX : ((lane < 10) ? NPS : NU);
    const bool isu = lane >= 10, has = (lane != 6) && (lane < 14);
#pragma unroll 3
    for (int k = 0; k <= N; k++) {
        const bool last_u = isu && (k == N);
        const T v = src[(last_u ? N - 1 : k) * str] - sY[k * SYS + lane];
        sY[k * SYS + lane] = (has && !last_u) ? v : T(0);
    }
}

// Gauss-Newton cost blocks in column form: lane j < 14 gets column j of the stage Hessian,
// lane 14 the gradient (yref quaternion part may differ from q_r).
template <typename T>
__device__ __forceinline__ void add_cost(T (&H)[14], const RtiCfg<T>& c, int j, int k, bool terminal, const T* __restrict__ sX,
                                         const T* __restrict__ sR, const T* __restrict__ sPar) {
    const T* wq = terminal ? c.Q : c.hQ;
    const T* xk = sX + k * NX;
    const bool g = (j == 14);
    T r[16];
    Vec4<T>::ld(sR + k * SYS, r[0], r[1], r[2], r[3]);
    Vec4<T>::ld(sR + k * SYS + 4, r[4], r[5], r[6], r[7]);
    Vec4<T>::ld(sR + k * SYS + 8, r[8], r[9], r[10], r[11]);
    Vec4<T>::ld(sR + k * SYS + 12, r[12], r[13], r[14], r[15]);
    T w, x, y, z;
    Vec4<T>::ld(sPar + k * NPS, w, x, y, z);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const T yi = g ? r[i] : ((j == i) ? T(1) : T(0));
        H[i] += wq[i] * yi;
    }
    T yq[4];
#pragma unroll
    for (int n = 0; n < 4; n++) yq[n] = g ? xk[6 + n] : ((j == 6 + n) ? T(1) : T(0));
    const T d1 = g ? r[7] : T(0), d2 = g ? r[8] : T(0), d3 = g ? r[9] : T(0);
    const T v1 = wq[7] * (-x * yq[0] + w * yq[1] - z * yq[2] + y * yq[3] + d1);
    const T v2 = wq[8] * (-y * yq[0] + z * yq[1] + w * yq[2] - x * yq[3] + d2);
    const T v3 = wq[9] * (-z * yq[0] - y * yq[1] + x * yq[2] + w * yq[3] + d3);
    H[6] += -x * v1 - y * v2 - z * v3;
    H[7] += w * v1 + z * v2 - y * v3;
    H[8] += -z * v1 + w * v2 + x * v3;
    H[9] += y * v1 - x * v2 + w * v3;
    if (!terminal) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const T yu = g ? r[10 + m] : ((j == 10 + m) ? T(1) : T(0));
            H[10 + m] += c.hR[m] * yu;
        }
    }
}

// Terminal stage: P_N = W_e, p_N = gradient (no bounds at the terminal node: acados lbx/ubx
// apply to intermediate nodes only).
template <typename T>
__device__ __forceinline__ void backward_terminal(const RtiCfg<T>& c, int N, int j, unsigned mask, T* sm, const SmemLayout& L, T (&pv)[10]) {
    T H[14];
#pragma unroll
    for (int i = 0; i < 14; i++) H[i] = T(0);
    add_cost<T>(H, c, j, N, true, sm + L.oX, sm + L.oY, sm + L.oPar);
    __syncwarp(mask);
    if (j < 10) {
#pragma unroll
        for (int i = 0; i < 10; i++)
            if (i >= j) sm[L.oP + i * (i + 1) / 2 + j] = H[i];
    } else if (j == 14) {
#pragma unroll
        for (int i = 0; i < 10; i++) sm[L.op + i] = H[i];
    }
#pragma unroll
    for (int i = 0; i < 10; i++) pv[i] = H[i];
    __syncwarp(mask);
}

// Active set of one lane's box (the input or velocity component this lane owns) over the horizon: bit k of
// `lo` / `hi` = pinned at the lower / upper bound at stage k.  Two 64-bit words each: N <= 128.
struct StageMask {
    unsigned long long w0, w1;
    __device__ __forceinline__ StageMask() : w0(0ull), w1(0ull) {}
    __device__ __forceinline__ bool test(int k) const { return (((k < 64) ? w0 : w1) >> (k & 63)) & 1ull; }
    __device__ __forceinline__ void set(int k) { const unsigned long long b = 1ull << (k & 63); if (k < 64) w0 |= b; else w1 |= b; }
    __device__ __forceinline__ void clear(int k) { const unsigned long long b = ~(1ull << (k & 63)); if (k < 64) w0 &= b; else w1 &= b; }
};
// active-set data of the rounds that keep it in registers (backward_stage kBar == 2)
template <typename T>
struct ActiveSet {
    StageMask lo_m, hi_m;
    T lo, hi;
};

// One backward Riccati stage on the tile `sT` ([10][TLD]: columns 6..13 of [A_k B_k], then b_k), which
// must be complete and visible.  colp = this lane's column inside the tile (or the constant tile of the
// trivial columns 0..5: dx+/dp = [I;0;0], dx+/dv = [hI;I;0]).  kBar: add the barrier / active-set
// diagonal and gradient (1: from the workspace arrays oBarD / oBarG; 2: pinned inputs of the register-held
// active set `as`); kRows: store the rows of [Hux Guu | g_u] needed by the active-set multiplier test.
// Returns false on a non-positive pivot.
//
// kBar == 2 also pins VELOCITY components exactly.  A pinned component a of x_{k+1} (lanes 3..5 carry those masks) is the
// stage-k mixed constraint  E (A dx_k + B du_k + b_k) = beta,  resolved in the range space of the inputs:
//   Y = G^-1 (E B)',  S = (E B) Y,   K = K_unc + Y S^-1 R_x,  kappa = kappa_unc + Y S^-1 r_0,
//   R = [0 | beta] - E [A | b] + (E B) G^-1 [H_ux | g_u]   (one column per lane),   P = P_unc + R_x' S^-1 R_x,  p = p_unc + R_x' S^-1 r_0
// -- an added positive semi-definite term, so nothing of size 1/mu is ever subtracted (the barrier-weighted recursion
// loses every fp32 digit on such problems; oracle/nmpc_oracle.c orc_eq_solve is the scalar restatement of this block).
// With kRows the rows T = S^-1 R are kept for the multiplier test  nu = -(T_x dx_k + t_0).
template <typename T, int kBar, bool kRows>
__device__ __forceinline__ bool backward_stage(const RtiCfg<T>& c, int N, int k, int j, unsigned mask, T* sm, const SmemLayout& L, T* ws,
                                               const WsLayout& WL, T* rec, const T* __restrict__ sT, const T* __restrict__ colp,
                                               const ActiveSet<T>* as, T (&pv)[10]) {
    const T* sX = sm + L.oX;
    const T* sU = sm + L.oU;
    const T* sPar = sm + L.oPar;
    T* sP = sm + L.oP;
    T* sp = sm + L.op;
    T* sHux = sm + L.oHux;
    T col[10];
#pragma unroll
    for (int r = 0; r < 10; r++) col[r] = colp[r * TLD];
    // W = P+ col (+ p+ on the gradient lane)
    // p+ lives in the gradient lane's registers between the stages of the nominal sweep (pv: this lane's Pn of the
    // previous stage, i.e. p+ on lane 14); the constrained sweeps keep it in shared memory, where the partial-sweep
    // save / restore finds (P | p) contiguous
    T W[10], spv[12];
    if (kBar != 0) {
        Vec4<T>::ld(sp, spv[0], spv[1], spv[2], spv[3]);
        Vec4<T>::ld(sp + 4, spv[4], spv[5], spv[6], spv[7]);
        Vec4<T>::ld(sp + 8, spv[8], spv[9], spv[10], spv[11]);
    } else {
#pragma unroll
        for (int i = 0; i < 10; i++) spv[i] = pv[i];
    }
    // P+ streams in once as its packed lower triangle (14 broadcast 128-bit loads instead of 30 for the full rows --
    // shared-memory wavefronts are what this kernel runs out of): element (a, b) feeds W[a] and, off the diagonal, W[b]
    T tri[PTRI];
#pragma unroll
    for (int q = 0; q < PTRI / 4; q++) Vec4<T>::ld(sP + 4 * q, tri[4 * q], tri[4 * q + 1], tri[4 * q + 2], tri[4 * q + 3]);
#pragma unroll
    for (int i = 0; i < 10; i++) W[i] = (j == 14) ? spv[i] : T(0);
#pragma unroll
    for (int a = 0; a < 10; a++) {
#pragma unroll
        for (int b = 0; b <= a; b++) {
            const T pv = tri[a * (a + 1) / 2 + b];
            W[a] += pv * col[b];
            if (b != a) W[b] += pv * col[a];
        }
    }
    // H[:,j] = [A B]' W  (columns 0..5 of [A B] are [I; 0; 0] and [hI; I; 0])
    T H[14];
    H[0] = W[0]; H[1] = W[1]; H[2] = W[2];
    H[3] = c.h * W[0] + W[3];
    H[4] = c.h * W[1] + W[4];
    H[5] = c.h * W[2] + W[5];
#pragma unroll
    for (int i = 6; i < 14; i++) H[i] = T(0);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        T a0, a1, a2, a3, a4, a5, a6, a7;
        Vec4<T>::ld(sT + r * TLD, a0, a1, a2, a3);
        Vec4<T>::ld(sT + r * TLD + 4, a4, a5, a6, a7);
        fma2<T>(H[6], H[7], a0, a1, W[r], W[r]);
        fma2<T>(H[8], H[9], a2, a3, W[r], W[r]);
        fma2<T>(H[10], H[11], a4, a5, W[r], W[r]);
        fma2<T>(H[12], H[13], a6, a7, W[r], W[r]);
    }
    add_cost<T>(H, c, j, k, false, sX, sm + L.oY, sPar);
    if (kRows) {
        // rows of the un-penalised [Hux Guu] (by symmetry: column 10+m) and g_u, for the multiplier test
        if (j >= 10 && j < 14) {
            T* hrow = ws + WL.oHrow + (k * 4 + (j - 10)) * 16;
            Vec4<T>::st(hrow, H[0], H[1], H[2], H[3]);
            Vec4<T>::st(hrow + 4, H[4], H[5], H[6], H[7]);
            Vec4<T>::st(hrow + 8, H[8], H[9], H[10], H[11]);
            hrow[12] = H[12];
            hrow[13] = H[13];
        } else if (j == 14) {
#pragma unroll
            for (int m = 0; m < 4; m++) ws[WL.oHrow + (k * 4 + m) * 16 + 14] = H[10 + m];
        }
    }
    // pinned inputs of the register-held active set (kBar == 2): flag and value (bound - iterate) of each input, known to
    // every lane; the elimination itself happens on the gathered 4x4 block below
    bool upin[4] = {false, false, false, false};
    T ubeta[4] = {T(0), T(0), T(0), T(0)};
    bool pin = false, mine = false;  // kBar == 2: this lane's input is pinned at stage k / its velocity component at stage k + 1
    T bnd = T(0);
    if (kBar == 2) {
        const bool at_lo = as->lo_m.test(k), at_hi = as->hi_m.test(k);
        pin = (j >= 10 && j < 14) && (at_lo || at_hi);
        bnd = (at_lo ? as->lo : as->hi) - sU[k * NU + ((j - 10) & 3)];
        mine = (j >= 3 && j < 6) && (k + 1 < N) && (as->lo_m.test(k + 1) || as->hi_m.test(k + 1));
    } else if (kBar == 1) {
        // barrier diagonal / gradient of the boxed variables (velocities: slots 0..2, inputs: 3..6 of the stage's row)
        const T* sBD = sm + L.oIpm + 6 * N * 8;
        const T* sBG = sBD + (N + 1) * 8;
        if (j < 14) {
            const bool boxl = (j >= 3 && j < 6) || j >= 10;
            const T d = boxl ? sBD[k * 8 + as_owner(j)] : T(0);
#pragma unroll
            for (int i = 0; i < 14; i++) H[i] += (j == i) ? d : T(0);
        } else if (j == 14) {
            T g0, g1, g2, g3, g4, g5, g6, g7;
            Vec4<T>::ld(sBG + k * 8, g0, g1, g2, g3);
            Vec4<T>::ld(sBG + k * 8 + 4, g4, g5, g6, g7);
            H[3] += g0; H[4] += g1; H[5] += g2;
            H[10] += g3; H[11] += g4; H[12] += g5; H[13] += g6;
        }
    }
    // 4x4 input block G = Huu from lanes 10..13, Cholesky, solve for this lane's column
    T g00, g10, g20, g30, g11, g21, g31, g22, g32, g33;
    unsigned vpin_bits = 0u;  // kBar == 2: velocity components of stage k + 1 pinned (bit a)
    if (kBar == 0) {
        // nominal kernel: whole-warp lockstep under a compile-time mask -- bare shuffles
        g00 = __shfl_sync(mask, H[10], 10, GL); g10 = __shfl_sync(mask, H[11], 10, GL);
        g20 = __shfl_sync(mask, H[12], 10, GL); g30 = __shfl_sync(mask, H[13], 10, GL);
        g11 = __shfl_sync(mask, H[11], 11, GL); g21 = __shfl_sync(mask, H[12], 11, GL);
        g31 = __shfl_sync(mask, H[13], 11, GL); g22 = __shfl_sync(mask, H[12], 12, GL);
        g32 = __shfl_sync(mask, H[13], 12, GL); g33 = __shfl_sync(mask, H[13], 13, GL);
    } else {
        // constrained kernel: the two halves of a warp diverge, so every shuffle on the run-time half mask is wrapped in a
        // WARPSYNC ... ENDCOLLECTIVE pair (18 shuffles + a ballot per stage).  One exchange through shared memory instead:
        // the Hux region is dead until this stage's rows are stored below.
        T* sx = sHux;
        if (j >= 10 && j < 14) {
            Vec4<T>::st(sx + (j - 10) * 4, H[10], H[11], H[12], H[13]);
            if (kBar == 2) { sx[16 + (j - 10)] = pin ? T(1) : T(0); sx[20 + (j - 10)] = bnd; }
        }
        if (kBar == 2 && j >= 3 && j < 6) sx[24 + (j - 3)] = mine ? T(1) : T(0);
        __syncwarp(mask);
        T t0, t1, t2, t3;
        Vec4<T>::ld(sx, g00, g10, g20, g30);
        Vec4<T>::ld(sx + 4, t0, g11, g21, g31);
        Vec4<T>::ld(sx + 8, t0, t1, g22, g32);
        Vec4<T>::ld(sx + 12, t0, t1, t2, g33);
        if (kBar == 2) {
            Vec4<T>::ld(sx + 16, t0, t1, t2, t3);
            upin[0] = t0 != T(0); upin[1] = t1 != T(0); upin[2] = t2 != T(0); upin[3] = t3 != T(0);
            Vec4<T>::ld(sx + 20, ubeta[0], ubeta[1], ubeta[2], ubeta[3]);
            Vec4<T>::ld(sx + 24, t0, t1, t2, t3);
            vpin_bits = (t0 != T(0) ? 1u : 0u) | (t1 != T(0) ? 2u : 0u) | (t2 != T(0) ? 4u : 0u);
        }
        __syncwarp(mask);
    }
    const T hu0 = H[10], hu1 = H[11], hu2 = H[12], hu3 = H[13];  // un-eliminated rows of this column (what sHux keeps)
    if (kBar == 2) {
        // EXACT elimination of the pinned inputs: du_m = beta_m.  Their rows / columns leave the 4x4 block (identity
        // instead), the free rows of the gradient column pick up G(f, m) beta_m, and the pinned rows of every column
        // become the constant -beta_m (gradient lane) or 0, so that the common solve returns kappa_m = beta_m, K(m, :) = 0.
        if (upin[0] | upin[1] | upin[2] | upin[3]) {
            if (j == 14) {
                H[10] += (upin[1] ? g10 * ubeta[1] : T(0)) + (upin[2] ? g20 * ubeta[2] : T(0)) + (upin[3] ? g30 * ubeta[3] : T(0));
                H[11] += (upin[0] ? g10 * ubeta[0] : T(0)) + (upin[2] ? g21 * ubeta[2] : T(0)) + (upin[3] ? g31 * ubeta[3] : T(0));
                H[12] += (upin[0] ? g20 * ubeta[0] : T(0)) + (upin[1] ? g21 * ubeta[1] : T(0)) + (upin[3] ? g32 * ubeta[3] : T(0));
                H[13] += (upin[0] ? g30 * ubeta[0] : T(0)) + (upin[1] ? g31 * ubeta[1] : T(0)) + (upin[2] ? g32 * ubeta[2] : T(0));
            }
#pragma unroll
            for (int m = 0; m < 4; m++)
                if (upin[m]) H[10 + m] = (j == 14) ? -ubeta[m] : T(0);
            if (upin[0]) { g00 = T(1); g10 = T(0); g20 = T(0); g30 = T(0); }
            if (upin[1]) { g11 = T(1); g10 = T(0); g21 = T(0); g31 = T(0); }
            if (upin[2]) { g22 = T(1); g20 = T(0); g21 = T(0); g32 = T(0); }
            if (upin[3]) { g33 = T(1); g30 = T(0); g31 = T(0); g32 = T(0); }
        }
    }
    const T i00 = trsqrt(g00);
    const T l10 = g10 * i00, l20 = g20 * i00, l30 = g30 * i00;
    const T e1 = g11 - l10 * l10;
    const T i11 = trsqrt(e1);
    const T l21 = (g21 - l20 * l10) * i11, l31 = (g31 - l30 * l10) * i11;
    const T e2 = g22 - l20 * l20 - l21 * l21;
    const T i22 = trsqrt(e2);
    const T l32 = (g32 - l30 * l20 - l31 * l21) * i22;
    const T e3 = g33 - l30 * l30 - l31 * l31 - l32 * l32;
    const T i33 = trsqrt(e3);
    const bool ok = (g00 > T(0)) && (e1 > T(0)) && (e2 > T(0)) && (e3 > T(0));
    if (kBar == 1 && j == 15) {
        // interior-point sweeps keep the factors of G: the corrector of the same iteration changes only the gradient
        // (same barrier weights), so it re-solves with them instead of factorising again (delta_backward)
        T* fq = ws + WL.oHrow + (long long)k * 64;
        Vec4<T>::st(fq, i00, l10, l20, l30);
        Vec4<T>::st(fq + 4, i11, l21, l31, i22);
        Vec4<T>::st(fq + 8, l32, i33, T(0), T(0));
    }
    const T y0 = H[10] * i00;
    const T y1 = (H[11] - l10 * y0) * i11;
    const T y2 = (H[12] - l20 * y0 - l21 * y1) * i22;
    const T y3 = (H[13] - l30 * y0 - l31 * y1 - l32 * y2) * i33;
    const T x3 = y3 * i33;
    const T x2 = (y2 - l32 * x3) * i22;
    const T x1 = (y1 - l21 * x2 - l31 * x3) * i11;
    const T x0 = (y0 - l10 * x1 - l20 * x2 - l30 * x3) * i00;
    T K0 = -x0, K1 = -x1, K2 = -x2, K3 = -x3;  // K[:,j] (j < 10) or kappa (j == 14)
    const T K0u = K0, K1u = K1, K2u = K2, K3u = K3;  // the unconstrained feedback column (K0..K3 may pick up the velocity-pin terms)
    bool ok_v = true;
    bool vpins = false;
    T tv0 = T(0), tv1 = T(0), tv2 = T(0);
    if (kBar == 2) {
        const int k1 = k + 1;
        const unsigned pins = vpin_bits;
        if (pins) {  // uniform over the group
            vpins = true;
            const T beta_own = mine ? ((as->lo_m.test(k1) ? as->lo : as->hi) - sX[k1 * NX + j]) : T(0);
            T D[3][4], Df[3][4], beta[3], Y[3][4];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const bool on = (pins >> a) & 1u;
                beta[a] = __shfl_sync(mask, beta_own, 3 + a, GL);
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const T d = __shfl_sync(mask, col[3 + a], 10 + m, GL);
                    Df[a][m] = on ? d : T(0);                 // E B, every input (right-hand side: pinned inputs move x+ by B beta)
                    D[a][m] = (on && !upin[m]) ? d : T(0);    // E B restricted to the free inputs (the constraint's handle)
                }
                // Y_a = G^-1 D_a' with the Cholesky factors of G
                const T f0 = D[a][0] * i00;
                const T f1 = (D[a][1] - l10 * f0) * i11;
                const T f2 = (D[a][2] - l20 * f0 - l21 * f1) * i22;
                const T f3 = (D[a][3] - l30 * f0 - l31 * f1 - l32 * f2) * i33;
                Y[a][3] = f3 * i33;
                Y[a][2] = (f2 - l32 * Y[a][3]) * i22;
                Y[a][1] = (f1 - l21 * Y[a][2] - l31 * Y[a][3]) * i11;
                Y[a][0] = (f0 - l10 * Y[a][1] - l20 * Y[a][2] - l30 * Y[a][3]) * i00;
            }
            auto dot4 = [](const T (&p)[4], const T (&q)[4]) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2] + p[3] * q[3]; };
            // S = D Y (identity on the rows that are not pinned), 3x3 Cholesky
            const T s00 = dot4(D[0], Y[0]) + ((pins & 1u) ? T(0) : T(1));
            const T s10 = dot4(D[1], Y[0]), s11 = dot4(D[1], Y[1]) + ((pins & 2u) ? T(0) : T(1));
            const T s20 = dot4(D[2], Y[0]), s21 = dot4(D[2], Y[1]), s22 = dot4(D[2], Y[2]) + ((pins & 4u) ? T(0) : T(1));
            const T piv = sizeof(T) == 4 ? T(1e-7) : T(1e-10);  // (E B_free) G^-1 (E B_free)' of a controllable component is >~ 1e-5
            const T j00 = trsqrt(s00);
            const T m10 = s10 * j00, m20 = s20 * j00;
            const T d1 = s11 - m10 * m10;
            const T j11 = trsqrt(d1);
            const T m21 = (s21 - m20 * m10) * j11;
            const T d2 = s22 - m20 * m20 - m21 * m21;
            const T j22 = trsqrt(d2);
            ok_v = (s00 > piv) && (d1 > piv) && (d2 > piv);
            // this lane's column of R, then t = S^-1 r
            const T xs[4] = {x0, x1, x2, x3};
            T r[3];
#pragma unroll
            for (int a = 0; a < 3; a++)
                r[a] = ((pins >> a) & 1u) ? (((j == 14) ? beta[a] : T(0)) - col[3 + a] + dot4(Df[a], xs)) : T(0);
            const T q0 = r[0] * j00;
            const T q1 = (r[1] - m10 * q0) * j11;
            const T q2 = (r[2] - m20 * q0 - m21 * q1) * j22;
            tv2 = q2 * j22;
            tv1 = (q1 - m21 * tv2) * j11;
            tv0 = (q0 - m10 * tv1 - m20 * tv2) * j00;
            K0 += Y[0][0] * tv0 + Y[1][0] * tv1 + Y[2][0] * tv2;
            K1 += Y[0][1] * tv0 + Y[1][1] * tv1 + Y[2][1] * tv2;
            K2 += Y[0][2] * tv0 + Y[1][2] * tv1 + Y[2][2] * tv2;
            K3 += Y[0][3] * tv0 + Y[1][3] * tv1 + Y[2][3] * tv2;
            // R rows through shared memory (the step region is dead during a backward sweep) for the P update
            T* sRv = sm + L.oDz;
            if (j < 10) { sRv[j] = r[0]; sRv[12 + j] = r[1]; sRv[24 + j] = r[2]; }
            if (kRows && (j < 10 || j == 14)) {
                T* tvp = ws + WL.oTv + (long long)k * 36 + ((j == 14) ? 10 : j);
                tvp[0] = tv0; tvp[12] = tv1; tvp[24] = tv2;
            }
        }
    }
    if (j < 10) Vec4<T>::st(sHux + j * 4, hu0, hu1, hu2, hu3);
    // feedback rows for the forward sweep: rec[k][10+m][j] = K(m, j), [10] = kappa_m
    if (j < 10 || j == 14) {
        T* kr = rec + ((long long)k * 14 + 10) * TLD + ((j == 14) ? 10 : j);
        kr[0] = K0; kr[TLD] = K1; kr[2 * TLD] = K2; kr[3 * TLD] = K3;
    }
    __syncwarp(mask);
    T Pn[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        T h0, h1, h2, h3;
        Vec4<T>::ld(sHux + i * 4, h0, h1, h2, h3);
        T pa = H[i], pb = T(0);  // unconstrained part: H_xx + H_xu K_unc
        fma2<T>(pa, pb, h0, h1, K0u, K1u);
        fma2<T>(pa, pb, h2, h3, K2u, K3u);
        Pn[i] = pa + pb;
    }
    if (kBar == 2 && vpins) {
        const T* sRv = sm + L.oDz;
#pragma unroll
        for (int i = 0; i < 10; i++) Pn[i] += sRv[i] * tv0 + sRv[12 + i] * tv1 + sRv[24 + i] * tv2;
    }
    if (j < 10) {
        // only the lower triangle is kept (entry (i, j), i >= j, from lane j): P stays exactly symmetric -- the
        // antisymmetric rounding part would be amplified by the recursion (1e-3 relative error at N = 80 in fp32)
#pragma unroll
        for (int i = 0; i < 10; i++) {
            if (i >= j) sP[i * (i + 1) / 2 + j] = Pn[i];
        }
    } else if (kBar != 0 && j == 14) {
        Vec4<T>::st(sp, Pn[0], Pn[1], Pn[2], Pn[3]);
        Vec4<T>::st(sp + 4, Pn[4], Pn[5], Pn[6], Pn[7]);
        Vec4<T>::st(sp + 8, Pn[8], Pn[9], T(0), T(0));
    }
#pragma unroll
    for (int i = 0; i < 10; i++) pv[i] = Pn[i];
    __syncwarp(mask);
    if (kBar == 2) {
        // keep (P_k | p_k) -- PSAVE contiguous elements from sP -- for partial sweeps of later rounds
        T* dst = ws + WL.oPs + (long long)k * PSAVE;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int idx = j + q * GL;
            if (idx < PSAVE / 4) {
                T a0, a1, a2, a3;
                Vec4<T>::ld(sP + idx * 4, a0, a1, a2, a3);
                Vec4<T>::st(dst + idx * 4, a0, a1, a2, a3);
            }
        }
    }
    return ok && ok_v;
}

// copy one finished tile ([10][TLD], 30 vectors of 4) to the workspace record of stage k
template <typename T>
__device__ __forceinline__ void tile_to_ws(const T* __restrict__ sT, T* __restrict__ rec, int lane) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int idx = lane + q * GL;
        if (idx < 30) {
            T a, b, cc, d;
            Vec4<T>::ld(sT + idx * 4, a, b, cc, d);
            Vec4<T>::st(rec + idx * 4, a, b, cc, d);
        }
    }
}

// Backward sweep.  kLin: linearise on the fly -- the 16 lanes integrate the 8 non-trivial sensitivity
// columns (q0, omega, c) of TWO intervals per pass (lanes 0-7: interval k, lanes 8-15: interval k-1; the
// state is integrated redundantly by every lane, so x+ comes for free) straight into the two tiles,
// which are also saved to the workspace for the forward sweep / later IPM sweeps.  !kLin: tiles are
// re-loaded from the workspace, one stage ahead of their use.
template <typename T, bool kLin, int kBar, bool kRows>
__device__ __forceinline__ bool backward_sweep(const RtiCfg<T>& c, int N, int j, unsigned mask, T* sm, const SmemLayout& L, T* ws,
                                               const WsLayout& WL, T* rec, const T* __restrict__ sTriv, const ActiveSet<T>* as, bool zero_b = false,
                                               int k_top = 1 << 30) {
    // k_top (!kLin, kBar == 2): the highest stage whose pins changed since the previous sweep of this problem -- the
    // stages above it are unchanged, so the recursion restarts from the saved (P, p) of stage k_top + 1
    const int k_first = (!kLin && kBar == 2 && k_top < N - 1) ? k_top : N - 1;
    T pv[10];
#pragma unroll
    for (int i = 0; i < 10; i++) pv[i] = T(0);
    if (k_first == N - 1) {
        backward_terminal<T>(c, N, j, mask, sm, L, pv);
    } else {
        const T* src = ws + WL.oPs + (long long)(k_first + 1) * PSAVE;
        T* sP = sm + L.oP;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int idx = j + q * GL;
            if (idx < PSAVE / 4) {
                T a0, a1, a2, a3;
                Vec4<T>::ld(src + idx * 4, a0, a1, a2, a3);
                Vec4<T>::st(sP + idx * 4, a0, a1, a2, a3);
            }
        }
        __syncwarp(mask);
    }
    bool ok = true;
    T* sT0 = sm + L.oT0;
    T* sT1 = sm + L.oT1;
    const int jc = (j >= 6 && j < 15) ? j - 6 : 9;  // lane 15 reads a zeroed pad column
    const T* col0 = (j < 6) ? sTriv + j : sT0 + jc;
    const T* col1 = (j < 6) ? sTriv + j : sT1 + jc;
    if (kLin) {
        const T* sX = sm + L.oX;
        const T* sU = sm + L.oU;
        const T* sPar = sm + L.oPar;
        const int half = j >> 3, cj = j & 7;
        for (int k = N - 1; k >= 0; k -= 2) {
            const int kk = (k - half >= 0) ? k - half : 0;
            {
                T xa[10], sa[10];
                const T* pr = sPar + kk * NPS;
                rk4_column<T>(c, 6 + cj, sX + kk * NX, sU + kk * NU, pr[4] * c.inv_mass, pr[5] * c.inv_mass, pr[6] * c.inv_mass, xa, sa);
                T* t = (half ? sT1 : sT0) + cj;
#pragma unroll
                for (int r = 0; r < 10; r++) t[r * TLD] = sa[r];
                // b = x+ - X_{k+1}: every lane holds x+; lane cj writes row cj, lanes 0/1 also rows 8/9
                T b0 = xa[0];
#pragma unroll
                for (int r = 1; r < 8; r++) b0 = (cj == r) ? xa[r] : b0;
                const T b1 = (cj == 0) ? xa[8] : xa[9];
                T* tb = (half ? sT1 : sT0) + 8;
                const T* xn = sX + (kk + 1) * NX;
                tb[cj * TLD] = b0 - xn[cj];
                if (cj < 2) tb[(8 + cj) * TLD] = b1 - xn[8 + cj];
            }
            __syncwarp(mask);
            tile_to_ws<T>(sT0, rec + (long long)k * 14 * TLD, j);
            ok &= backward_stage<T, kBar, kRows>(c, N, k, j, mask, sm, L, ws, WL, rec, sT0, col0, as, pv);
            if (k >= 1) {
                tile_to_ws<T>(sT1, rec + (long long)(k - 1) * 14 * TLD, j);
                ok &= backward_stage<T, kBar, kRows>(c, N, k - 1, j, mask, sm, L, ws, WL, rec, sT1, col1, as, pv);
            }
        }
    } else {
        // tile of stage k lives in sT[(N-1-k) & 1]; the loads of stage k-1 are issued before stage k runs
        T pre[2][4];
        auto fetch = [&](int k) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int idx = j + q * GL;
                if (idx < 30) Vec4<T>::ld(rec + (long long)k * 14 * TLD + idx * 4, pre[q][0], pre[q][1], pre[q][2], pre[q][3]);
            }
        };
        auto put = [&](T* t) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int idx = j + q * GL;
                // vector idx covers tile elements 4 idx .. 4 idx + 3; the b column is element 8 of each 12-wide row
                if (idx < 30) Vec4<T>::st(t + idx * 4, (zero_b && (idx % 3) == 2) ? T(0) : pre[q][0], pre[q][1], pre[q][2], pre[q][3]);
            }
        };
        fetch(k_first);
        put(sT0);
        __syncwarp(mask);
        int par = 0;
        for (int k = k_first; k >= 0; k--) {
            if (k >= 1) fetch(k - 1);
            ok &= backward_stage<T, kBar, kRows>(c, N, k, j, mask, sm, L, ws, WL, rec, par ? sT1 : sT0, par ? col1 : col0, as, pv);
            if (k >= 1) {
                put(par ? sT0 : sT1);
                __syncwarp(mask);
            }
            par ^= 1;
        }
    }
    return grp_all(mask, ok);
}

// Gradient-only backward sweep for a CHANGE dG of the stage gradients (dG[k][8]: the boxed variables, as_owner order) under the
// factorisation left by the last backward_sweep<.., kBar = 1, ..>: the Riccati matrices P_k, the feedback K_k and the
// factors of G_k do not depend on the gradient, so
//     dg = [A B]' dp+ + dG_k,   dkappa = -G^-1 dg_u,   dp = dg_x + K' dg_u
// and the step changes by the forward sweep of (K, dkappa) from dx_0 = 0 with b = 0.  Lane i < 14 carries dg_i, lanes
// 0..9 dp_i; dkappa replaces kappa in the feedback rows of `rec`.  About a fifth of a factorising sweep: this is the
// corrector of the Mehrotra iteration (HPIPM re-uses its factorisation in the same way).
template <typename T>
__device__ __forceinline__ void delta_backward(const RtiCfg<T>& c, int N, int j, unsigned mask, const T* __restrict__ ws, const WsLayout& WL,
                                               T* rec, const T* __restrict__ sTriv, const T* __restrict__ dG) {
    const int jc = (j >= 6 && j < 14) ? j - 6 : 0;
    const bool tcol = (j >= 6 && j < 14), xl = j < 10, ul = (j >= 10 && j < 14), boxl = (j >= 3 && j < 6) || ul;
    T triv[10];
#pragma unroll
    for (int r = 0; r < 10; r++) triv[r] = (j < 6) ? sTriv[r * TLD + j] : T(0);
    T col[10], fq[10], kf[4], dg0;
    auto fetch = [&](int k) {
        const T* rk = rec + (long long)k * 14 * TLD;
#pragma unroll
        for (int r = 0; r < 10; r++) col[r] = tcol ? rk[r * TLD + jc] : triv[r];
        const T* f = ws + WL.oHrow + (long long)k * 64;
        Vec4<T>::ld(f, fq[0], fq[1], fq[2], fq[3]);
        Vec4<T>::ld(f + 4, fq[4], fq[5], fq[6], fq[7]);
        T d0, d1;
        Vec4<T>::ld(f + 8, fq[8], fq[9], d0, d1);
#pragma unroll
        for (int m = 0; m < 4; m++) kf[m] = xl ? rk[(10 + m) * TLD + j] : T(0);
        dg0 = boxl ? dG[k * 8 + as_owner(j)] : T(0);
    };
    T dp = T(0);
    fetch(N - 1);
    for (int k = N - 1; k >= 0; k--) {
        T g = dg0;
#pragma unroll
        for (int r = 0; r < 10; r++) g += col[r] * __shfl_sync(mask, dp, r, GL);
        const T i00 = fq[0], l10 = fq[1], l20 = fq[2], l30 = fq[3], i11 = fq[4], l21 = fq[5], l31 = fq[6], i22 = fq[7], l32 = fq[8], i33 = fq[9];
        const T k0 = kf[0], k1 = kf[1], k2 = kf[2], k3 = kf[3];
        if (k > 0) fetch(k - 1);   // next stage's records travel while this one is solved
        const T gu0 = __shfl_sync(mask, g, 10, GL), gu1 = __shfl_sync(mask, g, 11, GL);
        const T gu2 = __shfl_sync(mask, g, 12, GL), gu3 = __shfl_sync(mask, g, 13, GL);
        const T y0 = gu0 * i00;
        const T y1 = (gu1 - l10 * y0) * i11;
        const T y2 = (gu2 - l20 * y0 - l21 * y1) * i22;
        const T y3 = (gu3 - l30 * y0 - l31 * y1 - l32 * y2) * i33;
        const T x3 = y3 * i33;
        const T x2 = (y2 - l32 * x3) * i22;
        const T x1 = (y1 - l21 * x2 - l31 * x3) * i11;
        const T x0 = (y0 - l10 * x1 - l20 * x2 - l30 * x3) * i00;
        if (ul) {
            const T xm = (j == 10) ? x0 : ((j == 11) ? x1 : ((j == 12) ? x2 : x3));
            rec[((long long)k * 14 + j) * TLD + 10] = -xm;
        }
        dp = xl ? g + k0 * gu0 + k1 * gu1 + k2 * gu2 + k3 * gu3 : T(0);
    }
    __syncwarp(mask);
}

// 4/8/16-byte asynchronous global -> shared copies (LDGSTS): the whole problem record is requested
// up front and lands while nothing else waits on it
template <int kBytes>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(kBytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

 // L1 prefetch of a global line that is read a little later
__device__ __forceinline__ void pf_l1(const void* g) { asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(g))); }

constexpr int FW_RING = 6;            // stages of forward-sweep records in flight (L2 latency / stage time ~ 4-5)
constexpr int FW_REC = 14 * TLD;      // one stage: 14 lanes x TLD
static_assert(SmemLayout(1, true).oHux - SmemLayout(1, true).oY >= FW_RING * FW_REC && SmemLayout(80, true).oHux - SmemLayout(80, true).oY >= FW_RING * FW_REC,
              "forward-sweep ring does not fit the dead shared-memory regions");

// Forward substitution: lane i < 10 carries dx_k[i], lanes 10..13 compute du_k[m].
// kFinal (the accepted sweep of the nominal path): the stage records stream from the L2-resident
// workspace through a FW_RING-deep shared-memory ring (cp.async groups; each lane copies and reads only
// its own record, so no barrier is needed) laid over the cost records, sDz, P+, p+ and the two stage tiles,
// which are all dead by then (the tiles' zero pad columns are restored afterwards); the
// new iterate (X + dx, U + du) replaces the old one IN SHARED MEMORY as it is produced (the caller copies it out with
// wide stores once the step is accepted -- global memory keeps the old iterate until then, which is what a failed or
// handed-over problem needs) and the box test / NaN test / active count are fused in (returned through viol / bad / nact).
// !kFinal (IPM sweeps): records are register-prefetched and the step goes to sDz[k][lane].
template <typename T, bool kFinal>
__device__ __forceinline__ void forward_sweep(const RtiCfg<T>& c, int N, int lane, unsigned mask, T dx0, T* sm, const SmemLayout& L,
                                              const T* rec_base, const WsLayout& WL, T lo, T hi, T* gX, T* gU, T* gu0, bool& viol, bool& bad,
                                              int& nact, bool rezero_pads = true, bool zero_b = false, bool accum = false) {
    T* sDz = sm + L.oDz;
    const bool isx = lane < 10, isv = (lane >= 3 && lane < 6);
    const T* rec = rec_base + (long long)((lane < 14) ? lane : 13) * TLD;
    T z = isx ? dx0 : T(0);
    // one stage: consumes this lane's record cf, the iterate value it_cur (kFinal)
    auto stage = [&](int k, const T (&cf)[12], T& du_out) {
        T xj[10];
#pragma unroll
        for (int jj = 0; jj < 10; jj++) xj[jj] = __shfl_sync(mask, z, jj, GL);
        const T zv = __shfl_sync(mask, z, (lane + 3) & 15, GL);
        T du = cf[10], du2 = T(0);
#pragma unroll
        for (int jj = 0; jj < 5; jj++) {
            du += cf[jj] * xj[jj];
            du2 += cf[5 + jj] * xj[5 + jj];
        }
        du += du2;
        T um[4];
#pragma unroll
        for (int m = 0; m < 4; m++) um[m] = __shfl_sync(mask, du, 10 + m, GL);
        T xn = ((zero_b && !kFinal) ? T(0) : cf[8]) + ((lane < 3) ? z + c.h * zv : ((lane < 6) ? z : T(0)));
        xn += cf[0] * xj[6] + cf[1] * xj[7] + cf[2] * xj[8] + cf[3] * xj[9];
        xn += cf[4] * um[0] + cf[5] * um[1] + cf[6] * um[2] + cf[7] * um[3];
        du_out = isx ? z : du;
        z = xn;
    };
    if (kFinal) {
        T* itp = isx ? sm + L.oX + lane : sm + L.oU + ((lane - 10) & 3);
        const int its = isx ? NX : NU;
        T* ring = sm + L.oY + ((lane < 14) ? lane : 13) * TLD;  // [FW_RING][14][TLD] over sY | sDz
        auto issue = [&](int k, int slot) {
            if (k < N && lane < 14) {
                const T* r = rec + (long long)k * FW_REC;
                T* d = ring + slot * FW_REC;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    if (sizeof(T) == 4) cp_async<16>(d + 4 * q, r + 4 * q);
                    else { cp_async<16>(d + 4 * q, r + 4 * q); cp_async<16>(d + 4 * q + 2, r + 4 * q + 2); }
                }
            }
            cp_async_commit();
        };
        __syncwarp(mask);  // every lane is done with the cost records
#pragma unroll
        for (int u = 0; u < FW_RING; u++) issue(u, u);
        bool v_l = false, b_l = false;
        int n_l = 0;
        T it_cur = itp[0];
        // ring slot of stage k = k % FW_RING: the loop runs in blocks of FW_RING stages so that the slot is a constant
        for (int k0 = 0; k0 < N; k0 += FW_RING) {
#pragma unroll
          for (int u = 0; u < FW_RING; u++) {
            const int k = k0 + u;
            if (k >= N) break;
            const T it_nxt = (k + 1 < N || isx) ? itp[(k + 1) * its] : T(0);
            cp_async_wait<FW_RING - 1>();
            T cf[12];
            const T* d = ring + u * FW_REC;
            Vec4<T>::ld(d, cf[0], cf[1], cf[2], cf[3]);
            Vec4<T>::ld(d + 4, cf[4], cf[5], cf[6], cf[7]);
            Vec4<T>::ld(d + 8, cf[8], cf[9], cf[10], cf[11]);
            issue(k + FW_RING, u);
            T dz;
            stage(k, cf, dz);
            {
                // branch-free: lanes without a box carry lo / hi = -+1e30 (lane_box), lanes 14 / 15 repeat lane 13's record;
                // the velocity box starts at stage 1
                const T v = it_cur + dz;
                if (lane < 14) itp[k * its] = v;
                b_l |= !(fabs(v) <= T(1e30));
                const bool chk = !(isv && k == 0);
                v_l |= chk && !(v >= lo && v <= hi);
                n_l += (int)(chk && v <= lo) + (int)(chk && v >= hi);
            }
            it_cur = it_nxt;
          }
        }
        cp_async_wait<0>();
        if (isx) {
            const T v = it_cur + z;
            itp[N * its] = v;
            b_l |= !(fabs(v) <= T(1e30));
        }
        __syncwarp(mask);
        // the ring ran over the stage tiles: restore their zero pad columns (9..11) for the next problem of the
        // persistent loop (the constrained path of this problem reloads whole tiles, pads included, from the workspace)
        if (rezero_pads)
            for (int i = lane; i < 20; i += GL) { T* q = sm + L.oT0 + i * TLD + 9; q[0] = T(0); q[1] = T(0); q[2] = T(0); }
        viol = grp_any(mask, v_l);
        bad = grp_any(mask, b_l);
        nact = n_l;
    } else {
        constexpr int kPf = 2;
        T buf[kPf][12];
#pragma unroll
        for (int u = 0; u < kPf; u++) {
            if (u < N) {
                const T* r = rec + (long long)u * FW_REC;
                Vec4<T>::ld(r, buf[u][0], buf[u][1], buf[u][2], buf[u][3]);
                Vec4<T>::ld(r + 4, buf[u][4], buf[u][5], buf[u][6], buf[u][7]);
                Vec4<T>::ld(r + 8, buf[u][8], buf[u][9], buf[u][10], buf[u][11]);
            }
        }
        for (int k0 = 0; k0 < N; k0 += kPf) {
#pragma unroll
            for (int u = 0; u < kPf; u++) {
                const int k = k0 + u;
                if (k < N) {
                    T dz;
                    stage(k, buf[u], dz);
                    if (lane < 14) sDz[k * 16 + lane] = accum ? sDz[k * 16 + lane] + dz : dz;
                    if (k + kPf < N) {
                        const T* r = rec + (long long)(k + kPf) * FW_REC;
                        Vec4<T>::ld(r, buf[u][0], buf[u][1], buf[u][2], buf[u][3]);
                        Vec4<T>::ld(r + 4, buf[u][4], buf[u][5], buf[u][6], buf[u][7]);
                        Vec4<T>::ld(r + 8, buf[u][8], buf[u][9], buf[u][10], buf[u][11]);
                    }
                }
            }
        }
        if (isx) sDz[N * 16 + lane] = accum ? sDz[N * 16 + lane] + z : z;
    }
    __syncwarp(mask);
}

#ifdef NDP_RTI_PROF
// diagnostics build: globaltimer stamps of every CTA's first pass (thread 0): entry, record staged, cost records,
// backward sweep, forward sweep, stores
static __device__ unsigned long long g_rti_prof[2048 * 8];
static __device__ unsigned long long g_con_prof[2 * 8192 + 2];  // constrained kernel: [2 p] start / [2 p + 1] end of problem p; [16384] first CTA entry
#define RTI_GT(i) do { if (threadIdx.x == 0 && blockIdx.x < 2048 && base == (int)blockIdx.x * ppc) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); g_rti_prof[blockIdx.x * 8 + (i)] = t_; } } while (0)
// interior-point phases of the constrained kernel (clock64 sums, lane 0 of the group): g_con_prof[100 + i]
#define IPM_T0() long long ipm_t0_ = clock64()
#define IPM_T(i) do { if (lane == 0) { const long long t_ = clock64(); g_con_prof[100 + (i)] += (unsigned long long)(t_ - ipm_t0_); ipm_t0_ = t_; } } while (0)
#define IPM_CNT(i, v) do { if (lane == 0) g_con_prof[100 + (i)] += (unsigned long long)(v); } while (0)
#else
#define IPM_T0() do { } while (0)
#define IPM_T(i) do { } while (0)
#define IPM_CNT(i, v) do { } while (0)
#define RTI_GT(i) do { } while (0)
#endif
enum { IPM_LL = 0, IPM_LU, IPM_TL, IPM_TU, IPM_CL, IPM_CU, IPM_ACT };

constexpr int AS_HIST = 6;  // active-set hashes remembered for the cycle test

// Constrained QP of one problem (the unconstrained step left its box): kept out of line so that the nominal
// path of rti_step_kernel keeps its register allocation (nothing but the problem loop's own state is live
// across the call): solves the QP, writes the new iterate, u0, status and statistics of the problem.
//
// Route: primal-dual active-set rounds with EXACT pins -- inputs and velocity components alike (backward_stage,
// kBar == 2) -- seeded by the bounds the unconstrained step violates; one Riccati factorisation per round, the active
// set lives in registers (two bit masks per owning lane: lanes 10..13 the inputs, lanes 3..5 the velocities).  A
// fixed point of the rounds satisfies the KKT conditions of the QP, i.e. it is the solution the reference's
// interior-point method converges to, without a barrier floor.  The plain update (release every wrong-signed
// multiplier, add every violated bound) can cycle; a repeated set (hash test) switches to a damped update (add every
// violated bound; release only the single worst multiplier, and only once nothing is violated).  If as_first_max
// rounds find no fixed point: Mehrotra predictor-corrector IPM (HPIPM's algorithm) for an active-set estimate, then
// the rounds again from that estimate.
template <typename T, int kN>
__device__ __forceinline__ void constrained_qp(const RtiCfg<T>& c, int lane, unsigned mask, T* sm, T* ws, T* rec, bool have_tiles, const T* sTriv,
                                               T dx0, T lo, T hi, T* gX, T* gU, T* gu0, int32_t* g_status, int32_t* g_status2,
                                               int32_t* g_stats, unsigned long long* g_as, bool keep_set, int n_fact0) {
    const int N = (kN > 0) ? kN : c.N;
    const SmemLayout L(N);
    const WsLayout WL(N);
    T* sX = sm + L.oX;
    T* sU = sm + L.oU;
    T* sDz = sm + L.oDz;
    const bool isx = lane < 10, isu = (lane >= 10 && lane < 14), isv = (lane >= 3 && lane < 6);
    auto iter_at = [&](int k) -> T { return isx ? sX[k * NX + lane] : (isu ? sU[k * NU + (lane - 10)] : T(0)); };
    auto has_box = [&](int k) -> bool { return (isu && k < N) || (isv && k >= 1 && k < N); };
    // The problem arrives staged in shared memory (iterate, cost records) with the set to start from in g_as; nothing
    // has been linearised in THIS kernel yet: the first round integrates and factorises with that set pinned.
    int status = 0, n_fact = n_fact0, n_ipm = 0, n_pol = 0;
    bool viol = false, bad = false;
    int nact_l = 0;
    T* wI = sm + L.oIpm;  // [field][k][8]: box lane bl = as_owner(lane)
    const int bl = (isu || isv) ? as_owner(lane) : 7;
    T* wZ = ws + WL.oZc;
    T* bD = sm + L.oIpm + 6 * N * 8;  // [k <= N][8], slot bl
    T* bG = bD + (N + 1) * 8;
    const int FS = N * 8;  // field stride of the interior-point vectors
    bool ipm_ok = false, pol_ok = false;
    ActiveSet<T> as;
    as.lo = lo; as.hi = hi;
    // [A B b] tiles of this iterate are in `rec`: left there by the nominal kernel's sweep when its workspace slot was
    // not reused for a later problem (have_tiles), else produced by the first sweep here
    bool lin_done = have_tiles;
    // a released variable that lands within a few ulps of its bound is not a violation (it would be re-pinned and
    // released for ever)
    const T feas_eps = (sizeof(T) == 4 ? T(5e-7) : T(1e-13)) * fmax(T(1), fmax(fabs(lo), fabs(hi)));

    // ---- primal-dual active-set rounds from the set in `as` ----
    auto pdas = [&](int max_rounds) -> bool {
        unsigned long long hist[AS_HIST];
#pragma unroll
        for (int q = 0; q < AS_HIST; q++) hist[q] = 0ull;
        bool damped = false;
        int k_top = N - 1;  // first round of a call: full sweep (nothing saved yet, or saved under other multipliers)
        for (int round = 0; round < max_rounds; round++) {
            // cycle test on a hash of the whole set
            {
                unsigned long long h = (as.lo_m.w0 * 0x9E3779B97F4A7C15ull) ^ (as.hi_m.w0 * 0xC2B2AE3D27D4EB4Full) ^
                                       (as.lo_m.w1 * 0x165667B19E3779F9ull) ^ (as.hi_m.w1 * 0x27D4EB2F165667C5ull);
                h *= (unsigned long long)(2 * lane + 1);
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) h += __shfl_xor_sync(mask, h, o, GL);
                h |= 1ull;
#pragma unroll
                for (int q = 0; q < AS_HIST; q++) damped |= (hist[q] == h);
#pragma unroll
                for (int q = AS_HIST - 1; q > 0; q--) hist[q] = hist[q - 1];
                hist[0] = h;
            }
            bool fact_ok;
            IPM_T0();
            if (!lin_done) fact_ok = backward_sweep<T, true, 2, true>(c, N, lane, mask, sm, L, ws, WL, rec, sTriv, &as);
            else fact_ok = backward_sweep<T, false, 2, true>(c, N, lane, mask, sm, L, ws, WL, rec, sTriv, &as, false, k_top);
            lin_done = true;
            if (!fact_ok) return false;
            n_fact++;
            n_pol++;
            IPM_T(10);
            IPM_CNT(20, 1);
            IPM_CNT(21, k_top + 1);
            forward_sweep<T, false>(c, N, lane, mask, dx0, sm, L, rec, WL, lo, hi, gX, gU, nullptr, viol, bad, nact_l);
            IPM_T(11);
            bool changed = false, any_viol = false;
            T worst = T(0);  // most wrong-signed multiplier of this lane's pinned bounds (damped mode releases one)
            int worst_k = -1;
            int top_l = -1;  // highest backward stage whose pins this lane changed (a velocity pin of stage k acts at stage k - 1)
            // Multiplier test and set update, spread over ALL 16 lanes.  (It used to run on the seven box-owning lanes, each
            // walking its N stages with two divergent dot-product paths and a ballot per stage: ~950 cycles per stage,
            // 10 us of a lone problem's 35 us round.)  The (stage, variable) items are dealt out -- input m = (lane - 10) & 3
            // at stages (lane >> 2) + 4 t, velocity component a = lane % 3 at stages 1 + lane / 3 + 5 t -- every lane
            // evaluates its items against a copy of the owner's masks and leaves one action code per item in shared
            // memory (the interior-point scratch fields are dead while the rounds run); the owners then apply the codes to
            // their register-held masks.  Same arithmetic per item, same resulting sets.
            {
                // (12 N elements over the fields CL | CU, 16 N: scratch of one interior-point iteration, rebuilt by the next)
                T* sNu = wI + IPM_CL * FS;  // [k][4]: multipliers of the velocity components of stage k + 1 pinned in this sweep
                T* sAct = sNu + 4 * N;      // [k][8]: per owner slot (as_owner order): 0 keep, 1 pin at lo, 2 pin at hi, 3 release
                const int um = (lane - 10) & 3, ug = lane >> 2, uo = 10 + um;  // this lane's input items: owner lane uo
                const int va = lane % 3, vg = lane / 3, vo = 3 + va;           // velocity items (lane 15: none)
                StageMask u_lo, u_hi, v_lo, v_hi, v_any;
                u_lo.w0 = __shfl_sync(mask, as.lo_m.w0, uo, GL); u_lo.w1 = __shfl_sync(mask, as.lo_m.w1, uo, GL);
                u_hi.w0 = __shfl_sync(mask, as.hi_m.w0, uo, GL); u_hi.w1 = __shfl_sync(mask, as.hi_m.w1, uo, GL);
                v_lo.w0 = __shfl_sync(mask, as.lo_m.w0, vo, GL); v_lo.w1 = __shfl_sync(mask, as.lo_m.w1, vo, GL);
                v_hi.w0 = __shfl_sync(mask, as.hi_m.w0, vo, GL); v_hi.w1 = __shfl_sync(mask, as.hi_m.w1, vo, GL);
                {
                    const unsigned long long p0 = isv ? (as.lo_m.w0 | as.hi_m.w0) : 0ull, p1 = isv ? (as.lo_m.w1 | as.hi_m.w1) : 0ull;
                    v_any.w0 = __shfl_sync(mask, p0, 3, GL) | __shfl_sync(mask, p0, 4, GL) | __shfl_sync(mask, p0, 5, GL);
                    v_any.w1 = __shfl_sync(mask, p1, 3, GL) | __shfl_sync(mask, p1, 4, GL) | __shfl_sync(mask, p1, 5, GL);
                }
                const T ulo = __shfl_sync(mask, lo, uo, GL), uhi = __shfl_sync(mask, hi, uo, GL);
                const T vlo = __shfl_sync(mask, lo, vo, GL), vhi = __shfl_sync(mask, hi, vo, GL);
                const T ueps = __shfl_sync(mask, feas_eps, uo, GL), veps = __shfl_sync(mask, feas_eps, vo, GL);
                int worst_key = 0x7fffffff;  // (owner lane << 8 | stage) of this lane's most wrong-signed multiplier
                // rows of the pinned items: requested now, read below (they were written through L2 by the backward sweep)
                for (int k1 = 1 + vg; k1 < N && lane < 15; k1 += 5)
                    if (v_lo.test(k1) || v_hi.test(k1))
                        pf_l1(ws + WL.oTv + (long long)(k1 - 1) * 36 + va * 12);
                for (int k = ug; k < N; k += 4)
                    if (u_lo.test(k) || u_hi.test(k)) {
                        const T* hr = ws + WL.oHrow + (k * 4 + um) * 16;
                        pf_l1(hr);
                        if (sizeof(T) == 8) pf_l1(hr + 8);
                        if (k + 1 < N && v_any.test(k + 1)) pf_l1(rec + ((long long)k * 14 + 3) * TLD + 4);
                    }
                // velocity items
                for (int k1 = 1 + vg; k1 < N && lane < 15; k1 += 5) {
                    const int k = k1 - 1;
                    const bool at_lo = v_lo.test(k1), at_hi = v_hi.test(k1);
                    T act = T(0), nu = T(0);
                    if (at_lo || at_hi) {
                        const T* tv = ws + WL.oTv + (long long)k * 36 + va * 12;
                        nu = tv[10];
#pragma unroll
                        for (int i = 0; i < 10; i++) nu += tv[i] * sDz[k * 16 + i];
                        nu = -nu;
                        const T lam = at_hi ? nu : -nu;  // >= 0 at a KKT point
                        if (lam < T(0)) {
                            if (!damped) act = T(3);
                            else if (lam < worst || (lam == worst && ((vo << 8) | k1) < worst_key)) { worst = lam; worst_key = (vo << 8) | k1; }
                        }
                    } else {
                        const T it_v = sX[k1 * NX + 3 + va], zn = sDz[k1 * 16 + 3 + va];
                        if (zn > vhi - it_v + veps) act = T(2);
                        else if (zn < vlo - it_v - veps) act = T(1);
                    }
                    sNu[k * 4 + va] = nu;
                    sAct[k1 * 8 + va] = act;
                }
                __syncwarp(mask);
                // input items
                for (int k = ug; k < N; k += 4) {
                    const bool at_lo = u_lo.test(k), at_hi = u_hi.test(k);
                    T act = T(0);
                    if (at_lo || at_hi) {
                        // multiplier of the pinned input from the un-penalised row of [Hux Guu | g_u] (+ the pinned
                        // velocity rows of this stage's mixed constraint)
                        const T* hr = ws + WL.oHrow + (k * 4 + um) * 16;
                        T gq = hr[14];
#pragma unroll
                        for (int i = 0; i < 14; i++) gq += hr[i] * sDz[k * 16 + i];
                        if (k + 1 < N && v_any.test(k + 1)) {
                            const T* eb = rec + ((long long)k * 14 + 3) * TLD + 4 + um;  // (E B)(a, m)
                            gq += eb[0] * sNu[k * 4] + eb[TLD] * sNu[k * 4 + 1] + eb[2 * TLD] * sNu[k * 4 + 2];
                        }
                        const T lam = at_hi ? -gq : gq;  // >= 0 at a KKT point
                        if (lam < T(0)) {
                            if (!damped) act = T(3);
                            else if (lam < worst || (lam == worst && ((uo << 8) | k) < worst_key)) { worst = lam; worst_key = (uo << 8) | k; }
                        }
                    } else {
                        const T it_v = sU[k * NU + um], zn = sDz[k * 16 + 10 + um];
                        if (zn > uhi - it_v + ueps) act = T(2);
                        else if (zn < ulo - it_v - ueps) act = T(1);
                    }
                    sAct[k * 8 + 3 + um] = act;
                }
                __syncwarp(mask);
                // the owners apply the codes of their variable (a velocity pin of stage k acts at backward stage k - 1)
                if (isu || isv) {
                    const int slot = as_owner(lane);
#pragma unroll 4
                    for (int k = isv ? 1 : 0; k < N; k++) {
                        const T act = sAct[k * 8 + slot];
                        if (act != T(0)) {
                            if (act == T(1)) { as.lo_m.set(k); any_viol = true; }
                            else if (act == T(2)) { as.hi_m.set(k); any_viol = true; }
                            else { as.lo_m.clear(k); as.hi_m.clear(k); }
                            changed = true;
                            top_l = isv ? k - 1 : k;
                        }
                    }
                }
                if (damped) {
                    // the problem's single worst multiplier: its owner releases it below, once nothing is violated
                    const T w_all = grp_min<T>(worst, mask);
                    int key = (worst == w_all && w_all < T(0)) ? worst_key : 0x7fffffff;
#pragma unroll
                    for (int o = 8; o >= 1; o >>= 1) {
                        const int other = __shfl_xor_sync(mask, key, o, GL);
                        key = other < key ? other : key;
                    }
                    worst = w_all;
                    worst_k = (key != 0x7fffffff && (key >> 8) == lane) ? (key & 255) : -1;
                }
                __syncwarp(mask);
            }
            any_viol = __any_sync(mask, any_viol);
            if (damped && !any_viol && worst_k >= 0) {
                // release the single worst multiplier of the problem (this lane owns it)
                as.lo_m.clear(worst_k); as.hi_m.clear(worst_k); changed = true;
                top_l = isv ? worst_k - 1 : worst_k;
            }
            changed = __any_sync(mask, changed);
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) {
                const int other = __shfl_xor_sync(mask, top_l, o, GL);
                top_l = other > top_l ? other : top_l;
            }
            k_top = top_l;
            __syncwarp(mask);
            IPM_T(12);
            if (!changed) {
                // pinned variables sit exactly on their bound
                if (isu || isv)
                    for (int k = 0; k < N; k++) {
                        const T it_v = iter_at(k);
                        if (as.lo_m.test(k)) sDz[k * 16 + lane] = lo - it_v;
                        if (as.hi_m.test(k)) sDz[k * 16 + lane] = hi - it_v;
                    }
                return true;
            }
        }
        return false;
    };

    // Phase 0: rounds from the handed-over set (the bounds the unconstrained step violates, or the previous solve's final set)
    if (isu || isv) {
        const unsigned long long* p = g_as + as_owner(lane) * 4;
        as.lo_m.w0 = p[0]; as.lo_m.w1 = p[1]; as.hi_m.w0 = p[2]; as.hi_m.w1 = p[3];
    }
    __syncwarp(mask);
    if (c.as_first_max > 0) pol_ok = pdas(c.as_first_max);

    if (!pol_ok) {
        // ================= Mehrotra IPM on the Riccati kernel: active-set estimate for the rounds =================
        // (t < lambda marks a bound as active).  The rounds are tried from the estimate after convergence and, since only
        // the active set has to be right, not the barrier iterate, every 3rd iteration before that.
        auto rounds_from_ipm = [&](int max_rounds) -> bool {
            as.lo_m = StageMask(); as.hi_m = StageMask();
            for (int k = 0; k < N; k++)
                if (has_box(k)) {
                    const int ei = k * 8 + bl;
                    if (wI[IPM_TL * FS + ei] < wI[IPM_LL * FS + ei]) as.lo_m.set(k);
                    else if (wI[IPM_TU * FS + ei] < wI[IPM_LU * FS + ei]) as.hi_m.set(k);
                }
            __syncwarp(mask);
            return pdas(max_rounds);
        };
        if (!lin_done) {
            // warm start without a usable guess: linearise with an empty set first
            as.lo_m = StageMask(); as.hi_m = StageMask();
            if (!backward_sweep<T, true, 2, false>(c, N, lane, mask, sm, L, ws, WL, rec, sTriv, &as)) status = 4;
            lin_done = true;
            n_fact++;
        }
        for (int k = 0; k <= N; k++) {
            if (lane < 8) { bD[k * 8 + lane] = T(0); bG[k * 8 + lane] = T(0); }
            wZ[k * 16 + lane] = (k == 0 && isx) ? dx0 : T(0);
        }
        __syncwarp(mask);
        int nb_l = 0;
        for (int k = 0; k < N; k++)
            if (has_box(k)) {
                const T it_v = iter_at(k);
                const T lb = lo - it_v, ub = hi - it_v;
                const T tl = fmax(-lb, c.t_floor), tu = fmax(ub, c.t_floor);
                wI[IPM_TL * FS + k * 8 + bl] = tl;
                wI[IPM_TU * FS + k * 8 + bl] = tu;
                wI[IPM_LL * FS + k * 8 + bl] = tdiv<T>(c.mu0, tl);
                wI[IPM_LU * FS + k * 8 + bl] = tdiv<T>(c.mu0, tu);
                nb_l++;
            }
        const T inv_m = T(1) / (T(2) * grp_sum<T>((T)nb_l, mask));
        __syncwarp(mask);
        T res_lin = T(1), mu_prev = T(1e30), mu = T(0);
        // a Cholesky pivot lost to the barrier weights (fp32 with active velocity bounds) ends the interior-point
        // iteration, not the solve: its active-set estimate still goes to the rounds
        bool ipm_broken = false;
        int it = 0;
        // element-wise passes over this lane's boxed variables: stages 0 (inputs) / 1 (velocities) .. N - 1.  Branch-free
        // bodies, unrolled, so that the loads and reciprocals of four stages overlap (these serial loops were ~30 % of a
        // lone problem's interior-point iteration)
        auto box_loop = [&](auto&& body) {
            if (isu || isv) {
#pragma unroll 4
                for (int k = isv ? 1 : 0; k < N; k++) body(k);
            }
        };
        IPM_T0();
        for (it = 0; status == 0 && it <= c.ipm_max_iter; it++) {
            if (it >= 3 && (it % 3) == 0 && res_lin <= T(1e-1) && c.polish_max > 0) {
                IPM_T(7);
                const bool r_ok = rounds_from_ipm(3);
                IPM_T(8);
                if (r_ok) { pol_ok = true; break; }
            }
            // complementarity, affine barrier terms
            IPM_T(0);
            T mu_l = T(0);
            box_loop([&](int k) {
                const int ei = k * 8 + bl;
                const T it_v = iter_at(k);
                const T lb = lo - it_v, ub = hi - it_v;
                const T tl = wI[IPM_TL * FS + ei], tu = wI[IPM_TU * FS + ei];
                const T ll = wI[IPM_LL * FS + ei], lu = wI[IPM_LU * FS + ei];
                mu_l += ll * tl + lu * tu;
                const T gl = tdiv<T>(ll, tl), gu = tdiv<T>(lu, tu);
                bD[ei] = gl + gu;
                bG[ei] = (-gu * ub + lu) - (gl * lb + ll);
            });
            mu = grp_sum<T>(mu_l, mask) * inv_m;
            if (it > 0 && res_lin <= c.tol_res && mu < c.tol_mu) { ipm_ok = true; break; }
            if (it > 3 && res_lin <= c.tol_res && mu > T(0.9) * mu_prev && mu < T(1e-2)) { ipm_ok = true; break; }
            if (!(mu == mu)) { ipm_broken = true; break; }  // NaN
            if (it == c.ipm_max_iter) break;
            mu_prev = mu;
            __syncwarp(mask);
            // ---- predictor ----
            IPM_T(1);
            if (!backward_sweep<T, false, 1, false>(c, N, lane, mask, sm, L, ws, WL, rec, sTriv, nullptr)) { ipm_broken = true; break; }
            n_fact++;
            IPM_T(2);
            forward_sweep<T, false>(c, N, lane, mask, dx0, sm, L, rec, WL, lo, hi, gX, gU, nullptr, viol, bad, nact_l);
            IPM_T(3);
            T amax = T(1e30);
            box_loop([&](int k) {
                const int e = k * 16 + lane, ei = k * 8 + bl;
                const T it_v = iter_at(k);
                const T lb = lo - it_v, ub = hi - it_v;
                const T zn = sDz[e];
                const T tl = wI[IPM_TL * FS + ei], tu = wI[IPM_TU * FS + ei];
                const T ll = wI[IPM_LL * FS + ei], lu = wI[IPM_LU * FS + ei];
                const T dtl = (zn - lb) - tl, dtu = (ub - zn) - tu;
                const T dll = -tdiv<T>(ll, tl) * dtl - ll, dlu = -tdiv<T>(lu, tu) * dtu - lu;
                wI[IPM_CL * FS + ei] = dll * dtl;
                wI[IPM_CU * FS + ei] = dlu * dtu;
                if (dtl < T(0)) amax = fmin(amax, tdiv<T>(-tl, dtl));
                if (dtu < T(0)) amax = fmin(amax, tdiv<T>(-tu, dtu));
                if (dll < T(0)) amax = fmin(amax, tdiv<T>(-ll, dll));
                if (dlu < T(0)) amax = fmin(amax, tdiv<T>(-lu, dlu));
            });
            const T a_aff = fmin(grp_min<T>(amax, mask), T(1));
            T mua_l = T(0);
            box_loop([&](int k) {
                const int e = k * 16 + lane, ei = k * 8 + bl;
                const T it_v = iter_at(k);
                const T lb = lo - it_v, ub = hi - it_v;
                const T zn = sDz[e];
                const T tl = wI[IPM_TL * FS + ei], tu = wI[IPM_TU * FS + ei];
                const T ll = wI[IPM_LL * FS + ei], lu = wI[IPM_LU * FS + ei];
                const T dtl = (zn - lb) - tl, dtu = (ub - zn) - tu;
                const T dll = -tdiv<T>(ll, tl) * dtl - ll, dlu = -tdiv<T>(lu, tu) * dtu - lu;
                mua_l += (ll + a_aff * dll) * (tl + a_aff * dtl) + (lu + a_aff * dlu) * (tu + a_aff * dtu);
            });
            const T mu_aff = grp_sum<T>(mua_l, mask) * inv_m;
            const T sg = tdiv<T>(mu_aff, mu);
            const T sigma_mu = sg * sg * sg * mu;
            // ---- corrector ----
            box_loop([&](int k) {
                const int ei = k * 8 + bl;
                const T tl = wI[IPM_TL * FS + ei], tu = wI[IPM_TU * FS + ei];
                const T cl = wI[IPM_CL * FS + ei], cu = wI[IPM_CU * FS + ei];
                // only the CHANGE of the barrier gradient against the predictor's: the weights bD stay, so the corrector
                // re-uses the predictor's factorisation (delta_backward) and adds its step to the affine one
                bG[ei] = tdiv<T>(sigma_mu - cu, tu) - tdiv<T>(sigma_mu - cl, tl);
            });
            __syncwarp(mask);
            IPM_T(4);
            delta_backward<T>(c, N, lane, mask, ws, WL, rec, sTriv, bG);
            IPM_T(5);
            forward_sweep<T, false>(c, N, lane, mask, T(0), sm, L, rec, WL, lo, hi, gX, gU, nullptr, viol, bad, nact_l, true, true, true);
            IPM_T(6);
            amax = T(1e30);
            box_loop([&](int k) {
                const int e = k * 16 + lane, ei = k * 8 + bl;
                const T it_v = iter_at(k);
                const T lb = lo - it_v, ub = hi - it_v;
                const T zn = sDz[e];
                const T tl = wI[IPM_TL * FS + ei], tu = wI[IPM_TU * FS + ei];
                const T ll = wI[IPM_LL * FS + ei], lu = wI[IPM_LU * FS + ei];
                const T cl = wI[IPM_CL * FS + ei], cu = wI[IPM_CU * FS + ei];
                const T dtl = (zn - lb) - tl, dtu = (ub - zn) - tu;
                const T dll = tdiv<T>(sigma_mu - cl - ll * dtl, tl) - ll;
                const T dlu = tdiv<T>(sigma_mu - cu - lu * dtu, tu) - lu;
                if (dtl < T(0)) amax = fmin(amax, tdiv<T>(-tl, dtl));
                if (dtu < T(0)) amax = fmin(amax, tdiv<T>(-tu, dtu));
                if (dll < T(0)) amax = fmin(amax, tdiv<T>(-ll, dll));
                if (dlu < T(0)) amax = fmin(amax, tdiv<T>(-lu, dlu));
            });
            const T alpha = fmin(T(1), T(0.995) * grp_min<T>(amax, mask));
            box_loop([&](int k) {
                const int e = k * 16 + lane, ei = k * 8 + bl;
                const T it_v = iter_at(k);
                const T lb = lo - it_v, ub = hi - it_v;
                const T zn = sDz[e];
                const T tl = wI[IPM_TL * FS + ei], tu = wI[IPM_TU * FS + ei];
                const T ll = wI[IPM_LL * FS + ei], lu = wI[IPM_LU * FS + ei];
                const T cl = wI[IPM_CL * FS + ei], cu = wI[IPM_CU * FS + ei];
                const T dtl = (zn - lb) - tl, dtu = (ub - zn) - tu;
                const T dll = tdiv<T>(sigma_mu - cl - ll * dtl, tl) - ll;
                const T dlu = tdiv<T>(sigma_mu - cu - lu * dtu, tu) - lu;
                wI[IPM_TL * FS + ei] = fmax(tl + alpha * dtl, c.t_min);
                wI[IPM_TU * FS + ei] = fmax(tu + alpha * dtu, c.t_min);
                wI[IPM_LL * FS + ei] = fmax(ll + alpha * dll, c.t_min);
                wI[IPM_LU * FS + ei] = fmax(lu + alpha * dlu, c.t_min);
            });
            if (lane < 14) {
                // (the iterate of the interior-point method stays in the workspace: four independent loads per round trip)
                const int kn = isx ? N + 1 : N;
                for (int k0 = 0; k0 < kn; k0 += 4) {
                    T zv[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) zv[q] = (k0 + q < kn) ? wZ[(k0 + q) * 16 + lane] : T(0);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (k0 + q < kn) wZ[(k0 + q) * 16 + lane] = zv[q] + alpha * (sDz[(k0 + q) * 16 + lane] - zv[q]);
                }
            }
            res_lin *= (T(1) - alpha);
            n_ipm++;
            __syncwarp(mask);
            IPM_T(7);
        }
        if (!pol_ok && status == 0 && c.polish_max > 0 && n_ipm > 0) pol_ok = rounds_from_ipm(c.polish_max);
        if (ipm_broken && !pol_ok) status = 4;
    }
    if (pol_ok && status == 0) {
        // Defect correction for pinned velocities.  Holding a velocity component through the previous stage's inputs is a
        // stiff feedback (gains ~ 1e2, cost-to-go ~ 1e4 against stage weights ~ 10), and the Schur complement of the next
        // stage loses about three digits to it: fp32 ends 2e-4 off the solution, fp64 5e-13.  One more sweep with the
        // same factorisation structure, linearised AT the computed step (iterate advanced in shared memory, stage
        // residuals b zeroed, cost gradient and pinned values re-evaluated there), solves for the small remainder.
        // The same sweep is taken by heavily saturated problems (a dozen or more pinned variables), whose fp32 error
        // otherwise compounds over warm-started steps.
        const bool vown = isv && ((as.lo_m.w0 | as.lo_m.w1 | as.hi_m.w0 | as.hi_m.w1) != 0ull);
        const int n_pin = (int)grp_sum<float>((float)(__popcll(as.lo_m.w0 | as.hi_m.w0) + __popcll(as.lo_m.w1 | as.hi_m.w1)), mask);
        if (sizeof(T) == 4 && (__any_sync(mask, vown) || n_pin >= 12)) {
            T* sY = sm + L.oY;
            if (lane < 14)
                for (int k = 0; k <= N; k++) {
                    if (k == N && !isx) break;
                    const T dz = sDz[k * 16 + lane];
                    if (isx) sX[k * NX + lane] += dz;
                    else sU[k * NU + (lane - 10)] += dz;
                    if (lane < 6 || lane >= 10) sY[k * SYS + lane] += dz;  // residual records are linear in the iterate
                }
            __syncwarp(mask);
            if (backward_sweep<T, false, 2, false>(c, N, lane, mask, sm, L, ws, WL, rec, sTriv, &as, true)) {
                n_fact++;
                forward_sweep<T, false>(c, N, lane, mask, T(0), sm, L, rec, WL, lo, hi, gX, gU, nullptr, viol, bad, nact_l, true, true);
                if (isu || isv)
                    for (int k = 0; k < N; k++) {
                        const T it_v = iter_at(k);
                        if (as.lo_m.test(k)) sDz[k * 16 + lane] = lo - it_v;
                        if (as.hi_m.test(k)) sDz[k * 16 + lane] = hi - it_v;
                    }
            } else {
                // keep the unrefined solution: the step relative to the advanced iterate is zero
                if (lane < 14)
                    for (int k = 0; k <= N; k++) sDz[k * 16 + lane] = T(0);
            }
            __syncwarp(mask);
        }
    }
    if (!pol_ok) {
        // fall back to the interior-point iterate
        if (lane < 14)
            for (int k = 0; k <= N; k++) {
                if (k == N && !isx) break;
                sDz[k * 16 + lane] = wZ[k * 16 + lane];
            }
        if (!ipm_ok && status == 0) status = 4;
    }
    __syncwarp(mask);
    // ---- write back: full step, overwrite the tentative iterate of the unconstrained sweep ----
    bool bad2 = false;
    nact_l = 0;
    StageMask fin_lo, fin_hi;  // variables that end on a bound: the next solve's first guess
    const T as_eps = (sizeof(T) == 4 ? T(1e-5) : T(1e-11)) * (T(1) + fabs(lo) + fabs(hi));
    if (lane < 14)
        for (int k = 0; k <= N; k++) {
            if (k == N && !isx) break;
            const T v = iter_at(k) + sDz[k * 16 + lane];
            bad2 |= !(fabs(v) <= T(1e30));
            if (has_box(k)) nact_l += (v <= lo) + (v >= hi);
            // (it + (bound - it) need not round back to the bound exactly: compare with a few ulps of slack)
            if (has_box(k) && v <= lo + as_eps) fin_lo.set(k);
            else if (has_box(k) && v >= hi - as_eps) fin_hi.set(k);
        }
    bad2 = __any_sync(mask, bad2);
    const int nact = (int)grp_sum<float>((float)nact_l, mask);
    if (bad2) status = 1;
    // A failed QP leaves the iterate as it was (acados SQP_RTI returns without updating it), so the next solve is
    // not warm-started from a poisoned point; u0 is then the previous first input.
    if (status == 0 && lane < 14)
        for (int k = 0; k <= N; k++) {
            if (k == N && !isx) break;
            const T v = iter_at(k) + sDz[k * 16 + lane];
            if (isx) sX[k * NX + lane] = v;
            else sU[k * NU + (lane - 10)] = v;
        }
    __syncwarp(mask);
    for (int i = lane; i < (N + 1) * NX; i += GL) gX[i] = sX[i];
    for (int i = lane; i < N * NU; i += GL) gU[i] = sU[i];
    if (gu0 && lane < NU) gu0[lane] = sU[lane];
    if (isu || isv) {
        const bool keep = keep_set && (status == 0);
        unsigned long long* p = g_as + as_owner(lane) * 4;
        p[0] = keep ? fin_lo.w0 : 0ull; p[1] = keep ? fin_lo.w1 : 0ull; p[2] = keep ? fin_hi.w0 : 0ull; p[3] = keep ? fin_hi.w1 : 0ull;
    }
    if (lane == 0) {
        *g_status = status;
        if (g_status2) *g_status2 = status;
        g_stats[0] = n_fact;
        g_stats[1] = n_ipm;
        g_stats[2] = n_pol;
        g_stats[3] = nact;
    }
    __syncwarp(mask);
}

constexpr int RTI_CTA = 64;  // threads per CTA of both launches (4 problems)
constexpr int QUEUE_SWEPT = 1 << 30;  // queue entry flag: the unconstrained sweep of this solve has run (statistics)
constexpr int QUEUE_HARD = 10;        // violated bounds (or stored pins) from which a problem joins the front of the queue

// Stage one problem record in shared memory with asynchronous copies (one wait): iterate, then either the stored
// yref / p (kFused == false) or -- controller.update()'s 42 solver.set calls, nmpc_body_rate_ctl.py:95-104 -- yref_k =
// [xr_k; ur_k], p_k = [xr_k[6:10]; f_k] built from (xr, ur, f) and persisted to yref_w / par_w as if set stage by stage.
template <typename T>
__device__ __forceinline__ void stage_problem(const RtiArgs<T>& a, int N, const SmemLayout& L, int lane, unsigned mask, T* sm, int prob,
                                              bool fused, bool persist = true) {
    constexpr int E2 = 8 / (int)sizeof(T);  // elements per 8-byte copy (records are 8-byte aligned)
    T* sX = sm + L.oX;
    T* sU = sm + L.oU;
    T* sY = sm + L.oY;
    T* sPar = sm + L.oPar;
    const T* gX = a.X + (size_t)prob * (N + 1) * NX;
    const T* gU = a.U + (size_t)prob * N * NU;
    for (int i = lane; i < (N + 1) * NX / E2; i += GL) cp_async<8>(sX + E2 * i, gX + E2 * i);
    for (int i = lane; i < N * NU / E2; i += GL) cp_async<8>(sU + E2 * i, gU + E2 * i);
    if (!fused) {
        const T* gY = a.yref + (size_t)prob * (N + 1) * NYS;
        const T* gP = a.par + (size_t)prob * (N + 1) * NPS;
        for (int i = lane; i < (N + 1) * (NYS / E2); i += GL) {
            const int k = i / (NYS / E2), q = i - k * (NYS / E2);
            cp_async<8>(sY + k * SYS + E2 * q, gY + k * NYS + E2 * q);
        }
        for (int i = lane; i < (N + 1) * NPS / E2; i += GL) cp_async<8>(sPar + E2 * i, gP + E2 * i);
    } else {
        const T* gxr = a.xr + (size_t)prob * (N + 1) * NX;
        const T* gur = a.ur + (size_t)prob * N * NU;
        for (int i = lane; i < (N + 1) * (NX / E2); i += GL) {
            const int k = i / (NX / E2), q = (i - k * (NX / E2)) * E2;
            cp_async<8>(sY + k * SYS + q, gxr + k * NX + q);
            if (q >= 6) cp_async<8>(sPar + k * NPS + q - 6, gxr + k * NX + q);
        }
        for (int i = lane; i < N * (NU / E2); i += GL) {
            const int k = i / (NU / E2), q = (i - k * (NU / E2)) * E2;
            cp_async<8>(sY + k * SYS + NX + q, gur + k * NU + q);
        }
        if (lane < NU) sY[N * SYS + NX + lane] = T(0);
        if (!a.f) {
            for (int i = lane; i < (N + 1) * 4; i += GL) sPar[(i >> 2) * NPS + 4 + (i & 3)] = T(0);
        }
    }
    for (int i = lane; i < (N + 1) * 2; i += GL) sY[(i >> 1) * SYS + NYS + (i & 1)] = T(0);
    if (fused && a.f) {
        // The forces come last: when this kernel was launched as a programmatic dependent of the kernel that
        // produces them (ndp_update_ex, NDP_UPDATE_F_FROM_PREVIOUS_KERNEL), everything above -- which only
        // reads older data -- has run under that kernel's tail; wait for its completion here (a no-op for an
        // ordinary launch).
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const T* gf = a.f + (size_t)prob * (N + 1) * 3;
        for (int i = lane; i < (N + 1) * 3; i += GL) {
            const int k = i / 3, m = i - k * 3;
            cp_async<(int)sizeof(T)>(sPar + k * NPS + 4 + m, gf + i);
        }
        for (int k = lane; k <= N; k += GL) sPar[k * NPS + 7] = T(0);
    }
    cp_async_wait_all();
    __syncwarp(mask);
    if (fused && persist) {
        // persist yref / p as if set stage by stage (a later plain solve, a get, or the constrained kernel sees them)
        T* wY = a.yref_w + (size_t)prob * (N + 1) * NYS;
        T* wP = a.par_w + (size_t)prob * (N + 1) * NPS;
        // 8-byte stores; two stage records per pass (lanes 0..6 and 8..14 carry the NYS / E2 vectors of one stage each)
        typedef typename std::conditional<sizeof(T) == 4, float2, double>::type V2;
        constexpr int VY = NYS / E2, VS = SYS / E2;
        static_assert(VY <= 8 || sizeof(T) == 8, "persist: one half group per stage record");
        if (sizeof(T) == 4) {
            const int q = lane & 7;
            for (int k = (lane >> 3); k <= N; k += 2)
                if (q < VY) reinterpret_cast<V2*>(wY)[k * VY + q] = reinterpret_cast<const V2*>(sm + L.oY)[k * VS + q];
        } else {
            for (int k = 0; k <= N; k++)
                if (lane < VY) reinterpret_cast<V2*>(wY)[k * VY + lane] = reinterpret_cast<const V2*>(sm + L.oY)[k * VS + lane];
        }
        for (int i = lane; i < (N + 1) * NPS / E2; i += GL) reinterpret_cast<V2*>(wP)[i] = reinterpret_cast<const V2*>(sm + L.oPar)[i];
    }
    if (fused) __syncwarp(mask);  // outside the `persist` test: the two halves of a warp may differ in it, and `mask` may name both
}

// per-lane box of the variable this lane owns (lanes 3..5: v, lanes 10..13: u)
template <typename T>
__device__ __forceinline__ void lane_box(const RtiCfg<T>& c, int lane, T& lo, T& hi) {
    lo = T(-1e30); hi = T(1e30);
#pragma unroll
    for (int m = 0; m < 3; m++)
        if (lane == 3 + m) { lo = c.vmin[m]; hi = c.vmax[m]; }
#pragma unroll
    for (int m = 0; m < 4; m++)
        if (lane == 10 + m) { lo = c.umin[m]; hi = c.umax[m]; }
}

// ---- nominal kernel: one RTI step per problem with the unconstrained Riccati sweep; a step that leaves its box is
// handed over to rti_constrained_kernel (queue), with global memory still holding the old iterate ----
// kLat: the latency build (fp32 only) -- same code with half the resident CTAs per SM, i.e. twice the registers,
// which ptxas spends on instruction-level parallelism.  Chosen when the whole batch is resident at that occupancy
// (B <= 148 * 4 * 4 problems).
#ifndef NDP_RTI_CTAS
#define NDP_RTI_CTAS 8
#endif
template <typename T, int kN, bool kLat>
__global__ void __launch_bounds__(RTI_CTA, (sizeof(T) == 4) ? (kLat ? 4 : NDP_RTI_CTAS) : 3) rti_step_kernel(const __grid_constant__ RtiCfg<T> c, const RtiArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = (kN > 0) ? kN : c.N;
    const SmemLayout L(N, true);
    const WsLayout WL(N);
    const int lane = threadIdx.x & 15;
    const int grp = threadIdx.x >> 4;
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);  // this problem's half of the warp
    // the two problems of a warp run the sweeps in lockstep (a half without a problem of its own repeats the last one
    // and stores nothing), so every shuffle / __syncwarp of the sweeps names the whole warp with a compile-time mask
    const unsigned mask = 0xffffffffu;
    const int ppc = blockDim.x >> 4;  // problems per CTA
    T* sTriv = reinterpret_cast<T*>(smem_raw);  // [10][TLD] constant tile of the trivial columns 0..5, shared by the CTA
    T* sm = sTriv + 10 * TLD + (size_t)grp * L.total;
    T* ws = a.ws + (size_t)(blockIdx.x * ppc + grp) * a.ws_stride;
    T* sX = sm + L.oX;
    T* sU = sm + L.oU;
    for (int i = threadIdx.x; i < 10 * TLD; i += blockDim.x) {
        const int r = i / TLD, cc = i - r * TLD;
        sTriv[i] = (cc < 6 && cc == r) ? T(1) : ((cc >= 3 && cc < 6 && cc - 3 == r) ? c.h : T(0));
    }
    for (int i = lane; i < 20 * TLD; i += GL) sm[L.oT0 + i] = T(0);
    __syncthreads();
    T lo, hi;
    lane_box<T>(c, lane, lo, hi);
    const bool isx = lane < 10, isu = (lane >= 10 && lane < 14), isv = (lane >= 3 && lane < 6);

    for (int base = blockIdx.x * ppc; base < a.B; base += gridDim.x * ppc) {
        const bool live = base + grp < a.B;
        const int prob = live ? base + grp : a.B - 1;
        const bool more = base + (int)gridDim.x * ppc < a.B;  // this CTA has another pass: the tiles' zero pads are needed again
        RTI_GT(0);
#ifdef NDP_RTI_PROF
        if (threadIdx.x == 0 && blockIdx.x < 2048 && base == (int)blockIdx.x * ppc) { unsigned sm_; asm volatile("mov.u32 %0, %smid;" : "=r"(sm_)); g_rti_prof[blockIdx.x * 8 + 6] = sm_; }
#endif
        T* gX = a.X + (size_t)prob * (N + 1) * NX;
        T* gU = a.U + (size_t)prob * N * NU;
        const T x0v = isx ? a.x0[(size_t)prob * NX + lane] : T(0);  // requested before the record's copies are waited for
        stage_problem<T>(a, N, L, lane, mask, sm, prob, a.xr != nullptr, live);
        RTI_GT(1);
        // a set left by the previous solve of this problem (active_set_warm): straight to the constrained kernel
        unsigned long long* g_as = a.as_store + (size_t)prob * (AS_OWNERS * 4);
        unsigned long long as_any = 0ull;
        if (a.as_warm && (isu || isv)) {
            const ulonglong2 m0 = *reinterpret_cast<const ulonglong2*>(g_as + as_owner(lane) * 4);
            const ulonglong2 m1 = *reinterpret_cast<const ulonglong2*>(g_as + as_owner(lane) * 4 + 2);
            as_any = m0.x | m0.y | m1.x | m1.y;
        }
        const bool warm = grp_any(mask, as_any != 0ull);
        bool ok = true, viol = false, bad = false;
        int nact_l = 0;
        if (__any_sync(0xffffffffu, !warm)) {  // a warm half rides along with its neighbour's sweeps (results unused)
            cost_records<T>(N, lane, sm + L.oY, sX, sU, sm + L.oPar);
            __syncwarp(mask);
            const T dx0 = isx ? x0v - sX[lane] : T(0);
            RTI_GT(2);
            // ---- preparation + unconstrained feedback; the step is accepted on the fly ----
            ok = backward_sweep<T, true, 0, false>(c, N, lane, mask, sm, L, ws, WL, ws, sTriv, nullptr);
            RTI_GT(3);
            forward_sweep<T, true>(c, N, lane, mask, dx0, sm, L, ws, WL, lo, hi, gX, gU, nullptr, viol, bad, nact_l,
                                   more);
        }
        const int nact = (int)grp_sum<float>((float)nact_l, mask);
        RTI_GT(4);
        if (!live) {
            // nothing of its own to store
        } else if (warm || (ok && viol && !bad)) {
            if (!warm && (isu || isv)) {
                // the bounds the unconstrained step violates seed the active-set rounds
                StageMask m_lo, m_hi;
                for (int k = isv ? 1 : 0; k < N; k++) {
                    const T v = isu ? sU[k * NU + (lane - 10)] : sX[k * NX + lane];
                    if (v < lo) m_lo.set(k);
                    else if (v > hi) m_hi.set(k);
                }
                unsigned long long* p = g_as + as_owner(lane) * 4;
                p[0] = m_lo.w0; p[1] = m_lo.w1; p[2] = m_hi.w0; p[3] = m_hi.w1;
            }
            // Hardest first: the launch of the constrained kernel lasts as long as its slowest problem, and the number of
            // bounds the unconstrained step violates (or the size of the stored set) predicts the sweeps a problem will
            // need (correlation 0.8 on the stress set; every problem that went on to the interior-point fallback had 10+).
            // Those problems fill the queue from the front, the others from the back; the consumer walks front to back.
            int hardness = nact;
            if (warm) hardness = (int)grp_sum<float>((float)__popcll(as_any), hmask);
            __syncwarp(hmask);
            if (lane == 0) {
                __threadfence();
                const bool hard = hardness >= QUEUE_HARD;
                const int slot = atomicAdd(a.qctl + (hard ? 3 : 0), 1);
                a.queue[hard ? slot : a.B - 1 - slot] = prob | (warm ? 0 : QUEUE_SWEPT);
            }
        } else {
            int status = 0;
            if (!ok) status = 4;
            if (bad) status = 1;
            if (status == 0) {
                // accepted: the new iterate goes out with 8-byte stores
                constexpr int E2 = 8 / (int)sizeof(T);
                typedef typename std::conditional<sizeof(T) == 4, float2, double>::type V2;
                for (int i = lane; i < (N + 1) * NX / E2; i += GL) reinterpret_cast<V2*>(gX)[i] = reinterpret_cast<const V2*>(sX)[i];
                for (int i = lane; i < N * NU / E2; i += GL) reinterpret_cast<V2*>(gU)[i] = reinterpret_cast<const V2*>(sU)[i];
                if (a.u0 && lane < NU) a.u0[(size_t)prob * NU + lane] = sU[lane];
            } else if (a.u0 && lane < NU) {
                // failed factorisation / NaN step: global memory still holds the previous iterate (acados SQP_RTI does not
                // update it when the QP fails); return the previous first input
                a.u0[(size_t)prob * NU + lane] = gU[lane];
            }
            if (lane == 0) {
                a.status[prob] = status;
                if (a.status2) a.status2[prob] = status;
                a.stats[prob * 4 + 0] = 1;
                a.stats[prob * 4 + 1] = 0;
                a.stats[prob * 4 + 2] = 0;
                a.stats[prob * 4 + 3] = nact;
            }
        }
        __syncwarp(mask);
        RTI_GT(5);
    }
}

// ---- constrained kernel: the problems the nominal kernel handed over, pulled from the queue by 16-lane groups (so a
// hard problem never blocks the others) with a register budget of its own ----
template <typename T, int kN>
__global__ void __launch_bounds__(RTI_CTA, (sizeof(T) == 4) ? 4 : 2) rti_constrained_kernel(const __grid_constant__ RtiCfg<T> c, const RtiArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // launched as a programmatic dependent of the nominal kernel: wait for its queue (a no-op for an ordinary launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int n_easy = *reinterpret_cast<volatile int*>(a.qctl), n_hard = *reinterpret_cast<volatile int*>(a.qctl + 3);
    const int count = n_easy + n_hard;
#ifdef NDP_RTI_PROF
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); g_con_prof[16384] = t_; }
#endif
    const int N = (kN > 0) ? kN : c.N;
    const SmemLayout L(N);
    const int lane = threadIdx.x & 15;
    const int grp = threadIdx.x >> 4;
    const unsigned mask = 0xFFFFu << (threadIdx.x & 16);
    const int ppc = blockDim.x >> 4;
    if (count > 0) {
        T* sTriv = reinterpret_cast<T*>(smem_raw);
        T* sm = sTriv + 10 * TLD + (size_t)grp * L.total;
        T* ws = a.ws + (size_t)(blockIdx.x * ppc + grp) * a.ws_stride;
        for (int i = threadIdx.x; i < 10 * TLD; i += blockDim.x) {
            const int r = i / TLD, cc = i - r * TLD;
            sTriv[i] = (cc < 6 && cc == r) ? T(1) : ((cc >= 3 && cc < 6 && cc - 3 == r) ? c.h : T(0));
        }
        for (int i = lane; i < 20 * TLD; i += GL) sm[L.oT0 + i] = T(0);
        __syncthreads();
        T lo, hi;
        lane_box<T>(c, lane, lo, hi);
        const bool isx = lane < 10;
        RtiArgs<T> as_plain = a;  // the references were persisted by the nominal kernel: stage the stored yref / p
        as_plain.xr = nullptr;
        for (;;) {
            int slot = 0;
            if (lane == 0) slot = atomicAdd(a.qctl + 1, 1);
            slot = __shfl_sync(mask, slot, 0, GL);
            if (slot >= count) break;
            const int entry = a.queue[slot < n_hard ? slot : a.B - 1 - (slot - n_hard)];
            const int prob = entry & (QUEUE_SWEPT - 1);
#ifdef NDP_RTI_PROF
            if (lane == 0 && prob < 8192) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); g_con_prof[2 * prob] = t_; }
#endif
            stage_problem<T>(as_plain, N, L, lane, mask, sm, prob, false);
            cost_records<T>(N, lane, sm + L.oY, sm + L.oX, sm + L.oU, sm + L.oPar);
            __syncwarp(mask);
            const T dx0 = isx ? a.x0[(size_t)prob * NX + lane] - sm[L.oX + lane] : T(0);
            // stage tiles [A B b]: the nominal kernel's, when its sweep ran (QUEUE_SWEPT) and its slot still holds them
            const bool have_tiles = (entry & QUEUE_SWEPT) && a.ws_n_slots >= a.B;
            T* rec = have_tiles ? a.ws_n + (size_t)prob * a.ws_n_stride : ws + WsLayout(N).oRec;
            constrained_qp<T, kN>(c, lane, mask, sm, ws, rec, have_tiles, sTriv, dx0, lo, hi, a.X + (size_t)prob * (N + 1) * NX, a.U + (size_t)prob * N * NU,
                                  a.u0 ? a.u0 + (size_t)prob * NU : nullptr, a.status + prob, a.status2 ? a.status2 + prob : nullptr,
                                  a.stats + (size_t)prob * 4,
                                  a.as_store + (size_t)prob * (AS_OWNERS * 4), a.as_warm != 0, (entry & QUEUE_SWEPT) ? 1 : 0);
#ifdef NDP_RTI_PROF
            if (lane == 0 && prob < 8192) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); g_con_prof[2 * prob + 1] = t_; }
#endif
            // the tiles' zero pad columns (the sweeps of this problem may have run the ring over them)
            for (int i = lane; i < 20; i += GL) { T* q = sm + L.oT0 + i * TLD + 9; q[0] = T(0); q[1] = T(0); q[2] = T(0); }
            __syncwarp(mask);
        }
    }
    // the last CTA out re-arms the queue for the next solve
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.qctl + 2, 1) == (int)gridDim.x - 1) {
            a.qctl[0] = 0; a.qctl[1] = 0; a.qctl[2] = 0; a.qctl[3] = 0;
            __threadfence();
        }
    }
}

}  // namespace ndp
