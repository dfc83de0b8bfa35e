/*
 * ORACLE -- test infrastructure only.  Never linked, imported or executed by the
 * product path (ndp_nmpc_qd_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C fp64 restatement of the reference's NMPC hot path as acados executes
 * it: SQP_RTI = [ERK(RK4, 1 step/interval) + forward sensitivities] ->
 * [NONLINEAR_LS Gauss-Newton cost blocks] -> [OCP-structured QP solved by a
 * Mehrotra predictor-corrector interior point method with a backward Riccati
 * factorisation, i.e. the published HPIPM algorithm] -> full step.
 *
 * PARITY UNPINNED at the acados boundary: acados / HPIPM / BLASFEO / CasADi are
 * un-vendored, un-pinned third-party dependencies of the reference (imported at
 * ndp_nmpc/scripts/nmpc_ctl/nmpc_body_rate_ctl.py:14-15) and are absent from
 * /root/reference and from this image; the reference has no test or golden
 * vector for u0.  This file is cross-checked against an independent dense-KKT
 * numpy solve (oracle/nmpc_numpy.py) and anchored on the reference's own OCP
 * definition, cited per function below.
 *
 * Build: oracle/Makefile  ->  oracle/_build/libnmpc_oracle.so
 * -DORC_REAL=float gives an fp32 build used only to study rounding offline.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>

/* debugging switch read once (getenv scans the environment: not something to do per interior-point iteration of a timed run) */
static int orc_trace_on(void) {
    static int on = -1;
    if (on < 0) on = getenv("ORC_TRACE") ? 1 : 0;
    return on;
}
#endif

#ifndef ORC_REAL
#define ORC_REAL double
#endif
typedef ORC_REAL real;

#define NX 10
#define NU 4
#define NZ 14
#define NBU 4 /* bounded inputs  (all)     nmpc_body_rate_ctl.py:56-58 */
#define NBX 3 /* bounded states  (vx,vy,vz) nmpc_body_rate_ctl.py:59-61 */
#define NMAX 128

typedef struct {
    int N;            /* shooting intervals            params/nmpc_params.py:9  */
    double h;         /* interval length T/N           params/nmpc_params.py:12 */
    double mass;      /* params/fhnp_params.py:9  */
    double gravity;   /* params/fhnp_params.py:12 */
    double Q[NX];     /* diag state weights            nmpc_body_rate_ctl.py:48 */
    double R[NU];     /* diag input weights            nmpc_body_rate_ctl.py:49 */
    double u_min[NU], u_max[NU]; /* nmpc_body_rate_ctl.py:56-58 */
    double v_min[NBX], v_max[NBX]; /* nmpc_body_rate_ctl.py:59-61 */
    double tol;       /* IPM residual tolerance */
    double tol_mu;    /* IPM complementarity tolerance */
    int max_iter;     /* acados qp_solver_iter_max default 50 */
    double mu0;       /* IPM cold start */
    double t_floor;   /* slack floor at cold start */
    int polish;       /* > 0: after the IPM, up to `polish` exact active-set rounds (see orc_polish) */
    int pdas_first;   /* > 0: up to `pdas_first` active-set rounds from the unconstrained step's violated bounds BEFORE
                         the IPM (the route the CUDA kernel takes first); the IPM runs only if they find no fixed point */
} orc_cfg;

/* xdot = f(x,u;fd)   ndp_nmpc_body_rate_ctl.py:151-162 */
void orc_f(const orc_cfg* c, const real* x, const real* u, const real* fd, real* xd) {
    real qw = x[6], qx = x[7], qy = x[8], qz = x[9];
    real wx = u[0], wy = u[1], wz = u[2], cc = u[3];
    real im = (real)(1.0 / c->mass);
    xd[0] = x[3];
    xd[1] = x[4];
    xd[2] = x[5];
    xd[3] = 2 * (qx * qz + qw * qy) * cc + fd[0] * im;
    xd[4] = 2 * (qy * qz - qw * qx) * cc + fd[1] * im;
    xd[5] = (1 - 2 * qx * qx - 2 * qy * qy) * cc - (real)c->gravity + fd[2] * im;
    xd[6] = (-wx * qx - wy * qy - wz * qz) * (real)0.5;
    xd[7] = (wx * qw + wz * qy - wy * qz) * (real)0.5;
    xd[8] = (wy * qw - wz * qx + wx * qz) * (real)0.5;
    xd[9] = (wz * qw + wy * qx - wx * qy) * (real)0.5;
}

/* dense analytic Jacobians of f (row-major A[10][10], B[10][4]) */
static void orc_jac(const real* x, const real* u, real* A, real* B) {
    real qw = x[6], qx = x[7], qy = x[8], qz = x[9];
    real wx = u[0], wy = u[1], wz = u[2], c = u[3];
    memset(A, 0, sizeof(real) * NX * NX);
    memset(B, 0, sizeof(real) * NX * NU);
    A[0 * NX + 3] = A[1 * NX + 4] = A[2 * NX + 5] = 1;
    A[3 * NX + 6] = 2 * c * qy; A[3 * NX + 7] = 2 * c * qz; A[3 * NX + 8] = 2 * c * qw; A[3 * NX + 9] = 2 * c * qx;
    A[4 * NX + 6] = -2 * c * qx; A[4 * NX + 7] = -2 * c * qw; A[4 * NX + 8] = 2 * c * qz; A[4 * NX + 9] = 2 * c * qy;
    A[5 * NX + 7] = -4 * c * qx; A[5 * NX + 8] = -4 * c * qy;
    A[6 * NX + 7] = -wx / 2; A[6 * NX + 8] = -wy / 2; A[6 * NX + 9] = -wz / 2;
    A[7 * NX + 6] = wx / 2; A[7 * NX + 8] = wz / 2; A[7 * NX + 9] = -wy / 2;
    A[8 * NX + 6] = wy / 2; A[8 * NX + 7] = -wz / 2; A[8 * NX + 9] = wx / 2;
    A[9 * NX + 6] = wz / 2; A[9 * NX + 7] = wy / 2; A[9 * NX + 8] = -wx / 2;
    B[3 * NU + 3] = 2 * (qx * qz + qw * qy);
    B[4 * NU + 3] = 2 * (qy * qz - qw * qx);
    B[5 * NU + 3] = 1 - 2 * qx * qx - 2 * qy * qy;
    B[6 * NU + 0] = -qx / 2; B[6 * NU + 1] = -qy / 2; B[6 * NU + 2] = -qz / 2;
    B[7 * NU + 0] = qw / 2; B[7 * NU + 1] = -qz / 2; B[7 * NU + 2] = qy / 2;
    B[8 * NU + 0] = qz / 2; B[8 * NU + 1] = qw / 2; B[8 * NU + 2] = -qx / 2;
    B[9 * NU + 0] = -qy / 2; B[9 * NU + 1] = qx / 2; B[9 * NU + 2] = qw / 2;
}

/* variational right-hand side at a stage state: k = f, kS = A_c S + [0 B_c], S = [Sx Su] (10x14) */
static void orc_vde(const orc_cfg* c, const real* x, const real* u, const real* fd, const real* S, real* k, real* kS) {
    real A[NX * NX], B[NX * NU];
    orc_f(c, x, u, fd, k);
    orc_jac(x, u, A, B);
    for (int i = 0; i < NX; i++)
        for (int j = 0; j < NZ; j++) {
            real s = (j >= NX) ? B[i * NU + (j - NX)] : 0;
            for (int r = 0; r < NX; r++) s += A[i * NX + r] * S[r * NZ + j];
            kS[i * NZ + j] = s;
        }
}

/* One RK4 step with forward sensitivities (ERK, 4 stages, 1 step: acados defaults for
 * integrator_type="ERK", nmpc_body_rate_ctl.py:76,80).  AB = [A_k B_k] row-major 10x14. */
void orc_rk4_sens(const orc_cfg* c, const real* x, const real* u, const real* fd, real* xn, real* AB) {
    real h = (real)c->h;
    real S0[NX * NZ], xs[NX], Ss[NX * NZ];
    real k1[NX], k2[NX], k3[NX], k4[NX];
    real K1[NX * NZ], K2[NX * NZ], K3[NX * NZ], K4[NX * NZ];
    memset(S0, 0, sizeof(S0));
    for (int i = 0; i < NX; i++) S0[i * NZ + i] = 1;
    orc_vde(c, x, u, fd, S0, k1, K1);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + h / 2 * k1[i];
    for (int i = 0; i < NX * NZ; i++) Ss[i] = S0[i] + h / 2 * K1[i];
    orc_vde(c, xs, u, fd, Ss, k2, K2);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + h / 2 * k2[i];
    for (int i = 0; i < NX * NZ; i++) Ss[i] = S0[i] + h / 2 * K2[i];
    orc_vde(c, xs, u, fd, Ss, k3, K3);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + h * k3[i];
    for (int i = 0; i < NX * NZ; i++) Ss[i] = S0[i] + h * K3[i];
    orc_vde(c, xs, u, fd, Ss, k4, K4);
    for (int i = 0; i < NX; i++) xn[i] = x[i] + h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    for (int i = 0; i < NX * NZ; i++) AB[i] = S0[i] + h / 6 * (K1[i] + 2 * K2[i] + 2 * K3[i] + K4[i]);
}

/* Gauss-Newton Hessian block of the quaternion-error output: s * M(qr)' D M(qr)
 * (nmpc_body_rate_ctl.py:164-179). */
static void orc_quat_hess(const orc_cfg* c, const real* qr, real s, real* Hqq /*4x4*/) {
    real w = qr[0], x = qr[1], y = qr[2], z = qr[3];
    real M[3][4] = {{-x, w, -z, y}, {-y, z, w, -x}, {-z, -y, x, w}};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            real a = 0;
            for (int m = 0; m < 3; m++) a += (real)c->Q[7 + m] * M[m][i] * M[m][j];
            Hqq[i * 4 + j] = s * a;
        }
}

typedef struct {
    real AB[NMAX][NX * NZ]; /* [A_k B_k] */
    real b[NMAX][NX];
    real Hxx[NMAX + 1][NX * NX]; /* cost Hessian, x block (without barrier) */
    real gx[NMAX + 1][NX];
    real Huu[NMAX][NU]; /* diagonal */
    real gu[NMAX][NU];
    real lbu[NMAX][NU], ubu[NMAX][NU];  /* bounds in delta space */
    real lbx[NMAX][NBX], ubx[NMAX][NBX]; /* stages 1..N-1 */
    /* Riccati factors */
    real P[NMAX + 1][NX * NX], p[NMAX + 1][NX];
    real K[NMAX][NU * NX], kap[NMAX][NU], Lg[NMAX][NU * NU];
    real Hux[NMAX][NU * NX];
    /* barrier-modified diagonals / gradients */
    real dQ[NMAX + 1][NX], dq[NMAX + 1][NX], dR[NMAX][NU], dr[NMAX][NU];
} orc_ws;

static int chol4(real* G, real* L) { /* lower Cholesky of 4x4, returns nonzero on failure */
    memset(L, 0, sizeof(real) * 16);
    for (int j = 0; j < 4; j++) {
        real s = G[j * 4 + j];
        for (int k = 0; k < j; k++) s -= L[j * 4 + k] * L[j * 4 + k];
        if (!(s > 0)) return 1;
        L[j * 4 + j] = sqrt(s);
        for (int i = j + 1; i < 4; i++) {
            real a = G[i * 4 + j];
            for (int k = 0; k < j; k++) a -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = a / L[j * 4 + j];
        }
    }
    return 0;
}
static void chol4_solve(const real* L, real* v) { /* v <- (L L')^-1 v */
    for (int i = 0; i < 4; i++) {
        real a = v[i];
        for (int k = 0; k < i; k++) a -= L[i * 4 + k] * v[k];
        v[i] = a / L[i * 4 + i];
    }
    for (int i = 3; i >= 0; i--) {
        real a = v[i];
        for (int k = i + 1; k < 4; k++) a -= L[k * 4 + i] * v[k];
        v[i] = a / L[i * 4 + i];
    }
}

/* Backward Riccati sweep on the barrier-augmented LQ problem.
 * factorise != 0: recompute P_k, K_k (matrix part); always recompute p_k, kappa_k. */
static int riccati_backward(const orc_cfg* c, orc_ws* w, int factorise) {
    int N = c->N;
    if (factorise) {
        memcpy(w->P[N], w->Hxx[N], sizeof(real) * NX * NX);
        for (int i = 0; i < NX; i++) w->P[N][i * NX + i] += w->dQ[N][i];
    }
    for (int i = 0; i < NX; i++) w->p[N][i] = w->gx[N][i] + w->dq[N][i];
    for (int k = N - 1; k >= 0; k--) {
        const real* AB = w->AB[k];
        const real* Pn = w->P[k + 1];
        real wv[NX]; /* P+ b + p+ */
        for (int i = 0; i < NX; i++) {
            real a = w->p[k + 1][i];
            for (int r = 0; r < NX; r++) a += Pn[i * NX + r] * w->b[k][r];
            wv[i] = a;
        }
        if (factorise) {
            real W[NX * NZ]; /* P+ [A B] */
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < NZ; j++) {
                    real a = 0;
                    for (int r = 0; r < NX; r++) a += Pn[i * NX + r] * AB[r * NZ + j];
                    W[i * NZ + j] = a;
                }
            real H[NZ * NZ]; /* [A B]' P+ [A B] */
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NZ; j++) {
                    real a = 0;
                    for (int r = 0; r < NX; r++) a += AB[r * NZ + i] * W[r * NZ + j];
                    H[i * NZ + j] = a;
                }
            real G[16];
            for (int i = 0; i < NU; i++)
                for (int j = 0; j < NU; j++) G[i * 4 + j] = H[(NX + i) * NZ + NX + j] + (i == j ? w->Huu[k][i] + w->dR[k][i] : 0);
            if (chol4(G, w->Lg[k])) return 1;
            for (int m = 0; m < NU; m++)
                for (int j = 0; j < NX; j++) w->Hux[k][m * NX + j] = H[(NX + m) * NZ + j];
            /* K = -G^-1 Hux, column by column */
            for (int j = 0; j < NX; j++) {
                real v[4];
                for (int m = 0; m < 4; m++) v[m] = w->Hux[k][m * NX + j];
                chol4_solve(w->Lg[k], v);
                for (int m = 0; m < 4; m++) w->K[k][m * NX + j] = -v[m];
            }
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < NX; j++) {
                    real a = H[i * NZ + j] + w->Hxx[k][i * NX + j] + (i == j ? w->dQ[k][i] : 0);
                    for (int m = 0; m < NU; m++) a += w->Hux[k][m * NX + i] * w->K[k][m * NX + j];
                    w->P[k][i * NX + j] = a;
                }
            /* symmetrise */
            for (int i = 0; i < NX; i++)
                for (int j = 0; j < i; j++) {
                    real a = (real)0.5 * (w->P[k][i * NX + j] + w->P[k][j * NX + i]);
                    w->P[k][i * NX + j] = w->P[k][j * NX + i] = a;
                }
        }
        real gfull[NZ];
        for (int i = 0; i < NZ; i++) {
            real a = 0;
            for (int r = 0; r < NX; r++) a += AB[r * NZ + i] * wv[r];
            gfull[i] = a;
        }
        real v[4];
        for (int m = 0; m < NU; m++) v[m] = gfull[NX + m] + w->gu[k][m] + w->dr[k][m];
        chol4_solve(w->Lg[k], v);
        for (int m = 0; m < NU; m++) w->kap[k][m] = -v[m];
        for (int i = 0; i < NX; i++) {
            real a = gfull[i] + w->gx[k][i] + w->dq[k][i];
            for (int m = 0; m < NU; m++) a += w->Hux[k][m * NX + i] * w->kap[k][m];
            w->p[k][i] = a;
        }
    }
    return 0;
}

/* forward substitution: dx[0] given; fills du[0..N-1], dx[1..N] */
static void riccati_forward(const orc_cfg* c, const orc_ws* w, real (*dx)[NX], real (*du)[NU]) {
    for (int k = 0; k < c->N; k++) {
        for (int m = 0; m < NU; m++) {
            real a = w->kap[k][m];
            for (int j = 0; j < NX; j++) a += w->K[k][m * NX + j] * dx[k][j];
            du[k][m] = a;
        }
        for (int i = 0; i < NX; i++) {
            real a = w->b[k][i];
            for (int j = 0; j < NX; j++) a += w->AB[k][i * NZ + j] * dx[k][j];
            for (int m = 0; m < NU; m++) a += w->AB[k][i * NZ + NX + m] * du[k][m];
            dx[k + 1][i] = a;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * Exact solve for a GIVEN active set (used to polish the interior-point iterate).
 *
 * An IPM whose barrier terms enter the Riccati recursion as diagonal weights lambda/t loses
 * accuracy on problems with active STATE bounds: the weight (~1/mu) on a velocity component at
 * stage k+1 reappears at stage k inside [A B]'P+[A B], and the Schur complement that eliminates
 * the inputs then subtracts two numbers of size 1/mu (measured here: u0 off by 3e-5 at mu = 1e-8,
 * by 1e-1 at mu = 1e-13, against the dense KKT solve of oracle/nmpc_numpy.py).  HPIPM meets the same
 * effect with iterative refinement and its LQ-factorisation fallback [EXT].  The restatement
 * removes it at the root: once the IPM has identified the active set, the QP is solved as an
 * equality-constrained LQ problem with
 *   - pinned inputs eliminated exactly (u_m = bound), and
 *   - pinned velocity components of x_{k+1} treated as the stage-k mixed constraint
 *     E (A dx_k + B du_k + b_k) = beta, resolved in the range/null space of the free inputs
 *     (needs E B_free of full row rank: collective thrust and the two tilt rates move the velocity),
 * followed by primal-dual active-set updates (release on a wrong multiplier sign, add on a violated
 * bound) until the set is a fixed point -- which is the KKT point of the QP, with no barrier floor. */
#ifndef ORC_PDAS_FREE
#define ORC_PDAS_FREE 1000
#endif
typedef struct {
    signed char au[NMAX][NU];       /* -1 lower, +1 upper, 0 free */
    signed char av[NMAX + 1][NBX];  /* stages 1..N-1 */
} orc_aset;

typedef struct {
    real K[NMAX][NU * NX], kap[NMAX][NU];
    real Tx[NMAX][NBX * NX], T0[NMAX][NBX]; /* nu_k = -(Tx dx_k + T0), rows = pinned velocity components of stage k+1 */
    int na[NMAX], ra[NMAX][NBX];             /* number / indices (0..2) of those components */
    real S[NMAX][NZ * NZ], g[NMAX][NZ];      /* stage quadratic incl. cost-to-go, [x;u] ordering */
} orc_eqws;

static int chol_n(int n, const real* G, real* L) { /* lower Cholesky n <= 4, row-major stride 4 */
    memset(L, 0, sizeof(real) * 16);
    for (int j = 0; j < n; j++) {
        real s = G[j * 4 + j];
        for (int k = 0; k < j; k++) s -= L[j * 4 + k] * L[j * 4 + k];
        if (!(s > 0)) return 1;
        L[j * 4 + j] = sqrt(s);
        for (int i = j + 1; i < n; i++) {
            real a = G[i * 4 + j];
            for (int k = 0; k < j; k++) a -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = a / L[j * 4 + j];
        }
    }
    return 0;
}
static void chol_n_solve(int n, const real* L, real* v) {
    for (int i = 0; i < n; i++) {
        real a = v[i];
        for (int k = 0; k < i; k++) a -= L[i * 4 + k] * v[k];
        v[i] = a / L[i * 4 + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        real a = v[i];
        for (int k = i + 1; k < n; k++) a -= L[k * 4 + i] * v[k];
        v[i] = a / L[i * 4 + i];
    }
}

/* backward + forward sweep of the equality-constrained LQ problem; fills dx, du and the equality-form
 * multipliers lam_u (pinned inputs) and nu_v[k] (pinned velocity components of stage k+1).  Nonzero: singular. */
static int orc_eq_solve(const orc_cfg* c, orc_ws* w, orc_eqws* e, const orc_aset* as, real (*dx)[NX], real (*du)[NU],
                        real (*lam_u)[NU], real (*nu_v)[NBX]) {
    int N = c->N;
    memcpy(w->P[N], w->Hxx[N], sizeof(real) * NX * NX);
    memcpy(w->p[N], w->gx[N], sizeof(real) * NX);
    for (int k = N - 1; k >= 0; k--) {
        const real* AB = w->AB[k];
        const real* Pn = w->P[k + 1];
        real* S = e->S[k];
        real* g = e->g[k];
        real W[NX * NZ], wv[NX];
        for (int i = 0; i < NX; i++) {
            real a = w->p[k + 1][i];
            for (int r = 0; r < NX; r++) a += Pn[i * NX + r] * w->b[k][r];
            wv[i] = a;
            for (int j = 0; j < NZ; j++) {
                real t = 0;
                for (int r = 0; r < NX; r++) t += Pn[i * NX + r] * AB[r * NZ + j];
                W[i * NZ + j] = t;
            }
        }
        for (int i = 0; i < NZ; i++) {
            for (int j = 0; j < NZ; j++) {
                real a = 0;
                for (int r = 0; r < NX; r++) a += AB[r * NZ + i] * W[r * NZ + j];
                if (i < NX && j < NX) a += w->Hxx[k][i * NX + j];
                if (i >= NX && i == j) a += w->Huu[k][i - NX];
                S[i * NZ + j] = a;
            }
            real a = 0;
            for (int r = 0; r < NX; r++) a += AB[r * NZ + i] * wv[r];
            g[i] = a + (i < NX ? w->gx[k][i] : w->gu[k][i - NX]);
        }
        /* pinned inputs (value bu) / pinned velocity components of x_{k+1} (value bv) */
        int pin[NU]; real bu[NU];
        for (int m = 0; m < NU; m++) {
            pin[m] = as->au[k][m] != 0;
            bu[m] = as->au[k][m] > 0 ? w->ubu[k][m] : w->lbu[k][m];
        }
        int na = 0; real bv[NBX];
        if (k + 1 <= N - 1)
            for (int m = 0; m < NBX; m++)
                if (as->av[k + 1][m]) { e->ra[k][na] = m; bv[na] = as->av[k + 1][m] > 0 ? w->ubx[k + 1][m] : w->lbx[k + 1][m]; na++; }
        e->na[k] = na;
        real G[16], hx[NU * NX], hg[NU];
        for (int i = 0; i < NU; i++) {
            for (int j = 0; j < NU; j++) G[i * 4 + j] = (pin[i] || pin[j]) ? (i == j ? 1 : 0) : S[(NX + i) * NZ + NX + j];
            for (int j = 0; j < NX; j++) hx[i * NX + j] = pin[i] ? 0 : S[(NX + i) * NZ + j];
            if (pin[i]) hg[i] = -bu[i];
            else {
                real a = g[NX + i];
                for (int m = 0; m < NU; m++) if (pin[m]) a += S[(NX + i) * NZ + NX + m] * bu[m];
                hg[i] = a;
            }
        }
        real L[16];
        if (chol_n(4, G, L)) return 1;
        real Ku[NU * NX], ku[NU];
        for (int j = 0; j < NX; j++) {
            real v[4];
            for (int m = 0; m < 4; m++) v[m] = hx[m * NX + j];
            chol_n_solve(4, L, v);
            for (int m = 0; m < 4; m++) Ku[m * NX + j] = -v[m];
        }
        { real v[4]; for (int m = 0; m < 4; m++) v[m] = hg[m]; chol_n_solve(4, L, v); for (int m = 0; m < 4; m++) ku[m] = -v[m]; }
        real* K = e->K[k];
        real* kap = e->kap[k];
        memcpy(K, Ku, sizeof(Ku));
        memcpy(kap, ku, sizeof(ku));
        if (na) {
            real D[NBX * NU], Y[NU * NBX], Sn[16], Ln[16];
            for (int a = 0; a < na; a++)
                for (int m = 0; m < NU; m++) D[a * NU + m] = pin[m] ? 0 : AB[(3 + e->ra[k][a]) * NZ + NX + m];
            for (int a = 0; a < na; a++) {
                real v[4];
                for (int m = 0; m < 4; m++) v[m] = D[a * NU + m];
                chol_n_solve(4, L, v);
                for (int m = 0; m < 4; m++) Y[m * NBX + a] = v[m];
            }
            for (int a = 0; a < na; a++)
                for (int b2 = 0; b2 < na; b2++) {
                    real t = 0;
                    for (int m = 0; m < NU; m++) t += D[a * NU + m] * Y[m * NBX + b2];
                    Sn[a * 4 + b2] = t;
                }
            if (chol_n(na, Sn, Ln)) return 2;
            /* right-hand sides d_x - D Ku (per x column) and d_0 - D ku */
            for (int j = 0; j <= NX; j++) {
                real v[4] = {0, 0, 0, 0};
                for (int a = 0; a < na; a++) {
                    int r = 3 + e->ra[k][a];
                    real t;
                    if (j < NX) {
                        t = -AB[r * NZ + j];
                        for (int m = 0; m < NU; m++) t -= D[a * NU + m] * Ku[m * NX + j];
                    } else {
                        t = bv[a] - w->b[k][r];
                        for (int m = 0; m < NU; m++) {
                            if (pin[m]) t -= AB[r * NZ + NX + m] * bu[m];
                            t -= D[a * NU + m] * ku[m];
                        }
                    }
                    v[a] = t;
                }
                chol_n_solve(na, Ln, v);
                for (int a = 0; a < na; a++) {
                    if (j < NX) e->Tx[k][a * NX + j] = v[a]; else e->T0[k][a] = v[a];
                }
            }
            for (int m = 0; m < NU; m++) {
                for (int j = 0; j < NX; j++) {
                    real t = 0;
                    for (int a = 0; a < na; a++) t += Y[m * NBX + a] * e->Tx[k][a * NX + j];
                    K[m * NX + j] += t;
                }
                real t = 0;
                for (int a = 0; a < na; a++) t += Y[m * NBX + a] * e->T0[k][a];
                kap[m] += t;
            }
        }
        /* value function under the affine policy du = K dx + kap (the pinned rows of K are 0, of kap the pinned value) */
        real SuK[NU * NX], gk[NU];
        for (int m = 0; m < NU; m++) {
            for (int j = 0; j < NX; j++) {
                real t = 0;
                for (int n = 0; n < NU; n++) t += S[(NX + m) * NZ + NX + n] * K[n * NX + j];
                SuK[m * NX + j] = t;
            }
            real t = g[NX + m];
            for (int n = 0; n < NU; n++) t += S[(NX + m) * NZ + NX + n] * kap[n];
            gk[m] = t;
        }
        for (int i = 0; i < NX; i++) {
            for (int j = 0; j < NX; j++) {
                real a = S[i * NZ + j];
                for (int m = 0; m < NU; m++) a += S[i * NZ + NX + m] * K[m * NX + j] + K[m * NX + i] * S[(NX + m) * NZ + j] + K[m * NX + i] * SuK[m * NX + j];
                w->P[k][i * NX + j] = a;
            }
            real a = g[i];
            for (int m = 0; m < NU; m++) a += S[i * NZ + NX + m] * kap[m] + K[m * NX + i] * gk[m];
            w->p[k][i] = a;
        }
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < i; j++) {
                real a = (real)0.5 * (w->P[k][i * NX + j] + w->P[k][j * NX + i]);
                w->P[k][i * NX + j] = w->P[k][j * NX + i] = a;
            }
    }
    for (int k = 0; k < N; k++) {
        for (int m = 0; m < NU; m++) {
            real a = e->kap[k][m];
            for (int j = 0; j < NX; j++) a += e->K[k][m * NX + j] * dx[k][j];
            du[k][m] = a;
        }
        for (int i = 0; i < NX; i++) {
            real a = w->b[k][i];
            for (int j = 0; j < NX; j++) a += w->AB[k][i * NZ + j] * dx[k][j];
            for (int m = 0; m < NU; m++) a += w->AB[k][i * NZ + NX + m] * du[k][m];
            dx[k + 1][i] = a;
        }
        real nu[NBX] = {0, 0, 0};
        for (int a = 0; a < e->na[k]; a++) {
            real t = e->T0[k][a];
            for (int j = 0; j < NX; j++) t += e->Tx[k][a * NX + j] * dx[k][j];
            nu[a] = -t;
            nu_v[k][e->ra[k][a]] = -t;
        }
        for (int m = 0; m < NU; m++) {
            real t = e->g[k][NX + m];
            for (int j = 0; j < NX; j++) t += e->S[k][(NX + m) * NZ + j] * dx[k][j];
            for (int n = 0; n < NU; n++) t += e->S[k][(NX + m) * NZ + NX + n] * du[k][n];
            for (int a = 0; a < e->na[k]; a++) t += w->AB[k][(3 + e->ra[k][a]) * NZ + NX + m] * nu[a];
            lam_u[k][m] = -t;
        }
    }
    return 0;
}

/* Primal-dual active-set rounds from the set `as`.  Returns 1 when a fixed point was reached (dx, du hold the QP
 * solution), 0 otherwise (dx, du untouched). */
static int orc_polish(const orc_cfg* c, orc_ws* w, orc_aset* as, real (*dx)[NX], real (*du)[NU], int max_rounds, int* rounds) {
    int N = c->N;
    orc_eqws* e = (orc_eqws*)malloc(sizeof(orc_eqws));
    real (*tx)[NX] = (real(*)[NX])malloc(sizeof(real) * (N + 1) * NX);
    real (*tu)[NU] = (real(*)[NU])malloc(sizeof(real) * N * NU * 2);
    real (*lam)[NU] = tu + N;
    real (*nu)[NBX] = (real(*)[NBX])malloc(sizeof(real) * N * NBX);
    memcpy(tx[0], dx[0], sizeof(real) * NX);
    const real eps = (sizeof(real) == 8) ? (real)1e-11 : (real)1e-5;
    int fixed = 0, r = 0;
    unsigned long long hist[64];
    int n_hist = 0, cycling = 0;
    for (r = 0; r < max_rounds; r++) {
        memset(nu, 0, sizeof(real) * N * NBX);
        if (orc_eq_solve(c, w, e, as, tx, tu, lam, nu)) break;
        int changed = 0;
        /* Rounds 0..ORC_PDAS_FREE-1: plain primal-dual update (release every wrong-signed multiplier, add every violated
         * bound).  That iteration can cycle between a few sets; later rounds therefore release only the single most
         * wrong-signed multiplier, and only once no bound is violated. */
        /* cycle detection: hash of the set each round; a repeat switches to the damped update for good */
        {
            unsigned long long hsh = 1469598103934665603ull;
            for (int k = 0; k < N; k++) {
                for (int m = 0; m < NU; m++) hsh = (hsh ^ (unsigned long long)(as->au[k][m] + 2)) * 1099511628211ull;
                for (int m = 0; m < NBX; m++) hsh = (hsh ^ (unsigned long long)(as->av[k][m] + 5)) * 1099511628211ull;
            }
            for (int q = 0; q < n_hist; q++) if (hist[q] == hsh) cycling = 1;
            if (n_hist < 64) hist[n_hist++] = hsh;
        }
        const int damped = cycling || r >= ORC_PDAS_FREE;
        int n_viol = 0;
        real worst = 0; int wk = -1, wm = -1, wx = 0;
        for (int k = 0; k < N; k++) {
            for (int m = 0; m < NU; m++) {
                if (as->au[k][m]) {
                    /* equality-form multiplier: upper bound needs lam >= 0, lower bound lam <= 0 */
                    real v = as->au[k][m] > 0 ? lam[k][m] : -lam[k][m];
                    if (v < 0) {
                        if (!damped) { as->au[k][m] = 0; changed = 1; }
                        else if (v < worst) { worst = v; wk = k; wm = m; wx = 0; }
                    }
                } else if (tu[k][m] > w->ubu[k][m] + eps) { as->au[k][m] = 1; changed = 1; n_viol++; }
                else if (tu[k][m] < w->lbu[k][m] - eps) { as->au[k][m] = -1; changed = 1; n_viol++; }
            }
            if (k + 1 <= N - 1)
                for (int m = 0; m < NBX; m++) {
                    if (as->av[k + 1][m]) {
                        real v = as->av[k + 1][m] > 0 ? nu[k][m] : -nu[k][m];
                        if (v < 0) {
                            if (!damped) { as->av[k + 1][m] = 0; changed = 1; }
                            else if (v < worst) { worst = v; wk = k + 1; wm = m; wx = 1; }
                        }
                    } else if (tx[k + 1][3 + m] > w->ubx[k + 1][m] + eps) { as->av[k + 1][m] = 1; changed = 1; n_viol++; }
                    else if (tx[k + 1][3 + m] < w->lbx[k + 1][m] - eps) { as->av[k + 1][m] = -1; changed = 1; n_viol++; }
                }
        }
        if (damped && n_viol == 0 && wk >= 0) {
            if (getenv("ORC_REL_ALL")) {
                for (int k = 0; k < N; k++) {
                    for (int m = 0; m < NU; m++) if (as->au[k][m] && (as->au[k][m] > 0 ? lam[k][m] : -lam[k][m]) < 0) as->au[k][m] = 0;
                    if (k + 1 <= N - 1) for (int m = 0; m < NBX; m++) if (as->av[k + 1][m] && (as->av[k + 1][m] > 0 ? nu[k][m] : -nu[k][m]) < 0) as->av[k + 1][m] = 0;
                }
            } else if (wx) as->av[wk][wm] = 0; else as->au[wk][wm] = 0;
            changed = 1;
        }
        if (getenv("ORC_KTOP")) {
            static signed char prev_u[NMAX][NU], prev_v[NMAX + 1][NBX];
            int ktop = -1;
            for (int k = 0; k < N; k++) {
                for (int m = 0; m < NU; m++) if (as->au[k][m] != prev_u[k][m]) ktop = k > ktop ? k : ktop;
                for (int m = 0; m < NBX; m++) if (as->av[k][m] != prev_v[k][m]) ktop = (k - 1) > ktop ? (k - 1) : ktop;
            }
            memcpy(prev_u, as->au, sizeof(prev_u)); memcpy(prev_v, as->av, sizeof(prev_v));
            fprintf(stderr, "KTOP %d %d\n", r, ktop);
        }
        if (!changed) { fixed = 1; r++; break; }
    }
    if (fixed) {
        memcpy(dx, tx, sizeof(real) * (N + 1) * NX);
        memcpy(du, tu, sizeof(real) * N * NU);
    }
    if (rounds) *rounds = r;
    free(e); free(tx); free(tu); free(nu);
    return fixed;
}

typedef struct {
    int status;    /* acados codes: 0 ok, 1 NaN, 4 QP failure */
    int n_iter;    /* IPM iterations */
    int n_active;  /* active one-sided bounds at the solution */
    double res;    /* final max residual (primal slack consistency) */
    double mu;     /* final complementarity */
} orc_stats;

/* One SQP_RTI step for one problem.  x0[10], xr[(N+1)*10], ur[N*4], fd[(N+1)*3] (N; may be
 * NULL = zeros), X[(N+1)*10], U[N*4] iterate in/out.  Protocol of
 * nmpc_body_rate_ctl.py:93-112 (update) with acados' solve_for_x0. */
int orc_rti_step(const orc_cfg* c, const real* x0, const real* xr, const real* ur, const real* fd, real* X, real* U,
                 orc_stats* st, orc_ws* w) {
    int N = c->N;
    int own = 0;
    if (!w) { w = (orc_ws*)malloc(sizeof(orc_ws)); own = 1; }
    real zero3[3] = {0, 0, 0};
    real h = (real)c->h;
    /* ---- preparation: linearise ---- */
    for (int k = 0; k < N; k++) {
        real xn[NX];
        orc_rk4_sens(c, X + k * NX, U + k * NU, fd ? fd + 3 * k : zero3, xn, w->AB[k]);
        for (int i = 0; i < NX; i++) w->b[k][i] = xn[i] - X[(k + 1) * NX + i];
    }
    for (int k = 0; k <= N; k++) {
        real s = (k < N) ? h : 1;
        const real* xk = X + k * NX;
        const real* xrk = xr + k * NX;
        real Hqq[16];
        orc_quat_hess(c, xrk + 6, s, Hqq);
        memset(w->Hxx[k], 0, sizeof(real) * NX * NX);
        for (int i = 0; i < 6; i++) {
            w->Hxx[k][i * NX + i] = s * (real)c->Q[i];
            w->gx[k][i] = s * (real)c->Q[i] * (xk[i] - xrk[i]);
        }
        for (int i = 0; i < 4; i++) {
            real a = 0;
            for (int j = 0; j < 4; j++) {
                w->Hxx[k][(6 + i) * NX + 6 + j] = Hqq[i * 4 + j];
                a += Hqq[i * 4 + j] * xk[6 + j];
            }
            w->gx[k][6 + i] = a;
        }
        if (k < N)
            for (int m = 0; m < NU; m++) {
                w->Huu[k][m] = s * (real)c->R[m];
                w->gu[k][m] = s * (real)c->R[m] * (U[k * NU + m] - ur[k * NU + m]);
                w->lbu[k][m] = (real)c->u_min[m] - U[k * NU + m];
                w->ubu[k][m] = (real)c->u_max[m] - U[k * NU + m];
            }
        if (k >= 1 && k < N)
            for (int m = 0; m < NBX; m++) {
                w->lbx[k][m] = (real)c->v_min[m] - xk[3 + m];
                w->ubx[k][m] = (real)c->v_max[m] - xk[3 + m];
            }
    }
    /* ---- feedback: Mehrotra IPM with Riccati ---- */
    static const int NB_MAX = NMAX * (NBU + NBX);
    int nb = 0;  /* number of two-sided boxes */
    /* box list: (stage, is_x, idx) flattened */
    int* bs = (int*)malloc(sizeof(int) * 3 * NB_MAX);
    for (int k = 0; k < N; k++) {
        for (int m = 0; m < NBU; m++) { bs[3 * nb] = k; bs[3 * nb + 1] = 0; bs[3 * nb + 2] = m; nb++; }
        if (k >= 1) for (int m = 0; m < NBX; m++) { bs[3 * nb] = k; bs[3 * nb + 1] = 1; bs[3 * nb + 2] = m; nb++; }
    }
    (void)NB_MAX;
    real* lb = (real*)malloc(sizeof(real) * nb * 12);
    real *ub = lb + nb, *tl = ub + nb, *tu = tl + nb, *ll = tu + nb, *lu = ll + nb;
    real *dtl = lu + nb, *dtu = dtl + nb, *dll = dtu + nb, *dlu = dll + nb, *zb = dlu + nb, *zbn = zb + nb;
    for (int i = 0; i < nb; i++) {
        int k = bs[3 * i], m = bs[3 * i + 2];
        lb[i] = bs[3 * i + 1] ? w->lbx[k][m] : w->lbu[k][m];
        ub[i] = bs[3 * i + 1] ? w->ubx[k][m] : w->ubu[k][m];
    }
    real (*dx)[NX] = (real(*)[NX])malloc(sizeof(real) * (N + 1) * NX * 2);
    real (*dxn)[NX] = dx + (N + 1);
    real (*du)[NU] = (real(*)[NU])malloc(sizeof(real) * N * NU * 2);
    real (*dun)[NU] = du + N;
    memset(dx, 0, sizeof(real) * (N + 1) * NX * 2);
    memset(du, 0, sizeof(real) * N * NU * 2);
    for (int i = 0; i < NX; i++) dx[0][i] = dxn[0][i] = x0[i] - X[i];
    /* cold start (HPIPM-style): z = 0, slacks floored, lam = mu0 / t */
    real mu0 = (real)c->mu0, thr = (real)c->t_floor;
    for (int i = 0; i < nb; i++) {
        zb[i] = 0;
        tl[i] = fmax(zb[i] - lb[i], thr);
        tu[i] = fmax(ub[i] - zb[i], thr);
        ll[i] = mu0 / tl[i];
        lu[i] = mu0 / tu[i];
    }
    int status = 4, it = 0;
    real mu = 0, res = 0;
    real tol = (real)c->tol;
    int pdas_rounds = 0;
    if (c->pdas_first > 0) {
        /* unconstrained step (empty active set), then rounds from the bounds it violates */
        orc_aset as;
        memset(&as, 0, sizeof(as));
        if (orc_polish(c, w, &as, dx, du, c->pdas_first + 1, &pdas_rounds)) {
            int na2 = 0;
            for (int k = 0; k < N; k++) {
                for (int m = 0; m < NU; m++) na2 += as.au[k][m] != 0;
                if (k >= 1) for (int m = 0; m < NBX; m++) na2 += as.av[k][m] != 0;
            }
            int nan2 = 0;
            for (int k = 0; k <= N; k++)
                for (int i = 0; i < NX; i++) { X[k * NX + i] += dx[k][i]; nan2 |= !isfinite(X[k * NX + i]); }
            for (int k = 0; k < N; k++)
                for (int m = 0; m < NU; m++) { U[k * NU + m] += du[k][m]; nan2 |= !isfinite(U[k * NU + m]); }
            if (st) { st->status = nan2 ? 1 : 0; st->n_iter = -pdas_rounds; st->n_active = na2; st->res = 0; st->mu = 0; }
            free(bs); free(lb); free(dx); free(du);
            if (own) free(w);
            return nan2 ? 1 : 0;
        }
    }
    /* residual tracking: the Newton system is solved exactly, so the stationarity and
     * dynamics residuals contract by (1-alpha) each iteration; res_lin is that factor times
     * the initial residual bound (checked explicitly below for the slack equations). */
    real res_lin = 1;
    real tol_mu = (real)c->tol_mu;
    int polish_tries = 0;
ipm_again:
    for (; it <= c->max_iter; it++) {
        mu = 0; res = 0;
        for (int i = 0; i < nb; i++) {
            mu += ll[i] * tl[i] + lu[i] * tu[i];
            res = fmax(res, fabs(tl[i] - (zb[i] - lb[i])));
            res = fmax(res, fabs(tu[i] - (ub[i] - zb[i])));
        }
        mu /= (2 * nb);
        if (it > 0 && res < tol && mu < tol_mu && res_lin < tol) { status = 0; break; }
        if (it == c->max_iter) break;
        real sigma_mu = 0;
        real alpha = 1;
        for (int pass = 0; pass < 2; pass++) {
            /* barrier terms -> diagonal Hessian / gradient updates */
            for (int k = 0; k <= N; k++) {
                memset(w->dQ[k], 0, sizeof(real) * NX);
                memset(w->dq[k], 0, sizeof(real) * NX);
                if (k < N) { memset(w->dR[k], 0, sizeof(real) * NU); memset(w->dr[k], 0, sizeof(real) * NU); }
            }
            for (int i = 0; i < nb; i++) {
                int k = bs[3 * i], isx = bs[3 * i + 1], m = bs[3 * i + 2];
                real gl = ll[i] / tl[i], gu_ = lu[i] / tu[i];
                real cl = pass ? dll[i] * dtl[i] : 0, cu = pass ? dlu[i] * dtu[i] : 0;
                real term_u = (sigma_mu - cu) / tu[i] - gu_ * ub[i] + lu[i];
                real term_l = (sigma_mu - cl) / tl[i] + gl * lb[i] + ll[i];
                if (isx) { w->dQ[k][3 + m] = gl + gu_; w->dq[k][3 + m] = term_u - term_l; }
                else { w->dR[k][m] = gl + gu_; w->dr[k][m] = term_u - term_l; }
            }
            if (riccati_backward(c, w, pass == 0)) { status = 4; goto done; }
            riccati_forward(c, w, dxn, dun);
            /* directions on the bounded variables */
            real a_max = 1e30;
            for (int i = 0; i < nb; i++) {
                int k = bs[3 * i], isx = bs[3 * i + 1], m = bs[3 * i + 2];
                real zn = isx ? dxn[k][3 + m] : dun[k][m];
                zbn[i] = zn;
                real gl = ll[i] / tl[i], gu_ = lu[i] / tu[i];
                real cl = pass ? dll[i] * dtl[i] : 0, cu = pass ? dlu[i] * dtu[i] : 0;
                real ntl = zn - lb[i], ntu = ub[i] - zn; /* full-step slacks */
                real nll = (sigma_mu - cl) / tl[i] - gl * (ntl - tl[i]);
                real nlu = (sigma_mu - cu) / tu[i] - gu_ * (ntu - tu[i]);
                /* lam+ = (sigma mu - corr)/t - Gamma*dt  with dt = t+ - t */
                dtl[i] = ntl - tl[i]; dtu[i] = ntu - tu[i];
                dll[i] = nll - ll[i]; dlu[i] = nlu - lu[i];
                if (dtl[i] < 0) a_max = fmin(a_max, -tl[i] / dtl[i]);
                if (dtu[i] < 0) a_max = fmin(a_max, -tu[i] / dtu[i]);
                if (dll[i] < 0) a_max = fmin(a_max, -ll[i] / dll[i]);
                if (dlu[i] < 0) a_max = fmin(a_max, -lu[i] / dlu[i]);
            }
            if (pass == 0) {
                real a_aff = fmin(a_max, 1);
                real mu_aff = 0;
                for (int i = 0; i < nb; i++)
                    mu_aff += (ll[i] + a_aff * dll[i]) * (tl[i] + a_aff * dtl[i]) + (lu[i] + a_aff * dlu[i]) * (tu[i] + a_aff * dtu[i]);
                mu_aff /= (2 * nb);
                real s = mu_aff / mu;
                sigma_mu = s * s * s * mu;
            } else {
                alpha = fmin(1, (real)0.995 * a_max);
            }
        }
        for (int i = 0; i < nb; i++) {
            zb[i] += alpha * (zbn[i] - zb[i]);
            tl[i] += alpha * dtl[i]; tu[i] += alpha * dtu[i];
            ll[i] += alpha * dll[i]; lu[i] += alpha * dlu[i];
        }
        for (int k = 0; k <= N; k++)
            for (int i = 0; i < NX; i++) dx[k][i] += alpha * (dxn[k][i] - dx[k][i]);
        for (int k = 0; k < N; k++)
            for (int m = 0; m < NU; m++) du[k][m] += alpha * (dun[k][m] - du[k][m]);
        res_lin *= (1 - alpha);
        if (orc_trace_on()) fprintf(stderr, "it %d mu %.3e res %.3e alpha %.4f res_lin %.3e sigma_mu %.3e\n", it, (double)mu, (double)res, (double)alpha, (double)res_lin, (double)sigma_mu);
    }
done:;
    int nact = 0;
    for (int i = 0; i < nb; i++) { nact += (tl[i] < ll[i]); nact += (tu[i] < lu[i]); }
    if (c->polish > 0 && status != 1) {
        /* exact solve on the IPM's active set (t < lambda marks a bound as active), see orc_polish */
        orc_aset as;
        memset(&as, 0, sizeof(as));
        for (int i = 0; i < nb; i++) {
            int k = bs[3 * i], isx = bs[3 * i + 1], m = bs[3 * i + 2];
            signed char a = (tl[i] < ll[i]) ? -1 : ((tu[i] < lu[i]) ? 1 : 0);
            if (isx) as.av[k][m] = a; else as.au[k][m] = a;
        }
        int rounds = 0;
        if (orc_polish(c, w, &as, dx, du, c->polish, &rounds)) {
            status = 0;
            nact = 0;
            for (int k = 0; k < N; k++) {
                for (int m = 0; m < NU; m++) nact += as.au[k][m] != 0;
                if (k >= 1) for (int m = 0; m < NBX; m++) nact += as.av[k][m] != 0;
            }
        }
        else if (polish_tries == 0 && status == 0 && tol > (real)1e-13 && sizeof(real) == 8) {
            /* no fixed point from this active-set estimate (the rounds can cycle on near-degenerate bounds): carry the
             * interior-point iteration on to its floor instead, then try the rounds once more */
            polish_tries = 1;
            tol = tol_mu = (real)1e-13;
            status = 4;
            goto ipm_again;
        }
        if (orc_trace_on()) fprintf(stderr, "polish: rounds %d status %d nact %d\n", rounds, status, nact);
    }
    /* ---- update: full step ---- */
    int nan = 0;
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < NX; i++) { X[k * NX + i] += dx[k][i]; nan |= !isfinite(X[k * NX + i]); }
    for (int k = 0; k < N; k++)
        for (int m = 0; m < NU; m++) { U[k * NU + m] += du[k][m]; nan |= !isfinite(U[k * NU + m]); }
    if (nan) status = 1;
    if (st) { st->status = status; st->n_iter = it; st->n_active = nact; st->res = res; st->mu = mu; }
    free(bs); free(lb); free(dx); free(du);
    if (own) free(w);
    return status;
}

/* Batched driver: problems are independent; OpenMP over the batch, one problem per thread.
 * Arrays are [B][...] contiguous per problem.  Returns the number of threads used. */
int orc_rti_batch(const orc_cfg* c, int B, const real* x0, const real* xr, const real* ur, const real* fd, real* X, real* U,
                  real* u0, int* status, int* n_iter, int* n_active, int nthreads) {
    int N = c->N;
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
        orc_ws* w = (orc_ws*)malloc(sizeof(orc_ws));
#pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < B; b++) {
            orc_stats st;
            orc_rti_step(c, x0 + (size_t)b * NX, xr + (size_t)b * (N + 1) * NX, ur + (size_t)b * N * NU,
                         fd ? fd + (size_t)b * (N + 1) * 3 : NULL, X + (size_t)b * (N + 1) * NX, U + (size_t)b * N * NU, &st, w);
            for (int m = 0; m < NU; m++) u0[(size_t)b * NU + m] = U[(size_t)b * N * NU + m];
            if (status) status[b] = st.status;
            if (n_iter) n_iter[b] = st.n_iter;
            if (n_active) n_active[b] = st.n_active;
        }
        free(w);
    }
#else
    orc_ws* w = (orc_ws*)malloc(sizeof(orc_ws));
    for (int b = 0; b < B; b++) {
        orc_stats st;
        orc_rti_step(c, x0 + (size_t)b * NX, xr + (size_t)b * (N + 1) * NX, ur + (size_t)b * N * NU,
                     fd ? fd + (size_t)b * (N + 1) * 3 : NULL, X + (size_t)b * (N + 1) * NX, U + (size_t)b * N * NU, &st, w);
        for (int m = 0; m < NU; m++) u0[(size_t)b * NU + m] = U[(size_t)b * N * NU + m];
        if (status) status[b] = st.status;
        if (n_iter) n_iter[b] = st.n_iter;
        if (n_active) n_active[b] = st.n_active;
    }
    free(w);
#endif
    return used;
}

int orc_sizeof_real(void) { return (int)sizeof(real); }
