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
    /* residual tracking: the Newton system is solved exactly, so the stationarity and
     * dynamics residuals contract by (1-alpha) each iteration; res_lin is that factor times
     * the initial residual bound (checked explicitly below for the slack equations). */
    real res_lin = 1;
    for (it = 0; it <= c->max_iter; it++) {
        mu = 0; res = 0;
        for (int i = 0; i < nb; i++) {
            mu += ll[i] * tl[i] + lu[i] * tu[i];
            res = fmax(res, fabs(tl[i] - (zb[i] - lb[i])));
            res = fmax(res, fabs(tu[i] - (ub[i] - zb[i])));
        }
        mu /= (2 * nb);
        if (it > 0 && res < tol && mu < (real)c->tol_mu && res_lin < tol) { status = 0; break; }
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
        if (getenv("ORC_TRACE")) fprintf(stderr, "it %d mu %.3e res %.3e alpha %.4f res_lin %.3e sigma_mu %.3e\n", it, (double)mu, (double)res, (double)alpha, (double)res_lin, (double)sigma_mu);
    }
done:;
    int nact = 0;
    for (int i = 0; i < nb; i++) { nact += (tl[i] < ll[i]); nact += (tu[i] < lu[i]); }
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
