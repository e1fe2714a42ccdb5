/*
 * pps_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see pps_oracle.h).  Compile with -O2 -ffp-contract=off: the reference binary
 * (g++ -O3, no -march) contains no FMA, and every expression below keeps the
 * reference's operand order so that results are bit-identical to oracle/_ref.
 *
 * Citations are to /root/reference/solverPoissonMPI_CPU/include/<file>:<lines>.
 */
#include "pps_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_PI 3.141592653589793 /* solverSetup.hpp:20 */

typedef struct {
    int rank, loc[3], n[3], ng[3];
    int ld[6], ls[6], hb[6], hc[6];
    long sj, sk, ntot;
    double eig_global[2], eig_local[2];
    double *x, *b;
    double *p, *r, *r0, *Mp, *AMp, *z, *Az; /* BiCGSTAB.hpp:23-29 ; CG uses p, r, AMp(=Apk), z */
    double *cy, *cz, *cw;                   /* chebyshevIteration.hpp:28-30 */
    double *lw[7];                          /* work arrays of a nested (local) Krylov preconditioner: p r r0 Mp AMp z Az */
    double theta, delta, sigma;             /* Chebyshev constants of the preconditioner on this block (global, or the block's own) */
} Block;

struct orc {
    orc_config c;
    int world;
    Block* blk;
    double theta, delta, sigma;             /* chebyshevIteration.hpp:22-26 */
    int iters;
    double err_iter, err_op, norm_b, norm_b_report, loop_seconds;
    double *hist, *h_alpha, *h_omega, *h_rho;
};

/* ---------------------------------------------------------------- manufactured problem */
double orc_exact_f(double x, double y, double z) { return -sin(x) - cos(y) - 3 * sin(z) + 2 * y * z + 2; }            /* solverSetup.hpp:48-50 */
double orc_exact_u(double x, double y, double z) { return sin(x) + cos(y) + 3 * sin(z) + x * x * y * z + x * x + 10; } /* solverSetup.hpp:56-59 */
double orc_exact_dudn(double x, double y, double z, int dir) {                                                          /* solverSetup.hpp:67-110 */
    if (dir == 0) return cos(x) + 2 * x * y * z + 2 * x;
    if (dir == 1) return -sin(y) + x * x * z;
    if (dir == 2) return 3 * cos(z) + x * x * y;
    return -100;
}

void orc_default_config(orc_config* c) {
    const int np[3] = {128, 128, 256};
    const int bcs[6] = {0, 1, 0, 1, 0, 1};
    for (int d = 0; d < 3; d++) { c->np[d] = np[d]; c->nranks[d] = 1; c->ds[d] = 0.1; c->origin[d] = 0; }
    for (int f = 0; f < 6; f++) c->bcs[f] = bcs[f];
    c->solver = ORC_SOLVER_BICGSTAB;
    c->precond = ORC_PRECOND_CHEBYSHEV;
    c->tolerance = 1e2 * 1e-10;
    c->max_iter = 1700;
    c->cheb_max = 11;
    c->cheb_epsilon = 1e-4;
    c->cheb_rescale_min = 500;
    c->cheb_rescale_max = 1 - 1e-4;
    c->precond_tolerance = 1e4 * 1e-10;
    c->precond_max_iter = 150;
    c->order_neumann = 2;
    c->dim = 3;
    c->cheb_eig_local = 0;
    c->cheb_f32 = 0;
    c->precond_comm = 0;
}

/* ---------------------------------------------------------------- geometry (blockGrid.hpp) */
static void eigen_pair(int dim, const double ds[3], const int n[3], double out[2]) {
    /* blockGrid.hpp:301-340 */
    double emin = 0, emax = 0;
    for (int i = 0; i < dim; i++) {
        double lo = 4 * sin(1 * ORC_PI / 2 / (n[i] + 1)) * sin(1 * ORC_PI / 2 / (n[i] + 1)) / (ds[i] * ds[i]);
        double hi = 4 * sin(n[i] * ORC_PI / 2 / (n[i] + 1)) * sin(n[i] * ORC_PI / 2 / (n[i] + 1)) / (ds[i] * ds[i]);
        emin += lo;
        emax += hi;
    }
    out[0] = emin;
    out[1] = emax;
}

static void block_init(const orc_config* c, int rank, Block* B) {
    memset(B, 0, sizeof(*B));
    B->rank = rank;
    B->loc[0] = rank % c->nranks[0];                                   /* blockGrid.hpp:151-158 */
    B->loc[1] = (rank / c->nranks[0]) % c->nranks[1];
    B->loc[2] = rank / (c->nranks[0] * c->nranks[1]);
    for (int d = 0; d < 3; d++) {
        B->n[d] = c->np[d] / c->nranks[d];                             /* :160-170, integer division */
        B->ng[d] = B->n[d] + 2;
        B->ld[2 * d] = 1;                                              /* :184-206 */
        B->ld[2 * d + 1] = B->n[d] + 1;
        const int first = B->loc[d] == 0, last = B->loc[d] == c->nranks[d] - 1, many = c->nranks[d] > 1;
        B->hb[2 * d] = first;                                          /* :234-254 */
        B->hb[2 * d + 1] = last;
        B->hc[2 * d] = many && !first;                                 /* :256-299 */
        B->hc[2 * d + 1] = many && !last;
        B->ls[2 * d] = B->ld[2 * d] + ((c->bcs[2 * d] == 0 && B->hb[2 * d]) ? 1 : 0);             /* :208-222 */
        B->ls[2 * d + 1] = B->ld[2 * d + 1] - ((c->bcs[2 * d + 1] == 0 && B->hb[2 * d + 1]) ? 1 : 0);
        if (d >= c->dim) {                                             /* :166-167,178-179,193-204,237,261: one point, no guards, no faces */
            B->n[d] = 1; B->ng[d] = 1;
            B->ld[2 * d] = 0; B->ld[2 * d + 1] = 1;
            B->ls[2 * d] = 0; B->ls[2 * d + 1] = 1;
            B->hb[2 * d] = B->hb[2 * d + 1] = 0;
            B->hc[2 * d] = B->hc[2 * d + 1] = 0;
        }
    }
    B->sj = B->ng[0];
    B->sk = (long)B->ng[0] * B->ng[1];
    B->ntot = B->sk * B->ng[2];
    int nl[3], ngl[3];
    for (int d = 0; d < 3; d++) {
        nl[d] = B->ls[2 * d + 1] - B->ls[2 * d];
        ngl[d] = c->np[d] - (c->bcs[2 * d] == 0) - (c->bcs[2 * d + 1] == 0);
    }
    eigen_pair(c->dim, c->ds, nl, B->eig_local);
    eigen_pair(c->dim, c->ds, ngl, B->eig_global);
}

static double* zalloc(long n) { return (double*)calloc((size_t)n, sizeof(double)); }

orc_t* orc_create(const orc_config* c) {
    orc_t* o = (orc_t*)calloc(1, sizeof(orc_t));
    o->c = *c;
    if (o->c.order_neumann != 1) o->c.order_neumann = 2;
    if (o->c.dim != 1 && o->c.dim != 2) o->c.dim = 3;
    for (int d = o->c.dim; d < 3; d++) o->c.nranks[d] = 1;                  /* main.cpp:40-48 */
    o->world = c->nranks[0] * c->nranks[1] * c->nranks[2];
    o->blk = (Block*)calloc((size_t)o->world, sizeof(Block));
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        block_init(&o->c, r, B);
        B->x = zalloc(B->ntot); B->b = zalloc(B->ntot);
        B->p = zalloc(B->ntot); B->r = zalloc(B->ntot); B->r0 = zalloc(B->ntot);
        B->Mp = zalloc(B->ntot); B->AMp = zalloc(B->ntot); B->z = zalloc(B->ntot); B->Az = zalloc(B->ntot);
        if (c->precond == ORC_PRECOND_CHEBYSHEV || c->precond == ORC_PRECOND_CG_CHEB_LOCAL || c->solver == ORC_SOLVER_CHEBYSHEV) { B->cy = zalloc(B->ntot); B->cz = zalloc(B->ntot); B->cw = zalloc(B->ntot); }
        if (c->precond == ORC_PRECOND_BICGSTAB_LOCAL || c->precond == ORC_PRECOND_CG_CHEB_LOCAL)
            for (int q = 0; q < 7; q++) B->lw[q] = zalloc(B->ntot);
    }
    const double* eg = o->blk[0].eig_global;
    /* chebyshevIteration.hpp:22-26 (global eigenvalues; delta is negative) */
    o->theta = (eg[0] * c->cheb_rescale_min + eg[1] * c->cheb_rescale_max) * 0.5 * (1.0 + c->cheb_epsilon);
    o->delta = (eg[0] * c->cheb_rescale_min - eg[1] * c->cheb_rescale_max) * 0.5;
    o->sigma = o->theta / o->delta;
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        B->theta = o->theta; B->delta = o->delta;
        if (c->cheb_eig_local) {
            /* alpaka tree, `local` (chebyshevIterationAlpaka.hpp:30-31,71-76): (l0 + l1) / 2 and (l1 - l0) / 2 of the rank's own block,
             * no rescaling, no epsilon; delta keeps the CPU tree's sign here (the alpaka sign is applied where its kernels are restated) */
            B->theta = (B->eig_local[0] + B->eig_local[1]) * 0.5;
            B->delta = (B->eig_local[0] - B->eig_local[1]) * 0.5;
        }
        B->sigma = B->theta / B->delta;
    }
    const long nh = (long)(c->max_iter > c->cheb_max ? c->max_iter : c->cheb_max) + 2;
    o->hist = zalloc(nh); o->h_alpha = zalloc(nh); o->h_omega = zalloc(nh); o->h_rho = zalloc(nh);
    o->norm_b = 1.0;
    o->err_iter = -1.0;
    o->err_op = -1.0;
    return o;
}

void orc_destroy(orc_t* o) {
    if (!o) return;
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        free(B->x); free(B->b); free(B->p); free(B->r); free(B->r0); free(B->Mp); free(B->AMp); free(B->z); free(B->Az);
        free(B->cy); free(B->cz); free(B->cw);
        for (int q = 0; q < 7; q++) free(B->lw[q]);
    }
    free(o->blk); free(o->hist); free(o->h_alpha); free(o->h_omega); free(o->h_rho);
    free(o);
}

int orc_world(const orc_t* o) { return o->world; }
double* orc_x(orc_t* o, int rank) { return o->blk[rank].x; }
double* orc_b(orc_t* o, int rank) { return o->blk[rank].b; }

void orc_block(const orc_t* o, int rank, orc_block_info* out) {
    const Block* B = &o->blk[rank];
    out->rank = rank;
    for (int d = 0; d < 3; d++) { out->loc[d] = B->loc[d]; out->nlocal[d] = B->n[d]; out->nguards[d] = B->ng[d]; }
    for (int f = 0; f < 6; f++) {
        out->limits_data[f] = B->ld[f]; out->limits_solver[f] = B->ls[f];
        out->has_boundary[f] = B->hb[f]; out->has_comm[f] = B->hc[f];
    }
    out->ntot = B->ntot;
}

void orc_eigenvalues(const orc_t* o, int rank, double g[2], double l[2]) {
    g[0] = o->blk[rank].eig_global[0]; g[1] = o->blk[rank].eig_global[1];
    l[0] = o->blk[rank].eig_local[0];  l[1] = o->blk[rank].eig_local[1];
}

/* coordinate of local index i along axis d  (iterativeSolverBase.hpp:547-549) */
static inline double coord(const orc_t* o, const Block* B, int d, int i) {
    return o->c.origin[d] + (i - B->ld[2 * d]) * o->c.ds[d] + B->loc[d] * (B->n[d]) * o->c.ds[d];
}

void orc_zero_fields(orc_t* o) {
    for (int r = 0; r < o->world; r++) {
        memset(o->blk[r].x, 0, sizeof(double) * (size_t)o->blk[r].ntot);
        memset(o->blk[r].b, 0, sizeof(double) * (size_t)o->blk[r].ntot);
    }
}

/* restrict `lim` (a copy of the data limits) to the boundary plane of `face` */
static void face_plane(const Block* B, int face, int lim[6]) {
    const int dir = face / 2;
    for (int q = 0; q < 6; q++) lim[q] = B->ld[q];
    if (face % 2 == 0) lim[2 * dir + 1] = lim[2 * dir] + 1;
    else lim[2 * dir] = lim[2 * dir + 1] - 1;
}

void orc_set_problem(orc_t* o) {
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        /* applyDirichletBCsFromFunction: iterativeSolverBase.hpp:557-603 */
        for (int face = 0; face < 6; face++) {
            if (!(B->hb[face] && o->c.bcs[face] == 0)) continue;
            int lim[6];
            face_plane(B, face, lim);
            for (int k = lim[4]; k < lim[5]; k++)
                for (int j = lim[2]; j < lim[3]; j++)
                    for (int i = lim[0]; i < lim[1]; i++)
                        B->x[i + B->sj * j + B->sk * k] = orc_exact_u(coord(o, B, 0, i), coord(o, B, 1, j), coord(o, B, 2, k));
        }
        /* setFieldValuefromFunction: iterativeSolverBase.hpp:537-555 */
        for (int k = B->ld[4]; k < B->ld[5]; k++)
            for (int j = B->ld[2]; j < B->ld[3]; j++)
                for (int i = B->ld[0]; i < B->ld[1]; i++)
                    B->b[i + B->sj * j + B->sk * k] = orc_exact_f(coord(o, B, 0, i), coord(o, B, 1, j), coord(o, B, 2, k));
    }
}

long orc_neumann_face(const orc_t* o, int rank, int face, double* out) {
    const Block* B = &o->blk[rank];
    int lim[6];
    face_plane(B, face, lim);
    long n = 0;
    for (int k = lim[4]; k < lim[5]; k++)
        for (int j = lim[2]; j < lim[3]; j++)
            for (int i = lim[0]; i < lim[1]; i++)
                out[n++] = orc_exact_dudn(coord(o, B, 0, i), coord(o, B, 1, j), coord(o, B, 2, k), face / 2);
    return n;
}

/* ---------------------------------------------------------------- operator and BCs */
/* matrixFreeOperatorA.hpp:33-38 */
static inline double stencil(const orc_t* o, const Block* B, const double* d, int i, int j, int k) {
    const double* ds = o->c.ds;
    const long sj = B->sj, sk = B->sk;
    if (o->c.dim == 1)                                                            /* :24-27 */
        return (d[i - 1] - 2 * d[i] + d[i + 1]) / (ds[0] * ds[0]);
    if (o->c.dim == 2)                                                            /* :28-32 */
        return (d[i - 1 + sj * j] - 2 * d[i + sj * j] + d[i + 1 + sj * j]) / (ds[0] * ds[0])
             + (d[i + sj * (j - 1)] - 2 * d[i + sj * j] + d[i + sj * (j + 1)]) / (ds[1] * ds[1]);
    return (d[i - 1 + sj * j + sk * k] - 2 * d[i + sj * j + sk * k] + d[i + 1 + sj * j + sk * k]) / (ds[0] * ds[0])
         + (d[i + sj * (j - 1) + sk * k] - 2 * d[i + sj * j + sk * k] + d[i + sj * (j + 1) + sk * k]) / (ds[1] * ds[1])
         + (d[i + sj * j + sk * (k - 1)] - 2 * d[i + sj * j + sk * k] + d[i + sj * j + sk * (k + 1)]) / (ds[2] * ds[2]);
}

#define FOR_SOLVER(B, i, j, k)                         \
    for (int k = (B)->ls[4]; k < (B)->ls[5]; k++)      \
        for (int j = (B)->ls[2]; j < (B)->ls[3]; j++)  \
            for (int i = (B)->ls[0]; i < (B)->ls[1]; i++)

void orc_apply(const orc_t* o, int rank, const double* in, double* out) {
    const Block* B = &o->blk[rank];
    FOR_SOLVER(B, i, j, k) out[i + B->sj * j + B->sk * k] = stencil(o, B, in, i, j, k);
}

/* communicationMPI.hpp:51-292.  Face planes only, tangential extent = data range (the z
 * message of the reference, :252, also drags x-guard cells into edge guards that no
 * 7-point stencil ever reads; they are not reproduced). */
void orc_halo_exchange(const orc_t* o, double* const* f) {
    for (int r = 0; r < o->world; r++) {
        const Block* B = &o->blk[r];
        for (int face = 0; face < 6; face++) {
            if (!B->hc[face]) continue;
            const int dir = face / 2, up = face % 2;
            int nloc[3] = {B->loc[0], B->loc[1], B->loc[2]};
            nloc[dir] += up ? 1 : -1;
            const int other = nloc[0] + nloc[1] * o->c.nranks[0] + nloc[2] * o->c.nranks[0] * o->c.nranks[1];
            const Block* N = &o->blk[other];
            int lim[6];
            face_plane(B, face, lim);
            int off[3] = {0, 0, 0};
            off[dir] = up ? 1 : -1; /* guard plane sits one cell outside my boundary plane */
            for (int k = lim[4]; k < lim[5]; k++)
                for (int j = lim[2]; j < lim[3]; j++)
                    for (int i = lim[0]; i < lim[1]; i++) {
                        int s[3] = {i, j, k};
                        /* neighbour's boundary data plane on the opposite side */
                        s[dir] = up ? N->ld[2 * dir] : N->ld[2 * dir + 1] - 1;
                        f[r][(i + off[0]) + B->sj * (j + off[1]) + B->sk * (k + off[2])] = f[other][s[0] + N->sj * s[1] + N->sk * s[2]];
                    }
        }
    }
}

/* resetNeumanBCs<isMainLoop, fieldData>: iterativeSolverBase.hpp:62-169.  orderNeumanBcs == 2 mirrors the first interior
 * plane (step 2 ds), orderNeumanBcs == 1 copies the boundary plane itself (step ds). */
void orc_reset_neumann(const orc_t* o, int rank, double* field, int with_bc_value, double norm_b) {
    const Block* B = &o->blk[rank];
    for (int dir = 0; dir < 3; dir++) {
        for (int up = 0; up < 2; up++) {
            const int face = 2 * dir + up;
            if (!(B->hb[face] && o->c.bcs[face] == 1)) continue;
            int lim[6];
            face_plane(B, face, lim);
            int adj[3] = {0, 0, 0};
            adj[dir] = up ? -1 : 1;
            for (int k = lim[4]; k < lim[5]; k++)
                for (int j = lim[2]; j < lim[3]; j++)
                    for (int i = lim[0]; i < lim[1]; i++) {
                        const long g = (i - adj[0]) + B->sj * (j - adj[1]) + B->sk * (k - adj[2]);
                        const long m = (i + adj[0]) + B->sj * (j + adj[1]) + B->sk * (k + adj[2]);
                        const long bd = i + B->sj * j + B->sk * k;
                        if (with_bc_value) {
                            const double dn = orc_exact_dudn(coord(o, B, 0, i), coord(o, B, 1, j), coord(o, B, 2, k), dir);
                            if (o->c.order_neumann == 1) {
                                if (!up) field[g] = field[bd] - o->c.ds[dir] * dn / norm_b;  /* :95 */
                                else     field[g] = field[bd] + o->c.ds[dir] * dn / norm_b;  /* :143 */
                            } else {
                                if (!up) field[g] = field[m] - 2 * o->c.ds[dir] * dn / norm_b;   /* :100 */
                                else     field[g] = field[m] + 2 * o->c.ds[dir] * dn / norm_b;   /* :148 */
                            }
                        } else {
                            field[g] = o->c.order_neumann == 1 ? field[bd] : field[m];       /* :108 / :113, :156 / :161 */
                        }
                    }
        }
    }
}

/* adjustFieldBForDirichletNeumanBCs: iterativeSolverBase.hpp:429-534 */
void orc_adjust_b(const orc_t* o, int rank, const double* x, double* b) {
    const Block* B = &o->blk[rank];
    for (int dir = 0; dir < 3; dir++) {
        for (int up = 0; up < 2; up++) {
            const int face = 2 * dir + up;
            if (!B->hb[face]) continue;
            int lim[6];
            face_plane(B, face, lim);
            int adj[3] = {0, 0, 0};
            adj[dir] = up ? -1 : 1;
            const double ds = o->c.ds[dir];
            for (int k = lim[4]; k < lim[5]; k++)
                for (int j = lim[2]; j < lim[3]; j++)
                    for (int i = lim[0]; i < lim[1]; i++) {
                        const long ix = i + B->sj * j + B->sk * k;
                        if (o->c.bcs[face] == 0) {
                            const long ib = (i + adj[0]) + B->sj * (j + adj[1]) + B->sk * (k + adj[2]);
                            b[ib] -= x[ix] / (ds * ds);                                      /* :454, :502 */
                        } else if (o->c.bcs[face] == 1) {
                            const double dn = orc_exact_dudn(coord(o, B, 0, i), coord(o, B, 1, j), coord(o, B, 2, k), dir);
                            if (o->c.order_neumann == 1) {
                                if (!up) b[ix] += dn / ds;                                   /* :475 */
                                else     b[ix] -= dn / ds;                                   /* :522 */
                            } else {
                                if (!up) b[ix] += 2 * dn / ds;                               /* :479 */
                                else     b[ix] -= 2 * dn / ds;                               /* :526 */
                            }
                        }
                    }
        }
    }
}

static double rank_sum(const orc_t* o, const double* partial) {
    /* MPI_Allreduce / Reduce+Bcast as the shim does them: rank order, starting from 0 */
    if (o->world == 1) return partial[0];
    double acc = 0;
    for (int r = 0; r < o->world; r++) acc += partial[r];
    return acc;
}

/* normalizeProblemToFieldBNorm<true, ON>: iterativeSolverBase.hpp:171-234 */
static double normalize_problem(orc_t* o) {
    double* part = (double*)calloc((size_t)o->world, sizeof(double));
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        double* tmp = (double*)malloc(sizeof(double) * (size_t)B->ntot);
        memcpy(tmp, B->b, sizeof(double) * (size_t)B->ntot);
        orc_adjust_b(o, r, B->x, tmp);
        double s = 0.0;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; s += tmp[q] * tmp[q]; }
        part[r] = s;
        free(tmp);
    }
    const double nrm = sqrt(rank_sum(o, part));
    free(part);
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        for (long q = 0; q < B->ntot; q++) { B->x[q] /= nrm; B->b[q] /= nrm; }
    }
    return nrm;
}

/* computeErrorOperatorA<true, ON>: iterativeSolverBase.hpp:236-280 */
static double residual_norm(orc_t* o, double norm_b) {
    double** xs = (double**)malloc(sizeof(double*) * (size_t)o->world);
    double* part = (double*)calloc((size_t)o->world, sizeof(double));
    for (int r = 0; r < o->world; r++) xs[r] = o->blk[r].x;
    orc_halo_exchange(o, xs);
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        orc_reset_neumann(o, r, B->x, 1, norm_b);
        double s = 0.0;
        FOR_SOLVER(B, i, j, k) {
            const long q = i + B->sj * j + B->sk * k;
            B->r[q] = B->b[q] - stencil(o, B, B->x, i, j, k);
            s += B->r[q] * B->r[q];
        }
        part[r] = s;
    }
    const double e = sqrt(rank_sum(o, part));
    free(part);
    free(xs);
    return e;
}

/* ---------------------------------------------------------------- preconditioners */
static void chebyshev_block(orc_t* o, int rank, double* X, double* Bf) {
    /* chebyshevIteration.hpp:48-140 with isMainLoop = false, communicationON = false */
    Block* B = &o->blk[rank];
    const double theta = B->theta, delta = B->delta, sigma = B->sigma;
    double rhoOld = 1 / sigma;
    double rhoCurr = 1 / (2 * sigma - rhoOld);
    orc_reset_neumann(o, rank, Bf, 0, 1.0);
    FOR_SOLVER(B, i, j, k) {
        const long q = i + B->sj * j + B->sk * k;
        B->cz[q] = Bf[q] / theta;
        B->cy[q] = 2 * rhoCurr / delta * (2 * Bf[q] + stencil(o, B, Bf, i, j, k) / theta);
    }
    for (int c = 2; c <= o->c.cheb_max; c++) {
        rhoOld = rhoCurr;
        rhoCurr = 1 / (2 * sigma - rhoOld);
        orc_reset_neumann(o, rank, B->cy, 0, 1.0);
        FOR_SOLVER(B, i, j, k) {
            const long q = i + B->sj * j + B->sk * k;
            B->cw[q] = rhoCurr * (2 * sigma * B->cy[q] + 2 / delta * (Bf[q] + stencil(o, B, B->cy, i, j, k)) - rhoOld * B->cz[q]);
        }
        double* t = B->cz; B->cz = B->cy; B->cy = t;   /* swap(fieldZ, fieldY) */
        t = B->cw; B->cw = B->cy; B->cy = t;           /* swap(fieldW, fieldY) */
    }
    FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; X[q] = (-1) * B->cw[q]; }
}

/* Mixed-precision Chebyshev preconditioner of the alpaka tree (T_data_chebyshev = float), communicationOFF:
 * chebyshevIterationAlpaka.hpp:116-310 (host loop, `cast` branch :149-190, X = -W :293-294) with the kernels
 * CastPrecisionFieldKernel (kernelsAlpakaChebyshev.hpp:8-18), Chebyshev1Kernel (:136-183), Chebyshev2Kernel (:233-270) and
 * AssignFieldWith1FieldKernel (:24-45).  PINNED bit for bit against the unmodified alpaka tree built with its OpenMP CPU accelerator
 * (oracle/build_ref_alpaka.py; tests/golden/alpaka/alp_f32*.npz, tests/test_oracle_alpaka.py).  Every operation is a float operation in the
 * kernels' own order, no contraction (this file is compiled with -ffp-contract=off).
 * The alpaka tree's delta has the opposite sign of the CPU tree's (chebyshevIterationAlpaka.hpp:29), hence -B->delta. */
static void chebyshev_block_alpaka_f32(orc_t* o, int rank, double* X, double* Bf) {
    Block* B = &o->blk[rank];
    /* theta_, delta_ and sigma_ are T_data_chebyshev members (:595-604): theta and delta are evaluated in fp64 and rounded, sigma is the
     * FLOAT quotient of the two rounded members (:34-35) -- pinned against the unmodified alpaka tree, tests/golden/alp_*.npz */
    const float theta = (float)B->theta, delta = (float)(-B->delta), sigma = theta / delta;                     /* :28-35,67-76,595-604 */
    float rhoOld = 1 / sigma;                                                                                  /* :123 */
    float rhoCurr = 1 / (2 * sigma - rhoOld);                                                                  /* :124 */
    orc_reset_neumann(o, rank, Bf, 0, 1.0);                                                                    /* :127, on the fp64 field */
    float* bt = (float*)calloc((size_t)B->ntot, sizeof(float));
    float* y = (float*)calloc((size_t)B->ntot, sizeof(float));
    float* z = (float*)calloc((size_t)B->ntot, sizeof(float));
    float* w = (float*)calloc((size_t)B->ntot, sizeof(float));
    for (long q = 0; q < B->ntot; q++) bt[q] = (float)Bf[q];                                                   /* CastPrecisionFieldKernel */
    /* alpaka vectors are ordered Z Y X: ds[2] of the kernels is dx */
    const float r0 = (float)(o->c.ds[0] * o->c.ds[0]), r1 = (float)(o->c.ds[1] * o->c.ds[1]), r2 = (float)(o->c.ds[2] * o->c.ds[2]);
    const long sj = B->sj, sk = B->sk;
    {
        const float f0 = 2 * rhoCurr / delta * (1 / r0 / theta);                                               /* :148-150 */
        const float f1 = 2 * rhoCurr / delta * (1 / r1 / theta);
        const float f2 = 2 * rhoCurr / delta * (1 / r2 / theta);
        const float fc0 = 2 * rhoCurr / delta * (2 - 2 * (1 / r0 + 1 / r1 + 1 / r2) / theta);                   /* :159 */
        FOR_SOLVER(B, i, j, k) {
            const long q = i + sj * j + sk * k;
            z[q] = bt[q] / theta;                                                                              /* :155 */
            float tmp = bt[q] * fc0 + (bt[q - 1] + bt[q + 1]) * f0;                                            /* :161 */
            tmp += (bt[q - sj] + bt[q + sj]) * f1 + (bt[q - sk] + bt[q + sk]) * f2;                            /* :162-163 */
            y[q] = tmp;
        }
    }
    for (int c = 2; c <= o->c.cheb_max; c++) {                                                                 /* :156-190 */
        rhoOld = rhoCurr;
        rhoCurr = 1 / (2 * sigma - rhoOld);
        /* resetNeumanBCsAlpakaCast<float, false, false>(bufY): plain mirrors on the float field (iterativeSolverBaseAlpaka.hpp:145-186) */
        for (int f = 0; f < 6; f++) {
            if (!(B->hb[f] && o->c.bcs[f] == 1)) continue;
            const int d = f / 2, up = f % 2;
            const int u = d == 0 ? 1 : 0, v = d == 2 ? 1 : 2;
            const long sd = d == 0 ? 1 : (d == 1 ? sj : sk), su = u == 0 ? 1 : sj, sv = v == 1 ? sj : sk;
            const int ghost = up ? B->n[d] + 1 : 0;
            const int src = o->c.order_neumann == 1 ? (up ? B->n[d] : 1) : (up ? B->n[d] - 1 : 2);
            for (int b2 = 1; b2 <= B->n[v]; b2++)
                for (int a2 = 1; a2 <= B->n[u]; a2++) y[ghost * sd + a2 * su + b2 * sv] = y[src * sd + a2 * su + b2 * sv];
        }
        const float f0 = rhoCurr * 2 / delta * (1 / r0);                                                       /* :240-246 */
        const float f1 = rhoCurr * 2 / delta * (1 / r1);
        const float f2 = rhoCurr * 2 / delta * (1 / r2);
        const float fB = rhoCurr * 2 / delta;
        const float fZ = -rhoCurr * rhoOld;
        const float fc0 = rhoCurr * (2 * sigma - 4 / delta * (1 / r0 + 1 / r1 + 1 / r2));                      /* :255 */
        FOR_SOLVER(B, i, j, k) {
            const long q = i + sj * j + sk * k;
            float tmp = y[q] * fc0 + (y[q - 1] + y[q + 1]) * f0;                                               /* :257 */
            tmp += (y[q - sj] + y[q + sj]) * f1 + (y[q - sk] + y[q + sk]) * f2;                                /* :258 */
            tmp += bt[q] * fB;                                                                                 /* :267 */
            tmp += z[q] * fZ;                                                                                  /* :268 */
            w[q] = tmp;
        }
        float* t = z; z = y; y = t;   /* swap(bufZ, bufY)  :184-185 */
        t = w; w = y; y = t;          /* swap(bufW, bufY)  :186-187 */
    }
    FOR_SOLVER(B, i, j, k) { const long q = i + sj * j + sk * k; X[q] = (double)((float)(-1.0) * w[q]); }     /* :293-294 */
    free(bt); free(y); free(z); free(w);
}

/* ---- nested Krylov preconditioners: isMainLoop = false, communicationON = false -> everything is rank-local.
 * X is zeroed, B is normalised by its own norm over the solver range and multiplied back at the end (so the CALLER's
 * vector changes in the last bits, as in the reference), Neumann ghosts are plain mirrors. */
static double local_norm_and_scale(const Block* B, double* X, double* Bf) {
    /* normalizeProblemToFieldBNorm<false, false>: iterativeSolverBase.hpp:171-234 */
    double s = 0.0;
    FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; s += Bf[q] * Bf[q]; }
    const double nrm = sqrt(s);
    for (long q = 0; q < B->ntot; q++) { X[q] /= nrm; Bf[q] /= nrm; }
    return nrm;
}
static double local_residual(const orc_t* o, int rank, double* X, const double* Bf, double* r) {
    /* computeErrorOperatorA<false, false>: iterativeSolverBase.hpp:236-280 */
    const Block* B = &o->blk[rank];
    orc_reset_neumann(o, rank, X, 0, 1.0);
    double s = 0.0;
    FOR_SOLVER(B, i, j, k) {
        const long q = i + B->sj * j + B->sk * k;
        r[q] = Bf[q] - stencil(o, B, X, i, j, k);
        s += r[q] * r[q];
    }
    return sqrt(s);
}

static void local_bicgstab(orc_t* o, int rank, double* X, double* Bf) {
    /* BiCGSTAB.hpp:55-322 with isMainLoop = false, communicationON = false, T_Preconditioner = NoneSolver */
    Block* B = &o->blk[rank];
    double *p = B->lw[0], *r = B->lw[1], *r0 = B->lw[2], *Mp = B->lw[3], *AMp = B->lw[4], *z = B->lw[5], *Az = B->lw[6];
    const size_t nb = sizeof(double) * (size_t)B->ntot;
    for (int q = 0; q < 7; q++) memset(B->lw[q], 0, nb);
    memset(X, 0, nb);                                                              /* :96 */
    const double nrm = local_norm_and_scale(B, X, Bf);
    const double err0 = local_residual(o, rank, X, Bf, r);
    if (err0 < o->c.precond_tolerance) return;                                     /* :118-122: returns without de-normalising */
    memcpy(p, r, nb);
    memcpy(r0, r, nb);
    double alphak = 1, omegak = 1, betak = 1, rho0 = 1, rho1 = 1, err = 0;
    int iter = 0;
    (void)betak;
    while (iter < o->c.precond_max_iter) {
        memcpy(Mp, p, nb);
        orc_reset_neumann(o, rank, Mp, 0, 1.0);
        double s1 = 0.0, s2 = 0.0;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; AMp[q] = stencil(o, B, Mp, i, j, k); s2 += r0[q] * AMp[q]; }
        alphak = rho0 / s2;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; r[q] = r[q] - alphak * AMp[q]; }
        memcpy(z, r, nb);
        orc_reset_neumann(o, rank, z, 0, 1.0);
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; Az[q] = stencil(o, B, z, i, j, k); }
        s1 = 0.0; s2 = 0.0;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; s1 += r[q] * Az[q]; s2 += Az[q] * Az[q]; }
        omegak = s1 / s2;
        for (long q = 0; q < B->ntot; q++) X[q] = X[q] + alphak * Mp[q] + omegak * z[q];
        s1 = 0.0; s2 = 0.0;
        FOR_SOLVER(B, i, j, k) {
            const long q = i + B->sj * j + B->sk * k;
            r[q] = r[q] - omegak * Az[q];
            s1 += r0[q] * r[q];
            s2 += r[q] * r[q];
        }
        err = sqrt(s2);
        rho1 = s1;
        betak = rho1 / rho0 * alphak / omegak;
        rho0 = rho1;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; p[q] = r[q] + betak * (p[q] - omegak * AMp[q]); }
        iter++;
        if (err < o->c.precond_tolerance) break;
    }
    orc_reset_neumann(o, rank, X, 0, 1.0);                                         /* :300 */
    for (long q = 0; q < B->ntot; q++) { X[q] *= nrm; Bf[q] *= nrm; }              /* :310-314 */
}

static void chebyshev_block(orc_t* o, int rank, double* X, double* Bf);

static void local_cg_chebyshev(orc_t* o, int rank, double* X, double* Bf) {
    /* baseCG.hpp:44-260 with isMainLoop = false, communicationON = false, T_Preconditioner = ChebyshevIteration */
    Block* B = &o->blk[rank];
    double *p = B->lw[0], *r = B->lw[1], *Ap = B->lw[2], *z = B->lw[3];
    const size_t nb = sizeof(double) * (size_t)B->ntot;
    for (int q = 0; q < 4; q++) memset(B->lw[q], 0, nb);
    memset(X, 0, nb);
    const double nrm = local_norm_and_scale(B, X, Bf);
    const double err0 = local_residual(o, rank, X, Bf, r);
    if (err0 < o->c.precond_tolerance) return;
    chebyshev_block(o, rank, z, r);
    memcpy(p, z, nb);
    double alphak = 1, betak = 1, s1 = 0, s2 = 0, srk = 1, err = 0;
    int iter = 0;
    while (iter < o->c.precond_max_iter) {
        s1 = 0.0; s2 = 0.0;
        FOR_SOLVER(B, i, j, k) {
            const long q = i + B->sj * j + B->sk * k;
            Ap[q] = stencil(o, B, p, i, j, k);
            s1 += r[q] * z[q];
            s2 += p[q] * Ap[q];
        }
        alphak = s1 / s2;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; X[q] = X[q] + alphak * p[q]; r[q] = r[q] - alphak * Ap[q]; }
        chebyshev_block(o, rank, z, r);
        s2 = 0.0; srk = 0.0;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; s2 += r[q] * z[q]; srk += r[q] * r[q]; }
        betak = s2 / s1;
        err = sqrt(srk);
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; p[q] = z[q] + betak * p[q]; }
        iter++;
        if (err < o->c.precond_tolerance) break;
    }
    for (long q = 0; q < B->ntot; q++) { X[q] *= nrm; Bf[q] *= nrm; }
}

/* ChebyshevIteration<..., isMainLoop = false, communicationON = true, ...> in the preconditioner slot (chebyshevIteration.hpp:48-140):
 * every rank exchanges the faces of B (:69-73) and of every iterate Y (:97-101) before the sweep -- all ranks in lock-step */
static void chebyshev_all_comm(orc_t* o, double* const* X, double* const* Bf) {
    const double theta = o->theta, delta = o->delta, sigma = o->sigma;
    double rhoOld = 1 / sigma;
    double rhoCurr = 1 / (2 * sigma - rhoOld);
    double** ys = (double**)malloc(sizeof(double*) * (size_t)o->world);
    orc_halo_exchange(o, (double* const*)Bf);
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        orc_reset_neumann(o, r, Bf[r], 0, 1.0);
        FOR_SOLVER(B, i, j, k) {
            const long q = i + B->sj * j + B->sk * k;
            B->cz[q] = Bf[r][q] / theta;
            B->cy[q] = 2 * rhoCurr / delta * (2 * Bf[r][q] + stencil(o, B, Bf[r], i, j, k) / theta);
        }
    }
    for (int c = 2; c <= o->c.cheb_max; c++) {
        rhoOld = rhoCurr;
        rhoCurr = 1 / (2 * sigma - rhoOld);
        for (int r = 0; r < o->world; r++) ys[r] = o->blk[r].cy;
        orc_halo_exchange(o, ys);
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            orc_reset_neumann(o, r, B->cy, 0, 1.0);
            FOR_SOLVER(B, i, j, k) {
                const long q = i + B->sj * j + B->sk * k;
                B->cw[q] = rhoCurr * (2 * sigma * B->cy[q] + 2 / delta * (Bf[r][q] + stencil(o, B, B->cy, i, j, k)) - rhoOld * B->cz[q]);
            }
            double* t = B->cz; B->cz = B->cy; B->cy = t;
            t = B->cw; B->cw = B->cy; B->cy = t;
        }
    }
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; X[r][q] = (-1) * B->cw[q]; }
    }
    free(ys);
}

/* BiCGSTAB<..., isMainLoop = false, communicationON = true, NoneSolver> in the preconditioner slot (BiCGSTAB.hpp:55-322; the alpaka tree's
 * T_PreconditionerBiCGStabGlobal, solverPoissonMPI_alpaka/include/inputParam.hpp:33): a GLOBAL nested Krylov solve -- face exchanges of
 * Mp, z and X and rank-ordered allreduces inside the preconditioner, all ranks in lock-step.  Same quirks as the local one: X zeroed (:96),
 * B divided by its (global) norm and multiplied back (:97,310-314), early return without multiplying back (:118-122). */
static void global_bicgstab(orc_t* o, double* const* X, double* const* Bf) {
    const int W = o->world;
    double* s1 = (double*)calloc((size_t)W, sizeof(double));
    double* s2 = (double*)calloc((size_t)W, sizeof(double));
    double** f = (double**)malloc(sizeof(double*) * (size_t)W);
    for (int r = 0; r < W; r++) {
        Block* B = &o->blk[r];
        const size_t nb = sizeof(double) * (size_t)B->ntot;
        for (int q = 0; q < 7; q++) memset(B->lw[q], 0, nb);                         /* :60-66 */
        memset(X[r], 0, nb);                                                         /* :96 */
    }
    /* normalizeProblemToFieldBNorm<false, true>: iterativeSolverBase.hpp:171-234 (Reduce + Bcast = rank-ordered sum) */
    for (int r = 0; r < W; r++) {
        Block* B = &o->blk[r];
        double s = 0.0;
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; s += Bf[r][q] * Bf[r][q]; }
        s1[r] = s;
    }
    const double nrm = sqrt(rank_sum(o, s1));
    for (int r = 0; r < W; r++) {
        Block* B = &o->blk[r];
        for (long q = 0; q < B->ntot; q++) { X[r][q] /= nrm; Bf[r][q] /= nrm; }
    }
    /* computeErrorOperatorA<false, true>: iterativeSolverBase.hpp:236-280 */
    orc_halo_exchange(o, X);
    for (int r = 0; r < W; r++) {
        Block* B = &o->blk[r];
        double* rk = B->lw[1];
        orc_reset_neumann(o, r, X[r], 0, 1.0);
        double s = 0.0;
        FOR_SOLVER(B, i, j, k) {
            const long q = i + B->sj * j + B->sk * k;
            rk[q] = Bf[r][q] - stencil(o, B, X[r], i, j, k);
            s += rk[q] * rk[q];
        }
        s1[r] = s;
    }
    const double err0 = sqrt(rank_sum(o, s1));
    if (err0 < o->c.precond_tolerance) { free(s1); free(s2); free(f); return; }      /* :118-122: returns without de-normalising */
    for (int r = 0; r < W; r++) {
        Block* B = &o->blk[r];
        const size_t nb = sizeof(double) * (size_t)B->ntot;
        memcpy(B->lw[0], B->lw[1], nb);                                              /* p = r    :125 */
        memcpy(B->lw[2], B->lw[1], nb);                                              /* r0 = r   :126 */
    }
    double alphak = 1, omegak = 1, betak = 1, rho0 = 1, rho1 = 1, err = 0;
    int iter = 0;
    (void)betak;
    while (iter < o->c.precond_max_iter) {
        for (int r = 0; r < W; r++) { Block* B = &o->blk[r]; memcpy(B->lw[3], B->lw[0], sizeof(double) * (size_t)B->ntot); f[r] = B->lw[3]; }   /* Mp = p (NoneSolver) */
        orc_halo_exchange(o, f);                                                     /* :135-139 */
        for (int r = 0; r < W; r++) {
            Block* B = &o->blk[r];
            double *r0 = B->lw[2], *Mp = B->lw[3], *AMp = B->lw[4];
            orc_reset_neumann(o, r, Mp, 0, 1.0);
            double s = 0.0;
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; AMp[q] = stencil(o, B, Mp, i, j, k); s += r0[q] * AMp[q]; }
            s2[r] = s;
        }
        alphak = rho0 / rank_sum(o, s2);                                             /* :156-160 */
        for (int r = 0; r < W; r++) {
            Block* B = &o->blk[r];
            double *rk = B->lw[1], *AMp = B->lw[4];
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; rk[q] = rk[q] - alphak * AMp[q]; }
            memcpy(B->lw[5], rk, sizeof(double) * (size_t)B->ntot);                  /* z = r (NoneSolver) */
            f[r] = B->lw[5];
        }
        orc_halo_exchange(o, f);                                                     /* :182-186 */
        for (int r = 0; r < W; r++) {
            Block* B = &o->blk[r];
            double *rk = B->lw[1], *z = B->lw[5], *Az = B->lw[6];
            orc_reset_neumann(o, r, z, 0, 1.0);
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; Az[q] = stencil(o, B, z, i, j, k); }
            double a = 0.0, b = 0.0;
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; a += rk[q] * Az[q]; b += Az[q] * Az[q]; }
            s1[r] = a; s2[r] = b;
        }
        omegak = rank_sum(o, s1) / rank_sum(o, s2);                                  /* :216-221 */
        for (int r = 0; r < W; r++) {
            Block* B = &o->blk[r];
            double *rk = B->lw[1], *r0 = B->lw[2], *Mp = B->lw[3], *z = B->lw[5], *Az = B->lw[6];
            for (long q = 0; q < B->ntot; q++) X[r][q] = X[r][q] + alphak * Mp[q] + omegak * z[q];
            double a = 0.0, b = 0.0;
            FOR_SOLVER(B, i, j, k) {
                const long q = i + B->sj * j + B->sk * k;
                rk[q] = rk[q] - omegak * Az[q];
                a += r0[q] * rk[q];
                b += rk[q] * rk[q];
            }
            s1[r] = a; s2[r] = b;
        }
        rho1 = rank_sum(o, s1);                                                      /* :247-252 */
        err = sqrt(rank_sum(o, s2));
        betak = rho1 / rho0 * alphak / omegak;
        rho0 = rho1;
        for (int r = 0; r < W; r++) {
            Block* B = &o->blk[r];
            double *p = B->lw[0], *rk = B->lw[1], *AMp = B->lw[4];
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; p[q] = rk[q] + betak * (p[q] - omegak * AMp[q]); }
        }
        iter++;
        if (err < o->c.precond_tolerance) break;
    }
    orc_halo_exchange(o, X);                                                         /* :294-298 */
    for (int r = 0; r < W; r++) {
        Block* B = &o->blk[r];
        orc_reset_neumann(o, r, X[r], 0, 1.0);                                       /* :300 */
        for (long q = 0; q < B->ntot; q++) { X[r][q] *= nrm; Bf[r][q] *= nrm; }      /* :310-314 */
    }
    orc_halo_exchange(o, X);                                                         /* :317-321 */
    free(s1); free(s2); free(f);
}

void orc_precondition(orc_t* o, double* const* X, double* const* Bf) {
    if (o->c.precond == ORC_PRECOND_CHEBYSHEV && o->c.precond_comm && !o->c.cheb_f32) { chebyshev_all_comm(o, X, Bf); return; }
    if (o->c.precond == ORC_PRECOND_BICGSTAB_LOCAL && o->c.precond_comm) { global_bicgstab(o, X, Bf); return; }
    for (int r = 0; r < o->world; r++) {
        if (o->c.precond == ORC_PRECOND_CHEBYSHEV && o->c.cheb_f32) chebyshev_block_alpaka_f32(o, r, X[r], Bf[r]);
        else if (o->c.precond == ORC_PRECOND_CHEBYSHEV) chebyshev_block(o, r, X[r], Bf[r]);
        else if (o->c.precond == ORC_PRECOND_BICGSTAB_LOCAL) local_bicgstab(o, r, X[r], Bf[r]);
        else if (o->c.precond == ORC_PRECOND_CG_CHEB_LOCAL) local_cg_chebyshev(o, r, X[r], Bf[r]);
        else memcpy(X[r], Bf[r], sizeof(double) * (size_t)o->blk[r].ntot);   /* noneSolver.hpp:24-27 */
    }
}

/* ---------------------------------------------------------------- solvers */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct { double **x, **p, **r, **Mp, **z; double *s1, *s2; } Ptrs;

static Ptrs ptrs_make(orc_t* o) {
    Ptrs P;
    const size_t n = (size_t)o->world;
    P.x = (double**)malloc(sizeof(double*) * n); P.p = (double**)malloc(sizeof(double*) * n);
    P.r = (double**)malloc(sizeof(double*) * n); P.Mp = (double**)malloc(sizeof(double*) * n);
    P.z = (double**)malloc(sizeof(double*) * n);
    P.s1 = (double*)calloc(n, sizeof(double)); P.s2 = (double*)calloc(n, sizeof(double));
    for (int r = 0; r < o->world; r++) {
        P.x[r] = o->blk[r].x; P.p[r] = o->blk[r].p; P.r[r] = o->blk[r].r; P.Mp[r] = o->blk[r].Mp; P.z[r] = o->blk[r].z;
    }
    return P;
}
static void ptrs_free(Ptrs* P) { free(P->x); free(P->p); free(P->r); free(P->Mp); free(P->z); free(P->s1); free(P->s2); }

static void clear_work(orc_t* o) {
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        const size_t nb = sizeof(double) * (size_t)B->ntot;
        memset(B->p, 0, nb); memset(B->r, 0, nb); memset(B->r0, 0, nb); memset(B->Mp, 0, nb);
        memset(B->AMp, 0, nb); memset(B->z, 0, nb); memset(B->Az, 0, nb);
    }
}

/* tail shared by both solvers: BiCGSTAB.hpp:293-321 / baseCG.hpp:231-259 */
static void finish_solve(orc_t* o, Ptrs* P, double t_start, int reset_neumann_x) {
    orc_halo_exchange(o, P->x);
    if (reset_neumann_x) for (int r = 0; r < o->world; r++) orc_reset_neumann(o, r, o->blk[r].x, 1, o->norm_b);
    o->loop_seconds = now_s() - t_start;
    o->err_op = residual_norm(o, o->norm_b);
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        for (long q = 0; q < B->ntot; q++) { B->x[q] *= o->norm_b; B->b[q] *= o->norm_b; }
    }
    o->norm_b = 1;
    orc_halo_exchange(o, P->x);
}

static int solve_bicgstab(orc_t* o) {
    /* BiCGSTAB.hpp:55-322 with isMainLoop = true, communicationON = true */
    Ptrs P = ptrs_make(o);
    const double tol = o->c.tolerance;
    int iter = 0;
    double alphak = 1, betak = 1, omegak = 1, rho0 = 1, rho1 = 1;
    (void)betak;
    clear_work(o);
    orc_halo_exchange(o, P.x);
    for (int r = 0; r < o->world; r++) orc_reset_neumann(o, r, o->blk[r].x, 1, o->norm_b /* == 1 here */);
    o->norm_b = normalize_problem(o);
    o->norm_b_report = o->norm_b;
    o->err_op = residual_norm(o, o->norm_b);
    o->hist[0] = o->err_op;
    if (o->err_op < tol) { o->iters = 0; ptrs_free(&P); return 0; }
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        memcpy(B->p, B->r, sizeof(double) * (size_t)B->ntot);
        memcpy(B->r0, B->r, sizeof(double) * (size_t)B->ntot);
    }
    rho0 = 1;
    rho1 = rho0;
    const double t0 = now_s();
    while (iter < o->c.max_iter) {
        orc_precondition(o, P.Mp, P.p);                                           /* :133 */
        orc_halo_exchange(o, P.Mp);                                               /* :135-139 */
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            orc_reset_neumann(o, r, B->Mp, 0, 1.0);                               /* :140 */
            double s2 = 0.0;
            FOR_SOLVER(B, i, j, k) {                                              /* :142-155 */
                const long q = i + B->sj * j + B->sk * k;
                B->AMp[q] = stencil(o, B, B->Mp, i, j, k);
                s2 += B->r0[q] * B->AMp[q];
            }
            P.s2[r] = s2;
        }
        alphak = rho0 / rank_sum(o, P.s2);                                        /* :156-164 */
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; B->r[q] = B->r[q] - alphak * B->AMp[q]; }  /* :168-178 */
        }
        orc_precondition(o, P.z, P.r);                                            /* :181 */
        orc_halo_exchange(o, P.z);                                                /* :182-186 */
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            orc_reset_neumann(o, r, B->z, 0, 1.0);                                /* :188 */
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; B->Az[q] = stencil(o, B, B->z, i, j, k); }  /* :189-199 */
            double s1 = 0.0, s2 = 0.0;
            FOR_SOLVER(B, i, j, k) {                                              /* :201-214 */
                const long q = i + B->sj * j + B->sk * k;
                s1 += B->r[q] * B->Az[q];
                s2 += B->Az[q] * B->Az[q];
            }
            P.s1[r] = s1;
            P.s2[r] = s2;
        }
        omegak = rank_sum(o, P.s1) / rank_sum(o, P.s2);                           /* :216-225 */
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            for (long q = 0; q < B->ntot; q++) B->x[q] = B->x[q] + alphak * B->Mp[q] + omegak * B->z[q];                 /* :227-230 */
            double s1 = 0.0, s2 = 0.0;
            FOR_SOLVER(B, i, j, k) {                                              /* :232-246 */
                const long q = i + B->sj * j + B->sk * k;
                B->r[q] = B->r[q] - omegak * B->Az[q];
                s1 += B->r0[q] * B->r[q];
                s2 += B->r[q] * B->r[q];
            }
            P.s1[r] = s1;
            P.s2[r] = s2;
        }
        rho1 = rank_sum(o, P.s1);                                                 /* :247-257 */
        o->err_iter = sqrt(rank_sum(o, P.s2));
        betak = rho1 / rho0 * alphak / omegak;                                    /* :258 */
        rho0 = rho1;
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            FOR_SOLVER(B, i, j, k) {                                              /* :262-272 */
                const long q = i + B->sj * j + B->sk * k;
                B->p[q] = B->r[q] + betak * (B->p[q] - omegak * B->AMp[q]);
            }
        }
        o->h_alpha[iter] = alphak; o->h_omega[iter] = omegak; o->h_rho[iter] = rho0;
        iter++;
        o->hist[iter] = o->err_iter;                                              /* :280 */
        if (o->err_iter < tol) break;                                             /* :288 */
    }
    o->iters = iter;
    finish_solve(o, &P, t0, 1);
    ptrs_free(&P);
    return 0;
}

static int solve_cg(orc_t* o) {
    /* baseCG.hpp:44-260 with isMainLoop = true, communicationON = true */
    Ptrs P = ptrs_make(o);
    const double tol = o->c.tolerance;
    int iter = 0;
    double alphak = 1, betak = 1, tot1 = 0, tot2 = 0, totRk = 1;
    clear_work(o);
    orc_halo_exchange(o, P.x);
    for (int r = 0; r < o->world; r++) orc_reset_neumann(o, r, o->blk[r].x, 1, o->norm_b);
    o->norm_b = normalize_problem(o);
    o->norm_b_report = o->norm_b;
    o->err_op = residual_norm(o, o->norm_b);
    o->hist[0] = o->err_op;
    if (o->err_op < tol) { o->iters = 0; ptrs_free(&P); return 0; }
    orc_precondition(o, P.z, P.r);                                                /* :109 */
    for (int r = 0; r < o->world; r++) memcpy(o->blk[r].p, o->blk[r].z, sizeof(double) * (size_t)o->blk[r].ntot);
    const double t0 = now_s();
    while (iter < o->c.max_iter) {
        orc_halo_exchange(o, P.p);                                                /* :118-122 */
        if (o->c.order_neumann == 1)                                              /* :123-124: only for order 1 */
            for (int r = 0; r < o->world; r++) orc_reset_neumann(o, r, o->blk[r].p, 0, 1.0);
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            double s1 = 0.0, s2 = 0.0;
            FOR_SOLVER(B, i, j, k) {                                              /* :126-140 */
                const long q = i + B->sj * j + B->sk * k;
                B->AMp[q] = stencil(o, B, B->p, i, j, k);
                s1 += B->r[q] * B->z[q];
                s2 += B->p[q] * B->AMp[q];
            }
            P.s1[r] = s1;
            P.s2[r] = s2;
        }
        tot1 = rank_sum(o, P.s1);
        tot2 = rank_sum(o, P.s2);
        alphak = tot1 / tot2;                                                     /* :145 */
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            FOR_SOLVER(B, i, j, k) {                                              /* :154-165 */
                const long q = i + B->sj * j + B->sk * k;
                B->x[q] = B->x[q] + alphak * B->p[q];
                B->r[q] = B->r[q] - alphak * B->AMp[q];
            }
        }
        orc_precondition(o, P.z, P.r);                                            /* :168 */
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            double s2 = 0.0, srk = 0.0;
            FOR_SOLVER(B, i, j, k) {                                              /* :171-182 */
                const long q = i + B->sj * j + B->sk * k;
                s2 += B->r[q] * B->z[q];
                srk += B->r[q] * B->r[q];
            }
            P.s1[r] = srk;
            P.s2[r] = s2;
        }
        tot2 = rank_sum(o, P.s2);
        totRk = rank_sum(o, P.s1);
        betak = tot2 / tot1;                                                      /* :187 */
        o->err_iter = sqrt(totRk);
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; B->p[q] = B->z[q] + betak * B->p[q]; }  /* :197-208 */
        }
        o->h_alpha[iter] = alphak; o->h_omega[iter] = betak; o->h_rho[iter] = tot2;
        iter++;
        o->hist[iter] = o->err_iter;
        if (o->err_iter < tol) break;
    }
    o->iters = iter;
    finish_solve(o, &P, t0, o->c.order_neumann == 1);                             /* :237-238 */
    ptrs_free(&P);
    return 0;
}

static int solve_chebyshev(orc_t* o) {
    /* chebyshevIteration.hpp:48-140 with isMainLoop = true, communicationON = true: cheb_max sweeps on the BC-adjusted
     * copy of b with a halo exchange before every sweep; no normalisation (normFieldB_ stays 1), no residual history,
     * x is written on the solver range only and neither de-normalised nor exchanged afterwards. */
    const double theta = o->theta, delta = o->delta, sigma = o->sigma;
    double rhoOld = 1 / sigma;
    double rhoCurr = 1 / (2 * sigma - rhoOld);
    double** bt = (double**)malloc(sizeof(double*) * (size_t)o->world);
    double** ys = (double**)malloc(sizeof(double*) * (size_t)o->world);
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        bt[r] = (double*)malloc(sizeof(double) * (size_t)B->ntot);
        memcpy(bt[r], B->b, sizeof(double) * (size_t)B->ntot);                    /* :63-65 */
        orc_adjust_b(o, r, B->x, bt[r]);                                          /* :66 */
    }
    orc_halo_exchange(o, bt);                                                     /* :69-73 */
    for (int r = 0; r < o->world; r++) orc_reset_neumann(o, r, bt[r], 0, 1.0);    /* :74: fieldData = false -> plain mirror */
    o->norm_b_report = 1.0;
    const double t0 = now_s();
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        FOR_SOLVER(B, i, j, k) {                                                  /* :79-91 */
            const long q = i + B->sj * j + B->sk * k;
            B->cz[q] = bt[r][q] / theta;
            B->cy[q] = 2 * rhoCurr / delta * (2 * bt[r][q] + stencil(o, B, bt[r], i, j, k) / theta);
        }
    }
    for (int c = 2; c <= o->c.cheb_max; c++) {                                    /* :94-116 */
        rhoOld = rhoCurr;
        rhoCurr = 1 / (2 * sigma - rhoOld);
        for (int r = 0; r < o->world; r++) ys[r] = o->blk[r].cy;
        orc_halo_exchange(o, ys);
        for (int r = 0; r < o->world; r++) {
            Block* B = &o->blk[r];
            orc_reset_neumann(o, r, B->cy, 0, 1.0);
            FOR_SOLVER(B, i, j, k) {
                const long q = i + B->sj * j + B->sk * k;
                B->cw[q] = rhoCurr * (2 * sigma * B->cy[q] + 2 / delta * (bt[r][q] + stencil(o, B, B->cy, i, j, k)) - rhoOld * B->cz[q]);
            }
            double* t = B->cz; B->cz = B->cy; B->cy = t;
            t = B->cw; B->cw = B->cy; B->cy = t;
        }
    }
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        FOR_SOLVER(B, i, j, k) { const long q = i + B->sj * j + B->sk * k; B->x[q] = (-1) * B->cw[q]; }   /* :118-128 */
        free(bt[r]);
    }
    o->loop_seconds = now_s() - t0;
    free(bt);
    free(ys);
    o->err_op = residual_norm(o, 1.0);                                            /* :134-135 */
    o->err_iter = o->err_op;                                                      /* :136 */
    o->iters = o->c.cheb_max;                                                     /* :137 */
    o->hist[0] = o->err_op;
    return 0;
}

int orc_solve(orc_t* o) {
    o->norm_b = 1.0;
    if (o->c.solver == ORC_SOLVER_CHEBYSHEV) return solve_chebyshev(o);
    return o->c.solver == ORC_SOLVER_CG ? solve_cg(o) : solve_bicgstab(o);
}

int orc_iters(const orc_t* o) { return o->iters; }
double orc_error_iteration(const orc_t* o) { return o->err_iter; }
double orc_error_operator(const orc_t* o) { return o->err_op; }
double orc_norm_b(const orc_t* o) { return o->norm_b_report; }
double orc_loop_seconds(const orc_t* o) { return o->loop_seconds; }
const double* orc_history(const orc_t* o) { return o->hist; }
const double* orc_alpha_history(const orc_t* o) { return o->h_alpha; }
const double* orc_omega_history(const orc_t* o) { return o->h_omega; }
const double* orc_rho_history(const orc_t* o) { return o->h_rho; }

void orc_check_solution(orc_t* o, double* sum_abs, double* max_abs) {
    /* checkSolutionLocalGlobal: iterativeSolverBase.hpp:283-408 (normFieldB_ is 1 after a solve) */
    double** xs = (double**)malloc(sizeof(double*) * (size_t)o->world);
    for (int r = 0; r < o->world; r++) xs[r] = o->blk[r].x;
    orc_halo_exchange(o, xs);
    free(xs);
    for (int r = 0; r < o->world; r++) {
        Block* B = &o->blk[r];
        orc_reset_neumann(o, r, B->x, 1, o->norm_b);
        double el = 0, em = -1;
        for (int k = B->ld[4]; k < B->ld[5]; k++)
            for (int j = B->ld[2]; j < B->ld[3]; j++)
                for (int i = B->ld[0]; i < B->ld[1]; i++) {
                    const double u = orc_exact_u(coord(o, B, 0, i), coord(o, B, 1, j), coord(o, B, 2, k));
                    const double e = fabs(B->x[i + B->sj * j + B->sk * k] - u);
                    el += e;
                    if (e > em) em = e;
                }
        sum_abs[r] = el;
        max_abs[r] = em;
    }
}
