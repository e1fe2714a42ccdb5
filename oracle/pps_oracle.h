/*
 * pps_oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, run-time configured restatement of solverPoissonMPI_CPU (the reference
 * is compile-time configured, one binary per problem).  All px*py*pz ranks of a
 * decomposition live in ONE process and are advanced in lock-step, with per-rank
 * partial sums accumulated in the reference's loop order (k, j, i) and combined in
 * rank order 0..n-1 -- the same arithmetic, operation for operation, as the
 * reference built against oracle/mpi_shim (whose Allreduce also sums in rank
 * order), so residual histories and solutions agree BIT FOR BIT with oracle/_ref
 * (pinned by tests/test_oracle.py on the fixtures in tests/golden/, which tests/golden/make_golden.py
 * generates from the unmodified reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * this.  The product (libpps_b200.so) never links or calls it.
 */
#ifndef PPS_ORACLE_H
#define PPS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SOLVER_BICGSTAB = 0, ORC_SOLVER_CG = 1,
       ORC_SOLVER_CHEBYSHEV = 2 /* ChebyshevIteration<..., isMainLoop = true, communicationON, NoneSolver>: chebyshevIteration.hpp:48-140; cheb_max sweeps, no normalisation, no history */ };
enum { ORC_PRECOND_NONE = 0,            /* T_NoneSolver,       inputParam.hpp:24 */
       ORC_PRECOND_CHEBYSHEV = 1,       /* T_Preconditioner2,  inputParam.hpp:28 */
       ORC_PRECOND_BICGSTAB_LOCAL = 2,  /* T_Preconditioner,   inputParam.hpp:31: BiCGSTAB, isMainLoop false, communicationOFF, NoneSolver inside
                                           (with orc_config.precond_comm = 1: communicationON, a global nested solve) */
       ORC_PRECOND_CG_CHEB_LOCAL = 3    /* T_Preconditioner3,  inputParam.hpp:29: BaseCG, isMainLoop false, communicationOFF, Chebyshev inside */ };

typedef struct orc_config {
    int np[3];            /* npglobal            inputParam.hpp:41 */
    int nranks[3];        /* argv px py pz       main.cpp:39-48    */
    double ds[3];         /*                     inputParam.hpp:42 */
    double origin[3];     /*                     inputParam.hpp:43 */
    int bcs[6];           /* 0 Dirichlet 1 Neumann, x- x+ y- y+ z- z+   inputParam.hpp:45 */
    int solver;           /* ORC_SOLVER_*        T_Solver, inputParam.hpp:33 */
    int precond;          /* ORC_PRECOND_*       T_Preconditioner2 / T_NoneSolver */
    double tolerance;     /* tollMainSolver * tollScalingFactor   solverSetup.hpp:22,27 */
    int max_iter;         /* iterMaxMainSolver   solverSetup.hpp:28 */
    int cheb_max;         /* chebyshevMax        solverSetup.hpp:40 */
    double cheb_epsilon;  /* epsilon             solverSetup.hpp:37 */
    double cheb_rescale_min; /* rescaleEigMin    solverSetup.hpp:38 */
    double cheb_rescale_max; /* rescaleEigMax    solverSetup.hpp:39 */
    double precond_tolerance; /* tollPreconditionerSolver * tollScalingFactor   solverSetup.hpp:31 (nested Krylov preconditioners) */
    int precond_max_iter;     /* iterMaxPreconditioner   solverSetup.hpp:32 */
    int order_neumann;        /* orderNeumanBcs: 2 (shipped) or 1   solverSetup.hpp:25 */
    int dim;                  /* DIM: 3 (shipped), 2 or 1   inputParam.hpp:16; axes >= dim hold one point, no guards (blockGrid.hpp:160-206) */
    /* alpaka-only configuration surface (solverPoissonMPI_alpaka).  PINNED: oracle/build_ref_alpaka.py builds that tree unmodified with
     * alpaka's OpenMP CPU accelerator (vendored alpaka 1.2.0 + mdspan, oracle/boost_shim, oracle/mpi_shim); the fp32 preconditioner agrees
     * bit for bit with it, the fp64 one to 2e-15 (tests/golden/alpaka/, tests/test_oracle_alpaka.py) */
    int cheb_eig_local;       /* 1: `local` of solverPoissonMPI_alpaka/include/inputParam.hpp:21-22: block-local, not rescaled eigenvalue bounds */
    int cheb_f32;             /* 1: T_data_chebyshev = float, solverPoissonMPI_alpaka/include/solverSetup.hpp:14 (mixed-precision preconditioner) */
    int precond_comm;         /* 1: the preconditioner runs with communicationON.  ORC_PRECOND_CHEBYSHEV: face exchange of B and of every iterate
                                 (chebyshevIteration.hpp:69-73,97-101) -- a GLOBAL polynomial preconditioner instead of block-Jacobi.
                                 ORC_PRECOND_BICGSTAB_LOCAL: the nested BiCGSTAB is ONE solve over all ranks (exchanges and allreduces inside the
                                 preconditioner; BiCGSTAB.hpp:55-322 with isMainLoop = false, communicationON = true) */
} orc_config;

typedef struct orc_block_info {
    int rank;
    int loc[3];           /* globalLocation_      blockGrid.hpp:151-158 */
    int nlocal[3];        /* nlocal_noguards_     blockGrid.hpp:160-170 */
    int nguards[3];       /* nlocal_guards_       */
    int limits_data[6];   /* indexLimitsData_     blockGrid.hpp:184-206 */
    int limits_solver[6]; /* indexLimitsSolver_   blockGrid.hpp:208-222 */
    int has_boundary[6];  /*                      blockGrid.hpp:234-254 */
    int has_comm[6];      /*                      blockGrid.hpp:256-299 */
    long ntot;            /* points incl. guards  */
} orc_block_info;

typedef struct orc orc_t;

void orc_default_config(orc_config* c);   /* the reference exactly as shipped */
orc_t* orc_create(const orc_config* c);
void orc_destroy(orc_t* o);
int orc_world(const orc_t* o);
void orc_block(const orc_t* o, int rank, orc_block_info* out);
void orc_eigenvalues(const orc_t* o, int rank, double global_minmax[2], double local_minmax[2]);

/* manufactured problem of solverSetup.hpp:44-111 */
double orc_exact_u(double x, double y, double z);
double orc_exact_f(double x, double y, double z);
double orc_exact_dudn(double x, double y, double z, int dir);

/* per-rank guard-padded arrays owned by the oracle (reference layout, x fastest) */
double* orc_x(orc_t* o, int rank);
double* orc_b(orc_t* o, int rank);
void orc_zero_fields(orc_t* o);
void orc_set_problem(orc_t* o);                          /* setProblem: iterativeSolverBase.hpp:51-55 */
/* values of du/dn on the boundary plane of `face` (tangential extent = data range,
 * row-major over the two tangential axes, lower axis fastest); returns the count */
long orc_neumann_face(const orc_t* o, int rank, int face, double* out);

/* building blocks (each follows the reference routine named beside it) */
void orc_apply(const orc_t* o, int rank, const double* in, double* out);     /* matrixFreeOperatorA.hpp:22-39 over the solver range */
void orc_halo_exchange(const orc_t* o, double* const* fields);               /* communicationMPI.hpp:51-292 (faces, data extent) */
void orc_reset_neumann(const orc_t* o, int rank, double* field, int with_bc_value, double norm_b); /* iterativeSolverBase.hpp:62-169 */
void orc_adjust_b(const orc_t* o, int rank, const double* x, double* b);     /* iterativeSolverBase.hpp:429-534 */
void orc_precondition(orc_t* o, double* const* X, double* const* B);         /* chebyshevIteration.hpp:48-140 or noneSolver.hpp:24-27 */

/* the solver: BiCGSTAB.hpp:55-322, baseCG.hpp:44-260 or chebyshevIteration.hpp:48-140 (main loop) on orc_x / orc_b */
int orc_solve(orc_t* o);
int orc_iters(const orc_t* o);
double orc_error_iteration(const orc_t* o);
double orc_error_operator(const orc_t* o);
double orc_norm_b(const orc_t* o);
double orc_loop_seconds(const orc_t* o);
const double* orc_history(const orc_t* o);   /* iters+1 entries */
/* per-iteration scalars of the last solve: alpha, omega (beta for CG), rho; iters entries each */
const double* orc_alpha_history(const orc_t* o);
const double* orc_omega_history(const orc_t* o);
const double* orc_rho_history(const orc_t* o);
/* checkSolutionLocalGlobal: iterativeSolverBase.hpp:283-408; per-rank sum|x-u| and max|x-u| */
void orc_check_solution(orc_t* o, double* sum_abs_per_rank, double* max_abs_per_rank);

#ifdef __cplusplus
}
#endif
#endif
