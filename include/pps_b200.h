/*
 * pps_b200.h -- C ABI of the B200-native Poisson hot path (libpps_b200.so).
 *
 * The reference (lucak17/ParallelPoissonSolver) has no FFI: its seam is the solver-class
 * concept used by solverPoissonMPI_CPU/src/main.cpp:83-126
 *     T_Solver solver(blockGrid, exactSolutionAndBCs, communicator);   main.cpp:83
 *     solver.setProblem(fieldX, fieldB);                               main.cpp:94
 *     solver(fieldX, fieldB, operatorA);                               main.cpp:99
 *     solver.getNumIterationFinal() / getErrorFromIteration() / ...    main.cpp:107-108
 * Every entry point below names the reference member it replaces.  Plain pointers and
 * sizes only; caller-owned HOST arrays use the reference's layout: dense fp64,
 * (nx+2)(ny+2)(nz+2) local points with one guard layer per side, x fastest,
 * idx = i + (nx+2)*j + (nx+2)(ny+2)*k   (blockGrid.hpp:172-182, matrixFreeOperatorA.hpp:18-19).
 *
 * A handle owns one or more BLOCKS (= reference MPI ranks, blockGrid.hpp:151-170) on ONE GPU:
 *   - world_size == 1: the handle hosts all px*py*pz blocks of the decomposition on its GPU
 *     ("virtual ranks": same arithmetic as a px*py*pz-rank reference run, one device);
 *   - world_size == px*py*pz: one process (or thread) per GPU, block `rank` lives here and
 *     faces / scalar sums travel over NCCL (NVLink) -- the replacement of CommunicatorMPI
 *     (communicationMPI.hpp:51-316) and of the inline MPI_Allreduce calls
 *     (BiCGSTAB.hpp:158,218-219,249-250).
 * All functions return 0 on success, non-zero on error (pps_last_error() explains); there is
 * no CPU fallback: without a CUDA device pps_create fails.
 * Not re-entrant per handle; distinct handles may be driven from distinct threads.
 */
#ifndef PPS_B200_H
#define PPS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPS_ABI_VERSION 1
#define PPS_UNIQUE_ID_BYTES 128

/* main solver: T_Solver of inputParam.hpp:33 */
enum { PPS_SOLVER_BICGSTAB = 0,  /* BiCGSTAB.hpp  */
       PPS_SOLVER_CG = 1,        /* baseCG.hpp    */
       PPS_SOLVER_CHEBYSHEV = 2  /* chebyshevIteration.hpp:48-140 with isMainLoop = true, communicationON: cheb_max_iter sweeps,
                                    no normalisation, no residual history (only the final ||b - A x||); precond must be NONE */ };
/* preconditioner slot: T_NoneSolver / T_Preconditioner2 / T_Preconditioner / T_Preconditioner3 of inputParam.hpp:24-31 */
enum { PPS_PRECOND_NONE = 0,            /* noneSolver.hpp */
       PPS_PRECOND_CHEBYSHEV = 1,       /* chebyshevIteration.hpp, communicationOFF (block-Jacobi) */
       PPS_PRECOND_BICGSTAB_LOCAL = 2,  /* BiCGSTAB<.., isMainLoop false, communicationOFF, NoneSolver>   inputParam.hpp:31 */
       PPS_PRECOND_CG_CHEB_LOCAL = 3    /* BaseCG<.., isMainLoop false, communicationOFF, Chebyshev>      inputParam.hpp:29 */ };
/* arithmetic of the operator: FAST = precomputed 1/ds^2 and FMA; PARITY = the reference's
 * expression order with IEEE division and no contraction (bit-identical per point to the
 * g++ build of matrixFreeOperatorA.hpp:33-38; slower, for kernel parity tests) */
enum { PPS_ARITH_FAST = 0, PPS_ARITH_PARITY = 1 };
/* kernel schedule of one Krylov iteration */
enum { PPS_FUSE_AUTO = 0,   /* best measured schedule */
       PPS_FUSE_SPLIT = 1,  /* one kernel per reference loop nest group (19 vector passes / BiCGSTAB iteration) */
       PPS_FUSE_FULL = 2    /* axpy updates fused into the operator kernels (17 passes) */ };

/* eigenvalue bounds of the Chebyshev preconditioner: GLOBAL = whole grid, rescaled by cheb_rescale_min/max and (1 + epsilon)
 * (chebyshevIteration.hpp:22-26); LOCAL = the block's own bounds, not rescaled (alpaka tree, chebyshevIterationAlpaka.hpp:30-31) */
enum { PPS_CHEB_EIG_GLOBAL = 0, PPS_CHEB_EIG_LOCAL = 1 };
/* precision of the Chebyshev iterates: FP32 = mixed-precision preconditioner (B is cast to float, the sweeps run in float with
 * the alpaka tree's folded 7-point form, X = -(float)W widened to double; kernelsAlpakaChebyshev.hpp:8-18,136-183,233-270) */
enum { PPS_CHEB_FP64 = 0, PPS_CHEB_FP32 = 1 };

typedef struct pps_config {
    int abi_version;          /* PPS_ABI_VERSION */
    int dim;                  /* DIM, inputParam.hpp:16 (3) */
    int npglobal[3];          /* inputParam.hpp:41 */
    int nranks[3];            /* argv px py pz, main.cpp:39-48 */
    double ds[3];             /* inputParam.hpp:42 */
    double origin[3];         /* inputParam.hpp:43 */
    int guards[3];            /* inputParam.hpp:44 (1,1,1) */
    int bcs_type[6];          /* inputParam.hpp:45: 0 Dirichlet, 1 Neumann; x- x+ y- y+ z- z+ */
    int solver;               /* PPS_SOLVER_* */
    int precond;              /* PPS_PRECOND_* */
    double tolerance;         /* tollMainSolver * tollScalingFactor, solverSetup.hpp:22,27 */
    int max_iter;             /* iterMaxMainSolver, solverSetup.hpp:28 */
    int cheb_max_iter;        /* chebyshevMax, solverSetup.hpp:40 */
    double cheb_epsilon;      /* epsilon, solverSetup.hpp:37 */
    double cheb_rescale_min;  /* rescaleEigMin, solverSetup.hpp:38 */
    double cheb_rescale_max;  /* rescaleEigMax, solverSetup.hpp:39 */
    int order_neumann;        /* orderNeumanBcs, solverSetup.hpp:25: 2 (shipped) or 1 */
    int arithmetic;           /* PPS_ARITH_* */
    int fusion;               /* PPS_FUSE_* */
    int device;               /* CUDA device ordinal, -1 = current */
    int flags;                /* PPS_FLAG_* */
    int precond_max_iter;     /* iterMaxPreconditioner, solverSetup.hpp:32 (nested Krylov preconditioners); 0 = 150 */
    double precond_tolerance; /* tollPreconditionerSolver * tollScalingFactor, solverSetup.hpp:31; 0 = 1e4 * 1e-10 */
    /* alpaka-only configuration surface; 0 = the CPU tree's behaviour */
    int cheb_eigenvalues;     /* PPS_CHEB_EIG_*: `global` / `local` of solverPoissonMPI_alpaka/include/inputParam.hpp:21-22,27-29 */
    int cheb_precision;       /* PPS_CHEB_FP*: T_data_chebyshev of solverPoissonMPI_alpaka/include/solverSetup.hpp:14 */
    int cheb_block;           /* Chebyshev sweeps advanced per HBM pass by the temporally blocked kernel (0 = one kernel per sweep, 1..4) */
    int precond_communication; /* communicationON / OFF template argument of the preconditioner (inputParam.hpp:20-21,28): 0 = OFF, block-Jacobi
                                  (shipped); 1 = ON: the Chebyshev preconditioner exchanges the faces of B and of every iterate
                                  (chebyshevIteration.hpp:69-73,97-101) -- a global polynomial preconditioner; with
                                  PPS_PRECOND_BICGSTAB_LOCAL the nested BiCGSTAB becomes ONE solve over all blocks / GPUs (face exchanges
                                  and allreduces inside the preconditioner, BiCGSTAB.hpp:135-139,156-164,182-186,216-225,247-257 with
                                  isMainLoop = false; solverPoissonMPI_alpaka/include/inputParam.hpp:33 T_PreconditionerBiCGStabGlobal) */
} pps_config;

/* pps_config.flags */
#define PPS_FLAG_OPERATOR_ONLY 1   /* allocate only what pps_bench_operator / pps_apply_operator need (3 vectors instead of 7+):
                                      the operator-apply bandwidth sweep up to the largest grid that fits the GPU; pps_solve fails */

#define PPS_FLAG_NO_DOT_VECTOR 2    /* with PPS_FLAG_OPERATOR_ONLY: 2 vectors only (x, y): pps_bench_operator without the fused dot, for the
                                      largest grid that fits (2 x 8 B x 2176^3 = 165 GB) */

typedef struct pps_block_info {   /* BlockGrid getters, blockGrid.hpp:40-145 */
    int rank;
    int global_location[3];
    int nlocal_noguards[3];
    int nlocal_guards[3];
    int limits_data[6];
    int limits_solver[6];
    int has_boundary[6];
    int has_communication[6];
    long long ntot_guards;
} pps_block_info;

typedef struct pps_handle pps_handle;

const char* pps_last_error(void);
int pps_version(void);
void pps_default_config(pps_config* cfg);   /* inputParam.hpp / solverSetup.hpp as shipped */

/* NCCL bootstrap for world_size > 1: rank 0 calls pps_get_unique_id and ships the bytes to
 * the other ranks by any means (torch.distributed, a file, MPI_Bcast in the reference driver). */
int pps_get_unique_id(unsigned char id[PPS_UNIQUE_ID_BYTES]);

/* replaces: BlockGrid ctor (main.cpp:58) + CommunicatorMPI ctor (main.cpp:78) + T_Solver ctor
 * (main.cpp:83; BiCGSTAB.hpp:18-39 allocates the work arrays).  unique_id may be NULL iff world_size == 1. */
int pps_create(const pps_config* cfg, int rank, int world_size, const unsigned char* unique_id, pps_handle** out);
int pps_destroy(pps_handle* h);                                   /* ~T_Solver, BiCGSTAB.hpp:41-51 */

int pps_num_local_blocks(const pps_handle* h);                    /* px*py*pz if world_size==1, else 1 */
int pps_block_info_get(const pps_handle* h, int rank, pps_block_info* out);
int pps_eigenvalues(const pps_handle* h, int rank, double global_min_max[2], double local_min_max[2]); /* blockGrid.hpp:301-340 */

/* replaces the hand-over of the caller's arrays to solver(fieldX, fieldB, ...) (main.cpp:94-99):
 * x_host = initial guess with Dirichlet boundary planes filled (applyDirichletBCsFromFunction,
 * iterativeSolverBase.hpp:557-603), b_host = right-hand side on the data range
 * (setFieldValuefromFunction, :537-555).  Copied to the device (pitched layout). */
int pps_set_fields(pps_handle* h, int rank, const double* x_host, const double* b_host);
/* values of ExactSolutionAndBCs::trueSolutionDdir (solverSetup.hpp:61-85) on the boundary plane of
 * `face` (0..5), tangential extent = data range, lower axis fastest; needed for Neumann faces by
 * resetNeumanBCs (iterativeSolverBase.hpp:100,148) and adjustFieldBForDirichletNeumanBCs (:480,527). */
int pps_set_neumann_face(pps_handle* h, int rank, int face, const double* dudn_host, size_t count);

/* replaces T_Solver::operator()(fieldX, fieldB, operatorA) (main.cpp:99; BiCGSTAB.hpp:55-322,
 * baseCG.hpp:44-260).  Collective over all ranks when world_size > 1. */
int pps_solve(pps_handle* h);
/* keep / restore a device-side copy of the fields given to pps_set_fields (repeat solves without H2D) */
int pps_save_fields(pps_handle* h);
int pps_restore_fields(pps_handle* h);

/* x after the solve: de-normalised, guards refreshed (BiCGSTAB.hpp:294-321) */
int pps_get_solution(pps_handle* h, int rank, double* x_host);
int pps_get_rhs(pps_handle* h, int rank, double* b_host);

int pps_get_iterations(const pps_handle* h);                 /* getNumIterationFinal, iterativeSolverBase.hpp:422-425 */
/* iterations of a nested Krylov preconditioner during the last solve, summed over its calls and over the local blocks
 * (getNumIterationPreconditionerFinal of the alpaka tree, iterativeSolverBaseAlpaka.hpp:615-618); 0 for the others */
long long pps_get_preconditioner_iterations(const pps_handle* h);
double pps_get_error_iteration(const pps_handle* h);         /* getErrorFromIteration,  :414-417 */
double pps_get_error_operator(const pps_handle* h);          /* getErrorComputeOperator, :418-421 */
double pps_get_norm_b(const pps_handle* h);                  /* normFieldB_ as printed at BiCGSTAB.hpp:108 */
double pps_get_solver_seconds(const pps_handle* h);          /* whole operator(): "Solver time", main.cpp:96-101,122 */
double pps_get_loop_seconds(const pps_handle* h);            /* getDurationSolver: "SolverInFunction time", BiCGSTAB.hpp:129,302-303 */
/* errorFromIterationHistory_[0..iters] (iterativeSolverBase.hpp:43; BiCGSTAB.hpp:114-117,278-281);
 * which: 0 residual, 1 alpha, 2 omega (beta for CG), 3 rho0 -- the columns of BiCGSTAB.hpp:285 */
int pps_get_history(const pps_handle* h, int which, double* out, int capacity);
/* checkSolutionLocalGlobal (iterativeSolverBase.hpp:283-408): sum|x-u| and max|x-u| over the data range
 * of block `rank` against u_exact_host (reference layout) */
int pps_check_solution(pps_handle* h, int rank, const double* u_exact_host, double* sum_abs, double* max_abs);

/* ---- building blocks, exported for parity tests and the stencil bandwidth sweep (BASELINE config 5) ---- */
/* out = A*in on the solver range of block `rank` (matrixFreeOperatorA.hpp:22-39 driven by the loop nest of
 * BiCGSTAB.hpp:189-199); host arrays in reference layout; cells outside the solver range of out are 0 */
int pps_apply_operator(pps_handle* h, int rank, const double* in_host, double* out_host);
/* X = M(B) for every local block (T_Preconditioner::operator(), chebyshevIteration.hpp:48-140) */
int pps_apply_preconditioner(pps_handle* h, int rank, const double* b_host, double* x_host);
/* device-resident timing of `reps` operator applies on block 0's work vectors; returns average ms.
 * with_dot: 0 y = A x; 1 fused with sum(w.y); bit 1 (value 2 or 3): every apply is preceded by the face halo exchange
 * of x with the neighbouring ranks (world_size > 1: the weak-scaling leg of the sweep) */
int pps_bench_operator(pps_handle* h, int reps, int with_dot, double* avg_ms);
/* average device time (ms) per launch of kernel class `which` during the last pps_solve, measured with
 * CUDA events on the launching stream when profiling is enabled: pps_set_profiling(h, 1) brackets every kernel
 * class, pps_set_profiling(h, 2 + c) only class c, 0 switches it off */
int pps_set_profiling(pps_handle* h, int enabled);
int pps_get_kernel_stats(const pps_handle* h, int which, double* avg_ms, long long* launches, const char** name);
long long pps_get_launch_count(const pps_handle* h);          /* kernels launched by the last pps_solve */
int pps_synchronize(pps_handle* h);
/* lower the iteration cap of later solves (<= the max_iter the handle was created with) */
int pps_set_max_iterations(pps_handle* h, int max_iter);
/* out[r*n .. r*n+n) = in of rank r, on every rank: the error gather of checkSolutionLocalGlobal
 * (iterativeSolverBase.hpp:320-373, point-to-point to rank 0 in the reference).  world_size == 1: a copy. */
int pps_allgather(pps_handle* h, const double* in_host, int n, double* out_host);
/* diagnostics of the 17-pass schedule: `reps` launches of ONE fused operator kernel (which: 0 = s-update + operator,
 * 1 = p-update + operator) on frozen pseudo-random inputs, outputs compared bit for bit with the split kernels.
 * variant = ring stages (3, 4, 6; 16 = 6 stages + cross-proxy fence).  out[0] differing operand entries (summed over
 * reps), out[1] differing A*operand entries, out[2] launches whose sums differ, out[3] launches with any difference,
 * out[4..7] pitch, plane, z-chunk, CTAs; out[8 + 5 e ..] = (rep, n_operand, first index, n_result, first index) of the
 * first four bad launches.  Needs a handle created with PPS_FUSE_FULL. */
int pps_debug_fused(pps_handle* h, int which, int variant, int reps, long long* out, int nout);
/* diagnostics: n doubles of a device array of block 0 from element `offset` of the pitched layout (0 s, 1 t, 2 p2, 3 v2, 4 r,
 * 5 v, 6 p, 7 r0, 8 x).  With a negative `reps` pps_debug_fused stops at the first bad launch and leaves the arrays in place. */
int pps_debug_peek(pps_handle* h, int array, long long offset, int n, double* out);
int pps_device_count(void);                                   /* CUDA devices visible to this process */

#ifdef __cplusplus
}
#endif
#endif
