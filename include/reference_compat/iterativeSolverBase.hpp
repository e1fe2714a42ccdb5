// iterativeSolverBase.hpp (reference_compat) -- the solver-class concept of the reference
// (solverPoissonMPI_CPU/include/iterativeSolverBase.hpp: ctor, setProblem :51-55, getters :410-425,
// checkSolutionLocalGlobal :283-408) re-expressed on top of the C ABI of libpps_b200.so.
//
// The class templates in BiCGSTAB.hpp / baseCG.hpp / chebyshevIteration.hpp / noneSolver.hpp keep the
// reference's template parameter lists, so `using T_Solver = BiCGSTAB<DIM, T_data, tollMainSolver, ...>` in an
// unmodified inputParam.hpp selects the same algorithm -- executed on the GPU.  Host work that stays here is
// exactly what the reference also does on the host outside its timed region: evaluating ExactSolutionAndBCs
// for setProblem(), for the Neumann face derivatives and for the post-solve error report.
#pragma once

#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../pps_b200.h"
#include "blockGrid.hpp"
#include "communicationMPI.hpp"
#include "matrixFreeOperatorA.hpp"
#include "mpi.h"
#include "output.hpp"
#include "solverSetup.hpp"

namespace pps_compat {

inline void die(const char* what) {
    std::cerr << "Error: " << what << ": " << pps_last_error() << std::endl;
    std::exit(-1);   // the reference's error style, main.cpp:51-55
}

// compile-time description of a level of the solver stack (what the T_Preconditioner slot carries)
struct StackInfo {
    int solver_kind;     // PPS_SOLVER_* or -1 (not a Krylov main solver)
    int precond_kind;    // PPS_PRECOND_* when used in the preconditioner slot, -1 = not implemented there
    int iterations;      // maxIteration template argument
    bool communication;  // communicationON template argument
    int tolerance = 0;         // tolerance template argument (times tollScalingFactor), used by nested Krylov preconditioners
    int inner_iterations = 0;  // maxIteration of the Chebyshev iteration inside a nested CG (T_Preconditioner3, inputParam.hpp:29)
};

// TimeCounter of the alpaka tree (solverPoissonMPI_alpaka/include/solverSetup.hpp:207-286), printed by its main.cpp:156 as
// `solver.timeCounter.printAverageTime(solver.getNumIterationFinal())`.  Here the phases are the kernel classes of libpps_b200.so,
// timed with CUDA events on the launching stream when the solve ran with phase timers on (environment PPS_PHASE_TIMERS=1 or
// -DPPS_PHASE_TIMERS; two event records per launch, a few per cent of overhead).  The aggregate rows keep the alpaka names.
class TimeCounterAdapter {
  public:
    void attach(pps_handle* h) { h_ = h; }
    void printAverageTime(const int numCycles) const {
        if (!h_ || numCycles <= 0) return;
        double precond = 0, comm = 0, allred = 0, kernels = 0, ghosts = 0, total = 0;
        std::cout << "Time in milliseconds " << std::endl;
        for (int k = 0;; k++) {
            double avg = 0;
            long long n = 0;
            const char* name = nullptr;
            if (pps_get_kernel_stats(h_, k, &avg, &n, &name)) break;
            if (n == 0) continue;
            const double per_cycle = avg * static_cast<double>(n) / numCycles;
            const std::string nm(name);
            std::cout << "timeTot[" << nm << "] " << per_cycle << "  (" << static_cast<double>(n) / numCycles << " launches per cycle)" << std::endl;
            if (nm.rfind("cheb", 0) == 0) precond += per_cycle;
            else if (nm == "halo") comm += per_cycle;
            else if (nm == "scalar_op") allred += per_cycle;
            else if (nm == "neumann_ghost") ghosts += per_cycle;
            else if (nm != "setup" && nm.rfind("residual", 0) != 0) kernels += per_cycle;
            total += per_cycle;
        }
        std::cout << "timeTotPreconditionerTot " << precond << std::endl;
        std::cout << "timeTotCommunicationTot " << comm << std::endl;
        std::cout << "timeTotAllReductionTot " << allred << std::endl;
        std::cout << "timeTotKernelsBBiCGstabTot " << kernels << std::endl;
        std::cout << "timeTotResetNeumanBCs " << ghosts << std::endl;
        std::cout << "timeTotal " << total << std::endl;
    }

  private:
    pps_handle* h_ = nullptr;
};

template <int DIM, typename T_data, int maxIteration>
class SolverAdapter {
  public:
    TimeCounterAdapter timeCounter;   // alpaka tree: iterativeSolverBaseAlpaka.hpp:640

    SolverAdapter(const BlockGrid<DIM, T_data>& blockGrid, const ExactSolutionAndBCs<DIM, T_data>& exact,
                  CommunicatorMPI<DIM, T_data>&, const StackInfo& self, const StackInfo& precond, int tolerance, const char* name)
        : grid_(blockGrid), exact_(exact), name_(name), rank_(blockGrid.getMyrank()) {
        static_assert(sizeof(T_data) == 8, "the B200 path computes in fp64 (T_data = double)");
        const auto nr = blockGrid.getNranks();
        nranksTot_ = nr[0] * nr[1] * nr[2];
        if (self.solver_kind < 0) { std::cerr << "Error: " << name << " is not available as the main solver" << std::endl; std::exit(-1); }
        if (precond.precond_kind < 0) {
            std::cerr << "Error: this preconditioner stack is not implemented on the B200 path (available: NoneSolver, ChebyshevIteration, "
                         "BiCGSTAB<.., false, communicationOFF | communicationON, NoneSolver>, BaseCG<.., false, communicationOFF, ChebyshevIteration>)" << std::endl;
            std::exit(-1);
        }
        if (self.solver_kind == PPS_SOLVER_CHEBYSHEV && precond.precond_kind != PPS_PRECOND_NONE) {
            std::cerr << "Error: ChebyshevIteration as main solver ignores its preconditioner slot: use NoneSolver there" << std::endl;
            std::exit(-1);
        }

        pps_config c;
        pps_default_config(&c);
        c.dim = DIM;
        for (int d = 0; d < 3; d++) {
            c.npglobal[d] = blockGrid.getNpglobal()[d];
            c.nranks[d] = nr[d];
            c.ds[d] = blockGrid.getDs()[d];
            c.origin[d] = blockGrid.getOrigin()[d];
            c.guards[d] = blockGrid.getGuards()[d];
        }
        for (int f = 0; f < 6; f++) c.bcs_type[f] = blockGrid.getBcsType()[f];
        c.solver = self.solver_kind;
        c.precond = precond.precond_kind;
        c.tolerance = static_cast<T_data>(tolerance) * tollScalingFactor;   // BiCGSTAB.hpp:22
        c.max_iter = maxIteration;
        c.cheb_max_iter = precond.precond_kind == PPS_PRECOND_CHEBYSHEV ? precond.iterations : chebyshevMax;
        if (self.solver_kind == PPS_SOLVER_CHEBYSHEV) c.cheb_max_iter = maxIteration;                 // chebyshevIteration.hpp:94,137
        if (precond.precond_kind == PPS_PRECOND_CG_CHEB_LOCAL) c.cheb_max_iter = precond.inner_iterations;
        nested_ = precond.precond_kind == PPS_PRECOND_BICGSTAB_LOCAL || precond.precond_kind == PPS_PRECOND_CG_CHEB_LOCAL;
        if (nested_) {
            c.precond_tolerance = static_cast<T_data>(precond.tolerance) * tollScalingFactor;         // BiCGSTAB.hpp:22 of the nested solver
            c.precond_max_iter = precond.iterations;
        }
        c.cheb_epsilon = epsilon;
        c.cheb_rescale_min = rescaleEigMin;
        c.cheb_rescale_max = rescaleEigMax;
        c.order_neumann = orderNeumanBcs;
        // communicationON in the preconditioner slot: the Chebyshev sweeps exchange faces (chebyshevIteration.hpp:69-73,97-101); the nested
        // BiCGSTAB becomes a global solve (BiCGSTAB.hpp:135-139,156-164,182-186,216-225,247-257 inside the preconditioner)
        c.precond_communication = ((precond.precond_kind == PPS_PRECOND_CHEBYSHEV || precond.precond_kind == PPS_PRECOND_BICGSTAB_LOCAL) &&
                                   precond.communication) ? 1 : 0;
        // alpaka-only switches, for a solverSetup.hpp / inputParam.hpp written for that tree:
        //   -DPPS_CHEBYSHEV_FLOAT       T_data_chebyshev = float   (solverPoissonMPI_alpaka/include/solverSetup.hpp:14)
        //   -DPPS_CHEBYSHEV_LOCAL_EIG   `local` eigenvalue bounds  (solverPoissonMPI_alpaka/include/inputParam.hpp:21-22,27)
#ifdef PPS_CHEBYSHEV_FLOAT
        c.cheb_precision = PPS_CHEB_FP32;
#endif
#ifdef PPS_CHEBYSHEV_LOCAL_EIG
        c.cheb_eigenvalues = PPS_CHEB_EIG_LOCAL;
#endif
        if (const char* a = std::getenv("PPS_ARITHMETIC")) c.arithmetic = std::atoi(a);
        cfg_ = c;
        precondIterations_ = precond.precond_kind == PPS_PRECOND_CHEBYSHEV ? precond.iterations : 0;
        World& w = world();
        const int ngpu = pps_device_count();
        if (nranksTot_ == 1) {
            if (pps_create(&c, 0, 1, nullptr, &h_)) die("pps_create");
        } else if (nranksTot_ <= ngpu) {
            // one rank-thread per GPU, NCCL between them
            if (rank_ == 0 && pps_get_unique_id(w.unique_id)) die("pps_get_unique_id");
            barrier();
            c.device = rank_;
            if (pps_create(&c, rank_, nranksTot_, w.unique_id, &h_)) die("pps_create");
        } else {
            // more ranks than GPUs: one handle hosts every block on GPU 0 ("virtual ranks", same arithmetic)
            shared_ = true;
            if (rank_ == 0) {
                if (pps_create(&c, 0, 1, nullptr, &w.shared_handle)) die("pps_create");
                w.gather.assign(2 * static_cast<size_t>(nranksTot_), 0.0);
            }
            barrier();
            h_ = w.shared_handle;
        }
    }

    ~SolverAdapter() {
        if (shared_) {
            barrier();
            if (rank_ == 0) pps_destroy(h_);
        } else if (h_) {
            pps_destroy(h_);
        }
    }

    // iterativeSolverBase.hpp:51-55: Dirichlet boundary planes of x from u_exact (:557-603), b = f on the data range (:537-555)
    void setProblem(T_data fieldX[], T_data fieldB[]) {
        const auto ld = grid_.getIndexLimitsData();
        const auto ng = grid_.getNlocalGuards();
        const auto hb = grid_.getHasBoundary();
        const auto bt = grid_.getBcsType();
        const long sj = ng[0], sk = static_cast<long>(ng[0]) * ng[1];
        for (int face = 0; face < 2 * DIM; face++) {
            if (!(hb[face] && bt[face] == 0)) continue;
            std::array<int, 6> lim = ld;
            facePlane(face, lim);
            for (int k = lim[4]; k < lim[5]; k++)
                for (int j = lim[2]; j < lim[3]; j++)
                    for (int i = lim[0]; i < lim[1]; i++)
                        fieldX[i + sj * j + sk * k] = exact_.trueSolutionFxyz(coord(0, i), coord(1, j), coord(2, k));
        }
        for (int k = ld[4]; k < ld[5]; k++)
            for (int j = ld[2]; j < ld[3]; j++)
                for (int i = ld[0]; i < ld[1]; i++)
                    fieldB[i + sj * j + sk * k] = exact_.setFieldB(coord(0, i), coord(1, j), coord(2, k));
    }

    // T_Solver::operator()(fieldX, fieldB, operatorA): BiCGSTAB.hpp:55-322 / baseCG.hpp:44-260 on the GPU
    void operator()(T_data fieldX[], T_data fieldB[], MatrixFreeOperatorA<DIM, T_data>&) {
        if (pps_set_fields(h_, rank_, fieldX, fieldB)) die("pps_set_fields");
        const auto hb = grid_.getHasBoundary();
        const auto bt = grid_.getBcsType();
        for (int face = 0; face < 2 * DIM; face++) {
            if (!(hb[face] && bt[face] == 1)) continue;
            std::array<int, 6> lim = grid_.getIndexLimitsData();
            facePlane(face, lim);
            std::vector<T_data> g;
            for (int k = lim[4]; k < lim[5]; k++)
                for (int j = lim[2]; j < lim[3]; j++)
                    for (int i = lim[0]; i < lim[1]; i++)
                        g.push_back(exact_.trueSolutionDdir(coord(0, i), coord(1, j), coord(2, k), face / 2));
            if (pps_set_neumann_face(h_, rank_, face, g.data(), g.size())) die("pps_set_neumann_face");
        }
        if (shared_) barrier();
        timeCounter.attach(h_);
#ifdef PPS_PHASE_TIMERS
        const bool phase_timers = true;
#else
        const bool phase_timers = std::getenv("PPS_PHASE_TIMERS") != nullptr && std::atoi(std::getenv("PPS_PHASE_TIMERS")) != 0;
#endif
        if (phase_timers && (!shared_ || rank_ == 0)) pps_set_profiling(h_, 1);
        if (!shared_ || rank_ == 0)
            if (pps_solve(h_)) die("pps_solve");
        if (shared_) barrier();
        if (pps_get_solution(h_, rank_, fieldX)) die("pps_get_solution");
        if (pps_get_rhs(h_, rank_, fieldB)) die("pps_get_rhs");
        numIterationFinal_ = pps_get_iterations(h_);
        errorFromIteration_ = pps_get_error_iteration(h_);
        errorComputeOperator_ = pps_get_error_operator(h_);
        normFieldB_ = pps_get_norm_b(h_);
        durationSolver_ = std::chrono::duration<double>(pps_get_loop_seconds(h_));
        precondTotal_ = pps_get_preconditioner_iterations(h_);
        if (rank_ == 0 && cfg_.solver != PPS_SOLVER_CHEBYSHEV) report();   // the Chebyshev main loop prints nothing (chebyshevIteration.hpp:48-140)
    }

    // iterativeSolverBase.hpp:283-408: per-rank sum and max of |x - u_exact| on the data range, gathered, two lines on rank 0
    T_data checkSolutionLocalGlobal(T_data fieldX[]) {
        const auto ld = grid_.getIndexLimitsData();
        const auto ng = grid_.getNlocalGuards();
        const long sj = ng[0], sk = static_cast<long>(ng[0]) * ng[1];
        std::vector<T_data> u(static_cast<size_t>(grid_.getNtotLocalGuards()), 0);
        for (int k = ld[4]; k < ld[5]; k++)
            for (int j = ld[2]; j < ld[3]; j++)
                for (int i = ld[0]; i < ld[1]; i++)
                    u[i + sj * j + sk * k] = exact_.trueSolutionFxyz(coord(0, i), coord(1, j), coord(2, k));
        double mine[2] = {0, 0};
        (void)fieldX;   // the device copy is the solution that pps_get_solution returned
        if (pps_check_solution(h_, rank_, u.data(), &mine[0], &mine[1])) die("pps_check_solution");
        std::vector<double> all(2 * static_cast<size_t>(nranksTot_), 0.0);
        if (shared_) {
            World& w = world();
            w.gather[2 * rank_] = mine[0];
            w.gather[2 * rank_ + 1] = mine[1];
            barrier();
            all = w.gather;
            barrier();
        } else if (pps_allgather(h_, mine, 2, all.data())) {
            die("pps_allgather");
        }
        if (rank_ == 0) {
            int ib = 0, ip = 0;
            for (int r = 0; r < nranksTot_; r++) {
                if (all[2 * r] > all[2 * ib]) ib = r;
                if (all[2 * r + 1] > all[2 * ip + 1]) ip = r;
            }
            std::cout << "Max error local block avg " << all[2 * ib] / grid_.getNtotLocalNoGuards() << " in rank " << ib << std::endl;
            std::cout << "Max error local point " << all[2 * ip + 1] << " in rank " << ip << std::endl;
        }
        return mine[0];
    }

    std::chrono::duration<double> getDurationSolver() const { return durationSolver_; }
    T_data getErrorFromIteration() const { return errorFromIteration_; }
    T_data getErrorComputeOperator() const { return errorComputeOperator_; }
    int getNumIterationFinal() const { return numIterationFinal_; }
    // extras of the alpaka tree (iterativeSolverBaseAlpaka.hpp:615-638)
    int getNumIterationPreconditionerFinal() const {
        return nested_ ? static_cast<int>(precondTotal_) : 2 * precondIterations_ * numIterationFinal_;
    }
    void writeResidualHistory() const {
        const std::vector<T_data> hst = getResidualHistory();
        write_residual_history("residualHistory.txt", durationSolver_.count(), numIterationFinal_, getNumIterationPreconditionerFinal(),
                               hst.data(), static_cast<int>(hst.size()), maxIteration);
    }
    std::vector<T_data> getResidualHistory() const {
        std::vector<T_data> hst(static_cast<size_t>(numIterationFinal_) + 1);
        pps_get_history(h_, 0, hst.data(), static_cast<int>(hst.size()));
        return hst;
    }
    pps_handle* handle() const { return h_; }

  private:
    T_data coord(int d, int i) const {   // iterativeSolverBase.hpp:547-549
        return grid_.getOrigin()[d] + (i - grid_.getIndexLimitsData()[2 * d]) * grid_.getDs()[d] +
               grid_.getGlobalLocation()[d] * (grid_.getNlocalNoGuards()[d]) * grid_.getDs()[d];
    }
    void facePlane(int face, std::array<int, 6>& lim) const {
        const int d = face / 2;
        if (face % 2 == 0) lim[2 * d + 1] = lim[2 * d] + grid_.getGuards()[d];
        else lim[2 * d] = lim[2 * d + 1] - grid_.getGuards()[d];
    }
    // the stdout lines of BiCGSTAB.hpp:105-108,285 / baseCG.hpp:88-91,213 (printed after the solve: the
    // iteration runs on the device without host round trips)
    void report() const {
        const auto loc = grid_.getGlobalLocation();
        const auto ld = grid_.getIndexLimitsData();
        const auto ls = grid_.getIndexLimitsSolver();
        std::cout << "Debug in " << name_ << " START " << " main loop " << 1 << " globalLocation " << loc[0] << " " << loc[1] << " " << loc[2]
                  << " indexLimitsData " << ld[0] << " " << ld[1] << " " << ld[2] << " " << ld[3] << " " << ld[4] << " " << ld[5]
                  << " indexLimitsSolver " << ls[0] << " " << ls[1] << " " << ls[2] << " " << ls[3] << " " << ls[4] << " " << ls[5]
                  << " norm fieldB " << normFieldB_ << std::endl;
        const int n = numIterationFinal_;
        if (n < 10) return;
        std::vector<double> err(n + 1), a(n), o(n), r(n);
        pps_get_history(h_, 0, err.data(), n + 1);
        pps_get_history(h_, 1, a.data(), n);
        pps_get_history(h_, 2, o.data(), n);
        pps_get_history(h_, 3, r.data(), n);
        const bool cg = cfg_.solver == PPS_SOLVER_CG;
        for (int it = 10; it <= n; it += 10) {
            if (cg) std::cout << " Debug in base CG iter " << it << " alpha " << a[it - 1] << " beta " << o[it - 1] << " error " << err[it] << std::endl;
            else std::cout << " Debug in BiCGSTAB iter " << it << " alpha " << a[it - 1] << " omega " << o[it - 1] << " rho0 " << r[it - 1]
                           << " error " << err[it] << std::endl;
        }
    }

    const BlockGrid<DIM, T_data>& grid_;
    const ExactSolutionAndBCs<DIM, T_data>& exact_;
    const char* name_;
    int rank_ = 0, nranksTot_ = 1, precondIterations_ = 0;
    bool shared_ = false, nested_ = false;
    long long precondTotal_ = 0;
    pps_config cfg_{};
    pps_handle* h_ = nullptr;
    T_data normFieldB_ = 1, errorFromIteration_ = -1, errorComputeOperator_ = -1;
    int numIterationFinal_ = 0;
    std::chrono::duration<double> durationSolver_{0};
};

}  // namespace pps_compat
