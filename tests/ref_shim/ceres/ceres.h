// stand-in for the Ceres declarations the reference's headers need in order to compile. The two call sites in the translation
// units built into oracle/_ref (Sprase_ImgAlign::CeresSolver, Feature_Alignment::Align2DCeres) are dead code in the reference
// (never called: ref src/Sprase_ImageAlign.cpp:50, src/Feature_alignment.cpp:153); ceres::Solve aborts if it is ever reached.
#ifndef MINI_CERES_H
#define MINI_CERES_H
#include <cstdlib>
#include <string>
#include <vector>
namespace ceres {
class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};
template <int kNumResiduals, int... Ns> class SizedCostFunction : public CostFunction {};
class LocalParameterization {
public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};
class LossFunction {
public:
    virtual ~LossFunction() {}
};
class Problem {
public:
    void AddParameterBlock(double*, int) {}
    void AddParameterBlock(double*, int, LocalParameterization*) {}
    template <class... P> void* AddResidualBlock(CostFunction*, LossFunction*, P...) { return nullptr; }
};
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
struct Solver {
    struct Options {
        TrustRegionStrategyType trust_region_strategy_type;
        LinearSolverType linear_solver_type;
        bool minimizer_progress_to_stdout;
        int max_num_iterations;
        int num_threads;
        Options() : trust_region_strategy_type(LEVENBERG_MARQUARDT), linear_solver_type(DENSE_QR), minimizer_progress_to_stdout(false), max_num_iterations(50), num_threads(1) {}
    };
    struct Summary {
        std::string FullReport() const { return std::string(); }
        std::string BriefReport() const { return std::string(); }
    };
};
inline void Solve(const Solver::Options&, Problem*, Solver::Summary*) { std::abort(); }
}  // namespace ceres
#endif
