// Base class of the nonlinear solvers: exit flag, the Solve() contract and the console
// reporting helpers (reference: AbstractNonlinearSolver.hpp:16-40, .cpp:11-95).
#ifndef ABSTRACTNONLINEARSOLVERHEADERDEF
#define ABSTRACTNONLINEARSOLVERHEADERDEF
#include <armadillo>
#include <string>

class AbstractNonlinearSolver {
 public:
  enum class ExitFlagType { converged, notConverged };

  virtual ~AbstractNonlinearSolver() {}

  // solution must arrive sized like the initial guess; residualHistory is sized by the solver;
  // pJacobianExternal (optional, n x n) receives the last Jacobian the solver used
  virtual void Solve(arma::vec& solution, arma::vec& residualHistory, ExitFlagType& exitFlag,
                     arma::mat* pJacobianExternal = NULL) = 0;

 protected:
  virtual void PrintHeader(const std::string solverName, int maxIterations, double tolerance) const;
  virtual void PrintFooter(const int iteration, const ExitFlagType exitFlag) const;
  virtual void PrintIteration(const int iteration, const double errorEstimate,
                              const bool initialise = false) const;
};
#endif
