// Damped Newton iteration on an AbstractNonlinearProblem; public surface of the reference
// (NewtonSolver.hpp:17-70): ParameterList defaults, both constructors, Solve, the setters.
#ifndef NEWTONSOLVERHEADERDEF
#define NEWTONSOLVERHEADERDEF
#include <armadillo>
#include "AbstractNonlinearProblem.hpp"
#include "AbstractNonlinearProblemJacobian.hpp"
#include "AbstractNonlinearProblemFused.hpp"
#include "AbstractNonlinearSolver.hpp"
#include "ConvergenceCriterion.hpp"

class NewtonSolver : public AbstractNonlinearSolver {
 public:
  struct ParameterList {
    ParameterList()
        : tolerance(1e-5), maxIterations(10), printOutput(true), finiteDifferenceEpsilon(1e-8), damping(1.0) {}
    double tolerance;
    int maxIterations;
    bool printOutput;
    double finiteDifferenceEpsilon;
    double damping;
  };

  // Jacobian by the solver's own forward differences: n sequential ComputeF calls per iteration
  NewtonSolver(AbstractNonlinearProblem* pProblem, const arma::vec* pInitialGuess,
               const ParameterList* pParameterList);
  // Jacobian supplied by the problem (EventDrivenMapB200: one batched launch, multi-GPU)
  NewtonSolver(AbstractNonlinearProblem* pProblem, AbstractNonlinearProblemJacobian* pProblemJacobian,
               const arma::vec* pInitialGuess, const ParameterList* pParameterList);
  ~NewtonSolver();

  void Solve(arma::vec& solution, arma::vec& residualHistory, ExitFlagType& exitFlag,
             arma::mat* pJacobianExternal = NULL);

  void SetInitialGuess(const arma::vec* pInitialGuess);
  void SetParameterList(const ParameterList* pParameterList);
  void SetProblem(AbstractNonlinearProblem* pProblem);
  void SetProblemJacobian(AbstractNonlinearProblemJacobian* pProblemJacobian);
  void PostProcess();
  // addition: when problem and Jacobian are ONE object that also implements AbstractNonlinearProblemFused, F(u) is not
  // evaluated twice per iteration (see that header; default on; off = the reference's call sequence)
  void SetFusedEvaluation(bool on) { mFuse = on; }

 private:
  NewtonSolver();
  void ComputeDFDU(const arma::vec& u, const arma::vec& f, arma::mat& jacobian);
  void Initialise();

  // none of these is owned, except the criterion (NewtonSolver.cpp:12-16, :36 of the reference)
  AbstractNonlinearProblem* mpProblem;
  AbstractNonlinearProblemJacobian* mpProblemJacobian;
  const arma::vec* mpInitialGuess;
  const ParameterList* mpParameterList;
  ConvergenceCriterion* mpConvergenceCriterion;
  int mMaxIterations;
  bool mPrintOutput;
  double mTolerance;
  bool mFuse;
};
#endif
