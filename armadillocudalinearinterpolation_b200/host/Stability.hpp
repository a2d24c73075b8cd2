// Linear stability of a fixed point from the spectrum of the (finite-difference) Jacobian;
// public surface of the reference (Stability.hpp:13-32).  Added: the finite-difference
// epsilon is initialised (1e-8, NewtonSolver's default) and settable — the reference never
// initialises mFiniteDifferenceEpsilon (Stability.hpp:50, Stability.cpp:6-20,90) — and the
// destructor the reference declares but never defines (Stability.hpp:28) exists.
#ifndef STABILITYHEADERDEF
#define STABILITYHEADERDEF
#include <armadillo>
#include <cassert>
#include "AbstractNonlinearProblem.hpp"
#include "AbstractNonlinearProblemJacobian.hpp"

class Stability {
 public:
  enum class ProblemType { flow, map, equationFree };

  Stability(ProblemType type, AbstractNonlinearProblem* pProblem);
  Stability(ProblemType type, AbstractNonlinearProblem* pProblem,
            AbstractNonlinearProblemJacobian* pProblemJacobian);
  ~Stability();

  int ComputeNumUnstableEigenvalues(const arma::vec& u);
  int ComputeNumUnstableEigenvalues(const arma::mat& jacobian);

  // additions
  void SetFiniteDifferenceEpsilon(double epsilon) { mFiniteDifferenceEpsilon = epsilon; }
  arma::cx_vec ComputeEigenvalues(const arma::vec& u);

 private:
  Stability();
  int CountUnstable(const arma::cx_vec& eigenvalues) const;
  void ComputeDFDU(const arma::vec& u, arma::mat& jacobian);

  AbstractNonlinearProblem* mpProblem;
  AbstractNonlinearProblemJacobian* mpProblemJacobian;
  ProblemType mProblemType;
  double mFiniteDifferenceEpsilon;
};
#endif
