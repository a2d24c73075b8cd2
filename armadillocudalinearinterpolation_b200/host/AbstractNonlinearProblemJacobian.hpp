// User-supplied Jacobian seam (reference: AbstractNonlinearProblemJacobian.hpp:6-13).
// NewtonSolver (4-argument constructor) and Stability (3-argument constructor) call this
// instead of their own sequential finite-difference loops; EventDrivenMapB200 implements it
// with ONE batched GPU launch for all n+1 evaluations.
#ifndef ABSTRACTCNONLINEARPROBLEMJACOBIANHEADERDEF
#define ABSTRACTCNONLINEARPROBLEMJACOBIANHEADERDEF
#include <armadillo>

class AbstractNonlinearProblemJacobian {
 public:
  virtual ~AbstractNonlinearProblemJacobian() {}
  // dfdu(:, i) <- dF/du_i at u; dfdu arrives pre-sized n x n (NewtonSolver.cpp:85, Stability.cpp:57)
  virtual void ComputeDFDU(const arma::vec& u, arma::mat& dfdu) = 0;
};
#endif
