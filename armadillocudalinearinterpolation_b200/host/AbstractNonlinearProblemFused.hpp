// Optional extension of the reference's problem interfaces (AbstractNonlinearProblem.hpp:6-14,
// AbstractNonlinearProblemJacobian.hpp:6-13) — NOT in the reference.  The reference's Newton loop evaluates the
// residual F(u) (NewtonSolver.cpp:110) and then asks for the Jacobian (:93 / :191), whose forward differences
// evaluate F(u) again.  A problem that implements this interface lets NewtonSolver avoid the repeat, either way
// with the same F, the same Jacobian and therefore the same iterates:
//   ComputeDFDUGivenF — the Jacobian from the residual the solver already holds (n evaluations instead of n + 1);
//   ComputeFAndDFDU   — F(u) and the Jacobian in ONE batch, asked for right after every update (the Jacobian of
//                       the converged iterate is then computed in vain); pays when a lone F(u) costs about as much
//                       as the whole batch, i.e. when the batch is spread over several GPUs.
#ifndef ABSTRACTNONLINEARPROBLEMFUSEDHEADERDEF
#define ABSTRACTNONLINEARPROBLEMFUSEDHEADERDEF
#include <armadillo>

class AbstractNonlinearProblemFused {
 public:
  // f = F(u) (what ComputeF returns), dfdu = what ComputeDFDU(u, .) returns; dfdu is pre-sized n x n by the caller
  virtual void ComputeFAndDFDU(const arma::vec& u, arma::vec& f, arma::mat& dfdu) = 0;
  // dfdu = what ComputeDFDU(u, .) returns, given f = F(u)
  virtual void ComputeDFDUGivenF(const arma::vec& u, const arma::vec& f, arma::mat& dfdu) = 0;
  // which of the two the solver should use
  virtual bool PrefersOneBatchPerIterate() const = 0;
  virtual ~AbstractNonlinearProblemFused() {}
};
#endif
