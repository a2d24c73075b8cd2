// Problem plug-in seam, kept signature-for-signature from the reference
// (AbstractNonlinearProblem.hpp:6-14) so existing solvers and user problems bind unchanged:
// a residual F(u) and an optional hook called once a solve has finished.
#ifndef ABSTRACTCNONLINEARPROBLEMHEADERDEF
#define ABSTRACTCNONLINEARPROBLEMHEADERDEF
#include <armadillo>

class AbstractNonlinearProblem {
 public:
  virtual ~AbstractNonlinearProblem() {}
  // f <- F(u); f is (re)sized by the implementation
  virtual void ComputeF(const arma::vec& u, arma::vec& f) = 0;
  // called by the solver after its last iteration (the event-driven map draws a new seed)
  virtual void PostProcess() {}
};
#endif
