// Reference behaviour: Stability.cpp:22-111.  flow: count Re(lambda) > 0; map and equationFree:
// count |lambda| > 1, where for equationFree the identity is added first because the problem
// returns F(u) = Phi(u) - u rather than the map Phi itself (Stability.cpp:68-71).
#include "Stability.hpp"
#include <cmath>

Stability::Stability(ProblemType type, AbstractNonlinearProblem* pProblem)
    : mpProblem(pProblem), mpProblemJacobian(NULL), mProblemType(type), mFiniteDifferenceEpsilon(1e-8) {}

Stability::Stability(ProblemType type, AbstractNonlinearProblem* pProblem,
                     AbstractNonlinearProblemJacobian* pProblemJacobian)
    : mpProblem(pProblem), mpProblemJacobian(pProblemJacobian), mProblemType(type),
      mFiniteDifferenceEpsilon(1e-8) {}

Stability::~Stability() {}

int Stability::CountUnstable(const arma::cx_vec& eigenvalues) const {
  if (mProblemType == ProblemType::flow) return (int)arma::accu(arma::real(eigenvalues) > 0.0);
  return (int)arma::accu(arma::abs(eigenvalues) > 1.0);
}

int Stability::ComputeNumUnstableEigenvalues(const arma::vec& u) { return CountUnstable(ComputeEigenvalues(u)); }

int Stability::ComputeNumUnstableEigenvalues(const arma::mat& jacobian) { return CountUnstable(arma::eig_gen(jacobian)); }

arma::cx_vec Stability::ComputeEigenvalues(const arma::vec& u) {
  const int n = (int)u.n_rows;
  arma::mat jacobian(n, n);
  if (mpProblemJacobian) mpProblemJacobian->ComputeDFDU(u, jacobian);
  else ComputeDFDU(u, jacobian);
  if (mProblemType == ProblemType::equationFree) jacobian += arma::mat(n, n, arma::fill::eye);
  return arma::eig_gen(jacobian);
}

// base evaluation + one ComputeF per column (Stability.cpp:76-111)
void Stability::ComputeDFDU(const arma::vec& u, arma::mat& jacobian) {
  const int n = (int)u.n_rows;
  const double epsilon = mFiniteDifferenceEpsilon;
  arma::vec base(n), shifted(n), perturbed(u);
  mpProblem->ComputeF(u, base);
  for (int i = 0; i < n; i++) {
    if (i > 0) perturbed(i - 1) = u(i - 1);
    perturbed(i) += epsilon;
    mpProblem->ComputeF(perturbed, shifted);
    jacobian.col(i) = (shifted - base) * std::pow(epsilon, -1);
  }
}
