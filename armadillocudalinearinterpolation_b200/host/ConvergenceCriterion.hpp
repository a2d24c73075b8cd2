// Residual-norm stopping test (reference: ConvergenceCriterion.hpp / .cpp:11-15).
#ifndef CONVERGENCECRITERIONHEADERDEF
#define CONVERGENCECRITERIONHEADERDEF

class ConvergenceCriterion {
 public:
  explicit ConvergenceCriterion(const double tolerance) : mTolerance(tolerance) {}
  // converged when ||r|| <= tolerance (NaN never converges)
  bool TestConvergence(const double residualNorm) const { return residualNorm <= mTolerance; }
  void SetTolerance(const double tolerance) { mTolerance = tolerance; }

 private:
  ConvergenceCriterion();
  double mTolerance;
};
#endif
