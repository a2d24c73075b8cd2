#include "AbstractNonlinearSolver.hpp"
#include <iomanip>
#include <iostream>

namespace {
const char* const kRule = "------------------------------------------------";
}

void AbstractNonlinearSolver::PrintHeader(const std::string solverName, int maxIterations,
                                          double tolerance) const {
  std::cout << kRule << "\n Attempt to solve nonlinear problem with " << solverName
            << "\n max number of iterations = " << maxIterations << "\n tolerance = " << tolerance
            << "\n" << kRule << std::endl;
}

void AbstractNonlinearSolver::PrintFooter(const int iteration, const ExitFlagType exitFlag) const {
  std::cout << kRule << "\n";
  if (exitFlag == ExitFlagType::converged)
    std::cout << "The method converged after " << iteration << " iterations" << std::endl;
  else if (exitFlag == ExitFlagType::notConverged)
    std::cout << "The method failed to converge after " << iteration << " iterations" << std::endl;
  else
    std::cout << "Exit flag not known" << std::endl;
}

void AbstractNonlinearSolver::PrintIteration(const int iteration, const double errorEstimate,
                                             const bool initialise) const {
  if (initialise)
    std::cout << std::setw(10) << "Iteration" << std::setw(25) << "error estimate" << std::endl;
  std::cout << std::setw(10) << iteration << std::scientific << std::setprecision(6) << std::setw(25)
            << errorEstimate << std::endl;
}
