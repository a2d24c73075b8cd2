// examples/driver.cpp — the reference's experiment (Driver.cu:11-126) on the B200 map, including
// the beta-continuation loop that the reference leaves commented out (Driver.cu:86-112):
// solve for the travelling wave, count unstable eigenvalues, step beta, reuse the solution.
//
//   driver_b200 [steps=3] [noReal=1000] [noNeurons=1024] [dbeta=0.1] [nGpus=1]
#include <armadillo>
#include <chrono>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include "EventDrivenMap.hpp"
#include "NewtonSolver.hpp"
#include "Stability.hpp"
#include "parameters.hpp"

int main(int argc, char* argv[]) {
  const int steps = argc > 1 ? std::atoi(argv[1]) : 3;
  const unsigned noReal = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1000;
  const unsigned noNeurons = argc > 3 ? (unsigned)std::atoi(argv[3]) : 1024;
  const double dbeta = argc > 4 ? std::atof(argv[4]) : 0.1;
  const int nGpus = argc > 5 ? std::atoi(argv[5]) : 1;

  arma::vec parameters(1);
  parameters << 13.0589f;                                  // Driver.cu:16
  EventDrivenMap map(&parameters, noReal);                 // Driver.cu:20
  if (noNeurons != 1024) map.SetNoThreads((int)noNeurons);
  map.SetFiniteDifferenceEpsilon(1e-2);                    // Driver.cu:37
  if (nGpus > 1) {                                         // every evaluation split over the GPUs of this process
    int ids[16];
    for (int i = 0; i < nGpus && i < 16; ++i) ids[i] = i;
    map.SetDevices(ids, (unsigned)(nGpus < 16 ? nGpus : 16));
  }

  arma::vec guess(noSpikes);
  guess << 0.3310f << 0.6914f << 1.3557f;                  // Driver.cu:24

  NewtonSolver::ParameterList pars;                        // Driver.cu:27-31
  pars.tolerance = 1e-4;
  pars.maxIterations = 10;
  pars.printOutput = true;
  pars.damping = 1.0;
  pars.finiteDifferenceEpsilon = 1e-2;

  // the map supplies its own Jacobian: n+1 evaluations in one batched launch
  NewtonSolver newton(&map, &map, &guess, &pars);
  Stability stability(Stability::ProblemType::equationFree, &map, &map);

  arma::vec solution(noSpikes), history;
  AbstractNonlinearSolver::ExitFlagType flag;
  for (int i = 0; i < steps; ++i) {
    auto t0 = std::chrono::steady_clock::now();
    newton.SetInitialGuess(&guess);
    newton.Solve(solution, history, flag);
    const int unstable = stability.ComputeNumUnstableEigenvalues(solution);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << std::setprecision(15) << "beta = " << parameters(0) << "  solution =";
    for (arma::uword j = 0; j < solution.n_elem; ++j) std::cout << " " << solution(j);
    std::cout << "  unstable eigenvalues = " << unstable << (unstable > 0 ? "  (unstable)" : "  (stable)")
              << "  [" << ms << " ms]" << std::endl;
    parameters += dbeta;                                   // Driver.cu:107-109
    map.SetParameters(0, (float)parameters(0));
    if (flag == AbstractNonlinearSolver::ExitFlagType::converged) guess = solution;
  }
  return 0;
}
