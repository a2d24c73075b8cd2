// EventDrivenMapB200 — the drop-in for the reference's EventDrivenMap
// (EventDrivenMap.hpp:11-121): same constructor, same public methods, same
// AbstractNonlinearProblem seam.  It additionally implements AbstractNonlinearProblemJacobian,
// so it can be handed to NewtonSolver's 4-argument constructor / Stability's 3-argument
// constructor and the whole finite-difference Jacobian becomes ONE batched GPU launch.
//
// Pure host C++ (g++): no CUDA or device type appears here — the class forwards to the C-ABI
// of include/b200_edm.h (libb200edm.so).  No Armadillo type crosses that boundary.
#ifndef EVENTDRIVENMAPB200HEADERDEF
#define EVENTDRIVENMAPB200HEADERDEF
#include <armadillo>
#include <stdexcept>
#include <string>
#include "AbstractNonlinearProblem.hpp"
#include "AbstractNonlinearProblemJacobian.hpp"
#include "AbstractNonlinearProblemFused.hpp"
#include "b200_edm.h"

class EventDrivenMapB200 : public AbstractNonlinearProblem, public AbstractNonlinearProblemJacobian,
                           public AbstractNonlinearProblemFused {
 public:
  // reference signature (EventDrivenMap.cu:57): parameters (p[0] = beta), realisations.
  // Defaults of the reference: 1024 neurons (mNoThreads, :70), noSpikes = 3 fronts
  // (parameters.hpp:12), FP64 arithmetic here (the reference's device math is FP32).
  EventDrivenMapB200(const arma::vec* pParameters, unsigned int noReal);
  EventDrivenMapB200(const arma::vec* pParameters, unsigned int noReal, unsigned int noNeurons,
                     unsigned int noFronts, b200_dtype precision = B200_F64);
  ~EventDrivenMapB200();

  // AbstractNonlinearProblem
  void ComputeF(const arma::vec& u, arma::vec& f);
  void PostProcess();
  // AbstractNonlinearProblemJacobian: forward differences with the epsilon set below
  void ComputeDFDU(const arma::vec& u, arma::mat& dfdu);
  // AbstractNonlinearProblemFused: the same Jacobian and the base evaluation F(u) it contains, one batch
  void ComputeFAndDFDU(const arma::vec& u, arma::vec& f, arma::mat& dfdu);
  void ComputeDFDUGivenF(const arma::vec& u, const arma::vec& f, arma::mat& dfdu);
  // one GPU: F alone is a quarter of the n = 3 batch -> reuse the residual; several GPUs: F alone and the batch
  // are both one ring's serial chain -> one batch per iterate (measured: tools/newton_fused_ab.py)
  bool PrefersOneBatchPerIterate() const { return mNoDevices > 1; }

  // reference setters (EventDrivenMap.hpp:27-51)
  void SetTimeHorizon(const float T);
  void SetNoRealisations(const int noReal);
  void SetNoThreads(const int noThreads);  // = neurons per ring (one thread per neuron there)
  void SetParameterStdDev(const float sigma);
  void SetParameters(const unsigned int parId, const float parVal);
  void ResetSeed();
  void SetNewSeed();
  void SetDebugFlag(const bool val);

  // additions
  void SetFiniteDifferenceEpsilon(double epsilon) { mEpsilon = epsilon; }  // Driver.cu:37 uses 1e-2
  void SetSeed(unsigned long long seed);
  // profile map (BASELINE config 5, see b200_edm_set_profile_mode): vectors become (V_c, S_c), n = 2 nCoarse
  void SetProfileMode(unsigned int nCoarse);
  // split every evaluation over several GPUs of this process (first id = the device the map was created on)
  void SetDevices(const int* deviceIds, unsigned int nDevices);
  void SetPrintOutput(bool on) { mPrint = on; }
  // n x ncols evaluation points -> n x ncols residuals, one launch
  void ComputeFBatch(const arma::mat& uCols, arma::mat& fCols);
  b200_edm* Handle() { return mpHandle; }

  struct firing { float time; unsigned int index; };  // EventDrivenMap.hpp:54-57 (kept for source compatibility)

 private:
  EventDrivenMapB200();
  EventDrivenMapB200(const EventDrivenMapB200&);
  void Check(int status, const char* what) const;
  b200_edm* mpHandle;
  double mEpsilon;
  bool mPrint;
  unsigned int mNoDevices;
};
#endif
