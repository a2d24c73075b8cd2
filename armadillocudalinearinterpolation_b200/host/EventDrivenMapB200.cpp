#include "EventDrivenMapB200.hpp"
#include <iostream>

void EventDrivenMapB200::Check(int status, const char* what) const {
  if (status != B200_OK)
    throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}

EventDrivenMapB200::EventDrivenMapB200(const arma::vec* pParameters, unsigned int noReal)
    : mpHandle(NULL), mEpsilon(1e-2), mPrint(true), mNoDevices(1) {
  Check(b200_edm_create(pParameters->memptr(), pParameters->n_elem, noReal, 1024, 3, B200_F64, &mpHandle),
        "EventDrivenMapB200");
}

EventDrivenMapB200::EventDrivenMapB200(const arma::vec* pParameters, unsigned int noReal,
                                       unsigned int noNeurons, unsigned int noFronts, b200_dtype precision)
    : mpHandle(NULL), mEpsilon(1e-2), mPrint(true), mNoDevices(1) {
  Check(b200_edm_create(pParameters->memptr(), pParameters->n_elem, noReal, noNeurons, noFronts, precision, &mpHandle),
        "EventDrivenMapB200");
}

EventDrivenMapB200::~EventDrivenMapB200() { b200_edm_destroy(mpHandle); }

void EventDrivenMapB200::ComputeF(const arma::vec& u, arma::vec& f) {
  f.set_size(u.n_elem);
  Check(b200_edm_compute_f(mpHandle, u.memptr(), u.n_elem, f.memptr()), "ComputeF");
}

void EventDrivenMapB200::ComputeFBatch(const arma::mat& uCols, arma::mat& fCols) {
  fCols.set_size(uCols.n_rows, uCols.n_cols);
  Check(b200_edm_compute_f_batch(mpHandle, uCols.memptr(), uCols.n_rows, uCols.n_cols, fCols.memptr()), "ComputeFBatch");
}

void EventDrivenMapB200::ComputeDFDU(const arma::vec& u, arma::mat& dfdu) {
  if (dfdu.n_rows != u.n_elem || dfdu.n_cols != u.n_elem) dfdu.set_size(u.n_elem, u.n_elem);
  Check(b200_edm_compute_dfdu(mpHandle, u.memptr(), u.n_elem, mEpsilon, dfdu.memptr(), NULL), "ComputeDFDU");
}

void EventDrivenMapB200::ComputeFAndDFDU(const arma::vec& u, arma::vec& f, arma::mat& dfdu) {
  if (dfdu.n_rows != u.n_elem || dfdu.n_cols != u.n_elem) dfdu.set_size(u.n_elem, u.n_elem);
  f.set_size(u.n_elem);
  Check(b200_edm_compute_dfdu(mpHandle, u.memptr(), u.n_elem, mEpsilon, dfdu.memptr(), f.memptr()), "ComputeFAndDFDU");
}

void EventDrivenMapB200::ComputeDFDUGivenF(const arma::vec& u, const arma::vec& f, arma::mat& dfdu) {
  if (dfdu.n_rows != u.n_elem || dfdu.n_cols != u.n_elem) dfdu.set_size(u.n_elem, u.n_elem);
  if (f.n_elem != u.n_elem) throw std::invalid_argument("ComputeDFDUGivenF: f and u differ in length");
  Check(b200_edm_compute_dfdu_given_f(mpHandle, u.memptr(), u.n_elem, mEpsilon, f.memptr(), dfdu.memptr()), "ComputeDFDUGivenF");
}

void EventDrivenMapB200::PostProcess() { SetNewSeed(); }

void EventDrivenMapB200::SetTimeHorizon(const float T) {
  Check(b200_edm_set_time_horizon(mpHandle, (double)T), "SetTimeHorizon");
  if (mPrint) std::cout << "Time horizon set to " << T << std::endl;
}
void EventDrivenMapB200::SetNoRealisations(const int noReal) {
  Check(noReal > 0 ? b200_edm_set_no_realisations(mpHandle, (unsigned)noReal) : b200_edm_set_no_realisations(mpHandle, 0),
        "SetNoRealisations");
  if (mPrint) std::cout << "Number of realisations set to " << noReal << std::endl;
}
void EventDrivenMapB200::SetNoThreads(const int noThreads) {
  Check(noThreads > 0 ? b200_edm_set_no_neurons(mpHandle, (unsigned)noThreads) : b200_edm_set_no_neurons(mpHandle, 0),
        "SetNoThreads");
  if (mPrint) std::cout << "Number of threads set to " << noThreads << std::endl;
}
void EventDrivenMapB200::SetParameterStdDev(const float sigma) {
  Check(b200_edm_set_param_stddev(mpHandle, (double)sigma), "SetParameterStdDev");
  if (mPrint) std::cout << "Parameter standard deviation set to " << sigma << std::endl;
}
void EventDrivenMapB200::SetParameters(const unsigned int parId, const float parVal) {
  Check(b200_edm_set_parameter(mpHandle, parId, (double)parVal), "SetParameters");
  if (mPrint) std::cout << "Parameter value set to " << parVal << std::endl;
}
// Every ComputeF of a solve sees the same ensemble (common random numbers); nothing to re-seed.
void EventDrivenMapB200::ResetSeed() {}
void EventDrivenMapB200::SetNewSeed() {
  Check(b200_edm_new_seed(mpHandle), "SetNewSeed");
  if (mPrint) std::cout << "New seed set" << std::endl;
}
void EventDrivenMapB200::SetSeed(unsigned long long seed) { Check(b200_edm_set_seed(mpHandle, seed), "SetSeed"); }
void EventDrivenMapB200::SetProfileMode(unsigned int nCoarse) {
  Check(b200_edm_set_profile_mode(mpHandle, nCoarse), "SetProfileMode");
}
void EventDrivenMapB200::SetDevices(const int* deviceIds, unsigned int nDevices) {
  Check(b200_edm_set_devices(mpHandle, deviceIds, nDevices), "SetDevices");
  mNoDevices = nDevices;
}
void EventDrivenMapB200::SetDebugFlag(const bool val) {
  Check(b200_edm_set_debug(mpHandle, val ? 1 : 0), "SetDebugFlag");
  if (mPrint) std::cout << (val ? "Debugging on" : "Debugging off") << std::endl;
}
