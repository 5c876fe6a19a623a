// lensing.qest on the device -- placeholder until the estimator kernels land.
#include "ox_common.cuh"
using namespace ox;
extern "C" {
int ox_qeplan_create(ox_geometry *, int, const double *, const double *, const double *, int, int, int, ox_qeplan **) {
  set_error("ox_qeplan_create: not implemented yet");
  return OX_ERR_UNSUPPORTED;
}
int ox_qeplan_destroy(ox_qeplan *) { return OX_OK; }
int ox_qe_reconstruct(ox_qeplan *, const void *, const void *, int, int, int, int, void *, int) {
  set_error("ox_qe_reconstruct: not implemented yet");
  return OX_ERR_UNSUPPORTED;
}
}
