// Library-wide pieces of libryolo_b200.so: error string, version, device sanity.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

extern "C" {

void ryolo_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

const char* ryolo_last_error(void) { return g_err; }

int ryolo_abi_version(void) { return 1; }

// Returns RYOLO_OK only on a compute-capability 10.x device (the library ships sm_100a SASS only).
int ryolo_check_device(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) { ryolo_set_error(cudaGetErrorString(e)); return RYOLO_ERR_CUDA; }
  if (p.major != 10) { ryolo_set_error("libryolo_b200 is built for sm_100a (B200) only"); return RYOLO_ERR_INVALID; }
  return RYOLO_OK;
}

}  // extern "C"
