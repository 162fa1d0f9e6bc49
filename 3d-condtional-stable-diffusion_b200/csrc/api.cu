// C-ABI plumbing: version, thread-local error text, device info.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void b200dm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool b200dm_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B200DM_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

extern "C" int b200dm_version(void) { return B200DM_VERSION; }

extern "C" const char* b200dm_last_error(void) { return g_err; }

extern "C" const char* b200dm_storage_dtype(void) { return B200DM_ACT_NAME; }

extern "C" int b200dm_device_info(int* sm_count, int* cc) {
  int dev = 0;
  B2_CHECK_CUDA(cudaGetDevice(&dev));
  int sms = 0, major = 0, minor = 0;
  B2_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B2_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B2_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc) *cc = major * 10 + minor;
  return B200DM_OK;
}
