// Error state and small queries of the C-ABI.
#include "tg_common.cuh"
#include <stdarg.h>

static thread_local char g_err[1024] = "";

void tg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* tg_last_error(void) { return g_err; }
extern "C" int tg_version(void) { return 200; }
// sizeof of the descriptor structs as THIS library was compiled: a binding whose struct
// definition is stale (too short / differently padded) can detect it before passing a pointer
extern "C" int64_t tg_sizeof_win(void) { return (int64_t)sizeof(tg_win); }
extern "C" int64_t tg_sizeof_basis(void) { return (int64_t)sizeof(tg_basis); }

extern "C" int tg_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}

// number of kernels this library has launched (bench.py's gpu_launches)
static long long g_launches = 0;
void tg_count_launch() { g_launches++; }
extern "C" int64_t tg_launch_count(void) { return (int64_t)g_launches; }
