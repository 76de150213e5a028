// abi.cu -- library identification, error string and launch counter of libb200pc.so.
#include <stdarg.h>
#include <string.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace b200

extern "C" int b200_abi_version(void) { return 1; }
extern "C" const char *b200_last_error(void) { return b200::g_err; }
extern "C" unsigned long long b200_launch_count(void) { return b200::g_launches.load(); }
