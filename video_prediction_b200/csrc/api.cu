// Error reporting + version for the C ABI.
#include "common.h"

namespace vp {
static thread_local char g_err[512] = "";
int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
}  // namespace vp

extern "C" long long vp_launch_count(void) { return vp::g_launches; }
extern "C" const char* vp_last_error(void) { return vp::g_err; }
extern "C" int vp_version(void) { return 100; }
