// Library-wide state of the C-ABI: thread-local error text and launch counter.  No other global state
// (re-entrant across devices, SURVEY 8b "Threading").
#include <stdlib.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
thread_local long long g_launches = 0;
}  // namespace

namespace sib {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
static int g_pdl = -1;   // -1: not decided yet (environment), 0 / 1: explicit
bool pdl_enabled() {
  if (g_pdl < 0) g_pdl = getenv("SIB_NO_PDL") == nullptr ? 1 : 0;
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
}  // namespace sib

extern "C" int sib_abi_version(void) { return SIB_ABI_VERSION; }
extern "C" const char* sib_last_error(void) { return g_err; }
extern "C" long long sib_launch_count(void) { return g_launches; }
extern "C" int sib_set_pdl(int enabled) {
  const int prev = sib::pdl_enabled() ? 1 : 0;
  sib::set_pdl(enabled);
  return prev;
}
