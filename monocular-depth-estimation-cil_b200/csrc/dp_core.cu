// ABI version, thread-local error string and the launch counter.
#include "common.cuh"
#include "../../include/depth_b200.h"
#include <atomic>
#include <cstring>
#include <cstdlib>

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int dp_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int dp_pdl_enabled(void) {
  static const int on = [] { const char* e = getenv("DP_PDL"); return e ? (atoi(e) != 0) : 1; }();
  return on;
}

static thread_local double g_pdl_hint = 0.0;
void dp_pdl_hint(double bytes_equivalent) { g_pdl_hint = bytes_equivalent; }
int dp_pdl_take(void) {
  static const double max_bytes = [] { const char* e = getenv("DP_PDL_MAX_MB"); return (e ? atof(e) : 256.0) * 1e6; }();
  const double h = g_pdl_hint;
  g_pdl_hint = 0.0;                      // one launch per hint; launches without a hint are short helper kernels
  return dp_pdl_enabled() && h <= max_bytes;
}

void dp_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

extern "C" {
int dp_abi_version(void) { return DP_ABI_VERSION; }
const char* dp_last_error(void) { return g_err; }
unsigned long long dp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}
