// Library-level entry points: ABI version, thread-local error text, launch counter.
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "common.h"

namespace vsb {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VSB_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace vsb

extern "C" int vsb_abi_version(void) { return VSB_ABI_VERSION; }
extern "C" const char* vsb_last_error(void) { return vsb::g_err; }
extern "C" uint64_t vsb_launch_count(void) { return vsb::g_launches.load(std::memory_order_relaxed); }
