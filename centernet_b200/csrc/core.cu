// Library-wide state: thread-local error text, launch counter, version.
#include "common.cuh"

namespace cnb {

static thread_local char t_err[512] = {0};
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

}  // namespace cnb

extern "C" {

int cnb_version(void) { return 100; /* 0.1.0 */ }
const char *cnb_last_error(void) { return cnb::t_err; }
unsigned long long cnb_launch_count(void) { return cnb::g_launches.load(); }

}  // extern "C"
