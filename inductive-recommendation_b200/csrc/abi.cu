// Library-wide state of libb200rec.so: error text, launch counter, version.
#include <stdlib.h>
#include "common.cuh"

namespace b200rec {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200REC_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}
}  // namespace b200rec

extern "C" const char* b200rec_last_error(void) { return b200rec::g_err; }
extern "C" int b200rec_version(void) { return 100; }
extern "C" uint64_t b200rec_launch_count(void) { return b200rec::g_launches.load(); }
