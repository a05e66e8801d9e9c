// Library-wide state of the C ABI: error string, launch counter, version.
#include "dm_common.cuh"

namespace dm {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace dm

extern "C" int dm_version(void) { return 100; }
extern "C" const char* dm_last_error(void) { return dm::g_err; }
extern "C" unsigned long long dm_launch_count(void) { return dm::g_launches.load(); }
extern "C" void dm_reset_launch_count(void) { dm::g_launches.store(0); }
