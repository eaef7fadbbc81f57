// Library-level entry points of libh3d: version, error text, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace h3d {

static thread_local char g_error[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }

}  // namespace h3d

extern "C" {

int h3d_version(void) { return 100; }

const char* h3d_last_error(void) { return h3d::g_error; }

unsigned long long h3d_launch_count(void) { return h3d::g_launches.load(); }

void h3d_reset_launch_count(void) { h3d::g_launches.store(0); }

}  // extern "C"
