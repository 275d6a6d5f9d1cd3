#include "lfi_common.cuh"
namespace lfi {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }
static long g_launches = 0;
void count_launches(long n) { g_launches += n; }
long get_launches() { return g_launches; }
}  // namespace lfi
extern "C" const char *lfi_last_error(void) { return lfi::get_error(); }
extern "C" int lfi_abi_version(void) { return LFI_ABI_VERSION; }
extern "C" long lfi_launch_count(void) { return lfi::get_launches(); }
