#include "scht_error.h"

#include "../../include/scht.h"

namespace scht {
std::string& last_error_slot() {
    static thread_local std::string s;
    return s;
}
int fail(int code, const std::string& msg) {
    last_error_slot() = msg;
    return code;
}
}  // namespace scht

extern "C" const char* scht_last_error(void) { return scht::last_error_slot().c_str(); }
extern "C" int scht_abi_version(void) { return SCHT_ABI_VERSION; }
