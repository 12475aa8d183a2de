// Thread-local last-error slot shared by the host builder and the CUDA side of the C ABI.
#ifndef SCHT_ERROR_H
#define SCHT_ERROR_H
#include <string>

namespace scht {
std::string& last_error_slot();
int fail(int code, const std::string& msg);
}  // namespace scht
#endif
