// emu_capi.cpp -- TEST INFRASTRUCTURE ONLY: the C ABI of include/mce_b200.h over the sequential emulation
// backend (see emu_backend.h).  Never linked into the product library.
#include "emu_backend.h"
#define MCE_BACKEND mce::EmuBackend
#define MCE_VERSION_STRING "mce-emu (test only)"
#include "../../cauchyfriendly_b200/csrc/mce_capi_impl.h"
