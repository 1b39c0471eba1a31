// Version / error strings of the C ABI.
#include "common.cuh"

namespace pqlb { unsigned long long g_launches = 0; int g_rec_stride_mode = 0; }

/* Measurement switch, to be set before any ring is created: 0 = power-of-two record stride (default, 1 KB for
 * AllegroHand), 1 = next multiple of 128 bytes (896 B).  DESIGN.md section 2 has the numbers behind the default. */
extern "C" void pqlb_record_stride_mode(int mode) { pqlb::g_rec_stride_mode = mode == 1 ? 1 : 0; }

extern "C" int pqlb_version(void) { return 100; }
extern "C" uint64_t pqlb_launch_count(void) { return pqlb::g_launches; }

extern "C" const char* pqlb_error_string(int code) {
  switch (code) {
    case PQLB_OK: return "ok";
    case PQLB_E_ARG: return "invalid argument (null pointer, non-positive size or bad enum)";
    case PQLB_E_SHAPE: return "inconsistent sizes";
    case PQLB_E_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case PQLB_E_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case PQLB_E_UNSUPPORTED: return "unsupported configuration";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}
