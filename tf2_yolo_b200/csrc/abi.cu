// ABI version + status strings.
#include "common.cuh"

namespace yb {
const char*& last_error_site() {
    static thread_local const char* site = "";
    return site;
}
}  // namespace yb

extern "C" int yb_abi_version(void) { return YB_ABI_VERSION; }

extern "C" const char* yb_last_error_site(void) { return yb::last_error_site(); }

extern "C" const char* yb_status_string(int status) {
    switch (status) {
        case YB_OK: return "ok";
        case YB_E_NULL: return "required pointer is NULL";
        case YB_E_SHAPE: return "invalid shape (grid / boxes / classes / counts)";
        case YB_E_PARAM: return "invalid parameter (version / mode / enum)";
        case YB_E_WORKSPACE: return "workspace too small or not 256-byte aligned";
        case YB_E_ALIGN: return "tensor pointer misaligned for its element type";
        case YB_E_CAPACITY: return "output capacity too small";
        default: break;
    }
    if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
    return "unknown yolo_b200 status";
}

// Device-side alias of a page-locked HOST allocation (cudaHostAlloc / cudaHostRegister), so that a
// kernel of this library can write its (small) result straight into host memory instead of the
// caller issuing a sized D2H copy after a round trip for the count.
extern "C" int yb_mapped_host_pointer(void* host_ptr, void** device_ptr) {
    if (host_ptr == nullptr || device_ptr == nullptr) return YB_E_NULL;
    return (int)cudaHostGetDevicePointer(device_ptr, host_ptr, 0);
}
