// ABI version probe and host-memory helper for the ctypes loader.
#include "common.cuh"
extern "C" int snb_abi_version(void) { return SNB_ABI_VERSION; }

// Device-visible alias of a PINNED host allocation (cudaHostAlloc / cudaHostRegister), so that a
// kernel can sample it in place over PCIe instead of staging the whole buffer in HBM.
extern "C" int snb_host_device_pointer(void* host_ptr, void** device_ptr) {
  if (!host_ptr || !device_ptr) return SNB_ERR_BAD_ARG;
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, host_ptr, 0) != cudaSuccess) {
    cudaGetLastError();  // clear the sticky error: the buffer is simply not pinned / mapped
    return SNB_ERR_UNSUPPORTED;
  }
  *device_ptr = d;
  return SNB_OK;
}

// sizeof(snb_bottomup_args), so a binding can verify its mirror of the struct layout.
extern "C" int snb_bottomup_args_size(void) { return (int)sizeof(snb_bottomup_args); }
