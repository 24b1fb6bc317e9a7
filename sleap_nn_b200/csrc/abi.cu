// ABI version probe for the ctypes loader.
#include "common.cuh"
extern "C" int snb_abi_version(void) { return SNB_ABI_VERSION; }
