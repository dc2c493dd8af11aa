// ABI bookkeeping for libcamli_b200.
#include "common.cuh"

extern "C" int camli_abi_version(void) { return 1; }

extern "C" const char* camli_strerror(int code) {
    if (code == CAMLI_OK) return "ok";
    if (code == CAMLI_EINVAL) return "invalid argument (size or null pointer)";
    if (code == CAMLI_EUNSUPPORTED) return "argument outside the limits of the sm_100a kernels";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown camli error";
}
