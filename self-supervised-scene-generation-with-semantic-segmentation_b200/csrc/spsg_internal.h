// Shared by the translation units of libspsg_raycast.so; not part of the C ABI.
#ifndef SPSG_INTERNAL_H_
#define SPSG_INTERNAL_H_
#include <cuda_runtime.h>

// record a thread-local error message (returned by spsg_last_error()) and hand back the code
int spsg_internal_fail(int code, const char *msg);
int spsg_internal_fail_cuda(cudaError_t e, const char *where);
// number of SMs of the current device (cached per device; 148 on a B200)
int spsg_internal_sm_count();

#define SPSG_CUDA_TRY(x)                                             \
    do {                                                             \
        cudaError_t e_ = (x);                                        \
        if (e_ != cudaSuccess) return spsg_internal_fail_cuda(e_, #x); \
    } while (0)
#endif
