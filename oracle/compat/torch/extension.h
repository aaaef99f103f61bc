// Build-recipe shim for oracle/build_ref.py (test infrastructure): lets the UNMODIFIED reference source
// torch/utils/depth_utils/depth_utils_cuda_kernel.cu compile against torch 2.11.  That file calls
// AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...), an idiom from torch <= 1.x: `Tensor::type()` now returns
// at::DeprecatedTypeProperties, which no longer converts to c10::ScalarType.  The macro is re-pointed at the scalar type of
// that object; nothing else changes (the kernels only instantiate float accessors).
#pragma once
#include_next <torch/extension.h>

namespace spsg_ref_compat {
inline c10::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
inline c10::ScalarType scalar_type(c10::ScalarType t) { return t; }
}  // namespace spsg_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
    AT_DISPATCH_SWITCH(spsg_ref_compat::scalar_type(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
