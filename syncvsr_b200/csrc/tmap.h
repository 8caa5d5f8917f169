// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled resolved through the runtime's
// driver entry point, so the .so carries no link-time libcuda dependency and loads on CPU boxes).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace svsr {

// bf16 tensor, up to 5 dims, dims[0] is the contiguous one. strides_bytes[i] is the byte stride of
// dims[i+1] (rank-1 entries). box/elem_strides are per-dim. swizzle128 selects CU_TENSOR_MAP_SWIZZLE_128B.
// Returns 0 or a negative svsr::Status (error text in svsr_last_error()).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                   bool swizzle128);

}  // namespace svsr
