// Device side of programmatic dependent launch (see launch_k in host_util.h): first statement of every kernel.
#pragma once

namespace gg {

// Wait until the preceding kernel of the stream has completed and its writes are visible, then allow the
// following kernel to be scheduled (its own pdl_entry() keeps it from touching memory until this one is done).
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
}

}  // namespace gg
