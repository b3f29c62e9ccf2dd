// One of the translation units libkucd.so is built from (see launch.cuh, KUCD_SPLIT_BUILD, and _lib.py: build): the explicit
// instantiations of the float32-grade chain kernels.  No code of its own.
#define KUCD_SPLIT_BUILD 1
#define KUCD_INST_UNIT 1
#include "chain.cuh"

namespace kucd {
template const void* chain_kernel_ptr<256, 2, false, KUCD_PRECISE_CH>();
template const void* chain_kernel_ptr<64, 1, false, KUCD_PRECISE_CH>();
}  // namespace kucd
