// One of the translation units libkucd.so is built from (see launch.cuh, KUCD_SPLIT_BUILD, and _lib.py: build): the explicit
// instantiations of the bf16 chain kernels.  No code of its own.
#define KUCD_SPLIT_BUILD 1
#define KUCD_INST_UNIT 1
#include "chain.cuh"

namespace kucd {
template const void* chain_kernel_ptr<256, 2, false, 0>();
template const void* chain_kernel_ptr<256, 2, true, 0>();
template const void* chain_kernel_ptr<64, 1, false, 0>();
template const void* chain_kernel_ptr<64, 1, true, 0>();
template const void* chain_kernel_ptr<256, 1, false, 0>();
template const void* chain_kernel_ptr<128, 1, false, 0>();
}  // namespace kucd
