// One of the translation units libkucd.so is built from (see launch.cuh, KUCD_SPLIT_BUILD, and _lib.py: build): the explicit
// instantiations of the contraction kernels with the kEpiProb epilogue.  No code of its own.
#define KUCD_SPLIT_BUILD 1
#define KUCD_INST_UNIT 1
#include "launch.cuh"

namespace kucd {
template cudaError_t launch_bn<kEpiProb>(const GemmParams&, int, bool, int, bool, bool, int, cudaStream_t);
}  // namespace kucd
