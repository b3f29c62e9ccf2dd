// Host-side print-outs of the data-path helpers of csrc/aux_kernels.cuh - the very functions the kernels call on the
// device - so that tests/test_data_path.py can compare them with the oracle's restatement without a GPU.
//   data_path_host perm <rows> <seed> <epoch>   -> the shuffling permutation (make_feistel_key + feistel_permute)
//   data_path_host bits                         -> for every byte value b and valid count 0..8:
//                                                "b valid w0 w1 w2 w3 back" (bits_to_bf16x8, bf16x8_to_bits round trip)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../keras_unsupervised_b200/csrc/aux_kernels.cuh"

int main(int argc, char** argv) {
  if (argc == 5 && strcmp(argv[1], "perm") == 0) {
    const long long rows = atoll(argv[2]);
    const unsigned long long seed = strtoull(argv[3], nullptr, 10), epoch = strtoull(argv[4], nullptr, 10);
    const kucd::FeistelKey key = kucd::make_feistel_key(seed, epoch, rows);
    for (long long i = 0; i < rows; ++i)
      printf("%llu\n", static_cast<unsigned long long>(
                           kucd::feistel_permute(static_cast<uint64_t>(i), static_cast<uint64_t>(rows), key)));
    return 0;
  }
  if (argc == 2 && strcmp(argv[1], "bits") == 0) {
    for (unsigned b = 0; b < 256; ++b)
      for (int valid = 0; valid <= 8; ++valid) {
        const kucd::Bf16x8 v = kucd::bits_to_bf16x8(b, valid);
        printf("%u %d %u %u %u %u %u\n", b, valid, v.w[0], v.w[1], v.w[2], v.w[3], kucd::bf16x8_to_bits(v, 8));
      }
    return 0;
  }
  return 2;
}
