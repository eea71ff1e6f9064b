// IEEE half <-> float helpers of include/ldpc_b200_dropin.hpp (used for channels converted on the host in an fp16
// build) against numpy: usage  cpp_half_test floats.bin out.bin   (float32 in, uint16 out, then 65536 float32 bit patterns)
#include "../include/ldpc_b200_dropin.hpp"

#include <cstdio>
#include <vector>

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  std::vector<float> in;
  float x;
  while (std::fread(&x, 4, 1, f) == 1) in.push_back(x);
  std::fclose(f);
  std::FILE *o = std::fopen(argv[2], "wb");
  if (!o) return 4;
  for (float v : in) {
    const uint16_t h = ldpc_b200_dropin::float_to_half_bits(v);
    std::fwrite(&h, 2, 1, o);
  }
  for (uint32_t h = 0; h < 65536; ++h) {
    const float v = ldpc_b200_dropin::half_bits_to_float(uint16_t(h));
    std::fwrite(&v, 4, 1, o);
  }
  std::fclose(o);
  return 0;
}
