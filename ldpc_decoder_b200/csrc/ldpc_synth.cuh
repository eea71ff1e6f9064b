// ldpc_synth.cuh -- on-device synthesis of the reference harness' test frames.
// Same streams as the reference generator (src/main.cpp:450-538): ChaCha8 keyed by the absolute frame index
// (src/prng_chacha.cpp:18-67, 1536-byte refills, nonce = refill counter), one PRNG word per (variable, group of 32
// frames) for the bits, a per-frame stream (seed | 2^32) for the noise, unit() / polar-method gaussian() of
// h/rng.h:38-70, BSC / BI-AWGN noise of src/channel.cpp:26-30, 65-68, punctured tail = 0, syndromes = H x.
// Bits, syndromes and BSC outputs are bit-identical to the CPU generator; AWGN samples differ only where CUDA's
// logf (1 ulp) differs from glibc's.
#pragma once

#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ldpc_b200 {
namespace dev {

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }

#define LDPC_QR(a, b, c, d)                                                                                            \
  a += b; d = rotl32(d ^ a, 16); c += d; b = rotl32(b ^ c, 12); a += b; d = rotl32(d ^ a, 8); c += d; b = rotl32(b ^ c, 7);

// one 64-byte ChaCha8 block: key words 0,1 = seed, counter = block index inside the refill, nonce = refill index
__device__ __forceinline__ void chacha8_block(uint64_t seed, uint32_t block, uint64_t refill, uint32_t out[16]) {
  const uint32_t in[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u, uint32_t(seed), uint32_t(seed >> 32), 0, 0,
                           0, 0, 0, 0, block, 0, uint32_t(refill), uint32_t(refill >> 32)};
  uint32_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3], x4 = in[4], x5 = in[5], x6 = in[6], x7 = in[7], x8 = in[8],
           x9 = in[9], x10 = in[10], x11 = in[11], x12 = in[12], x13 = in[13], x14 = in[14], x15 = in[15];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    LDPC_QR(x0, x4, x8, x12) LDPC_QR(x1, x5, x9, x13) LDPC_QR(x2, x6, x10, x14) LDPC_QR(x3, x7, x11, x15)
    LDPC_QR(x0, x5, x10, x15) LDPC_QR(x1, x6, x11, x12) LDPC_QR(x2, x7, x8, x13) LDPC_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + in[0]; out[1] = x1 + in[1]; out[2] = x2 + in[2]; out[3] = x3 + in[3];
  out[4] = x4 + in[4]; out[5] = x5 + in[5]; out[6] = x6 + in[6]; out[7] = x7 + in[7];
  out[8] = x8 + in[8]; out[9] = x9 + in[9]; out[10] = x10 + in[10]; out[11] = x11 + in[11];
  out[12] = x12 + in[12]; out[13] = x13 + in[13]; out[14] = x14 + in[14]; out[15] = x15 + in[15];
}

// bit-sliced random words: sliced[i * n_groups + grp] = i-th PRNG word of the stream seeded first + 32*grp
// (main.cpp:476-487).  One thread = one ChaCha block (16 consecutive variables of one group).
__global__ void synth_bits_kernel(uint32_t *__restrict__ sliced, uint64_t first, uint32_t n_vars, uint32_t n_groups) {
  const uint32_t blocks_per_group = (n_vars + 15) / 16;
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= uint64_t(blocks_per_group) * n_groups) return;
  const uint32_t grp = uint32_t(t % n_groups), blk = uint32_t(t / n_groups);
  uint32_t w[16];
  chacha8_block(first + 32ull * grp, blk % 24u, blk / 24u, w);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t i = blk * 16u + k;
    if (i < n_vars) sliced[size_t(i) * n_groups + grp] = w[k];
  }
}

struct FrameRng { // per-thread sequential stream (rng.h state machine)
  uint64_t seed, refill;
  uint32_t block, pos;
  uint32_t w[16];
  bool has_next;
  float next;
  __device__ void init(uint64_t s) {
    seed = s;
    refill = 0;
    block = 0;
    pos = 0;
    has_next = false;
    chacha8_block(seed, 0, 0, w);
  }
  __device__ uint32_t random_int() {
    if (pos == 16) {
      pos = 0;
      if (++block == 24) {
        block = 0;
        ++refill;
      }
      chacha8_block(seed, block, refill, w);
    }
    uint32_t v = 0; // w[pos] with a register-resident array: select instead of dynamic indexing
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k == int(pos)) v = w[k];
    ++pos;
    return v;
  }
  __device__ float unit() { return __fmul_rn(__fadd_rn(float(random_int()), 0.5f), 2.3283064365386963e-10f); }
  __device__ float gaussian() {
    float result;
    if (has_next) {
      result = next;
    } else {
      float x, y, sq;
      do {
        x = __fadd_rn(__fmul_rn(2.f, unit()), -1.f);
        y = __fadd_rn(__fmul_rn(2.f, unit()), -1.f);
        sq = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
      } while (sq >= 1.f || sq == 0.f);
      const float modulus = __fsqrt_rn(__fdiv_rn(__fmul_rn(-2.f, logf(sq)), sq));
      result = __fmul_rn(x, modulus);
      next = __fmul_rn(y, modulus);
    }
    has_next = !has_next;
    return result;
  }
};

// One thread = one frame: channel outputs in the API layout input[v + n_frames*i] and the packed transmitted word.
template <typename Elem>
__global__ void synth_noise_kernel(const uint32_t *__restrict__ sliced, Elem *__restrict__ input,
                                   uint32_t *__restrict__ ref_bits, uint64_t first, uint32_t n_frames, uint32_t n_vars,
                                   uint32_t n_tx, uint32_t n_groups, int channel, float noise) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_frames) return;
  FrameRng r;
  r.init((first + v) | (1ull << 32)); // main.cpp:522
  const uint32_t W = n_vars >> 5;
  uint32_t word = 0;
  for (uint32_t i = 0; i < n_vars; ++i) {
    const uint32_t bit = (__ldg(sliced + size_t(i) * n_groups + (v >> 5)) >> (v & 31u)) & 1u;
    word |= bit << (i & 31u);
    if ((i & 31u) == 31u) {
      ref_bits[size_t(v) * W + (i >> 5)] = word;
      word = 0;
    }
    float out = 0.f;
    if (i < n_tx) {
      const float val = bit ? 1.f : -1.f;
      if (channel == 1) { // BSC
        out = (r.unit() < noise) ? -val : val;
      } else if (sizeof(Elem) == 4) {
        out = __fadd_rn(val, __fmul_rn(r.gaussian(), noise));
      } else { // fp16 build of the reference: gaussian() is rounded to half before the multiply
        out = __fadd_rn(val, __fmul_rn(__half2float(__float2half_rn(r.gaussian())), noise));
      }
    }
    if (sizeof(Elem) == 4) reinterpret_cast<float *>(input)[size_t(v) + size_t(n_frames) * i] = out;
    else reinterpret_cast<__half *>(input)[size_t(v) + size_t(n_frames) * i] = __float2half_rn(out);
  }
}

// H x on the bit-sliced words, 32 frames per XOR (ldpc_code.cpp:256-286)
__global__ void synth_syndrome_kernel(const uint32_t *__restrict__ sliced, uint32_t *__restrict__ syn_sliced,
                                      const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ edge_var,
                                      uint32_t n_checks, uint32_t n_groups) {
  const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= uint64_t(n_checks) * n_groups) return;
  const uint32_t c = uint32_t(t / n_groups), grp = uint32_t(t % n_groups);
  uint32_t x = 0;
  for (uint32_t e = __ldg(row_ptr + c); e < __ldg(row_ptr + c + 1); ++e)
    x ^= __ldg(sliced + size_t(__ldg(edge_var + e)) * n_groups + grp);
  syn_sliced[size_t(c) * n_groups + grp] = x;
}

// bit-sliced [check][group] -> frame-major packed u32[frame][S]: one warp = (block of 32 checks, group)
__global__ void synth_syndrome_transpose_kernel(const uint32_t *__restrict__ syn_sliced, uint32_t *__restrict__ syn,
                                                uint32_t n_checks, uint32_t n_groups, uint32_t n_frames,
                                                uint32_t syn_words) {
  const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (warp >= uint64_t(syn_words) * n_groups) return;
  const uint32_t cb = uint32_t(warp / n_groups), grp = uint32_t(warp % n_groups);
  const uint32_t c = cb * 32u + lane;
  const uint32_t x = c < n_checks ? __ldg(syn_sliced + size_t(c) * n_groups + grp) : 0u;
  uint32_t mine = 0;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const uint32_t r = __ballot_sync(0xFFFFFFFFu, (x >> b) & 1u);
    if (int(lane) == b) mine = r;
  }
  const uint32_t frame = grp * 32u + lane;
  if (frame < n_frames) syn[size_t(frame) * syn_words + cb] = mine;
}

} // namespace dev
} // namespace ldpc_b200
