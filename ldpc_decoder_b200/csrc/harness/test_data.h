// test_data.h -- test-frame generation of the CLI harness (C++14): the reference's create_data
// (src/main.cpp:450-538), channel noise (src/channel.cpp:20-30, 55-68) and capacities (:70-101).
#pragma once

#include "../ldpc_code.h"

#include <cstdint>
#include <vector>

namespace ldpc_b200 {

// kind: 0 = BSC (noise = flip probability), 1 = BI-AWGN (noise = sigma)           [main.cpp:196-215]
float channel_capacity(int kind, float noise);

// Frames start_idx + batch_idx*n_vec + [0, n_vec): random words (NOT codewords), BPSK 1 -> +1, channel noise,
// punctured tail = 0, target syndromes H x.  Seeds are absolute frame indices, so any frame can be regenerated.
//   noisy      float[N * n_vec], element [v + n_vec*i]                 (or IEEE half bits when fp16)
//   ref_frames u32[n_vec * N/32] frame-major, LSB first
//   syndromes  u32[n_vec * ceil(M/32)] frame-major
void create_data(const ldpc_code &code, int kind, float noise, uint32_t start_idx, uint32_t n_vec, uint32_t batch_idx,
                 bool fp16, void *noisy, uint32_t *ref_frames, uint32_t *syndromes);

} // namespace ldpc_b200
