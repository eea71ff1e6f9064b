// test_report.h -- result accumulation and the summary text of the CLI harness (C++14).
// Same fields, formulas and labels as the reference (h/test_report.h:13-53, src/test_report.cpp:96-135), so the
// output of a run can be diffed against the reference binary's.
#pragma once

#include "../ldpc_code.h"
#include "../ldpc_decoder_gpu.h"

#include <cstdint>
#include <string>
#include <vector>

namespace ldpc_b200 {

struct run_summary {
  std::string code_and_channel_specs;
  uint32_t num_vectors_per_run = 0, num_runs = 0, frame_size = 0, target_errors = 0;
  uint32_t min_iter = uint32_t(-1), max_iter = 0;
  float avg_iter = 0, iter_time_per_vector = 0;
  double elapsed_time = 0;
  uint32_t vectors_with_errors = 0, vectors_with_error_above_target = 0, max_bit_error = 0, num_bit_errors = 0;
  std::string text() const; // gen_summary()
};

std::string describe_code_and_channel(const ldpc_code &code, int channel_kind, float noise, float snr_or_zero);
std::string describe_error_stats(uint32_t n_frames, uint32_t offset, const std::vector<uint32_t> &errors);

} // namespace ldpc_b200
