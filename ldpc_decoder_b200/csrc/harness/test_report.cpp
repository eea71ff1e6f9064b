#include "test_report.h"

#include "test_data.h"

#include <algorithm>
#include <iomanip>
#include <sstream>

namespace ldpc_b200 {

std::string describe_code_and_channel(const ldpc_code &code, int kind, float noise, float snr) {
  std::ostringstream o;
  const float capacity = channel_capacity(kind, noise);
  o << "Channel:\n";
  if (kind == 0) o << "Binary channel with bit error probability: " << noise << "\n";            // channel.cpp:16-19
  else o << "Binary channel with Gaussian noise of std. deviation " << noise << "; SNR = " << snr << "\n"; // :49-52
  o << "capacity: " << capacity << " bits/symbol\n\n";
  o << "Error-correcting code:\n"
    << code.n_inputs() << " variables\n"
    << code.n_outputs() << " parity bits\n"
    << code.n_erased_inputs() << " erased variables (not sent, but recovered)\n"
    << "maximum input bit arity: " << code.max_degree_in() << "\n"
    << "maximum output/check bit arity: " << code.max_degree_out() << "\n"
    << "Rate = " << rate(code) << "\n\n";
  const float eff = rate(code) / capacity * 100;
  std::ostringstream e;
  e << std::fixed << std::setprecision(2) << eff;
  o << "Code efficiency over channel = rate/channel capacity = " << e.str() << "%\n";
  return o.str();
}

std::string describe_error_stats(uint32_t n, uint32_t offset, const std::vector<uint32_t> &errors) {
  std::ostringstream o;
  if (n > 1) {
    uint32_t lo = uint32_t(-1), hi = 0;
    double total = 0;
    for (uint32_t v = 0; v < n; ++v) {
      total += errors[v];
      lo = std::min(lo, errors[v]);
      hi = std::max(hi, errors[v]);
    }
    o << "on vectors " << offset << " ... " << offset + n - 1 << ":\n"
      << "  total = " << total << ", average = " << total / n << ", min = " << lo << ", max = " << hi << "\n";
  } else {
    o << "on frame " << offset << ": " << errors[0] << "\n";
  }
  return o.str();
}

std::string run_summary::text() const {
  std::ostringstream r;
  const size_t bits_per_run = size_t(frame_size) * size_t(num_vectors_per_run);
  const size_t bits = size_t(num_runs) * bits_per_run;
  const double ber = double(num_bit_errors) / double(bits);
  const double mbits = double(bits >> 20);
  const uint32_t frames = num_runs * num_vectors_per_run;
  r << "                                            ***\n"
    << "                                          Summary \n\n"
    << "* Channel and code description\n\n"
    << code_and_channel_specs << "\n\n"
    << "* Test result\n\n"
    << "# of frames decoded:              " << frames << "\n"
    << "Frame size:                       " << frame_size << " bits\n"
    << "Total # of errors:                " << num_bit_errors << "\n"
    << "Bit error rate (BER):             " << ber << "\n"
    << "Maximum # of errors / frame:      " << max_bit_error << "\n";
  if (target_errors > 0)
    r << "Frames with more than " << target_errors << " errors:  " << vectors_with_error_above_target
      << " (corresponding FER: " << double(vectors_with_error_above_target) / double(frames) << ")\n";
  r << "Frames with at least one error:   " << vectors_with_errors
    << " (corresponding FER: " << double(vectors_with_errors) / double(frames) << ")\n\n"
    << "Mbits processed:                  " << mbits << "\n"
    << "Elapsed system time:              " << elapsed_time << " sec.\n"
    << "Throughput including transfers and finish: " << mbits / elapsed_time << " Mbits/sec.\n"
    << "Max/min/average number of iterations per vector: " << max_iter << "/" << min_iter << "/" << avg_iter << "\n"
    << "Iteration time per vector (i.e. iteration time / vector batch size): " << iter_time_per_vector << " sec\n"
    << "Decoding throughput: " << frame_size / (avg_iter * iter_time_per_vector * 1048576.) << " Mbits/sec.\n\n";
  return r.str();
}

} // namespace ldpc_b200
