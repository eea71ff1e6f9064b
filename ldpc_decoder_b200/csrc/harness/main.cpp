// main.cpp -- command-line self-test of the B200 decoder with the reference's flags and output
// (src/main.cpp:54-563): -f alist -c channel -n noise -p log2(frames in flight) -m loading factor -i max iterations
// -e / -b error target -r runs -s first frame index -l log level, plus -x 16|32 (message precision; the reference
// chooses it at compile time with USE_FLOAT16_COMPUTE).  Frames are generated exactly as the reference generates
// them (ChaCha8 streams keyed by the absolute frame index), so a run can be compared with the reference binary.
#include "../ldpc_decoder_gpu.h"
#include "half.h"
#include "test_data.h"
#include "test_report.h"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

using namespace ldpc_b200;
using std::cout;
using std::endl;

static void print_usage() {
  cout << "options: " << endl;
  cout << " -b f where f is the bit error rate above which a frame is considered to be in error; alternative to -e; default is 0" << endl;
  cout << " -c n where n defines the channel: 0 for bsc, 1 for awgn" << endl;
  cout << " -e n where n is the number of bit errors above which a frame is considered to be in error; alternative to -b; default is 0" << endl;
  cout << " -f s where s is the name of the code file" << endl;
  cout << " -h to display this help" << endl;
  cout << " -i n where n is the maximum number of iterations per vector of the decoding algorithm; default is 100" << endl;
  cout << " -l n where n is the log level, from 1 to 3 included. default 1." << endl;
  cout << " -m n where, if k vectors are decoded in parallel by the GPU, n*k vectors are decoded in each run; default is 4" << endl;
  cout << " -n f where f is the noise level of the simulated channel" << endl;
  cout << " -p n where n is the log2 of the maximum number of vectors decoded in parallel by the GPU; default is 5" << endl;
  cout << " -r n where n is the number of decoding runs; default is 1" << endl;
  cout << " -s n where n is the first vector sequence index (seed for rngs), in order to reproduce a test" << endl;
  cout << " -x n where n is 16 or 32: message precision (default 16, the reference's default build)" << endl;
  cout << " Option parameters are either i(n)tegers, (f)loating-point values or (s)trings" << endl;
}

int main(int argc, char **argv) {
  std::string code_filename;
  float noise = 0;
  uint32_t num_runs = 1, start_index = 0, target_errors = 0;
  int channel_idx = 0, log_level = 1, precision_bits = 16;
  double target_ber = 0;
  ldpc_decoder_gpu_static_parameters static_p;
  ldpc_decoder_gpu_dynamic_parameters dyn_p;
  bool channel_defined = false, noise_defined = false, error_defined = false, ber_defined = false, bad = false;
  double noise_arg = 0;

  for (int i = 1; i < argc && !bad; ++i) {
    if (std::strlen(argv[i]) != 2 || argv[i][0] != '-') {
      bad = true;
      break;
    }
    const char opt = argv[i][1];
    if (opt == 'h') {
      print_usage();
      return EXIT_SUCCESS;
    }
    const char *param = i + 1 < argc ? argv[i + 1] : nullptr;
    if (!std::strchr("bcefilmnprsx", opt)) {
      cout << "unrecognized argument" << endl;
      return EXIT_FAILURE;
    }
    if (!param) {
      bad = true;
      break;
    }
    ++i;
    switch (opt) {
    case 'b': ber_defined = true; target_ber = atof(param); break;
    case 'c': channel_defined = true; channel_idx = atoi(param); break;
    case 'e': error_defined = true; target_errors = uint32_t(atoi(param)); break;
    case 'f': code_filename = param; break;
    case 'i': dyn_p.m_num_iter_max = uint32_t(atoi(param)); break;
    case 'l': log_level = atoi(param); bad = log_level < 1 || log_level > 3; break;
    case 'm': dyn_p.m_loading_factor = uint32_t(atoi(param)); break;
    case 'n': noise_defined = true; noise_arg = atof(param); break;
    case 'p': static_p.m_max_log_parallel_factor_user = uint32_t(atoi(param)); break;
    case 'r': num_runs = uint32_t(atoi(param)); break;
    case 's': start_index = uint32_t(atoi(param)); break;
    case 'x': precision_bits = atoi(param); bad = precision_bits != 16 && precision_bits != 32; break;
    }
  }
  if (bad) {
    print_usage();
    return EXIT_FAILURE;
  }
  const bool fp16 = precision_bits == 16;
  static_p.m_use_float16 = fp16;
  // the reference parses the noise level into transfer_llr_t: in its fp16 build sigma itself is a half
  // (main.cpp:57,163 -- README prints 0.939941 for -n 0.94)
  noise = fp16 ? half_bits_to_float(float_to_half_bits(float(noise_arg))) : float(noise_arg);

  cout << "Code file name:" << code_filename << endl;
  if (num_runs == 0) {
    cout << "0 runs to perform, exiting" << endl;
    return EXIT_SUCCESS;
  }
  bool user_error = false;
  if (error_defined && ber_defined) {
    cout << "Cannot define both bit error rate and bit error count" << endl;
    user_error = true;
  }
  if (dyn_p.m_loading_factor == 0) {
    cout << "Invalid overloading factor" << endl;
    user_error = true;
  }
  if (!channel_defined || !noise_defined) {
    cout << "Missing mode and/or channel parameters" << endl;
    user_error = true;
  }
  if (code_filename.empty()) {
    cout << "You have to enter a filename with option -f (filename)." << endl;
    user_error = true;
  }
  std::unique_ptr<noisy_channel> channel;
  if (channel_idx == 0) channel.reset(new bsc_channel(noise));
  else if (channel_idx == 1) channel.reset(new biawgn_channel(noise));
  else {
    cout << "Unknown channel type specified" << endl;
    user_error = true;
  }
  if (user_error) {
    print_usage();
    return EXIT_FAILURE;
  }

  try {
    const ldpc_code code(code_filename, true);
    const uint32_t frame_sz = uint32_t(code.n_inputs());
    dyn_p.m_target_errors = target_errors > 0 ? target_errors : uint32_t(double(frame_sz) * target_ber);
    cout << "Target number of errors per frame: " << dyn_p.m_target_errors << endl << endl;

    // do_test (main.cpp:301-448)
    ldpc_decoder_gpu dec(code, *channel, static_p);
    cout << "Chosen parallel factor: " << dec.parallel_factor() << " vectors decoded in parallel" << endl;
    dyn_p.m_num_vectors_per_run = dec.parallel_factor() * dyn_p.m_loading_factor;
    const uint32_t n_vec = dyn_p.m_num_vectors_per_run;
    const int64_t W = (frame_sz + 31) >> 5, S = ((code.n_outputs() - code.n_erased_outputs()) + 31) >> 5;
    run_summary rep;
    const float snr = channel_idx == 1 ? 1 / (noise * noise) : 0.f;
    rep.code_and_channel_specs = describe_code_and_channel(code, channel_idx, noise, snr);
    rep.num_runs = num_runs;
    rep.num_vectors_per_run = n_vec;
    rep.frame_size = frame_sz;
    rep.target_errors = dyn_p.m_target_errors;
    std::vector<uint32_t> ref_frames(size_t(W) * n_vec), results(size_t(W) * n_vec), syndromes(size_t(S) * n_vec);
    std::vector<uint8_t> noisy(size_t(frame_sz) * n_vec * (fp16 ? 2 : 4));
    cout << "Performing a test with " << num_runs << " run(s)" << endl;
    cout << "Number of vectors (or frames) per run: " << n_vec << endl;
    cout << "Total syndrome size per batch: " << (code.n_outputs() - code.n_erased_outputs()) * int64_t(n_vec) << " bits" << endl;
    cout << "Total data size per batch: " << code.n_inputs() * int64_t(n_vec) << " bits" << endl << endl;
    for (uint32_t run = 0; run < num_runs; ++run) {
      cout << "Creating and processing frame batch " << run << " / " << num_runs << endl;
      cout << " Creating test vectors" << endl;
      auto t0 = std::chrono::steady_clock::now();
      create_data(code, channel_idx, noise, start_index, n_vec, run, fp16, noisy.data(), ref_frames.data(), syndromes.data());
      cout << " Test vector computation time: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << endl;
      cout << " Decoding" << endl;
      test_report tr;
      t0 = std::chrono::steady_clock::now();
      dec.decode(dyn_p, n_vec, noisy.data(), syndromes.data(), results.data(), tr, uint32_t(log_level));
      rep.elapsed_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      rep.max_iter = tr.max_iter;
      rep.min_iter = tr.min_iter;
      rep.avg_iter = tr.avg_iter;
      rep.iter_time_per_vector = tr.iter_time_per_vector;
      if (log_level >= 1)
        cout << "Iterations (avg / max / min): " << tr.avg_iter << " " << tr.max_iter << " " << tr.min_iter << endl;
      cout << " Computing errors after EC" << endl;
      std::vector<uint32_t> errors(n_vec, 0);
      for (uint32_t v = 0; v < n_vec; ++v)
        for (int64_t w = 0; w < W; ++w) {
          const uint32_t cnt = uint32_t(__builtin_popcount(ref_frames[size_t(v) * W + w] ^ results[size_t(v) * W + w]));
          errors[v] += cnt;
          rep.num_bit_errors += cnt;
        }
      cout << "  Errors after error correction " << describe_error_stats(n_vec, start_index + n_vec * run, errors);
      for (uint32_t v = 0; v < n_vec; ++v) {
        if (errors[v] > 0) rep.vectors_with_errors++;
        if (errors[v] > rep.target_errors) rep.vectors_with_error_above_target++;
        rep.max_bit_error = std::max(rep.max_bit_error, errors[v]);
      }
      cout << endl;
    }
    cout << "End of decoding test" << endl << endl;
    cout << rep.text();
  } catch (std::exception &e) {
    cout << e.what() << endl;
  }
  return EXIT_SUCCESS;
}
