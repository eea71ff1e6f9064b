// ldpc_decoder_gpu.h -- drop-in for the reference's h/ldpc_decoder_gpu.h AND h/ldpc_decoder_gpu_cuda.h.
//
// Put this directory BEFORE the reference's h/ on the include path (-I<b200>/include -I<reference>/h): the reference's
// own sources (src/main.cpp:4-5, 315-320, 403-405) then compile unchanged against the B200 decoder; the only other
// change is to link libldpc_b200.so instead of src/ldpc_decoder_gpu.cu, src/cuda/flood.cu and
// src/cuda/cuda_manager.cpp.  Everything else -- ldpc_code, noisy_channel / bsc_channel / biawgn_channel, the two
// parameter classes, test_report, error -- is taken from the headers found next on the include path (the reference's
// own h/channel.h, h/ldpc_code.h, h/ldpc_decoder_gpu_common.h, h/test_report.h, h/common.h), in the global namespace.
// Precision follows the reference's compile-time macro USE_FLOAT16_COMPUTE.  INTEGRATION.md section 2;
// tests/test_gpu_parity.py::test_reference_main_cpp_builds_and_runs_against_the_dropin_header.
#ifndef LDPC_B200_DROPIN_LDPC_DECODER_GPU_H
#define LDPC_B200_DROPIN_LDPC_DECODER_GPU_H

#include "channel.h"
#include "ldpc_code.h"
#include "ldpc_decoder_gpu_common.h"
#include "test_report.h"

#include "ldpc_b200_dropin.hpp"

struct ldpc_b200_reference_tree_types {
  typedef ::ldpc_code code_t;
  typedef ::noisy_channel channel_t;
  typedef ::bsc_channel bsc_t;
  typedef ::biawgn_channel awgn_t;
  typedef ::ldpc_decoder_gpu_static_parameters static_params_t;
  typedef ::ldpc_decoder_gpu_dynamic_parameters dynamic_params_t;
  typedef ::test_report report_t;
  typedef ::error error_t;
  static const int bsc_id = int(::bsc), awgn_id = int(::awgn);
};

typedef ldpc_b200_dropin::decoder<ldpc_b200_reference_tree_types> ldpc_decoder_gpu;
// the CUDA-flavoured name do_test instantiates under CUDA_DECODER (main.cpp:315-319)
typedef ldpc_decoder_gpu ldpc_decoder_gpu_cuda;

#endif
