// ldpc_decoder_gpu_cuda.h -- drop-in for the reference's h/ldpc_decoder_gpu_cuda.h (src/main.cpp:5): the same class as
// ldpc_decoder_gpu.h of this directory (the reference's two decoder classes have identical public interfaces).
#include "ldpc_decoder_gpu.h"
