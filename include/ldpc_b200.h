/*
 * ldpc_b200.h -- C ABI of the B200-native LDPC flood decoder (libldpc_b200.so).
 *
 * This is the drop-in boundary for the reference's decoder class
 *   h/ldpc_decoder_gpu.h:15-139  (CUDA twin h/ldpc_decoder_gpu_cuda.h:11-145)
 * of kunzjacq/ldpc_decoder.  The reference has no FFI layer of its own; the
 * header-compatible C++14 class (ldpc_b200_dropin.hpp, instantiated in the global
 * namespace by ldpc_decoder_gpu.h / ldpc_decoder_gpu_cuda.h of this directory for the
 * reference tree's own types, and in namespace ldpc_b200 by
 * ldpc_decoder_b200/csrc/ldpc_decoder_gpu.h) forwards to these entry points (see
 * INTEGRATION.md).
 *
 * Plain pointers and sizes only; no exceptions cross this boundary: every
 * function returns an ldpc_b200_status and ldpc_b200_last_error() gives text.
 * There is no CPU fallback: without a CUDA device every call that needs one
 * returns LDPC_B200_ERR_CUDA.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ldpc_b200_status {
  LDPC_B200_OK = 0,
  LDPC_B200_ERR_INVALID = 1,     /* bad argument                                              */
  LDPC_B200_ERR_CODE = 2,        /* "Incorrect code structure" (ldpc_decoder_gpu.cu:40-58)    */
  LDPC_B200_ERR_MULTIPLE_32 = 3, /* N % 32 != 0 (ldpc_decoder_gpu.cu:30-32)                   */
  LDPC_B200_ERR_CUDA = 4,        /* CUDA runtime failure / no device                          */
  LDPC_B200_ERR_NOMEM = 5,       /* does not fit in device memory                             */
  LDPC_B200_ERR_IO = 6,          /* alist file could not be opened (ldpc_code.cpp:26-29)      */
  LDPC_B200_ERR_ALIST = 7        /* malformed alist (ldpc_code.cpp:114-117)                   */
} ldpc_b200_status;

typedef enum ldpc_b200_precision {
  LDPC_B200_FP32 = 0, /* llr_t = float  (reference built without USE_FLOAT16_COMPUTE) */
  LDPC_B200_FP16 = 1  /* llr_t = __half (reference CMake default, common.h:17-19)     */
} ldpc_b200_precision;

typedef enum ldpc_b200_channel {
  LDPC_B200_AWGN = 0, /* llr = x * param,            param = 2/sigma^2  (flood.cu:62-75, channel.h:70-73) */
  LDPC_B200_BSC = 1,  /* llr = copysign(param, x),   param = ln((1-p)/p) (flood.cu:47-60, channel.cpp:6-9) */
  LDPC_B200_LLR = 2   /* input already holds LLRs (decoding_input_is_llr() == true path)                  */
} ldpc_b200_channel;

/* ---- code graph (replaces ldpc_code, h/ldpc_code.h:10-54) ---------------------------- */
typedef struct ldpc_b200_code ldpc_b200_code;

/* Parse the reference's alist dialect (src/ldpc_code.cpp:45-152). */
ldpc_b200_status ldpc_b200_code_load_alist(const char *path, ldpc_b200_code **out);
ldpc_b200_status ldpc_b200_code_parse_alist(const char *text, size_t len, ldpc_b200_code **out);
/* Build from CSR arrays: row_ptr[M+1], col_idx[E] = variable of each check-major edge. */
ldpc_b200_status ldpc_b200_code_from_csr(uint32_t n_checks, uint32_t n_vars, const uint32_t *row_ptr,
                                         const uint32_t *col_idx, uint32_t n_erased_vars, ldpc_b200_code **out);
void ldpc_b200_code_destroy(ldpc_b200_code *code);

typedef struct ldpc_b200_code_info {
  uint32_t n_checks, n_vars, n_edges, n_erased_vars, n_erased_checks;
  uint32_t max_deg_var, max_deg_check;
  const uint32_t *row_ptr;         /* [M+1] out_bit_to_edge  (check-major CSR offsets)  */
  const uint32_t *col_ptr;         /* [N+1] in_bit_to_edge   (variable-major offsets)   */
  const uint32_t *in_to_out;       /* [E]   edge_in_to_out                              */
  const uint32_t *out_to_in;       /* [E]   edge_out_to_in                              */
  const uint32_t *out_edge_to_var; /* [E]   variable of each check-major edge           */
} ldpc_b200_code_info;
ldpc_b200_status ldpc_b200_code_get_info(const ldpc_b200_code *code, ldpc_b200_code_info *info);
void ldpc_b200_code_set_erased_vars(ldpc_b200_code *code, uint32_t n_erased); /* ldpc_code.cpp:164-167 */

/* ---- decoder (replaces ldpc_decoder_gpu / ldpc_decoder_gpu_cuda) ------------------- */
typedef struct ldpc_b200_decoder ldpc_b200_decoder;

typedef struct ldpc_b200_config {
  int32_t precision;            /* ldpc_b200_precision                                            */
  uint32_t max_log2_frames;     /* m_max_log_parallel_factor_user (ldpc_decoder_gpu_common.h:13)  */
  int32_t device;               /* CUDA device ordinal; -1 = current device                       */
  uint32_t mem_headroom_pct;    /* device memory left unused; 0 -> 10 (ldpc_decoder_gpu.cu:85)    */
  uint32_t storage_order;       /* LDPC_B200_ORDER_*: how the decoder lays its arrays out in HBM  */
  uint32_t reserved[3];
} ldpc_b200_config;

enum {
  LDPC_B200_ORDER_NATURAL = 0,  /* message rows in the alist's check order (the reference's layout, flood.cu:98)        */
  LDPC_B200_ORDER_LOCALITY = 1  /* EXTENSION: checks and variables are STORED in a Cuthill-McKee      *
                                 * order of the graph, so that the message rows a variable-node update gathers are      *
                                 * close together (worth ~3 % of that pass on structured codes -- quasi-cyclic,          *
                                 * spatially coupled, banded; nothing on a random graph).  Inputs, syndromes, results    *
                                 * and every sum's order are unchanged: results are bit-identical to ORDER_NATURAL.      */
};

/* The storage order LDPC_B200_ORDER_LOCALITY uses for this code (host only, no device needed): check_order[M] and
 * var_order[N] list the original check / variable kept at each storage position. */
ldpc_b200_status ldpc_b200_code_locality_order(const ldpc_b200_code *code, uint32_t *check_order, uint32_t *var_order);

/* Constructor: validates the graph, sizes frames-in-flight P = min(2^floor(log2 fit), 2^max_log2_frames)
 * from free device memory (ldpc_decoder_gpu.cu:67-99), allocates and uploads. */
ldpc_b200_status ldpc_b200_decoder_create(const ldpc_b200_code *code, const ldpc_b200_config *cfg,
                                          ldpc_b200_decoder **out);
void ldpc_b200_decoder_destroy(ldpc_b200_decoder *dec);
/* parallel_factor() (ldpc_decoder_gpu.h:120-123) */
uint32_t ldpc_b200_decoder_frames_in_flight(const ldpc_b200_decoder *dec);
/* Allocate the pinned staging mirror, the device ring and the result ring used when decode() is given HOST buffers
 * (otherwise they are allocated by the first such call).  The reference allocates its pinned buffers in the
 * constructor (ldpc_decoder_gpu.cu:119-129); the C++ class calls this from its constructor for the same reason. */
ldpc_b200_status ldpc_b200_decoder_prepare_host_io(ldpc_b200_decoder *dec);
/* set_erased_variables() (ldpc_decoder_gpu.h:125-128) */
void ldpc_b200_decoder_set_erased_vars(ldpc_b200_decoder *dec, uint32_t n_erased);

enum {
  LDPC_B200_LAYOUT_INTERLEAVED = 0, /* element [v + n_frames*i] (reference API, ldpc_decoder_gpu.h:93-94) */
  LDPC_B200_LAYOUT_FRAME_MAJOR = 1, /* element [v*frame_stride + i] (extension)         */
  LDPC_B200_LAYOUT_PACKED_BITS = 2  /* hard-decision input, 1 bit per transmitted variable: u32 words,     *
                                     * frame-major [v*frame_stride + i/32], bit i%32, 1 <-> +1 (extension   *
                                     * f3: 32x less H2D than one float per bit for a BSC, main.cpp:526-527) */
};

typedef struct ldpc_b200_decode_params {
  int32_t channel;                /* ldpc_b200_channel                                                */
  float channel_param;            /* ref_llr() or factor()                                            */
  uint32_t num_iter_max;          /* m_num_iter_max          (default 100)                            */
  uint32_t num_iter_check_parity; /* m_num_iter_check_parity (default 10)                             */
  int32_t input_layout;           /* LDPC_B200_LAYOUT_*                                               */
  uint64_t frame_stride;          /* FRAME_MAJOR / PACKED_BITS: elements (words) between frames;      *
                                   * 0 -> N - n_erased (resp. its word count)                          */
  uint32_t input_pool_frames;     /* 0: input holds n_frames frames.  k>0: input/syndromes hold k      *
                                   * frames and frame f reads column f % k (benchmark recycling pool)  */
  uint32_t flags;                 /* LDPC_B200_FLAG_*                                                  */
  uint32_t queue_lookahead_frames; /* ldpc_b200_decode_from_queue / _multi: frames a decoder keeps reserved   *
                                   * beyond those it has loaded; 0 -> its frames in flight (every check can   *
                                   * then load as many frames as it frees, as in a classic call)              */
  uint32_t reserved[3];
} ldpc_b200_decode_params;

enum {
  LDPC_B200_FLAG_SIMD_PAIRS = 1u,  /* refill counts rounded down to even, as the reference's OpenCL *
                                    * backend does (ldpc_decoder_gpu.cpp:623)                       */
  LDPC_B200_FLAG_TIME_KERNELS = 2u, /* bracket every check-node / variable-node launch with CUDA    *
                                     * events on the launching stream and report cn_s / vn_s        *
                                     * (implies plain launches instead of the CUDA graph)            */
  LDPC_B200_FLAG_NO_GRAPH = 4u,     /* launch the per-period kernel sequence directly instead of     *
                                     * replaying its CUDA graph                                      */
  LDPC_B200_FLAG_HALF_INPUT = 16u,  /* EXTENSION ("fp16 transfer with fp32 compute"; the          *
                                     * reference sets transfer_llr_t = llr_t "for simplicity", h/common.h:35-37):   *
                                     * `input` holds IEEE halves although the decoder is FP32; they are widened on   *
                                     * the device when a frame is loaded.  Same layouts, half the H2D bytes; results  *
                                     * equal the FP32 decode of the half-rounded inputs bit for bit.                  */
  LDPC_B200_FLAG_EARLY_RETIRE = 8u  /* EXTENSION (off by default; the reference keeps every slot      *
                                     * iterating until the last frame of the call stops,              *
                                     * ldpc_decoder_gpu.cu:419-432): once the input is exhausted, a      *
                                     * frame is retired at the check where it stops and the frames      *
                                     * still running are compacted into the low slots, so the tail of   *
                                     * a call touches only as many slots as it needs.  Decisions of      *
                                     * frames that hit num_iter_max are then taken at that check.       */
};

typedef struct ldpc_b200_report {
  uint32_t max_iter, min_iter; /* test_report fields written by decode (ldpc_decoder_gpu.cu:616-628) */
  float avg_iter;
  float iter_time_per_vector;
  uint32_t global_iters;       /* flood iterations executed                                  */
  uint32_t n_checks_run;       /* syndrome checks executed                                   */
  double elapsed_s;            /* wall time of the call                                      */
  double device_s;             /* CUDA-event time first kernel -> last result                */
  double cn_s, vn_s;           /* accumulated CUDA-event time of the CN / VN kernels         */
  uint64_t cn_launches, vn_launches, other_launches;
  uint64_t h2d_bytes, d2h_bytes;
  /* LDPC_B200_FLAG_TIME_KERNELS only: host wall time (stream-synchronised) of the per-check service steps */
  double check_s;  /* syndrome check kernel + flag read-back                         */
  double retire_s; /* extraction of retiring frames (+ D2H when results are on host) */
  double refill_s; /* staging (+ H2D) and refill kernels of new frames               */
  uint64_t frames_refilled;
  uint64_t graph_launches; /* CUDA-graph replays of the per-period sequence (kernels inside are counted above) */
  uint64_t slot_iterations; /* sum over flood iterations of the slots processed (P each, fewer once EARLY_RETIRE     *
                             * has compacted the tail): x bytes per slot-iteration = algorithmic bytes moved         */
  /* The first check-node pass of every check period also builds the start messages of the frames loaded at that   *
   * check (fused refill).  Those launches are included in cn_launches; their CUDA-event time is kept apart from    *
   * cn_s so that cn_s / (cn_launches - cn_refill_launches) is the plain check-node pass.                           */
  uint64_t cn_refill_launches;
  double cn_refill_s;
  uint64_t frames_decoded; /* frames this call decoded (n_frames; less for a decoder sharing a queue with others) */
  /* host-resident buffers: wall time the calling thread spent gathering frames into the pinned staging mirror and   *
   * enqueueing their H2D copies (the reference's prepare_vectors + transfer_vectors), and the part of it that came    *
   * before the first kernel could start (the first batch)                                                            */
  double host_stage_s, first_batch_stage_s;
} ldpc_b200_report;

/* decode() (ldpc_decoder_gpu.h:102-110; semantics of ldpc_decoder_gpu.cu:283-634).
 * input     : channel values, llr_t elements (float, or IEEE half bits when FP16); host or device pointer.
 * syndromes : u32[n_frames * ceil(M/32)] frame-major, bit i of word w = check 32w+i; host or device.
 * results   : u32[n_frames * N/32] frame-major, LSB-first; host or device.
 * iter_counts (nullable): per-frame iter_end - iter_start; host.
 * n_frames == 0 returns immediately (ldpc_decoder_gpu.cu:293-294). */
ldpc_b200_status ldpc_b200_decode(ldpc_b200_decoder *dec, const ldpc_b200_decode_params *params, uint32_t n_frames,
                                  const void *input, const uint32_t *syndromes, uint32_t *results,
                                  ldpc_b200_report *report, uint32_t *iter_counts);

/* ---- one frame queue, several GPUs --------------------------------------------------------------
 * The reference drives device 0 only (h/cuda_manager.h:52).  Frames are independent, so several decoders -- one per
 * GPU -- can work through ONE queue of frames: a queue is an atomic cursor over the rows of the caller's arrays; each
 * decoder reserves whole blocks of `block_frames` consecutive frames as its slots free up and decodes them as one
 * continuous call (the reference's schedule: first batch, frame replacement at every check, one tail at the end of the
 * queue), writing results / iter_counts rows of exactly the frames it decoded.  No collective on the data path.
 *   shm_name == NULL : queue private to this process (threads, one per decoder: ldpc_b200_decode_multi does that);
 *   shm_name != NULL : POSIX shared-memory object shared by the processes of one host (one process per GPU): one
 *                      process passes create = 1 (with the sizes), the others open it afterwards with create = 0. */
typedef struct ldpc_b200_queue ldpc_b200_queue;
ldpc_b200_status ldpc_b200_queue_create(const char *shm_name, int32_t create, uint32_t n_frames, uint32_t block_frames,
                                        ldpc_b200_queue **out);
void ldpc_b200_queue_destroy(ldpc_b200_queue *queue);
uint32_t ldpc_b200_queue_size(const ldpc_b200_queue *queue);
void ldpc_b200_queue_reset(ldpc_b200_queue *queue); /* cursor back to frame 0 (no decoder may be drawing from it) */
/* Reserve the next block by hand (what a decoder does internally; for callers that schedule frames themselves and for
 * tests): returns the number of frames reserved (0: queue exhausted) and their first frame index. */
uint32_t ldpc_b200_queue_take(ldpc_b200_queue *queue, uint32_t *first_frame);
/* Like ldpc_b200_decode, with input / syndromes / results / iter_counts sized for the WHOLE queue (or a pool, see
 * input_pool_frames); only the rows of the frames this decoder took are read and written; report->frames_decoded
 * says how many those were.  Call it concurrently, once per decoder. */
ldpc_b200_status ldpc_b200_decode_from_queue(ldpc_b200_decoder *dec, const ldpc_b200_decode_params *params,
                                             ldpc_b200_queue *queue, const void *input, const uint32_t *syndromes,
                                             uint32_t *results, ldpc_b200_report *report, uint32_t *iter_counts);
/* One process, n_decs decoders (one per device, same code and precision): a private queue over the n_frames frames, one
 * host thread per decoder, a merged report (counters summed, longest elapsed / device time, iteration statistics over
 * all frames) and optionally each decoder's own (per_decoder[n_decs], nullable).  block_frames = 0 picks a default. */
ldpc_b200_status ldpc_b200_decode_multi(ldpc_b200_decoder *const *decs, uint32_t n_decs,
                                        const ldpc_b200_decode_params *params, uint32_t n_frames, const void *input,
                                        const uint32_t *syndromes, uint32_t *results, ldpc_b200_report *report,
                                        ldpc_b200_report *per_decoder, uint32_t *iter_counts, uint32_t block_frames);

/* ---- kernel-level entry points (parity tests and profiling; same kernels decode() runs) ------ */
/* Load cnt <= P frames into slots [0,cnt): staged[j + cnt*i] channel values (ldpc_decoder_gpu.cu:218-273). */
ldpc_b200_status ldpc_b200_debug_refill(ldpc_b200_decoder *dec, int32_t channel, float channel_param, uint32_t cnt,
                                        const void *staged, const uint32_t *syndromes);
/* n flood iterations (check-node pass then variable-node pass); the last variable-node pass also emits hard
 * decisions when final_bits != 0.  Optionally returns the accumulated CUDA-event time of each pass. */
ldpc_b200_status ldpc_b200_debug_iterate(ldpc_b200_decoder *dec, uint32_t n_iters, int32_t final_bits,
                                         float *cn_ms, float *vn_ms);
/* one check-node / variable-node pass only */
ldpc_b200_status ldpc_b200_debug_cn_pass(ldpc_b200_decoder *dec);
ldpc_b200_status ldpc_b200_debug_vn_pass(ldpc_b200_decoder *dec, int32_t final_bits);
/* syndrome check on the current hard decisions: violated[P] bytes (flood.cu:191-223) */
ldpc_b200_status ldpc_b200_debug_check(ldpc_b200_decoder *dec, uint8_t *violated);
/* messages of one slot in check-major edge order, widened to float; priors likewise */
ldpc_b200_status ldpc_b200_debug_read_messages(ldpc_b200_decoder *dec, uint32_t slot, float *out_e);
ldpc_b200_status ldpc_b200_debug_read_priors(ldpc_b200_decoder *dec, uint32_t slot, float *out_n);
/* packed hard decisions of one slot, u32[N/32] LSB-first (flood.cu:277-295) */
ldpc_b200_status ldpc_b200_debug_read_hard(ldpc_b200_decoder *dec, uint32_t slot, uint32_t *out_words);
/* the 32768-entry fp16 phi_abs table the FP16 kernels use (index = |x| half bits) */
ldpc_b200_status ldpc_b200_phi_table_fp16(ldpc_b200_decoder *dec, uint16_t *table);
/* element-wise phi_abs on the device, fp32 path (flood.cu:31-37) */
ldpc_b200_status ldpc_b200_debug_phi_abs_f32(const float *x, float *y, size_t n);

/* ---- on-device synthesis of test frames --------------------------------------------------- */
/* Generates, in DEVICE memory, the frames the reference harness would generate on the host for absolute frame
 * indices start_idx .. start_idx+n_frames-1 (src/main.cpp:450-538: ChaCha8 streams keyed by the frame index, random
 * words, BPSK, channel noise, punctured tail = 0, syndromes = H x), ready to be passed to ldpc_b200_decode().
 *   channel: LDPC_B200_AWGN (noise = sigma) or LDPC_B200_BSC (noise = flip probability)
 *   d_input     llr_t[N * n_frames], element [v + n_frames*i]     (float, or half when the decoder is FP16)
 *   d_syndromes u32[n_frames * ceil(M/32)] frame-major
 *   d_ref_bits  u32[n_frames * N/32] frame-major, LSB first (the transmitted words, to count errors against) */
ldpc_b200_status ldpc_b200_synthesize_frames(ldpc_b200_decoder *dec, int32_t channel, float noise, uint64_t start_idx,
                                             uint32_t n_frames, void *d_input, uint32_t *d_syndromes,
                                             uint32_t *d_ref_bits);

/* algorithmic bytes of one flood iteration of ONE frame: (4E+N)*s + 4*ceil(M/32) */
uint64_t ldpc_b200_bytes_per_frame_iteration(const ldpc_b200_decoder *dec);

const char *ldpc_b200_last_error(void);
const char *ldpc_b200_version(void);
/* sizeof of the four public structs as this library was built, in the order ldpc_b200_code_info, ldpc_b200_config,
 * ldpc_b200_decode_params, ldpc_b200_report: a binding checks its own mirror of the layouts against them. */
void ldpc_b200_abi_sizes(uint32_t sizes[4]);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
