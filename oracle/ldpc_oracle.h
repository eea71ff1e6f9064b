/*
 * ldpc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference decoder's flood (sum-product) hot
 * path (kunzjacq/ldpc_decoder).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may link or call this.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Data layouts are the reference's own:
 *   messages      llr[E*P]   index slot + P*out_edge      (check-major edge id)
 *   priors        llr[N*P]   index slot + P*var
 *   syndromes     u32[S*P]   index slot + P*word          (S = ceil(M/32))
 *   final bits    u8[N*P]    index slot + P*var
 *
 * Parity pinning: the reference holds no golden vectors, so this
 * oracle is pinned against the reference's own CUDA backend compiled from
 * /root/reference into oracle/_ref (see oracle/Makefile, oracle/ref_driver.cu)
 * and run on a B200; the resulting fixtures live in tests/golden/.
 */
#ifndef LDPC_ORACLE_H
#define LDPC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- code graph: src/ldpc_code.cpp:45-152 --------------------------------- */
typedef struct oracle_code {
  int64_t n_checks;        /* M  (m_n_outputs) */
  int64_t n_vars;          /* N  (m_n_inputs)  */
  int64_t n_edges;         /* E  */
  int64_t n_erased_vars;   /* "#e=" header     */
  int64_t n_erased_checks; /* "#ec=" header    */
  int32_t max_deg_var, max_deg_check;
  uint32_t *out_bit_to_edge; /* [M+1] CSR offsets (check-major)        */
  uint32_t *in_bit_to_edge;  /* [N+1] CSC offsets (variable-major)     */
  uint32_t *edge_out_to_in;  /* [E]                                    */
  uint32_t *edge_in_to_out;  /* [E]                                    */
  uint32_t *in_edge_to_bit;  /* [E] variable owning each in-edge       */
  uint32_t *out_edge_to_bit; /* [E] check owning each out-edge         */
  uint32_t *out_edge_to_in_bit; /* [E] variable of each out-edge (ldpc_decoder_gpu.cu:60-65) */
} oracle_code;

/* returns NULL on failure; *err (if non-NULL) gets 1 = cannot open, 2 = malformed */
oracle_code *oracle_code_from_alist_file(const char *path, int *err);
oracle_code *oracle_code_from_alist_text(const char *text, int *err);
void oracle_code_free(oracle_code *c);
float oracle_code_rate(const oracle_code *c); /* ldpc_code.cpp:241-254 */

/* H * bits for one frame; bits/syndrome packed LSB-first u32 (ldpc_code.cpp:256-286) */
void oracle_compute_syndrome(const oracle_code *c, const uint32_t *bits_packed, uint32_t *syndrome_packed);

/* ---- numerics ------------------------------------------------------------- */
/* pre_threshold: 1e-5f for the CUDA fp32 backend (flood.cu:14), 2*exp(-(thr+1))
 * for the OpenCL backend (flood_vec2.cl:187). */
float oracle_phi_abs(float x, float pre_threshold); /* flood.cu:31-37  */
float oracle_phi(float x, float pre_threshold);     /* flood.cu:40-45  */

/* fp16 flavour (flood.cu:21-29): phi_abs on raw half bits, CPU-evaluated
 * (float libm, one rounding to half per reference operation). */
uint16_t oracle_phi_abs_h(uint16_t x_bits);
/* fills table[32768] with oracle_phi_abs_h(i) */
void oracle_phi_table_h(uint16_t *table);

/* ---- decoder -------------------------------------------------------------- */
enum { ORACLE_AWGN = 0, ORACLE_BSC = 1 };
enum { ORACLE_FP32 = 0, ORACLE_FP16 = 1 };

typedef struct oracle_decoder oracle_decoder;

/* precision: ORACLE_FP32 / ORACLE_FP16.  log2_frames: P = 2^log2_frames slots.
 * pre_threshold only matters for fp32.  phi_table_h (fp16 only) may be NULL
 * (CPU table) or a 32768-entry table exported by the device library so that
 * both sides evaluate the identical phi. */
oracle_decoder *oracle_decoder_create(const oracle_code *c, int precision, uint32_t log2_frames,
                                      float pre_threshold, const uint16_t *phi_table_h, int n_threads);
void oracle_decoder_free(oracle_decoder *d);
/* fp32 only: switch the CN/VN passes to the SIMD-over-frames kernels (ldpc_cpu_fast.c).  Same algorithm and
 * schedule; libmvec/fast-math arithmetic, so results agree with the exact kernels to ~1e-5, not bit for bit.
 * Used for the CPU baseline timing (what a vectorising OpenCL CPU runtime would do with flood_vec2.cl). */
void oracle_decoder_set_fast(oracle_decoder *d, int fast);

/* Kernel-level entry points (each is one reference kernel over all P slots). */
/* transfer_vectors (ldpc_decoder_gpu.cu:218-273): stage cnt new frames, convert to LLR, refill slots [0,cnt).
 * input: element [j + cnt*i] = i-th channel value of new frame j (float, or half bits when fp16), i < N-n_erased.
 * syndromes: frame-major u32[cnt*S]. */
void oracle_refill(oracle_decoder *d, int channel, float channel_param, uint32_t cnt, const void *input,
                   const uint32_t *syndromes);
void oracle_backward(oracle_decoder *d);                /* flood.cu:77-115   */
void oracle_forward(oracle_decoder *d, int final_bits); /* flood.cu:117-189  */
void oracle_check_parity(oracle_decoder *d, uint8_t *violated /*[P]*/); /* flood.cu:191-223 */
void oracle_permute(oracle_decoder *d, uint32_t n, const uint32_t *origin, const uint32_t *dest); /* flood.cu:225-275 */
void oracle_deinterlace(oracle_decoder *d, uint32_t *packed /*[P*(N/32)]*/); /* flood.cu:277-295 */
/* copy one slot's messages out as float (fp16 is widened) / raw */
void oracle_read_messages(const oracle_decoder *d, uint32_t slot, float *out /*[E]*/);
void oracle_read_priors(const oracle_decoder *d, uint32_t slot, float *out /*[N]*/);

typedef struct oracle_report {
  uint32_t max_iter, min_iter;
  float avg_iter;
  float iter_time_per_vector;
  uint32_t global_iters;
  double elapsed_s;
} oracle_report;

/* decode(): ldpc_decoder_gpu.cu:283-634 (simd_pairs=1 adds the OpenCL even-count
 * refill rule, ldpc_decoder_gpu.cpp:623).  input: element [v + n_frames*i]
 * (reference API layout).  iter_counts (nullable): per-frame iter_end-iter_start. */
void oracle_decode(oracle_decoder *d, int channel, float channel_param, uint32_t num_iter_max,
                   uint32_t num_iter_check_parity, uint32_t n_frames, const void *input,
                   const uint32_t *syndromes, uint32_t *results, oracle_report *report, uint32_t *iter_counts,
                   int simd_pairs);

#ifdef __cplusplus
}
#endif
#endif
