// ldpc_b200.cu -- C ABI (include/ldpc_b200.h) and host scheduler of the B200 flood decoder.
//
// Replaces, for the decode hot path, the reference's
//   src/ldpc_decoder_gpu.cu:20-634 (constructor, prepare_vectors, transfer_vectors, decode)
//   h/cuda_manager.h + src/cuda/cuda_manager.cpp (runtime wrapper)
// with: 64-bit-indexed [E][P] message arrays, slot indirection instead of flood_permute_vecs,
// bit-sliced syndromes / hard decisions, and refills written straight into freed slots.
#include "../../include/ldpc_b200.h"

#include "ldpc_code.h"
#include "ldpc_kernels.cuh"
#include "ldpc_synth.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

using namespace ldpc_b200;

// ------------------------------------------------------------------ error plumbing
namespace {

thread_local std::string g_last_error;

ldpc_b200_status fail(ldpc_b200_status st, const std::string &msg) {
  g_last_error = msg;
  return st;
}

#define CU_TRY(expr)                                                                                                   \
  do {                                                                                                                 \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess)                                                                                             \
      return fail(LDPC_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                             \
  } while (0)

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T> struct dev_buf { // RAII device allocation
  T *p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    const cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count; // a failed allocation leaves an empty buffer (p == nullptr, n == 0)
    else p = nullptr;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~dev_buf() { release(); }
};

template <typename T> struct pinned_buf { // RAII pinned host allocation
  T *p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    const cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&p), std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = count;
    else p = nullptr;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
  ~pinned_buf() { release(); }
};

bool is_device_pointer(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

uint32_t ilog2(uint64_t x) {
  uint32_t r = 0;
  while ((uint64_t(1) << (r + 1)) <= x) ++r;
  return r;
}

// One 64-byte run of the staging gather.  The pinned mirror is written once and then read by the DMA engine only, so
// the stores bypass the cache (no read-for-ownership of the destination line): measured 1.8x the memcpy rate on the
// strided gather (4 threads, 873813 rows x 64 B runs).  dst must be 16-byte aligned; callers fence before the H2D copy.
inline void copy64_streaming(uint8_t *dst, const uint8_t *src) {
#if defined(__SSE2__)
  const __m128i a0 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
  const __m128i a1 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
  const __m128i a2 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
  const __m128i a3 = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
  _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a0);
  _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), a1);
  _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), a2);
  _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), a3);
#else
  memcpy(dst, src, 64);
#endif
}
inline void streaming_fence() {
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

// host worker threads for the strided gather (prepare_vectors, ldpc_decoder_gpu.cu:199-207)
void parallel_for(size_t n, unsigned n_threads, const std::function<void(size_t, size_t)> &fn) {
  n_threads = std::max(1u, std::min<unsigned>(n_threads, unsigned((n + 4095) / 4096)));
  if (n_threads == 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  const size_t chunk = (n + n_threads - 1) / n_threads;
  for (unsigned t = 0; t < n_threads; ++t) {
    const size_t a = std::min(n, t * chunk), b = std::min(n, a + chunk);
    if (a < b) th.emplace_back(fn, a, b);
  }
  for (auto &t : th) t.join();
}
} // namespace

// ------------------------------------------------------------------ code object
struct ldpc_b200_code {
  std::unique_ptr<ldpc_code> code;
};

extern "C" {

const char *ldpc_b200_last_error(void) { return g_last_error.c_str(); }
const char *ldpc_b200_version(void) { return "ldpc_b200 0.1 (sm_100a)"; }
void ldpc_b200_abi_sizes(uint32_t sizes[4]) {
  if (!sizes) return;
  sizes[0] = sizeof(ldpc_b200_code_info);
  sizes[1] = sizeof(ldpc_b200_config);
  sizes[2] = sizeof(ldpc_b200_decode_params);
  sizes[3] = sizeof(ldpc_b200_report);
}

static ldpc_b200_status wrap_code(std::function<ldpc_code *()> make, ldpc_b200_code **out) {
  if (!out) return fail(LDPC_B200_ERR_INVALID, "null output pointer");
  *out = nullptr;
  try {
    std::unique_ptr<ldpc_code> c(make());
    *out = new ldpc_b200_code{std::move(c)};
    return LDPC_B200_OK;
  } catch (const error &e) {
    const std::string msg = e.what();
    return fail(msg.find("could not be opened") != std::string::npos ? LDPC_B200_ERR_IO : LDPC_B200_ERR_ALIST, msg);
  } catch (const std::exception &e) {
    return fail(LDPC_B200_ERR_INVALID, e.what());
  }
}

ldpc_b200_status ldpc_b200_code_load_alist(const char *path, ldpc_b200_code **out) {
  if (!path) return fail(LDPC_B200_ERR_INVALID, "null path");
  return wrap_code([&] { return new ldpc_code(std::string(path), true); }, out);
}

ldpc_b200_status ldpc_b200_code_parse_alist(const char *text, size_t len, ldpc_b200_code **out) {
  if (!text) return fail(LDPC_B200_ERR_INVALID, "null text");
  return wrap_code([&] { return new ldpc_code(std::string(text, len), false); }, out);
}

ldpc_b200_status ldpc_b200_code_from_csr(uint32_t n_checks, uint32_t n_vars, const uint32_t *row_ptr,
                                         const uint32_t *col_idx, uint32_t n_erased_vars, ldpc_b200_code **out) {
  if (!row_ptr || !col_idx || !n_checks || !n_vars) return fail(LDPC_B200_ERR_INVALID, "bad CSR arguments");
  return wrap_code([&] { return new ldpc_code(n_checks, n_vars, row_ptr, col_idx, n_erased_vars); }, out);
}

void ldpc_b200_code_destroy(ldpc_b200_code *code) { delete code; }

ldpc_b200_status ldpc_b200_code_get_info(const ldpc_b200_code *code, ldpc_b200_code_info *info) {
  if (!code || !info) return fail(LDPC_B200_ERR_INVALID, "null argument");
  const ldpc_code &c = *code->code;
  info->n_checks = uint32_t(c.n_outputs());
  info->n_vars = uint32_t(c.n_inputs());
  info->n_edges = c.n_edges();
  info->n_erased_vars = uint32_t(c.n_erased_inputs());
  info->n_erased_checks = uint32_t(c.n_erased_outputs());
  info->max_deg_var = uint32_t(c.max_degree_in());
  info->max_deg_check = uint32_t(c.max_degree_out());
  info->row_ptr = c.row_ptr().data();
  info->col_ptr = c.col_ptr().data();
  info->in_to_out = c.in_to_out().data();
  info->out_to_in = c.out_to_in().data();
  info->out_edge_to_var = c.out_edge_to_var().data();
  return LDPC_B200_OK;
}

void ldpc_b200_code_set_erased_vars(ldpc_b200_code *code, uint32_t n_erased) {
  if (code) code->code->set_n_erased_in_bits(int32_t(n_erased));
}

} // extern "C"

// ------------------------------------------------------------------ decoder object
struct ldpc_b200_decoder {
  int device = 0;
  int precision = LDPC_B200_FP32;
  uint32_t elem = 4; // sizeof(llr)
  uint32_t M = 0, N = 0, E = 0, S = 0, W = 0, n_erased = 0;
  uint32_t frames_in_flight = 1; // P exposed to callers (parallel_factor)
  uint32_t P = 32, log2P = 5;    // allocated slots (>= 32)
  uint32_t log2G = 3;            // 16-byte vectors per row
  uint32_t log2PW = 0;           // 32-slot words per row
  int n_sms = 1;
  unsigned host_threads = 1;

  dev_buf<uint32_t> d_row_ptr, d_col_ptr, d_in_to_out, d_edge_var;
  dev_buf<uint32_t> d_vn_ell; // padded gather list per variable of the variable-node pass (dev::vn_fetch_list)
  // Locality storage order (ldpc_b200_config::storage_order = 1): message rows, priors, decisions and syndrome rows are
  // kept in a permuted order; all device tables above are then in STORAGE order and these map original indices to it
  // (empty / null: identity).  Host copies serve the read-back entry points and the frame synthesis.
  dev_buf<uint32_t> d_var_pos, d_check_pos, d_row_ptr_orig, d_edge_var_orig;
  std::vector<uint32_t> h_edge_pos, h_var_pos;
  bool reordered = false;
  dev_buf<uint8_t> d_msg, d_llr0, d_stage;
  dev_buf<uint32_t> d_syn, d_hard, d_viol, d_stage_syn, d_retire_mask, d_new_mask, d_slot_entry, d_results;
  dev_buf<dev::RefillEntry> d_refill;
  dev_buf<dev::ExtractEntry> d_extract;
  dev_buf<uint16_t> d_phi_table;
  // device-resident scheduler state (dev::sched_kernel)
  dev_buf<dev::SchedState> d_state;
  dev_buf<dev::SchedSummary> d_summary;
  dev_buf<uint32_t> d_n_new, d_in_gpu, d_virt, d_retired, d_iter; // d_iter: [2][n_frames] start/end, grown on demand
  dev_buf<uint32_t> d_scratch, d_n_moves;                        // early-retire compaction
  dev_buf<dev::SlotMove> d_moves;
  pinned_buf<dev::SchedSummary> h_summary; // [2]
  pinned_buf<uint32_t> h_retired;          // [2][P]
  cudaStream_t copy_stream = nullptr;      // H2D of new frames / D2H of results, overlapped with decoding
  cudaEvent_t ev_staged = nullptr, ev_refilled = nullptr, ev_period[2] = {nullptr, nullptr}, ev_extracted[2] = {nullptr, nullptr};
  pinned_buf<uint8_t> h_stage;
  pinned_buf<uint32_t> h_stage_syn, h_viol, h_results, h_mask, h_retire_mask, h_slot_entry;
  pinned_buf<dev::RefillEntry> h_refill;
  pinned_buf<dev::ExtractEntry> h_extract;

  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  int grid_cn = 1, grid_cn_refill = 1, grid_vn = 1, grid_vnb = 1, grid_misc = 1;
  size_t smem_tab = 0;
  uint32_t tab_last = dev::kPhiTableSize - 1; // first index from which the fp16 phi table is all zero

  ~ldpc_b200_decoder() {
    for (cudaEvent_t e : {ev_a, ev_b, ev_staged, ev_refilled, ev_period[0], ev_period[1], ev_extracted[0], ev_extracted[1]})
      if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }
  const uint32_t *done_flag() const { return &d_state.p->done; }

  // Staging for host-resident buffers is allocated on first use: a decoder fed from device memory never pays for
  // the pinned mirror (N*P*s bytes) nor the device ring.
  // ring_frames: frame positions of the staging ring (P plus the look-ahead slack of host-fed calls, stage_slack)
  ldpc_b200_status ensure_input_staging(uint32_t ring_frames = 0) {
    const size_t need = size_t(N) * std::max<size_t>(P, ring_frames) * elem;
    if (!d_stage.p || d_stage.n < need) CU_TRY(d_stage.alloc(need));
    if (!h_stage.p || h_stage.n < need) CU_TRY(h_stage.alloc(need));
    return LDPC_B200_OK;
  }
  ldpc_b200_status ensure_syndrome_staging(uint32_t ring_frames = 0) {
    const size_t need = size_t(S) * std::max<size_t>(P, ring_frames);
    if (!d_stage_syn.p || d_stage_syn.n < need) CU_TRY(d_stage_syn.alloc(need));
    if (!h_stage_syn.p || h_stage_syn.n < need) CU_TRY(h_stage_syn.alloc(need));
    return LDPC_B200_OK;
  }
  // Look-ahead slack of the staging ring of host-fed calls (LDPC_B200_STAGE_SLACK_PCT, % of the frames in flight): the
  // ring holds frames_in_flight + slack frames, so the frames the NEXT check may load are normally staged a period in
  // advance and the host thread can enqueue the next period's kernels first and gather afterwards -- the strided
  // gather (27 ms per period alone on a host, ~90 ms with peaks above a period when 8 ranks share one,
  // profiles/r06_bench_8gpu_feed.json) leaves the critical path of the launches.  (A first version that kept the gather
  // before the launches and merely made its passes bigger was slower: profiles/README.md.)
  uint32_t stage_slack_pct = 25;
  uint32_t stage_slack(uint32_t batch, uint32_t ring_block) const {
    return uint32_t(uint64_t(batch) * stage_slack_pct / 100) / ring_block * ring_block;
  }
  ldpc_b200_status ensure_result_staging() {
    if (!d_results.p) CU_TRY(d_results.alloc(size_t(W) * P));
    if (!h_results.p) CU_TRY(h_results.alloc(size_t(W) * P));
    return LDPC_B200_OK;
  }

  // ---- launches -----------------------------------------------------------
  // kernel instantiations are chosen once per decoder (precision, register-resident degree D)
  typedef void (*cn_fn)(uint4 *, const uint32_t *, const uint8_t *, const uint16_t *, uint32_t, uint32_t, uint32_t,
                        const uint32_t *, uint32_t);
  typedef void (*vn_fn)(uint4 *, const uint4 *, const uint32_t *, const uint32_t *, uint8_t *, const uint16_t *,
                        uint32_t, uint32_t, uint32_t, const uint32_t *);
  typedef void (*cn_refill_fn)(uint4 *, const uint32_t *, const uint8_t *, const uint16_t *, uint32_t, uint32_t,
                               uint32_t, const uint32_t *, const dev::CnRefillArgs, uint32_t);
  cn_fn k_cn = nullptr;
  cn_refill_fn k_cn_refill = nullptr; // fused refill: first check-node pass of a period (null: refill writes the messages)
  vn_fn k_vn = nullptr, k_vnb = nullptr;
  int deg_cn = 8, deg_vn = 8;
  int cn_parts = 1; // > 1: check rows are split over this many lanes
  int fused_mode = -1; // fused refill: -1 chosen per check period on the device, 0 never, 1 always (LDPC_B200_FUSED_REFILL)
  uint32_t fused_min_pct = 6; // -1 mode: a period fuses when it loads at least this share of the slots (LDPC_B200_FUSED_MIN_PCT)
  int threads_cn = dev::kThreads, threads_vn = dev::kThreads;
  bool cn_bulk = false;   // bulk-async (TMA) staged check-node kernel (LDPC_B200_CN_BULK)
  bool vn_bulk = false;   // bulk-async (TMA) staged variable-node kernel (LDPC_B200_VN_BULK)
  size_t smem_cn = 0;     // dynamic shared memory of the check-node kernel
  size_t smem_vn = 0;     // ... of the variable-node kernel
  // gate (dev::kGate*): run always / only in a period whose refill is fused / only in a period whose refill is not
  void launch_cn(uint32_t gate = dev::kGateNone) {
    k_cn<<<grid_cn, threads_cn, smem_cn, stream>>>(reinterpret_cast<uint4 *>(d_msg.p), d_row_ptr.p,
                                                      reinterpret_cast<const uint8_t *>(d_syn.p), d_phi_table.p,
                                                      tab_last, M, log2G, done_flag(), gate);
  }
  // first check-node pass after refill_kernel<.., kMsgRows = false>: start messages of the new slots come from the priors
  void launch_cn_refill(uint32_t gate = dev::kGateNone) {
    const dev::CnRefillArgs ra{reinterpret_cast<const uint4 *>(d_llr0.p), d_edge_var.p,
                               reinterpret_cast<const uint8_t *>(d_new_mask.p)};
    k_cn_refill<<<grid_cn_refill, dev::kThreads, smem_tab, stream>>>(
        reinterpret_cast<uint4 *>(d_msg.p), d_row_ptr.p, reinterpret_cast<const uint8_t *>(d_syn.p), d_phi_table.p,
        tab_last, M, log2G, done_flag(), ra, gate);
  }
  void launch_vn(bool bits) {
    (bits ? k_vnb : k_vn)<<<bits ? grid_vnb : grid_vn, threads_vn, smem_vn, stream>>>(
        reinterpret_cast<uint4 *>(d_msg.p), reinterpret_cast<const uint4 *>(d_llr0.p), d_vn_ell.p, d_in_to_out.p,
        reinterpret_cast<uint8_t *>(d_hard.p), d_phi_table.p, tab_last, N, log2G, done_flag());
  }
  void launch_check() {
    cudaMemsetAsync(d_viol.p, 0, sizeof(uint32_t) << log2PW, stream);
    dev::syndrome_check_kernel<<<grid_misc, dev::kThreads, 0, stream>>>(d_syn.p, d_hard.p, d_row_ptr.p, d_edge_var.p,
                                                                        d_viol.p, M, log2PW, done_flag());
  }
  // refill kernels: entries, count, masks are already in device memory (scheduler kernel or host upload)
  // staged_src: the source is interleaved (element [i][frame], or the staging ring's blocks of 32 frames)
  // LDPC_B200_REFILL_STAGED=1: stage the channel values of interleaved sources through shared memory (coalesced source
  // reads).  Measured in bench steps (profiles/r06_ab_fused_refill.txt): 0.394 s of refill per step against 0.383 s with
  // per-lane loads -- the refill is bound by the scattered 16-byte read-modify-writes, not by its source reads.  Off.
  bool refill_staged = false;
  void launch_refill_kernels(const void *src, const uint32_t *src_syn, int channel, float param, int src_kind = 0,
                             bool msg_rows = true, uint32_t gate = dev::kGateNone, bool with_syndromes = true,
                             bool staged_src = false) {
    const uint32_t n_regular = N - n_erased;
    uint4 *msg = reinterpret_cast<uint4 *>(d_msg.p), *l0 = reinterpret_cast<uint4 *>(d_llr0.p);
    auto launch = [&](auto kernel, const uint16_t *table) {
      kernel<<<grid_misc, dev::kThreads, 0, stream>>>(msg, l0, src, d_refill.p, d_n_new.p, d_new_mask.p, d_slot_entry.p,
                                                      d_col_ptr.p, d_in_to_out.p, table, N, n_regular, log2G, channel,
                                                      param, done_flag(), gate, d_var_pos.p);
    };
    // priors-only refill of an interleaved source: channel values staged through shared memory (coalesced source reads)
    if (!msg_rows && staged_src && src_kind != dev::kSrcPackedBits && refill_staged && (1u << log2G) >= uint32_t(dev::kThreads)) {
      const size_t es = (precision == LDPC_B200_FP16 || src_kind == dev::kSrcHalf) ? 2 : 4;
      uint32_t rows = 4;
      while (rows > 1 && size_t(rows) * P * es > 64 * 1024) rows >>= 1;
      const size_t smem = size_t(rows) * P * es;
      auto launch_staged = [&](auto kernel) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        kernel<<<grid_misc, dev::kThreads, smem, stream>>>(l0, src, d_refill.p, d_n_new.p, d_new_mask.p, d_slot_entry.p, N,
                                                           n_regular, log2G, channel, param, done_flag(), gate,
                                                           d_var_pos.p, rows);
      };
      if (precision == LDPC_B200_FP16) launch_staged(dev::refill_priors_staged_kernel<dev::TraitsF16, dev::kSrcNative>);
      else if (src_kind == dev::kSrcHalf) launch_staged(dev::refill_priors_staged_kernel<dev::TraitsF32, dev::kSrcHalf>);
      else launch_staged(dev::refill_priors_staged_kernel<dev::TraitsF32, dev::kSrcNative>);
    } else
    // src_kind: dev::kSrcNative / kSrcPackedBits / kSrcHalf (half elements into an fp32 decoder)
    if (precision == LDPC_B200_FP32) {
      if (src_kind == dev::kSrcPackedBits)
        msg_rows ? launch(dev::refill_kernel<dev::TraitsF32, dev::kSrcPackedBits, true>, nullptr)
                 : launch(dev::refill_kernel<dev::TraitsF32, dev::kSrcPackedBits, false>, nullptr);
      else if (src_kind == dev::kSrcHalf)
        msg_rows ? launch(dev::refill_kernel<dev::TraitsF32, dev::kSrcHalf, true>, nullptr)
                 : launch(dev::refill_kernel<dev::TraitsF32, dev::kSrcHalf, false>, nullptr);
      else
        msg_rows ? launch(dev::refill_kernel<dev::TraitsF32, dev::kSrcNative, true>, nullptr)
                 : launch(dev::refill_kernel<dev::TraitsF32, dev::kSrcNative, false>, nullptr);
    } else {
      if (src_kind == dev::kSrcPackedBits)
        msg_rows ? launch(dev::refill_kernel<dev::TraitsF16, dev::kSrcPackedBits, true>, d_phi_table.p)
                 : launch(dev::refill_kernel<dev::TraitsF16, dev::kSrcPackedBits, false>, d_phi_table.p);
      else
        msg_rows ? launch(dev::refill_kernel<dev::TraitsF16, dev::kSrcNative, true>, d_phi_table.p)
                 : launch(dev::refill_kernel<dev::TraitsF16, dev::kSrcNative, false>, d_phi_table.p);
    }
    if (with_syndromes)
      dev::refill_syndrome_kernel<<<grid_misc, dev::kThreads, 0, stream>>>(d_syn.p, src_syn, d_refill.p, d_n_new.p,
                                                                           d_new_mask.p, d_slot_entry.p, M, log2PW,
                                                                           d_check_pos.p);
  }
  // host-built entries (kernel-level entry points): h_refill[0..n) -> device, then the same kernels
  void launch_refill(uint32_t n_new, const void *src, const uint32_t *src_syn, int channel, float param) {
    const uint32_t PW = 1u << log2PW;
    std::fill(h_mask.p, h_mask.p + PW, 0u);
    for (uint32_t t = 0; t < n_new; ++t) {
      const uint32_t s = h_refill.p[t].slot;
      h_mask.p[s >> 5] |= 1u << (s & 31);
      h_slot_entry.p[s] = t;
    }
    h_viol.p[0] = n_new; // pinned scratch word
    cudaMemcpyAsync(d_n_new.p, h_viol.p, sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_refill.p, h_refill.p, sizeof(dev::RefillEntry) * n_new, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_new_mask.p, h_mask.p, sizeof(uint32_t) * PW, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_slot_entry.p, h_slot_entry.p, sizeof(uint32_t) * P, cudaMemcpyHostToDevice, stream);
    launch_refill_kernels(src, src_syn, channel, param);
  }
  // h_extract[slot].dst_offset valid for every slot set in h_retire_mask
  void launch_extract(uint32_t *dst) {
    const uint32_t PW = 1u << log2PW;
    cudaMemcpyAsync(d_retire_mask.p, h_retire_mask.p, sizeof(uint32_t) * PW, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_extract.p, h_extract.p, sizeof(dev::ExtractEntry) * P, cudaMemcpyHostToDevice, stream);
    launch_extract_kernel(dst);
  }
  void launch_extract_kernel(uint32_t *dst) { // masks and destinations already on the device
    dev::extract_kernel<<<grid_misc, 128, 0, stream>>>(d_hard.p, dst, d_retire_mask.p, d_extract.p, W, log2PW,
                                                       d_var_pos.p);
  }
};

namespace {

// smallest D in {4,6,8} such that rows of degree <= D carry >= 90 % of the edges (others take the re-read path)
int pick_reg_degree(const std::vector<uint32_t> &offsets) {
  const size_t n = offsets.size() - 1;
  uint64_t le4 = 0, le6 = 0;
  const uint64_t total = offsets[n];
  for (size_t i = 0; i < n; ++i) {
    const uint32_t deg = offsets[i + 1] - offsets[i];
    if (deg <= 4) le4 += deg;
    if (deg <= 6) le6 += deg;
  }
  if (le4 * 10 >= total * 9) return 4;
  if (le6 * 10 >= total * 9) return 6;
  return 8;
}

// Long check rows (high-rate codes): if more than 10 % of the edges sit in rows longer than 8 and those rows
// average >= 14 edges, split every row over the fewest lanes (2, 4 or 8, eight edges each) that hold the longest
// row; beyond 64 keep the re-read path.  Measured (scripts/kernel_perf.py): (3,32) fp32 4.55 -> 5.33 TB/s,
// (4,16) 5.9 TB/s; at degree 12 the second lane is half empty and the re-read path wins (5.4 vs 4.7 TB/s).
int pick_cn_parts(const std::vector<uint32_t> &row_ptr, int max_deg) {
  const size_t n = row_ptr.size() - 1;
  uint64_t long_edges = 0, long_rows = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t deg = row_ptr[i + 1] - row_ptr[i];
    if (deg > 8) {
      long_edges += deg;
      ++long_rows;
    }
  }
  if (long_edges * 10 <= uint64_t(row_ptr[n]) || long_edges < 14 * long_rows) return 1;
  for (int parts = 2; parts <= 8; parts *= 2)
    if (max_deg <= 8 * parts) return parts;
  return 1;
}

template <class Tr, class Idx> void select_for_precision(ldpc_b200_decoder *d) {
  // check-node pass: split-row kernel for long rows, else the bulk-async experiment if asked for, else the
  // register-resident kernel (with its fused-refill twin)
  if (d->cn_parts > 1) {
    switch (d->cn_parts) {
    case 2: d->k_cn = dev::cn_update_split_kernel<Tr, 2>; break;
    case 4: d->k_cn = dev::cn_update_split_kernel<Tr, 4>; break;
    default: d->k_cn = dev::cn_update_split_kernel<Tr, 8>; break;
    }
  } else if (d->cn_bulk) { // plain passes staged by the bulk-copy engine; the fused-refill pass stays the register kernel
    switch (d->deg_cn) {
    case 4: d->k_cn = dev::cn_update_bulk_kernel<Tr, 4>; d->k_cn_refill = dev::cn_update_refill_kernel<Tr, 4, Idx>; break;
    case 6: d->k_cn = dev::cn_update_bulk_kernel<Tr, 6>; d->k_cn_refill = dev::cn_update_refill_kernel<Tr, 6, Idx>; break;
    default: d->k_cn = dev::cn_update_bulk_kernel<Tr, 8>; d->k_cn_refill = dev::cn_update_refill_kernel<Tr, 8, Idx>; break;
    }
  } else {
    switch (d->deg_cn) {
    case 4: d->k_cn = dev::cn_update_kernel<Tr, 4, Idx>; d->k_cn_refill = dev::cn_update_refill_kernel<Tr, 4, Idx>; break;
    case 6: d->k_cn = dev::cn_update_kernel<Tr, 6, Idx>; d->k_cn_refill = dev::cn_update_refill_kernel<Tr, 6, Idx>; break;
    default: d->k_cn = dev::cn_update_kernel<Tr, 8, Idx>; d->k_cn_refill = dev::cn_update_refill_kernel<Tr, 8, Idx>; break;
    }
  }
  if (d->vn_bulk) {
    switch (d->deg_vn) {
    case 4: d->k_vn = dev::vn_update_bulk_kernel<Tr, false, 4>; d->k_vnb = dev::vn_update_bulk_kernel<Tr, true, 4>; break;
    case 6: d->k_vn = dev::vn_update_bulk_kernel<Tr, false, 6>; d->k_vnb = dev::vn_update_bulk_kernel<Tr, true, 6>; break;
    default: d->k_vn = dev::vn_update_bulk_kernel<Tr, false, 8>; d->k_vnb = dev::vn_update_bulk_kernel<Tr, true, 8>; break;
    }
  } else {
    switch (d->deg_vn) {
    case 4: d->k_vn = dev::vn_update_kernel<Tr, false, 4, Idx>; d->k_vnb = dev::vn_update_kernel<Tr, true, 4, Idx>; break;
    case 6: d->k_vn = dev::vn_update_kernel<Tr, false, 6, Idx>; d->k_vnb = dev::vn_update_kernel<Tr, true, 6, Idx>; break;
    default: d->k_vn = dev::vn_update_kernel<Tr, false, 8, Idx>; d->k_vnb = dev::vn_update_kernel<Tr, true, 8, Idx>; break;
    }
  }
}

ldpc_b200_status select_kernels(ldpc_b200_decoder *d) {
  // 32-bit vector indices when every (row, vector) index of the message and prior arrays fits (see cn_update_kernel)
  const bool idx32 = (uint64_t(std::max(d->E, d->N)) << d->log2G) < (1ull << 32) && !std::getenv("LDPC_B200_IDX64");
  if (d->precision == LDPC_B200_FP32) {
    if (idx32) select_for_precision<dev::TraitsF32, uint32_t>(d);
    else select_for_precision<dev::TraitsF32, size_t>(d);
  } else {
    if (idx32) select_for_precision<dev::TraitsF16, uint32_t>(d);
    else select_for_precision<dev::TraitsF16, size_t>(d);
  }
  if (d->smem_tab) {
    CU_TRY(cudaFuncSetAttribute(d->k_cn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_tab)));
    CU_TRY(cudaFuncSetAttribute(d->k_vn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_tab)));
    CU_TRY(cudaFuncSetAttribute(d->k_vnb, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_tab)));
    if (d->k_cn_refill)
      CU_TRY(cudaFuncSetAttribute(d->k_cn_refill, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_tab)));
  }
  // Fused refill (refill writes only the priors, the first check-node pass of the period builds the start messages):
  // the fused pass holds message and prior vectors of a row at once and costs a few ms more than the plain pass, while
  // it saves ~9 ms in a period that refills 11 % of the slots (measurements: profiles/README.md).  It is therefore
  // chosen PER CHECK PERIOD ON THE DEVICE: the scheduler kernel knows how many frames the period loads and sets
  // SchedState::fused_now when that is at least fused_min_pct % of the slots; both variants of the refill and of the
  // period's first check-node pass are enqueued and the one not chosen exits at once (dev::gated_off).
  // LDPC_B200_FUSED_REFILL=0 / 1 forces it off / on for every period that loads frames.
  d->fused_mode = -1;
  if (const char *fused = std::getenv("LDPC_B200_FUSED_REFILL")) d->fused_mode = atoi(fused) != 0 ? 1 : 0;
  if (const char *pct = std::getenv("LDPC_B200_FUSED_MIN_PCT")) d->fused_min_pct = uint32_t(std::max(0, atoi(pct)));
  if (const char *e = std::getenv("LDPC_B200_REFILL_STAGED")) d->refill_staged = atoi(e) != 0;
  if (d->fused_mode == 0) d->k_cn_refill = nullptr;
  return LDPC_B200_OK;
}

// Locality order of checks and variables: Cuthill-McKee breadth-first numbering of the bipartite graph (neighbours
// by increasing degree, components started from a lowest-degree node), reversed.  Variables whose checks are close
// get close numbers, so the message rows a variable-node update gathers sit in nearby memory.
void locality_order(const ldpc_code &code, std::vector<uint32_t> &check_order, std::vector<uint32_t> &var_order) {
  const uint32_t M = uint32_t(code.n_outputs()), N = uint32_t(code.n_inputs());
  const std::vector<uint32_t> &rp = code.row_ptr(), &ev = code.out_edge_to_var(), &cp = code.col_ptr(), &i2o = code.in_to_out();
  std::vector<uint32_t> edge_check(code.n_edges());
  for (uint32_t c = 0; c < M; ++c)
    for (uint32_t e = rp[c]; e < rp[c + 1]; ++e) edge_check[e] = c;
  auto degree = [&](uint32_t node) { return node < M ? rp[node + 1] - rp[node] : cp[node - M + 1] - cp[node - M]; };
  std::vector<uint32_t> order, by_degree(size_t(M) + N), nb;
  order.reserve(size_t(M) + N);
  std::vector<uint8_t> seen(size_t(M) + N, 0);
  for (uint32_t i = 0; i < M + N; ++i) by_degree[i] = i;
  std::stable_sort(by_degree.begin(), by_degree.end(), [&](uint32_t a, uint32_t b) { return degree(a) < degree(b); });
  for (uint32_t start : by_degree) {
    if (seen[start]) continue;
    seen[start] = 1;
    size_t head = order.size();
    order.push_back(start);
    while (head < order.size()) {
      const uint32_t node = order[head++];
      nb.clear();
      if (node < M) {
        for (uint32_t e = rp[node]; e < rp[node + 1]; ++e) nb.push_back(M + ev[e]);
      } else {
        const uint32_t v = node - M;
        for (uint32_t k = cp[v]; k < cp[v + 1]; ++k) nb.push_back(edge_check[i2o[k]]);
      }
      std::stable_sort(nb.begin(), nb.end(), [&](uint32_t a, uint32_t b) { return degree(a) < degree(b); });
      for (uint32_t x : nb)
        if (!seen[x]) {
          seen[x] = 1;
          order.push_back(x);
        }
    }
  }
  check_order.clear();
  var_order.clear();
  for (size_t i = order.size(); i-- > 0;) { // reversed
    if (order[i] < M) check_order.push_back(order[i]);
    else var_order.push_back(order[i] - M);
  }
}

// The decoder's graph tables in STORAGE order: check_order[c'] / var_order[v'] = original index kept at storage position
// c' / v'.  Edges keep their order inside a check (check-node sums) and the gather lists keep the original CSC order
// (variable-node sums), so every floating-point sum is evaluated in the reference's order: results are bit-identical
// to the natural order, only the memory addresses change.
struct storage_tables {
  std::vector<uint32_t> row_ptr, edge_var, col_ptr, in_to_out, edge_pos, var_pos, check_pos;
};
storage_tables build_storage_tables(const ldpc_code &code, const std::vector<uint32_t> &check_order,
                                    const std::vector<uint32_t> &var_order) {
  const uint32_t M = uint32_t(code.n_outputs()), N = uint32_t(code.n_inputs()), E = code.n_edges();
  const std::vector<uint32_t> &rp = code.row_ptr(), &ev = code.out_edge_to_var(), &cp = code.col_ptr(), &i2o = code.in_to_out();
  storage_tables t;
  t.var_pos.resize(N);
  t.check_pos.resize(M);
  for (uint32_t i = 0; i < N; ++i) t.var_pos[var_order[i]] = i;
  for (uint32_t i = 0; i < M; ++i) t.check_pos[check_order[i]] = i;
  t.row_ptr.resize(size_t(M) + 1);
  t.edge_var.resize(E);
  t.edge_pos.resize(E);
  uint32_t row = 0;
  for (uint32_t i = 0; i < M; ++i) {
    const uint32_t c = check_order[i];
    t.row_ptr[i] = row;
    for (uint32_t e = rp[c]; e < rp[c + 1]; ++e) {
      t.edge_pos[e] = row;
      t.edge_var[row] = t.var_pos[ev[e]];
      ++row;
    }
  }
  t.row_ptr[M] = row;
  t.col_ptr.resize(size_t(N) + 1);
  t.in_to_out.resize(E);
  uint32_t k2 = 0;
  for (uint32_t i = 0; i < N; ++i) {
    const uint32_t v = var_order[i];
    t.col_ptr[i] = k2;
    for (uint32_t k = cp[v]; k < cp[v + 1]; ++k) t.in_to_out[k2++] = t.edge_pos[i2o[k]];
  }
  t.col_ptr[N] = k2;
  return t;
}

template <typename K> int occupancy_grid(K kernel, int threads, size_t smem, int n_sms, uint64_t items) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const uint64_t need = (items + threads - 1) / threads;
  return int(std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(per_sm) * n_sms, need)));
}

} // namespace

extern "C" {

ldpc_b200_status ldpc_b200_decoder_create(const ldpc_b200_code *code_h, const ldpc_b200_config *cfg,
                                          ldpc_b200_decoder **out) {
  if (!code_h || !cfg || !out) return fail(LDPC_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  const ldpc_code &full_code = *code_h->code;
  if (full_code.n_inputs() & 0x1F) // ldpc_decoder_gpu.cu:30-32
    return fail(LDPC_B200_ERR_MULTIPLE_32, "This decoder only handles input sizes that are multiple of 32");
  if (!full_code.structure_ok()) return fail(LDPC_B200_ERR_CODE, "Incorrect code structure\n"); // :40-58
  // Erased checks (`#ec=` in the alist header, ldpc_code.cpp:65-69): the LAST n_erased_checks checks carry no target
  // syndrome bit -- the reference's harness sizes the per-frame syndromes for the effective checks only
  // (main.cpp:343, n_effective_outputs) while its decoder goes on reading ceil(M/32) words (ldpc_decoder_gpu.cu:104),
  // i.e. it does not support them.  A check of unknown parity carries no information, so the decoder is built on the
  // graph WITHOUT those checks, and p_syndromes holds ceil((M - n_erased_checks)/32) words per frame as the harness
  // allocates them.  Variables left without checks keep their prior.
  std::unique_ptr<ldpc_code> reduced;
  if (full_code.n_erased_outputs() > 0) {
    if (full_code.n_erased_outputs() >= full_code.n_outputs())
      return fail(LDPC_B200_ERR_INVALID, "every check of the code is erased");
    const uint32_t m_eff = uint32_t(full_code.n_outputs() - full_code.n_erased_outputs());
    try {
      reduced.reset(new ldpc_code(m_eff, uint32_t(full_code.n_inputs()), full_code.row_ptr().data(),
                                  full_code.out_edge_to_var().data(), uint32_t(full_code.n_erased_inputs())));
    } catch (const std::exception &e) {
      return fail(LDPC_B200_ERR_CODE, e.what());
    }
  }
  const ldpc_code &code = reduced ? *reduced : full_code;
  if (cfg->precision != LDPC_B200_FP32 && cfg->precision != LDPC_B200_FP16)
    return fail(LDPC_B200_ERR_INVALID, "unknown precision");

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(LDPC_B200_ERR_CUDA, "no CUDA device: the B200 decoder has no CPU fallback");
  }
  std::unique_ptr<ldpc_b200_decoder> d(new ldpc_b200_decoder);
  if (cfg->device >= 0) {
    CU_TRY(cudaSetDevice(cfg->device));
    d->device = cfg->device;
  } else {
    CU_TRY(cudaGetDevice(&d->device));
  }
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, d->device));
  d->n_sms = prop.multiProcessorCount;
  // host worker threads for the staging gather; ranks of one node share the cores (torchrun exports LOCAL_WORLD_SIZE)
  unsigned local_world = 1;
  if (const char *lw = std::getenv("LOCAL_WORLD_SIZE")) local_world = std::max(1, atoi(lw));
  d->host_threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency() / local_world));
  if (const char *ht = std::getenv("LDPC_B200_HOST_THREADS")) d->host_threads = std::max(1, atoi(ht)); // feed experiments
  if (const char *sl = std::getenv("LDPC_B200_STAGE_SLACK_PCT")) d->stage_slack_pct = uint32_t(std::min(100, std::max(0, atoi(sl))));
  d->precision = cfg->precision;
  d->elem = cfg->precision == LDPC_B200_FP32 ? 4 : 2;
  d->M = uint32_t(code.n_outputs());
  d->N = uint32_t(code.n_inputs());
  d->E = code.n_edges();
  d->S = (d->M + 31) >> 5;
  d->W = d->N >> 5;
  d->n_erased = uint32_t(code.n_erased_inputs());
  if (d->n_erased > d->N) return fail(LDPC_B200_ERR_INVALID, "more punctured variables than variables");

  // frames in flight from device memory (ldpc_decoder_gpu.cu:67-99), with this layout's own per-frame cost:
  // messages + priors + staging + bit-sliced syndromes/decisions + packed results
  size_t free_b = 0, total_b = 0;
  CU_TRY(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t graph_bytes = 4ull * (uint64_t(d->M) + 13ull * d->N + 2ull * d->E + 2);
  const uint64_t per_frame = uint64_t(d->elem) * (uint64_t(d->E) + 2ull * d->N) + 2ull * (d->M >> 3) + 8ull * d->S +
                             (d->N >> 3) * 2ull + 64;
  const uint32_t headroom = cfg->mem_headroom_pct ? cfg->mem_headroom_pct : 10;
  const uint64_t usable = uint64_t(double(free_b) * (100.0 - std::min(headroom, 90u)) / 100.0);
  if (usable < graph_bytes + per_frame * 32) return fail(LDPC_B200_ERR_NOMEM, "code does not fit in device memory");
  const uint64_t fit = (usable - graph_bytes) / per_frame;
  uint32_t log2_frames = std::min<uint32_t>(ilog2(fit), std::min<uint32_t>(cfg->max_log2_frames, 13u));
  d->frames_in_flight = 1u << log2_frames;
  d->log2P = std::max<uint32_t>(log2_frames, 5u);
  d->P = 1u << d->log2P;
  d->log2G = d->log2P - (d->precision == LDPC_B200_FP32 ? 2 : 3);
  d->log2PW = d->log2P - 5;

  const size_t P = d->P;
  CU_TRY(d->d_row_ptr.alloc(d->M + 1));
  CU_TRY(d->d_col_ptr.alloc(d->N + 1));
  CU_TRY(d->d_in_to_out.alloc(d->E));
  CU_TRY(d->d_edge_var.alloc(d->E));
  // graph tables, in storage order (identity unless cfg->storage_order asks for the locality order)
  storage_tables tabs;
  d->reordered = cfg->storage_order == LDPC_B200_ORDER_LOCALITY;
  if (cfg->storage_order != LDPC_B200_ORDER_NATURAL && cfg->storage_order != LDPC_B200_ORDER_LOCALITY)
    return fail(LDPC_B200_ERR_INVALID, "unknown storage_order");
  if (d->reordered) {
    std::vector<uint32_t> check_order, var_order;
    locality_order(code, check_order, var_order);
    tabs = build_storage_tables(code, check_order, var_order);
    d->h_edge_pos = tabs.edge_pos;
    d->h_var_pos = tabs.var_pos;
    CU_TRY(d->d_var_pos.alloc(d->N));
    CU_TRY(d->d_check_pos.alloc(d->M));
    CU_TRY(d->d_row_ptr_orig.alloc(d->M + 1));
    CU_TRY(d->d_edge_var_orig.alloc(d->E));
    CU_TRY(cudaMemcpy(d->d_var_pos.p, tabs.var_pos.data(), 4ull * d->N, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d->d_check_pos.p, tabs.check_pos.data(), 4ull * d->M, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d->d_row_ptr_orig.p, code.row_ptr().data(), 4ull * (d->M + 1), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(d->d_edge_var_orig.p, code.out_edge_to_var().data(), 4ull * d->E, cudaMemcpyHostToDevice));
  }
  const std::vector<uint32_t> &t_row_ptr = d->reordered ? tabs.row_ptr : code.row_ptr();
  const std::vector<uint32_t> &t_col_ptr = d->reordered ? tabs.col_ptr : code.col_ptr();
  const std::vector<uint32_t> &t_in_to_out = d->reordered ? tabs.in_to_out : code.in_to_out();
  const std::vector<uint32_t> &t_edge_var = d->reordered ? tabs.edge_var : code.out_edge_to_var();
  CU_TRY(cudaMemcpy(d->d_row_ptr.p, t_row_ptr.data(), 4ull * (d->M + 1), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(d->d_col_ptr.p, t_col_ptr.data(), 4ull * (d->N + 1), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(d->d_in_to_out.p, t_in_to_out.data(), 4ull * d->E, cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(d->d_edge_var.p, t_edge_var.data(), 4ull * d->E, cudaMemcpyHostToDevice));
  CU_TRY(d->d_msg.alloc(size_t(d->E) * P * d->elem));
  CU_TRY(d->d_llr0.alloc(size_t(d->N) * P * d->elem));
  CU_TRY(d->d_syn.alloc(size_t(d->M) << d->log2PW));
  CU_TRY(d->d_hard.alloc(size_t(d->N) << d->log2PW));
  CU_TRY(d->d_viol.alloc(P / 32));
  CU_TRY(d->d_retire_mask.alloc(P / 32));
  CU_TRY(d->d_new_mask.alloc(P / 32));
  CU_TRY(d->d_slot_entry.alloc(P));
  CU_TRY(d->d_refill.alloc(P));
  CU_TRY(d->d_extract.alloc(P));
  CU_TRY(d->d_phi_table.alloc(dev::kPhiTableSize));
  CU_TRY(d->d_state.alloc(1));
  CU_TRY(d->d_summary.alloc(1));
  CU_TRY(d->d_n_new.alloc(1));
  CU_TRY(d->d_in_gpu.alloc(P));
  CU_TRY(d->d_virt.alloc(P));
  CU_TRY(d->d_retired.alloc(P));
  CU_TRY(d->d_scratch.alloc(2 * P));
  CU_TRY(d->d_n_moves.alloc(1));
  CU_TRY(d->d_moves.alloc(P));
  CU_TRY(cudaMemset(d->d_n_moves.p, 0, sizeof(uint32_t)));
  CU_TRY(cudaMemset(d->d_state.p, 0, sizeof(dev::SchedState)));
  CU_TRY(d->h_summary.alloc(2));
  CU_TRY(d->h_retired.alloc(2 * P));
  CU_TRY(cudaMemset(d->d_msg.p, 0, size_t(d->E) * P * d->elem));
  CU_TRY(cudaMemset(d->d_llr0.p, 0, size_t(d->N) * P * d->elem));
  CU_TRY(cudaMemset(d->d_syn.p, 0, (size_t(d->M) << d->log2PW) * 4));
  CU_TRY(cudaMemset(d->d_hard.p, 0, (size_t(d->N) << d->log2PW) * 4));
  CU_TRY(d->h_viol.alloc(P / 32));
  CU_TRY(d->h_mask.alloc(P / 32));
  CU_TRY(d->h_retire_mask.alloc(P / 32));
  CU_TRY(d->h_slot_entry.alloc(P));
  CU_TRY(d->h_refill.alloc(P));
  CU_TRY(d->h_extract.alloc(P));
  CU_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&d->ev_staged, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&d->ev_refilled, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    CU_TRY(cudaEventCreateWithFlags(&d->ev_period[i], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&d->ev_extracted[i], cudaEventDisableTiming));
  }
  CU_TRY(cudaEventCreate(&d->ev_a));
  CU_TRY(cudaEventCreate(&d->ev_b));

  dev::build_phi_table_kernel<<<dev::kPhiTableSize / 256, 256, 0, d->stream>>>(d->d_phi_table.p);
  CU_TRY(cudaGetLastError());

  // only the non-zero head of the fp16 table is staged in shared memory (2e^-x underflows above x ~ 17.3)
  if (d->precision == LDPC_B200_FP16) {
    std::vector<uint16_t> tab(dev::kPhiTableSize);
    CU_TRY(cudaStreamSynchronize(d->stream));
    CU_TRY(cudaMemcpy(tab.data(), d->d_phi_table.p, tab.size() * 2, cudaMemcpyDeviceToHost));
    uint32_t last = 0x7C00; // +inf: phi = 0; NaN patterns never reach the lookup (hmax2 removes them)
    while (last > 0 && tab[last - 1] == 0) --last;
    if (tab[last] != 0) last = dev::kPhiTableSize - 1;
    d->tab_last = last;
  }
  // persistent grids: SMs x resident CTAs
  ldpc_b200_status st_sel = LDPC_B200_OK;
  d->smem_tab = d->precision == LDPC_B200_FP16 ? size_t(d->tab_last / 8 + 1) * 16 : 0;
  const uint64_t cn_items = uint64_t(d->M) << d->log2G, vn_items = uint64_t(d->N) << d->log2G;
  d->deg_cn = pick_reg_degree(code.row_ptr());
  d->deg_vn = pick_reg_degree(code.col_ptr());
  d->cn_parts = pick_cn_parts(code.row_ptr(), code.max_degree_out());
  if (const char *e = std::getenv("LDPC_B200_CN_BULK"))
    d->cn_bulk = atoi(e) != 0 && d->cn_parts == 1 && code.max_degree_out() <= d->deg_cn &&
                 (1u << d->log2G) >= uint32_t(dev::kBulkTileVec);
  if (const char *e = std::getenv("LDPC_B200_VN_BULK"))
    d->vn_bulk = atoi(e) != 0 && code.max_degree_in() <= d->deg_vn && (1u << d->log2G) >= uint32_t(dev::kBulkTileVec);
  if (const char *e = std::getenv("LDPC_B200_DEG_VN")) d->deg_vn = atoi(e); // kernel experiments only
  if (const char *e = std::getenv("LDPC_B200_DEG_CN")) d->deg_cn = atoi(e);
  if (d->cn_parts > 1 && d->log2G < 5) { // a warp must stay inside one row: at least 32/G parts
    d->cn_parts = std::max(d->cn_parts, 32 >> d->log2G);
    if (d->cn_parts > 8) d->cn_parts = 1;
  }
  st_sel = select_kernels(d.get());
  if (st_sel != LDPC_B200_OK) return st_sel;
  { // gather list of the variable-node pass: [degree, rows of the first D edges (CSC order), ..., col_ptr[v]] per variable
    const int D = d->deg_vn == 4 ? 4 : (d->deg_vn == 6 ? 6 : 8), W = D <= 6 ? 8 : 12; // as select_for_precision
    const std::vector<uint32_t> &cp = t_col_ptr, &i2o = t_in_to_out; // storage order
    std::vector<uint32_t> ell(size_t(d->N) * W, 0u);
    for (uint32_t v = 0; v < d->N; ++v) {
      uint32_t *e = &ell[size_t(v) * W];
      const uint32_t deg = cp[v + 1] - cp[v];
      e[0] = deg;
      for (uint32_t k = 0; k < std::min<uint32_t>(deg, uint32_t(D)); ++k) e[1 + k] = i2o[cp[v] + k];
      e[W - 1] = cp[v];
    }
    CU_TRY(d->d_vn_ell.alloc(ell.size()));
    CU_TRY(cudaMemcpy(d->d_vn_ell.p, ell.data(), ell.size() * 4, cudaMemcpyHostToDevice));
  }
  d->threads_cn = d->cn_parts > 1 ? dev::kThreads : dev::cn_threads_for(d->deg_cn);
  d->smem_cn = d->smem_tab;
  if (d->cn_bulk) {
    d->threads_cn = dev::kBulkTileVec;
    d->smem_cn = d->smem_tab + size_t(dev::kBulkStages) * d->deg_cn * dev::kBulkTileVec * 16 + 64;
    CU_TRY(cudaFuncSetAttribute(d->k_cn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_cn)));
  }
  d->grid_cn = occupancy_grid(d->k_cn, d->threads_cn, d->smem_cn, d->n_sms, cn_items);
  if (d->k_cn_refill) d->grid_cn_refill = occupancy_grid(d->k_cn_refill, dev::kThreads, d->smem_tab, d->n_sms, cn_items);
  d->threads_vn = dev::vn_threads_for(d->deg_vn);
  d->smem_vn = d->smem_tab;
  if (d->vn_bulk) {
    const int D = d->deg_vn == 4 ? 4 : (d->deg_vn == 6 ? 6 : 8), W = D <= 6 ? 8 : 12;
    d->threads_vn = dev::kVnBulkThreads;
    d->smem_vn = d->smem_tab + size_t(2) * dev::kVnBulkStages * (D + 1) * dev::kBulkTileVec * 16 +
                 size_t(2) * dev::kVnBulkStages * W * 4 + size_t(2) * dev::kVnBulkStages * 8 + 64;
    CU_TRY(cudaFuncSetAttribute(d->k_vn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_vn)));
    CU_TRY(cudaFuncSetAttribute(d->k_vnb, cudaFuncAttributeMaxDynamicSharedMemorySize, int(d->smem_vn)));
  }
  d->grid_vn = occupancy_grid(d->k_vn, d->threads_vn, d->smem_vn, d->n_sms, vn_items);
  d->grid_vnb = occupancy_grid(d->k_vnb, d->threads_vn, d->smem_vn, d->n_sms, vn_items);
  d->grid_misc = d->n_sms * 8; // x 256 threads: a multiple of every row length G <= 2048 (refill_kernel keeps its g)
  CU_TRY(cudaStreamSynchronize(d->stream));
  *out = d.release();
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_code_locality_order(const ldpc_b200_code *code, uint32_t *check_order, uint32_t *var_order) {
  if (!code || !check_order || !var_order) return fail(LDPC_B200_ERR_INVALID, "null argument");
  std::vector<uint32_t> co, vo;
  locality_order(*code->code, co, vo);
  std::copy(co.begin(), co.end(), check_order);
  std::copy(vo.begin(), vo.end(), var_order);
  return LDPC_B200_OK;
}

void ldpc_b200_decoder_destroy(ldpc_b200_decoder *dec) {
  if (!dec) return;
  cudaSetDevice(dec->device);
  delete dec;
}

uint32_t ldpc_b200_decoder_frames_in_flight(const ldpc_b200_decoder *dec) { return dec ? dec->frames_in_flight : 0; }

ldpc_b200_status ldpc_b200_decoder_prepare_host_io(ldpc_b200_decoder *dec) {
  if (!dec) return fail(LDPC_B200_ERR_INVALID, "null decoder");
  CU_TRY(cudaSetDevice(dec->device));
  const uint32_t ring = dec->P + dec->stage_slack(dec->frames_in_flight, 32) + 32;
  ldpc_b200_status st = dec->ensure_input_staging(ring);
  if (st == LDPC_B200_OK) st = dec->ensure_syndrome_staging(ring);
  if (st == LDPC_B200_OK) st = dec->ensure_result_staging();
  return st;
}

void ldpc_b200_decoder_set_erased_vars(ldpc_b200_decoder *dec, uint32_t n_erased) {
  if (dec && n_erased <= dec->N) dec->n_erased = n_erased;
}

uint64_t ldpc_b200_bytes_per_frame_iteration(const ldpc_b200_decoder *dec) {
  if (!dec) return 0;
  return (4ull * dec->E + dec->N) * dec->elem + 4ull * dec->S;
}

} // extern "C"

// ------------------------------------------------------------------ staging of new frames
namespace {

struct input_view {
  const uint8_t *base;     // host or device
  const uint32_t *syn;     // host or device
  bool on_device, syn_on_device;
  int layout;
  uint64_t frame_stride;   // FRAME_MAJOR
  uint32_t n_cols;         // frames physically present (pool or n_frames)
};

// Fill h_refill[t] for new frames [first, first+cnt) going to slots[t]; stage host data if needed.
// Returns device pointers the refill kernels should read.
ldpc_b200_status stage_frames(ldpc_b200_decoder *d, const input_view &in, uint32_t first, uint32_t cnt,
                              const uint32_t *slots, const void **src_dev, const uint32_t **syn_dev,
                              ldpc_b200_report *rep) {
  const uint32_t n_reg = d->N - d->n_erased;
  const size_t es = d->elem;
  if (!in.on_device) {
    const ldpc_b200_status se = d->ensure_input_staging();
    if (se != LDPC_B200_OK) return se;
  }
  if (!in.syn_on_device) {
    const ldpc_b200_status se = d->ensure_syndrome_staging();
    if (se != LDPC_B200_OK) return se;
  }
  for (uint32_t t = 0; t < cnt; ++t) {
    const uint32_t col = (first + t) % in.n_cols;
    dev::RefillEntry &en = d->h_refill.p[t];
    en.slot = slots[t];
    en.pad = 0;
    if (in.on_device) {
      if (in.layout == LDPC_B200_LAYOUT_INTERLEAVED) {
        en.src_offset = col;
        en.src_stride = in.n_cols;
      } else {
        en.src_offset = uint64_t(col) * in.frame_stride;
        en.src_stride = 1;
      }
    } else { // staged as [n_reg][cnt] (interleaved) or [cnt][n_reg] (frame-major)
      if (in.layout == LDPC_B200_LAYOUT_INTERLEAVED) {
        en.src_offset = t;
        en.src_stride = cnt;
      } else {
        en.src_offset = uint64_t(t) * n_reg;
        en.src_stride = 1;
      }
    }
    en.syn_offset = in.syn_on_device ? uint64_t(col) * d->S : uint64_t(t) * d->S;
  }
  if (in.on_device) {
    *src_dev = in.base;
  } else {
    uint8_t *stage = d->h_stage.p;
    if (in.layout == LDPC_B200_LAYOUT_INTERLEAVED) {
      // prepare_vectors (ldpc_decoder_gpu.cu:199-207): rows of cnt consecutive frames, split where the pool wraps
      const uint32_t c0 = first % in.n_cols;
      parallel_for(n_reg, d->host_threads, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
          uint8_t *dst = stage + i * cnt * es;
          const uint8_t *row = in.base + i * size_t(in.n_cols) * es;
          uint32_t done = 0, col = c0;
          while (done < cnt) {
            const uint32_t run = std::min(cnt - done, in.n_cols - col);
            memcpy(dst + size_t(done) * es, row + size_t(col) * es, size_t(run) * es);
            done += run;
            col = 0;
          }
        }
      });
    } else {
      parallel_for(cnt, d->host_threads, [&](size_t a, size_t b) {
        for (size_t t = a; t < b; ++t) {
          const uint32_t col = uint32_t((first + t) % in.n_cols);
          memcpy(stage + t * size_t(n_reg) * es, in.base + size_t(col) * in.frame_stride * es, size_t(n_reg) * es);
        }
      });
    }
    const size_t bytes = size_t(n_reg) * cnt * es;
    CU_TRY(cudaMemcpyAsync(d->d_stage.p, stage, bytes, cudaMemcpyHostToDevice, d->stream));
    if (rep) rep->h2d_bytes += bytes;
    *src_dev = d->d_stage.p;
  }
  if (in.syn_on_device) {
    *syn_dev = in.syn;
  } else {
    for (uint32_t t = 0; t < cnt; ++t) {
      const uint32_t col = (first + t) % in.n_cols;
      memcpy(d->h_stage_syn.p + size_t(t) * d->S, in.syn + size_t(col) * d->S, size_t(d->S) * 4);
    }
    CU_TRY(cudaMemcpyAsync(d->d_stage_syn.p, d->h_stage_syn.p, size_t(d->S) * cnt * 4, cudaMemcpyHostToDevice,
                           d->stream));
    if (rep) rep->h2d_bytes += size_t(d->S) * cnt * 4;
    *syn_dev = d->d_stage_syn.p;
  }
  return LDPC_B200_OK;
}

// Retire the frames sitting in slots[0..cnt) -> results rows frames[0..cnt)
ldpc_b200_status retire_frames(ldpc_b200_decoder *d, uint32_t cnt, const uint32_t *slots, const uint32_t *frames,
                               uint32_t *results, bool results_on_device, ldpc_b200_report *rep) {
  const uint32_t PW = 1u << d->log2PW;
  if (!results_on_device) {
    const ldpc_b200_status se = d->ensure_result_staging();
    if (se != LDPC_B200_OK) return se;
  }
  std::fill(d->h_retire_mask.p, d->h_retire_mask.p + PW, 0u);
  for (uint32_t t = 0; t < cnt; ++t) {
    const uint32_t s = slots[t];
    d->h_retire_mask.p[s >> 5] |= 1u << (s & 31);
    d->h_extract.p[s].dst_offset = results_on_device ? uint64_t(frames[t]) * d->W : uint64_t(t) * d->W;
  }
  d->launch_extract(results_on_device ? results : d->d_results.p);
  CU_TRY(cudaGetLastError());
  if (!results_on_device) {
    const size_t bytes = size_t(d->W) * cnt * 4;
    CU_TRY(cudaMemcpyAsync(d->h_results.p, d->d_results.p, bytes, cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(cudaStreamSynchronize(d->stream));
    for (uint32_t t = 0; t < cnt; ++t)
      memcpy(results + size_t(frames[t]) * d->W, d->h_results.p + size_t(t) * d->W, size_t(d->W) * 4);
    if (rep) rep->d2h_bytes += bytes;
  }
  return LDPC_B200_OK;
}

} // namespace

// ------------------------------------------------------------------ per-launch timing
// LDPC_B200_FLAG_TIME_KERNELS: CUDA events on the launching stream around every check-node / variable-node launch and
// around the per-check service steps.  Two pools (periods alternate: the device runs one period ahead of the host), 3
// events per flood iteration + 4 per service step; flush(h) adds the elapsed times of pool h to the report once the
// host has synchronised with that period.  Without the flag every call is a no-op.
namespace {

struct launch_timer {
  std::vector<cudaEvent_t> iter_ev[2], service_ev[2];
  uint32_t timed_iters[2] = {0, 0};
  cudaStream_t stream = nullptr;
  ldpc_b200_report *rep = nullptr;
  bool on = false;

  ldpc_b200_status create(uint32_t max_iters, cudaStream_t s, ldpc_b200_report *r) {
    stream = s;
    rep = r;
    for (int h = 0; h < 2; ++h) {
      iter_ev[h].assign(3 * size_t(max_iters), nullptr);
      service_ev[h].assign(4, nullptr);
      for (cudaEvent_t &e : iter_ev[h]) CU_TRY(cudaEventCreate(&e));
      for (cudaEvent_t &e : service_ev[h]) CU_TRY(cudaEventCreate(&e));
    }
    on = true;
    return LDPC_B200_OK;
  }
  ~launch_timer() {
    for (int h = 0; h < 2; ++h) {
      for (cudaEvent_t e : iter_ev[h])
        if (e) cudaEventDestroy(e);
      for (cudaEvent_t e : service_ev[h])
        if (e) cudaEventDestroy(e);
    }
  }
  // which: 0 before the check-node launch, 1 between the two launches, 2 after the variable-node launch
  void mark_iter(int h, uint32_t i, int which) {
    if (on) cudaEventRecord(iter_ev[h][3 * i + which], stream);
  }
  // which: 0 before the syndrome check, 1 after it, 2 after scheduler + extraction, 3 after the refill
  void mark_service(int h, int which) {
    if (on) cudaEventRecord(service_ev[h][which], stream);
  }
  void set_iters(int h, uint32_t n) {
    if (on) timed_iters[h] = n;
  }
  // first_cn_is_fused: the first check-node pass of these iterations was the fused one (timed apart)
  void flush(int h, bool first_cn_is_fused) {
    if (!on) return;
    for (uint32_t i = 0; i < timed_iters[h]; ++i) {
      float cn = 0, vn = 0;
      cudaEventElapsedTime(&cn, iter_ev[h][3 * i], iter_ev[h][3 * i + 1]);
      cudaEventElapsedTime(&vn, iter_ev[h][3 * i + 1], iter_ev[h][3 * i + 2]);
      if (i == 0 && first_cn_is_fused) rep->cn_refill_s += cn * 1e-3; // the fused pass is timed apart from the plain ones
      else rep->cn_s += cn * 1e-3;
      rep->vn_s += vn * 1e-3;
    }
    timed_iters[h] = 0;
    float check = 0, retire = 0, refill = 0;
    cudaEventElapsedTime(&check, service_ev[h][0], service_ev[h][1]);
    cudaEventElapsedTime(&retire, service_ev[h][1], service_ev[h][2]);
    cudaEventElapsedTime(&refill, service_ev[h][2], service_ev[h][3]);
    rep->check_s += check * 1e-3;
    rep->retire_s += retire * 1e-3;
    rep->refill_s += refill * 1e-3;
  }
};

} // namespace

// ------------------------------------------------------------------ scheduler constants
// Everything dev::sched_kernel needs to know about one decode() call: the reference's batch / stop-rule parameters
// (ldpc_decoder_gpu.cu:296-315) plus where new frames come from (input layout, staging-ring geometry).
namespace {

dev::SchedConst make_sched_const(const ldpc_b200_decoder *d, const ldpc_b200_decode_params *prm, uint32_t n_frames,
                                 bool in_dev, bool syn_dev, bool res_dev, uint32_t slack_frames) {
  dev::SchedConst sc;
  memset(&sc, 0, sizeof(sc));
  const uint32_t B = std::min(n_frames, d->frames_in_flight); // num_vectors_gpu_batch
  const uint32_t n_reg = d->N - d->n_erased;
  sc.n_frames = n_frames;
  sc.batch = B;
  sc.log2_p = d->log2P;
  sc.k_check = prm->num_iter_check_parity;
  sc.imax = prm->num_iter_max;
  sc.simd_pairs = (prm->flags & LDPC_B200_FLAG_SIMD_PAIRS) ? 1u : 0u;
  sc.syn_words = d->S;
  sc.res_words = d->W;
  sc.n_cols = prm->input_pool_frames ? prm->input_pool_frames : n_frames;
  sc.n_regular = n_reg;
  const bool packed = prm->input_layout == LDPC_B200_LAYOUT_PACKED_BITS;
  const bool frame_major = prm->input_layout == LDPC_B200_LAYOUT_FRAME_MAJOR || packed;
  sc.row_elems = packed ? (n_reg + 31) / 32 : n_reg; // per-frame elements of a frame-major input
  sc.frame_stride = prm->frame_stride ? prm->frame_stride : sc.row_elems;
  sc.ring_block = frame_major ? 1u : 32u; // frames per interleaved staging block (64-128 B runs)
  // live staged frames are always inside [next_to_load, next_to_load + B) (+ block rounding)
  // (with no slack this is min(P, B rounded up to blocks + 1 block), as before)
  sc.ring_frames = slack_frames ? ((B + slack_frames + sc.ring_block - 1) / sc.ring_block + 1) * sc.ring_block
                                : std::min<uint32_t>(d->P, ((B + sc.ring_block - 1) / sc.ring_block + 1) * sc.ring_block);
  sc.input_mode = in_dev ? (frame_major ? 1u : 0u) : (frame_major ? 3u : 2u);
  sc.early_retire = (prm->flags & LDPC_B200_FLAG_EARLY_RETIRE) ? 1u : 0u;
  sc.log2_v = d->precision == LDPC_B200_FP32 ? 2u : 3u;
  // the flood kernels need at least 32 slots, and a warp of the split-row kernel must stay inside one row
  sc.min_log2_p_active = 5;
  if (d->cn_parts > 1) sc.min_log2_p_active = std::max<uint32_t>(5, ilog2(32u / uint32_t(d->cn_parts)) + sc.log2_v);
  if (d->cn_bulk || d->vn_bulk) sc.min_log2_p_active = ilog2(uint32_t(dev::kBulkTileVec)) + sc.log2_v; // whole tiles only
  sc.syn_ring = syn_dev ? 0u : 1u;
  sc.results_ring = res_dev ? 0u : 1u;
  return sc;
}

} // namespace

// ------------------------------------------------------------------ shared frame queue
// "a single host-side atomic cursor over the frame queue gives dynamic load balance" (the reference
// hard-codes device 0, h/cuda_manager.h:52).  A queue is an atomic cursor over the rows of the caller's arrays; every
// decoder drawing from it (one per GPU: threads of one process, or one process per GPU through POSIX shared memory)
// reserves whole blocks of consecutive frames and decodes them as ONE continuous call -- frames keep being loaded into
// freed slots across block boundaries, and only the end of the queue has a tail.
struct queue_shared { // lives in the process heap or in a shared-memory mapping
  std::atomic<uint64_t> cursor;
  uint32_t n_total, block, magic, pad;
};
constexpr uint32_t kQueueMagic = 0x4C445043u;

struct ldpc_b200_queue {
  queue_shared *sh = nullptr;
  queue_shared local;
  size_t map_len = 0;
  std::string shm_name;
  bool owner = false;
  ~ldpc_b200_queue() {
    if (map_len) {
      munmap(sh, map_len);
      if (owner) shm_unlink(shm_name.c_str());
    }
  }
};

namespace {

// one decoder's view of a queue: the blocks it has reserved so far, mirrored into mapped host memory for the
// scheduler kernel (dev::FrameQueueView)
struct local_queue {
  ldpc_b200_queue *q = nullptr;
  uint32_t block = 0, n_total = 0;
  uint32_t avail = 0; // local frames reserved so far
  bool final = false; // the cursor has passed the end: nothing more will come
  std::vector<uint32_t> starts;
  uint32_t *h_blk = nullptr;          // mapped pinned [max_blocks], 0xFFFFFFFF = not written
  volatile uint32_t *h_ctl = nullptr; // mapped pinned [2]: avail, final
  uint32_t max_blocks = 0;

  ldpc_b200_status init(ldpc_b200_queue *queue) {
    q = queue;
    block = queue->sh->block;
    n_total = queue->sh->n_total;
    max_blocks = (n_total + block - 1) / block + 1;
    CU_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h_blk), sizeof(uint32_t) * (size_t(max_blocks) + 2),
                         cudaHostAllocPortable | cudaHostAllocMapped));
    std::fill(h_blk, h_blk + max_blocks + 2, 0xFFFFFFFFu);
    h_ctl = h_blk + max_blocks; // the two control words share the allocation
    h_ctl[0] = 0;
    h_ctl[1] = 0;
    return LDPC_B200_OK;
  }
  ~local_queue() {
    if (h_blk) cudaFreeHost(h_blk);
  }
  // reserve blocks until `want` local frames are available (or the queue is exhausted)
  void reserve_upto(uint64_t want) {
    while (!final && avail < want) {
      const uint64_t start = q->sh->cursor.fetch_add(block, std::memory_order_relaxed);
      if (start >= n_total) {
        final = true;
        break;
      }
      const uint32_t cnt = uint32_t(std::min<uint64_t>(block, n_total - start));
      h_blk[starts.size()] = uint32_t(start);
      starts.push_back(uint32_t(start));
      std::atomic_thread_fence(std::memory_order_release); // the block start is visible before the count that covers it
      avail += cnt;
      h_ctl[0] = avail;
      if (cnt < block) final = true; // a partial block is the last one of the queue
    }
    if (final) {
      std::atomic_thread_fence(std::memory_order_release);
      h_ctl[1] = 1;
    }
  }
  uint32_t global_of(uint32_t f) const { return starts[f / block] + f % block; }
  dev::FrameQueueView view() const { return dev::FrameQueueView{h_ctl, h_blk}; }
};

} // namespace

// ------------------------------------------------------------------ host feed
// Staging of host-resident frames into the device ring over the copy stream (the reference's prepare_vectors +
// transfer_vectors, ldpc_decoder_gpu.cu:199-273, but gathered by all host threads, off the critical path and in
// sub-ranges so that the H2D of one chunk overlaps the gather of the next).  Frames [staged, upto) are gathered into a
// pinned mirror laid out like the ring and copied block by block; ring position of frame f: f % ring_frames.
namespace {

struct host_feeder {
  ldpc_b200_decoder *d = nullptr;
  ldpc_b200_report *rep = nullptr;
  const uint8_t *in_host = nullptr;   // null: the channel values are device-resident (read in place)
  const uint32_t *syn_host = nullptr; // null: the syndromes are device-resident
  bool frame_major = false;
  uint32_t n_frames = 0, n_cols = 0, n_reg = 0, ring_block = 32, ring_frames = 0, row_elems = 0;
  size_t es = 4, es_fm = 4;  // element size of the interleaved / frame-major (or packed-bit) layouts
  uint64_t frame_stride = 0; // frame-major inputs: elements between frames
  uint32_t staged = 0;       // frames [0, staged) have been handed to the copy stream
  bool refill_recorded = false;
  const local_queue *lq = nullptr; // frames are drawn from a shared queue: local index -> row of the caller's arrays

  uint32_t col_of(uint32_t f) const { return (lq ? lq->global_of(f) : f) % n_cols; }
  // consecutive local frames are consecutive rows of the caller's arrays only inside a queue block
  uint32_t contiguous_from(uint32_t f) const { return lq ? lq->block - f % lq->block : 0xFFFFFFFFu; }

  ldpc_b200_status stage_upto(uint32_t upto) {
    if (!in_host && !syn_host) return LDPC_B200_OK;
    upto = std::min(lq ? lq->avail : n_frames, upto);
    if (upto <= staged) return LDPC_B200_OK;
    CU_TRY(cudaStreamSynchronize(d->copy_stream)); // earlier copies out of the pinned mirror are done
    // ring positions being overwritten belong to frames already loaded; their refill kernels must have read them
    if (refill_recorded) CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_refilled, 0));
    const uint32_t sub = std::max(8 * ring_block, 256u);
    for (uint32_t f0 = staged; f0 < upto;) {
      const uint32_t f1 = std::min(upto, (f0 / sub + 1) * sub);
      ldpc_b200_status st = LDPC_B200_OK;
      if (in_host) st = frame_major ? stage_frame_major(f0, f1) : stage_interleaved(f0, f1);
      if (st == LDPC_B200_OK && syn_host) st = stage_syndromes(f0, f1);
      if (st != LDPC_B200_OK) return st;
      f0 = f1;
    }
    staged = upto;
    CU_TRY(cudaEventRecord(d->ev_staged, d->copy_stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_staged, 0)); // the refill kernels read the ring
    return LDPC_B200_OK;
  }

private:
  // the reference's prepare_vectors (ldpc_decoder_gpu.cu:199-207) in blocks of ring_block frames:
  // pinned[blk][i][0..ring_block) <- input[i][col..col+run)
  ldpc_b200_status stage_interleaved(uint32_t f0, uint32_t f1) {
    struct chunk_t {
      size_t dst, src, bytes; // dst: offset of row 0 inside the pinned mirror; src: byte offset inside a row
    };
    std::vector<chunk_t> chunks; // the same for every variable row: computed once
    for (uint32_t f = f0; f < f1;) {
      const uint32_t r = f % ring_frames, in_blk = r % ring_block, col = col_of(f);
      const uint32_t run = std::min({f1 - f, ring_block - in_blk, n_cols - col, contiguous_from(f)});
      chunks.push_back({(size_t(r / ring_block) * n_reg * ring_block + in_blk) * es, size_t(col) * es, size_t(run) * es});
      f += run;
    }
    const size_t row_dst_stride = size_t(ring_block) * es, row_src_stride = size_t(n_cols) * es;
    uint8_t *mirror = d->h_stage.p;
    const uint8_t *src = in_host;
    // Rows of the caller's array are row_src_stride apart (8 KB for 4096 fp16 frames): every row is a new page and a
    // new DRAM stream that the hardware prefetcher only picks up after a few lines, so the first lines of the rows a
    // few iterations ahead are requested by hand.
    const size_t first_src = chunks.empty() ? 0 : chunks.front().src;
    const size_t run_bytes = chunks.empty() ? 0 : chunks.back().src + chunks.back().bytes - first_src;
    constexpr size_t kAhead = 8;
    parallel_for(n_reg, d->host_threads, [&](size_t ia, size_t ib) {
      for (size_t i = ia; i < ib; ++i) {
        const uint8_t *row = src + i * row_src_stride;
        uint8_t *dst_row = mirror + i * row_dst_stride;
#if defined(__SSE2__)
        if (i + kAhead < ib) {
          const char *pf = reinterpret_cast<const char *>(src + (i + kAhead) * row_src_stride + first_src);
          _mm_prefetch(pf, _MM_HINT_T0);
          if (run_bytes > 64) _mm_prefetch(pf + 64, _MM_HINT_T0);
          if (run_bytes > 128) _mm_prefetch(pf + 128, _MM_HINT_T0);
        }
#endif
        for (const chunk_t &c : chunks) {
          // whole blocks (32 frames = 64 or 128 bytes) start on a 64-byte boundary of the mirror
          if (c.bytes == 64 && c.dst % 64 == 0) {
            copy64_streaming(dst_row + c.dst, row + c.src);
          } else if (c.bytes == 128 && c.dst % 128 == 0) {
            copy64_streaming(dst_row + c.dst, row + c.src);
            copy64_streaming(dst_row + c.dst + 64, row + c.src + 64);
          } else {
            memcpy(dst_row + c.dst, row + c.src, c.bytes);
          }
        }
      }
      streaming_fence(); // the streaming stores are globally visible before the DMA engine reads the mirror
    });
    // one H2D per touched block range (contiguous blocks in the ring)
    for (uint32_t f = f0; f < f1;) {
      const uint32_t r = f % ring_frames, blk0 = r / ring_block;
      const uint32_t run = std::min(f1 - f, ring_frames - r); // up to the end of the ring
      const uint32_t n_blk = (r % ring_block + run + ring_block - 1) / ring_block;
      const size_t off = size_t(blk0) * n_reg * ring_block * es, bytes = size_t(n_blk) * n_reg * ring_block * es;
      CU_TRY(cudaMemcpyAsync(d->d_stage.p + off, d->h_stage.p + off, bytes, cudaMemcpyHostToDevice, d->copy_stream));
      rep->h2d_bytes += bytes;
      f += run;
    }
    return LDPC_B200_OK;
  }

  // frame-major (or packed-bit) input: one contiguous row per frame
  ldpc_b200_status stage_frame_major(uint32_t f0, uint32_t f1) {
    uint8_t *mirror = d->h_stage.p;
    parallel_for(f1 - f0, d->host_threads, [&](size_t a, size_t b) {
      for (size_t t = a; t < b; ++t) {
        const uint32_t f = f0 + uint32_t(t);
        memcpy(mirror + size_t(f % ring_frames) * row_elems * es_fm, in_host + size_t(col_of(f)) * frame_stride * es_fm,
               size_t(row_elems) * es_fm);
      }
    });
    for (uint32_t f = f0; f < f1;) {
      const uint32_t r = f % ring_frames, run = std::min(f1 - f, ring_frames - r);
      const size_t off = size_t(r) * row_elems * es_fm, bytes = size_t(run) * row_elems * es_fm;
      CU_TRY(cudaMemcpyAsync(d->d_stage.p + off, d->h_stage.p + off, bytes, cudaMemcpyHostToDevice, d->copy_stream));
      rep->h2d_bytes += bytes;
      f += run;
    }
    return LDPC_B200_OK;
  }

  ldpc_b200_status stage_syndromes(uint32_t f0, uint32_t f1) {
    const uint32_t S = d->S;
    for (uint32_t f = f0; f < f1; ++f)
      memcpy(d->h_stage_syn.p + size_t(f % ring_frames) * S, syn_host + size_t(col_of(f)) * S, size_t(S) * 4);
    for (uint32_t f = f0; f < f1;) {
      const uint32_t r = f % ring_frames, run = std::min(f1 - f, ring_frames - r);
      CU_TRY(cudaMemcpyAsync(d->d_stage_syn.p + size_t(r) * S, d->h_stage_syn.p + size_t(r) * S, size_t(run) * S * 4,
                             cudaMemcpyHostToDevice, d->copy_stream));
      rep->h2d_bytes += size_t(run) * S * 4;
      f += run;
    }
    return LDPC_B200_OK;
  }
};

} // namespace

// queue == nullptr: the classic call, frames 0..n_frames-1.  Otherwise n_frames is the size of the queue (= rows of the
// caller's arrays) and this decoder decodes whatever blocks of it it manages to reserve.
static ldpc_b200_status decode_impl(ldpc_b200_decoder *d, const ldpc_b200_decode_params *prm, uint32_t n_frames,
                                    const void *input, const uint32_t *syndromes, uint32_t *results,
                                    ldpc_b200_report *report, uint32_t *iter_counts, ldpc_b200_queue *queue) {
  if (!d || !prm) return fail(LDPC_B200_ERR_INVALID, "null argument");
  if (n_frames == 0) return LDPC_B200_OK; // ldpc_decoder_gpu.cu:293-294
  if (!input || !syndromes || !results) return fail(LDPC_B200_ERR_INVALID, "null buffer");
  if (prm->num_iter_check_parity == 0) return fail(LDPC_B200_ERR_INVALID, "num_iter_check_parity must be > 0");
  CU_TRY(cudaSetDevice(d->device));
  const double t0 = now_s();
  ldpc_b200_report rep;
  memset(&rep, 0, sizeof(rep));
  if (report) *report = rep;

  // ---- shared queue: reserve the first batch now; later reservations keep `ahead` frames beyond what is loaded
  local_queue lq_store, *lq = nullptr;
  if (queue) {
    if (prm->flags & LDPC_B200_FLAG_SIMD_PAIRS)
      return fail(LDPC_B200_ERR_INVALID, "SIMD_PAIRS is not available when decoding from a shared queue");
    if (queue->sh->n_total != n_frames) return fail(LDPC_B200_ERR_INVALID, "queue size differs from n_frames");
    const ldpc_b200_status qs = lq_store.init(queue);
    if (qs != LDPC_B200_OK) return qs;
    lq = &lq_store;
    lq->reserve_upto(d->frames_in_flight);
    if (lq->avail == 0) return LDPC_B200_OK; // the other decoders have taken the whole queue
  }
  const uint32_t queue_ahead = prm->queue_lookahead_frames ? prm->queue_lookahead_frames : d->frames_in_flight;

  const bool in_dev = is_device_pointer(input), syn_dev = is_device_pointer(syndromes);
  const bool res_dev = is_device_pointer(results);
  const bool simd_pairs = (prm->flags & LDPC_B200_FLAG_SIMD_PAIRS) != 0;
  const bool time_kernels = (prm->flags & LDPC_B200_FLAG_TIME_KERNELS) != 0;
  const uint32_t B = std::min(lq ? lq->avail : n_frames, d->frames_in_flight); // num_vectors_gpu_batch
  const uint32_t k_check = prm->num_iter_check_parity;
  // even-count refills can never load an odd remainder (the reference's OpenCL harness only ever feeds
  // parallel_factor * loading_factor frames, ldpc_decoder_gpu.cpp:623 / main.cpp:320)
  if (simd_pairs && ((n_frames - B) & 1u))
    return fail(LDPC_B200_ERR_INVALID, "SIMD_PAIRS needs an even number of frames beyond the first batch");
  // with one frame in flight an even-count refill is always empty: nothing beyond the first frame could ever be loaded
  if (simd_pairs && B < 2 && n_frames > B)
    return fail(LDPC_B200_ERR_INVALID, "SIMD_PAIRS needs at least two frames in flight to load frames beyond the first");
  // no frame may stay longer than num_iter_max + one check period, so this bounds the whole run
  const uint64_t period_limit =
      (uint64_t(n_frames) / B + 2) * ((uint64_t(prm->num_iter_max) + 2ull * k_check + 2) / k_check + 2);

  // ---- scheduler constants (also the geometry of the staging ring)
  const bool packed = prm->input_layout == LDPC_B200_LAYOUT_PACKED_BITS;
  const bool frame_major = prm->input_layout == LDPC_B200_LAYOUT_FRAME_MAJOR || packed;
  // fp16 transfer with fp32 compute: the input holds IEEE halves, the refill kernel widens them
  const bool half_input = (prm->flags & LDPC_B200_FLAG_HALF_INPUT) != 0 && !packed;
  if (half_input && d->precision != LDPC_B200_FP32)
    return fail(LDPC_B200_ERR_INVALID, "HALF_INPUT is the transfer format of an FP32 decoder (an FP16 decoder takes halves anyway)");
  const int src_kind = packed ? dev::kSrcPackedBits : (half_input ? dev::kSrcHalf : dev::kSrcNative);
  // look-ahead slack of the staging ring (host-fed classic calls only: a decoder on a shared queue reserves, and so
  // stages, only what the next check may load)
  const bool frame_major_in = prm->input_layout != LDPC_B200_LAYOUT_INTERLEAVED;
  const uint32_t slack = (in_dev || lq || n_frames <= B) ? 0u : d->stage_slack(B, frame_major_in ? 1u : 32u);
  dev::SchedConst sc = make_sched_const(d, prm, n_frames, in_dev, syn_dev, res_dev, slack);
  sc.batch = B;
  sc.queue_block = lq ? lq->block : 0u;
  const dev::FrameQueueView qview = lq ? lq->view() : dev::FrameQueueView{nullptr, nullptr};
  // fused refill, decided per period by the scheduler kernel (see select_kernels): threshold in frames
  sc.fused_min_new = 0xFFFFFFFFu;
  if (d->k_cn_refill) sc.fused_min_new = d->fused_mode == 1 ? 1u : std::max<uint32_t>(1u, uint32_t(uint64_t(d->P) * d->fused_min_pct / 100));
  const bool fused_any = sc.fused_min_new != 0xFFFFFFFFu; // both variants are enqueued, the device picks
  if (!in_dev && sc.ring_frames < std::min<uint32_t>(d->P, B + sc.ring_block) && sc.ring_frames < B)
    return fail(LDPC_B200_ERR_INVALID, "staging ring too small (internal error)");
  if (!in_dev) {
    const ldpc_b200_status se = d->ensure_input_staging(sc.ring_frames);
    if (se != LDPC_B200_OK) return se;
  }
  if (!syn_dev) {
    const ldpc_b200_status se = d->ensure_syndrome_staging(sc.ring_frames);
    if (se != LDPC_B200_OK) return se;
  }
  if (!res_dev) {
    const ldpc_b200_status se = d->ensure_result_staging();
    if (se != LDPC_B200_OK) return se;
  }
  const void *src_dev = in_dev ? input : static_cast<const void *>(d->d_stage.p);
  const uint32_t *syn_src_dev = syn_dev ? syndromes : d->d_stage_syn.p;
  uint32_t *res_dst_dev = res_dev ? results : d->d_results.p;

  if (!d->d_iter.p || d->d_iter.n < 2ull * n_frames) CU_TRY(d->d_iter.alloc(2ull * n_frames));
  uint32_t *d_iter_start = d->d_iter.p, *d_iter_end = d->d_iter.p + n_frames;

  // ---- staging of host-resident frames into the device ring (copy stream): see host_feeder
  host_feeder feed;
  feed.d = d;
  feed.rep = &rep;
  feed.in_host = in_dev ? nullptr : static_cast<const uint8_t *>(input);
  feed.syn_host = syn_dev ? nullptr : syndromes;
  feed.frame_major = frame_major;
  feed.n_frames = n_frames;
  feed.n_cols = sc.n_cols;
  feed.n_reg = sc.n_regular;
  feed.ring_block = sc.ring_block;
  feed.ring_frames = sc.ring_frames;
  feed.row_elems = sc.row_elems;
  feed.es = half_input ? 2 : d->elem;
  feed.es_fm = packed ? 4 : (half_input ? 2 : d->elem);
  feed.frame_stride = sc.frame_stride;
  feed.lq = lq;
  auto stage_upto = [&](uint32_t upto) {
    const double ts = now_s();
    const ldpc_b200_status r = feed.stage_upto(upto);
    rep.host_stage_s += now_s() - ts;
    return r;
  };

  // ---- optional per-launch timing (LDPC_B200_FLAG_TIME_KERNELS): see launch_timer
  launch_timer timer;
  if (time_kernels) {
    const ldpc_b200_status ts = timer.create(k_check + 1, d->stream, &rep);
    if (ts != LDPC_B200_OK) return ts;
  }
  // whether the first check-node pass of the iterations in timer pool h was the fused one (known from the summaries)
  bool first_cn_fused[2] = {B >= sc.fused_min_new, false};
  auto flush_timing = [&](int h) { timer.flush(h, first_cn_fused[h]); };

  cudaEvent_t ev_first, ev_last;
  CU_TRY(cudaEventCreate(&ev_first));
  CU_TRY(cudaEventCreate(&ev_last));
  struct ev_guard {
    cudaEvent_t a, b;
    ~ev_guard() {
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } guard{ev_first, ev_last};

  // ---- first batch: frames 0..B-1 into slots 0..B-1
  ldpc_b200_status st = stage_upto(B);
  if (st != LDPC_B200_OK) return st;
  rep.first_batch_stage_s = rep.host_stage_s; // the device has nothing to do until the first batch is staged
  CU_TRY(cudaEventRecord(ev_first, d->stream));
  dev::sched_init_kernel<<<1, dev::kSchedThreads, 0, d->stream>>>(sc, d->d_state.p, d->d_in_gpu.p, d->d_virt.p,
                                                                  d_iter_start, d_iter_end, d->d_refill.p, d->d_n_new.p,
                                                                  d->d_new_mask.p, d->d_retire_mask.p, d->d_slot_entry.p,
                                                                  qview);
  auto enqueue_refill = [&]() {
    if (fused_any) {
      d->launch_refill_kernels(src_dev, syn_src_dev, prm->channel, prm->channel_param, src_kind, false, dev::kGateFused, true,
                               !frame_major);
      d->launch_refill_kernels(src_dev, syn_src_dev, prm->channel, prm->channel_param, src_kind, true, dev::kGatePlain, false);
    } else {
      d->launch_refill_kernels(src_dev, syn_src_dev, prm->channel, prm->channel_param, src_kind, true);
    }
  };
  enqueue_refill();
  CU_TRY(cudaEventRecord(d->ev_refilled, d->stream));
  feed.refill_recorded = true;
  rep.other_launches += 3 + (fused_any ? 1 : 0);
  rep.cn_refill_launches += first_cn_fused[0] ? 1 : 0;
  CU_TRY(cudaGetLastError());

  // launch accounting of one period (shared by the plain-launch and graph-replay paths): the gated launch that
  // exits at once is counted with the service launches
  auto count_iterations = [&](uint32_t n_it) {
    rep.cn_launches += n_it;
    rep.vn_launches += n_it;
    rep.other_launches += 1 + (fused_any ? 1 : 0); // syndrome check (+ the gated-off check-node variant)
    rep.n_checks_run++;
  };
  auto count_service = [&]() { rep.other_launches += 4 + sc.early_retire + (fused_any ? 1 : 0); };
  auto enqueue_iterations = [&](uint64_t period) {
    const int h = int(period & 1);
    const uint32_t n_it = period == 0 ? k_check + 1 : k_check; // g = 0..k, then k more per period (:350)
    for (uint32_t i = 0; i < n_it; ++i) {
      const bool last = i + 1 == n_it;
      timer.mark_iter(h, i, 0);
      if (fused_any && i == 0) { // the device runs exactly one of the two (SchedState::fused_now)
        d->launch_cn_refill(dev::kGateFused);
        d->launch_cn(dev::kGatePlain);
      } else {
        d->launch_cn();
      }
      timer.mark_iter(h, i, 1);
      d->launch_vn(last);
      timer.mark_iter(h, i, 2);
    }
    timer.set_iters(h, n_it);
    count_iterations(n_it);
    timer.mark_service(h, 0);
    d->launch_check();
    timer.mark_service(h, 1);
  };
  // while capturing, host-visible events must be recorded as external event nodes
  bool capturing = false;
  auto record = [&](cudaEvent_t ev) {
    return cudaEventRecordWithFlags(ev, d->stream, capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
  };
  auto enqueue_service = [&](uint64_t period) -> ldpc_b200_status {
    const int h = int(period & 1);
    dev::sched_kernel<<<1, dev::kSchedThreads, 0, d->stream>>>(
        sc, d->d_state.p, d->d_viol.p, d->d_in_gpu.p, d->d_virt.p, d_iter_start, d_iter_end, d->d_refill.p,
        d->d_extract.p, d->d_n_new.p, d->d_new_mask.p, d->d_retire_mask.p, d->d_slot_entry.p, d->d_retired.p,
        d->d_summary.p, d->d_scratch.p, d->d_moves.p, d->d_n_moves.p, qview);
    d->launch_extract_kernel(res_dst_dev);
    if (sc.early_retire) { // tail compaction of the frames still running (no-op unless the scheduler asked for moves)
      if (d->elem == 4)
        dev::move_slots_kernel<float><<<d->grid_misc, dev::kThreads, 0, d->stream>>>(
            reinterpret_cast<float *>(d->d_msg.p), reinterpret_cast<float *>(d->d_llr0.p), d->d_syn.p, d->d_moves.p,
            d->d_n_moves.p, d->E, d->N, d->M, d->log2P);
      else
        dev::move_slots_kernel<uint16_t><<<d->grid_misc, dev::kThreads, 0, d->stream>>>(
            reinterpret_cast<uint16_t *>(d->d_msg.p), reinterpret_cast<uint16_t *>(d->d_llr0.p), d->d_syn.p,
            d->d_moves.p, d->d_n_moves.p, d->E, d->N, d->M, d->log2P);
    }
    timer.mark_service(h, 2);
    CU_TRY(record(d->ev_extracted[h]));
    enqueue_refill();
    CU_TRY(record(d->ev_refilled));
    timer.mark_service(h, 3);
    CU_TRY(cudaMemcpyAsync(d->h_summary.p + h, d->d_summary.p, sizeof(dev::SchedSummary), cudaMemcpyDeviceToHost,
                           d->stream));
    if (!res_dev)
      CU_TRY(cudaMemcpyAsync(d->h_retired.p + size_t(h) * d->P, d->d_retired.p, sizeof(uint32_t) * B,
                             cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(record(d->ev_period[h]));
    count_service();
    return LDPC_B200_OK;
  };

  // ---- main loop: the device is always one period ahead of the host
  struct graph_pair {
    cudaGraphExec_t exec[2] = {nullptr, nullptr};
    ~graph_pair() {
      for (cudaGraphExec_t e : exec)
        if (e) cudaGraphExecDestroy(e);
    }
  } graphs;
  bool use_graph = !time_kernels && !(prm->flags & LDPC_B200_FLAG_NO_GRAPH);
  uint64_t period = 0;
  uint32_t next_known = B; // next_to_load as of the last summary read
  rep.slot_iterations = uint64_t(k_check + 1) * d->P; // period 0
  enqueue_iterations(0);
  for (;;) {
    if (period > 0) {
      const int hp = int((period - 1) & 1);
      CU_TRY(cudaEventSynchronize(d->ev_period[hp])); // the host sleeps here while the next iterations run
      flush_timing(hp);
      const dev::SchedSummary sm = d->h_summary.p[hp];
      rep.d2h_bytes += sizeof(dev::SchedSummary);
      rep.frames_refilled += sm.n_new;
      next_known = sm.next_to_load;
      first_cn_fused[int(period & 1)] = sm.fused != 0; // the iterations that follow this service
      rep.cn_refill_launches += sm.fused ? 1 : 0;
      if (!res_dev && sm.n_retired) { // drain the result ring of period-1 on the copy stream
        const size_t bytes = size_t(sm.n_retired) * d->W * 4;
        CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_extracted[hp], 0));
        CU_TRY(cudaMemcpyAsync(d->h_results.p, d->d_results.p, bytes, cudaMemcpyDeviceToHost, d->copy_stream));
        CU_TRY(cudaStreamSynchronize(d->copy_stream));
        const uint32_t *frames = d->h_retired.p + size_t(hp) * d->P;
        parallel_for(sm.n_retired, d->host_threads, [&](size_t a, size_t b2) {
          for (size_t t = a; t < b2; ++t)
            memcpy(results + size_t(lq ? lq->global_of(frames[t]) : frames[t]) * d->W, d->h_results.p + t * d->W,
                   size_t(d->W) * 4);
        });
        rep.d2h_bytes += bytes + sizeof(uint32_t) * B;
      }
      if (sm.done) {
        rep.global_iters = sm.g; // the loop index at which the reference breaks (:414-462)
        break;
      }
      rep.slot_iterations += uint64_t(k_check) << sm.log2_p_active; // the k iterations that follow this check
    }
    // frames the next refill may need: at most B beyond what has been loaded so far
    if (lq) lq->reserve_upto(uint64_t(next_known) + queue_ahead);
    {
      // The next check may load up to B frames beyond next_known: they must be staged before it is enqueued.  With a
      // look-ahead slack in the ring this is normally already the case (the top-up below ran a period ago), so the
      // host enqueues the next period at once and gathers AFTERWARDS, while the device works through it.
      const uint64_t need = std::min<uint64_t>(n_frames, uint64_t(next_known) + B);
      if (uint64_t(feed.staged) < need) st = stage_upto(uint32_t(need));
      if (st != LDPC_B200_OK) return st;
    }
    // service of this period + the next period's flood iterations (run ahead: no-ops once the scheduler has
    // set `done`).  The sequence is fixed, so it is captured once per parity into a CUDA graph and replayed.
    const int hg = int(period & 1);
    if (use_graph && !graphs.exec[hg]) {
      cudaGraph_t g = nullptr;
      if (cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        capturing = true;
        st = enqueue_service(period);
        enqueue_iterations(period + 1);
        capturing = false;
        const cudaError_t ce = cudaStreamEndCapture(d->stream, &g);
        if (st != LDPC_B200_OK || ce != cudaSuccess || !g ||
            cudaGraphInstantiate(&graphs.exec[hg], g, 0) != cudaSuccess) {
          graphs.exec[hg] = nullptr;
          use_graph = false; // fall back to plain launches
        }
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        // the counters were bumped while capturing; the launch below is what actually runs
      } else {
        cudaGetLastError();
        use_graph = false;
      }
      if (!use_graph) {
        st = enqueue_service(period);
        if (st != LDPC_B200_OK) return st;
        enqueue_iterations(period + 1);
      } else {
        CU_TRY(cudaGraphLaunch(graphs.exec[hg], d->stream));
        rep.graph_launches++;
      }
    } else if (use_graph) {
      CU_TRY(cudaGraphLaunch(graphs.exec[hg], d->stream));
      rep.graph_launches++;
      count_service();
      count_iterations(k_check);
    } else {
      st = enqueue_service(period);
      if (st != LDPC_B200_OK) return st;
      enqueue_iterations(period + 1);
    }
    CU_TRY(cudaGetLastError());
    if (slack) { // top the ring up for the checks after this one: off the critical path of the launches above
      st = stage_upto(uint32_t(std::min<uint64_t>(n_frames, uint64_t(next_known) + B + slack)));
      if (st != LDPC_B200_OK) return st;
    }
    ++period;
    if (period > period_limit) return fail(LDPC_B200_ERR_INVALID, "scheduler made no progress (internal error)");
  }
  CU_TRY(cudaEventRecord(ev_last, d->stream));
  CU_TRY(cudaStreamSynchronize(d->stream));
  CU_TRY(cudaStreamSynchronize(d->copy_stream));
  float ms = 0;
  CU_TRY(cudaEventElapsedTime(&ms, ev_first, ev_last));
  rep.device_s = ms * 1e-3;
  // the run-ahead period was enqueued but did no work
  rep.cn_launches -= k_check;
  rep.vn_launches -= k_check;
  rep.other_launches -= 1 + (fused_any ? 1 : 0);
  rep.n_checks_run -= 1;

  // per-frame iteration counters live on the device (:616-628), indexed by this decoder's local frame numbers
  const uint32_t n_local = lq ? next_known : n_frames; // frames this decoder loaded (all of them, for a classic call)
  std::vector<uint32_t> it(2ull * n_frames);
  CU_TRY(cudaMemcpy(it.data(), d->d_iter.p, sizeof(uint32_t) * 2ull * n_frames, cudaMemcpyDeviceToHost));
  rep.max_iter = 0;
  rep.min_iter = uint32_t(-1);
  double avg = 0;
  for (uint32_t j = 0; j < n_local; ++j) {
    const uint32_t n = it[n_frames + j] - it[j];
    rep.max_iter = std::max(rep.max_iter, n);
    rep.min_iter = std::min(rep.min_iter, n);
    avg += n;
    if (iter_counts) iter_counts[lq ? lq->global_of(j) : j] = n;
  }
  rep.frames_decoded = n_local;
  rep.avg_iter = float(avg / n_local);
  rep.iter_time_per_vector = rep.global_iters ? float(rep.device_s / (double(rep.global_iters) * B)) : 0.f;
  rep.elapsed_s = now_s() - t0;
  if (report) *report = rep;
  return LDPC_B200_OK;
}

extern "C" {

// decode(): the reference's scheduling semantics (ldpc_decoder_gpu.cu:283-634) with the decisions taken on
// the device.  Per check period the compute stream runs
//     [CN, VN] x k  ->  syndrome check  ->  sched_kernel  ->  extract  ->  refill
// without waiting for the host; the host only (a) keeps one period of work queued ahead, (b) stages the next
// frames into a device ring over pinned cudaMemcpyAsync on the copy stream while the flood iterations run,
// (c) drains retired frames' packed decisions on the copy stream, (d) watches the summary for the end.
ldpc_b200_status ldpc_b200_decode(ldpc_b200_decoder *d, const ldpc_b200_decode_params *prm, uint32_t n_frames,
                                  const void *input, const uint32_t *syndromes, uint32_t *results,
                                  ldpc_b200_report *report, uint32_t *iter_counts) {
  return decode_impl(d, prm, n_frames, input, syndromes, results, report, iter_counts, nullptr);
}

// ------------------------------------------------------------------ shared frame queue: C ABI
ldpc_b200_status ldpc_b200_queue_create(const char *shm_name, int32_t create, uint32_t n_frames, uint32_t block_frames,
                                        ldpc_b200_queue **out) {
  if (!out) return fail(LDPC_B200_ERR_INVALID, "null output pointer");
  *out = nullptr;
  std::unique_ptr<ldpc_b200_queue> q(new ldpc_b200_queue);
  if (!shm_name) { // private to this process
    if (!n_frames || !block_frames) return fail(LDPC_B200_ERR_INVALID, "queue needs a size and a block size");
    q->sh = &q->local;
    q->sh->cursor.store(0);
    q->sh->n_total = n_frames;
    q->sh->block = block_frames;
    q->sh->magic = kQueueMagic;
  } else {
    q->shm_name = shm_name[0] == '/' ? shm_name : std::string("/") + shm_name;
    if (create && (!n_frames || !block_frames)) return fail(LDPC_B200_ERR_INVALID, "queue needs a size and a block size");
    if (create) shm_unlink(q->shm_name.c_str()); // a stale object of a crashed run
    const int fd = shm_open(q->shm_name.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return fail(LDPC_B200_ERR_IO, "shm_open(" + q->shm_name + ") failed");
    if (create && ftruncate(fd, sizeof(queue_shared)) != 0) {
      close(fd);
      return fail(LDPC_B200_ERR_IO, "ftruncate on the shared queue failed");
    }
    void *m = mmap(nullptr, sizeof(queue_shared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(LDPC_B200_ERR_IO, "mmap of the shared queue failed");
    q->sh = static_cast<queue_shared *>(m);
    q->map_len = sizeof(queue_shared);
    q->owner = create != 0;
    if (create) {
      new (&q->sh->cursor) std::atomic<uint64_t>(0);
      q->sh->n_total = n_frames;
      q->sh->block = block_frames;
      std::atomic_thread_fence(std::memory_order_release);
      q->sh->magic = kQueueMagic;
    } else if (q->sh->magic != kQueueMagic) {
      return fail(LDPC_B200_ERR_INVALID, "shared queue is not initialised (open it after the creating process returned)");
    }
  }
  *out = q.release();
  return LDPC_B200_OK;
}

void ldpc_b200_queue_destroy(ldpc_b200_queue *q) { delete q; }

uint32_t ldpc_b200_queue_size(const ldpc_b200_queue *q) { return q ? q->sh->n_total : 0; }

uint32_t ldpc_b200_queue_take(ldpc_b200_queue *q, uint32_t *first_frame) {
  if (!q || !first_frame) return 0;
  const uint64_t start = q->sh->cursor.fetch_add(q->sh->block, std::memory_order_relaxed);
  if (start >= q->sh->n_total) return 0;
  *first_frame = uint32_t(start);
  return uint32_t(std::min<uint64_t>(q->sh->block, q->sh->n_total - start));
}

void ldpc_b200_queue_reset(ldpc_b200_queue *q) {
  if (q) q->sh->cursor.store(0);
}

ldpc_b200_status ldpc_b200_decode_from_queue(ldpc_b200_decoder *d, const ldpc_b200_decode_params *prm,
                                             ldpc_b200_queue *queue, const void *input, const uint32_t *syndromes,
                                             uint32_t *results, ldpc_b200_report *report, uint32_t *iter_counts) {
  if (!queue) return fail(LDPC_B200_ERR_INVALID, "null queue");
  return decode_impl(d, prm, queue->sh->n_total, input, syndromes, results, report, iter_counts, queue);
}

// One queue, one host thread per decoder (= per GPU).  The merged report sums the counters, takes the longest
// elapsed / device time and the iteration statistics over all frames.
ldpc_b200_status ldpc_b200_decode_multi(ldpc_b200_decoder *const *decs, uint32_t n_decs,
                                        const ldpc_b200_decode_params *prm, uint32_t n_frames, const void *input,
                                        const uint32_t *syndromes, uint32_t *results, ldpc_b200_report *report,
                                        ldpc_b200_report *per_decoder, uint32_t *iter_counts, uint32_t block_frames) {
  if (!decs || !n_decs || !prm) return fail(LDPC_B200_ERR_INVALID, "null argument");
  if (n_frames == 0) return LDPC_B200_OK;
  uint32_t blk = block_frames;
  if (!blk) { // default: fine enough to balance, coarse enough that a block is a few host-staging runs
    blk = 256;
    for (uint32_t i = 0; i < n_decs; ++i)
      if (decs[i]) blk = std::min(blk, std::max(32u, decs[i]->frames_in_flight));
  }
  ldpc_b200_queue *q = nullptr;
  ldpc_b200_status st = ldpc_b200_queue_create(nullptr, 1, n_frames, blk, &q);
  if (st != LDPC_B200_OK) return st;
  std::unique_ptr<ldpc_b200_queue> guard(q);
  std::vector<ldpc_b200_report> reps(n_decs);
  std::vector<ldpc_b200_status> sts(n_decs, LDPC_B200_OK);
  std::vector<std::string> errs(n_decs);
  std::vector<std::thread> threads;
  for (uint32_t i = 0; i < n_decs; ++i)
    threads.emplace_back([&, i] {
      memset(&reps[i], 0, sizeof(ldpc_b200_report));
      sts[i] = ldpc_b200_decode_from_queue(decs[i], prm, q, input, syndromes, results, &reps[i], iter_counts);
      if (sts[i] != LDPC_B200_OK) errs[i] = g_last_error; // thread-local: carry it to the caller's thread
    });
  for (std::thread &t : threads) t.join();
  for (uint32_t i = 0; i < n_decs; ++i)
    if (sts[i] != LDPC_B200_OK) return fail(sts[i], "decoder " + std::to_string(i) + ": " + errs[i]);
  ldpc_b200_report m;
  memset(&m, 0, sizeof(m));
  m.min_iter = uint32_t(-1);
  double avg = 0;
  for (uint32_t i = 0; i < n_decs; ++i) {
    const ldpc_b200_report &r = reps[i];
    if (per_decoder) per_decoder[i] = r;
    if (!r.frames_decoded) continue;
    m.max_iter = std::max(m.max_iter, r.max_iter);
    m.min_iter = std::min(m.min_iter, r.min_iter);
    avg += double(r.avg_iter) * double(r.frames_decoded);
    m.frames_decoded += r.frames_decoded;
    m.global_iters = std::max(m.global_iters, r.global_iters);
    m.n_checks_run += r.n_checks_run;
    m.elapsed_s = std::max(m.elapsed_s, r.elapsed_s);
    m.device_s = std::max(m.device_s, r.device_s);
    m.cn_s += r.cn_s, m.vn_s += r.vn_s, m.check_s += r.check_s, m.retire_s += r.retire_s, m.refill_s += r.refill_s;
    m.cn_refill_s += r.cn_refill_s;
    m.cn_launches += r.cn_launches, m.vn_launches += r.vn_launches, m.other_launches += r.other_launches;
    m.cn_refill_launches += r.cn_refill_launches;
    m.h2d_bytes += r.h2d_bytes, m.d2h_bytes += r.d2h_bytes;
    m.frames_refilled += r.frames_refilled, m.graph_launches += r.graph_launches, m.slot_iterations += r.slot_iterations;
    m.host_stage_s = std::max(m.host_stage_s, r.host_stage_s);
    m.first_batch_stage_s = std::max(m.first_batch_stage_s, r.first_batch_stage_s);
  }
  m.avg_iter = m.frames_decoded ? float(avg / double(m.frames_decoded)) : 0.f;
  // the reference's per-frame iteration time (ldpc_decoder_gpu.cu:627) over the frames in flight on all devices
  double slots = 0;
  for (uint32_t i = 0; i < n_decs; ++i)
    if (reps[i].iter_time_per_vector > 0) slots += 1.0 / reps[i].iter_time_per_vector;
  m.iter_time_per_vector = slots > 0 ? float(1.0 / slots) : 0.f;
  if (report) *report = m;
  return LDPC_B200_OK;
}

// ------------------------------------------------------------------ kernel-level entry points
ldpc_b200_status ldpc_b200_debug_refill(ldpc_b200_decoder *d, int32_t channel, float channel_param, uint32_t cnt,
                                        const void *staged, const uint32_t *syndromes) {
  if (!d || !staged || !syndromes || cnt == 0 || cnt > d->P) return fail(LDPC_B200_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(d->device));
  input_view in;
  in.base = static_cast<const uint8_t *>(staged);
  in.syn = syndromes;
  in.on_device = is_device_pointer(staged);
  in.syn_on_device = is_device_pointer(syndromes);
  in.layout = LDPC_B200_LAYOUT_INTERLEAVED;
  in.frame_stride = 0;
  in.n_cols = cnt;
  std::vector<uint32_t> slots(cnt);
  for (uint32_t j = 0; j < cnt; ++j) slots[j] = j;
  CU_TRY(cudaMemsetAsync(d->d_state.p, 0, sizeof(dev::SchedState), d->stream)); // a finished decode leaves done = 1
  const void *src_dev = nullptr;
  const uint32_t *syn_dev = nullptr;
  ldpc_b200_status st = stage_frames(d, in, 0, cnt, slots.data(), &src_dev, &syn_dev, nullptr);
  if (st != LDPC_B200_OK) return st;
  d->launch_refill(cnt, src_dev, syn_dev, channel, channel_param);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(d->stream));
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_debug_iterate(ldpc_b200_decoder *d, uint32_t n_iters, int32_t final_bits, float *cn_ms,
                                         float *vn_ms) {
  if (!d) return fail(LDPC_B200_ERR_INVALID, "null decoder");
  CU_TRY(cudaSetDevice(d->device));
  float cn_acc = 0, vn_acc = 0;
  const bool timed = cn_ms || vn_ms;
  cudaEvent_t e0 = d->ev_a, e1 = d->ev_b, e2 = nullptr;
  if (timed) CU_TRY(cudaEventCreate(&e2));
  for (uint32_t it = 0; it < n_iters; ++it) {
    if (timed) CU_TRY(cudaEventRecord(e0, d->stream));
    d->launch_cn();
    if (timed) CU_TRY(cudaEventRecord(e1, d->stream));
    d->launch_vn(final_bits && it + 1 == n_iters);
    if (timed) {
      CU_TRY(cudaEventRecord(e2, d->stream));
      CU_TRY(cudaEventSynchronize(e2));
      float a = 0, b = 0;
      CU_TRY(cudaEventElapsedTime(&a, e0, e1));
      CU_TRY(cudaEventElapsedTime(&b, e1, e2));
      cn_acc += a;
      vn_acc += b;
    }
  }
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(d->stream));
  if (e2) cudaEventDestroy(e2);
  if (cn_ms) *cn_ms = cn_acc;
  if (vn_ms) *vn_ms = vn_acc;
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_debug_cn_pass(ldpc_b200_decoder *d) {
  if (!d) return fail(LDPC_B200_ERR_INVALID, "null decoder");
  CU_TRY(cudaSetDevice(d->device));
  d->launch_cn();
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(d->stream));
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_debug_vn_pass(ldpc_b200_decoder *d, int32_t final_bits) {
  if (!d) return fail(LDPC_B200_ERR_INVALID, "null decoder");
  CU_TRY(cudaSetDevice(d->device));
  d->launch_vn(final_bits != 0);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(d->stream));
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_debug_check(ldpc_b200_decoder *d, uint8_t *violated) {
  if (!d || !violated) return fail(LDPC_B200_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(d->device));
  d->launch_check();
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(d->h_viol.p, d->d_viol.p, sizeof(uint32_t) << d->log2PW, cudaMemcpyDeviceToHost, d->stream));
  CU_TRY(cudaStreamSynchronize(d->stream));
  for (uint32_t s = 0; s < d->P; ++s) violated[s] = (d->h_viol.p[s >> 5] >> (s & 31)) & 1u;
  return LDPC_B200_OK;
}

// pos: storage row of each original row (null: identity)
static ldpc_b200_status read_column(ldpc_b200_decoder *d, const uint8_t *src, uint64_t rows, uint32_t slot, float *out,
                                    const std::vector<uint32_t> *pos = nullptr) {
  if (slot >= d->P) return fail(LDPC_B200_ERR_INVALID, "slot out of range");
  CU_TRY(cudaSetDevice(d->device));
  dev_buf<uint8_t> tmp;
  CU_TRY(tmp.alloc(rows * d->elem));
  const unsigned blocks = unsigned((rows + 255) / 256);
  std::vector<uint8_t> host(rows * d->elem);
  if (d->elem == 4)
    dev::gather_slot_kernel<float><<<blocks, 256, 0, d->stream>>>(reinterpret_cast<const float *>(src),
                                                                  reinterpret_cast<float *>(tmp.p), rows, d->log2P, slot);
  else
    dev::gather_slot_kernel<uint16_t><<<blocks, 256, 0, d->stream>>>(
        reinterpret_cast<const uint16_t *>(src), reinterpret_cast<uint16_t *>(tmp.p), rows, d->log2P, slot);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(host.data(), tmp.p, rows * d->elem, cudaMemcpyDeviceToHost, d->stream));
  CU_TRY(cudaStreamSynchronize(d->stream));
  const bool perm = pos && !pos->empty();
  if (d->elem == 4) {
    const float *f = reinterpret_cast<const float *>(host.data());
    for (uint64_t i = 0; i < rows; ++i) out[i] = f[perm ? (*pos)[i] : i];
  } else {
    const __half *h = reinterpret_cast<const __half *>(host.data());
    for (uint64_t i = 0; i < rows; ++i) out[i] = __half2float(h[perm ? (*pos)[i] : i]);
  }
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_debug_read_messages(ldpc_b200_decoder *d, uint32_t slot, float *out_e) {
  if (!d || !out_e) return fail(LDPC_B200_ERR_INVALID, "null argument");
  return read_column(d, d->d_msg.p, d->E, slot, out_e, &d->h_edge_pos);
}

ldpc_b200_status ldpc_b200_debug_read_priors(ldpc_b200_decoder *d, uint32_t slot, float *out_n) {
  if (!d || !out_n) return fail(LDPC_B200_ERR_INVALID, "null argument");
  return read_column(d, d->d_llr0.p, d->N, slot, out_n, &d->h_var_pos);
}

ldpc_b200_status ldpc_b200_debug_read_hard(ldpc_b200_decoder *d, uint32_t slot, uint32_t *out_words) {
  if (!d || !out_words || slot >= d->P) return fail(LDPC_B200_ERR_INVALID, "bad argument");
  CU_TRY(cudaSetDevice(d->device));
  uint32_t frame0 = 0;
  ldpc_b200_status st = retire_frames(d, 1, &slot, &frame0, out_words, false, nullptr);
  return st;
}

ldpc_b200_status ldpc_b200_phi_table_fp16(ldpc_b200_decoder *d, uint16_t *table) {
  if (!d || !table) return fail(LDPC_B200_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(d->device));
  CU_TRY(cudaMemcpy(table, d->d_phi_table.p, dev::kPhiTableSize * 2, cudaMemcpyDeviceToHost));
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_synthesize_frames(ldpc_b200_decoder *d, int32_t channel, float noise, uint64_t start_idx,
                                             uint32_t n_frames, void *d_input, uint32_t *d_syndromes,
                                             uint32_t *d_ref_bits) {
  if (!d || !d_input || !d_syndromes || !d_ref_bits) return fail(LDPC_B200_ERR_INVALID, "null argument");
  if (channel != LDPC_B200_AWGN && channel != LDPC_B200_BSC) return fail(LDPC_B200_ERR_INVALID, "unknown channel");
  if (n_frames == 0) return LDPC_B200_OK;
  if (!is_device_pointer(d_input) || !is_device_pointer(d_syndromes) || !is_device_pointer(d_ref_bits))
    return fail(LDPC_B200_ERR_INVALID, "synthesize_frames writes device memory");
  CU_TRY(cudaSetDevice(d->device));
  const uint32_t n_groups = (n_frames + 31) / 32;
  dev_buf<uint32_t> sliced, syn_sliced;
  CU_TRY(sliced.alloc(size_t(d->N) * n_groups));
  CU_TRY(syn_sliced.alloc(size_t(d->M) * n_groups));
  const uint64_t n_blocks = uint64_t((d->N + 15) / 16) * n_groups;
  dev::synth_bits_kernel<<<unsigned((n_blocks + 255) / 256), 256, 0, d->stream>>>(sliced.p, start_idx, d->N, n_groups);
  const uint32_t n_tx = d->N - d->n_erased;
  if (d->elem == 4)
    dev::synth_noise_kernel<float><<<(n_frames + 63) / 64, 64, 0, d->stream>>>(
        sliced.p, static_cast<float *>(d_input), d_ref_bits, start_idx, n_frames, d->N, n_tx, n_groups, channel, noise);
  else
    dev::synth_noise_kernel<__half><<<(n_frames + 63) / 64, 64, 0, d->stream>>>(
        sliced.p, static_cast<__half *>(d_input), d_ref_bits, start_idx, n_frames, d->N, n_tx, n_groups, channel, noise);
  const uint64_t n_syn = uint64_t(d->M) * n_groups;
  // the syndromes are those of the code as given: original check and variable numbering
  dev::synth_syndrome_kernel<<<unsigned((n_syn + 255) / 256), 256, 0, d->stream>>>(
      sliced.p, syn_sliced.p, d->reordered ? d->d_row_ptr_orig.p : d->d_row_ptr.p,
      d->reordered ? d->d_edge_var_orig.p : d->d_edge_var.p, d->M, n_groups);
  const uint64_t n_warps = uint64_t(d->S) * n_groups;
  dev::synth_syndrome_transpose_kernel<<<unsigned((n_warps * 32 + 255) / 256), 256, 0, d->stream>>>(
      syn_sliced.p, d_syndromes, d->M, n_groups, n_frames, d->S);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(d->stream));
  return LDPC_B200_OK;
}

ldpc_b200_status ldpc_b200_debug_phi_abs_f32(const float *x, float *y, size_t n) {
  if (!x || !y) return fail(LDPC_B200_ERR_INVALID, "null argument");
  dev_buf<float> dx, dy;
  CU_TRY(dx.alloc(n));
  CU_TRY(dy.alloc(n));
  CU_TRY(cudaMemcpy(dx.p, x, n * 4, cudaMemcpyHostToDevice));
  dev::phi_abs_f32_kernel<<<unsigned((n + 255) / 256), 256>>>(dx.p, dy.p, n);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpy(y, dy.p, n * 4, cudaMemcpyDeviceToHost));
  return LDPC_B200_OK;
}

} // extern "C"
