// ldpc_kernels.cuh -- sm_100a device kernels of the B200 flood decoder.
//
// Data layout in HBM (P = frame slots, power of two >= 32; s = sizeof(llr)):
//   msg   [E][P]      llr   check-major edge id, frame slot fastest (one array, updated in place:
//                           v2c in the "phi domain" after a variable-node pass, c2v after a check-node pass)
//   llr0  [N][P]      llr   channel priors
//   syn   [M][P/32]   u32   target syndromes, bit-sliced over slots (slot s <-> byte s>>3, bit s&7)
//   hard  [N][P/32]   u32   posterior hard decisions, bit-sliced the same way
//   viol  [P/32]      u32   "some check violated" bit per slot
// A thread owns one 16-byte vector of a row: 4 frames in fp32, 8 frames in fp16 (4 half2 words).
// All per-word arithmetic goes through a Traits struct so both precisions share the kernels.
//
// What each kernel replaces in the reference (src/cuda/flood.cu):
//   cn_update_kernel      flood_backward              :77-115
//   vn_update_kernel      flood_forward[_w_final_bits] :117-189
//   syndrome_check_kernel check_parity                :191-223 (on bit-sliced words, 32 frames per XOR)
//   refill_kernel         llr_bsc/llr_biawgn + flood_refill :47-75, 297-329 (direct to freed slots)
//   refill_syndrome_kernel flood_refill syndrome part :324-328
//   extract_kernel        deinterlace_output          :277-295 (only retiring slots)
// flood_permute_vecs (:225-275) has no equivalent: slots are indirect, frames never move.
#pragma once

#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ldpc_b200 {
namespace dev {

#ifndef LDPC_THREADS
#define LDPC_THREADS 256
#endif
#ifndef LDPC_HMIN_CLAMP
#define LDPC_HMIN_CLAMP 1 // fp16 table index: packed-half clamp + byte offsets (A/B on B200: CN +2.4 %, VN +5.1 %)
#endif
#ifndef LDPC_REFILL_UNROLL
#define LDPC_REFILL_UNROLL 1 // variables per loop trip of the priors-only refill kernel (A/B: 1 -> 0.377 s, 4 -> 0.480 s, 8 -> 0.571 s per bench step)
#endif
#ifndef LDPC_MERGE_DECISION_STORES
#define LDPC_MERGE_DECISION_STORES 1 // variable-node pass with decisions: one 32-bit store per 32 slots (see vn_update_kernel)
#endif
constexpr int kThreads = LDPC_THREADS; // CTA size of the streaming kernels
// Rows of degree <= D are held in registers (one read, one write per message); D is a template parameter chosen
// per code from {4, 6, 8} (host: pick_reg_degree).  Fewer registers buy a resident CTA: measured on B200
// (scripts/kernel_perf.py, README-shaped code) D=6 at 4 CTAs/SM is 4-9 % faster than D=8 at 3 CTAs/SM.
#ifndef LDPC_MINBLOCKS_D4
#define LDPC_MINBLOCKS_D4 4
#endif
constexpr int min_blocks_for(int D) { return D <= 4 ? LDPC_MINBLOCKS_D4 : (D <= 6 ? 4 : 3); }
// The check-node kernel with D <= 6 runs 512-thread CTAs (2 per SM instead of 4 x 256): same resident threads, half
// the table staging and CTA turnover -- +2-3 % measured; larger D and the variable-node kernel showed no gain.
constexpr int cn_threads_for(int D) { return D <= 6 ? 512 : kThreads; }
#ifndef LDPC_VN_THREADS
#define LDPC_VN_THREADS kThreads
#endif
constexpr int vn_threads_for(int D) { return D <= 6 ? LDPC_VN_THREADS : kThreads; }
#ifndef LDPC_PREFETCH_F32_VN
#define LDPC_PREFETCH_F32_VN false
#endif
#ifndef LDPC_PREFETCH_F16_CN
#define LDPC_PREFETCH_F16_CN false
#endif
#ifndef LDPC_PREFETCH_F32_CN
#define LDPC_PREFETCH_F32_CN false
#endif
constexpr int kPhiTableSize = 32768; // fp16 phi_abs table entries (all non-negative half bit patterns)

// ---------------------------------------------------------------- memory helpers
// Cache policy of the streaming accesses (A/B-tested on B200 with scripts/ab_kernels.py; the defaults are what won):
//   LDPC_LD_MODE 0 = ld.global.cs (evict-first), 1 = ld.global (default), 2 = ld.global.cg, 3 = ld.global.nc.L1::no_allocate
//   LDPC_ST_MODE 0 = st.global.cs,               1 = st.global (write-back), 2 = st.global.cg
#ifndef LDPC_LD_MODE
#define LDPC_LD_MODE 0
#endif
#ifndef LDPC_ST_MODE
#define LDPC_ST_MODE 0
#endif
#if LDPC_LD_MODE == 0
#define LDPC_LD_OP "ld.global.cs.v4.u32"
#elif LDPC_LD_MODE == 1
#define LDPC_LD_OP "ld.global.v4.u32"
#elif LDPC_LD_MODE == 2
#define LDPC_LD_OP "ld.global.cg.v4.u32"
#else
#define LDPC_LD_OP "ld.global.nc.L1::no_allocate.v4.u32"
#endif
#if LDPC_ST_MODE == 0
#define LDPC_ST_OP "st.global.cs.v4.u32"
#elif LDPC_ST_MODE == 1
#define LDPC_ST_OP "st.global.v4.u32"
#else
#define LDPC_ST_OP "st.global.cg.v4.u32"
#endif
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) { // read once: evict-first
  uint4 r;
  asm volatile(LDPC_LD_OP " {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
  asm volatile(LDPC_ST_OP " [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_cached(const uint4 *p) { // re-read soon (large-degree rows): keep in L1/L2
  uint4 r;
  asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---------------------------------------------------------------- phi, fp32
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// phi_abs(x) = -ln tanh(x/2) on x >= 0, with the reference's clamp and Taylor switch
// (flood.cu:14, 31-37): xm = max(x,1e-5); xm > 5 ? 2 e^-xm : ln((1+e^-xm)/(1-e^-xm)).
// 3 MUFU (ex2, 2x lg2) instead of libm exp/expm1/log + IEEE divide.  The cancellation in
// 1-e^-x for small x (expm1 in the reference) is avoided with a degree-6 series below 1/8.
__device__ __forceinline__ float phi_abs_f32(float x) {
  const float xm = fmaxf(x, 1.0e-5f);
  const float e = ex2_approx(xm * -1.4426950408889634f);
  // 1 - e^-x = x(1 - x/2 + x^2/6 - x^3/24 + x^4/120 - x^5/720), |rel err| < 1e-9 for x < 1/8
  float q = fmaf(xm, -1.0f / 720.0f, 1.0f / 120.0f);
  q = fmaf(xm, q, -1.0f / 24.0f);
  q = fmaf(xm, q, 1.0f / 6.0f);
  q = fmaf(xm, q, -0.5f);
  q = fmaf(xm, q, 1.0f);
  const float den = xm < 0.125f ? xm * q : 1.0f - e;
  const float num = 2.0f - den; // 1 + e^-x
  const float r = 0.6931471805599453f * (lg2_approx(num) - lg2_approx(den));
  return xm > 5.0f ? e + e : r;
}

// ---------------------------------------------------------------- per-word traits
// A "word" is 32 bits of a message vector: one float, or two halves.
struct TraitsF32 {
  static constexpr int kFramesPerWord = 1;
  static constexpr int kFramesPerVec = 4;
  static constexpr uint32_t kSignMask = 0x80000000u;
  static constexpr bool kUsesTable = false;
  // software prefetch of the next item's indices: measured on B200 (scripts/kernel_perf.py), it only pays in the
  // fp16 variable-node pass (+3-4 %); elsewhere the extra registers cost a resident CTA
  static constexpr bool kPrefetchCN = LDPC_PREFETCH_F32_CN, kPrefetchVN = LDPC_PREFETCH_F32_VN;

  __device__ static __forceinline__ uint32_t abs_w(uint32_t w) { return w & 0x7FFFFFFFu; }
  __device__ static __forceinline__ uint32_t add_w(uint32_t a, uint32_t b) {
    return __float_as_uint(__uint_as_float(a) + __uint_as_float(b));
  }
  __device__ static __forceinline__ uint32_t sub_w(uint32_t a, uint32_t b) {
    return __float_as_uint(__uint_as_float(a) - __uint_as_float(b));
  }
  // phi_abs of a word that may be negative (clamped like the reference's fmax)
  __device__ static __forceinline__ uint32_t phi_abs_w(uint32_t w, const uint16_t *, uint32_t) {
    return __float_as_uint(phi_abs_f32(__uint_as_float(w)));
  }
  // magnitude r (any sign bit) with the sign of s
  __device__ static __forceinline__ uint32_t with_sign(uint32_t r, uint32_t s) {
    return (r & ~kSignMask) | (s & kSignMask);
  }
  // sign word with bit set where the frame's bit is 1 (bits = low kFramesPerWord bits)
  __device__ static __forceinline__ uint32_t sign_from_bits(uint32_t bits) { return bits << 31; }
  // all-ones in the lanes of the word whose frame bit is set
  __device__ static __forceinline__ uint32_t lane_mask(uint32_t bits) { return 0u - (bits & 1u); }
  // hard decision bits of a posterior word: bit=1 <=> sign bit clear (flood.cu:180)
  __device__ static __forceinline__ uint32_t hard_bits(uint32_t w) { return (~w) >> 31; }
  // channel value -> llr (flood.cu:58, :73)
  __device__ static __forceinline__ uint32_t to_llr(uint32_t raw, int channel, float param) {
    const float x = __uint_as_float(raw);
    if (channel == 1) return (__float_as_uint(param) & 0x7FFFFFFFu) | (raw & 0x80000000u);
    if (channel == 0) return __float_as_uint(x * param);
    return raw;
  }
};

struct TraitsF16 {
  static constexpr int kFramesPerWord = 2;
  static constexpr int kFramesPerVec = 8;
  static constexpr uint32_t kSignMask = 0x80008000u;
  static constexpr bool kUsesTable = true;
  static constexpr bool kPrefetchCN = LDPC_PREFETCH_F16_CN, kPrefetchVN = true;

  __device__ static __forceinline__ __half2 h2(uint32_t w) { return *reinterpret_cast<__half2 *>(&w); }
  __device__ static __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
  __device__ static __forceinline__ uint32_t abs_w(uint32_t w) { return u32(__habs2(h2(w))); } // clears both sign bits
  // one IEEE half rounding per add/sub, as the reference's scalar __half arithmetic (flood.cu:95-99, 134-147)
  __device__ static __forceinline__ uint32_t add_w(uint32_t a, uint32_t b) { return u32(__hadd2_rn(h2(a), h2(b))); }
  __device__ static __forceinline__ uint32_t sub_w(uint32_t a, uint32_t b) { return u32(__hsub2_rn(h2(a), h2(b))); }
  // `last` = index from which every table entry is zero (2e^-x underflows in half above x ~ 17.3): only
  // [0, last] is staged in shared memory, larger indices clamp to it.
  // The argument is never negative: the check-node pass looks up S - |m_e| with S = the round-to-nearest sum of all
  // |m_e| (monotone rounding: S >= every term), the variable-node pass |T - m_e|.  Non-negative halves order like their
  // bit patterns, so the clamp is one packed half min (a NaN, which only a NaN input can produce, clamps to `last`);
  // doubling both indices at once (2*last < 2^16: no carry) gives the two byte offsets into the table.
  __device__ static __forceinline__ uint32_t phi_abs_w(uint32_t w, const uint16_t *tab, uint32_t last) {
#if LDPC_HMIN_CLAMP
    const uint32_t c = u32(__hmin2(h2(w), h2(last * 0x10001u)));
    const uint32_t c2 = c + c;
    const uint8_t *tb = reinterpret_cast<const uint8_t *>(tab);
    const uint32_t lo = *reinterpret_cast<const uint16_t *>(tb + (c2 & 0xFFFFu));
    const uint32_t hi = *reinterpret_cast<const uint16_t *>(tb + (c2 >> 16));
    return __byte_perm(lo, hi, 0x5410);
#else
    // first version: negative / NaN preimages clamp to the table's first entry (flood.cu:10, 26)
    const uint32_t c = u32(__hmax2(h2(w), h2(0u)));
    const uint32_t lo = tab[min(c & 0x7FFFu, last)];
    const uint32_t hi = tab[min(c >> 16, last)];
#endif
    return lo | (hi << 16);
  }
  // table entries are non-negative halves: nothing to clear in r
  __device__ static __forceinline__ uint32_t with_sign(uint32_t r, uint32_t s) { return r | (s & kSignMask); }
  __device__ static __forceinline__ uint32_t sign_from_bits(uint32_t bits) {
    return ((bits & 1u) << 15) | ((bits & 2u) << 30);
  }
  __device__ static __forceinline__ uint32_t lane_mask(uint32_t bits) {
    return ((bits & 1u) | ((bits & 2u) << 15)) * 0xFFFFu;
  }
  __device__ static __forceinline__ uint32_t hard_bits(uint32_t w) {
    const uint32_t n = ~w;
    return ((n >> 15) & 1u) | ((n >> 30) & 2u);
  }
  __device__ static __forceinline__ uint16_t to_llr(uint16_t raw, int channel, float param) {
    const __half ph = __float2half_rn(param); // float -> llr_t kernel argument
    const uint16_t pb = __half_as_ushort(ph);
    if (channel == 1) return (pb & 0x7FFFu) | (raw & 0x8000u);
    if (channel == 0) return __half_as_ushort(__hmul(__ushort_as_half(raw), ph));
    return raw;
  }
};

// reference fp16 phi_abs (flood.cu:21-29), used once to fill the table
__device__ __forceinline__ __half phi_abs_f16_exact(__half x) {
  const __half limit = __ushort_as_half(0x4500); // 5
  const __half pre = __ushort_as_half(0x003f);
  const __half two = __ushort_as_half(0x4000);
  const __half half_one = __ushort_as_half(0x3800);
  const __half xm = (x > pre) ? x : pre;
  return (xm > limit) ? __hmul(two, hexp(__hneg(xm))) : __hneg(hlog(htanh(__hmul(xm, half_one))));
}

__global__ void build_phi_table_kernel(uint16_t *table) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kPhiTableSize) table[i] = __half_as_ushort(phi_abs_f16_exact(__ushort_as_half(uint16_t(i))));
}

__global__ void phi_abs_f32_kernel(const float *x, float *y, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = phi_abs_f32(x[i]);
}

template <class Tr>
__device__ __forceinline__ const uint16_t *stage_phi_table(const uint16_t *g_table, uint32_t last) {
  if (Tr::kUsesTable) {
    extern __shared__ __align__(16) uint16_t s_table[];
    const uint4 *src = reinterpret_cast<const uint4 *>(g_table);
    uint4 *dst = reinterpret_cast<uint4 *>(s_table);
    for (int i = threadIdx.x; i < int(last / 8 + 1); i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    return s_table;
  }
  return nullptr;
}

// control block handed to the streaming kernels: ctl[0] = done flag, ctl[1] = log2 of the active vectors per row
// (0 = all of them), ctl[4] = "this period's refill is fused into its first check-node pass" (SchedState::fused_now)
__device__ __forceinline__ uint32_t active_log2(const uint32_t *ctl, uint32_t log2_g) {
  const uint32_t a = ctl ? ctl[1] : 0u;
  return a ? a : log2_g;
}
// Fused refill is chosen per check period ON THE DEVICE (the scheduler knows how many frames it loads): the host
// enqueues both variants of the refill and of the period's first check-node pass, and the variant that was not chosen
// exits at its first instruction.  gate: 0 = always run, 1 = run only in a fused period, 2 = only in a plain period.
enum : uint32_t { kGateNone = 0, kGateFused = 1, kGatePlain = 2 };
constexpr int kCtlFused = 4;
__device__ __forceinline__ bool gated_off(const uint32_t *ctl, uint32_t gate) {
  if (gate == kGateNone || !ctl) return false;
  return (ctl[kCtlFused] != 0u) != (gate == kGateFused);
}

__device__ __forceinline__ uint32_t &word(uint4 &v, int i) { return reinterpret_cast<uint32_t *>(&v)[i]; }
__device__ __forceinline__ uint32_t word(const uint4 &v, int i) { return reinterpret_cast<const uint32_t *>(&v)[i]; }

// bits of this thread's frames out of a bit-sliced row (row_bytes = &row[0] as bytes)
template <class Tr> __device__ __forceinline__ uint32_t load_frame_bits(const uint8_t *row_bytes, uint32_t g) {
  if (Tr::kFramesPerVec == 8) return row_bytes[g];
  return (row_bytes[g >> 1] >> ((g & 1u) * 4u)) & 0xFu;
}

// ---------------------------------------------------------------- output vectors of the two passes
// check node:    o = phi_abs(S - |m|) with sign(m) xor parity          (flood.cu:100-113)
// variable node: o = phi_abs(|T - m|) with the sign of T - m           (flood.cu:141-155)
template <class Tr>
__device__ __forceinline__ uint4 cn_out_vec(const uint4 &S, const uint4 &m, const uint4 &X, const uint16_t *tab,
                                            uint32_t tab_last) {
  uint4 o;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const uint32_t mw = word(m, w);
    const uint32_t r = Tr::phi_abs_w(Tr::sub_w(word(S, w), Tr::abs_w(mw)), tab, tab_last);
    word(o, w) = Tr::with_sign(r, mw ^ word(X, w));
  }
  return o;
}

template <class Tr>
__device__ __forceinline__ uint4 vn_out_vec(const uint4 &T, const uint4 &m, const uint16_t *tab, uint32_t tab_last) {
  uint4 o;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const uint32_t pre = Tr::sub_w(word(T, w), word(m, w));
    const uint32_t r = Tr::phi_abs_w(Tr::abs_w(pre), tab, tab_last);
    word(o, w) = Tr::with_sign(r, pre);
  }
  return o;
}

// ---------------------------------------------------------------- degree dispatch
// A row of degree <= kMaxRegDeg is processed by code specialised on its exact degree (no per-edge `k < deg` branch):
// run<D>(deg, f) calls f(std::integral_constant<int, deg>) for deg in [1, D].  All threads of a warp share the row,
// so the if-chain is uniform.  Measured on B200: fp32 variable-node pass +4.7 %, fp16 +0.9 %.
template <int D> struct DegDispatch {
  template <class F> __device__ static __forceinline__ void run(uint32_t deg, F &&f) {
    if (deg == uint32_t(D)) f(std::integral_constant<int, D>{});
    else DegDispatch<D - 1>::run(deg, f);
  }
};
template <> struct DegDispatch<0> {
  template <class F> __device__ static __forceinline__ void run(uint32_t, F &&) {}
};

// ---------------------------------------------------------------- check-node pass
// One thread = (check c, frame vector g).  S = sum |m_e|; every edge gets phi_abs(S - |m_e|) with
// sign = sign(m_e) xor parity, parity = syndrome bit xor (number of non-negative inputs) (flood.cu:77-115).
// The row offsets and syndrome bits of the thread's NEXT item are fetched while the current row's messages are
// in flight, so the dependent index -> message load chain is off the critical path.
//
// kRefill (cn_update_refill_kernel, the first check-node pass after a refill): the refill kernel has only written the
// new frames' PRIORS; the messages a new frame starts from, phi(prior of the edge's variable) (flood.cu:297-323), are
// produced here, in the pass that reads and rewrites every message row anyway -- a gather of the prior vector per edge
// for the vectors that hold new slots instead of a separate read-modify-write of those message vectors.
struct CnRefillArgs {
  const uint4 *llr0;         // priors [N][P]
  const uint32_t *edge_var;  // variable of each check-major edge
  const uint8_t *new_mask;   // bit per slot: refilled since the last check-node pass
};

// message vector a new-frame lane starts from: phi_abs(|prior|) with the prior's sign
template <class Tr>
__device__ __forceinline__ uint4 start_message(const uint4 &prior, const uint16_t *tab, uint32_t tab_last) {
  uint4 o;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const uint32_t pw = word(prior, w);
    const uint32_t r = Tr::phi_abs_w(Tr::abs_w(pw), tab, tab_last);
    word(o, w) = Tr::with_sign(r, pw);
  }
  return o;
}

template <class Tr, int kMaxRegDeg, bool kRefill, class Idx>
__device__ __forceinline__ void
cn_update_body(uint4 *__restrict__ msg, const uint32_t *__restrict__ row_ptr, const uint8_t *__restrict__ syn,
               const uint16_t *__restrict__ phi_table, uint32_t tab_last, uint32_t n_checks, uint32_t log2_g,
               const uint32_t *__restrict__ done, const CnRefillArgs ra, uint32_t gate) {
  if (done && *done) return; // the device-side scheduler has retired every frame: run-ahead launches are no-ops
  if (gated_off(done, gate)) return;
  constexpr uint32_t kFull = (1u << Tr::kFramesPerVec) - 1u;
  const uint16_t *tab = stage_phi_table<Tr>(phi_table, tab_last);
  const uint32_t G = 1u << log2_g; // row stride in vectors
  // active vectors per row: the scheduler may have compacted the live frames into the low slots (early retire)
  const uint32_t log2_ga = active_log2(done, log2_g);
  const uint32_t GA = 1u << log2_ga;
  const size_t syn_row_bytes = (size_t(G) * Tr::kFramesPerVec) >> 3;
  const uint64_t total = uint64_t(n_checks) << log2_ga;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  uint64_t item = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t a_next = 0, b_next = 0, sbits_next = 0;
  if (Tr::kPrefetchCN && item < total) {
    const uint32_t c = uint32_t(item >> log2_ga);
    a_next = __ldg(row_ptr + c);
    b_next = __ldg(row_ptr + c + 1);
    sbits_next = load_frame_bits<Tr>(syn + size_t(c) * syn_row_bytes, uint32_t(item) & (GA - 1u));
  }
  for (; item < total; item += stride) {
    const uint32_t g = uint32_t(item) & (GA - 1u);
    if (!Tr::kPrefetchCN) {
      const uint32_t c = uint32_t(item >> log2_ga);
      a_next = __ldg(row_ptr + c);
      b_next = __ldg(row_ptr + c + 1);
      sbits_next = load_frame_bits<Tr>(syn + size_t(c) * syn_row_bytes, g);
    }
    const uint32_t a = a_next, deg = b_next - a_next, sbits = sbits_next;
    uint4 *p = msg + ((Idx(a) << log2_g) + g);
    // parity seed: syndrome bit xor (deg & 1); xor-ing the raw sign bits of all inputs then gives
    // syndrome ^ #(non-negative inputs)
    const uint32_t odd = (deg & 1u) ? Tr::kSignMask : 0u;
    uint4 S = make_uint4(0, 0, 0, 0), X;
#pragma unroll
    for (int w = 0; w < 4; ++w)
      word(X, w) = Tr::sign_from_bits(sbits >> (w * Tr::kFramesPerWord)) ^ odd;
    // slots of this vector refilled since the last pass, and their lanes
    uint32_t nm = 0;
    uint4 SEL = make_uint4(0, 0, 0, 0);
    if (kRefill) {
      nm = load_frame_bits<Tr>(ra.new_mask, g);
#pragma unroll
      for (int w = 0; w < 4; ++w) word(SEL, w) = Tr::lane_mask(nm >> (w * Tr::kFramesPerWord));
    }
    if (deg <= kMaxRegDeg) {
      DegDispatch<kMaxRegDeg>::run(deg, [&](auto deg_c) {
        constexpr int DEG = decltype(deg_c)::value;
        uint4 m[DEG];
#pragma unroll
        for (int k = 0; k < DEG; ++k)
          m[k] = (kRefill && nm == kFull) ? make_uint4(0, 0, 0, 0) : ld_stream(p + (Idx(k) << log2_g));
        if (kRefill && nm) {
          uint4 pr[DEG];
#pragma unroll
          for (int k = 0; k < DEG; ++k)
            pr[k] = ld_stream(ra.llr0 + ((Idx(__ldg(ra.edge_var + a + k)) << log2_g) + g));
#pragma unroll
          for (int k = 0; k < DEG; ++k) {
            const uint4 o = start_message<Tr>(pr[k], tab, tab_last);
#pragma unroll
            for (int w = 0; w < 4; ++w) word(m[k], w) = (word(m[k], w) & ~word(SEL, w)) | (word(o, w) & word(SEL, w));
          }
        }
        if (Tr::kPrefetchCN && item + stride < total) { // next item's indices behind the message loads
          const uint32_t c = uint32_t((item + stride) >> log2_ga);
          a_next = __ldg(row_ptr + c);
          b_next = __ldg(row_ptr + c + 1);
          sbits_next = load_frame_bits<Tr>(syn + size_t(c) * syn_row_bytes, g);
        }
#pragma unroll
        for (int k = 0; k < DEG; ++k) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            word(S, w) = Tr::add_w(word(S, w), Tr::abs_w(word(m[k], w)));
            word(X, w) ^= word(m[k], w);
          }
        }
#pragma unroll
        for (int k = 0; k < DEG; ++k) {
          const uint4 o = cn_out_vec<Tr>(S, m[k], X, tab, tab_last);
          st_stream(p + (Idx(k) << log2_g), o);
        }
      });
    } else {
      if (Tr::kPrefetchCN && item + stride < total) {
        const uint32_t c = uint32_t((item + stride) >> log2_ga);
        a_next = __ldg(row_ptr + c);
        b_next = __ldg(row_ptr + c + 1);
        sbits_next = load_frame_bits<Tr>(syn + size_t(c) * syn_row_bytes, g);
      }
      // long rows: two passes; a refilled lane's start message is recomputed from the prior in both
      auto load_row = [&](uint32_t k) {
        uint4 mk = ld_cached(p + (Idx(k) << log2_g));
        if (kRefill && nm) {
          const uint4 o =
              start_message<Tr>(ld_cached(ra.llr0 + ((Idx(__ldg(ra.edge_var + a + k)) << log2_g) + g)), tab, tab_last);
#pragma unroll
          for (int w = 0; w < 4; ++w) word(mk, w) = (word(mk, w) & ~word(SEL, w)) | (word(o, w) & word(SEL, w));
        }
        return mk;
      };
      for (uint32_t k = 0; k < deg; ++k) {
        uint4 mk = load_row(k);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          word(S, w) = Tr::add_w(word(S, w), Tr::abs_w(word(mk, w)));
          word(X, w) ^= word(mk, w);
        }
      }
      for (uint32_t k = 0; k < deg; ++k) {
        uint4 mk = load_row(k);
        const uint4 o = cn_out_vec<Tr>(S, mk, X, tab, tab_last);
        st_stream(p + (Idx(k) << log2_g), o);
      }
    }
  }
}

// Idx = uint32_t when every vector index (E * vectors per row) fits 32 bits: one IMAD.WIDE per address instead of a
// 64-bit shift/add chain (the host picks; size_t otherwise, e.g. 2^13 slots of a 4M-bit code)
template <class Tr, int kMaxRegDeg, class Idx>
__global__ void __launch_bounds__(cn_threads_for(kMaxRegDeg), min_blocks_for(kMaxRegDeg) * kThreads / cn_threads_for(kMaxRegDeg))
    cn_update_kernel(uint4 *__restrict__ msg, const uint32_t *__restrict__ row_ptr, const uint8_t *__restrict__ syn,
                     const uint16_t *__restrict__ phi_table, uint32_t tab_last, uint32_t n_checks, uint32_t log2_g,
                     const uint32_t *__restrict__ done, uint32_t gate) {
  cn_update_body<Tr, kMaxRegDeg, false, Idx>(msg, row_ptr, syn, phi_table, tab_last, n_checks, log2_g, done, CnRefillArgs{},
                                             gate);
}

// first check-node pass after a refill (one launch per check period that loads frames): message and prior vectors of a
// row are both live.  A/B on B200 in whole bench steps (profiles/r06_ab_fused_refill.txt): 2 CTAs/SM (122 registers) 16.4
// ms per fused pass, 3 CTAs/SM (80 registers, prior gathers partly serialised) 15.2 ms, 4 CTAs/SM (64 registers) 17.3 ms.
#ifndef LDPC_CN_REFILL_MINBLOCKS
#define LDPC_CN_REFILL_MINBLOCKS 3
#endif
template <class Tr, int kMaxRegDeg, class Idx>
__global__ void __launch_bounds__(kThreads, LDPC_CN_REFILL_MINBLOCKS)
    cn_update_refill_kernel(uint4 *__restrict__ msg, const uint32_t *__restrict__ row_ptr,
                            const uint8_t *__restrict__ syn, const uint16_t *__restrict__ phi_table, uint32_t tab_last,
                            uint32_t n_checks, uint32_t log2_g, const uint32_t *__restrict__ done, const CnRefillArgs ra,
                            uint32_t gate) {
  cn_update_body<Tr, kMaxRegDeg, true, Idx>(msg, row_ptr, syn, phi_table, tab_last, n_checks, log2_g, done, ra, gate);
}

// ---------------------------------------------------------------- check-node pass, long rows (high-rate codes)
// Rate >= 0.8 codes have check degrees of 16-64.  Holding such a row in one thread's registers leaves too few
// resident threads, and the two-pass loop reads every message twice.  Here a row is split over kParts lanes of
// the same warp: lane part l keeps edges [8l, 8l+8) in registers (one read, one write per message), the sum of
// magnitudes is CARRIED from part to part with shuffles so the additions happen in exactly the reference's edge
// order (flood.cu:95-99; this keeps fp16 bit-exact), and the sign parity is an xor butterfly.
// Warp layout: lane = part * (32/kParts) + j, j indexing 32/kParts consecutive 16-byte frame vectors, so each
// part still issues full 128-byte (kParts = 4) row segments.
template <class Tr, int kParts>
__global__ void __launch_bounds__(kThreads, 3)
    cn_update_split_kernel(uint4 *__restrict__ msg, const uint32_t *__restrict__ row_ptr,
                           const uint8_t *__restrict__ syn, const uint16_t *__restrict__ phi_table, uint32_t tab_last,
                           uint32_t n_checks, uint32_t log2_g, const uint32_t *__restrict__ done, uint32_t gate) {
  if (done && *done) return;
  if (gated_off(done, gate)) return;
  constexpr int kChunk = 8;
  constexpr int kVecPerWarp = 32 / kParts;
  const uint16_t *tab = stage_phi_table<Tr>(phi_table, tab_last);
  const uint32_t G = 1u << log2_g;
  const uint32_t GA = 1u << active_log2(done, log2_g);
  const size_t syn_row_bytes = (size_t(G) * Tr::kFramesPerVec) >> 3;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t part = lane / kVecPerWarp, j = lane % kVecPerWarp;
  const uint32_t groups_per_row = GA / kVecPerWarp; // active G >= kVecPerWarp (host guarantees)
  const uint64_t total = uint64_t(n_checks) * groups_per_row; // warp-items
  const uint64_t warp_stride = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  for (uint64_t wi = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; wi < total; wi += warp_stride) {
    const uint32_t c = uint32_t(wi / groups_per_row);
    const uint32_t g = uint32_t(wi - uint64_t(c) * groups_per_row) * kVecPerWarp + j;
    const uint32_t a = __ldg(row_ptr + c), deg = __ldg(row_ptr + c + 1) - a;
    const uint32_t k0 = part * kChunk; // this lane's edges: [k0, min(deg, k0 + 8))
    uint4 *p = msg + (size_t(a + k0) << log2_g) + g;
    uint4 m[kChunk];
#pragma unroll
    for (int k = 0; k < kChunk; ++k)
      if (k0 + k < deg) m[k] = ld_stream(p + (size_t(k) << log2_g));
    const uint32_t sbits = load_frame_bits<Tr>(syn + size_t(c) * syn_row_bytes, g);
    const uint32_t odd = (deg & 1u) ? Tr::kSignMask : 0u;
    uint4 S = make_uint4(0, 0, 0, 0), X = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < kChunk; ++k)
      if (k0 + k < deg) {
#pragma unroll
        for (int w = 0; w < 4; ++w) word(X, w) ^= word(m[k], w);
      }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int d = kVecPerWarp; d < 32; d <<= 1) word(X, w) ^= __shfl_xor_sync(0xFFFFFFFFu, word(X, w), d);
      word(X, w) ^= Tr::sign_from_bits(sbits >> (w * Tr::kFramesPerWord)) ^ odd;
    }
    // carried sum: part 0 adds its magnitudes, hands S to part 1, ... -> sequential order of the reference
#pragma unroll
    for (int q = 0; q < kParts; ++q) {
      if (part == uint32_t(q)) {
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k0 + k < deg) {
#pragma unroll
            for (int w = 0; w < 4; ++w) word(S, w) = Tr::add_w(word(S, w), Tr::abs_w(word(m[k], w)));
          }
      }
#pragma unroll
      for (int w = 0; w < 4; ++w) word(S, w) = __shfl_sync(0xFFFFFFFFu, word(S, w), q * kVecPerWarp + j);
    }
#pragma unroll
    for (int k = 0; k < kChunk; ++k)
      if (k0 + k < deg) {
        const uint4 o = cn_out_vec<Tr>(S, m[k], X, tab, tab_last);
        st_stream(p + (size_t(k) << log2_g), o);
      }
  }
}

// ---------------------------------------------------------------- check-node pass, bulk-async (TMA) staged  [EXPERIMENT]
// Same arithmetic as cn_update_kernel, but the row slices of a tile (one check x 256 frame vectors = D x 4 KB) are
// brought into shared memory by the bulk-copy engine (cp.async.bulk + mbarrier complete_tx), updated in place by the
// 256 threads and written back with cp.async.bulk (bulk_group), three tiles in flight per CTA.  Selected with
// LDPC_B200_CN_BULK=1; see DESIGN.md for the measurement that decides whether it replaces the register kernel.
namespace bulk {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 24)) __trap(); // never hang the GPU on a protocol bug
}
__device__ __forceinline__ void load_bulk(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void store_bulk(void *gmem_dst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
} // namespace bulk

constexpr int kBulkStages = 3;
constexpr int kBulkTileVec = 256; // vectors (threads) per tile
template <class Tr, int kMaxRegDeg>
__global__ void __launch_bounds__(kBulkTileVec, 2)
    cn_update_bulk_kernel(uint4 *__restrict__ msg, const uint32_t *__restrict__ row_ptr, const uint8_t *__restrict__ syn,
                          const uint16_t *__restrict__ phi_table, uint32_t tab_last, uint32_t n_checks, uint32_t log2_g,
                          const uint32_t *__restrict__ done, uint32_t gate) {
  if (done && *done) return;
  if (gated_off(done, gate)) return;
  extern __shared__ __align__(16) uint16_t s_table[]; // [table | stages | barriers]
  const uint32_t table_bytes = Tr::kUsesTable ? (tab_last / 8 + 1) * 16 : 0;
  uint4 *stages = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(s_table) + table_bytes);
  uint64_t *full = reinterpret_cast<uint64_t *>(stages + size_t(kBulkStages) * kMaxRegDeg * kBulkTileVec);
  const uint16_t *tab = stage_phi_table<Tr>(phi_table, tab_last);
  const uint32_t tid = threadIdx.x;
  const uint32_t G = 1u << log2_g;
  const uint32_t GA = 1u << active_log2(done, log2_g);
  const size_t syn_row_bytes = (size_t(G) * Tr::kFramesPerVec) >> 3;
  const uint32_t tiles_per_row = GA / kBulkTileVec;
  const uint64_t total = uint64_t(n_checks) * tiles_per_row;
  if (tid == 0)
    for (int sidx = 0; sidx < kBulkStages; ++sidx) bulk::mbar_init(full + sidx, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  auto tile_of = [&](uint64_t i) { return uint64_t(blockIdx.x) + i * gridDim.x; };
  auto issue = [&](uint64_t i) { // thread 0: bring tile i's rows into stage i % kBulkStages
    const uint64_t t = tile_of(i);
    if (t >= total) return;
    const uint32_t c = uint32_t(t / tiles_per_row), gb = uint32_t(t - uint64_t(c) * tiles_per_row);
    const uint32_t a = __ldg(row_ptr + c), deg = __ldg(row_ptr + c + 1) - a;
    const int sidx = int(i % kBulkStages);
    bulk::mbar_expect_tx(full + sidx, deg * kBulkTileVec * 16u);
    for (uint32_t k = 0; k < deg; ++k)
      bulk::load_bulk(stages + (size_t(sidx) * kMaxRegDeg + k) * kBulkTileVec,
                      msg + (size_t(a + k) << log2_g) + size_t(gb) * kBulkTileVec, kBulkTileVec * 16u, full + sidx);
  };
  if (tid == 0)
    for (int i = 0; i < kBulkStages - 1; ++i) issue(uint64_t(i));
  for (uint64_t i = 0; tile_of(i) < total; ++i) {
    const uint64_t t = tile_of(i);
    const uint32_t c = uint32_t(t / tiles_per_row), gb = uint32_t(t - uint64_t(c) * tiles_per_row);
    const uint32_t g = gb * kBulkTileVec + tid;
    const uint32_t a = __ldg(row_ptr + c), deg = __ldg(row_ptr + c + 1) - a;
    const uint32_t sbits = load_frame_bits<Tr>(syn + size_t(c) * syn_row_bytes, g);
    const int sidx = int(i % kBulkStages);
    uint4 *tile = stages + size_t(sidx) * kMaxRegDeg * kBulkTileVec;
    bulk::mbar_wait(full + sidx, uint32_t((i / kBulkStages) & 1));
    const uint32_t odd = (deg & 1u) ? Tr::kSignMask : 0u;
    uint4 S = make_uint4(0, 0, 0, 0), X;
#pragma unroll
    for (int w = 0; w < 4; ++w) word(X, w) = Tr::sign_from_bits(sbits >> (w * Tr::kFramesPerWord)) ^ odd;
    uint4 m[kMaxRegDeg];
#pragma unroll
    for (int k = 0; k < kMaxRegDeg; ++k)
      if (k < deg) m[k] = tile[size_t(k) * kBulkTileVec + tid];
#pragma unroll
    for (int k = 0; k < kMaxRegDeg; ++k)
      if (k < deg) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          word(S, w) = Tr::add_w(word(S, w), Tr::abs_w(word(m[k], w)));
          word(X, w) ^= word(m[k], w);
        }
      }
#pragma unroll
    for (int k = 0; k < kMaxRegDeg; ++k)
      if (k < deg) {
        const uint4 o = cn_out_vec<Tr>(S, m[k], X, tab, tab_last);
        tile[size_t(k) * kBulkTileVec + tid] = o;
      }
    bulk::fence_async_smem(); // generic-proxy writes -> visible to the bulk-copy (async) proxy
    __syncthreads();
    if (tid == 0) {
      for (uint32_t k = 0; k < deg; ++k)
        bulk::store_bulk(msg + (size_t(a + k) << log2_g) + size_t(gb) * kBulkTileVec,
                         tile + size_t(k) * kBulkTileVec, kBulkTileVec * 16u);
      bulk::store_commit();
      // the stage tile i-1 used may be refilled once its write-back has finished READING shared memory
      bulk::store_wait_read<1>();
      issue(i + kBulkStages - 1);
    }
  }
  if (tid == 0) bulk::store_wait_all();
}

// ---------------------------------------------------------------- variable-node pass
// One thread = (variable v, frame vector g).  T = llr0 + sum of incoming c2v in CSC order; every edge gets
// phi(T - m_e); optionally the hard decision signbit(T)==0 is written bit-sliced (flood.cu:117-189).
//
// Gather list: the reference reads in_bit_to_edge[v], then in_to_out_edge[...] (two DEPENDENT index loads,
// flood.cu:131-137).  Under a saturated memory system each of them is a long L2 round trip that a warp sits through
// with nothing else in flight (ncu, profiles/r05: more stall samples on the two index loads than on the message loads
// themselves).  Here the list is a padded row per variable ("ELL"): VnEll::kWords u32 = [degree, row of edge 0..D-1,
// ..., first CSC edge], ONE independent load level, fetched for the thread's NEXT item right after the current item's
// message loads are issued, so it arrives with them.  A degree above kMaxRegDeg marks a long variable: its list is
// read from col_ptr / in_to_out (the last word = col_ptr[v]).
template <int kMaxRegDeg> struct VnEll {
  static constexpr int kWords = kMaxRegDeg <= 6 ? 8 : 12;
};

template <int kMaxRegDeg>
__device__ __forceinline__ void vn_fetch_list(const uint32_t *__restrict__ vn_ell, uint32_t v, uint32_t &deg,
                                              uint32_t &first_edge, uint32_t (&rows)[kMaxRegDeg]) {
  constexpr int W = VnEll<kMaxRegDeg>::kWords;
  const uint4 *e = reinterpret_cast<const uint4 *>(vn_ell) + size_t(v) * (W / 4);
  uint32_t w[W];
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    const uint4 x = __ldg(e + q);
    w[4 * q] = x.x, w[4 * q + 1] = x.y, w[4 * q + 2] = x.z, w[4 * q + 3] = x.w;
  }
  deg = w[0];
  first_edge = w[W - 1];
#pragma unroll
  for (int k = 0; k < kMaxRegDeg; ++k) rows[k] = w[1 + k];
}

template <class Tr, bool kFinalBits, int kMaxRegDeg, class Idx>
__global__ void __launch_bounds__(vn_threads_for(kMaxRegDeg), min_blocks_for(kMaxRegDeg) * kThreads / vn_threads_for(kMaxRegDeg))
    vn_update_kernel(uint4 *__restrict__ msg, const uint4 *__restrict__ llr0, const uint32_t *__restrict__ vn_ell,
                     const uint32_t *__restrict__ in_to_out, uint8_t *__restrict__ hard,
                     const uint16_t *__restrict__ phi_table, uint32_t tab_last, uint32_t n_vars, uint32_t log2_g,
                     const uint32_t *__restrict__ done) {
  if (done && *done) return;
  const uint16_t *tab = stage_phi_table<Tr>(phi_table, tab_last);
  const uint32_t G = 1u << log2_g;
  const uint32_t log2_ga = active_log2(done, log2_g);
  const uint32_t GA = 1u << log2_ga;
  const size_t hard_row_bytes = (size_t(G) * Tr::kFramesPerVec) >> 3;
  const uint64_t total = uint64_t(n_vars) << log2_ga;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  uint64_t item = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t a_next = 0, deg_next = 0;
  uint32_t rows_next[kMaxRegDeg];
  if (Tr::kPrefetchVN && item < total) vn_fetch_list<kMaxRegDeg>(vn_ell, uint32_t(item >> log2_ga), deg_next, a_next, rows_next);
  for (; item < total; item += stride) {
    const uint32_t v = uint32_t(item >> log2_ga);
    const uint32_t g = uint32_t(item) & (GA - 1u);
    if (!Tr::kPrefetchVN) vn_fetch_list<kMaxRegDeg>(vn_ell, v, deg_next, a_next, rows_next);
    const uint32_t a = a_next, deg = deg_next;
    uint4 T = ld_stream(llr0 + ((Idx(v) << log2_g) + g));
    if (deg <= kMaxRegDeg) {
      if (deg == 0 && Tr::kPrefetchVN && item + stride < total) // isolated variable: still fetch the next item
        vn_fetch_list<kMaxRegDeg>(vn_ell, uint32_t((item + stride) >> log2_ga), deg_next, a_next, rows_next);
      DegDispatch<kMaxRegDeg>::run(deg, [&](auto deg_c) {
        constexpr int DEG = decltype(deg_c)::value;
        uint4 m[DEG];
        Idx at[DEG]; // vector index of each edge row: one IMAD.WIDE away from its address, for the load and the store
#pragma unroll
        for (int k = 0; k < DEG; ++k) {
          at[k] = (Idx(rows_next[k]) << log2_g) + g;
          m[k] = ld_stream(msg + at[k]);
        }
        if (Tr::kPrefetchVN && item + stride < total) // next item's gather list behind the message loads
          vn_fetch_list<kMaxRegDeg>(vn_ell, uint32_t((item + stride) >> log2_ga), deg_next, a_next, rows_next);
#pragma unroll
        for (int k = 0; k < DEG; ++k) {
#pragma unroll
          for (int w = 0; w < 4; ++w) word(T, w) = Tr::add_w(word(T, w), word(m[k], w));
        }
#pragma unroll
        for (int k = 0; k < DEG; ++k) {
          const uint4 o = vn_out_vec<Tr>(T, m[k], tab, tab_last);
          st_stream(msg + at[k], o);
        }
      });
    } else {
      if (Tr::kPrefetchVN && item + stride < total)
        vn_fetch_list<kMaxRegDeg>(vn_ell, uint32_t((item + stride) >> log2_ga), deg_next, a_next, rows_next);
      for (uint32_t k = 0; k < deg; ++k) {
        const uint4 mk = ld_cached(msg + ((Idx(__ldg(in_to_out + a + k)) << log2_g) + g));
#pragma unroll
        for (int w = 0; w < 4; ++w) word(T, w) = Tr::add_w(word(T, w), word(mk, w));
      }
      for (uint32_t k = 0; k < deg; ++k) {
        uint4 *pk = msg + ((Idx(__ldg(in_to_out + a + k)) << log2_g) + g);
        const uint4 mk = ld_cached(pk);
        const uint4 o = vn_out_vec<Tr>(T, mk, tab, tab_last);
        st_stream(pk, o);
      }
    }
    if (kFinalBits) {
      uint32_t bits = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) bits |= Tr::hard_bits(word(T, w)) << (w * Tr::kFramesPerWord);
      uint8_t *row = hard + size_t(v) * hard_row_bytes;
#if LDPC_MERGE_DECISION_STORES
      // The lanes of 32 consecutive slots (4 threads in fp16, 8 in fp32: always the same variable, a row holds at least
      // 32 slots) merge their decision bits into ONE 32-bit store with shuffles instead of four one-byte stores.
      if (Tr::kFramesPerVec == 8) {
        uint32_t b = bits | (__shfl_down_sync(0xFFFFFFFFu, bits, 1) << 8);
        b |= __shfl_down_sync(0xFFFFFFFFu, b, 2) << 16;
        if ((g & 3u) == 0) *reinterpret_cast<uint32_t *>(row + g) = b;
      } else {
        uint32_t b = bits | (__shfl_down_sync(0xFFFFFFFFu, bits, 1) << 4);
        b |= __shfl_down_sync(0xFFFFFFFFu, b, 2) << 8;
        b |= __shfl_down_sync(0xFFFFFFFFu, b, 4) << 16;
        if ((g & 7u) == 0) *reinterpret_cast<uint32_t *>(row + (g >> 1)) = b;
      }
#else
      if (Tr::kFramesPerVec == 8) {
        row[g] = uint8_t(bits);
      } else {
        // two neighbouring threads (g even/odd, same variable) share a byte
        const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, bits, 1);
        if ((g & 1u) == 0) row[g >> 1] = uint8_t(bits | (other << 4));
      }
#endif
    }
  }
}

// ---------------------------------------------------------------- variable-node pass, bulk-async (TMA) staged
// Same arithmetic as vn_update_kernel, but a tile's rows (one variable x 256 frame vectors: the prior row and its
// deg message rows, 4 KB each, wherever they lie) are gathered into shared memory by the bulk-copy engine
// (cp.async.bulk + mbarrier complete_tx), updated in place by 256 threads and scattered back with cp.async.bulk
// (bulk_group).  The copy engine moves a whole 4 KB row piece with one request instead of eight 512-byte warp loads,
// and the threads never wait on global memory.
// One CTA = 512 threads = TWO independent halves of 256 threads, each with its own 3-stage tile pipeline, mbarriers and
// named barrier, sharing ONE copy of the phi table: (1 + D) x 4 KB x 3 stages x 2 halves + 39 KB table = 207 KB for
// D = 6 -- two separate CTAs with a table each would only fit two stages, i.e. no load/compute overlap inside a CTA.
// Needs every variable degree <= kMaxRegDeg and at least 256 vectors per row; selected by the host.
constexpr int kVnBulkStages = 3;
constexpr int kVnBulkThreads = 2 * kBulkTileVec;

template <class Tr, bool kFinalBits, int kMaxRegDeg>
__global__ void __launch_bounds__(kVnBulkThreads, 1)
    vn_update_bulk_kernel(uint4 *__restrict__ msg, const uint4 *__restrict__ llr0, const uint32_t *__restrict__ vn_ell,
                          const uint32_t *__restrict__, uint8_t *__restrict__ hard,
                          const uint16_t *__restrict__ phi_table, uint32_t tab_last, uint32_t n_vars, uint32_t log2_g,
                          const uint32_t *__restrict__ done) {
  if (done && *done) return;
  constexpr int kStages = kVnBulkStages;
  constexpr int kRows = kMaxRegDeg + 1; // prior + messages
  constexpr int kListWords = VnEll<kMaxRegDeg>::kWords;
  extern __shared__ __align__(16) uint16_t s_table[]; // [table | 2 x stages | 2 x lists | 2 x barriers]
  const uint32_t table_bytes = Tr::kUsesTable ? (tab_last / 8 + 1) * 16 : 0;
  const uint32_t half = threadIdx.x / kBulkTileVec, tid = threadIdx.x % kBulkTileVec;
  uint4 *stages_all = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(s_table) + table_bytes);
  uint4 *stages = stages_all + size_t(half) * kStages * kRows * kBulkTileVec;
  uint32_t *lists_all = reinterpret_cast<uint32_t *>(stages_all + size_t(2) * kStages * kRows * kBulkTileVec);
  uint32_t *lists = lists_all + half * kStages * kListWords;
  uint64_t *full = reinterpret_cast<uint64_t *>(lists_all + 2 * kStages * kListWords) + half * kStages;
  const uint16_t *tab = stage_phi_table<Tr>(phi_table, tab_last); // all 512 threads; ends with __syncthreads if staged
  const uint32_t G = 1u << log2_g;
  const uint32_t GA = 1u << active_log2(done, log2_g);
  const size_t hard_row_bytes = (size_t(G) * Tr::kFramesPerVec) >> 3;
  const uint32_t tiles_per_row = GA / kBulkTileVec;
  const uint64_t total = uint64_t(n_vars) * tiles_per_row;
  if (tid == 0)
    for (int sidx = 0; sidx < kStages; ++sidx) bulk::mbar_init(full + sidx, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  auto half_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1u + half), "n"(kBulkTileVec) : "memory"); };
  const uint64_t lane_tile = uint64_t(blockIdx.x) * 2 + half, tile_stride = uint64_t(gridDim.x) * 2;
  auto tile_of = [&](uint64_t i) { return lane_tile + i * tile_stride; };
  auto issue = [&](uint64_t i) { // elected thread: gather tile i's rows into stage i % kStages
    const uint64_t t = tile_of(i);
    if (t >= total) return;
    const uint32_t v = uint32_t(t / tiles_per_row), gb = uint32_t(t - uint64_t(v) * tiles_per_row);
    const int sidx = int(i % kStages);
    uint32_t *list = lists + sidx * kListWords;
    const uint4 *e = reinterpret_cast<const uint4 *>(vn_ell) + size_t(v) * (kListWords / 4);
#pragma unroll
    for (int q = 0; q < kListWords / 4; ++q) reinterpret_cast<uint4 *>(list)[q] = __ldg(e + q);
    const uint32_t deg = list[0];
    uint4 *tile = stages + size_t(sidx) * kRows * kBulkTileVec;
    bulk::mbar_expect_tx(full + sidx, (deg + 1) * kBulkTileVec * 16u); // release: the list is visible to the waiters
    bulk::load_bulk(tile, llr0 + (size_t(v) << log2_g) + size_t(gb) * kBulkTileVec, kBulkTileVec * 16u, full + sidx);
    for (uint32_t k = 0; k < deg; ++k)
      bulk::load_bulk(tile + size_t(1 + k) * kBulkTileVec, msg + (size_t(list[1 + k]) << log2_g) + size_t(gb) * kBulkTileVec,
                      kBulkTileVec * 16u, full + sidx);
  };
  if (tid == 0)
    for (int i = 0; i < kStages - 1; ++i) issue(uint64_t(i));
  for (uint64_t i = 0; tile_of(i) < total; ++i) {
    const uint64_t t = tile_of(i);
    const uint32_t v = uint32_t(t / tiles_per_row), gb = uint32_t(t - uint64_t(v) * tiles_per_row);
    const uint32_t g = gb * kBulkTileVec + tid;
    const int sidx = int(i % kStages);
    uint4 *tile = stages + size_t(sidx) * kRows * kBulkTileVec;
    const uint32_t *list = lists + sidx * kListWords;
    bulk::mbar_wait(full + sidx, uint32_t((i / kStages) & 1));
    const uint32_t deg = list[0];
    uint4 T = tile[tid];
    uint4 m[kMaxRegDeg];
#pragma unroll
    for (int k = 0; k < kMaxRegDeg; ++k)
      if (k < int(deg)) m[k] = tile[size_t(1 + k) * kBulkTileVec + tid];
#pragma unroll
    for (int k = 0; k < kMaxRegDeg; ++k)
      if (k < int(deg)) {
#pragma unroll
        for (int w = 0; w < 4; ++w) word(T, w) = Tr::add_w(word(T, w), word(m[k], w));
      }
#pragma unroll
    for (int k = 0; k < kMaxRegDeg; ++k)
      if (k < int(deg)) tile[size_t(1 + k) * kBulkTileVec + tid] = vn_out_vec<Tr>(T, m[k], tab, tab_last);
    if (kFinalBits) {
      uint32_t bits = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) bits |= Tr::hard_bits(word(T, w)) << (w * Tr::kFramesPerWord);
      uint8_t *row = hard + size_t(v) * hard_row_bytes;
      if (Tr::kFramesPerVec == 8) {
        uint32_t b = bits | (__shfl_down_sync(0xFFFFFFFFu, bits, 1) << 8);
        b |= __shfl_down_sync(0xFFFFFFFFu, b, 2) << 16;
        if ((g & 3u) == 0) *reinterpret_cast<uint32_t *>(row + g) = b;
      } else {
        uint32_t b = bits | (__shfl_down_sync(0xFFFFFFFFu, bits, 1) << 4);
        b |= __shfl_down_sync(0xFFFFFFFFu, b, 2) << 8;
        b |= __shfl_down_sync(0xFFFFFFFFu, b, 4) << 16;
        if ((g & 7u) == 0) *reinterpret_cast<uint32_t *>(row + (g >> 1)) = b;
      }
    }
    bulk::fence_async_smem(); // generic-proxy writes -> visible to the bulk-copy (async) proxy
    half_sync();
    if (tid == 0) {
      for (uint32_t k = 0; k < deg; ++k)
        bulk::store_bulk(msg + (size_t(list[1 + k]) << log2_g) + size_t(gb) * kBulkTileVec,
                         tile + size_t(1 + k) * kBulkTileVec, kBulkTileVec * 16u);
      bulk::store_commit();
      // the stage tile i-1 used (and its list) may be refilled once its write-back has finished READING shared memory
      bulk::store_wait_read<1>();
      issue(i + kStages - 1);
    }
  }
  if (tid == 0) bulk::store_wait_all();
}

// ---------------------------------------------------------------- syndrome check (bit-sliced)
// One thread = (check c, 32-slot word w): x = syn[c][w] xor hard[var(e)][w] over the row; any set bit marks
// that slot's frame as violated (flood.cu:191-223, 32 frames per XOR).
__global__ void __launch_bounds__(kThreads)
    syndrome_check_kernel(const uint32_t *__restrict__ syn, const uint32_t *__restrict__ hard,
                          const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ edge_var,
                          uint32_t *__restrict__ viol, uint32_t n_checks, uint32_t log2_pw,
                          const uint32_t *__restrict__ done) {
  if (done && *done) return;
  const uint32_t PW = 1u << log2_pw;
  const uint64_t total = uint64_t(n_checks) << log2_pw;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x; // multiple of PW: a thread keeps its word w
  uint32_t acc = 0;
  const uint64_t first = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t w = uint32_t(first) & (PW - 1u);
  for (uint64_t item = first; item < total; item += stride) {
    const uint32_t c = uint32_t(item >> log2_pw);
    const uint32_t a = __ldg(row_ptr + c), b = __ldg(row_ptr + c + 1);
    uint32_t x = __ldg(syn + (size_t(c) << log2_pw) + w);
    for (uint32_t e = a; e < b; ++e) x ^= __ldg(hard + (size_t(__ldg(edge_var + e)) << log2_pw) + w);
    acc |= x;
  }
  if (acc) atomicOr(viol + w, acc);
}

// ---------------------------------------------------------------- refill of freed slots
struct RefillEntry {     // one new frame
  uint64_t src_offset;   // element offset of its first channel value in the staging/input array
  uint64_t src_stride;   // element stride between consecutive variables
  uint64_t syn_offset;   // u32 offset of its packed syndrome (frame-major, S words)
  uint32_t slot;         // destination slot
  uint32_t pad;
};

// One thread = (16-byte slot vector g, every variable i of its stride).  New frames land in scattered slots, so a 2- or
// 4-byte store per (row, slot) would cost a 32-byte sector read-modify-write each.  Instead every vector that contains
// at least one new slot is rewritten whole: read 16 B, replace the new lanes, write 16 B -- for the prior row and, with
// kMsgRows, for every edge row of the variable.  prior = llr(channel value) (0 for punctured variables), edges get
// phi(prior) (flood.cu:47-75, 297-323).  Vectors without new slots are skipped; a vector of only new slots is not read.
// The grid stride is a multiple of the row length (host: grid_misc), so a thread keeps its vector g: the source
// addresses of its new lanes are looked up once, not once per variable.
// kMsgRows = false (decode(), fused path): only the priors are written here; the first check-node pass of the period
// (cn_update_refill_kernel) derives the start messages from them.
// kSrc: what the source array holds -- kSrcNative: llr elements of the decoder's precision; kSrcPackedBits: one bit
// per transmitted variable (extension f3); kSrcHalf (fp32 decoders only): IEEE halves, widened here -- "fp16 transfer
// with fp32 compute", the split the reference leaves open (h/common.h:35-37 sets transfer_llr_t = llr_t "for
// simplicity"): half the H2D bytes of the fp32 path, arithmetic unchanged.
enum : int { kSrcNative = 0, kSrcPackedBits = 1, kSrcHalf = 2 };
template <class Tr, int kSrc, bool kMsgRows>
__global__ void __launch_bounds__(kThreads)
    refill_kernel(uint4 *__restrict__ msg, uint4 *__restrict__ llr0, const void *__restrict__ src,
                  const RefillEntry *__restrict__ entries, const uint32_t *__restrict__ n_new_ptr,
                  const uint32_t *__restrict__ new_mask, const uint32_t *__restrict__ slot_entry,
                  const uint32_t *__restrict__ col_ptr, const uint32_t *__restrict__ in_to_out,
                  const uint16_t *__restrict__ phi_table, uint32_t n_vars, uint32_t n_regular, uint32_t log2_g,
                  int channel, float param, const uint32_t *__restrict__ ctl, uint32_t gate,
                  const uint32_t *__restrict__ var_pos) {
  // var_pos: storage row of each variable when the decoder keeps its arrays in a locality order (nullptr: identity);
  // col_ptr / in_to_out are then the storage-order tables
  constexpr bool packed_bits = kSrc == kSrcPackedBits;
  static_assert(kSrc != kSrcHalf || Tr::kFramesPerVec == 4, "half transfer elements are widened by fp32 decoders only");
  if (*n_new_ptr == 0) return; // written by the scheduler kernel (or the host for kernel-level calls)
  if (gated_off(ctl, gate)) return;
  constexpr int V = Tr::kFramesPerVec;
  constexpr uint32_t kFull = (1u << V) - 1u;
  const uint32_t G = 1u << log2_g;
  const uint64_t total = uint64_t(n_vars) << log2_g;
  const uint64_t first = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint32_t g = uint32_t(first) & (G - 1u);
  const uint32_t nm = load_frame_bits<Tr>(reinterpret_cast<const uint8_t *>(new_mask), g);
  if (nm == 0) return; // this thread's vector holds no new frame
  uint64_t off[V];     // element offset of each new lane's first channel value
  uint64_t src_stride = 0;
  uint4 SEL;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    off[j] = 0;
    if ((nm >> j) & 1u) {
      const RefillEntry en = entries[__ldg(slot_entry + g * V + j)];
      off[j] = en.src_offset;
      src_stride = en.src_stride; // the same for every frame of a launch
    }
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) word(SEL, w) = Tr::lane_mask(nm >> (w * Tr::kFramesPerWord));
  // kUnroll variables per trip (priors-only variant).  More than one per trip turned out SLOWER (profiles/
  // r06_ab_fused_refill.txt): the launch is bound by the sector traffic of its scattered source reads, and a wider
  // window of variables in flight only evicts the sectors neighbouring threads are about to use.
  constexpr int kUnroll = kMsgRows ? 1 : LDPC_REFILL_UNROLL;
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t item = first; item < total; item += kUnroll * stride) {
    uint4 L[kUnroll], old[kUnroll];
    uint32_t iv[kUnroll], vrow[kUnroll];
    bool valid[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t it = item + uint64_t(u) * stride;
      valid[u] = it < total;
      iv[u] = uint32_t(it >> log2_g);
      vrow[u] = (valid[u] && var_pos) ? __ldg(var_pos + iv[u]) : iv[u];
      old[u] = make_uint4(0, 0, 0, 0);
      if (valid[u] && nm != kFull) old[u] = ld_stream(llr0 + (size_t(vrow[u]) << log2_g) + g);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t i = iv[u];
      L[u] = make_uint4(0, 0, 0, 0);
      if (valid[u] && i < n_regular) {
        const uint64_t at_i = uint64_t(i) * src_stride;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if ((nm >> j) & 1u) {
            // packed hard-decision input (1 bit per transmitted variable, frame-major words): bit 1 <-> +1
            uint32_t pbit = 0;
            if (packed_bits) pbit = (__ldg(static_cast<const uint32_t *>(src) + off[j] + (i >> 5)) >> (i & 31u)) & 1u;
            if (V == 4) {
              uint32_t raw;
              if (packed_bits) raw = pbit ? 0x3F800000u : 0xBF800000u;
              else if (kSrc == kSrcHalf)
                raw = __float_as_uint(__half2float(__ushort_as_half(__ldg(static_cast<const uint16_t *>(src) + off[j] + at_i))));
              else raw = __ldg(static_cast<const uint32_t *>(src) + off[j] + at_i);
              word(L[u], j & 3) = TraitsF32::to_llr(raw, channel, param);
            } else {
              const uint32_t raw = TraitsF16::to_llr(packed_bits ? uint16_t(pbit ? 0x3C00u : 0xBC00u)
                                                                 : __ldg(static_cast<const uint16_t *>(src) + off[j] + at_i),
                                                     channel, param);
              word(L[u], (j >> 1) & 3) |= raw << ((j & 1) * 16);
            }
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (!valid[u]) continue;
#pragma unroll
      for (int w = 0; w < 4; ++w) word(old[u], w) = (word(old[u], w) & ~word(SEL, w)) | word(L[u], w);
      st_stream(llr0 + (size_t(vrow[u]) << log2_g) + g, old[u]);
    }
    if (kMsgRows) {
      // start messages: phi_abs(|prior|) with the prior's sign (punctured variables: phi_abs(0))
      const uint32_t i = vrow[0];
      uint4 O = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t lw = word(L[0], w);
        uint32_t r;
        if (V == 4) r = __float_as_uint(phi_abs_f32(__uint_as_float(lw & 0x7FFFFFFFu)));
        else r = uint32_t(__ldg(phi_table + (lw & 0x7FFFu))) | (uint32_t(__ldg(phi_table + ((lw >> 16) & 0x7FFFu))) << 16);
        word(O, w) = ((r & ~Tr::kSignMask) | (lw & Tr::kSignMask)) & word(SEL, w);
      }
      const uint32_t a = __ldg(col_ptr + i), b = __ldg(col_ptr + i + 1);
      for (uint32_t ie = a; ie < b; ++ie) {
        uint4 *p = msg + (size_t(__ldg(in_to_out + ie)) << log2_g) + g;
        uint4 oldm = make_uint4(0, 0, 0, 0);
        if (nm != kFull) oldm = ld_stream(p);
#pragma unroll
        for (int w = 0; w < 4; ++w) word(oldm, w) = (word(oldm, w) & ~word(SEL, w)) | word(O, w);
        st_stream(p, oldm);
      }
    }
  }
}

// Priors-only refill with the new frames' channel values STAGED THROUGH SHARED MEMORY (interleaved sources, rows of
// at least 256 vectors).  In refill_kernel every new lane reads its 2- or 4-byte channel value with a load of its own:
// one 32-byte sector request per (new lane, variable) -- 8-12 ms per refilling period at 2^12 slots, more than the
// read-modify-write of the prior vectors itself.  The frames a check loads are CONSECUTIVE frames, i.e. consecutive
// columns of a variable's row in the caller's array (or runs of 32 in the staging ring), so here a CTA first copies
// the values of rows_per_trip variables for ALL new frames into shared memory with coalesced loads (entry t = t-th
// new frame), then every thread picks the values of its new lanes by entry index.  One CTA = 256 consecutive vectors
// of a row; the CTAs of a row walk the variables together.
template <class Tr, int kSrc>
__global__ void __launch_bounds__(kThreads)
    refill_priors_staged_kernel(uint4 *__restrict__ llr0, const void *__restrict__ src,
                                const RefillEntry *__restrict__ entries, const uint32_t *__restrict__ n_new_ptr,
                                const uint32_t *__restrict__ new_mask, const uint32_t *__restrict__ slot_entry,
                                uint32_t n_vars, uint32_t n_regular, uint32_t log2_g, int channel, float param,
                                const uint32_t *__restrict__ ctl, uint32_t gate, const uint32_t *__restrict__ var_pos,
                                uint32_t rows_per_trip) {
  const uint32_t n_new = *n_new_ptr;
  if (n_new == 0) return;
  if (gated_off(ctl, gate)) return;
  constexpr int V = Tr::kFramesPerVec;
  constexpr uint32_t kFull = (1u << V) - 1u;
  constexpr bool kHalfElems = V == 8 || kSrc == kSrcHalf;
  typedef typename std::conditional<kHalfElems, uint16_t, uint32_t>::type src_t;
  extern __shared__ __align__(16) uint8_t s_stage_raw[];
  src_t *s_val = reinterpret_cast<src_t *>(s_stage_raw); // [rows_per_trip][n_new]
  const uint32_t ctas_per_row = (1u << log2_g) / kThreads;  // >= 1 (host)
  const uint32_t g = (blockIdx.x % ctas_per_row) * kThreads + threadIdx.x;
  const uint32_t group = blockIdx.x / ctas_per_row, n_groups = gridDim.x / ctas_per_row;
  const uint32_t nm = load_frame_bits<Tr>(reinterpret_cast<const uint8_t *>(new_mask), g);
  uint32_t ent[V];
  uint4 SEL;
#pragma unroll
  for (int j = 0; j < V; ++j) ent[j] = ((nm >> j) & 1u) ? __ldg(slot_entry + g * V + j) : 0u;
#pragma unroll
  for (int w = 0; w < 4; ++w) word(SEL, w) = Tr::lane_mask(nm >> (w * Tr::kFramesPerWord));
  const uint64_t src_stride = entries[0].src_stride; // the same for every frame of a launch
  const src_t *sp = static_cast<const src_t *>(src);
  for (uint32_t i0 = group * rows_per_trip; i0 < n_vars; i0 += n_groups * rows_per_trip) {
    for (uint32_t t = threadIdx.x; t < n_new; t += kThreads) {
      const uint64_t off = entries[t].src_offset;
      for (uint32_t r = 0; r < rows_per_trip; ++r) {
        const uint32_t i = i0 + r;
        if (i < n_regular) s_val[r * n_new + t] = __ldg(sp + off + uint64_t(i) * src_stride);
      }
    }
    __syncthreads();
    if (nm) {
      for (uint32_t r = 0; r < rows_per_trip; ++r) {
        const uint32_t i = i0 + r;
        if (i >= n_vars) break;
        uint4 L = make_uint4(0, 0, 0, 0);
        if (i < n_regular) {
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if ((nm >> j) & 1u) {
              const src_t raw = s_val[r * n_new + ent[j]];
              if (V == 4) {
                const uint32_t wide = kSrc == kSrcHalf ? __float_as_uint(__half2float(__ushort_as_half(uint16_t(raw)))) : uint32_t(raw);
                word(L, j & 3) = TraitsF32::to_llr(wide, channel, param);
              } else {
                word(L, (j >> 1) & 3) |= uint32_t(TraitsF16::to_llr(uint16_t(raw), channel, param)) << ((j & 1) * 16);
              }
            }
          }
        }
        const uint32_t vrow = var_pos ? __ldg(var_pos + i) : i;
        uint4 *p = llr0 + (size_t(vrow) << log2_g) + g;
        uint4 old = make_uint4(0, 0, 0, 0);
        if (nm != kFull) old = ld_stream(p);
#pragma unroll
        for (int w = 0; w < 4; ++w) word(old, w) = (word(old, w) & ~word(SEL, w)) | word(L, w);
        st_stream(p, old);
      }
    }
    __syncthreads(); // the staged values are overwritten by the next trip
  }
}

// Target syndromes of the new frames: frame-major packed words (bit c%32 of word c/32) -> bit-sliced syn[c][slot word].
// One thread = (group of 32 checks, slot word w) for words that received new frames: it loads ONE source word per new
// slot of its word (the 32 target bits of that frame for its checks) and then rewrites its 32 syn words, inserting bit
// cc of every source word at the slot's position -- 32x fewer source and entry loads than one thread per check
// (which spent 3 ms per refill on dependent scalar loads).  New slots are handled 8 at a time (registers).
__global__ void __launch_bounds__(kThreads)
    refill_syndrome_kernel(uint32_t *__restrict__ syn, const uint32_t *__restrict__ src_syn,
                           const RefillEntry *__restrict__ entries, const uint32_t *__restrict__ n_new_ptr,
                           const uint32_t *__restrict__ new_mask, const uint32_t *__restrict__ slot_entry,
                           uint32_t n_checks, uint32_t log2_pw, const uint32_t *__restrict__ check_pos) {
  if (*n_new_ptr == 0) return;
  constexpr int kBatch = 8;
  const uint32_t PW = 1u << log2_pw;
  const uint32_t n_groups = (n_checks + 31u) >> 5;
  const uint64_t total = uint64_t(n_groups) << log2_pw;
  for (uint64_t item = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; item < total;
       item += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t cg = uint32_t(item >> log2_pw);
    const uint32_t w = uint32_t(item) & (PW - 1u);
    uint32_t mask = __ldg(new_mask + w);
    const uint32_t c0 = cg << 5, cn = min(32u, n_checks - c0);
    while (mask) {
      uint32_t src[kBatch], pos[kBatch], batch_mask = 0;
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        src[j] = 0;
        pos[j] = 0;
        if (mask) {
          const uint32_t bit = __ffs(mask) - 1;
          mask &= mask - 1;
          batch_mask |= 1u << bit;
          pos[j] = bit;
          src[j] = __ldg(src_syn + entries[__ldg(slot_entry + w * 32 + bit)].syn_offset + cg);
        }
      }
      for (uint32_t cc = 0; cc < cn; ++cc) {
        uint32_t *p = syn + (size_t(check_pos ? __ldg(check_pos + c0 + cc) : c0 + cc) << log2_pw) + w;
        uint32_t x = *p & ~batch_mask;
#pragma unroll
        for (int j = 0; j < kBatch; ++j) x |= ((src[j] >> cc) & 1u) << pos[j]; // unused entries: src = 0
        *p = x;
      }
    }
  }
}

// ---------------------------------------------------------------- extraction of retired frames
// One warp = (slot word w, tile of 32 result words).  Lane i reads hard[32j+i][w]; a ballot per retiring
// slot of that word assembles result word j for that slot's frame; each warp then writes 128 B per frame
// (LSB-first packing as flood.cu:277-295, but only for slots that actually retire).
struct ExtractEntry {
  uint64_t dst_offset; // u32 offset of the frame's packed result
};

__global__ void __launch_bounds__(128)
    extract_kernel(const uint32_t *__restrict__ hard, uint32_t *__restrict__ dst, const uint32_t *__restrict__ retire_mask,
                   const ExtractEntry *__restrict__ slot_dst, uint32_t n_words /* N/32 */, uint32_t log2_pw,
                   const uint32_t *__restrict__ var_pos) {
  __shared__ uint32_t s_out[4][32][33];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t PW = 1u << log2_pw;
  const uint32_t n_tiles = (n_words + 31u) >> 5;
  const uint64_t total = uint64_t(n_tiles) << log2_pw;
  for (uint64_t wi = uint64_t(blockIdx.x) * 4 + warp; wi < total; wi += uint64_t(gridDim.x) * 4) {
    const uint32_t w = uint32_t(wi) & (PW - 1u);
    const uint32_t tile = uint32_t(wi >> log2_pw);
    const uint32_t mask = __ldg(retire_mask + w);
    if (!mask) continue; // warp-uniform
    const uint32_t j0 = tile << 5;
    const uint32_t jn = min(32u, n_words - j0);
    for (uint32_t jj = 0; jj < jn; ++jj) {
      const uint32_t var = (j0 + jj) * 32 + lane;
      const uint32_t x = __ldg(hard + (size_t(var_pos ? __ldg(var_pos + var) : var) << log2_pw) + w);
      uint32_t mm = mask, k = 0;
      while (mm) {
        const uint32_t bit = __ffs(mm) - 1;
        mm &= mm - 1;
        const uint32_t r = __ballot_sync(0xFFFFFFFFu, (x >> bit) & 1u);
        if (lane == 0) s_out[warp][k][jj] = r;
        ++k;
      }
    }
    __syncwarp();
    uint32_t mm = mask, k = 0;
    while (mm) {
      const uint32_t bit = __ffs(mm) - 1;
      mm &= mm - 1;
      if (lane < jn) dst[slot_dst[w * 32 + bit].dst_offset + j0 + lane] = s_out[warp][k][lane];
      ++k;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- device-side scheduler
// The reference decides on the HOST, after a blocking read of P flag bytes, which frames stop, which are
// retired and where the new ones go (ldpc_decoder_gpu.cu:374-413, 464-516, 604-609).  Here one CTA does it on
// the device right after the syndrome check, so the next flood iterations never wait for the host:
//   stop rule -> n_new -> the reference's swap pairing (as a prefix-sum compaction of finished positions)
//   -> extraction list, refill list, slot masks, per-frame iteration counters, a summary for the host.
struct SchedConst {
  uint32_t n_frames, batch, log2_p, k_check, imax, simd_pairs, syn_words, res_words;
  uint32_t input_mode;   // 0 interleaved [i][n_cols], 1 frame-major, 2 ring of interleaved blocks, 3 frame-major ring
  uint32_t n_cols;       // frames physically present in the input (pool or n_frames)
  uint32_t ring_frames, ring_block, n_regular;
  uint32_t syn_ring;     // syndromes staged in a ring (position f % ring_frames) instead of read in place
  uint32_t results_ring; // results go to a [P][W] device ring (host drains it) instead of results[frame]
  uint32_t row_elems;    // staged elements per frame in the frame-major ring (n_regular, or its word count when packed)
  uint64_t frame_stride;
  // LDPC_B200_FLAG_EARLY_RETIRE: once the input is exhausted, retire frames as soon as they stop and compact the
  // survivors into the low slots so that the flood kernels touch only 2^log2_ga vectors per row
  uint32_t early_retire, log2_v /* log2 frames per 16-byte vector */, min_log2_p_active;
  // a period that loads at least this many frames fuses its refill into its first check-node pass (0xFFFFFFFF: never)
  uint32_t fused_min_new;
  // Shared frame queue (ldpc_b200_decode_from_queue: several decoders, one per GPU, draw from one queue): this decoder
  // numbers the frames it has been handed 0, 1, 2, ... ("local" index: iteration counters, staging ring); they are
  // whole blocks of queue_block consecutive frames of the caller's arrays, block i starting at global frame
  // FrameQueueView::block_start[i].  queue_block = 0: classic call, local index = global frame index.
  uint32_t queue_block, pad3;
};
// what the scheduler kernel reads of the queue: HOST memory mapped into the device (cudaHostAllocMapped), so that a
// scheduler kernel enqueued one period ahead still sees what the host thread has reserved since
struct FrameQueueView {
  const volatile uint32_t *ctl;     // [0] = local frames reserved so far, [1] = 1 once no more frames will come
  const volatile uint32_t *block_start; // [i] = first global frame of local block i (0xFFFFFFFF: not written yet)
};
struct SchedState {
  uint32_t next_to_load, g;
  uint32_t done, log2_ga; // control block read by the streaming kernels (ctl[0], ctl[1]): keep adjacent
  uint32_t period, batch_active;
  uint32_t fused_now, pad; // ctl[kCtlFused]: the refill of this period is fused into its first check-node pass
};
static_assert(offsetof(SchedState, fused_now) - offsetof(SchedState, done) == 4 * kCtlFused, "control block layout");
struct SchedSummary {
  uint32_t n_new, n_retired, next_to_load, done, g, n_stop, period, n_moves;
  uint32_t log2_p_active; // slots the flood kernels process from now on
  uint32_t fused;         // this period's refill is fused into the next check-node pass
  uint32_t pad[2];
};
struct SlotMove {
  uint32_t src, dst;
};

// global frame index (row of the caller's arrays) of local frame f
__device__ __forceinline__ uint32_t global_frame(const SchedConst &c, const FrameQueueView &q, uint32_t f) {
  if (c.queue_block == 0) return f;
  return q.block_start[f / c.queue_block] + f % c.queue_block;
}
// frames of the queue this decoder may load now, and whether that is all it will ever get
__device__ __forceinline__ void queue_state(const SchedConst &c, const FrameQueueView &q, uint32_t &avail, uint32_t &final) {
  avail = c.n_frames;
  final = 1;
  if (c.queue_block) {
    avail = q.ctl[0];
    final = q.ctl[1];
    // a block whose start has not become visible yet is not available (the host writes the start first)
    while (avail && q.block_start[(avail - 1) / c.queue_block] == 0xFFFFFFFFu) {
      avail = ((avail - 1) / c.queue_block) * c.queue_block;
      final = 0;
    }
  }
}

__device__ __forceinline__ void make_refill_entry(const SchedConst &c, const FrameQueueView &q, uint32_t frame, uint32_t slot,
                                                  RefillEntry &en) {
  const uint32_t col = global_frame(c, q, frame) % c.n_cols;
  if (c.input_mode == 0) {
    en.src_offset = col;
    en.src_stride = c.n_cols;
  } else if (c.input_mode == 1) {
    en.src_offset = uint64_t(col) * c.frame_stride;
    en.src_stride = 1;
  } else if (c.input_mode == 2) {
    const uint32_t r = frame % c.ring_frames;
    en.src_offset = uint64_t(r / c.ring_block) * c.n_regular * c.ring_block + (r % c.ring_block);
    en.src_stride = c.ring_block;
  } else {
    en.src_offset = uint64_t(frame % c.ring_frames) * c.row_elems;
    en.src_stride = 1;
  }
  en.syn_offset = uint64_t(c.syn_ring ? frame % c.ring_frames : col) * c.syn_words;
  en.slot = slot;
  en.pad = 0;
}

constexpr int kSchedThreads = 1024;
constexpr int kMaxSlots = 8192;

// block-wide exclusive scan of one value per thread (1024 threads)
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t &total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, d);
      if (lane >= d) w += y;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  total = s_warp[31];
  const uint32_t base = warp ? s_warp[warp - 1] : 0u;
  __syncthreads();
  return base + x - v;
}

// First batch: frames 0..batch-1 into slots 0..batch-1 (ldpc_decoder_gpu.cu:300-315).
__global__ void __launch_bounds__(kSchedThreads)
    sched_init_kernel(SchedConst c, SchedState *st, uint32_t *in_gpu, uint32_t *virt, uint32_t *iter_start,
                      uint32_t *iter_end, RefillEntry *entries, uint32_t *n_new_out, uint32_t *new_mask,
                      uint32_t *retire_mask, uint32_t *slot_entry, FrameQueueView q) {
  const uint32_t PW = (1u << c.log2_p) >> 5;
  for (uint32_t i = threadIdx.x; i < c.n_frames; i += blockDim.x) iter_start[i] = iter_end[i] = 0xFFFFFFFFu;
  for (uint32_t w = threadIdx.x; w < PW; w += blockDim.x) {
    const uint32_t lo = w * 32u;
    new_mask[w] = c.batch >= lo + 32u ? 0xFFFFFFFFu : (c.batch > lo ? (1u << (c.batch - lo)) - 1u : 0u);
    retire_mask[w] = 0;
  }
  for (uint32_t j = threadIdx.x; j < c.batch; j += blockDim.x) {
    in_gpu[j] = virt[j] = j;
    slot_entry[j] = j;
    make_refill_entry(c, q, j, j, entries[j]);
  }
  if (threadIdx.x == 0) {
    st->next_to_load = c.batch;
    st->g = c.k_check; // the first check happens at global iteration k (after k+1 flood iterations)
    st->done = 0;
    st->log2_ga = 0; // all vectors of a row are live
    st->period = 0;
    st->batch_active = c.batch;
    st->fused_now = c.batch >= c.fused_min_new ? 1u : 0u;
    st->pad = 0;
    *n_new_out = c.batch;
  }
}

__global__ void __launch_bounds__(kSchedThreads)
    sched_kernel(SchedConst c, SchedState *st, const uint32_t *__restrict__ viol, uint32_t *in_gpu, uint32_t *virt,
                 uint32_t *iter_start, uint32_t *iter_end, RefillEntry *entries, ExtractEntry *slot_dst,
                 uint32_t *n_new_out, uint32_t *new_mask, uint32_t *retire_mask, uint32_t *slot_entry,
                 uint32_t *retired_frames, SchedSummary *summary, uint32_t *scratch /* [2P] */, SlotMove *moves,
                 uint32_t *n_moves_out, FrameQueueView q) {
  __shared__ uint8_t s_stop[kMaxSlots];
  __shared__ uint32_t s_partner[kMaxSlots]; // finished positions >= n_new, compacted in position order
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_n_new, s_all_done, s_avail, s_final;
  const uint32_t B = st->batch_active, PW = (1u << c.log2_p) >> 5, tid = threadIdx.x;
  if (tid == 0) {
    *n_moves_out = 0;
    uint32_t avail, final;
    queue_state(c, q, avail, final);
    s_avail = avail;
    s_final = final;
  }
  __syncthreads();
  // frames this decoder can load at this check (all of the call's frames unless it draws from a shared queue), and
  // whether the input is exhausted once they are loaded
  const uint32_t avail = s_avail;
  const bool input_final = s_final != 0;
  if (st->done) { // run-ahead period after the end: nothing to retire or load
    for (uint32_t w = tid; w < PW; w += blockDim.x) new_mask[w] = retire_mask[w] = 0;
    if (tid == 0) {
      *n_new_out = 0;
      SchedSummary sm = {0, 0, st->next_to_load, 1, st->g, 0, st->period, 0, c.log2_p, 0, {0, 0}};
      *summary = sm;
      st->period++;
      st->fused_now = 0;
    }
    return;
  }
  const uint32_t g = st->g, next = st->next_to_load, period = st->period;
  // stop rule (ldpc_decoder_gpu.cu:395-403)
  const uint32_t chunk = (B + blockDim.x - 1) / blockDim.x;
  const uint32_t j0 = min(B, tid * chunk), j1 = min(B, j0 + chunk);
  uint32_t my_stop = 0;
  for (uint32_t j = j0; j < j1; ++j) {
    const uint32_t s = virt[j], f = in_gpu[j];
    const bool violated = (viol[s >> 5] >> (s & 31u)) & 1u;
    const uint32_t n_it = g - iter_start[f];
    const bool stop = !violated || n_it >= c.imax;
    s_stop[j] = stop;
    if (stop) {
      ++my_stop;
      if (iter_end[f] == 0xFFFFFFFFu) iter_end[f] = g;
    }
  }
  if (c.early_retire && input_final && next == avail) {
    // ---- input exhausted, early-retire mode: every stopped frame leaves now; survivors are compacted
    __shared__ uint32_t s_occ[kMaxSlots / 32];
    __shared__ uint32_t s_p_new;
    for (uint32_t w = tid; w < PW; w += blockDim.x) new_mask[w] = retire_mask[w] = 0;
    for (uint32_t w = tid; w < kMaxSlots / 32; w += blockDim.x) s_occ[w] = 0;
    uint32_t tot_stop;
    uint32_t rank_stop = block_exclusive_scan(my_stop, s_warp, tot_stop);
    uint32_t tot_keep;
    uint32_t rank_keep = block_exclusive_scan((j1 - j0) - my_stop, s_warp, tot_keep);
    uint32_t *keep_frame = scratch, *keep_slot = scratch + (1u << c.log2_p);
    for (uint32_t j = j0; j < j1; ++j) {
      const uint32_t s = virt[j], f = in_gpu[j];
      if (s_stop[j]) {
        const uint32_t t = rank_stop++;
        atomicOr(retire_mask + (s >> 5), 1u << (s & 31u));
        slot_dst[s].dst_offset = c.results_ring ? uint64_t(t) * c.res_words : uint64_t(global_frame(c, q, f)) * c.res_words;
        retired_frames[t] = f;
      } else {
        const uint32_t k = rank_keep++;
        keep_frame[k] = f;
        keep_slot[k] = s;
      }
    }
    __syncthreads();
    const uint32_t Bn = tot_keep;
    // target slot range: the smallest power of two holding the survivors (never below the kernels' minimum)
    if (tid == 0) {
      uint32_t lp = c.min_log2_p_active;
      while ((1u << lp) < Bn) ++lp;
      const uint32_t cur = st->log2_ga ? st->log2_ga + c.log2_v : c.log2_p;
      s_p_new = lp < cur ? lp : cur;
    }
    __syncthreads();
    const uint32_t log2_pn = s_p_new, Pn = 1u << log2_pn;
    const uint32_t cur_log2_p = st->log2_ga ? st->log2_ga + c.log2_v : c.log2_p;
    const uint32_t chunk_k = (Bn + blockDim.x - 1) / blockDim.x;
    const uint32_t k0 = min(Bn, tid * chunk_k), k1 = min(Bn, k0 + chunk_k);
    uint32_t n_moves = 0;
    if (log2_pn < cur_log2_p && Bn > 0) {
      uint32_t my_movers = 0;
      for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t s = keep_slot[k];
        if (s < Pn) atomicOr(s_occ + (s >> 5), 1u << (s & 31u));
        else ++my_movers;
      }
      __syncthreads();
      uint32_t tot_movers;
      uint32_t rank_m = block_exclusive_scan(my_movers, s_warp, tot_movers);
      // free slots below Pn, in increasing order
      const uint32_t chunk_s = (Pn + blockDim.x - 1) / blockDim.x;
      const uint32_t s0 = min(Pn, tid * chunk_s), s1 = min(Pn, s0 + chunk_s);
      uint32_t my_free = 0;
      for (uint32_t s2 = s0; s2 < s1; ++s2) my_free += !((s_occ[s2 >> 5] >> (s2 & 31u)) & 1u);
      uint32_t tot_free;
      uint32_t rank_f = block_exclusive_scan(my_free, s_warp, tot_free);
      for (uint32_t s2 = s0; s2 < s1; ++s2)
        if (!((s_occ[s2 >> 5] >> (s2 & 31u)) & 1u)) s_partner[rank_f++] = s2;
      __syncthreads();
      for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t s = keep_slot[k];
        if (s >= Pn) {
          const uint32_t dst = s_partner[rank_m];
          moves[rank_m].src = s;
          moves[rank_m].dst = dst;
          keep_slot[k] = dst;
          ++rank_m;
        }
      }
      n_moves = tot_movers;
      __syncthreads();
    }
    for (uint32_t k = k0; k < k1; ++k) {
      in_gpu[k] = keep_frame[k];
      virt[k] = keep_slot[k];
    }
    if (tid == 0) {
      *n_new_out = 0;
      *n_moves_out = n_moves;
      st->g = g + c.k_check;
      st->done = Bn == 0 ? 1u : 0u;
      if (Bn > 0 && log2_pn < cur_log2_p) st->log2_ga = log2_pn - c.log2_v;
      st->batch_active = Bn;
      st->period = period + 1;
      st->fused_now = 0;
      SchedSummary sm = {0, tot_stop, next, Bn == 0 ? 1u : 0u, g, tot_stop, period, n_moves,
                         (Bn > 0 && log2_pn < cur_log2_p) ? log2_pn : cur_log2_p, 0, {0, 0}};
      *summary = sm;
    }
    return;
  }
  uint32_t n_stop;
  block_exclusive_scan(my_stop, s_warp, n_stop);
  if (tid == 0) {
    const uint32_t remaining = avail > next ? avail - next : 0u;
    s_all_done = (input_final && remaining == 0 && n_stop == B); // :414
    uint32_t n_new = min(remaining, n_stop);      // :464
    if (c.simd_pairs) n_new &= ~1u;               // ldpc_decoder_gpu.cpp:623
    s_n_new = s_all_done ? 0u : n_new;
  }
  for (uint32_t w = tid; w < PW; w += blockDim.x) new_mask[w] = retire_mask[w] = 0;
  __syncthreads();
  const uint32_t n_new = s_n_new;
  const bool all_done = s_all_done;
  if (all_done) { // retire everything, in position order (:437-460)
    for (uint32_t j = j0; j < j1; ++j) {
      const uint32_t s = virt[j], f = in_gpu[j];
      atomicOr(retire_mask + (s >> 5), 1u << (s & 31u));
      slot_dst[s].dst_offset = c.results_ring ? uint64_t(j) * c.res_words : uint64_t(global_frame(c, q, f)) * c.res_words;
      retired_frames[j] = f;
    }
  } else if (n_new > 0) {
    // the reference pairs, in position order, every still-running position o < n_new with a finished position
    // q >= n_new and swaps them (:487-516): rank both sets with a scan, match equal ranks
    uint32_t my_a = 0, my_b = 0;
    for (uint32_t j = j0; j < j1; ++j) {
      if (j < n_new) my_a += !s_stop[j];
      else my_b += s_stop[j];
    }
    uint32_t tot_a, tot_b;
    uint32_t rank_a = block_exclusive_scan(my_a, s_warp, tot_a);
    uint32_t rank_b = block_exclusive_scan(my_b, s_warp, tot_b);
    for (uint32_t j = j0; j < j1; ++j)
      if (j >= n_new && s_stop[j]) s_partner[rank_b++] = j;
    __syncthreads();
    for (uint32_t j = j0; j < j1; ++j)
      if (j < n_new && !s_stop[j]) {
        const uint32_t q = s_partner[rank_a++];
        const uint32_t f = in_gpu[j], s = virt[j];
        in_gpu[j] = in_gpu[q];
        virt[j] = virt[q];
        in_gpu[q] = f;
        virt[q] = s;
      }
    __syncthreads();
    // positions [0,n_new) now hold finished frames: retire them and load frames next.. into their slots
    for (uint32_t j = tid; j < n_new; j += blockDim.x) {
      const uint32_t s = virt[j], f_old = in_gpu[j], f_new = next + j;
      atomicOr(retire_mask + (s >> 5), 1u << (s & 31u));
      atomicOr(new_mask + (s >> 5), 1u << (s & 31u));
      slot_dst[s].dst_offset =
          c.results_ring ? uint64_t(j) * c.res_words : uint64_t(global_frame(c, q, f_old)) * c.res_words;
      retired_frames[j] = f_old;
      slot_entry[s] = j;
      make_refill_entry(c, q, f_new, s, entries[j]);
      in_gpu[j] = f_new; // :604-609
      iter_start[f_new] = g;
    }
  }
  if (tid == 0) {
    *n_new_out = n_new;
    st->next_to_load = next + n_new;
    st->g = g + c.k_check;
    st->done = all_done ? 1u : 0u;
    st->period = period + 1;
    const uint32_t fused = (n_new > 0 && n_new >= c.fused_min_new) ? 1u : 0u;
    st->fused_now = fused;
    SchedSummary sm = {n_new, all_done ? B : n_new, next + n_new, all_done ? 1u : 0u, g, n_stop, period, 0,
                       c.log2_p, fused, {0, 0}};
    *summary = sm;
  }
}

// Early-retire compaction: copy the state of the frames named in `moves` (src slot -> dst slot): messages, priors
// and target-syndrome bits.  One thread = (row, move).  (The reference moves frames with flood_permute_vecs on every
// refill, flood.cu:225-275; here it only happens in the tail of a call, for the few frames still running.)
template <typename Elem>
__global__ void __launch_bounds__(kThreads)
    move_slots_kernel(Elem *__restrict__ msg, Elem *__restrict__ llr0, uint32_t *__restrict__ syn,
                      const SlotMove *__restrict__ moves, const uint32_t *__restrict__ n_moves_ptr, uint32_t n_edges,
                      uint32_t n_vars, uint32_t n_checks, uint32_t log2_p) {
  const uint32_t n_moves = *n_moves_ptr;
  if (n_moves == 0) return;
  const uint64_t rows = uint64_t(n_edges) + n_vars + n_checks;
  const uint64_t total = rows * n_moves;
  const uint32_t log2_pw = log2_p - 5;
  for (uint64_t item = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; item < total;
       item += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t row = item / n_moves;
    const SlotMove mv = moves[item - row * n_moves];
    if (row < n_edges) {
      msg[(size_t(row) << log2_p) + mv.dst] = msg[(size_t(row) << log2_p) + mv.src];
    } else if (row < uint64_t(n_edges) + n_vars) {
      const size_t v = size_t(row - n_edges);
      llr0[(v << log2_p) + mv.dst] = llr0[(v << log2_p) + mv.src];
    } else {
      const size_t c = size_t(row - n_edges - n_vars);
      const uint32_t bit = (syn[(c << log2_pw) + (mv.src >> 5)] >> (mv.src & 31u)) & 1u;
      uint32_t *w = syn + (c << log2_pw) + (mv.dst >> 5);
      if (bit) atomicOr(w, 1u << (mv.dst & 31u));
      else atomicAnd(w, ~(1u << (mv.dst & 31u)));
    }
  }
}

// gather one slot's column of a [rows][P] array (debug read-back)
template <typename Elem>
__global__ void gather_slot_kernel(const Elem *__restrict__ src, Elem *__restrict__ dst, uint64_t rows, uint32_t log2_p,
                                   uint32_t slot) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < rows) dst[i] = src[(i << log2_p) + slot];
}

} // namespace dev
} // namespace ldpc_b200
