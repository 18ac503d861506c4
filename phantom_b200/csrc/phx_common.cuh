// phx_common.cuh -- shared host/device plumbing of libphx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/phx.h"

namespace phx {

// ------------------------------------------------------------------ host error plumbing
void set_error(const std::string& msg);  // thread-local, read by phx_last_error()

#define PHX_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t err__ = (call);                                                     \
    if (err__ != cudaSuccess) {                                                     \
      ::phx::set_error(std::string(#call) + ": " + cudaGetErrorString(err__));      \
      (void)cudaGetLastError(); /* do not leave it for an unrelated later check */  \
      return PHX_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define PHX_REQUIRE(cond, code, msg)   \
  do {                                 \
    if (!(cond)) {                     \
      ::phx::set_error(msg);           \
      return (code);                   \
    }                                  \
  } while (0)

inline bool mask_bit(const uint32_t* m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; }

// --------------------------------------------------------------------- per-env bookkeeping
// Env header, one int4 per env:  x = PhantomEnv.current_step (phantom/env.py:252)
//                                y = episode (RNG contract coordinate)
//                                z = FSM stage index (phantom/fsm.py:171)
//                                w = reserved
// PhantomEnv._terminations / _truncations (phantom/env.py:71-72) are bitmasks over the
// strategic index, stored as uint32 [W][E].

struct StepIO {
  // inputs
  const float* actions;        // [T,E,S,A]
  const uint8_t* action_mask;  // [T,E,S] or null
  // outputs (any may be null = not wanted)
  float* obs;            // [T,E,S,O]
  uint8_t* obs_mask;     // [T,E,S]
  float* reward;         // [T,E,S]
  uint8_t* reward_mask;  // [T,E,S]
  uint8_t* term;         // [T,E,S]
  uint8_t* trunc;        // [T,E,S]
  uint8_t* all_done;     // [T,E,2]
};

// Faults: sticky per-env error word, first code wins; a global counter lets
// phx_poll_errors skip the scan when nothing happened.
struct FaultSink {
  uint32_t* err;       // [E]
  uint32_t* n_faults;  // [1]
};

__device__ __forceinline__ void raise_fault(const FaultSink& f, int env, uint32_t code) {
  if (code != 0u && f.err[env] == 0u) {  // env is owned by exactly one tile: no race
    f.err[env] = code;
    atomicAdd(f.n_faults, 1u);
  }
}

// Message trace (Resolver.tracked_messages, phantom/resolvers.py:41-60): rows of the last
// step, [E, cap] int4 = (sender | recv << 8 | type << 16, payload0, payload1, round).
struct TraceSink {
  int4* rows;    // [E, cap]
  int32_t* cnt;  // [E]
  int32_t cap;
};

__device__ __forceinline__ int4 trace_row(int sender, int recv, int type, int p0, int p1,
                                          int round) {
  return make_int4(sender | (recv << 8) | (type << 16), p0, p1, round);
}

// float32(num / den) as the reference computes it (Python float64 division, then a float32
// cast): Markstein's FMA refinement with the correctly rounded reciprocal rcp = RN(1 / den)
// yields the correctly rounded float32 quotient in four instructions (I2F, FMUL, 2 x FFMA) where
// __fdiv_rn is a ~25-instruction subroutine -- a third of all instructions of the C4 step.
// RN32(RN64(n / d)) == RN32(n / d) for the small integers of these env classes (the float64
// rounding cannot land on a float32 tie).  Checked exhaustively on the device for |num| <= 2^21
// and every denominator the shipped env classes use (tests/test_gpu_supply_chain.py
// test_ratio_exhaustive).
__device__ __forceinline__ float ratio_rn(int num, float den, float rcp) {
  const float x = (float)num;
  const float q = __fmul_rn(x, rcp);
  const float r = __fmaf_rn(-den, q, x);
  return __fmaf_rn(r, rcp, q);
}

// ------------------------------------------------------------------------ memory helpers
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

// cp.async (SASS LDGSTS): 4-byte global -> shared copy that bypasses the register file, so a
// prefetch has no register scoreboard to trip over; completion is tracked per commit group.
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// cp.async.bulk shared::cta -> global (TMA bulk store, SASS UBLKCP).  `bytes` % 16 == 0,
// both addresses 16-byte aligned.  Issued by ONE thread after the writers of `smem_src`
// have fenced (fence.proxy.async) and synchronised.
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src,
                                           uint32_t bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst),
               "r"(s), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // smem source reusable
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {  // writes complete
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}
// Programmatic dependent launch (griddepcontrol, sm_90+): launch_dependents lets the NEXT kernel
// on the stream (if it was launched with cudaLaunchAttributeProgrammaticStreamSerialization)
// start before this grid completes; wait blocks until every earlier grid has completed and its
// memory operations are visible.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

}  // namespace phx
