// phx_user.cuh -- device programs that are NOT compiled into libphx.so.
//
// The reference's plugin API is "subclass Agent and write handlers" (phantom/agents.py:48-60,
// 122-155): any experiment may bring its own agent classes.  Python handlers cannot run on the
// device, but a new env class must not need a rebuild of the library either.  A user writes the
// callbacks of their agent classes as a device program -- the same interface the shipped
// families implement (view / act / pre / handle / post / encode / reward / terminated /
// truncated / reset_agent, see fam_stackelberg.cu for a small one) -- in their own .cu file:
//
//     #include "phx_user.cuh"
//     namespace phx { namespace { struct MyProgram { ... }; } }
//     PHX_USER_PROGRAM(phx::MyProgram)
//
// (constants a program declares: PW payload words (1-2), NWORDS int32 state words per agent, VW
// view words (0-1), OBS_DIM, ACT_DIM, Q1CAP = messages in flight per resolver round, ACTCAP /
// RESPCAP / RECVCAP = messages one agent sends in the acting phase / in a response round /
// receives in a round, and the flags BATCHED, HAS_PRE, HAS_POST)
// and names it from Python (`FamilyInfo(program_source="my_family.cu", ...)`, the agent classes
// carry `__phx_family__` / `__phx_kind__` as usual).  phantom_b200 compiles the file once with
// `nvcc -cubin` for sm_100a (cached by content hash, phantom_b200/jit.py) and libphx loads the
// cubin through the runtime's library API (`phx_create_user`): the kernels below are the
// thread-per-env engine (env classes of at most 8 agents) instantiated for the user's program,
// and `phx_user_desc` tells the library how much state and shared memory the program needs.
//
// Env classes of 9..128 agents run on the 128-lane block engine (phx_engine_wide.cuh).  A program
// opts in by being WIDTH INDEPENDENT: every callback a template over the context type
// (`template <class C> ... const C& c`; Ctx on the thread engine, WCtx on the block engine),
// neighbours iterated with c.next_neighbour() / c.next_of_kind(), never a raw mask word, and
//     static constexpr bool WIDE_OK = true;
// Its queue segments are sized per agent as min(cap, degree in the graph), cap = ACTCAP / RESPCAP
// or, if declared, WIDE_ACTCAP / WIDE_RESPCAP (<= 255).
#pragma once
#define PHX_JIT_TU 1  // (a program unit, not libphx: the families' host classes stay out)
#include "phx_engine1.cuh"
#include "phx_engine_wide.cuh"

namespace phx {
constexpr int32_t PHX_USER_MAGIC = 0x50485855;  // "PHXU"

// block-engine bodies of a user program: empty unless the program declares WIDE_OK
template <class P, bool TRACK>
__device__ __forceinline__ void user_wide_step(const WideArgs<P>& a) {
  if constexpr (IsWideOk<P>::value) wide_step_body<P, TRACK>(a);
}
template <class P>
__device__ __forceinline__ void user_wide_reset(const WideArgs<P>& a, const uint8_t* env_mask,
                                                float* obs, uint8_t* obs_mask, bool agents_only) {
  if constexpr (IsWideOk<P>::value) wide_reset_body<P>(a, env_mask, obs, obs_mask, agents_only);
}
template <class P>
constexpr int32_t user_wide_smem() {
  if constexpr (IsWideOk<P>::value) return (int32_t)sizeof(WideSmem<P>);
  return 0;
}
}  // namespace phx

#define PHX_USER_PROGRAM(Prog)                                                                  \
  extern "C" __global__ void __launch_bounds__(::phx::ENGINE1_BLOCK)                            \
  phx_user_step(const ::phx::EngineArgs<Prog> a) {                                              \
    ::phx::engine1_step_body<Prog, false, ::phx::SpecFromArgs>(a);                              \
  }                                                                                             \
  extern "C" __global__ void __launch_bounds__(::phx::ENGINE1_BLOCK)                            \
  phx_user_step_tracked(const ::phx::EngineArgs<Prog> a) {                                      \
    ::phx::engine1_step_body<Prog, true, ::phx::SpecFromArgs>(a);                               \
  }                                                                                             \
  extern "C" __global__ void __launch_bounds__(::phx::ENGINE_BLOCK)                             \
  phx_user_reset(const ::phx::EngineArgs<Prog> a, const uint8_t* env_mask, float* obs,          \
                 uint8_t* obs_mask, bool agents_only) {                                         \
    ::phx::engine_reset_body<Prog, 8>(a, env_mask, obs, obs_mask, agents_only);                 \
  }                                                                                             \
  extern "C" __global__ void __launch_bounds__(::phx::WIDE_G)                                   \
  phx_user_wide_step(const ::phx::WideArgs<Prog> a) {                                           \
    ::phx::user_wide_step<Prog, false>(a);                                                      \
  }                                                                                             \
  extern "C" __global__ void __launch_bounds__(::phx::WIDE_G)                                   \
  phx_user_wide_step_tracked(const ::phx::WideArgs<Prog> a) {                                   \
    ::phx::user_wide_step<Prog, true>(a);                                                       \
  }                                                                                             \
  extern "C" __global__ void __launch_bounds__(::phx::WIDE_G)                                   \
  phx_user_wide_reset(const ::phx::WideArgs<Prog> a, const uint8_t* env_mask, float* obs,       \
                      uint8_t* obs_mask, bool agents_only) {                                    \
    ::phx::user_wide_reset<Prog>(a, env_mask, obs, obs_mask, agents_only);                      \
  }                                                                                             \
  extern "C" __device__ const int32_t phx_user_desc[12] = {                                     \
      ::phx::PHX_USER_MAGIC, Prog::NWORDS,     Prog::VW,      Prog::PW,                         \
      Prog::ACT_DIM,         Prog::OBS_DIM,    Prog::Q1CAP,   ::phx::EnvWords<Prog>::value,     \
      (int32_t)sizeof(::phx::BlockSmem<Prog, 8>), Prog::BATCHED ? 1 : 0,                        \
      ::phx::user_wide_smem<Prog>(),                                                            \
      ::phx::WideCapConst<Prog>::act | (::phx::WideCapConst<Prog>::resp << 16)};
