// phx_family.h -- host-side base class of a device program family.
//
// A family owns the HBM-resident state of E env instances of one env class and launches
// the fused reset / step kernels for it.  The generic per-env bookkeeping (clock, episode,
// stage, done sets, fault words, message trace) lives here; the agent state columns and the
// kernels live in fam_<name>.cu.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "phx_common.cuh"
#include "phx_hostpool.h"

namespace phx {

class Family {
 public:
  virtual ~Family();

  // Validates the spec against the family's domain and allocates state.
  virtual int32_t init(const phx_spec& spec) = 0;
  virtual int32_t reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                        cudaStream_t stream) = 0;
  virtual int32_t rollout(int32_t T, const StepIO& io, cudaStream_t stream) = 0;
  // Run-time specialisation (include/phx.h: phx_jit_source / phx_load_specialised): the text of a
  // translation unit whose step kernel has THIS handle's lowered env class as a compile-time
  // constant, and the loader of its compiled cubin.  Families without one: PHX_ERR_UNSUPPORTED.
  virtual int32_t jit_source(std::string& out) {
    (void)out;
    set_error("this env class / kernel variant has no run-time specialisation");
    return PHX_ERR_UNSUPPORTED;
  }
  // jit_source for a family object that was never initialised on a device (spec / E / seed /
  // env_offset set by hand): the self-test hook phx_selftest_jit_source.
  virtual int32_t jit_source_offline(std::string& out) { return jit_source(out); }
  virtual int32_t load_specialised(const char* cubin_path) {
    (void)cubin_path;
    set_error("this env class / kernel variant has no run-time specialisation");
    return PHX_ERR_UNSUPPORTED;
  }
  // phx_rollout_host: every pointer of `host` is HOST memory.  The default stages the planes in
  // device memory and runs a chunked three-stream pipeline (H2D | kernel | D2H); a family may
  // override it with a cheaper wire format for its results.
  virtual int32_t rollout_host(int32_t T, const StepIO& host);
  // Family state columns (field >= PHX_FIELD_FAMILY); returns PHX_ERR_INVALID if unknown.
  virtual int32_t family_field(int32_t field, int32_t index, void** dev_ptr, size_t* bytes) = 0;
  virtual const char* exec_name() const = 0;

  int32_t base_init(const phx_spec& spec, int32_t num_envs, int32_t device, uint64_t seed,
                    int64_t env_offset);
  int32_t field_ptr(int32_t field, int32_t index, void** dev_ptr, size_t* bytes);

  phx_spec spec{};
  int32_t E = 0;
  int32_t device = 0;
  uint64_t seed = 0;
  int64_t env_offset = 0;
  int32_t mask_words = 1;  // W = ceil(S / 32)

  int4* d_hdr = nullptr;       // [E]
  uint32_t* d_term = nullptr;  // [W][E]
  uint32_t* d_trunc = nullptr; // [W][E]
  uint32_t* d_err = nullptr;   // [E]
  uint32_t* d_nfaults = nullptr;  // [1]
  int4* d_trace = nullptr;     // [trace_T, E, cap]  the message trace of the last launch
  int32_t* d_trace_cnt = nullptr;  // [trace_T, E]
  int32_t trace_T = 1;         // steps of the last tracked launch
  int32_t trace_T_cap = 1;     // steps the trace buffers can hold
  int32_t ensure_trace(int32_t T);  // grows the trace buffers to T steps; records trace_T

  // staging for the *_host entry points
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;  // phx_rollout_host pipeline
  cudaEvent_t ev_in[16] = {}, ev_k[16] = {}, ev_out[16] = {};
  int32_t ensure_stage(size_t bytes);  // (re)allocates d_stage to at least `bytes`
  HostPool& host_pool() {              // host threads of the *_host entry points, built on demand
    if (!pool_) pool_.reset(new HostPool(HostPool::default_threads()));
    return *pool_;
  }
  std::unique_ptr<HostPool> pool_;

  FaultSink fault_sink() const { return FaultSink{d_err, d_nfaults}; }
  TraceSink trace_sink() const { return TraceSink{d_trace, d_trace_cnt, spec.trace_capacity}; }
  bool tracking() const { return (spec.flags & PHX_FLAG_TRACK_MESSAGES) != 0; }
};

Family* make_supply_chain_family(const phx_spec& spec);
Family* make_mock_family(const phx_spec& spec);
Family* make_market_family(const phx_spec& spec);
Family* make_stackelberg_family(const phx_spec& spec);
Family* make_dense_family(const phx_spec& spec);
Family* make_supply_chain2_family(const phx_spec& spec);
Family* make_simple_market_family(const phx_spec& spec);
Family* make_digital_ads_family(const phx_spec& spec);
Family* make_user_family(const char* cubin_path, const phx_spec& spec);
int32_t selftest_ratio(int32_t device, int32_t den, int32_t lo, int32_t count, float* host_out);

}  // namespace phx
