// phx_engine_wide_host.cuh -- host side of the 128-lane block engine (phx_engine_wide.cuh): a
// Family that owns the state of E envs of a device program P with 33..128 agents (or fewer, when
// PHX_EXEC_WIDE is forced) and launches wide_step_kernel<P>, one block per env.
// Field mapping: PHX_FIELD_FAMILY + w = state word w, int32 [E, 128] (slot-major inside an env);
// PHX_FIELD_TERMINATED / _TRUNCATED = uint32 [E, 4] bitmasks over agent slots;
// PHX_FIELD_ADJACENCY = uint32 [E, 128, 4] rows.
#pragma once
#include "phx_engine_host.cuh"
#include "phx_engine_wide.cuh"

namespace phx {

static __global__ void wide_init_kernel(int E, int4* hdr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) hdr[i] = make_int4(0, -1, 0, 0);  // episode becomes 0 on the first reset
}

template <class P>
class WideFamily : public Family {
 public:
  ~WideFamily() override {
    cudaFree(d_spec);
    cudaFree(d_state);
    cudaFree(d_rcache);
    cudaFree(d_rnone);
    cudaFree(d_ocache);
    cudaFree(d_ocached);
    cudaFree(d_adj);
    cudaFree(d_base);
    cudaFree(d_env);
  }

  int32_t init(const phx_spec& s) override {
    PHX_REQUIRE(!(s.flags & PHX_FLAG_SHUFFLE_BATCHES), PHX_ERR_UNSUPPORTED,
                "shuffle_batches is not available on the 128-lane block engine");
    int32_t rc = make_engine_spec(s, E, seed, env_offset, &wspec, P::NWORDS, EnvWords<P>::value);
    if (rc != PHX_OK) return rc;
    rc = P::validate(s);
    if (rc != PHX_OK) return rc;
    PHX_REQUIRE(s.obs_dim <= P::OBS_DIM, PHX_ERR_INVALID, "obs_dim exceeds the family's OBS_DIM");
    int act_total = 0, resp_total = 0;
    for (int i = 0; i < s.n_agents; ++i) {
      int deg = 0;
      for (int r = 0; r < s.n_agents; ++r) deg += mask_bit(s.adjacency[i], r);
      const int ca = wide_cap<P>(true, s.agent_kind[i], deg, s.n_agents);
      const int cr = wide_cap<P>(false, s.agent_kind[i], deg, s.n_agents);
      PHX_REQUIRE(ca >= 0 && ca <= 255 && cr >= 0 && cr <= 255, PHX_ERR_UNSUPPORTED,
                  "an agent's per-round fan-out exceeds 255 messages");
      act_total += ca;
      resp_total += cr;
    }
    lay = wide_layout<P>(act_total, resp_total);
    smem = ((sizeof(WideSmem<P>) + 15) & ~(size_t)15) + (size_t)lay.bytes;
    PHX_REQUIRE(smem <= 200 * 1024, PHX_ERR_UNSUPPORTED,
                "message queues of this env class exceed the shared memory of a block");

    // done sets over agent SLOTS: four words per env (the base class sized them by n_strategic)
    cudaFree(d_term);
    cudaFree(d_trunc);
    d_term = d_trunc = nullptr;
    mask_words = WIDE_MW;
    const size_t mw = sizeof(uint32_t) * (size_t)E * WIDE_MW;
    PHX_CUDA(cudaMalloc(&d_term, mw));
    PHX_CUDA(cudaMalloc(&d_trunc, mw));
    PHX_CUDA(cudaMemset(d_term, 0, mw));
    PHX_CUDA(cudaMemset(d_trunc, 0, mw));

    PHX_CUDA(cudaMalloc(&d_spec, sizeof(WideSpec)));
    PHX_CUDA(cudaMemcpy(d_spec, &wspec, sizeof(WideSpec), cudaMemcpyHostToDevice));
    const size_t n = (size_t)E * WIDE_G;
    const size_t nw = P::NWORDS > 0 ? P::NWORDS : 1;
    PHX_CUDA(cudaMalloc(&d_state, sizeof(int32_t) * n * nw));
    PHX_CUDA(cudaMemset(d_state, 0, sizeof(int32_t) * n * nw));
    if (s.env_kind != PHX_ENV_BASE) {
      PHX_CUDA(cudaMalloc(&d_rcache, sizeof(float) * n));
      PHX_CUDA(cudaMemset(d_rcache, 0, sizeof(float) * n));
      PHX_CUDA(cudaMalloc(&d_rnone, mw));
      PHX_CUDA(cudaMemset(d_rnone, 0, mw));
      if (s.env_kind == PHX_ENV_FSM) {
        PHX_CUDA(cudaMalloc(&d_ocache, sizeof(float) * n * s.obs_dim));
        PHX_CUDA(cudaMemset(d_ocache, 0, sizeof(float) * n * s.obs_dim));
        PHX_CUDA(cudaMalloc(&d_ocached, mw));
        PHX_CUDA(cudaMemset(d_ocached, 0, mw));
      }
    }
    if (EnvWords<P>::value > 0) {  // env-level words start at zero (e.g. avg_price = 0.0)
      PHX_CUDA(cudaMalloc(&d_env, sizeof(int32_t) * (size_t)E * EnvWords<P>::value));
      PHX_CUDA(cudaMemset(d_env, 0, sizeof(int32_t) * (size_t)E * EnvWords<P>::value));
    }
    if (s.flags & PHX_FLAG_STOCHASTIC_NETWORK) {
      PHX_REQUIRE(s.n_base_connections >= 0 && s.n_base_connections <= PHX_MAX_BASE_CONNECTIONS,
                  PHX_ERR_INVALID, "n_base_connections out of range");
      std::vector<uint2> base((size_t)s.n_base_connections + 1);
      for (int c = 0; c < s.n_base_connections; ++c) {
        PHX_REQUIRE(s.base_u[c] < s.n_agents && s.base_v[c] < s.n_agents, PHX_ERR_INVALID,
                    "base connection names an agent slot outside the env");
        const double r = s.base_rate[c];
        PHX_REQUIRE(r == r, PHX_ERR_INVALID, "base connection rate is NaN");
        // uniform01 < r  <=>  d24 < ceil(r * 2^24); r * 2^24 is exact in float64
        const double scaled = std::ceil(std::min(std::max(r, 0.0), 1.0) * 16777216.0);
        base[c] = make_uint2((uint32_t)s.base_u[c] | ((uint32_t)s.base_v[c] << 8), (uint32_t)scaled);
      }
      n_base = s.n_base_connections;
      PHX_CUDA(cudaMalloc(&d_base, sizeof(uint2) * base.size()));
      PHX_CUDA(cudaMemcpy(d_base, base.data(), sizeof(uint2) * base.size(), cudaMemcpyHostToDevice));
      PHX_CUDA(cudaMalloc(&d_adj, sizeof(uint32_t) * n * WIDE_MW));
      PHX_CUDA(cudaMemset(d_adj, 0, sizeof(uint32_t) * n * WIDE_MW));
    }
    PHX_CUDA(cudaFuncSetAttribute(wide_step_kernel<P, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHX_CUDA(cudaFuncSetAttribute(wide_step_kernel<P, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHX_CUDA(cudaFuncSetAttribute(wide_reset_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    wide_init_kernel<<<(E + 255) / 256, 256>>>(E, d_hdr);
    PHX_CUDA(cudaGetLastError());
    // constructor-time agent state (PhantomEnv.__init__ ends with agent.reset(), env.py:122-124)
    rc = launch_reset(nullptr, nullptr, nullptr, 0, /*agents_only=*/true);
    if (rc != PHX_OK) return rc;
    PHX_CUDA(cudaDeviceSynchronize());
    name = "wide(G=128)";
    return PHX_OK;
  }

  WideArgs<P> make_args(int32_t T, const StepIO& io) const {
    WideArgs<P> a;
    a.spec = d_spec;
    a.T = T;
    a.lay = lay;
    a.hdr = d_hdr;
    a.term = d_term;
    a.trunc = d_trunc;
    a.state = d_state;
    a.reward_cache = d_rcache;
    a.reward_none = d_rnone;
    a.obs_cache = d_ocache;
    a.obs_cached = d_ocached;
    a.env_state = d_env;
    a.adj_env = d_adj;
    a.base_conn = d_base;
    a.n_base = n_base;
    a.io = io;
    a.faults = fault_sink();
    a.trace = trace_sink();
    return a;
  }

  int32_t launch_reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask, cudaStream_t stream,
                       bool agents_only) {
    StepIO io{};
    const WideArgs<P> a = make_args(1, io);
    wide_reset_kernel<P><<<E, WIDE_G, smem, stream>>>(a, env_mask, obs, obs_mask, agents_only);
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  int32_t reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                cudaStream_t stream) override {
    return launch_reset(env_mask, obs, obs_mask, stream, false);
  }

  int32_t rollout(int32_t T, const StepIO& io, cudaStream_t stream) override {
    if (tracking()) {
      const int32_t rc = ensure_trace(T);
      if (rc != PHX_OK) return rc;
    }
    const WideArgs<P> a = make_args(T, io);
    if (tracking())
      wide_step_kernel<P, true><<<E, WIDE_G, smem, stream>>>(a);
    else
      wide_step_kernel<P, false><<<E, WIDE_G, smem, stream>>>(a);
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  int32_t family_field(int32_t field, int32_t index, void** p, size_t* bytes) override {
    if (field == PHX_FIELD_ENV_STATE) {  // int32 [E]: env-level word `index`
      PHX_REQUIRE(index >= 0 && index < EnvWords<P>::value, PHX_ERR_INVALID,
                  "PHX_FIELD_ENV_STATE: this env class has no such env-level word");
      *p = d_env + (size_t)index * E;
      *bytes = sizeof(int32_t) * (size_t)E;
      return PHX_OK;
    }
    if (field == PHX_FIELD_ADJACENCY) {  // uint32 [E, 128, 4]
      PHX_REQUIRE(d_adj != nullptr, PHX_ERR_INVALID,
                  "PHX_FIELD_ADJACENCY needs PHX_FLAG_STOCHASTIC_NETWORK");
      *p = d_adj;
      *bytes = sizeof(uint32_t) * (size_t)E * WIDE_G * WIDE_MW;
      return PHX_OK;
    }
    const int w = field - PHX_FIELD_FAMILY;
    if (w >= 0 && w < P::NWORDS) {
      *p = d_state + (size_t)w * E * WIDE_G;
      *bytes = sizeof(int32_t) * (size_t)E * WIDE_G;
      return PHX_OK;
    }
    set_error("unknown family field " + std::to_string(field));
    return PHX_ERR_INVALID;
  }

  const char* exec_name() const override { return name.c_str(); }

  WideSpec wspec{};
  WideSpec* d_spec = nullptr;
  WideLayout lay{};
  size_t smem = 0;
  int32_t* d_state = nullptr;
  float* d_rcache = nullptr;
  uint32_t* d_rnone = nullptr;
  float* d_ocache = nullptr;
  uint32_t* d_ocached = nullptr;
  uint32_t* d_adj = nullptr;
  uint2* d_base = nullptr;
  int32_t* d_env = nullptr;  // [ENVW][E]
  int32_t n_base = 0;
  std::string name = "wide(G=128)";
};

// The family object of an engine-backed program: the block engine for env classes wider than a
// warp (or when PHX_EXEC_WIDE asks for it), the tile / thread engines otherwise.
template <class P>
inline Family* make_engine_family(const phx_spec& s) {
  if (s.n_agents > ENGINE_MAX_AGENTS || s.exec_mode == PHX_EXEC_WIDE) return new WideFamily<P>();
  return new EngineFamily<P>();
}

}  // namespace phx
