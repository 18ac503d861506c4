// phx_engine_wide_host.cuh -- host side of the 128-lane block engine (phx_engine_wide.cuh): a
// Family that owns the state of E envs of a device program with 33..128 agents (or fewer, when
// PHX_EXEC_WIDE is forced) and launches wide_step_kernel, one block per env.
//   WideFamilyCore  everything that does not depend on the program type: buffers, launches,
//                   fields.  What it needs to know about the program is a WideProgramInfo
//                   (sizes, queue capacities, kernel handles) -- so the same class serves the
//                   families compiled into libphx (WideFamily<P>) and a user's program that
//                   arrives as a cubin (fam_user.cu).
// Field mapping: PHX_FIELD_FAMILY + w = state word w, int32 [E, 128] (slot-major inside an env);
// PHX_FIELD_TERMINATED / _TRUNCATED = uint32 [E, 4] bitmasks over agent slots;
// PHX_FIELD_ADJACENCY = uint32 [E, 128, 4] rows; PHX_FIELD_ENV_STATE = int32 [E] per word.
#pragma once
#include <functional>

#include "phx_engine_host.cuh"
#include "phx_engine_wide.cuh"

namespace phx {

static __global__ void wide_init_kernel(int E, int4* hdr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) hdr[i] = make_int4(0, -1, 0, 0);  // episode becomes 0 on the first reset
}

struct WideTag {};  // WideArgs<P> has the same layout for every P

struct WideProgramInfo {
  int nwords = 0, envw = 0, pw = 1, obs_dim = 1;
  size_t smem_fixed = 0;  // sizeof(WideSmem<P>)
  // segment capacity of one agent: (acting phase?, kind, degree in the full graph, n_agents)
  std::function<int(bool, int, int, int)> cap;
  const void* k_step = nullptr;          // wide_step_kernel<P, false> (or a cudaKernel_t)
  const void* k_step_tracked = nullptr;  // wide_step_kernel<P, true>: tracking and / or shuffle_batches
  const void* k_reset = nullptr;         // wide_reset_kernel<P>
};

class WideFamilyCore : public Family {
 public:
  ~WideFamilyCore() override {
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

  // `info` is filled by the subclass before this runs.
  int32_t core_init(const phx_spec& s) {
    int32_t rc = make_engine_spec(s, E, seed, env_offset, &wspec, info.nwords, info.envw);
    if (rc != PHX_OK) return rc;
    PHX_REQUIRE(s.obs_dim <= info.obs_dim, PHX_ERR_INVALID, "obs_dim exceeds the family's OBS_DIM");
    int act_total = 0, resp_total = 0;
    for (int i = 0; i < s.n_agents; ++i) {
      int deg = 0;
      for (int r = 0; r < s.n_agents; ++r) deg += mask_bit(s.adjacency[i], r);
      const int ca = info.cap(true, s.agent_kind[i], deg, s.n_agents);
      const int cr = info.cap(false, s.agent_kind[i], deg, s.n_agents);
      PHX_REQUIRE(ca >= 0 && ca <= 255 && cr >= 0 && cr <= 255, PHX_ERR_UNSUPPORTED,
                  "an agent's per-round fan-out exceeds 255 messages");
      act_total += ca;
      resp_total += cr;
    }
    lay = wide_layout_rt(info.pw, act_total, resp_total);
    smem = ((info.smem_fixed + 15) & ~(size_t)15) + (size_t)lay.bytes;
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
    const size_t nw = info.nwords > 0 ? info.nwords : 1;
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
    if (info.envw > 0) {  // env-level words start at zero (e.g. avg_price = 0.0)
      PHX_CUDA(cudaMalloc(&d_env, sizeof(int32_t) * (size_t)E * info.envw));
      PHX_CUDA(cudaMemset(d_env, 0, sizeof(int32_t) * (size_t)E * info.envw));
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
    for (const void* k : {info.k_step, info.k_step_tracked, info.k_reset})
      PHX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wide_init_kernel<<<(E + 255) / 256, 256>>>(E, d_hdr);
    PHX_CUDA(cudaGetLastError());
    // constructor-time agent state (PhantomEnv.__init__ ends with agent.reset(), env.py:122-124)
    rc = launch_reset(nullptr, nullptr, nullptr, 0, /*agents_only=*/true);
    if (rc != PHX_OK) return rc;
    PHX_CUDA(cudaDeviceSynchronize());
    return PHX_OK;
  }

  WideArgs<WideTag> make_args(int32_t T, const StepIO& io) const {
    WideArgs<WideTag> a;
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
    WideArgs<WideTag> a = make_args(1, io);
    void* args[] = {(void*)&a, (void*)&env_mask, (void*)&obs, (void*)&obs_mask, (void*)&agents_only};
    PHX_CUDA(cudaLaunchKernel(info.k_reset, dim3(E), dim3(WIDE_G), args, smem, stream));
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
    WideArgs<WideTag> a = make_args(T, io);
    void* args[] = {(void*)&a};
    const bool full = tracking() || (spec.flags & PHX_FLAG_SHUFFLE_BATCHES) != 0;
    PHX_CUDA(cudaLaunchKernel(full ? info.k_step_tracked : info.k_step, dim3(E), dim3(WIDE_G), args,
                              smem, stream));
    return PHX_OK;
  }

  int32_t family_field(int32_t field, int32_t index, void** p, size_t* bytes) override {
    if (field == PHX_FIELD_ENV_STATE) {  // int32 [E]: env-level word `index`
      PHX_REQUIRE(index >= 0 && index < info.envw, PHX_ERR_INVALID,
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
    if (w >= 0 && w < info.nwords) {
      *p = d_state + (size_t)w * E * WIDE_G;
      *bytes = sizeof(int32_t) * (size_t)E * WIDE_G;
      return PHX_OK;
    }
    set_error("unknown family field " + std::to_string(field));
    return PHX_ERR_INVALID;
  }

  const char* exec_name() const override { return name.c_str(); }

  WideProgramInfo info;
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

// The block engine for a program compiled into libphx.
template <class P>
class WideFamily : public WideFamilyCore {
 public:
  int32_t init(const phx_spec& s) override {
    const int32_t rc = P::validate(s);
    if (rc != PHX_OK) return rc;
    info.nwords = P::NWORDS;
    info.envw = EnvWords<P>::value;
    info.pw = P::PW;
    info.obs_dim = P::OBS_DIM;
    info.smem_fixed = sizeof(WideSmem<P>);
    info.cap = [](bool acting, int kind, int deg, int n) { return wide_cap<P>(acting, kind, deg, n); };
    info.k_step = (const void*)wide_step_kernel<P, false>;
    info.k_step_tracked = (const void*)wide_step_kernel<P, true>;
    info.k_reset = (const void*)wide_reset_kernel<P>;
    return core_init(s);
  }
};

// The family object of an engine-backed program: the block engine for env classes wider than a
// warp (or when PHX_EXEC_WIDE asks for it), the tile / thread engines otherwise.
template <class P>
inline Family* make_engine_family(const phx_spec& s) {
  if (s.n_agents > ENGINE_MAX_AGENTS || s.exec_mode == PHX_EXEC_WIDE) return new WideFamily<P>();
  return new EngineFamily<P>();
}

}  // namespace phx
