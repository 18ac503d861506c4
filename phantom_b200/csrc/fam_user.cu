// fam_user.cu -- PHX_FAMILY_USER: env classes whose device program lives in a cubin the CALLER
// compiled (phx_user.cuh), not in libphx.so.  This is what keeps the reference's plugin contract
// -- "an experiment brings its own agent classes" (phantom/agents.py:48-60) -- without a rebuild
// of the library: libphx owns the HBM state, the ABI and the launch logic, the user's unit owns
// the callbacks.  The kernels in the cubin are the thread-per-env engine (phx_engine1.cuh: at
// most 8 agents per env) and -- for width-independent programs (WIDE_OK, phx_user.cuh) -- the
// 128-lane block engine (phx_engine_wide.cuh: 9..128 agents) instantiated for the user's
// program; `phx_user_desc` says how many state words / view words / payload words it uses.
#include <cstring>
#include <string>

#include "phx_engine_host.cuh"
#include "phx_engine_wide_host.cuh"

namespace phx {
namespace {

struct UserTag {};  // EngineArgs<P> has the same layout for every P

class UserFamily final : public Family {
 public:
  explicit UserFamily(const char* cubin) : cubin_path(cubin ? cubin : "") {}
  ~UserFamily() override {
    cudaFree(d_state);
    cudaFree(d_rcache);
    cudaFree(d_rnone);
    cudaFree(d_ocache);
    cudaFree(d_ocached);
    cudaFree(d_env);
    cudaFree(d_carry);
    cudaFree(d_carry_n);
    if (lib) cudaLibraryUnload(lib);
  }

  int32_t init(const phx_spec& s) override {
    PHX_REQUIRE(!cubin_path.empty(), PHX_ERR_INVALID, "phx_create_user: cubin path is NULL");
    PHX_REQUIRE(s.n_agents <= ENGINE1_SLOTS, PHX_ERR_UNSUPPORTED,
                "user device programs run on the thread-per-env engine: at most 8 agents per env");
    PHX_REQUIRE(!(s.flags & (PHX_FLAG_STOCHASTIC_NETWORK | PHX_FLAG_SHUFFLE_BATCHES)),
                PHX_ERR_UNSUPPORTED,
                "user device programs: StochasticNetwork / shuffle_batches are not supported");
    PHX_CUDA(cudaLibraryLoadFromFile(&lib, cubin_path.c_str(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    PHX_CUDA(cudaLibraryGetKernel(&k_step, lib, "phx_user_step"));
    PHX_CUDA(cudaLibraryGetKernel(&k_step_tracked, lib, "phx_user_step_tracked"));
    PHX_CUDA(cudaLibraryGetKernel(&k_reset, lib, "phx_user_reset"));
    void* dptr = nullptr;
    size_t dbytes = 0;
    PHX_CUDA(cudaLibraryGetGlobal(&dptr, &dbytes, lib, "phx_user_desc"));
    PHX_REQUIRE(dbytes >= sizeof(desc), PHX_ERR_INVALID, "phx_user_desc has the wrong size");
    PHX_CUDA(cudaMemcpy(desc, dptr, sizeof(desc), cudaMemcpyDeviceToHost));
    PHX_REQUIRE(desc[0] == 0x50485855, PHX_ERR_INVALID,
                "the cubin does not end with PHX_USER_PROGRAM(...) of this library version");
    nwords = desc[1]; vw = desc[2]; pw = desc[3]; act_dim = desc[4]; obs_dim = desc[5];
    q1cap = desc[6]; envw = desc[7]; reset_smem = desc[8];
    PHX_REQUIRE(desc[9] == 0, PHX_ERR_UNSUPPORTED,
                "user device programs with a handle_batch override (BATCHED) need the tile engine");
    PHX_REQUIRE(vw <= 1 && q1cap > 0, PHX_ERR_UNSUPPORTED,
                "user device programs: VW <= 1 and Q1CAP > 0 (thread-per-env engine)");
    PHX_REQUIRE(s.obs_dim <= obs_dim && s.act_dim == act_dim, PHX_ERR_INVALID,
                "spec obs_dim / act_dim do not match the device program");
    int32_t rc = make_engine_spec(s, E, seed, env_offset, &espec, nwords, envw);
    if (rc != PHX_OK) return rc;
    const size_t n = (size_t)E * ENGINE1_SLOTS;
    const size_t words = (size_t)(nwords > 0 ? nwords : 1);
    PHX_CUDA(cudaMalloc(&d_state, sizeof(int32_t) * n * words));
    PHX_CUDA(cudaMemset(d_state, 0, sizeof(int32_t) * n * words));
    if (s.env_kind != PHX_ENV_BASE) {
      PHX_CUDA(cudaMalloc(&d_rcache, sizeof(float) * n));
      PHX_CUDA(cudaMemset(d_rcache, 0, sizeof(float) * n));
      PHX_CUDA(cudaMalloc(&d_rnone, sizeof(uint32_t) * (size_t)E));
      PHX_CUDA(cudaMemset(d_rnone, 0, sizeof(uint32_t) * (size_t)E));
      if (s.env_kind == PHX_ENV_FSM) {
        PHX_CUDA(cudaMalloc(&d_ocache, sizeof(float) * n * s.obs_dim));
        PHX_CUDA(cudaMemset(d_ocache, 0, sizeof(float) * n * s.obs_dim));
        PHX_CUDA(cudaMalloc(&d_ocached, sizeof(uint32_t) * (size_t)E));
        PHX_CUDA(cudaMemset(d_ocached, 0, sizeof(uint32_t) * (size_t)E));
      }
    }
    if (envw > 0) {
      PHX_CUDA(cudaMalloc(&d_env, sizeof(int32_t) * (size_t)E * envw));
      PHX_CUDA(cudaMemset(d_env, 0, sizeof(int32_t) * (size_t)E * envw));
    }
    if (waiting_mail_possible(s)) {  // mail that waits across steps (see EngineFamily::init)
      PHX_CUDA(cudaMalloc(&d_carry_n, sizeof(int32_t) * (size_t)E));
      PHX_CUDA(cudaMemset(d_carry_n, 0, sizeof(int32_t) * (size_t)E));
      PHX_CUDA(cudaMalloc(&d_carry, sizeof(int32_t) * (size_t)E * q1cap * (1 + pw)));
    }
    {  // env header: episode becomes 0 on the first reset
      std::vector<int4> h((size_t)E, make_int4(0, -1, 0, 0));
      PHX_CUDA(cudaMemcpy(d_hdr, h.data(), sizeof(int4) * (size_t)E, cudaMemcpyHostToDevice));
    }
    for (cudaKernel_t k : {k_step, k_step_tracked})
      PHX_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)step_smem()));
    PHX_CUDA(cudaFuncSetAttribute((const void*)k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  reset_smem));
    // constructor-time agent state (PhantomEnv.__init__ ends with agent.reset(), env.py:122-124)
    rc = launch_reset(nullptr, nullptr, nullptr, 0, /*agents_only=*/true);
    if (rc != PHX_OK) return rc;
    PHX_CUDA(cudaDeviceSynchronize());
    name = "thread-per-env(G=8, user program)";
    return PHX_OK;
  }

  size_t step_smem() const {
    const Engine1Layout lay =
        engine1_layout_rt(nwords, vw, act_dim, pw, spec.n_agents, spec.n_strategic, q1cap,
                          spec.env_kind != PHX_ENV_BASE, spec.obs_dim, /*with_stage=*/false);
    return sizeof(int32_t) * (size_t)lay.words;
  }

  EngineArgs<UserTag> make_args(int32_t T, const StepIO& io) const {
    EngineArgs<UserTag> a;
    std::memset(&a, 0, sizeof(a));
    a.spec = espec;
    a.T = T;
    a.qcap = q1cap;
    a.hdr = d_hdr;
    a.term = d_term;
    a.trunc = d_trunc;
    a.state = d_state;
    a.reward_cache = d_rcache;
    a.reward_none = d_rnone;
    a.obs_cache = d_ocache;
    a.obs_cached = d_ocached;
    a.env_state = d_env;
    a.io = io;
    a.faults = fault_sink();
    a.trace = trace_sink();
    a.carry_n = d_carry_n;
    a.carry = d_carry;
    return a;
  }

  int32_t launch_reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask, cudaStream_t stream,
                       bool agents_only) {
    StepIO io{};
    EngineArgs<UserTag> a = make_args(1, io);
    constexpr int TPB = ENGINE_BLOCK / 8;
    void* args[] = {(void*)&a, (void*)&env_mask, (void*)&obs, (void*)&obs_mask, (void*)&agents_only};
    PHX_CUDA(cudaLaunchKernel((const void*)k_reset, dim3((E + TPB - 1) / TPB), dim3(ENGINE_BLOCK),
                              args, (size_t)reset_smem, stream));
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
    EngineArgs<UserTag> a = make_args(T, io);
    void* args[] = {(void*)&a};
    const int grid = (E + ENGINE1_BLOCK - 1) / ENGINE1_BLOCK;
    PHX_CUDA(cudaLaunchKernel((const void*)(tracking() ? k_step_tracked : k_step), dim3(grid),
                              dim3(ENGINE1_BLOCK), args, step_smem(), stream));
    return PHX_OK;
  }

  int32_t family_field(int32_t field, int32_t index, void** p, size_t* bytes) override {
    if (field == PHX_FIELD_ENV_STATE) {
      PHX_REQUIRE(index >= 0 && index < envw, PHX_ERR_INVALID, "no such env-level word");
      *p = d_env + (size_t)index * E;
      *bytes = sizeof(int32_t) * (size_t)E;
      return PHX_OK;
    }
    const int w = field - PHX_FIELD_FAMILY;
    if (w >= 0 && w < nwords) {
      *p = d_state + (size_t)w * E * ENGINE1_SLOTS;
      *bytes = sizeof(int32_t) * (size_t)E * ENGINE1_SLOTS;
      return PHX_OK;
    }
    set_error("user family: unknown field " + std::to_string(field));
    return PHX_ERR_INVALID;
  }

  const char* exec_name() const override { return name.c_str(); }

 private:
  std::string cubin_path, name;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t k_step = nullptr, k_step_tracked = nullptr, k_reset = nullptr;
  int32_t desc[12] = {};
  int nwords = 0, vw = 0, pw = 1, act_dim = 1, obs_dim = 1, q1cap = 0, envw = 0, reset_smem = 0;
  EngineSpec espec{};
  int32_t* d_state = nullptr;
  float* d_rcache = nullptr;
  uint32_t* d_rnone = nullptr;
  float* d_ocache = nullptr;
  uint32_t* d_ocached = nullptr;
  int32_t* d_env = nullptr;
  int32_t* d_carry_n = nullptr;
  int32_t* d_carry = nullptr;
};

// A user's program on the block engine: WideFamilyCore with the kernels of the cubin.
class UserWideFamily final : public WideFamilyCore {
 public:
  explicit UserWideFamily(const char* cubin) : cubin_path(cubin ? cubin : "") {}
  ~UserWideFamily() override {
    if (lib) cudaLibraryUnload(lib);
  }

  int32_t init(const phx_spec& s) override {
    PHX_REQUIRE(!cubin_path.empty(), PHX_ERR_INVALID, "phx_create_user: cubin path is NULL");
    PHX_CUDA(cudaLibraryLoadFromFile(&lib, cubin_path.c_str(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaKernel_t ks = nullptr, kt = nullptr, kr = nullptr;
    PHX_CUDA(cudaLibraryGetKernel(&ks, lib, "phx_user_wide_step"));
    PHX_CUDA(cudaLibraryGetKernel(&kt, lib, "phx_user_wide_step_tracked"));
    PHX_CUDA(cudaLibraryGetKernel(&kr, lib, "phx_user_wide_reset"));
    void* dptr = nullptr;
    size_t dbytes = 0;
    int32_t desc[12] = {};
    PHX_CUDA(cudaLibraryGetGlobal(&dptr, &dbytes, lib, "phx_user_desc"));
    PHX_REQUIRE(dbytes >= sizeof(desc), PHX_ERR_INVALID, "phx_user_desc has the wrong size");
    PHX_CUDA(cudaMemcpy(desc, dptr, sizeof(desc), cudaMemcpyDeviceToHost));
    PHX_REQUIRE(desc[0] == 0x50485855, PHX_ERR_INVALID,
                "the cubin does not end with PHX_USER_PROGRAM(...) of this library version");
    PHX_REQUIRE(desc[10] > 0, PHX_ERR_UNSUPPORTED,
                "env classes of more than 8 agents run on the 128-lane block engine: the user "
                "program must be width independent and declare WIDE_OK (csrc/phx_user.cuh)");
    PHX_REQUIRE(s.obs_dim <= desc[5] && s.act_dim == desc[4], PHX_ERR_INVALID,
                "spec obs_dim / act_dim do not match the device program");
    info.nwords = desc[1];
    info.pw = desc[3];
    info.obs_dim = desc[5];
    info.envw = desc[7];
    info.smem_fixed = (size_t)desc[10];
    const int actcap = desc[11] & 0xFFFF, respcap = (desc[11] >> 16) & 0xFFFF;
    info.cap = [actcap, respcap](bool acting, int, int deg, int) {
      const int cap = acting ? actcap : respcap, want = deg > 0 ? deg : 1;
      return want < cap ? want : cap;  // wide_cap's default: one message per neighbour and round
    };
    info.k_step = (const void*)ks;
    info.k_step_tracked = (const void*)kt;
    info.k_reset = (const void*)kr;
    name = "wide(G=128, user program)";
    return core_init(s);
  }

 private:
  std::string cubin_path;
  cudaLibrary_t lib = nullptr;
};

}  // namespace

Family* make_user_family(const char* cubin_path, const phx_spec& s) {
  if (s.n_agents > ENGINE1_SLOTS || s.exec_mode == PHX_EXEC_WIDE) return new UserWideFamily(cubin_path);
  return new UserFamily(cubin_path);
}

}  // namespace phx
