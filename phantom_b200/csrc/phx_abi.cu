// phx_abi.cu -- the C ABI of libphx (include/phx.h): handle management, argument checks,
// dispatch to the device program families, host-buffer variants, field/trace/fault access.
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "phx_family.h"
#include "phx_sc_wire.h"

namespace phx {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

Family::~Family() {
  if (device < 0) return;  // never initialised on a device (phx_selftest_jit_source)
  cudaSetDevice(device);
  cudaFree(d_hdr);
  cudaFree(d_term);
  cudaFree(d_trunc);
  cudaFree(d_err);
  cudaFree(d_nfaults);
  cudaFree(d_trace);
  cudaFree(d_trace_cnt);
  cudaFree(d_stage);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (copy_in) cudaStreamDestroy(copy_in);
  if (copy_out) cudaStreamDestroy(copy_out);
  for (int i = 0; i < 16; ++i) {
    if (ev_in[i]) cudaEventDestroy(ev_in[i]);
    if (ev_k[i]) cudaEventDestroy(ev_k[i]);
    if (ev_out[i]) cudaEventDestroy(ev_out[i]);
  }
}

int32_t Family::base_init(const phx_spec& s, int32_t num_envs, int32_t dev, uint64_t sd,
                          int64_t off) {
  spec = s;
  E = num_envs;
  device = dev;
  seed = sd;
  env_offset = off;
  mask_words = (s.n_strategic + 31) / 32;
  if (mask_words < 1) mask_words = 1;
  PHX_CUDA(cudaSetDevice(device));
  PHX_CUDA(cudaMalloc(&d_hdr, sizeof(int4) * (size_t)E));
  PHX_CUDA(cudaMalloc(&d_term, sizeof(uint32_t) * (size_t)E * mask_words));
  PHX_CUDA(cudaMalloc(&d_trunc, sizeof(uint32_t) * (size_t)E * mask_words));
  PHX_CUDA(cudaMalloc(&d_err, sizeof(uint32_t) * (size_t)E));
  PHX_CUDA(cudaMalloc(&d_nfaults, sizeof(uint32_t)));
  PHX_CUDA(cudaMemset(d_hdr, 0, sizeof(int4) * (size_t)E));
  PHX_CUDA(cudaMemset(d_term, 0, sizeof(uint32_t) * (size_t)E * mask_words));
  PHX_CUDA(cudaMemset(d_trunc, 0, sizeof(uint32_t) * (size_t)E * mask_words));
  PHX_CUDA(cudaMemset(d_err, 0, sizeof(uint32_t) * (size_t)E));
  PHX_CUDA(cudaMemset(d_nfaults, 0, sizeof(uint32_t)));
  if (tracking()) {
    PHX_REQUIRE(spec.trace_capacity > 0, PHX_ERR_INVALID,
                "PHX_FLAG_TRACK_MESSAGES needs trace_capacity > 0");
    PHX_CUDA(cudaMalloc(&d_trace, sizeof(int4) * (size_t)E * spec.trace_capacity));
    PHX_CUDA(cudaMalloc(&d_trace_cnt, sizeof(int32_t) * (size_t)E));
    PHX_CUDA(cudaMemset(d_trace_cnt, 0, sizeof(int32_t) * (size_t)E));
  }
  PHX_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  PHX_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
  PHX_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
  for (int i = 0; i < 16; ++i) {
    PHX_CUDA(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
    PHX_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
    PHX_CUDA(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
  }
  return PHX_OK;
}

int32_t Family::field_ptr(int32_t field, int32_t index, void** p, size_t* bytes) {
  switch (field) {
    case PHX_FIELD_TERMINATED:
      *p = d_term; *bytes = sizeof(uint32_t) * (size_t)E * mask_words; return PHX_OK;
    case PHX_FIELD_TRUNCATED:
      *p = d_trunc; *bytes = sizeof(uint32_t) * (size_t)E * mask_words; return PHX_OK;
    case PHX_FIELD_ERROR:
      *p = d_err; *bytes = sizeof(uint32_t) * (size_t)E; return PHX_OK;
    default:
      break;
  }
  if (field >= PHX_FIELD_FAMILY || field == PHX_FIELD_ADJACENCY || field == PHX_FIELD_ENV_STATE)
    return family_field(field, index, p, bytes);
  set_error("unknown field id " + std::to_string(field));
  return PHX_ERR_INVALID;
}

// Message tracing inside a T-step launch: one slab of `trace_capacity` rows per (step, env).
int32_t Family::ensure_trace(int32_t T) {
  PHX_REQUIRE(tracking() && T >= 1, PHX_ERR_INVALID, "handle was created without tracking");
  const size_t rows = (size_t)T * E * spec.trace_capacity;
  PHX_REQUIRE(rows * sizeof(int4) <= ((size_t)8 << 30), PHX_ERR_INVALID,
              "message trace of this launch exceeds 8 GiB: track fewer envs or steps");
  if (T > trace_T_cap) {
    PHX_CUDA(cudaDeviceSynchronize());
    cudaFree(d_trace);
    cudaFree(d_trace_cnt);
    d_trace = nullptr;
    d_trace_cnt = nullptr;
    PHX_CUDA(cudaMalloc(&d_trace, sizeof(int4) * rows));
    PHX_CUDA(cudaMalloc(&d_trace_cnt, sizeof(int32_t) * (size_t)T * E));
    PHX_CUDA(cudaMemset(d_trace_cnt, 0, sizeof(int32_t) * (size_t)T * E));
    trace_T_cap = T;
  }
  trace_T = T;
  return PHX_OK;
}

int32_t Family::ensure_stage(size_t total) {
  if (total > stage_bytes) {
    PHX_CUDA(cudaStreamSynchronize(own_stream));
    PHX_CUDA(cudaStreamSynchronize(copy_in));
    PHX_CUDA(cudaStreamSynchronize(copy_out));
    if (d_stage) PHX_CUDA(cudaFree(d_stage));
    d_stage = nullptr;
    stage_bytes = 0;
    PHX_CUDA(cudaMalloc(&d_stage, total));
    stage_bytes = total;
  }
  return PHX_OK;
}

int32_t Family::rollout_host(int32_t T, const StepIO& h) {
  const float* actions = h.actions;
  const uint8_t* action_mask = h.action_mask;
  float *obs = h.obs, *reward = h.reward;
  uint8_t *obs_mask = h.obs_mask, *reward_mask = h.reward_mask, *term = h.term, *trunc = h.trunc,
          *all_done = h.all_done;
  Family* f = this;
  const size_t n = (size_t)T * f->E * (size_t)(f->spec.n_strategic > 0 ? f->spec.n_strategic : 1);
  const size_t S = f->spec.n_strategic;
  const size_t TE = (size_t)T * f->E;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  // carve one staging block: actions | action_mask | obs | obs_mask | reward | reward_mask |
  // term | trunc | all_done
  const size_t b_act = up(TE * S * f->spec.act_dim * sizeof(float));
  const size_t b_am = up(TE * S);
  const size_t b_obs = up(TE * S * f->spec.obs_dim * sizeof(float));
  const size_t b_u8 = up(TE * S);
  const size_t b_rew = up(TE * S * sizeof(float));
  const size_t b_all = up(TE * 2);
  const size_t total = b_act + b_am + b_obs + 4 * b_u8 + b_rew + b_all;
  (void)n;
  {
    const int32_t rc = ensure_stage(total);
    if (rc != PHX_OK) return rc;
  }
  uint8_t* p = (uint8_t*)f->d_stage;
  float* d_act = (float*)p; p += b_act;
  uint8_t* d_am = p; p += b_am;
  float* d_obs = (float*)p; p += b_obs;
  uint8_t* d_om = p; p += b_u8;
  float* d_rew = (float*)p; p += b_rew;
  uint8_t* d_rm = p; p += b_u8;
  uint8_t* d_term = p; p += b_u8;
  uint8_t* d_trunc = p; p += b_u8;
  uint8_t* d_all = p;
  phx::StepIO io{d_act,
                 action_mask ? d_am : nullptr,
                 obs ? d_obs : nullptr,
                 obs_mask ? d_om : nullptr,
                 reward ? d_rew : nullptr,
                 reward_mask ? d_rm : nullptr,
                 term ? d_term : nullptr,
                 trunc ? d_trunc : nullptr,
                 all_done ? d_all : nullptr};
  const size_t E = (size_t)f->E, A = (size_t)f->spec.act_dim, O = (size_t)f->spec.obs_dim;
  // Chunked 3-stage pipeline (H2D | kernel | D2H on three streams), chunked along TIME: the rows
  // [t0, t1) of every [T, E, ...] plane are one contiguous block, so each transfer is a single
  // large 1-D copy (env-range chunks needed pitched 2-D copies of 16-98 KB rows, which PCIe
  // moves ~10 % slower), and every family can do it -- a chunk is just a shorter rollout that
  // continues from the state the previous launch left in HBM.  PCIe is full duplex: the input
  // copy of chunk c+1 and the output copy of chunk c-1 overlap the kernel of chunk c.
  const int chunks = (T >= 8 && TE >= 65536 && !f->tracking()) ? 8 : 1;
  auto copy1d = [&](void* dst, const void* src, size_t row_bytes, size_t t0, size_t nt,
                    cudaMemcpyKind k, cudaStream_t st) {
    if (row_bytes == 0 || nt == 0) return cudaSuccess;
    return cudaMemcpyAsync((char*)dst + t0 * row_bytes, (const char*)src + t0 * row_bytes,
                           nt * row_bytes, k, st);
  };
  // Chunk boundaries grow geometrically at the front: the first output copy can only start after
  // the first chunk's input copy and kernel, so a short first chunk shortens the pipeline fill
  // (uniform eighths: 3.3 MB of actions = 80 us before the D2H engine has anything to do).
  static const int kFrac[9] = {0, 2, 6, 14, 28, 46, 64, 82, 100};  // percent of T, cumulative
  auto bound = [&](int c) {
    return chunks == 1 ? (size_t)(c ? T : 0) : ((size_t)T * kFrac[c] + 50) / 100;
  };
  for (int c = 0; c < chunks; ++c) {
    const size_t t0 = bound(c), t1 = bound(c + 1);
    const size_t nt = t1 - t0;
    if (nt == 0) continue;
    if (actions)
      PHX_CUDA(copy1d(d_act, actions, E * S * A * sizeof(float), t0, nt, cudaMemcpyHostToDevice, f->copy_in));
    if (action_mask)
      PHX_CUDA(copy1d(d_am, action_mask, E * S, t0, nt, cudaMemcpyHostToDevice, f->copy_in));
    PHX_CUDA(cudaEventRecord(f->ev_in[c], f->copy_in));
    PHX_CUDA(cudaStreamWaitEvent(f->own_stream, f->ev_in[c], 0));
    phx::StepIO ic = io;  // the chunk's rows of every plane
    ic.actions = io.actions + t0 * E * S * A;
    if (ic.action_mask) ic.action_mask += t0 * E * S;
    if (ic.obs) ic.obs += t0 * E * S * O;
    if (ic.obs_mask) ic.obs_mask += t0 * E * S;
    if (ic.reward) ic.reward += t0 * E * S;
    if (ic.reward_mask) ic.reward_mask += t0 * E * S;
    if (ic.term) ic.term += t0 * E * S;
    if (ic.trunc) ic.trunc += t0 * E * S;
    if (ic.all_done) ic.all_done += t0 * E * 2;
    const int32_t rc = f->rollout((int32_t)nt, ic, f->own_stream);
    if (rc != PHX_OK) return rc;
    PHX_CUDA(cudaEventRecord(f->ev_k[c], f->own_stream));
    PHX_CUDA(cudaStreamWaitEvent(f->copy_out, f->ev_k[c], 0));
    cudaStream_t so = f->copy_out;
    if (obs) PHX_CUDA(copy1d(obs, d_obs, E * S * O * sizeof(float), t0, nt, cudaMemcpyDeviceToHost, so));
    if (obs_mask) PHX_CUDA(copy1d(obs_mask, d_om, E * S, t0, nt, cudaMemcpyDeviceToHost, so));
    if (reward) PHX_CUDA(copy1d(reward, d_rew, E * S * sizeof(float), t0, nt, cudaMemcpyDeviceToHost, so));
    if (reward_mask) PHX_CUDA(copy1d(reward_mask, d_rm, E * S, t0, nt, cudaMemcpyDeviceToHost, so));
    if (term) PHX_CUDA(copy1d(term, d_term, E * S, t0, nt, cudaMemcpyDeviceToHost, so));
    if (trunc) PHX_CUDA(copy1d(trunc, d_trunc, E * S, t0, nt, cudaMemcpyDeviceToHost, so));
    if (all_done) PHX_CUDA(copy1d(all_done, d_all, E * 2, t0, nt, cudaMemcpyDeviceToHost, so));
  }
  PHX_CUDA(cudaStreamSynchronize(f->copy_out));
  PHX_CUDA(cudaStreamSynchronize(f->own_stream));
  return PHX_OK;
}


}  // namespace phx

using phx::Family;
using phx::set_error;

// K8 field_reduce: word `col` of every `width`-word row of an int32 column -> sum / min / max.
// Grid-stride, warp-shuffle reduction, one atomic triple per block.
__global__ void phx_field_reduce_kernel(const int32_t* col_base, int E, int width, int col,
                                        unsigned long long* sum, int32_t* mn, int32_t* mx) {
  long long s = 0;
  int32_t lo = INT32_MAX, hi = INT32_MIN;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int32_t v = col_base[(size_t)e * width + col];
    s += v;
    lo = min(lo, v);
    hi = max(hi, v);
  }
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
    lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, off));
    hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sum, (unsigned long long)s);
    atomicMin(mn, lo);
    atomicMax(mx, hi);
  }
}

struct phx_env {
  Family* fam;
};

static int32_t check_spec(const phx_spec* s) {
  PHX_REQUIRE(s != nullptr, PHX_ERR_INVALID, "spec is NULL");
  PHX_REQUIRE(s->struct_size == sizeof(phx_spec), PHX_ERR_INVALID,
              "phx_spec.struct_size mismatch (ABI skew): got " + std::to_string(s->struct_size) +
                  ", library has " + std::to_string(sizeof(phx_spec)));
  PHX_REQUIRE(s->n_agents >= 1 && s->n_agents <= PHX_MAX_AGENTS, PHX_ERR_INVALID,
              "n_agents out of range");
  PHX_REQUIRE(s->n_strategic >= 0 && s->n_strategic <= s->n_agents, PHX_ERR_INVALID,
              "n_strategic out of range");
  PHX_REQUIRE(s->n_payload_types >= 0 && s->n_payload_types <= PHX_MAX_TYPES, PHX_ERR_INVALID,
              "n_payload_types out of range");
  PHX_REQUIRE(s->num_steps >= 1, PHX_ERR_INVALID, "num_steps must be >= 1");
  PHX_REQUIRE(s->round_limit >= -1, PHX_ERR_INVALID, "round_limit must be >= -1");
  int n_strat = 0;
  for (int i = 0; i < s->n_agents; ++i) {
    if (s->strategic_index[i] >= 0) {
      PHX_REQUIRE(s->strategic_index[i] == n_strat, PHX_ERR_INVALID,
                  "strategic_index must number strategic slots 0..S-1 in slot order");
      ++n_strat;
    }
  }
  PHX_REQUIRE(n_strat == s->n_strategic, PHX_ERR_INVALID, "n_strategic != #strategic slots");
  if (s->env_kind == PHX_ENV_FSM) {
    PHX_REQUIRE(s->n_stages >= 1 && s->n_stages <= PHX_MAX_STAGES, PHX_ERR_INVALID,
                "n_stages out of range");
    PHX_REQUIRE(s->initial_stage >= 0 && s->initial_stage < s->n_stages, PHX_ERR_INVALID,
                "initial_stage out of range");
  }
  if (s->flags & PHX_FLAG_STOCHASTIC_NETWORK) {
    // (33..128 agents: the block engine, for the families that run on it -- the tile engine
    // refuses wider specs itself)
    PHX_REQUIRE(s->n_base_connections >= 0 && s->n_base_connections <= PHX_MAX_BASE_CONNECTIONS,
                PHX_ERR_INVALID, "n_base_connections out of range");
  }
  return PHX_OK;
}

extern "C" {

int32_t phx_abi_version(void) { return PHX_ABI_VERSION; }

uint32_t phx_sizeof_spec(void) { return (uint32_t)sizeof(phx_spec); }

const char* phx_last_error(void) { return phx::g_last_error.c_str(); }

int32_t phx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static Family* make_family(const phx_spec* spec) {
  switch (spec->family) {
    case PHX_FAMILY_SUPPLY_CHAIN: return phx::make_supply_chain_family(*spec);
    case PHX_FAMILY_MOCK: return phx::make_mock_family(*spec);
    case PHX_FAMILY_MARKET: return phx::make_market_family(*spec);
    case PHX_FAMILY_STACKELBERG: return phx::make_stackelberg_family(*spec);
    case PHX_FAMILY_DENSE: return phx::make_dense_family(*spec);
    case PHX_FAMILY_SUPPLY_CHAIN2: return phx::make_supply_chain2_family(*spec);
    case PHX_FAMILY_SIMPLE_MARKET: return phx::make_simple_market_family(*spec);
    case PHX_FAMILY_DIGITAL_ADS: return phx::make_digital_ads_family(*spec);
    default: return nullptr;
  }
}

int32_t phx_selftest_jit_source(const phx_spec* spec, int32_t num_envs, uint64_t seed, char* buf,
                                uint64_t buf_bytes, uint64_t* needed) {
  int32_t rc = check_spec(spec);
  if (rc != PHX_OK) return rc;
  Family* fam = make_family(spec);
  PHX_REQUIRE(fam != nullptr, PHX_ERR_UNSUPPORTED, "no device program for this family");
  fam->spec = *spec;  // (no base_init / init: nothing here touches a device)
  fam->E = num_envs;
  fam->seed = seed;
  std::string text;
  rc = fam->jit_source_offline(text);
  fam->device = -1;
  delete fam;
  if (rc != PHX_OK) return rc;
  if (needed) *needed = (uint64_t)text.size() + 1;
  if (buf == nullptr || buf_bytes == 0) return PHX_OK;
  PHX_REQUIRE(buf_bytes >= text.size() + 1, PHX_ERR_INVALID, "buffer too small for the source text");
  std::memcpy(buf, text.c_str(), text.size() + 1);
  return PHX_OK;
}

static int32_t create_common(const phx_spec* spec, const char* cubin_path, int32_t num_envs,
                             int32_t device, uint64_t seed, int64_t env_offset, phx_env** out) {
  PHX_REQUIRE(out != nullptr, PHX_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int32_t rc = check_spec(spec);
  if (rc != PHX_OK) return rc;
  PHX_REQUIRE(num_envs >= 1, PHX_ERR_INVALID, "num_envs must be >= 1");
  PHX_REQUIRE(env_offset >= 0 && env_offset + num_envs <= 0xFFFFFFFFll, PHX_ERR_INVALID,
              "env_offset + num_envs must fit the 32-bit env coordinate of the RNG contract");
  int ndev = phx_device_count();
  PHX_REQUIRE(ndev > 0, PHX_ERR_NO_DEVICE,
              "no CUDA device visible: libphx has no CPU fallback by design");
  PHX_REQUIRE(device >= 0 && device < ndev, PHX_ERR_INVALID, "device index out of range");
  Family* fam = spec->family == PHX_FAMILY_USER ? phx::make_user_family(cubin_path, *spec) : make_family(spec);
  if (fam == nullptr) {
    set_error("no device program for family " + std::to_string(spec->family));
    return PHX_ERR_UNSUPPORTED;
  }
  rc = fam->base_init(*spec, num_envs, device, seed, env_offset);
  if (rc == PHX_OK) rc = fam->init(*spec);
  if (rc != PHX_OK) {
    delete fam;
    return rc;
  }
  phx_env* h = new (std::nothrow) phx_env{fam};
  if (h == nullptr) {
    delete fam;
    set_error("out of host memory allocating the env handle");
    return PHX_ERR_INVALID;
  }
  *out = h;
  return PHX_OK;
}

int32_t phx_create(const phx_spec* spec, int32_t num_envs, int32_t device, uint64_t seed,
                   int64_t env_offset, phx_env** out) {
  PHX_REQUIRE(spec == nullptr || spec->family != PHX_FAMILY_USER, PHX_ERR_INVALID,
              "PHX_FAMILY_USER needs phx_create_user (a cubin path)");
  return create_common(spec, nullptr, num_envs, device, seed, env_offset, out);
}

int32_t phx_create_user(const phx_spec* spec, const char* cubin_path, int32_t num_envs,
                        int32_t device, uint64_t seed, int64_t env_offset, phx_env** out) {
  PHX_REQUIRE(spec != nullptr && spec->family == PHX_FAMILY_USER, PHX_ERR_INVALID,
              "phx_create_user: spec->family must be PHX_FAMILY_USER");
  PHX_REQUIRE(cubin_path != nullptr, PHX_ERR_INVALID, "cubin_path is NULL");
  return create_common(spec, cubin_path, num_envs, device, seed, env_offset, out);
}

void phx_destroy(phx_env* env) {
  if (!env) return;
  cudaSetDevice(env->fam->device);
  cudaDeviceSynchronize();
  delete env->fam;
  delete env;
}

int32_t phx_num_envs(const phx_env* env) { return env ? env->fam->E : 0; }

const char* phx_exec_name(const phx_env* env) { return env ? env->fam->exec_name() : ""; }

int32_t phx_reset(phx_env* env, const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                  void* stream) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  PHX_CUDA(cudaSetDevice(env->fam->device));
  return env->fam->reset(env_mask, obs, obs_mask, (cudaStream_t)stream);
}

int32_t phx_rollout(phx_env* env, int32_t T, const float* actions, const uint8_t* action_mask,
                    float* obs, uint8_t* obs_mask, float* reward, uint8_t* reward_mask,
                    uint8_t* term, uint8_t* trunc, uint8_t* all_done, void* stream) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  PHX_REQUIRE(T >= 1, PHX_ERR_INVALID, "T must be >= 1");
  PHX_REQUIRE(actions != nullptr || env->fam->spec.n_strategic == 0 || action_mask != nullptr,
              PHX_ERR_INVALID, "actions is NULL");
  PHX_CUDA(cudaSetDevice(env->fam->device));
  phx::StepIO io{actions, action_mask, obs, obs_mask, reward, reward_mask, term, trunc, all_done};
  return env->fam->rollout(T, io, (cudaStream_t)stream);
}

int32_t phx_step(phx_env* env, const float* actions, const uint8_t* action_mask, float* obs,
                 uint8_t* obs_mask, float* reward, uint8_t* reward_mask, uint8_t* term,
                 uint8_t* trunc, uint8_t* all_done, void* stream) {
  return phx_rollout(env, 1, actions, action_mask, obs, obs_mask, reward, reward_mask, term,
                     trunc, all_done, stream);
}

void* phx_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void phx_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int32_t phx_rollout_host(phx_env* env, int32_t T, const float* actions,
                         const uint8_t* action_mask, float* obs, uint8_t* obs_mask,
                         float* reward, uint8_t* reward_mask, uint8_t* term, uint8_t* trunc,
                         uint8_t* all_done) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  PHX_REQUIRE(T >= 1, PHX_ERR_INVALID, "T must be >= 1");
  Family* f = env->fam;
  PHX_CUDA(cudaSetDevice(f->device));
  phx::StepIO h{actions, action_mask, obs, obs_mask, reward, reward_mask, term, trunc, all_done};
  return f->rollout_host(T, h);
}

static int32_t field_copy(phx_env* env, int32_t field, int32_t index, void* host, uint64_t bytes,
                          bool set) {
  PHX_REQUIRE(env != nullptr && host != nullptr, PHX_ERR_INVALID, "NULL argument");
  Family* f = env->fam;
  PHX_CUDA(cudaSetDevice(f->device));
  PHX_CUDA(cudaDeviceSynchronize());
  if (field == PHX_FIELD_STEP || field == PHX_FIELD_EPISODE || field == PHX_FIELD_STAGE) {
    PHX_REQUIRE(bytes == sizeof(int32_t) * (uint64_t)f->E, PHX_ERR_INVALID,
                "buffer must hold int32 [E]");
    std::vector<int4> h((size_t)f->E);
    PHX_CUDA(cudaMemcpy(h.data(), f->d_hdr, sizeof(int4) * (size_t)f->E, cudaMemcpyDeviceToHost));
    int32_t* io = (int32_t*)host;
    for (int e = 0; e < f->E; ++e) {
      int32_t* w = field == PHX_FIELD_STEP ? &h[e].x : field == PHX_FIELD_EPISODE ? &h[e].y : &h[e].z;
      if (set) *w = io[e]; else io[e] = *w;
    }
    if (set)
      PHX_CUDA(cudaMemcpy(f->d_hdr, h.data(), sizeof(int4) * (size_t)f->E, cudaMemcpyHostToDevice));
    return PHX_OK;
  }
  void* d = nullptr;
  size_t n = 0;
  int32_t rc = f->field_ptr(field, index, &d, &n);
  if (rc != PHX_OK) return rc;
  PHX_REQUIRE(bytes == n, PHX_ERR_INVALID,
              "field size mismatch: column has " + std::to_string(n) + " bytes, buffer " +
                  std::to_string(bytes));
  if (set) PHX_CUDA(cudaMemcpy(d, host, n, cudaMemcpyHostToDevice));
  else PHX_CUDA(cudaMemcpy(host, d, n, cudaMemcpyDeviceToHost));
  return PHX_OK;
}

int32_t phx_get_field(phx_env* env, int32_t field, int32_t index, void* host_out,
                      uint64_t out_bytes) {
  return field_copy(env, field, index, host_out, out_bytes, false);
}

int32_t phx_set_field(phx_env* env, int32_t field, int32_t index, const void* host_in,
                      uint64_t in_bytes) {
  return field_copy(env, field, index, const_cast<void*>(host_in), in_bytes, true);
}

int32_t phx_reduce_field(phx_env* env, int32_t field, int32_t index, int32_t width, int32_t col,
                         int64_t* host_sum, int32_t* host_min, int32_t* host_max) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  Family* f = env->fam;
  PHX_CUDA(cudaSetDevice(f->device));
  void* d = nullptr;
  size_t n = 0;
  int32_t rc = f->field_ptr(field, index, &d, &n);
  if (rc != PHX_OK) return rc;
  PHX_REQUIRE(width >= 1 && col >= 0 && col < width &&
                  n == sizeof(int32_t) * (size_t)f->E * (size_t)width,
              PHX_ERR_INVALID, "width / col do not match the column's layout");
  struct Acc { unsigned long long sum; int32_t mn, mx; } h{0ull, INT32_MAX, INT32_MIN};
  Acc* dacc = nullptr;
  PHX_CUDA(cudaMalloc(&dacc, sizeof(Acc)));
  // steps issued on non-blocking streams are not ordered before the legacy default stream
  PHX_CUDA(cudaDeviceSynchronize());
  PHX_CUDA(cudaMemcpy(dacc, &h, sizeof(Acc), cudaMemcpyHostToDevice));
  const int blocks = (f->E + 255) / 256 < 1184 ? (f->E + 255) / 256 : 1184;  // 8 blocks per SM
  phx_field_reduce_kernel<<<blocks, 256>>>((const int32_t*)d, f->E, width, col, &dacc->sum,
                                            &dacc->mn, &dacc->mx);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpy(&h, dacc, sizeof(Acc), cudaMemcpyDeviceToHost);
  cudaFree(dacc);
  PHX_CUDA(err);
  if (host_sum) *host_sum = (int64_t)h.sum;
  if (host_min) *host_min = h.mn;
  if (host_max) *host_max = h.mx;
  return PHX_OK;
}

int32_t phx_get_trace_step(phx_env* env, int32_t step, int32_t env_begin, int32_t env_end,
                           int32_t* host_counts, int32_t* host_msgs) {
  PHX_REQUIRE(env != nullptr && host_counts != nullptr && host_msgs != nullptr, PHX_ERR_INVALID,
              "NULL argument");
  Family* f = env->fam;
  PHX_REQUIRE(f->tracking(), PHX_ERR_INVALID, "handle was created without PHX_FLAG_TRACK_MESSAGES");
  PHX_REQUIRE(0 <= env_begin && env_begin <= env_end && env_end <= f->E, PHX_ERR_INVALID,
              "env range out of bounds");
  PHX_REQUIRE(step >= 0 && step < f->trace_T, PHX_ERR_INVALID,
              "step outside the last tracked launch (it recorded " + std::to_string(f->trace_T) +
                  " step(s))");
  PHX_CUDA(cudaSetDevice(f->device));
  PHX_CUDA(cudaDeviceSynchronize());
  const size_t n = (size_t)(env_end - env_begin);
  const size_t first = (size_t)step * f->E + env_begin;
  PHX_CUDA(cudaMemcpy(host_counts, f->d_trace_cnt + first, sizeof(int32_t) * n,
                      cudaMemcpyDeviceToHost));
  PHX_CUDA(cudaMemcpy(host_msgs, f->d_trace + first * f->spec.trace_capacity,
                      sizeof(int4) * n * f->spec.trace_capacity, cudaMemcpyDeviceToHost));
  return PHX_OK;
}

int32_t phx_get_trace(phx_env* env, int32_t env_begin, int32_t env_end, int32_t* host_counts,
                      int32_t* host_msgs) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  return phx_get_trace_step(env, env->fam->trace_T - 1, env_begin, env_end, host_counts, host_msgs);
}

int32_t phx_trace_steps(const phx_env* env) { return env ? env->fam->trace_T : 0; }

int32_t phx_jit_source(phx_env* env, char* buf, uint64_t buf_bytes, uint64_t* needed) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  std::string text;
  const int32_t rc = env->fam->jit_source(text);
  if (rc != PHX_OK) return rc;
  if (needed) *needed = (uint64_t)text.size() + 1;
  if (buf == nullptr || buf_bytes == 0) return PHX_OK;  // size query
  PHX_REQUIRE(buf_bytes >= text.size() + 1, PHX_ERR_INVALID, "buffer too small for the source text");
  std::memcpy(buf, text.c_str(), text.size() + 1);
  return PHX_OK;
}

int32_t phx_load_specialised(phx_env* env, const char* cubin_path) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  Family* f = env->fam;
  PHX_CUDA(cudaSetDevice(f->device));
  PHX_CUDA(cudaDeviceSynchronize());
  return f->load_specialised(cubin_path);
}

int32_t phx_poll_errors(phx_env* env, int32_t* n_bad, int32_t* first_env, int32_t* code,
                        int32_t clear) {
  PHX_REQUIRE(env != nullptr, PHX_ERR_INVALID, "env is NULL");
  Family* f = env->fam;
  PHX_CUDA(cudaSetDevice(f->device));
  PHX_CUDA(cudaDeviceSynchronize());
  uint32_t n = 0;
  PHX_CUDA(cudaMemcpy(&n, f->d_nfaults, sizeof(n), cudaMemcpyDeviceToHost));
  int32_t first = -1, c = 0;
  if (n != 0) {
    std::vector<uint32_t> h((size_t)f->E);
    PHX_CUDA(cudaMemcpy(h.data(), f->d_err, sizeof(uint32_t) * (size_t)f->E,
                        cudaMemcpyDeviceToHost));
    for (int e = 0; e < f->E; ++e)
      if (h[e] != 0) { first = e; c = (int32_t)h[e]; break; }
    if (clear) {
      PHX_CUDA(cudaMemset(f->d_err, 0, sizeof(uint32_t) * (size_t)f->E));
      PHX_CUDA(cudaMemset(f->d_nfaults, 0, sizeof(uint32_t)));
    }
  }
  if (n_bad) *n_bad = (int32_t)n;
  if (first_env) *first_env = first;
  if (code) *code = c;
  return PHX_OK;
}

int32_t phx_selftest_ratio(int32_t device, int32_t den, int32_t lo, int32_t count,
                           float* host_out) {
  return phx::selftest_ratio(device, den, lo, count, host_out);
}

int32_t phx_selftest_wire_expand(int32_t max_stock, int32_t cap, const uint32_t* wire, uint64_t n,
                                 int32_t threads, float* obs, float* reward, uint8_t* all_done) {
  PHX_REQUIRE(max_stock > 0 && cap > 0 && threads >= 1 && wire && obs && reward && all_done,
              PHX_ERR_INVALID, "bad arguments");
  phx::HostPool pool(threads);
  phx::sc_wire_expand(pool, phx::ScWireParams{max_stock, cap}, wire, (size_t)n, obs, reward,
                      all_done);
  return PHX_OK;
}

uint32_t phx_selftest_wire_pack(int32_t stock, int32_t sales, int32_t missed, int32_t truncated,
                                int32_t was_reset) {
  return phx::scw_pack(stock, sales, missed, truncated != 0, was_reset != 0);
}

}  // extern "C"
