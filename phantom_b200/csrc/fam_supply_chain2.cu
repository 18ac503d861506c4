// fam_supply_chain2.cu -- device program family PHX_FAMILY_SUPPLY_CHAIN2: the multi-shop supply
// chain with agent SUPERTYPES of the reference's tutorial (docs/user/tutorial2.rst:90-128 several
// shops, customers pick one at random; :236-345 the `excess_stock_weight` supertype sampled per
// episode by a UniformFloatSampler), as defined by oracle/workloads/supply_chain2.py.
//
// What is new relative to fam_supply_chain.cu: per-env, per-episode sampled TYPE parameters
// (phantom/supertype.py:14-30, phantom/utils/samplers.py:119-147, sampled by the env at reset,
// phantom/env.py:212-216, copied into agent.type by Agent.reset, phantom/agents.py:166-168) live
// in the agent's state columns and feed reward and observation in float64.
//
// Agent kinds: 0 ShopAgent (strategic), 1 FactoryAgent, 2 CustomerAgent.
// Payload types: 0 OrderRequest, 1 OrderResponse, 2 StockRequest, 3 StockResponse.
// State words (shop): 0 stock, 1 sales, 2 missed_sales, 3 delivered, 4/5 excess_stock_weight
//                     (float64 bits, low / high word).
// iparams: 0 max_order, 1 max_stock, 2 n_shops, 3..8 shop slots (the customers' shop_ids list),
//          9 n_customers.
// dparams: 0 MAX_EXCESS_STOCK_WEIGHT (obs scale).
// agent_iparam[slot] = {shop: factory slot | customer: ordinal, -, shop: sampler index or -1}
// agent_fparam[slot] = {sampler low (or the constant value), sampler high}
// RNG: stream 0 order size, stream 4 shop choice (idx = customer ordinal); stream 3 samplers
// (step 0, idx = position in the env's sampler list).
#include "phx_engine_host.cuh"
#ifndef PHX_JIT_TU
#include "phx_engine_wide_host.cuh"
#endif

namespace phx {
namespace {

enum { S2_SHOP = 0, S2_FACTORY = 1, S2_CUSTOMER = 2 };
enum { S2_ORDER_REQUEST = 0, S2_ORDER_RESPONSE = 1, S2_STOCK_REQUEST = 2, S2_STOCK_RESPONSE = 3 };
constexpr int S2_STREAM_ORDER = 0, S2_STREAM_SAMPLER = 3, S2_STREAM_CHOICE = 4;

struct Sc2Program {
  // run-time specialisation (phx_jit.cuh): where this program lives and what it is called
  static constexpr const char* JIT_SOURCE = "fam_supply_chain2.cu";
  static constexpr const char* JIT_NAME = "Sc2Program";
  static constexpr int PW = 1, NWORDS = 6, VW = 0, ACTCAP = 1, RESPCAP = 8, OBS_DIM = 4,
                       ACT_DIM = 1, Q1CAP = 8;
  static constexpr int RECVCAP = 8;  // max messages one agent receives in a round
  static constexpr bool BATCHED = false, HAS_PRE = true, HAS_POST = false;
  static constexpr bool WIDE_OK = true;  // every callback is a template over the context type

  static int q1_cap(const phx_spec& s) { return s.n_agents; }

  static int32_t validate(const phx_spec& s) {
    PHX_REQUIRE(s.env_kind == PHX_ENV_BASE, PHX_ERR_UNSUPPORTED,
                "supply-chain-2 family runs under PhantomEnv (PHX_ENV_BASE) only");
    PHX_REQUIRE(s.n_agents <= 8, PHX_ERR_UNSUPPORTED, "supply-chain-2 family: up to 8 agents");
    PHX_REQUIRE(s.obs_dim == 4 && s.act_dim == 1 && s.n_payload_types == 4, PHX_ERR_INVALID,
                "supply-chain-2 family: obs_dim 4, act_dim 1, 4 payload types");
    PHX_REQUIRE(s.iparams[0] >= 1 && s.iparams[0] <= 255 && s.iparams[1] >= 1 &&
                    s.iparams[1] <= (1 << 20) && s.iparams[2] >= 1 && s.iparams[2] <= 6,
                PHX_ERR_INVALID, "supply-chain-2 family: parameters out of range");
    PHX_REQUIRE(s.fparams[0] > 0.0, PHX_ERR_INVALID, "MAX_EXCESS_STOCK_WEIGHT must be > 0");
    return PHX_OK;
  }

  __device__ static double weight(const int* st) {
    return __hiloint2double(st[5], st[4]);
  }

  template <class C, class E>
  __device__ static void act(const C& c, int* st, bool has_action, const float* action, E& out) {
    const auto& sp = *c.spec;
    if (c.kind == S2_SHOP) {
      if (!has_action) return;
      const float a0 = action[0];
      if (!(fabsf(a0) <= 1048576.0f)) {
        out.fault = PHX_FAULT_INVALID_ACTION;
        return;
      }
      out.send(sp.agent_iparam[c.slot][0], S2_STOCK_REQUEST, min(__float2int_rn(a0), sp.iparams[1] - st[0]));
    } else if (c.kind == S2_CUSTOMER) {  // tutorial2.rst:90-97
      const uint32_t ord = (uint32_t)sp.agent_iparam[c.slot][0];
      const int size = rng_randint(c.rand24_hi(S2_STREAM_ORDER, ord), (uint32_t)sp.iparams[0]);
      const int pick = rng_randint(c.rand24_hi(S2_STREAM_CHOICE, ord), (uint32_t)sp.iparams[2]);
      out.send(sp.iparams[3 + pick], S2_ORDER_REQUEST, size);
    }
  }
  template <class C>
  __device__ static void view(const C&, const int*, int*) {}
  template <class C>
  __device__ static void pre(const C& c, int* st) {
    if (c.kind == S2_SHOP) st[1] = st[2] = 0;
  }
  template <class C>
  __device__ static void post(const C&, int*) {}

  template <class C, class E>
  __device__ static bool handle(const C& c, int* st, const Msg& m, E& out) {
    if (c.kind == S2_SHOP) {
      if (m.type == S2_STOCK_RESPONSE) {
        st[3] = m.p[0];
        st[0] = min(st[0] + m.p[0], c.spec->iparams[1]);
        return true;
      }
      if (m.type == S2_ORDER_REQUEST) {
        const int sold = min(m.p[0], st[0]);
        st[2] += m.p[0] - sold;
        st[0] -= sold;
        st[1] += sold;
        out.send(m.sender, S2_ORDER_RESPONSE, sold);
        return true;
      }
      return false;
    }
    if (c.kind == S2_FACTORY) {
      if (m.type != S2_STOCK_REQUEST) return false;
      out.send(m.sender, S2_STOCK_RESPONSE, m.p[0]);
      return true;
    }
    return m.type == S2_ORDER_RESPONSE;
  }

  template <class C>
  __device__ static bool encode(const C& c, int* st, float* obs) {
    const auto& sp = *c.spec;
    const float cap = (float)(sp.iparams[9] * sp.iparams[0]);  // n_customers * max_order
    obs[0] = __fdiv_rn((float)st[0], (float)sp.iparams[1]);
    obs[1] = __fdiv_rn((float)st[1], cap);
    obs[2] = __fdiv_rn((float)st[2], cap);
    obs[3] = (float)__ddiv_rn(weight(st), sp.dparams[0]);  // float32(type.w / MAX_W), float64 division
    return true;
  }
  template <class C>
  __device__ static float reward(const C&, int* st) {
    // sales - type.excess_stock_weight * stock in float64, two roundings (no contraction)
    return (float)__dsub_rn((double)st[1], __dmul_rn(weight(st), (double)st[0]));
  }
  template <class C>
  __device__ static bool terminated(const C&, const int*) { return false; }
  template <class C>
  __device__ static bool truncated(const C&, const int*) { return false; }

  template <class C>
  __device__ static void reset_agent(const C& c, int* st) {
    if (c.kind != S2_SHOP) return;
    // Agent.reset(): self.type = self.supertype.sample() -- the env-managed sampler's value
    const auto& sp = *c.spec;
    const int idx = sp.agent_iparam[c.slot][2];
    double w = sp.agent_fparam[c.slot][0];
    if (idx >= 0) {  // UniformFloatSampler: low + (high - low) * u, u = d24 / 2^24, all float64
      const double u = (double)(c.rand24_hi(S2_STREAM_SAMPLER, (uint32_t)idx) >> 8) * (1.0 / 16777216.0);
      w = __dadd_rn(sp.agent_fparam[c.slot][0],
                    __dmul_rn(__dsub_rn(sp.agent_fparam[c.slot][1], sp.agent_fparam[c.slot][0]), u));
    }
    st[4] = __double2loint(w);
    st[5] = __double2hiint(w);
    st[0] = 0;  // ShopAgent.reset: stock = 0 (sales / missed survive)
  }
};

}  // namespace

#ifndef PHX_JIT_TU  // a specialised translation unit only needs the program above
Family* make_supply_chain2_family(const phx_spec& s) { return make_engine_family<Sc2Program>(s); }
#endif

}  // namespace phx
