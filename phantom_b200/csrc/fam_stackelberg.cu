// fam_stackelberg.cu -- device program family PHX_FAMILY_STACKELBERG (BASELINE config C4): the
// leader-follower pricing game of oracle/workloads/stackelberg.py under StackelbergEnv
// (/root/reference/phantom/stackelberg.py:111-196: leaders act on odd steps, followers on
// even steps, rewards cached per acting group).
//
// Agent kinds: 0 LeaderAgent, 1 FollowerAgent (both strategic).
// Payload types: 0 Price(ticks), 1 Demand(qty), 2 Ack(filled).
// State words: leader   0 price, 1 remaining, 2 revenue_round, 3 demand_round
//              follower 0 value, 1 seen_price, 2 last_filled, 3 utility_round
// iparams: 0 CAPACITY.  agent_iparam[slot] = {follower ordinal (RNG idx), leader slot}.
#include "phx_engine_host.cuh"

namespace phx {
namespace {

enum { SK_LEADER = 0, SK_FOLLOWER = 1 };
enum { SK_PRICE = 0, SK_DEMAND = 1, SK_ACK = 2 };
constexpr int SK_STREAM_VALUE = 2;

struct StackelbergProgram {
  // run-time specialisation (phx_jit.cuh): where this program lives and what it is called
  static constexpr const char* JIT_SOURCE = "fam_stackelberg.cu";
  static constexpr const char* JIT_NAME = "StackelbergProgram";
  static constexpr int PW = 1, NWORDS = 4, VW = 0, ACTCAP = 8, RESPCAP = 8, OBS_DIM = 2,
                       ACT_DIM = 1, Q1CAP = 8;
  static constexpr int RECVCAP = 8;  // max messages one agent receives in a round
  static constexpr bool BATCHED = false, HAS_PRE = true, HAS_POST = false;

  static int q1_cap(const phx_spec& s) { return s.n_agents - 1; }  // one message per follower
  static int32_t validate(const phx_spec& s) {
    PHX_REQUIRE(s.env_kind == PHX_ENV_STACKELBERG, PHX_ERR_UNSUPPORTED,
                "stackelberg family runs under StackelbergEnv only");
    PHX_REQUIRE(s.n_agents <= 8, PHX_ERR_UNSUPPORTED, "stackelberg family: up to 7 followers");
    PHX_REQUIRE(s.obs_dim == 2 && s.act_dim == 1 && s.n_payload_types == 3, PHX_ERR_INVALID,
                "stackelberg family: obs_dim 2, act_dim 1, 3 payload types");
    PHX_REQUIRE(s.iparams[0] >= 1 && s.iparams[0] <= 100000, PHX_ERR_INVALID, "CAPACITY range");
    return PHX_OK;
  }

  __device__ static void view(const Ctx&, const int*, int*) {}

#ifndef PHX_JIT_TU
  // Static send signature (phx_engine_host.cuh build_static_plan): what act() / handle() below
  // may send, in emission order.
  static void act_sends(const phx_spec& s, int slot, int, std::vector<SendSig>& out) {
    if (s.agent_kind[slot] == SK_LEADER) {  // Price to every follower neighbour, slot order
      for (int r = 0; r < s.n_agents; ++r)
        if (s.agent_kind[r] == SK_FOLLOWER && mask_bit(s.adjacency[slot], r))
          out.push_back(SendSig{r, SK_PRICE});
    } else {
      out.push_back(SendSig{s.agent_iparam[slot][1], SK_DEMAND});
    }
  }
  static void handle_sends(const phx_spec& s, int slot, int type, int sender,
                           std::vector<SendSig>& out) {
    if (s.agent_kind[slot] == SK_LEADER && type == SK_DEMAND) out.push_back(SendSig{sender, SK_ACK});
  }
#endif

  template <class E>
  __device__ static void act(const Ctx& c, int* st, bool has_action, const float* action, E& out) {
    const EngineSpec& sp = *c.spec;
    if (!has_action) return;
    const float a0 = action[0];
    if (!(fabsf(a0) <= 1048576.0f)) {
      out.fault = PHX_FAULT_INVALID_ACTION;
      return;
    }
    if (c.kind == SK_LEADER) {
      st[0] = max(0, min(100, __float2int_rn(__fmul_rn(a0, 100.0f))));
#ifdef PHX_JIT_TU  // the neighbour mask is a constant: a countable loop unrolls to straight-line sends
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if ((c.neighbours_of_kind(SK_FOLLOWER) >> r) & 1u) out.send(r, SK_PRICE, st[0]);
#else
      for (uint32_t m = c.neighbours_of_kind(SK_FOLLOWER); m; m &= m - 1)
        out.send(__ffs(m) - 1, SK_PRICE, st[0]);
#endif
    } else {
      const int qty = max(0, min(10, __float2int_rn(__fmul_rn(a0, 10.0f))));
      out.send(sp.agent_iparam[c.slot][1], SK_DEMAND, qty);
    }
  }

  __device__ static void pre(const Ctx& c, int* st) {
    if ((c.step & 1) != 0) return;  // only on the followers' turn
    if (c.kind == SK_LEADER) {
      st[1] = c.spec->iparams[0];
      st[2] = 0;
      st[3] = 0;
    } else {
      st[2] = 0;
      st[3] = 0;
    }
  }
  __device__ static void post(const Ctx&, int*) {}

  template <class E>
  __device__ static bool handle(const Ctx& c, int* st, const Msg& m, E& out) {
    if (c.kind == SK_LEADER) {
      if (m.type != SK_DEMAND) return false;
      const int filled = min(m.p[0], st[1]);
      st[1] -= filled;
      st[2] += filled * st[0];
      st[3] += m.p[0];
      out.send(m.sender, SK_ACK, filled);
      return true;
    }
    if (m.type == SK_PRICE) {
      st[1] = m.p[0];
      return true;
    }
    if (m.type == SK_ACK) {
      st[2] = m.p[0];
      st[3] = m.p[0] * (st[0] - st[1]);
      return true;
    }
    return false;
  }

  __device__ static bool encode(const Ctx& c, int* st, float* obs) {
    if (c.kind == SK_LEADER) {
      const float cap = (float)c.spec->iparams[0];  // (a constant in a specialised unit)
      obs[0] = ratio_rn(st[3], 30.0f, 1.0f / 30.0f);
      obs[1] = ratio_rn(st[1], cap, 1.0f / cap);
    } else {
      obs[0] = ratio_rn(st[1], 100.0f, 1.0f / 100.0f);
      obs[1] = ratio_rn(st[2], 10.0f, 1.0f / 10.0f);
    }
    return true;
  }
  __device__ static float reward(const Ctx& c, int* st) {
    return ratio_rn(c.kind == SK_LEADER ? st[2] : st[3], 100.0f, 1.0f / 100.0f);
  }
  __device__ static bool terminated(const Ctx&, const int*) { return false; }
  __device__ static bool truncated(const Ctx&, const int*) { return false; }
  __device__ static void reset_agent(const Ctx& c, int* st) {
    st[0] = st[1] = st[2] = st[3] = 0;
    if (c.kind == SK_LEADER) {
      st[1] = c.spec->iparams[0];
    } else {
      st[0] = 50 + rng_randint(c.rand24_hi(SK_STREAM_VALUE, (uint32_t)c.spec->agent_iparam[c.slot][0]), 51u);
    }
  }
};

}  // namespace

#ifndef PHX_JIT_TU  // a specialised translation unit only needs the program above
Family* make_stackelberg_family(const phx_spec&) { return new EngineFamily<StackelbergProgram>(); }
#endif

}  // namespace phx
