// fam_market.cu -- device program family PHX_FAMILY_MARKET (BASELINE config C3): the
// three-stage FiniteStateMachineEnv market of oracle/workloads/market.py, modelled on the
// reference's examples/environments/simple_market/ (2-stage FSM, Price/Order messages) and
// the digital-ads auction.  MAKER -> TAKER -> CLEARING -> MAKER, handler-less stages.
//
// Agent kinds: 0 MakerAgent (strategic), 1 TakerAgent (strategic), 2 ClearingAgent.
// Payload types: 0 Quote(price), 1 Order(maker ordinal, price), 2 Fill(units, notional).
// State words (per slot):
//   maker     0 inventory, 1 cash, 2 last_price, 3 last_notional
//   taker     0 value, 1 best_price, 2 best_maker, 3 holdings, 4 last_surplus
//   clearing  0..6 per maker: units | notional << 8;  7 accepted-taker bitmask
// Views (network.py:208-222): a maker's public state is its inventory (VW = 1).
// iparams: 0 n_makers, 1 n_takers, 2 MAKER_INVENTORY, 3 MAKER_CAPACITY, 4 stage index of
// MAKER, 5 stage index of CLEARING, 6 slot of the clearing agent.
// agent_iparam[slot][0] = maker / taker ordinal.
#include "phx_engine_host.cuh"

namespace phx {
namespace {

enum { MKT_MAKER = 0, MKT_TAKER = 1, MKT_CLEARING = 2 };
enum { MKT_QUOTE = 0, MKT_ORDER = 1, MKT_FILL = 2 };
constexpr int MKT_NO_QUOTE = 1 << 20;
constexpr int MKT_STREAM_VALUE = 1;

struct MarketProgram {
  // run-time specialisation: where this program lives and what it is called
  static constexpr const char* JIT_SOURCE = "fam_market.cu";
  static constexpr const char* JIT_NAME = "MarketProgram";
  // a maker quotes up to 24 takers and the clearing agent settles with 31 agents in the acting
  // phase; no handler of this market ever answers a message
  static constexpr int PW = 2, NWORDS = 8, VW = 1, ACTCAP = 32, RESPCAP = 1, OBS_DIM = 3,
                       ACT_DIM = 1, Q1CAP = 0;  // 32 agents: tile engine only
  static constexpr int RECVCAP = 32;  // max messages one agent receives in a round
  // compact acting queue: a maker quotes its taker neighbours, a taker sends at most one order,
  // the clearing agent settles with every neighbour -- 7 x 25 + 24 + 31 = 230 entries at most
  static constexpr int ACTTOTAL = 256;
  __host__ __device__ static int act_cap(int kind, int out_degree) {
    return kind == 1 /* MKT_TAKER */ ? 1 : out_degree;
  }
  static constexpr bool BATCHED = false, HAS_PRE = true, HAS_POST = true;

  static int q1_cap(const phx_spec&) { return 0; }
  static int32_t validate(const phx_spec& s) {
    PHX_REQUIRE(s.env_kind == PHX_ENV_FSM, PHX_ERR_UNSUPPORTED,
                "market family runs under FiniteStateMachineEnv only");
    PHX_REQUIRE(s.iparams[0] >= 1 && s.iparams[0] <= 7 && s.iparams[1] >= 1 && s.iparams[1] <= 24,
                PHX_ERR_UNSUPPORTED, "market family: up to 7 makers and 24 takers");
    PHX_REQUIRE(s.obs_dim == 3 && s.act_dim == 1 && s.n_payload_types == 3, PHX_ERR_INVALID,
                "market family: obs_dim 3, act_dim 1, 3 payload types");
    int clearing = 0;
    for (int i = 0; i < s.n_agents; ++i) clearing += s.agent_kind[i] == MKT_CLEARING;
    PHX_REQUIRE(clearing == 1, PHX_ERR_UNSUPPORTED, "exactly one ClearingAgent per env");
    PHX_REQUIRE(s.iparams[3] >= 0 && s.iparams[3] <= 255, PHX_ERR_INVALID, "MAKER_CAPACITY range");
    return PHX_OK;
  }

  // The clearing agent's per-maker words are addressed by a run-time maker ordinal; a plain
  // st[k] would push the whole state array of EVERY lane to local memory (the C3 profile showed
  // the spill stores among the top stalls), a compare chain keeps it in registers.
  __device__ static int maker_word(const int* st, int k) {
    int v = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j == k) v = st[j];
    return v;
  }

  __device__ static void view(const Ctx& c, const int* st, int* v) {
    v[0] = c.kind == MKT_MAKER ? st[0] : 0;  // MakerView(inventory)
  }

  template <class E>
  __device__ static void act(const Ctx& c, int* st, bool has_action, const float* action, E& out) {
    const EngineSpec& sp = *c.spec;
    if (c.kind == MKT_MAKER) {
      if (!has_action) return;
      const float a0 = action[0];
      if (!(fabsf(a0) <= 1048576.0f)) {
        out.fault = PHX_FAULT_INVALID_ACTION;
        return;
      }
      st[2] = max(0, min(100, __float2int_rn(__fmul_rn(a0, 100.0f))));
      for (uint32_t m = c.neighbours_of_kind(MKT_TAKER); m; m &= m - 1)  // taker_ids, agent order
        out.send(__ffs(m) - 1, MKT_QUOTE, st[2]);
    } else if (c.kind == MKT_TAKER) {
      if (has_action && __float2int_rn(action[0]) == 1 && st[2] >= 0)
        out.send(sp.iparams[6], MKT_ORDER, st[2], st[1]);
    } else if (c.stage == sp.iparams[5]) {  // ClearingAgent.generate_messages, CLEARING stage
      for (uint32_t m = c.neighbours_of_kind(MKT_MAKER); m; m &= m - 1) {
        const int r = __ffs(m) - 1;
        const int w = maker_word(st, c.iparam0_of(r));
        out.send(r, MKT_FILL, w & 0xFF, w >> 8);
      }
      for (uint32_t m = c.neighbours_of_kind(MKT_TAKER); m; m &= m - 1) {
        const int r = __ffs(m) - 1;
        out.send(r, MKT_FILL, (st[7] >> c.iparam0_of(r)) & 1, 0);
      }
    }
  }

  __device__ static void pre(const Ctx& c, int* st) {
    if (c.kind == MKT_TAKER && c.stage == c.spec->iparams[4]) {  // a new cycle
      st[1] = MKT_NO_QUOTE;
      st[2] = -1;
      st[4] = 0;
    }
  }
  __device__ static void post(const Ctx& c, int* st) {
    if (c.kind == MKT_CLEARING && c.stage == c.spec->iparams[5])
      for (int w = 0; w < 8; ++w) st[w] = 0;
  }

  template <class E>
  __device__ static bool handle(const Ctx& c, int* st, const Msg& m, E&) {
    const EngineSpec& sp = *c.spec;
    if (c.kind == MKT_MAKER) {
      if (m.type != MKT_FILL) return false;
      st[0] -= m.p[0];
      st[1] += m.p[1];
      st[3] = m.p[1];
      return true;
    }
    if (c.kind == MKT_TAKER) {
      if (m.type == MKT_QUOTE) {
        if (c.view_of(m.sender)[0] <= 0) return true;  // maker had no inventory at step start
        if (m.p[0] < st[1]) {
          st[1] = m.p[0];
          st[2] = c.iparam0_of(m.sender);
        }
        return true;
      }
      if (m.type == MKT_FILL) {
        if (m.p[0] > 0) {
          st[3] += 1;
          st[4] = st[0] - st[1];
        }
        return true;
      }
      return false;
    }
    if (m.type != MKT_ORDER) return false;
    const int mk = m.p[0];
    if ((maker_word(st, mk) & 0xFF) < sp.iparams[3]) {  // first come, first served
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (j == mk) st[j] += 1 + (m.p[1] << 8);
      st[7] |= 1 << c.iparam0_of(m.sender);
    }
    return true;
  }

  __device__ static bool encode(const Ctx& c, int* st, float* obs) {
    if (c.kind == MKT_MAKER) {
      const float inv = (float)c.spec->iparams[2];  // (a constant in a specialised unit)
      obs[0] = ratio_rn(st[0], inv, 1.0f / inv);
      obs[1] = ratio_rn(st[2], 100.0f, 1.0f / 100.0f);
      obs[2] = 0.f;
    } else {
      obs[0] = ratio_rn(st[0], 100.0f, 1.0f / 100.0f);
      obs[1] = ratio_rn(st[2] < 0 ? 100 : st[1], 100.0f, 1.0f / 100.0f);
      obs[2] = ratio_rn(st[3], 33.0f, 1.0f / 33.0f);
    }
    return true;
  }
  __device__ static float reward(const Ctx& c, int* st) {
    return ratio_rn(c.kind == MKT_MAKER ? st[3] : st[4], 100.0f, 1.0f / 100.0f);
  }
  __device__ static bool terminated(const Ctx& c, const int* st) {
    return c.kind == MKT_MAKER && st[0] <= 0;
  }
  __device__ static bool truncated(const Ctx&, const int*) { return false; }
  __device__ static void reset_agent(const Ctx& c, int* st) {
    for (int w = 0; w < NWORDS; ++w) st[w] = 0;
    if (c.kind == MKT_MAKER) {
      st[0] = c.spec->iparams[2];
    } else if (c.kind == MKT_TAKER) {
      st[0] = rng_randint(c.rand24_hi(MKT_STREAM_VALUE, (uint32_t)c.spec->agent_iparam[c.slot][0]), 101u);
      st[1] = MKT_NO_QUOTE;
      st[2] = -1;
    }
  }
};

}  // namespace

#ifndef PHX_JIT_TU
Family* make_market_family(const phx_spec&) { return new EngineFamily<MarketProgram>(); }
#endif

}  // namespace phx
