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

#ifndef PHX_JIT_TU
  // Host: is this env class one the collective resolve below was written for?  Every potential
  // send must pass Network.send's checks on the static graph (so that no NO_EDGE /
  // BAD_PAYLOAD_TYPE fault can occur at run time), the stages must be handler-less, and no
  // data-dependent routing option may be on.
  static bool collective_ok(const phx_spec& s) {
    if (s.env_kind != PHX_ENV_FSM || s.n_agents > 32 || s.round_limit == 0) return false;
    if (s.flags & (PHX_FLAG_STOCHASTIC_NETWORK | PHX_FLAG_SHUFFLE_BATCHES)) return false;
    for (int k = 0; k < s.n_stages; ++k)
      if (s.stages[k].handler != 0) return false;
    const int cl = s.iparams[6];
    for (int a = 0; a < s.n_agents; ++a) {
      for (int r = 0; r < s.n_agents; ++r) {
        if (!mask_bit(s.adjacency[a], r)) continue;
        if (mask_bit(s.adjacency[a], r) != mask_bit(s.adjacency[r], a)) return false;
        if (s.agent_kind[a] == MKT_MAKER && s.agent_kind[r] == MKT_TAKER &&
            plan_send_check(s, a, r, MKT_QUOTE))
          return false;
        if (s.agent_kind[a] == MKT_CLEARING && s.agent_kind[r] != MKT_CLEARING &&
            plan_send_check(s, a, r, MKT_FILL))
          return false;
      }
      if (s.agent_kind[a] == MKT_TAKER && plan_send_check(s, a, cl, MKT_ORDER)) return false;
    }
    return true;
  }
#endif

  // COLLECTIVE RESOLVE of one step (see phx_engine.cuh HasCollective): acting phase, pre hook and
  // the (single) resolver round of the stage, for the whole tile, with the receivers pulling.
  //   Quotes   a taker walks the quoting makers in slot order (= push order of its batch) and
  //            reads each price with a shuffle: 7 shuffles instead of 7 x 24 queue entries.
  //   Orders   the clearing agent admits orders first come first served per maker, up to
  //            MAKER_CAPACITY (market.py: order dependent): the takers ordering maker k are a
  //            ballot, a taker's place in that queue is a popcount of the lower lanes, and the
  //            clearing agent's totals are a popcount and a REDUX.ADD of the admitted prices.
  //   Fills    every maker / taker pulls its Fill from the clearing agent's registers.
  // Handlers of different receivers only touch their own agent, and nobody answers a message in
  // this market, so the order of the three groups is free -- except that a Fill's payload is
  // what the clearing agent held in the acting phase, hence fills first.
  static constexpr bool HAS_COLLECTIVE = true;
  template <int G>
  __device__ static void step_collective(const Ctx& c, int* st, bool has_ctx, bool acts,
                                         bool has_action, const float* action, uint32_t tmask,
                                         uint32_t& fault) {
    const EngineSpec& sp = *c.spec;
    const int shift = G >= 32 ? 0 : ((threadIdx.x & 31) / G * G);  // the tile's first lane
    const uint32_t gmask = G >= 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    const int cl = sp.iparams[6];
    const int ord = c.slot < sp.n_agents ? c.iparam0_of(c.slot) : 0;  // maker / taker ordinal
    // ---- acting phase (env.py:320-336): who sends, with the payloads as of now
    bool quote = false, order = false, fills = false;
    if (acts) {
      if (c.kind == MKT_MAKER) {
        if (has_action) {
          const float a0 = action[0];
          if (!(fabsf(a0) <= 1048576.0f)) {
            fault = PHX_FAULT_INVALID_ACTION;
          } else {
            st[2] = max(0, min(100, __float2int_rn(__fmul_rn(a0, 100.0f))));
            quote = true;
          }
        }
      } else if (c.kind == MKT_TAKER) {
        order = has_action && __float2int_rn(action[0]) == 1 && st[2] >= 0 &&
                ((c.out_mask >> cl) & 1u);  // (no edge: only reachable with ignore_connection_errors)
      } else if (c.stage == sp.iparams[5]) {
        fills = true;
      }
    }
    const int q_price = st[2];
    const int o_maker = st[2], o_price = st[1];
    // ---- pre_message_resolution (env.py:170-173)
    if (has_ctx) pre(c, st);
    // ---- Fills: pulled from the clearing agent's registers
    if (__shfl_sync(tmask, (int)fills, cl + shift)) {  // tile-uniform
      int w_mine = 0;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int wj = __shfl_sync(tmask, st[j], cl + shift);
        if (j == ord) w_mine = wj;
      }
      const int accepted = __shfl_sync(tmask, st[7], cl + shift);
      if (has_ctx && ((c.in_mask >> cl) & 1u)) {
        if (c.kind == MKT_MAKER) {
          st[0] -= w_mine & 0xFF;
          st[1] += w_mine >> 8;
          st[3] = w_mine >> 8;
        } else if (c.kind == MKT_TAKER && ((accepted >> ord) & 1)) {
          st[3] += 1;
          st[4] = st[0] - st[1];
        }
      }
    }
    // ---- Quotes: the quoting makers in slot order
    for (uint32_t qm = (__ballot_sync(tmask, quote) >> shift) & gmask; qm; qm &= qm - 1) {
      const int m = __ffs(qm) - 1;
      const int pm = __shfl_sync(tmask, q_price, m + shift);
      if (has_ctx && c.kind == MKT_TAKER && ((c.in_mask >> m) & 1u) && c.view_of(m)[0] > 0 &&
          pm < st[1]) {
        st[1] = pm;
        st[2] = c.iparam0_of(m);
      }
    }
    // ---- Orders: first come (lowest taker slot), first served, per maker
    if (__ballot_sync(tmask, order) != 0u) {  // tile-uniform
      const uint32_t below = (1u << c.slot) - 1u;
      const bool cl_live = __shfl_sync(tmask, (int)has_ctx, cl + shift) != 0;
#pragma unroll
      for (int mk = 0; mk < 7; ++mk) {
        const bool mine = order && o_maker == mk;
        const uint32_t queue = (__ballot_sync(tmask, mine) >> shift) & gmask;
        if (queue == 0u) continue;  // tile-uniform
        const int cur = __shfl_sync(tmask, st[mk], cl + shift);
        const int room = sp.iparams[3] - (cur & 0xFF);
        const bool admitted = mine && cl_live && __popc(queue & below) < room;
        const int n_adm = __popc((__ballot_sync(tmask, admitted) >> shift) & gmask);
        const int sum_p = __reduce_add_sync(tmask, admitted ? o_price : 0);
        const uint32_t bits = __reduce_or_sync(tmask, admitted ? (1u << ord) : 0u);
        if (c.slot == cl && n_adm > 0) {
          st[mk] += n_adm + (sum_p << 8);
          st[7] |= (int)bits;
        }
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
