// phx_engine.cuh -- the generic message-queue engine: one fused kernel that runs the whole
// reference step loop for a batch of envs, for ANY agent order / topology of a family.
//
// What it replaces (all under /root/reference/phantom/):
//   env.py:239-303        PhantomEnv.step            (clock, contexts, acting phase, outputs)
//   env.py:320-348        _handle_acting_agents / _make_ctxs
//   env.py:170-183        pre/post_message_resolution, resolve_network
//   network.py:233-254    Network.send               (edge check, payload whitelist, push)
//   network.py:208-222    Network.context_for        (start-of-step neighbour views)
//   resolvers.py:128-163  BatchResolver.resolve      (rounds, first-arrival receiver order,
//                                                     delivery-time edge filter, round limit)
//   agents.py:96-155      Agent.handle_batch / handle_message dispatch
//   fsm.py:253-380        FiniteStateMachineEnv.step (stage-gated acting set, caches)
//   stackelberg.py:111-196 StackelbergEnv.step       (leader / follower turns, reward cache)
//   env.py:185-237, fsm.py:195-251, stackelberg.py:53-109   reset
//
// Mapping: a TILE of G lanes (G = 8, 16 or 32, lane == agent slot) owns one env; a block of
// 128 threads holds 128/G envs.  Agent state lives in registers for the whole launch (T
// steps); the per-env message queue, the neighbour views and the output rows are staged in
// shared memory; nothing but actions, outputs and (once per launch) state touches HBM.
//
// The queue reproduces the reference's ordering rules without any sorting of
// messages: it is SEGMENTED by producing agent.  In the acting phase agent i appends to
// segment i, and segments are read in agent order, which is the global push order
// (env.py:324).  In a round every receiver lane scans the current queue in that order,
// handles the messages addressed to it (push order within its batch, agents.py:110-120) and
// appends its responses to ITS segment of the next queue; the next queue's segment order is
// the receivers sorted by the position of their first message (dict insertion order of
// `messages[receiver]`, resolvers.py:126,142).  Handlers of different receivers only touch
// their own agent's state, so all receivers of a round run in parallel.
#pragma once
#include "phx_common.cuh"
#include "phx_rng.cuh"

namespace phx {

constexpr int ENGINE_BLOCK = 128;
constexpr int ENGINE_MAX_AGENTS = 32;

// Generic part of phx_spec in kernel-parameter (constant bank) form.
struct EngineSpec {
  int32_t E, n_agents, n_strategic, num_steps, round_limit, env_kind;
  uint32_t flags;
  int32_t obs_dim, act_dim;
  int8_t kind[ENGINE_MAX_AGENTS];
  int8_t sidx[ENGINE_MAX_AGENTS];      // strategic index or -1
  uint32_t adj[ENGINE_MAX_AGENTS];     // bit r of adj[s]: edge s -> r
  uint32_t sender_ok[PHX_MAX_TYPES];   // bit s: slot s may send this payload type
  uint32_t receiver_ok[PHX_MAX_TYPES];
  uint32_t strategic_mask;             // over slots
  int32_t n_stages, initial_stage;
  uint32_t stage_acting[PHX_MAX_STAGES];
  uint32_t stage_rewarded[PHX_MAX_STAGES];
  uint8_t stage_rewarded_none[PHX_MAX_STAGES];
  int8_t stage_next[PHX_MAX_STAGES];
  uint32_t leaders, followers;         // over slots
  uint64_t seed;
  uint32_t env_offset;
  int32_t iparams[PHX_MAX_PARAMS];
  float fparams[PHX_MAX_PARAMS];
  int32_t agent_iparam[ENGINE_MAX_AGENTS][4];
  float agent_fparam[ENGINE_MAX_AGENTS][2];
  int32_t codec_op[ENGINE_MAX_AGENTS][PHX_MAX_CODEC_OPS];  // opcode | length << 8
  float codec_val[ENGINE_MAX_AGENTS][PHX_MAX_CODEC_OPS];
};

struct Msg {
  int sender, type;
  int p[2];
};

// What a device program sees of "its" agent and env: phantom.Context (context.py:11-40) with
// the env view (views.py:27-34; fsm.py:66-73 adds the stage) and the neighbour views.
struct Ctx {
  const EngineSpec* spec;
  int slot, kind;
  int step;         // EnvView.current_step (already incremented, env.py:252)
  int stage;        // FSMEnvView.stage index
  uint32_t env_id;  // global env index (RNG contract)
  uint32_t episode;
  const int* views; // start-of-step snapshot, [slot][VW] (this env's tile)
  int view_stride;
  __device__ __forceinline__ float proportion_time_elapsed() const {
    // EnvView.proportion_time_elapsed = current_step / num_steps in float64 (env.py:166-168)
    return (float)((double)step / (double)spec->num_steps);
  }
  __device__ __forceinline__ const int* view_of(int other_slot) const {
    return views + other_slot * view_stride;
  }
  __device__ __forceinline__ bool has_neighbour(int other_slot) const {
    return (spec->adj[slot] >> other_slot) & 1u;
  }
  __device__ __forceinline__ uint32_t rand24_hi(uint32_t stream, uint32_t idx) const {
    return rng_d24_hi(spec->seed, env_id, episode, (uint32_t)step, stream, idx);
  }
};

// Segmented queue of one env in shared memory.  Entry (seg, k): head word + PW payload
// words.  Layout is word-major so that the lanes of a tile writing their own segments hit
// different banks.
template <int G, int SEGCAP, int PW>
struct TileQueue {
  uint32_t head[SEGCAP][G];   // sender | recv << 8 | type << 16
  int32_t pay[PW][SEGCAP][G];
  uint8_t cnt[G];             // entries per segment
  uint8_t order[G];           // segment visiting order
  int32_t nseg;
};

// Emission cursor of one lane: appends to the lane's own segment after the reference's send
// checks (network.py:246-254).
template <int G, int SEGCAP, int PW>
struct Emit {
  TileQueue<G, SEGCAP, PW>* q;
  const EngineSpec* spec;
  int slot;
  int n;
  uint32_t fault;
  __device__ __forceinline__ void send(int recv, int type, int p0, int p1 = 0) {
    if (fault) return;
    const bool edge = (spec->adj[slot] >> recv) & 1u;
    if (!(spec->flags & PHX_FLAG_IGNORE_CONNECTION_ERRORS) && !edge) {
      fault = PHX_FAULT_NO_EDGE;
      return;
    }
    if (!(spec->flags & PHX_FLAG_NO_PAYLOAD_CHECKS)) {
      if (!((spec->sender_ok[type] >> slot) & 1u) || !((spec->receiver_ok[type] >> recv) & 1u)) {
        fault = PHX_FAULT_BAD_PAYLOAD_TYPE;
        return;
      }
    }
    if (n >= SEGCAP) {
      fault = PHX_FAULT_QUEUE_OVERFLOW;
      return;
    }
    q->head[n][slot] = (uint32_t)slot | ((uint32_t)recv << 8) | ((uint32_t)type << 16);
    q->pay[0][n][slot] = p0;
    if (PW > 1) q->pay[PW > 1 ? 1 : 0][n][slot] = p1;
    ++n;
  }
};

template <class P>
struct EngineArgs {
  EngineSpec spec;
  int32_t T;
  int4* hdr;            // [E] step, episode, stage, -
  uint32_t* term;       // [E] bitmask over strategic index
  uint32_t* trunc;      // [E]
  int32_t* state;       // [NWORDS][E][G]
  float* reward_cache;  // [E][G]   FSM / Stackelberg `_rewards` (by slot)
  uint32_t* reward_none;// [E]      bit slot: cached reward is None
  float* obs_cache;     // [E][G][O] FSM `_observations` (by slot)
  uint32_t* obs_cached; // [E]      bit slot: an obs is cached
  StepIO io;
  FaultSink faults;
  TraceSink trace;
};

__device__ __forceinline__ uint32_t tile_mask(int G) {
  return G >= 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
}

// Shared-memory footprint of one env tile.
template <class P, int G>
struct TileSmem {
  TileQueue<G, P::SEGCAP, P::PW> q[2];
  int32_t views[G][P::VW > 0 ? P::VW : 1];
  int32_t first_idx[G];
};

// The fused step kernel.  P is the device program of a family (see fam_*.cu for the
// interface: view / act / pre / handle / post / encode / reward / terminated / truncated /
// reset_agent).  encode and reward receive the mutable agent state: the reference's
// callbacks may have side effects (the KAT agents count their calls).
template <class P, int G, bool TRACK>
__global__ void __launch_bounds__(ENGINE_BLOCK) engine_step_kernel(const EngineArgs<P> a) {
  constexpr int TPB = ENGINE_BLOCK / G;  // env tiles per block
  constexpr int INF = 0x7FFFFFFF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<P, G>* tiles = reinterpret_cast<TileSmem<P, G>*>(smem_raw);

  const EngineSpec& sp = a.spec;
  const int tb = threadIdx.x / G;
  const int slot = threadIdx.x % G;
  const int env = blockIdx.x * TPB + tb;
  const bool env_live = env < sp.E;
  const bool is_agent = env_live && slot < sp.n_agents;
  const uint32_t tmask = tile_mask(G);
  TileSmem<P, G>& ts = tiles[tb];
  const int e = env_live ? env : sp.E - 1;

  const int kind = slot < sp.n_agents ? sp.kind[slot] : -1;
  const int sidx = slot < sp.n_agents ? sp.sidx[slot] : -1;
  const bool strategic = sidx >= 0;
  const int S = sp.n_strategic, O = sp.obs_dim, A = sp.act_dim;

  // ---- load env header, done sets, agent state, caches (once per launch)
  int4 h = a.hdr[e];
  uint32_t term = a.term[e], trunc = a.trunc[e];
  int st[P::NWORDS > 0 ? P::NWORDS : 1];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) st[w] = a.state[((size_t)w * sp.E + e) * G + slot];
  const bool cached_env = sp.env_kind != PHX_ENV_BASE;
  float rcache = 0.f;
  uint32_t rnone = 0, ocached = 0;
  if (cached_env) {
    rcache = a.reward_cache[(size_t)e * G + slot];
    rnone = a.reward_none[e];
    if (sp.env_kind == PHX_ENV_FSM) ocached = a.obs_cached[e];
  }
  uint32_t fault_key = 0xFFFFFFFFu;  // (phase << 16 | slot << 8 | code), smallest wins

  Ctx ctx;
  ctx.spec = &sp;
  ctx.slot = slot;
  ctx.kind = kind;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.views = &ts.views[0][0];
  ctx.view_stride = P::VW > 0 ? P::VW : 1;

  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * sp.E + e;  // [T,E] row of this env
    h.x += 1;                                 // env.py:252
    ctx.step = h.x;
    ctx.episode = (uint32_t)h.y;
    ctx.stage = h.z;
    const uint32_t done_bits = term | trunc;
    const bool was_done = strategic && ((done_bits >> sidx) & 1u);
    const bool has_ctx = is_agent && !was_done;  // env.py:344-348: no context for done agents

    // ---- start-of-step snapshot of every agent's public state (network.py:208-222)
    if (P::VW > 0) {
      if (is_agent) P::view(ctx, st, &ts.views[slot][0]);
      __syncwarp(tmask);
    }

    // ---- acting phase (env.py:320-336; fsm.py:276-277; stackelberg.py:133-140)
    uint32_t acting = 0xFFFFFFFFu, observing = sp.strategic_mask, rewarded = sp.strategic_mask;
    int next_stage = h.z;
    if (sp.env_kind == PHX_ENV_FSM) {
      acting = sp.stage_acting[h.z];
      next_stage = sp.stage_next[h.z];
      if (!sp.stage_rewarded_none[h.z]) {  // fsm.py:315-320
        rewarded = sp.stage_rewarded[h.z];
        observing = sp.stage_acting[next_stage];
      }
    } else if (sp.env_kind == PHX_ENV_STACKELBERG) {
      const bool leaders_turn = (h.x & 1) == 1;
      acting = leaders_turn ? sp.leaders : sp.followers;
      observing = leaders_turn ? sp.followers : sp.leaders;
      rewarded = acting;
    }
    int cur = 0;
    Emit<G, P::SEGCAP, P::PW> out{&ts.q[cur], &sp, slot, 0, 0u};
    if (has_ctx && ((acting >> slot) & 1u)) {
      bool has_action = false;
      const float* act = nullptr;
      if (strategic) {
        has_action = a.io.action_mask ? a.io.action_mask[row * S + sidx] != 0 : true;
        act = a.io.actions + (row * S + sidx) * A;
      }
      P::act(ctx, st, has_action, act, out);
    }
    ts.q[cur].cnt[slot] = (uint8_t)out.n;
    ts.q[cur].order[slot] = (uint8_t)slot;
    if (slot == 0) ts.q[cur].nseg = sp.n_agents;
    if (out.fault) fault_key = min(fault_key, (0u << 16) | ((uint32_t)slot << 8) | out.fault);
    __syncwarp(tmask);

    int traced = 0;
    if (TRACK && env_live && slot == 0) {  // pushes of the acting phase, in global push order
      for (int si = 0; si < sp.n_agents; ++si)
        for (int k = 0; k < ts.q[cur].cnt[si]; ++k) {
          const uint32_t hd = ts.q[cur].head[k][si];
          if (traced < a.trace.cap)
            a.trace.rows[(size_t)e * a.trace.cap + traced] =
                make_int4((int)hd, ts.q[cur].pay[0][k][si],
                          P::PW > 1 ? ts.q[cur].pay[P::PW > 1 ? 1 : 0][k][si] : 0, 0);
          ++traced;
        }
    }

    // ---- pre_message_resolution for every live context, in agent order (env.py:170-173);
    // hooks only touch their own agent, so order across agents is immaterial
    if (has_ctx) P::pre(ctx, st);

    // ---- BatchResolver.resolve (resolvers.py:128-163)
    for (int round = 0;; ++round) {
      TileQueue<G, P::SEGCAP, P::PW>& qc = ts.q[cur];
      TileQueue<G, P::SEGCAP, P::PW>& qn = ts.q[cur ^ 1];
      int total = 0;
      const int nseg = qc.nseg;
      for (int si = 0; si < nseg; ++si) total += qc.cnt[qc.order[si]];
      if (total == 0) break;
      if (sp.round_limit >= 0 && round >= sp.round_limit) {  // resolvers.py:160-163
        fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | (0xFFu << 8) | PHX_FAULT_ROUND_LIMIT);
        break;
      }
      Emit<G, P::SEGCAP, P::PW> resp{&qn, &sp, slot, 0, 0u};
      int first = INF, pos = 0;
      bool bad_type = false;
      if constexpr (P::BATCHED) {
        if (has_ctx) P::batch_begin(ctx, st);
      }
      for (int si = 0; si < nseg; ++si) {
        const int seg = qc.order[si];
        const int c = qc.cnt[seg];
        for (int k = 0; k < c; ++k, ++pos) {
          const uint32_t hd = qc.head[k][seg];
          if ((int)((hd >> 8) & 0xFFu) != slot) continue;
          if (first == INF) first = pos;  // first-arrival position of this receiver
          if (!has_ctx) continue;         // done agent: mail dropped (resolvers.py:143-144)
          const int sender = (int)(hd & 0xFFu);
          // delivery-time edge filter (resolvers.py:146-148)
          if (!((sp.adj[sender] >> slot) & 1u)) continue;
          Msg m;
          m.sender = sender;
          m.type = (int)((hd >> 16) & 0xFFu);
          m.p[0] = qc.pay[0][k][seg];
          m.p[1] = P::PW > 1 ? qc.pay[P::PW > 1 ? 1 : 0][k][seg] : 0;
          if (!P::handle(ctx, st, m, resp)) bad_type = true;  // agents.py:140-143
        }
      }
      if constexpr (P::BATCHED) {
        if (has_ctx && first != INF) P::batch_end(ctx, st, resp);
      }
      if (bad_type)
        fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) |
                                       PHX_FAULT_UNKNOWN_MSG_TYPE);
      if (resp.fault)
        fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) | resp.fault);
      qn.cnt[slot] = (uint8_t)resp.n;
      ts.first_idx[slot] = first;
      __syncwarp(tmask);
      // next queue's segment order = receivers by first-arrival position
      int rank = 0, nrecv = 0;
      for (int j = 0; j < G; ++j) {
        const int fj = ts.first_idx[j];
        nrecv += fj != INF;
        rank += fj < first;
      }
      if (first != INF) qn.order[rank] = (uint8_t)slot;
      if (slot == 0) qn.nseg = nrecv;
      __syncwarp(tmask);
      if (TRACK && env_live && slot == 0) {
        for (int si = 0; si < nrecv; ++si) {
          const int seg = qn.order[si];
          for (int k = 0; k < qn.cnt[seg]; ++k) {
            if (traced < a.trace.cap)
              a.trace.rows[(size_t)e * a.trace.cap + traced] =
                  make_int4((int)qn.head[k][seg], qn.pay[0][k][seg],
                            P::PW > 1 ? qn.pay[P::PW > 1 ? 1 : 0][k][seg] : 0, round + 1);
            ++traced;
          }
        }
      }
      cur ^= 1;
    }
    if (TRACK && env_live && slot == 0) a.trace.cnt[e] = traced;

    // ---- post_message_resolution (env.py:175-178)
    if (has_ctx) P::post(ctx, st);

    // ---- outputs for strategic agents (env.py:273-303; fsm.py:322-378;
    // stackelberg.py:149-194)
    bool obs_now = false, rew_now = false;
    float obs_val[P::OBS_DIM] = {};
    float rew_val = 0.f;
    bool t_flag = false, u_flag = false;
    if (strategic && has_ctx) {
      if ((observing >> slot) & 1u) obs_now = P::encode(ctx, st, obs_val);  // None -> false
      if (sp.env_kind == PHX_ENV_BASE) {
        if (obs_now) {  // env.py:281-284: reward only travels with an observation
          rew_val = P::reward(ctx, st);
          rew_now = true;
        }
      } else if ((rewarded >> slot) & 1u) {
        rew_val = P::reward(ctx, st);
        rew_now = true;
        rcache = rew_val;
      }
      t_flag = P::terminated(ctx, st);
      u_flag = P::truncated(ctx, st);
    }
    const uint32_t lane_bit = strategic ? (1u << sidx) : 0u;
    // tile-wide: newly done agents, cache bookkeeping
    const uint32_t t_ballot = __ballot_sync(tmask, t_flag);
    const uint32_t u_ballot = __ballot_sync(tmask, u_flag);
    const uint32_t o_ballot = __ballot_sync(tmask, obs_now);
    const uint32_t r_ballot = __ballot_sync(tmask, rew_now);
    const int tile_shift = G >= 32 ? 0 : ((threadIdx.x & 31) / G * G);
    // ballots are over lanes (= slots); convert the done ballots to strategic-index masks
    uint32_t t_new = 0, u_new = 0;
    {
      const uint32_t tb_ = (t_ballot >> tile_shift), ub_ = (u_ballot >> tile_shift);
      for (int j = 0; j < sp.n_agents; ++j) {
        const int sj = sp.sidx[j];
        if (sj >= 0) {
          t_new |= ((tb_ >> j) & 1u) << sj;
          u_new |= ((ub_ >> j) & 1u) << sj;
        }
      }
    }
    term |= t_new;
    trunc |= u_new;
    const uint32_t obs_slots = (o_ballot >> tile_shift), rew_slots = (r_ballot >> tile_shift);
    if (cached_env) {
      rnone &= ~rew_slots;  // _rewards.update(rewards)
      if (sp.env_kind == PHX_ENV_FSM) ocached |= obs_slots;
    }
    const bool all_term = __popc(term) == S;                           // env.py:308-310
    const bool all_trunc = (h.x == sp.num_steps) || __popc(trunc) == S;  // env.py:312-318
    const bool terminal = all_term || all_trunc;
    if (sp.env_kind == PHX_ENV_FSM) h.z = next_stage;  // fsm.py:355

    if (strategic && env_live) {
      const size_t orow = row * S + sidx;
      // observation + masks, per step-loop kind
      uint8_t om = 0, rm = 0;
      float r_out = 0.f;
      if (sp.env_kind == PHX_ENV_BASE) {
        om = obs_now;
        rm = rew_now;
        r_out = rew_val;
      } else if (sp.env_kind == PHX_ENV_FSM) {
        if (obs_now && a.obs_cache)
          for (int j = 0; j < O; ++j) a.obs_cache[((size_t)e * G + slot) * O + j] = obs_val[j < P::OBS_DIM ? j : 0];
        if (terminal) {  // fsm.py:360-375: flush the caches
          om = (ocached >> slot) & 1u;
          if (om && !obs_now)
            for (int j = 0; j < O; ++j) obs_val[j < P::OBS_DIM ? j : 0] = a.obs_cache[((size_t)e * G + slot) * O + j];
          rm = ((rnone >> slot) & 1u) ? 2 : 1;
          r_out = rcache;
        } else {  // fsm.py:378: last computed reward of every agent observing now
          om = obs_now;
          if (obs_now) {
            rm = ((rnone >> slot) & 1u) ? 2 : 1;
            r_out = rcache;
          }
        }
      } else {  // Stackelberg
        om = obs_now;
        if (terminal) {  // stackelberg.py:180-187: the whole reward cache
          rm = ((rnone >> slot) & 1u) ? 2 : 1;
          r_out = rcache;
        } else if (obs_now && !((rnone >> slot) & 1u)) {  // stackelberg.py:190-194
          rm = 1;
          r_out = rcache;
        }
      }
      if (a.io.obs && om)
        for (int j = 0; j < O; ++j) a.io.obs[orow * O + j] = obs_val[j < P::OBS_DIM ? j : 0];
      if (a.io.obs_mask) a.io.obs_mask[orow] = om;
      if (a.io.reward) a.io.reward[orow] = rm == 1 ? r_out : 0.f;
      if (a.io.reward_mask) a.io.reward_mask[orow] = rm;
      if (a.io.term) a.io.term[orow] = was_done ? 255 : (uint8_t)t_flag;
      if (a.io.trunc) a.io.trunc[orow] = was_done ? 255 : (uint8_t)u_flag;
    }
    if (env_live && slot == 0 && a.io.all_done)
      reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(all_term, all_trunc);
    (void)lane_bit;

    // ---- PHX_FLAG_AUTO_RESET: the step that ends the episode also resets the env
    if ((sp.flags & PHX_FLAG_AUTO_RESET) && terminal) {
      h.x = 0;
      h.y += 1;
      h.z = sp.initial_stage;
      term = trunc = 0;
      if (is_agent) P::reset_agent(ctx, st);
      rnone = cached_env ? sp.strategic_mask : 0u;
      ctx.step = 0;
      ctx.episode = (uint32_t)h.y;
      ctx.stage = h.z;
      if (P::VW > 0) {
        if (is_agent) P::view(ctx, st, &ts.views[slot][0]);
        __syncwarp(tmask);
      }
      uint32_t first_obs = sp.strategic_mask;
      if (sp.env_kind == PHX_ENV_FSM) first_obs &= sp.stage_acting[sp.initial_stage];
      if (sp.env_kind == PHX_ENV_STACKELBERG) first_obs &= sp.leaders;
      if (strategic && env_live) {
        const size_t orow = row * S + sidx;
        bool got = false;
        if ((first_obs >> slot) & 1u) got = P::encode(ctx, st, obs_val);
        if (a.io.obs && got)
          for (int j = 0; j < O; ++j) a.io.obs[orow * O + j] = obs_val[j < P::OBS_DIM ? j : 0];
        if (a.io.obs_mask) a.io.obs_mask[orow] = got;
      }
    }
  }

  // ---- write back
  if (env_live) {
    if (slot == 0) {
      a.hdr[e] = h;
      a.term[e] = term;
      a.trunc[e] = trunc;
      if (cached_env) {
        a.reward_none[e] = rnone;
        if (sp.env_kind == PHX_ENV_FSM) a.obs_cached[e] = ocached;
      }
    }
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * G + slot] = st[w];
    if (cached_env) a.reward_cache[(size_t)e * G + slot] = rcache;
  }
  // first fault of the env in event order (phase, then agent order)
  uint32_t fk = fault_key;
  for (int off = G / 2; off > 0; off >>= 1) fk = min(fk, __shfl_xor_sync(tmask, fk, off, G));
  if (env_live && slot == 0 && fk != 0xFFFFFFFFu) raise_fault(a.faults, e, fk & 0xFFu);
}

// PhantomEnv.reset / FiniteStateMachineEnv.reset / StackelbergEnv.reset for masked envs.
template <class P, int G>
__global__ void __launch_bounds__(ENGINE_BLOCK)
engine_reset_kernel(const EngineArgs<P> a, const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                    bool agents_only) {
  constexpr int TPB = ENGINE_BLOCK / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<P, G>* tiles = reinterpret_cast<TileSmem<P, G>*>(smem_raw);
  const EngineSpec& sp = a.spec;
  const int tb = threadIdx.x / G, slot = threadIdx.x % G;
  const int env = blockIdx.x * TPB + tb;
  const uint32_t tmask = tile_mask(G);
  const bool env_live = env < sp.E && (env_mask == nullptr || env_mask[env < sp.E ? env : 0] != 0);
  const int e = env < sp.E ? env : sp.E - 1;
  const bool is_agent = slot < sp.n_agents;
  TileSmem<P, G>& ts = tiles[tb];
  int4 h = a.hdr[e];
  int st[P::NWORDS > 0 ? P::NWORDS : 1];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) st[w] = a.state[((size_t)w * sp.E + e) * G + slot];
  h.x = 0;
  h.y += 1;
  h.z = sp.env_kind == PHX_ENV_FSM ? sp.initial_stage : 0;
  Ctx ctx;
  ctx.spec = &sp;
  ctx.slot = slot;
  ctx.kind = is_agent ? sp.kind[slot] : -1;
  ctx.step = 0;
  ctx.stage = h.z;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.episode = (uint32_t)h.y;
  ctx.views = &ts.views[0][0];
  ctx.view_stride = P::VW > 0 ? P::VW : 1;
  if (is_agent) P::reset_agent(ctx, st);  // Network.reset -> agent.reset() (network.py:179-184)
  if (agents_only) {  // PhantomEnv.__init__ ends with agent.reset() only (env.py:122-124)
    if (env_live) {
#pragma unroll
      for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * G + slot] = st[w];
    }
    return;
  }
  if (P::VW > 0) {
    if (is_agent) P::view(ctx, st, &ts.views[slot][0]);
    __syncwarp(tmask);
  }
  const int sidx = is_agent ? sp.sidx[slot] : -1;
  // who observes at reset: all strategic (env.py:227), the initial stage's acting agents
  // (fsm.py:238-243) or the leaders (stackelberg.py:97-101)
  uint32_t first_obs = sp.strategic_mask;
  if (sp.env_kind == PHX_ENV_FSM) first_obs &= sp.stage_acting[sp.initial_stage];
  if (sp.env_kind == PHX_ENV_STACKELBERG) first_obs &= sp.leaders;
  if (env_live) {
    if (sidx >= 0) {
      float obs_val[P::OBS_DIM] = {};
      bool got = false;
      if ((first_obs >> slot) & 1u) got = P::encode(ctx, st, obs_val);
      const size_t orow = (size_t)e * sp.n_strategic + sidx;
      if (obs && got)
        for (int j = 0; j < sp.obs_dim; ++j) obs[orow * sp.obs_dim + j] = obs_val[j < P::OBS_DIM ? j : 0];
      if (obs_mask) obs_mask[orow] = got;
    }
    if (slot == 0) {
      a.hdr[e] = h;
      a.term[e] = 0;
      a.trunc[e] = 0;
      if (sp.env_kind != PHX_ENV_BASE) a.reward_none[e] = sp.strategic_mask;  // _rewards = None
    }
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * G + slot] = st[w];
  }
}

}  // namespace phx
