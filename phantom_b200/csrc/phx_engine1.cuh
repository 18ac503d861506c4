// phx_engine1.cuh -- the generic message-queue engine with ONE THREAD PER ENV (tile size 1).
//
// Same semantics, same device-program interface and same HBM layout as the tile engine
// (phx_engine.cuh, which see for the reference lines replaced), for env classes with at most
// 8 agents.  There the lane-per-agent mapping wastes the machine: a 4-agent env keeps 4 of 32
// lanes busy and agents of different kinds serialise inside the warp (measured on C4: ~250
// warp-instructions per env-step, 17 % of all samples waiting at tile reductions).  With a
// thread per env every lane advances its own env; the reference's sequential semantics are
// restated literally -- agents act in order, a round visits receivers in first-arrival order
// and each handles its batch in push order -- so no ordering machinery is needed at all.
//
// Per-thread working set in shared memory (word-major, thread-minor: conflict free):
//   agent state   [NWORDS][8][threads]      views [VW][8][threads]
//   two queues    [Q1CAP][1 + PW][threads]  (current round / next round)
#pragma once
#include "phx_engine.cuh"

namespace phx {

constexpr int ENGINE1_BLOCK = 128;
constexpr int ENGINE1_SLOTS = 8;  // == the G of the HBM state layout [NWORDS][E][8]

template <class P>
struct Engine1Smem {
  int32_t state[P::NWORDS > 0 ? P::NWORDS : 1][ENGINE1_SLOTS][ENGINE1_BLOCK];
  int32_t views[P::VW > 0 ? P::VW : 1][ENGINE1_SLOTS][ENGINE1_BLOCK];
  uint32_t qhead[2][P::Q1CAP][ENGINE1_BLOCK];
  int32_t qpay[2][P::Q1CAP][P::PW][ENGINE1_BLOCK];
  int8_t kind_tab[ENGINE_MAX_AGENTS];
  int32_t ip0_tab[ENGINE_MAX_AGENTS];
};

// Emission cursor of one env: appends to the queue in program order (== global push order).
template <class P>
struct Emit1 {
  Engine1Smem<P>* sm;
  const EngineSpec* spec;
  int which, tid;
  int slot;
  uint32_t out_mask;
  int n;
  uint32_t fault;
  __device__ __forceinline__ void send(int recv, int type, int p0, int p1 = 0) {
    if (fault) return;
    if (!(spec->flags & PHX_FLAG_IGNORE_CONNECTION_ERRORS) && !((out_mask >> recv) & 1u)) {
      fault = PHX_FAULT_NO_EDGE;
      return;
    }
    if (!(spec->flags & PHX_FLAG_NO_PAYLOAD_CHECKS)) {
      if (!((spec->sender_ok[type] >> slot) & 1u) || !((spec->receiver_ok[type] >> recv) & 1u)) {
        fault = PHX_FAULT_BAD_PAYLOAD_TYPE;
        return;
      }
    }
    if (n >= P::Q1CAP) {
      fault = PHX_FAULT_QUEUE_OVERFLOW;
      return;
    }
    sm->qhead[which][n][tid] = (uint32_t)slot | ((uint32_t)recv << 8) | ((uint32_t)type << 16);
    sm->qpay[which][n][0][tid] = p0;
    if (P::PW > 1) sm->qpay[which][n][P::PW > 1 ? 1 : 0][tid] = p1;
    ++n;
  }
};

template <class P, bool TRACK>
__global__ void __launch_bounds__(ENGINE1_BLOCK) engine1_step_kernel(const EngineArgs<P> a) {
  static_assert(P::VW <= 1, "thread-per-env engine: views of at most one word per agent");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Engine1Smem<P>& sm = *reinterpret_cast<Engine1Smem<P>*>(smem_raw);
  const EngineSpec& sp = a.spec;
  const int tid = threadIdx.x;
  const int env = blockIdx.x * ENGINE1_BLOCK + tid;
  const bool env_live = env < sp.E;
  const int e = env_live ? env : sp.E - 1;
  const int n = sp.n_agents, S = sp.n_strategic, O = sp.obs_dim;

  if (tid < ENGINE_MAX_AGENTS) {
    sm.kind_tab[tid] = sp.kind[tid];
    sm.ip0_tab[tid] = sp.agent_iparam[tid][0];
  }
  // ---- load env header, done sets, agent state (once per launch)
  int4 h = a.hdr[e];
  uint32_t term = a.term[e], trunc = a.trunc[e];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) {
    const int4* src = reinterpret_cast<const int4*>(a.state + ((size_t)w * sp.E + e) * ENGINE1_SLOTS);
    const int4 lo = src[0], hi = src[1];
    sm.state[w][0][tid] = lo.x; sm.state[w][1][tid] = lo.y; sm.state[w][2][tid] = lo.z; sm.state[w][3][tid] = lo.w;
    sm.state[w][4][tid] = hi.x; sm.state[w][5][tid] = hi.y; sm.state[w][6][tid] = hi.z; sm.state[w][7][tid] = hi.w;
  }
  const bool cached_env = sp.env_kind != PHX_ENV_BASE;
  uint32_t rnone = 0, ocached = 0;
  if (cached_env) {
    rnone = a.reward_none[e];
    if (sp.env_kind == PHX_ENV_FSM) ocached = a.obs_cached[e];
  }
  __syncthreads();
  uint32_t fault = 0;  // first fault in event order: events ARE sequential here

  Ctx ctx;
  ctx.spec = &sp;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.views = &sm.views[0][0][tid];
  ctx.view_stride = ENGINE1_BLOCK;  // view word 0 of slot s at views[0][s][tid]
  ctx.kind_tab = sm.kind_tab;
  ctx.ip0_tab = sm.ip0_tab;

  auto bind = [&](int slot) {
    ctx.slot = slot;
    ctx.kind = sp.kind[slot];
    ctx.out_mask = sp.adj[slot];
  };
  auto in_mask_of = [&](int slot) {
    uint32_t in = 0;
    for (int s = 0; s < n; ++s) in |= ((sp.adj[s] >> slot) & 1u) << s;
    return in;
  };
  auto load_state = [&](int slot, int* st) {
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) st[w] = sm.state[w][slot][tid];
  };
  auto store_state = [&](int slot, const int* st) {
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) sm.state[w][slot][tid] = st[w];
  };

  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * sp.E + e;
    h.x += 1;  // env.py:252
    ctx.step = h.x;
    ctx.episode = (uint32_t)h.y;
    ctx.stage = h.z;
    const uint32_t done = term | trunc;  // agents without a context this step (env.py:344-348)
    int st[P::NWORDS > 0 ? P::NWORDS : 1];

    // ---- start-of-step snapshot of every agent's public state (network.py:208-222)
    if (P::VW > 0) {
      for (int s = 0; s < n; ++s) {
        bind(s);
        load_state(s, st);
        int v[P::VW > 0 ? P::VW : 1];
        P::view(ctx, st, v);
#pragma unroll
        for (int w = 0; w < P::VW; ++w) sm.views[w][s][tid] = v[w];
      }
    }

    // ---- acting phase, agents in order (env.py:320-336; fsm.py:276-277; stackelberg.py:133-140)
    uint32_t acting = 0xFFFFFFFFu, observing = sp.strategic_mask, rewarded = sp.strategic_mask;
    int next_stage = h.z;
    if (sp.env_kind == PHX_ENV_FSM) {
      acting = sp.stage_acting[h.z];
      next_stage = sp.stage_next[h.z];
      if (!sp.stage_rewarded_none[h.z]) {
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
    Emit1<P> out{&sm, &sp, cur, tid, 0, 0u, 0, 0u};
    for (int s = 0; s < n; ++s) {
      if (!((acting >> s) & 1u) || ((done >> s) & 1u)) continue;
      bind(s);
      out.slot = s;
      out.out_mask = ctx.out_mask;
      load_state(s, st);
      const int sidx = sp.sidx[s];
      bool has_action = false;
      float act[P::ACT_DIM];
#pragma unroll
      for (int j = 0; j < P::ACT_DIM; ++j) act[j] = 0.f;
      if (sidx >= 0) {
        has_action = a.io.action_mask ? a.io.action_mask[row * S + sidx] != 0 : true;
#pragma unroll
        for (int j = 0; j < P::ACT_DIM; ++j) act[j] = a.io.actions[(row * S + sidx) * P::ACT_DIM + j];
      }
      P::act(ctx, st, has_action, act, out);
      store_state(s, st);
    }
    if (out.fault && !fault) fault = out.fault;
    int n_cur = out.n;
    int traced = 0;
    if (TRACK && env_live) {
      for (int i = 0; i < n_cur; ++i, ++traced)
        if (traced < a.trace.cap)
          a.trace.rows[(size_t)e * a.trace.cap + traced] =
              make_int4((int)sm.qhead[cur][i][tid], sm.qpay[cur][i][0][tid],
                        P::PW > 1 ? sm.qpay[cur][i][P::PW > 1 ? 1 : 0][tid] : 0, 0);
    }

    // ---- pre_message_resolution (env.py:170-173)
    for (int s = 0; s < n; ++s) {
      if ((done >> s) & 1u) continue;
      bind(s);
      load_state(s, st);
      P::pre(ctx, st);
      store_state(s, st);
    }

    // ---- BatchResolver.resolve (resolvers.py:128-163), literally
    for (int round = 0; n_cur > 0; ++round) {
      if (sp.round_limit >= 0 && round >= sp.round_limit) {
        if (!fault) fault = PHX_FAULT_ROUND_LIMIT;
        break;
      }
      Emit1<P> resp{&sm, &sp, cur ^ 1, tid, 0, 0u, 0, 0u};
      uint32_t seen = 0;
      for (int i = 0; i < n_cur; ++i) {
        const int r = (int)((sm.qhead[cur][i][tid] >> 8) & 0xFFu);
        if ((seen >> r) & 1u) continue;
        seen |= 1u << r;  // receivers in first-arrival order (resolvers.py:126,142)
        if ((done >> r) & 1u) continue;  // no context: mail dropped (:143-144)
        bind(r);
        resp.slot = r;
        resp.out_mask = ctx.out_mask;
        ctx.in_mask = in_mask_of(r);
        load_state(r, st);
        if constexpr (P::BATCHED) P::batch_begin(ctx, st);
        for (int j = i; j < n_cur; ++j) {  // its batch, in push order
          const uint32_t hd = sm.qhead[cur][j][tid];
          if ((int)((hd >> 8) & 0xFFu) != r) continue;
          const int sender = (int)(hd & 0xFFu);
          if (!((ctx.in_mask >> sender) & 1u)) continue;  // delivery-time edge filter (:146-148)
          Msg m;
          m.sender = sender;
          m.type = (int)((hd >> 16) & 0xFFu);
          m.p[0] = sm.qpay[cur][j][0][tid];
          m.p[1] = P::PW > 1 ? sm.qpay[cur][j][P::PW > 1 ? 1 : 0][tid] : 0;
          if (!P::handle(ctx, st, m, resp) && !fault && !resp.fault) fault = PHX_FAULT_UNKNOWN_MSG_TYPE;
        }
        if constexpr (P::BATCHED) P::batch_end(ctx, st, resp);
        store_state(r, st);
        if (resp.fault && !fault) fault = resp.fault;
      }
      if (TRACK && env_live) {
        for (int i = 0; i < resp.n; ++i, ++traced)
          if (traced < a.trace.cap)
            a.trace.rows[(size_t)e * a.trace.cap + traced] =
                make_int4((int)sm.qhead[cur ^ 1][i][tid], sm.qpay[cur ^ 1][i][0][tid],
                          P::PW > 1 ? sm.qpay[cur ^ 1][i][P::PW > 1 ? 1 : 0][tid] : 0, round + 1);
      }
      cur ^= 1;
      n_cur = resp.n;
    }
    if (TRACK && env_live) a.trace.cnt[e] = traced;

    // ---- post_message_resolution (env.py:175-178)
    for (int s = 0; s < n; ++s) {
      if ((done >> s) & 1u) continue;
      bind(s);
      load_state(s, st);
      P::post(ctx, st);
      store_state(s, st);
    }

    // ---- outputs, strategic agents in order (env.py:273-303; fsm.py:322-378;
    // stackelberg.py:149-194).  Pass 1: callbacks + caches; pass 2 (needs the terminal flag): rows.
    uint32_t obs_slots = 0, rew_slots = 0, t_slots = 0, u_slots = 0;
    for (int s = 0; s < n; ++s) {
      const int sidx = sp.sidx[s];
      if (sidx < 0 || ((done >> s) & 1u)) continue;
      bind(s);
      load_state(s, st);
      float obs_val[P::OBS_DIM] = {};
      bool obs_now = false, rew_now = false;
      float rew_val = 0.f;
      if ((observing >> s) & 1u) obs_now = P::encode(ctx, st, obs_val);
      if (sp.env_kind == PHX_ENV_BASE) {
        if (obs_now) { rew_val = P::reward(ctx, st); rew_now = true; }
      } else if ((rewarded >> s) & 1u) {
        rew_val = P::reward(ctx, st);
        rew_now = true;
        a.reward_cache[(size_t)e * ENGINE1_SLOTS + s] = rew_val;
      }
      if (P::terminated(ctx, st)) t_slots |= 1u << s;
      if (P::truncated(ctx, st)) u_slots |= 1u << s;
      store_state(s, st);
      obs_slots |= (uint32_t)obs_now << s;
      rew_slots |= (uint32_t)rew_now << s;
      if (env_live) {
        const size_t orow = row * S + sidx;
        if (obs_now) {
          if (a.io.obs) {
#pragma unroll
            for (int j = 0; j < P::OBS_DIM; ++j)
              if (j < O) a.io.obs[orow * O + j] = obs_val[j];
          }
          if (sp.env_kind == PHX_ENV_FSM) {
#pragma unroll
            for (int j = 0; j < P::OBS_DIM; ++j)
              if (j < O) a.obs_cache[((size_t)e * ENGINE1_SLOTS + s) * O + j] = obs_val[j];
          }
        }
        if (sp.env_kind == PHX_ENV_BASE && a.io.reward) a.io.reward[orow] = rew_now ? rew_val : 0.f;
      }
    }
    term |= t_slots;
    trunc |= u_slots;
    if (cached_env) {
      rnone &= ~rew_slots;
      if (sp.env_kind == PHX_ENV_FSM) ocached |= obs_slots;
    }
    const bool all_term = __popc(term) == S;
    const bool all_trunc = (h.x == sp.num_steps) || __popc(trunc) == S;
    const bool terminal = all_term || all_trunc;
    if (sp.env_kind == PHX_ENV_FSM) h.z = next_stage;

    if (env_live) {
      for (int s = 0; s < n; ++s) {
        const int sidx = sp.sidx[s];
        if (sidx < 0) continue;
        const size_t orow = row * S + sidx;
        const bool was_done = (done >> s) & 1u;
        const bool obs_now = (obs_slots >> s) & 1u;
        uint8_t om = obs_now, rm = 0;
        if (sp.env_kind == PHX_ENV_BASE) {
          rm = (rew_slots >> s) & 1u;
        } else {
          const bool none = (rnone >> s) & 1u;
          if (sp.env_kind == PHX_ENV_FSM) {
            if (terminal) {  // fsm.py:360-375: flush the caches
              om = (ocached >> s) & 1u;
              if (om && !obs_now && a.io.obs)
                for (int j = 0; j < O; ++j)
                  a.io.obs[orow * O + j] = a.obs_cache[((size_t)e * ENGINE1_SLOTS + s) * O + j];
              rm = none ? 2 : 1;
            } else if (obs_now) {  // fsm.py:378
              rm = none ? 2 : 1;
            }
          } else if (terminal) {  // stackelberg.py:180-187
            rm = none ? 2 : 1;
          } else if (obs_now && !none) {  // stackelberg.py:190-194
            rm = 1;
          }
          if (a.io.reward)
            a.io.reward[orow] = rm == 1 ? a.reward_cache[(size_t)e * ENGINE1_SLOTS + s] : 0.f;
        }
        if (a.io.obs_mask) a.io.obs_mask[orow] = om;
        if (a.io.reward_mask) a.io.reward_mask[orow] = rm;
        if (a.io.term) a.io.term[orow] = was_done ? 255 : (uint8_t)((t_slots >> s) & 1u);
        if (a.io.trunc) a.io.trunc[orow] = was_done ? 255 : (uint8_t)((u_slots >> s) & 1u);
      }
      if (a.io.all_done)
        reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(all_term, all_trunc);
    }

    // ---- PHX_FLAG_AUTO_RESET
    if ((sp.flags & PHX_FLAG_AUTO_RESET) && terminal) {
      h.x = 0;
      h.y += 1;
      h.z = sp.initial_stage;
      term = trunc = 0;
      ctx.step = 0;
      ctx.episode = (uint32_t)h.y;
      ctx.stage = h.z;
      rnone = cached_env ? sp.strategic_mask : 0u;
      for (int s = 0; s < n; ++s) {
        bind(s);
        load_state(s, st);
        P::reset_agent(ctx, st);
        store_state(s, st);
      }
      if (P::VW > 0) {
        for (int s = 0; s < n; ++s) {
          bind(s);
          load_state(s, st);
          int v[P::VW > 0 ? P::VW : 1];
          P::view(ctx, st, v);
#pragma unroll
          for (int w = 0; w < P::VW; ++w) sm.views[w][s][tid] = v[w];
        }
      }
      uint32_t first_obs = sp.strategic_mask;
      if (sp.env_kind == PHX_ENV_FSM) first_obs &= sp.stage_acting[sp.initial_stage];
      if (sp.env_kind == PHX_ENV_STACKELBERG) first_obs &= sp.leaders;
      for (int s = 0; s < n; ++s) {
        const int sidx = sp.sidx[s];
        if (sidx < 0) continue;
        bind(s);
        load_state(s, st);
        float obs_val[P::OBS_DIM] = {};
        bool got = false;
        if ((first_obs >> s) & 1u) got = P::encode(ctx, st, obs_val);
        store_state(s, st);
        if (env_live) {
          const size_t orow = row * S + sidx;
          if (a.io.obs && got) {
#pragma unroll
            for (int j = 0; j < P::OBS_DIM; ++j)
              if (j < O) a.io.obs[orow * O + j] = obs_val[j];
          }
          if (a.io.obs_mask) a.io.obs_mask[orow] = got;
        }
      }
    }
  }

  // ---- write back
  if (env_live) {
    a.hdr[e] = h;
    a.term[e] = term;
    a.trunc[e] = trunc;
    if (cached_env) {
      a.reward_none[e] = rnone;
      if (sp.env_kind == PHX_ENV_FSM) a.obs_cached[e] = ocached;
    }
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) {
      int4* dst = reinterpret_cast<int4*>(a.state + ((size_t)w * sp.E + e) * ENGINE1_SLOTS);
      dst[0] = make_int4(sm.state[w][0][tid], sm.state[w][1][tid], sm.state[w][2][tid], sm.state[w][3][tid]);
      dst[1] = make_int4(sm.state[w][4][tid], sm.state[w][5][tid], sm.state[w][6][tid], sm.state[w][7][tid]);
    }
    if (fault) raise_fault(a.faults, e, fault);
  }
}

}  // namespace phx
