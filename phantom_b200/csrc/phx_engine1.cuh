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
//   agent state [NWORDS][n][threads], views [n][threads], reward cache [n][threads],
//   action ring [2][S][A][threads] (cp.async, one step ahead), two queues [qcap][1+PW][threads]
// sized at launch from the env class actually lowered (Engine1Layout).
#pragma once
#include "phx_engine.cuh"

namespace phx {

// Agent loops: a specialised unit (PHX_JIT_TU) knows the agent count at compile time and unrolls
// them completely, so that every agent's callbacks are compiled for ITS kind, slot and masks;
// the generic build leaves the choice to the compiler, as before.
#ifdef PHX_JIT_TU
#define PHX_AGENT_UNROLL _Pragma("unroll")
#else
#define PHX_AGENT_UNROLL
#endif

constexpr int ENGINE1_BLOCK = 128;
constexpr int ENGINE1_SLOTS = 8;  // == the G of the HBM state layout [NWORDS][E][8]

// Runtime-sized shared-memory layout (word index = (row * ENGINE1_BLOCK + tid)): sized by the
// env class actually lowered (n agents, S strategic, queue bound), not by the family maximum,
// so that small env classes keep enough blocks resident to run in a single wave.
struct Engine1Layout {
  int n, S, qcap;
  int off_state, off_views, off_rcache, off_act, off_qhead, off_qpay, off_tables, words;
  // output staging (bytes from the start of the stage area, which begins at word `off_stage`):
  // one step's rows of a full block of envs in the layout of the global planes
  int off_stage, st_rew, st_om, st_rm, st_term, st_trunc, st_all, stage_bytes;
};

// (run-time form: programs loaded from a user cubin are only known by their descriptor)
__host__ __device__ inline Engine1Layout engine1_layout_rt(int nwords, int vw, int act_dim, int pw,
                                                           int n_agents, int n_strategic, int qcap,
                                                           bool cached_env, int obs_dim,
                                                           bool with_stage);

template <class P>
__host__ __device__ inline Engine1Layout engine1_layout(int n_agents, int n_strategic, int qcap,
                                                        bool cached_env, int obs_dim,
                                                        bool with_stage) {
  return engine1_layout_rt(P::NWORDS, P::VW, P::ACT_DIM, P::PW, n_agents, n_strategic, qcap,
                           cached_env, obs_dim, with_stage);
}

__host__ __device__ inline Engine1Layout engine1_layout_rt(int nwords, int vw, int act_dim, int pw,
                                                           int n_agents, int n_strategic, int qcap,
                                                           bool cached_env, int obs_dim,
                                                           bool with_stage) {
  Engine1Layout L;
  L.n = n_agents; L.S = n_strategic > 0 ? n_strategic : 1; L.qcap = qcap;
  int rows = 0;
  L.off_state = rows; rows += nwords * n_agents;
  L.off_views = rows; rows += vw > 0 ? n_agents : 0;
  L.off_rcache = rows; rows += cached_env ? n_agents : 0;
  L.off_act = rows; rows += 2 * L.S * act_dim;
  L.off_qhead = rows; rows += 2 * qcap;
  L.off_qpay = rows; rows += 2 * qcap * pw;
  L.off_tables = rows * ENGINE1_BLOCK;           // kind_tab / ip0_tab / in_tab, 32 words each
  L.off_stage = (L.off_tables + 3 * ENGINE_MAX_AGENTS + 3) & ~3;  // 16-byte aligned
  const int rowS = ENGINE1_BLOCK * L.S;          // strategic rows of a block
  L.st_rew = rowS * obs_dim * 4;
  L.st_om = L.st_rew + rowS * 4;
  L.st_rm = L.st_om + rowS;
  L.st_term = L.st_rm + rowS;
  L.st_trunc = L.st_term + rowS;
  L.st_all = L.st_trunc + rowS;
  L.stage_bytes = L.st_all + ENGINE1_BLOCK * 2;
  L.words = L.off_stage + (with_stage ? (L.stage_bytes + 3) / 4 : 0);
  return L;
}

// Emission cursor of one env: appends to the queue in program order (== global push order).
template <class P>
struct Emit1 {
  int32_t* sm;
  const EngineSpec* spec;
  Engine1Layout L;
  int which, tid;
  int slot;
  uint32_t out_mask;
  int n;
  uint32_t fault;
  __device__ __forceinline__ int32_t& head(int w, int i) const {
    return sm[(L.off_qhead + w * L.qcap + i) * ENGINE1_BLOCK + tid];
  }
  __device__ __forceinline__ int32_t& pay(int w, int i, int k) const {
    return sm[(L.off_qpay + (w * L.qcap + i) * P::PW + k) * ENGINE1_BLOCK + tid];
  }
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
    if (n >= L.qcap) {
      fault = PHX_FAULT_QUEUE_OVERFLOW;
      return;
    }
    head(which, n) = (int32_t)((uint32_t)slot | ((uint32_t)recv << 8) | ((uint32_t)type << 16));
    pay(which, n, 0) = p0;
    if (P::PW > 1) pay(which, n, P::PW > 1 ? 1 : 0) = p1;
    ++n;
  }
};

// Does the spec source of a specialised build carry a static message schedule?
template <class SP, class = void>
struct SpHasPlan : std::false_type {};
template <class SP>
struct SpHasPlan<SP, std::void_t<decltype(SP::plan())>> : std::true_type {};

template <class P, bool TRACK, class SP>
__device__ __forceinline__ void engine1_step_body(const EngineArgs<P>& a) {
  static_assert(P::VW <= 1, "thread-per-env engine: views of at most one word per agent");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* sm = reinterpret_cast<int32_t*>(smem_raw);
  const EngineSpec& sp = SP::get(a);
  const int tid = threadIdx.x;
  const Engine1Layout L = engine1_layout<P>(sp.n_agents, sp.n_strategic, a.qcap,
                                            sp.env_kind != PHX_ENV_BASE, sp.obs_dim,
                                            a.stage_out != 0);
  int8_t* kind_tab = reinterpret_cast<int8_t*>(sm + L.off_tables);
  int32_t* ip0_tab = sm + L.off_tables + ENGINE_MAX_AGENTS;
  uint32_t* in_tab = reinterpret_cast<uint32_t*>(sm + L.off_tables + 2 * ENGINE_MAX_AGENTS);
  auto ST = [&](int w, int slot) -> int32_t& { return sm[(L.off_state + w * L.n + slot) * ENGINE1_BLOCK + tid]; };
  auto VIEW = [&](int slot) -> int32_t& { return sm[(L.off_views + slot) * ENGINE1_BLOCK + tid]; };
  auto RC = [&](int slot) -> float& {
    return reinterpret_cast<float*>(sm)[(L.off_rcache + slot) * ENGINE1_BLOCK + tid];
  };
  auto ACT = [&](int buf, int k, int j) -> float& {
    return reinterpret_cast<float*>(sm)[(L.off_act + (buf * L.S + k) * P::ACT_DIM + j) * ENGINE1_BLOCK + tid];
  };
  auto QH = [&](int w, int i) -> int32_t& { return sm[(L.off_qhead + w * L.qcap + i) * ENGINE1_BLOCK + tid]; };
  auto QP = [&](int w, int i, int k) -> int32_t& {
    return sm[(L.off_qpay + (w * L.qcap + i) * P::PW + k) * ENGINE1_BLOCK + tid];
  };
  const int env = blockIdx.x * ENGINE1_BLOCK + tid;
  const bool env_live = env < sp.E;
  const int e = env_live ? env : sp.E - 1;
  const int n = sp.n_agents, S = sp.n_strategic, O = sp.obs_dim;

  if (tid < ENGINE_MAX_AGENTS) {
    kind_tab[tid] = sp.kind[tid];
    ip0_tab[tid] = sp.agent_iparam[tid][0];
    uint32_t in = 0;
    for (int s = 0; s < n; ++s) in |= ((sp.adj[s] >> tid) & 1u) << s;
    in_tab[tid] = in;
  }
  // ---- load env header, done sets, agent state (once per launch)
  int4 h = a.hdr[e];
  uint32_t term = a.term[e], trunc = a.trunc[e];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) {
    const int4* src = reinterpret_cast<const int4*>(a.state + ((size_t)w * sp.E + e) * ENGINE1_SLOTS);
    const int4 lo = src[0], hi = src[1];
    const int v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) ST(w, k) = v[k];
  }
  const bool cached_env = sp.env_kind != PHX_ENV_BASE;
  uint32_t rnone = 0, ocached = 0;
  if (cached_env) {
    rnone = a.reward_none[e];
    if (sp.env_kind == PHX_ENV_FSM) ocached = a.obs_cached[e];
    const float4* rc = reinterpret_cast<const float4*>(a.reward_cache + (size_t)e * ENGINE1_SLOTS);
    const float4 lo = rc[0], hi = rc[1];
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) RC(k) = v[k];
  }
  // actions of step t+1 are fetched with cp.async while step t runs
  auto prefetch_actions = [&](int t_next) {
    if (t_next < a.T) {
      const float* src = a.io.actions + ((size_t)t_next * sp.E + e) * S * P::ACT_DIM;
      for (int k = 0; k < S; ++k)
#pragma unroll
        for (int j = 0; j < P::ACT_DIM; ++j)
          cp_async4(&ACT(t_next & 1, k, j), src + k * P::ACT_DIM + j);
    }
    cp_async_commit();
  };
  prefetch_actions(0);
  __syncthreads();
  uint32_t fault = 0;  // first fault in event order: events ARE sequential here

  // ---- OUTPUT STAGING.  With a thread per env a direct store of one float touches its own
  // 32-byte sector (the rows of consecutive envs are S * O floats apart): C4 issued 29 sector
  // writes per env-step, an order of magnitude more L1->L2 traffic than payload.  A block whose
  // 128 envs are all live instead builds the step's rows in shared memory, in the layout of the
  // global planes (the rows of consecutive envs are contiguous there), and thread 0 writes each
  // plane segment with ONE TMA bulk store (cp.async.bulk, SASS UBLKCP).  Rows the reference
  // would not return (mask 0) are written as zeros.
  const bool staged = a.stage_out != 0 && (blockIdx.x + 1) * ENGINE1_BLOCK <= sp.E;  // block-uniform
  unsigned char* const stage = reinterpret_cast<unsigned char*>(sm + L.off_stage);
  auto put_obs = [&](size_t row, int sidx, int j, float v) {
    if (staged) reinterpret_cast<float*>(stage)[(tid * S + sidx) * O + j] = v;
    else if (env_live && a.io.obs) a.io.obs[(row * S + sidx) * O + j] = v;
  };
  auto put_rew = [&](size_t row, int sidx, float v) {
    if (staged) reinterpret_cast<float*>(stage + L.st_rew)[tid * S + sidx] = v;
    else if (env_live && a.io.reward) a.io.reward[row * S + sidx] = v;
  };
  auto put_u8 = [&](uint8_t* plane, int st_off, size_t row, int sidx, uint8_t v) {
    if (staged) stage[st_off + tid * S + sidx] = v;
    else if (env_live && plane) plane[row * S + sidx] = v;
  };

  constexpr int EW = EnvWords<P>::value;
  int envw[EW > 0 ? EW : 1], envsnap[EW > 0 ? EW : 1];
  if constexpr (EW > 0) {
#pragma unroll
    for (int w = 0; w < EW; ++w) envw[w] = a.env_state[(size_t)w * sp.E + e];
  }

  Ctx ctx;
  ctx.spec = &sp;
  ctx.env = envsnap;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.views = &VIEW(0);
  ctx.view_stride = ENGINE1_BLOCK;  // view word 0 of slot s at views[0][s][tid]
  ctx.kind_tab = kind_tab;
  ctx.ip0_tab = ip0_tab;

  // StochasticNetwork: the env's own 8 x 8 adjacency, one byte per row (rows == columns)
  const bool stochastic = a.adj_env != nullptr;
  uint64_t adj64 = 0;
  if (stochastic) {
    const uint4* src = reinterpret_cast<const uint4*>(a.adj_env + (size_t)e * ENGINE1_SLOTS);
    const uint4 lo = src[0], hi = src[1];
    const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) adj64 |= (uint64_t)(v[k] & 0xFFu) << (8 * k);
  }
  auto row_of = [&](int slot) { return (uint32_t)(adj64 >> (8 * slot)) & 0xFFu; };
  auto bind = [&](int slot) {
    ctx.slot = slot;
    ctx.kind = sp.kind[slot];
    ctx.out_mask = stochastic ? row_of(slot) : sp.adj[slot];
  };
  auto in_mask_of = [&](int slot) { return stochastic ? row_of(slot) : in_tab[slot]; };
  auto load_state = [&](int slot, int* st) {
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) st[w] = ST(w, slot);
  };
  auto store_state = [&](int slot, const int* st) {
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) ST(w, slot) = st[w];
  };

  // Mail that waits for a later step's resolve (a stage handler that does not call
  // resolve_network(), fsm.py:280-283): the resolver's queue outlives the step in the reference.
  // Here it stays in this thread's queue between steps and in HBM between launches.
  int carry_cnt = 0, carry_q = 0;
  if (a.carry != nullptr) {
    carry_cnt = min(a.carry_n[e], a.qcap);
    for (int i = 0; i < carry_cnt; ++i) {
      const int32_t* src = a.carry + ((size_t)e * a.qcap + i) * (1 + P::PW);
      QH(0, i) = src[0];
#pragma unroll
      for (int k = 0; k < P::PW; ++k) QP(0, i, k) = src[1 + k];
    }
  }

  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * sp.E + e;
    h.x += 1;  // env.py:252
    ctx.step = h.x;
    ctx.episode = (uint32_t)h.y;
    ctx.stage = h.z;
    const uint32_t done = term | trunc;  // agents without a context this step (env.py:344-348)
    int st[P::NWORDS > 0 ? P::NWORDS : 1];
    prefetch_actions(t + 1);
    // This step's actions (issued one step ago) have landed.  The wait is unconditional: a step
    // in which no strategic agent acts (a stage of other agents, everybody done) used to skip it,
    // which left two copies into the same ring slot formally in flight (racecheck: "invalid
    // memcpy_async synchronization").
    cp_async_wait<1>();
    if constexpr (EW > 0) {  // the EnvView of this step (env.py:340)
#pragma unroll
      for (int w = 0; w < EW; ++w) envsnap[w] = envw[w];
    }

    // ---- start-of-step snapshot of every agent's public state (network.py:208-222)
    if (P::VW > 0) {
      PHX_AGENT_UNROLL
      for (int s = 0; s < n; ++s) {
        bind(s);
        load_state(s, st);
        int v[P::VW > 0 ? P::VW : 1];
        P::view(ctx, st, v);
#pragma unroll
        for (int w = 0; w < P::VW; ++w) VIEW(s) = v[w];
      }
    }

    // ---- acting phase, agents in order (env.py:320-336; fsm.py:276-277; stackelberg.py:133-140)
    uint32_t acting = 0xFFFFFFFFu, observing = sp.strategic_mask, rewarded = sp.strategic_mask;
    int next_stage = h.z;
    // a stage WITH an env handler is resolved only if the handler does it (fsm.py:280-283), and
    // its next stage is known only after that (fsm.py:294-302)
    bool handled = false, resolves = true;
    if (sp.env_kind == PHX_ENV_FSM) {
      acting = sp.stage_acting[h.z];
      next_stage = sp.stage_next[h.z];
      handled = sp.stage_rule[h.z][SR_HANDLER] != 0;
      resolves = !handled || sp.stage_rule[h.z][SR_RESOLVES] != 0;
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
    if constexpr (SpHasPlan<SP>::value && !TRACK) {
      // ---- STATIC SCHEDULE (see StaticPlan): acting phase, hooks and resolver rounds of this
      // step's phase with every potential message in a fixed slot
      const int phase = sp.env_kind == PHX_ENV_FSM ? h.z
                        : sp.env_kind == PHX_ENV_STACKELBERG ? ((h.x & 1) == 1 ? 0 : 1) : 0;
      auto route = [&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if constexpr (PH >= SP::N_PHASES) return;  // (keeps the unused phases out of the unit)
        const StaticPlan& pl = SP::plan();
        int p0[SPL_ROUNDS][SPL_MSGS], p1[SPL_ROUNDS][SPL_MSGS];
        uint32_t valid[SPL_ROUNDS];
#pragma unroll
        for (int r = 0; r < SPL_ROUNDS; ++r) valid[r] = 0u;
        // acting phase, agents in order (env.py:320-336)
#pragma unroll
        for (int s = 0; s < SPL_AGENTS; ++s) {
          if (s >= n || !((acting >> s) & 1u) || ((done >> s) & 1u)) continue;
          bind(s);
          load_state(s, st);
          const int sidx = sp.sidx[s];
          bool has_action = false;
          float act[P::ACT_DIM];
#pragma unroll
          for (int j = 0; j < P::ACT_DIM; ++j) act[j] = 0.f;
          if (sidx >= 0) {
            has_action = a.io.action_mask ? a.io.action_mask[row * S + sidx] != 0 : true;
#pragma unroll
            for (int j = 0; j < P::ACT_DIM; ++j) act[j] = ACT(t & 1, sidx, j);
          }
          PlanEmit em{pl.act_slot[PH][s], pl.act_n[PH][s], pl.recv[PH][0], pl.type[PH][0],
                      p0[0], p1[0], &valid[0], 0, 0u};
          P::act(ctx, st, has_action, act, em);
          store_state(s, st);
          if (em.fault && !fault) fault = em.fault;
        }
        if (!resolves) {  // the mail would wait for a later step's resolve
          if (valid[0] && !fault) fault = PHX_FAULT_UNRESOLVED_MAIL;
          return;
        }
        // pre_message_resolution (env.py:170-173)
        if (P::HAS_PRE) {
#pragma unroll
          for (int s = 0; s < SPL_AGENTS; ++s) {
            if (s >= n || ((done >> s) & 1u)) continue;
            bind(s);
            load_state(s, st);
            P::pre(ctx, st);
            store_state(s, st);
          }
        }
        // BatchResolver.resolve (resolvers.py:128-163) along the plan
#pragma unroll
        for (int r = 0; r < SPL_ROUNDS; ++r) {
          if (r >= pl.n_rounds[PH]) break;
          if (valid[r] == 0u) break;  // queue empty: resolved
          if (sp.round_limit >= 0 && r >= sp.round_limit) {
            if (!fault) fault = PHX_FAULT_ROUND_LIMIT;
            break;
          }
#pragma unroll
          for (int i = 0; i < SPL_MSGS; ++i) {
            if (i >= pl.n_msg[PH][r]) break;
            const int rc = pl.recv[PH][r][i];
            // (mail of done agents is dropped, resolvers.py:143-144; a message that was not
            // sent leaves its handler's response slots invalid)
            if (!((valid[r] >> i) & 1u) || ((done >> rc) & 1u)) continue;
            // delivery-time edge filter (resolvers.py:146-148; matters with ignore_connection_errors)
            if (!((in_mask_of(rc) >> pl.sender[PH][r][i]) & 1u)) continue;
            bind(rc);
            ctx.in_mask = in_mask_of(rc);
            load_state(rc, st);
            Msg m;
            m.sender = pl.sender[PH][r][i];
            m.type = pl.type[PH][r][i];
            m.p[0] = p0[r][i];
            m.p[1] = p1[r][i];
            constexpr int RN = SPL_ROUNDS - 1;
            const int rn = r + 1 < SPL_ROUNDS ? r + 1 : RN;
            PlanEmit em{pl.resp_slot[PH][r][i], r + 1 < SPL_ROUNDS ? pl.resp_n[PH][r][i] : 0,
                        pl.recv[PH][rn], pl.type[PH][rn], p0[rn], p1[rn], &valid[rn], 0, 0u};
            if (!P::handle(ctx, st, m, em) && !fault && !em.fault) fault = PHX_FAULT_UNKNOWN_MSG_TYPE;
            store_state(rc, st);
            if (em.fault && !fault) fault = em.fault;
          }
        }
        // (a plan never has more rounds than SPL_ROUNDS: the host refuses to build it otherwise)
      };
      switch (phase) {
        case 0: route(std::integral_constant<int, 0>{}); break;
        case 1: route(std::integral_constant<int, 1>{}); break;
        case 2: route(std::integral_constant<int, 2>{}); break;
        case 3: route(std::integral_constant<int, 3>{}); break;
        case 4: route(std::integral_constant<int, 4>{}); break;
        case 5: route(std::integral_constant<int, 5>{}); break;
        case 6: route(std::integral_constant<int, 6>{}); break;
        default: route(std::integral_constant<int, 7>{}); break;
      }
    } else {
    int cur = carry_q;  // (0 unless mail is waiting: this step's sends go behind it)
    Emit1<P> out{sm, &sp, L, cur, tid, 0, 0u, carry_cnt, 0u};
    auto act_one = [&](const int s) {
      if (!((acting >> s) & 1u) || ((done >> s) & 1u)) return;
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
        for (int j = 0; j < P::ACT_DIM; ++j) act[j] = ACT(t & 1, sidx, j);
      }
      P::act(ctx, st, has_action, act, out);
      store_state(s, st);
    };
    if (!sp.any_act_order) {  // slot order (the base env; stage lists that ascend in slot)
      PHX_AGENT_UNROLL
      for (int s = 0; s < n; ++s) act_one(s);
    } else {  // the stage's own acting order (fsm.py:276-277, stackelberg.py:133-140)
      const int phase = sp.env_kind == PHX_ENV_FSM ? h.z
                        : sp.env_kind == PHX_ENV_STACKELBERG ? ((h.x & 1) == 1 ? 0 : 1) : 0;
      const int n_act = sp.n_act[phase];
      for (int i = 0; i < n_act; ++i) act_one(sp.act_order[phase][i]);
    }
    if (out.fault && !fault) fault = out.fault;
    int n_cur = out.n;
    int traced = 0;
    if (TRACK && env_live) {  // this step's pushes (waiting mail was traced when it was pushed)
      for (int i = carry_cnt; i < n_cur; ++i, ++traced)
        if (traced < a.trace.cap)
          a.trace.rows[row * a.trace.cap + traced] =
              make_int4(QH(cur, i), QP(cur, i, 0), P::PW > 1 ? QP(cur, i, P::PW > 1 ? 1 : 0) : 0, 0);
    }

    if (!resolves && n_cur > 0) {  // the mail waits for a later step's resolve
      if (a.carry != nullptr) {
        carry_cnt = n_cur;
        carry_q = cur;
      } else if (!fault) {
        fault = PHX_FAULT_UNRESOLVED_MAIL;
      }
      n_cur = 0;
    } else {
      carry_cnt = 0;  // resolved below (or there was nothing)
      carry_q = 0;
    }

    // ---- pre_message_resolution (env.py:170-173)
    if (P::HAS_PRE && resolves)
      PHX_AGENT_UNROLL
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
      Emit1<P> resp{sm, &sp, L, cur ^ 1, tid, 0, 0u, 0, 0u};
      uint32_t seen = 0;
      for (int i = 0; i < n_cur; ++i) {
        const int r = (int)(((uint32_t)QH(cur, i) >> 8) & 0xFFu);
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
          const uint32_t hd = (uint32_t)QH(cur, j);
          if ((int)((hd >> 8) & 0xFFu) != r) continue;
          const int sender = (int)(hd & 0xFFu);
          if (!((ctx.in_mask >> sender) & 1u)) continue;  // delivery-time edge filter (:146-148)
          Msg m;
          m.sender = sender;
          m.type = (int)((hd >> 16) & 0xFFu);
          m.p[0] = QP(cur, j, 0);
          m.p[1] = P::PW > 1 ? QP(cur, j, P::PW > 1 ? 1 : 0) : 0;
          if (!P::handle(ctx, st, m, resp) && !fault && !resp.fault) fault = PHX_FAULT_UNKNOWN_MSG_TYPE;
        }
        if constexpr (P::BATCHED) P::batch_end(ctx, st, resp);
        store_state(r, st);
        if (resp.fault && !fault) fault = resp.fault;
      }
      if (TRACK && env_live) {
        for (int i = 0; i < resp.n; ++i, ++traced)
          if (traced < a.trace.cap)
            a.trace.rows[row * a.trace.cap + traced] =
                make_int4(QH(cur ^ 1, i), QP(cur ^ 1, i, 0),
                          P::PW > 1 ? QP(cur ^ 1, i, P::PW > 1 ? 1 : 0) : 0, round + 1);
      }
      cur ^= 1;
      n_cur = resp.n;
    }
    if (TRACK && env_live) a.trace.cnt[row] = traced;
    }

    // ---- post_message_resolution (env.py:175-178)
    if (P::HAS_POST && resolves)
      PHX_AGENT_UNROLL
      for (int s = 0; s < n; ++s) {
      if ((done >> s) & 1u) continue;
      bind(s);
      load_state(s, st);
      P::post(ctx, st);
      store_state(s, st);
    }
    if constexpr (EW > 0) {  // the env class's own post_message_resolution override
      if (resolves)
        P::env_post(ctx, envw, [&](int s_, auto w_) { return ST(decltype(w_)::value, s_); });
    }

    // ---- the stage's env handler picks the next stage (fsm.py:294-307)
    if (handled) {
      next_stage = stage_rule_pick(sp, h.z, [&](int kind, int rs, int rw, int constant) {
        int v = constant;
        if (kind == PHX_RULE_STEP) {
          v = h.x;
        } else if (kind == PHX_RULE_AGENT_WORD) {
#pragma unroll
          for (int w = 0; w < P::NWORDS; ++w)
            if (w == rw) v = ST(w, rs);
        } else if (kind == PHX_RULE_ENV_WORD) {
          if constexpr (EW > 0) {
#pragma unroll
            for (int w = 0; w < EW; ++w)
              if (w == rw) v = envw[w];
          }
        }
        return v;
      });
      if (!((sp.stage_allowed[h.z] >> next_stage) & 1u)) {
        if (!fault) fault = PHX_FAULT_BAD_TRANSITION;
        next_stage = h.z;
      }
      if (!sp.stage_rewarded_none[h.z]) observing = sp.stage_acting[next_stage];
    }

    // ---- outputs, strategic agents in order (env.py:273-303; fsm.py:322-378;
    // stackelberg.py:149-194).  Pass 1: callbacks + caches; pass 2 (needs the terminal flag): rows.
    uint32_t obs_slots = 0, rew_slots = 0, t_slots = 0, u_slots = 0;
    if (staged) {  // the previous step's bulk stores have read the stage area
      if (tid == 0) bulk_wait_read<0>();
      __syncthreads();
    }
    PHX_AGENT_UNROLL
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
        RC(s) = rew_val;
      }
      if (P::terminated(ctx, st)) t_slots |= 1u << s;
      if (P::truncated(ctx, st)) u_slots |= 1u << s;
      store_state(s, st);
      obs_slots |= (uint32_t)obs_now << s;
      rew_slots |= (uint32_t)rew_now << s;
      if (obs_now) {
#pragma unroll
        for (int j = 0; j < P::OBS_DIM; ++j)
          if (j < O) put_obs(row, sidx, j, obs_val[j]);
        if (env_live && sp.env_kind == PHX_ENV_FSM) {
#pragma unroll
          for (int j = 0; j < P::OBS_DIM; ++j)
            if (j < O) a.obs_cache[((size_t)e * ENGINE1_SLOTS + s) * O + j] = obs_val[j];
        }
      }
      if (sp.env_kind == PHX_ENV_BASE) put_rew(row, sidx, rew_now ? rew_val : 0.f);
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

    if (env_live || staged) {
      PHX_AGENT_UNROLL
      for (int s = 0; s < n; ++s) {
        const int sidx = sp.sidx[s];
        if (sidx < 0) continue;
        const bool was_done = (done >> s) & 1u;
        const bool obs_now = (obs_slots >> s) & 1u;
        bool obs_written = obs_now;
        uint8_t om = obs_now, rm = 0;
        if (sp.env_kind == PHX_ENV_BASE) {
          rm = (rew_slots >> s) & 1u;
          if (was_done && staged) put_rew(row, sidx, 0.f);  // (pass 1 skips done agents)
        } else {
          const bool none = (rnone >> s) & 1u;
          if (sp.env_kind == PHX_ENV_FSM) {
            if (terminal) {  // fsm.py:360-375: flush the caches
              om = (ocached >> s) & 1u;
              if (om && !obs_now) {
                for (int j = 0; j < O; ++j)
                  put_obs(row, sidx, j, a.obs_cache[((size_t)e * ENGINE1_SLOTS + s) * O + j]);
                obs_written = true;
              }
              rm = none ? 2 : 1;
            } else if (obs_now) {  // fsm.py:378
              rm = none ? 2 : 1;
            }
          } else if (terminal) {  // stackelberg.py:180-187
            rm = none ? 2 : 1;
          } else if (obs_now && !none) {  // stackelberg.py:190-194
            rm = 1;
          }
          put_rew(row, sidx, rm == 1 ? RC(s) : 0.f);
        }
        if (staged && !obs_written)  // a staged row is always written: zeros where mask == 0
          for (int j = 0; j < O; ++j) put_obs(row, sidx, j, 0.f);
        put_u8(a.io.obs_mask, L.st_om, row, sidx, om);
        put_u8(a.io.reward_mask, L.st_rm, row, sidx, rm);
        put_u8(a.io.term, L.st_term, row, sidx, was_done ? 255 : (uint8_t)((t_slots >> s) & 1u));
        put_u8(a.io.trunc, L.st_trunc, row, sidx, was_done ? 255 : (uint8_t)((u_slots >> s) & 1u));
      }
      if (staged)
        reinterpret_cast<uchar2*>(stage + L.st_all)[tid] = make_uchar2(all_term, all_trunc);
      else if (a.io.all_done)
        reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(all_term, all_trunc);
    }

    // ---- PHX_FLAG_AUTO_RESET
    if ((sp.flags & PHX_FLAG_AUTO_RESET) && terminal) {
      h.x = 0;
      h.y += 1;
      h.z = sp.initial_stage;
      term = trunc = 0;
      carry_cnt = carry_q = 0;  // Network.reset -> resolver.reset() drops waiting mail
      ctx.step = 0;
      ctx.episode = (uint32_t)h.y;
      ctx.stage = h.z;
      rnone = cached_env ? sp.strategic_mask : 0u;
      if (stochastic) {  // StochasticNetwork.reset: resample, then reset agents (network.py:450-453)
        adj64 = 0;
        for (int c = 0; c < a.n_base; ++c) {
          const uint2 bc = a.base_conn[c];
          const int u = bc.x & 0xFF, v = (bc.x >> 8) & 0xFF;
          if (base_connection_exists(sp.seed, ctx.env_id, ctx.episode, c, bc.y))
            adj64 |= (1ull << (8 * u + v)) | (1ull << (8 * v + u));
        }
      }
      PHX_AGENT_UNROLL
      for (int s = 0; s < n; ++s) {
        bind(s);
        load_state(s, st);
        P::reset_agent(ctx, st);
        store_state(s, st);
      }
      if constexpr (EW > 0) {  // reset() builds a fresh EnvView (fsm.py:232-236)
#pragma unroll
        for (int w = 0; w < EW; ++w) envsnap[w] = envw[w];
      }
      if (P::VW > 0) {
        PHX_AGENT_UNROLL
        for (int s = 0; s < n; ++s) {
          bind(s);
          load_state(s, st);
          int v[P::VW > 0 ? P::VW : 1];
          P::view(ctx, st, v);
#pragma unroll
          for (int w = 0; w < P::VW; ++w) VIEW(s) = v[w];
        }
      }
      uint32_t first_obs = sp.strategic_mask;
      if (sp.env_kind == PHX_ENV_FSM) first_obs &= sp.stage_acting[sp.initial_stage];
      if (sp.env_kind == PHX_ENV_STACKELBERG) first_obs &= sp.leaders;
      PHX_AGENT_UNROLL
      for (int s = 0; s < n; ++s) {
        const int sidx = sp.sidx[s];
        if (sidx < 0) continue;
        bind(s);
        load_state(s, st);
        float obs_val[P::OBS_DIM] = {};
        bool got = false;
        if ((first_obs >> s) & 1u) got = P::encode(ctx, st, obs_val);
        store_state(s, st);
        if (got) {
#pragma unroll
          for (int j = 0; j < P::OBS_DIM; ++j)
            if (j < O) put_obs(row, sidx, j, obs_val[j]);
        }
        put_u8(a.io.obs_mask, L.st_om, row, sidx, got);
      }
    }
    if (staged) {  // this step's rows: one bulk store per plane segment of the block
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        const size_t r0 = ((size_t)t * sp.E + (size_t)blockIdx.x * ENGINE1_BLOCK) * S;  // first row
        const uint32_t rows = (uint32_t)(ENGINE1_BLOCK * S);
        if (a.io.obs) bulk_store(a.io.obs + r0 * O, stage, rows * (uint32_t)O * 4u);
        if (a.io.reward) bulk_store(a.io.reward + r0, stage + L.st_rew, rows * 4u);
        if (a.io.obs_mask) bulk_store(a.io.obs_mask + r0, stage + L.st_om, rows);
        if (a.io.reward_mask) bulk_store(a.io.reward_mask + r0, stage + L.st_rm, rows);
        if (a.io.term) bulk_store(a.io.term + r0, stage + L.st_term, rows);
        if (a.io.trunc) bulk_store(a.io.trunc + r0, stage + L.st_trunc, rows);
        if (a.io.all_done)
          bulk_store(a.io.all_done + ((size_t)t * sp.E + (size_t)blockIdx.x * ENGINE1_BLOCK) * 2,
                     stage + L.st_all, (uint32_t)ENGINE1_BLOCK * 2u);
        bulk_commit();
      }
    }
  }
  if (staged && tid == 0) bulk_wait<0>();

  // ---- write back
  if (env_live) {
    if (a.carry != nullptr) {
      a.carry_n[e] = carry_cnt;
      for (int i = 0; i < carry_cnt; ++i) {
        int32_t* dst = a.carry + ((size_t)e * a.qcap + i) * (1 + P::PW);
        dst[0] = QH(carry_q, i);
#pragma unroll
        for (int k = 0; k < P::PW; ++k) dst[1 + k] = QP(carry_q, i, k);
      }
    }
    a.hdr[e] = h;
    a.term[e] = term;
    a.trunc[e] = trunc;
    if (cached_env) {
      a.reward_none[e] = rnone;
      if (sp.env_kind == PHX_ENV_FSM) a.obs_cached[e] = ocached;
      float4* rc = reinterpret_cast<float4*>(a.reward_cache + (size_t)e * ENGINE1_SLOTS);
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = k < n ? RC(k) : 0.f;
      rc[0] = make_float4(v[0], v[1], v[2], v[3]);
      rc[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if constexpr (EW > 0) {
#pragma unroll
      for (int w = 0; w < EW; ++w) a.env_state[(size_t)w * sp.E + e] = envw[w];
    }
    if (stochastic) {
      uint4* dst = reinterpret_cast<uint4*>(a.adj_env + (size_t)e * ENGINE1_SLOTS);
      dst[0] = make_uint4(row_of(0), row_of(1), row_of(2), row_of(3));
      dst[1] = make_uint4(row_of(4), row_of(5), row_of(6), row_of(7));
    }
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) {
      int4* dst = reinterpret_cast<int4*>(a.state + ((size_t)w * sp.E + e) * ENGINE1_SLOTS);
      int v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = k < n ? ST(w, k) : 0;
      dst[0] = make_int4(v[0], v[1], v[2], v[3]);
      dst[1] = make_int4(v[4], v[5], v[6], v[7]);
    }
    if (fault) raise_fault(a.faults, e, fault);
  }
}

template <class P, bool TRACK>
__global__ void __launch_bounds__(ENGINE1_BLOCK) engine1_step_kernel(const EngineArgs<P> a) {
  engine1_step_body<P, TRACK, SpecFromArgs>(a);
}

}  // namespace phx
