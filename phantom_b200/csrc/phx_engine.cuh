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
#include <type_traits>

#include "phx_common.cuh"
#include "phx_rng.cuh"

namespace phx {

// ENV-LEVEL state and hooks (optional part of the device-program interface).  An env class may
// keep state of its own, publish it to the agents as extra EnvView fields and update it in an
// override of PhantomEnv.post_message_resolution (phantom/env.py:175-178) -- the reference's
// simple_market example does (simple_mkt_env.py:46-58: avg_price = mean of the sellers'
// prices).  A program opts in with
//     static constexpr int ENVW = <words>;                         // persisted int32 [ENVW][E]
//     template <class Acc> static void env_post(const Ctx&, int* env, Acc agent_word);
// `env_post` runs after the agents' post hooks; `agent_word(slot, w)` reads word w of any
// agent's state (w must be a compile-time constant at the call site).  `Ctx::env` is the
// START-OF-STEP snapshot of the env words, because the reference builds the EnvView once, in
// _make_ctxs (env.py:338-348), before any message is handled.
template <class P, class = void>
struct EnvWords {
  static constexpr int value = 0;
};
template <class P>
struct EnvWords<P, std::void_t<decltype(P::ENVW)>> {
  static constexpr int value = P::ENVW;
};

constexpr int ENGINE_BLOCK = 128;
constexpr int ENGINE_MAX_AGENTS = 32;

// Generic part of phx_spec in device form.  One definition for both widths of an env tile:
//   EngineSpec  <= 32 agents, a mask over slots is ONE word; passed as a kernel parameter
//               (constant bank) or, in a specialised build, a compile-time constant
//   WideSpec    <= 128 agents, masks of PHX_MASK_WORDS words; lives in global memory
//               (phx_engine_wide.cuh)
struct WMask {
  uint32_t w[PHX_MASK_WORDS];
};
template <int MAXA, class M>
struct EngineSpecT {
  int32_t E, n_agents, n_strategic, num_steps, round_limit, env_kind;
  uint32_t flags;
  int32_t obs_dim, act_dim;
  int8_t kind[MAXA];
  int8_t sidx[MAXA];            // strategic index or -1
  M adj[MAXA];                  // bit r of adj[s]: edge s -> r
  M sender_ok[PHX_MAX_TYPES];   // bit s: slot s may send this payload type
  M receiver_ok[PHX_MAX_TYPES];
  M strategic_mask;             // over slots
  M kind_mask[8];               // slots of each agent kind
  int32_t n_stages, initial_stage;
  M stage_acting[PHX_MAX_STAGES];
  M stage_rewarded[PHX_MAX_STAGES];
  uint8_t stage_rewarded_none[PHX_MAX_STAGES];
  int8_t stage_next[PHX_MAX_STAGES];
  M leaders, followers;         // over slots
  uint64_t seed;
  uint32_t env_offset;
  int32_t iparams[PHX_MAX_PARAMS];
  float fparams[PHX_MAX_PARAMS];
  double dparams[4];                   // family parameters that must stay float64
  int32_t agent_iparam[MAXA][4];
  double agent_fparam[MAXA][4];
  int32_t codec_op[MAXA][PHX_MAX_CODEC_OPS];  // opcode | length << 8
  float codec_val[MAXA][PHX_MAX_CODEC_OPS];
  // device form of the stages' env handlers (fsm.py:294-307; include/phx.h phx_stage.rule_*):
  // an if / elif / else chain, each branch a conjunction of comparisons between two operands
  int8_t stage_rule[PHX_MAX_STAGES][4];  // {handler, resolves, n_branches, else}
  int8_t rule_branch[PHX_MAX_STAGES][PHX_RULE_BRANCHES][2];                    // {n_terms, then}
  int8_t rule_term[PHX_MAX_STAGES][PHX_RULE_BRANCHES][PHX_RULE_TERMS][8];      // RT_*
  int32_t rule_rhs[PHX_MAX_STAGES][PHX_RULE_BRANCHES][PHX_RULE_TERMS];         // PHX_RULE_CONST
  uint8_t stage_allowed[PHX_MAX_STAGES];  // FSMStage.next_stages as a stage bitmask
  // Acting ORDER per phase (FSM: stage; Stackelberg: 0 leaders' / 1 followers' turn; else 0): the
  // reference walks the user's list (fsm.py:276-277, stackelberg.py:133-140), the push order of
  // a step's mail follows it.  any_act_order == 0: every list is ascending in slot (the engines
  // keep their slot-order loops); else the first n_act[ph] entries of act_order[ph].
  uint8_t any_act_order;
  uint8_t n_act[PHX_MAX_STAGES];
  int8_t act_order[PHX_MAX_STAGES][MAXA];
};
using EngineSpec = EngineSpecT<ENGINE_MAX_AGENTS, uint32_t>;
using WideSpec = EngineSpecT<PHX_MAX_AGENTS, WMask>;

enum { SR_HANDLER = 0, SR_RESOLVES, SR_BRANCHES, SR_ELSE };
enum { RT_LHS = 0, RT_SLOT, RT_WORD, RT_CMP, RT_RHS, RT_RHS_SLOT, RT_RHS_WORD };

__device__ __forceinline__ bool rule_compare(int cmp, int lhs, int rhs) {
  if (cmp & PHX_CMP_F32) {  // float32 operands (IEEE comparisons, as numpy's float32 scalars)
    const float a = __int_as_float(lhs), b = __int_as_float(rhs);
    switch (cmp & 7) {
      case PHX_CMP_LT: return a < b;
      case PHX_CMP_LE: return a <= b;
      case PHX_CMP_EQ: return a == b;
      case PHX_CMP_NE: return a != b;
      case PHX_CMP_GE: return a >= b;
      default: return a > b;
    }
  }
  switch (cmp) {
    case PHX_CMP_LT: return lhs < rhs;
    case PHX_CMP_LE: return lhs <= rhs;
    case PHX_CMP_EQ: return lhs == rhs;
    case PHX_CMP_NE: return lhs != rhs;
    case PHX_CMP_GE: return lhs >= rhs;
    default: return lhs > rhs;
  }
}

// `if <terms of branch 0>: return then_0 / elif ...: return then_1 / ... / return else` of a
// stage handler.  `operand(kind, slot, word, constant)` reads one operand from the env's state
// (uniform over the threads of an env: the rule is part of the spec).
// (a specialised unit folds the whole chain: the spec is a constant there, so unroll it)
#ifdef PHX_JIT_TU
#define PHX_RULE_UNROLL _Pragma("unroll")
#else
#define PHX_RULE_UNROLL _Pragma("unroll 1")
#endif
template <class Spec, class Operand>
__device__ __forceinline__ int stage_rule_pick(const Spec& sp, int stage, Operand&& operand) {
  const int nb = sp.stage_rule[stage][SR_BRANCHES];
  PHX_RULE_UNROLL
  for (int b = 0; b < PHX_RULE_BRANCHES; ++b) {
    if (b >= nb) break;
    bool all = true;
    const int nt = sp.rule_branch[stage][b][0];
    PHX_RULE_UNROLL
    for (int k = 0; k < PHX_RULE_TERMS; ++k) {
      if (k >= nt) break;
      const int8_t* t = sp.rule_term[stage][b][k];
      if (t[RT_LHS] == PHX_RULE_ALWAYS) continue;
      const int lhs = operand((int)t[RT_LHS], (int)t[RT_SLOT], (int)t[RT_WORD], 0);
      const int rhs = operand((int)t[RT_RHS], (int)t[RT_RHS_SLOT], (int)t[RT_RHS_WORD],
                              sp.rule_rhs[stage][b][k]);
      all = all && rule_compare(t[RT_CMP], lhs, rhs);
    }
    if (all) return sp.rule_branch[stage][b][1];
  }
  return sp.stage_rule[stage][SR_ELSE];
}

struct Msg {
  int sender, type;
  int p[2];
};

// STATIC MESSAGE SCHEDULE of one env class (thread-per-env engine, specialised builds only).
// On most env classes WHO may send WHAT to WHOM in which resolver round does not depend on the
// data: it follows from the graph, the stage / turn and the device program.  A program that
// declares its potential sends (act_sends / handle_sends, see fam_stackelberg.cu) lets the host
// walk the reference's routing rules (SURVEY.md A.1 rules 2-7: acting order, Network.send checks,
// first-arrival receiver order, delivery filter, responses into the next round) ONCE per phase
// (PhantomEnv: one; FSM: per stage; Stackelberg: leaders' / followers' turn) into this table.
// The specialised kernel then keeps every potential message in a FIXED slot -- registers after
// unrolling -- with a valid bit, and calls the handlers in the planned order: no queue in shared
// memory, no receiver scan, no send checks at run time (they were evaluated on the plan).  A send
// the plan does not contain raises PHX_FAULT_PLAN_MISMATCH instead of being mis-routed.
// Messages of a round are stored grouped by receiver, receivers in first-arrival order, a
// receiver's batch in push order.  The plan is only built if no receiver of a response round
// can hear from two different senders (then the order in which the responders of the previous
// round are visited cannot be observed); every other env class keeps the dynamic queue.
constexpr int SPL_PHASES = 8, SPL_ROUNDS = 4, SPL_MSGS = 16, SPL_AGENTS = 8, SPL_RESP = 4;
struct StaticPlan {
  int32_t n_phases;
  int8_t n_rounds[SPL_PHASES];
  int8_t n_msg[SPL_PHASES][SPL_ROUNDS];
  int8_t sender[SPL_PHASES][SPL_ROUNDS][SPL_MSGS];
  int8_t recv[SPL_PHASES][SPL_ROUNDS][SPL_MSGS];
  int8_t type[SPL_PHASES][SPL_ROUNDS][SPL_MSGS];
  // the slots of round r+1 that the handler of message i of round r may fill, in emission order
  int8_t resp_n[SPL_PHASES][SPL_ROUNDS][SPL_MSGS];
  int8_t resp_slot[SPL_PHASES][SPL_ROUNDS][SPL_MSGS][SPL_RESP];
  // the slots of round 0 that agent s's acting-phase sends may fill, in emission order (one
  // agent's sends are contiguous in push order but not in the receiver-grouped storage)
  int8_t act_n[SPL_PHASES][SPL_AGENTS];
  int8_t act_slot[SPL_PHASES][SPL_AGENTS][SPL_MSGS];
};

// Emission cursor of one callback under a static plan: the callback's potential sends are the
// slots `slots[0..n)` of the target round, in emission order; an actual send takes the next
// potential send with the same (receiver, type) -- actual sends are a subsequence of the
// potential ones -- and anything else is a plan mismatch.  After inlining into an unrolled
// specialised kernel every index here is a compile-time constant.
struct PlanEmit {
  const int8_t* slots;
  int n;
  const int8_t* recv;
  const int8_t* type;
  int* p0;
  int* p1;
  uint32_t* valid;
  int k;
  uint32_t fault;
  __device__ __forceinline__ void send(int rcv, int typ, int a, int b = 0) {
    if (fault) return;
#pragma unroll
    for (int j = 0; j < SPL_MSGS; ++j) {
      if (j < k || j >= n) continue;
      const int sl = slots[j];
      if (recv[sl] == rcv && type[sl] == typ) {
        p0[sl] = a;
        p1[sl] = b;
        *valid |= 1u << sl;
        k = j + 1;
        return;
      }
    }
    fault = PHX_FAULT_PLAN_MISMATCH;
  }
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
  uint32_t out_mask;   // adjacency row of this agent: bit r = edge slot -> r
  uint32_t in_mask;    // adjacency column: bit s = edge s -> slot
  const int* views;    // start-of-step snapshot, [slot][VW] (this env's tile, shared memory)
  int view_stride;
  const int* env;      // start-of-step snapshot of the env-level words (custom EnvView fields)
  const int8_t* kind_tab;   // per-slot tables staged in shared memory (dynamic lookups by
  const int32_t* ip0_tab;   // sender / receiver slot would serialise on the constant bank)
  __device__ __forceinline__ float proportion_time_elapsed() const {
    // EnvView.proportion_time_elapsed = current_step / num_steps in float64 (env.py:166-168)
    return (float)((double)step / (double)spec->num_steps);
  }
  __device__ __forceinline__ const int* view_of(int other_slot) const {
    return views + other_slot * view_stride;
  }
  __device__ __forceinline__ bool has_neighbour(int other_slot) const {
    return (out_mask >> other_slot) & 1u;
  }
  // neighbours of a given agent kind, as a slot bitmask (iterate with __ffs)
  __device__ __forceinline__ uint32_t neighbours_of_kind(int k) const {
    return out_mask & spec->kind_mask[k];
  }
  __device__ __forceinline__ int kind_of(int other_slot) const { return kind_tab[other_slot]; }
  __device__ __forceinline__ int iparam0_of(int other_slot) const { return ip0_tab[other_slot]; }
  __device__ __forceinline__ uint32_t rand24_hi(uint32_t stream, uint32_t idx) const {
    return rng_d24_hi(spec->seed, env_id, episode, (uint32_t)step, stream, idx);
  }
  // Draw i of a PACKED site (phx_rng.cuh rng_packed_randint: K base-n digits per 32-bit word, the
  // words of a stream four to a Philox block) through a one-entry cache of the last block: the
  // customers of an env read the SAME word of a step and four consecutive steps the same block,
  // so a thread that plays several agents (thread-per-env engine) evaluates Philox once per four
  // steps instead of once per customer and step (C2 on the generic engine: -40 % instructions).
  mutable uint32_t pk_blk = 0xFFFFFFFFu, pk_stream = 0u, pk_episode = 0u;
  mutable Philox4 pk = {};
  __device__ __forceinline__ int packed_randint(uint32_t stream, uint32_t n, uint32_t kpw,
                                                uint32_t W, uint32_t i) const {
    const uint32_t g = (uint32_t)step * W + i / kpw;
    if (pk_blk != (g >> 2) || pk_stream != stream || pk_episode != episode) {
      pk = rng_word_block(spec->seed, env_id, episode, g >> 2, stream);
      pk_blk = g >> 2;
      pk_stream = stream;
      pk_episode = episode;
    }
    const uint32_t q = g & 3u;
    uint32_t x = pk.w[0];
    if (q == 1u) x = pk.w[1];
    if (q == 2u) x = pk.w[2];
    if (q == 3u) x = pk.w[3];
    int d = 0;
    for (uint32_t r = 0; r <= i % kpw; ++r) d = rng_next_digit(x, n);
    return d;
  }
  // Width-independent iteration (the same program text runs on the 128-lane block engine, whose
  // masks are PHX_MASK_WORDS words -- phx_engine_wide.cuh WCtx):
  //     for (int r = c.next_neighbour(-1); r >= 0; r = c.next_neighbour(r)) ...
  static constexpr int MASK_WORDS = 1;
  __device__ __forceinline__ static int next_bit(uint32_t m, int after) {
    m = after >= 31 ? 0u : (m & ~((2u << after) - 1u));  // after == -1: (2u << -1) is avoided below
    return m ? __ffs(m) - 1 : -1;
  }
  __device__ __forceinline__ int next_neighbour(int after) const {
    return after < 0 ? (out_mask ? __ffs(out_mask) - 1 : -1) : next_bit(out_mask, after);
  }
  // agents of a kind in the env class (whether connected or not), in slot order
  __device__ __forceinline__ int next_of_kind(int k, int after) const {
    const uint32_t m = spec->kind_mask[k];
    return after < 0 ? (m ? __ffs(m) - 1 : -1) : next_bit(m, after);
  }
};

// Segmented queue of one env in shared memory.  Entry (seg, k): head word + PW payload
// words; the segment is the SENDER's slot.  Two layouts behind one accessor interface:
//   TileQueue     every segment has SEGCAP entries, word-major ([k][seg]): the lanes of a tile
//                 writing their own segments hit different banks
//   CompactQueue  segment s occupies [base[s], base[s+1]) of one flat array; the per-segment
//                 capacities come from the lowered env class (P::act_cap(kind, out-degree)).
//                 A 32-agent market step holds at most 230 acting-phase messages, the
//                 fixed layout reserves 32 x 32 = 1024 entries for them: compacting takes the
//                 tile from 13 KB to 6 KB of shared memory and doubles the resident warps.
template <int G, int SEGCAP, int PW>
struct TileQueue {
  uint16_t head[SEGCAP][G];   // recv | type << 8 (the sender is the segment)
  int32_t pay[PW][SEGCAP][G];
  uint8_t cnt[G];             // entries per segment
  uint8_t order[G];           // segment visiting order
  uint16_t ordcnt[G];         // the same, with the segment's length: order[i] | cnt[order[i]] << 8
                              // (one load per lane instead of two dependent ones)
  int32_t nseg;
  __device__ __forceinline__ uint16_t& hd(int k, int seg) { return head[k][seg]; }
  __device__ __forceinline__ int32_t& py(int w, int k, int seg) { return pay[w][k][seg]; }
  __device__ __forceinline__ int cap_of(int) const { return SEGCAP; }
};

template <int G, int TOTAL, int PW>
struct CompactQueue {
  uint16_t head[TOTAL];
  int32_t pay[PW][TOTAL];
  uint16_t base[G + 1];       // set once per launch from the lowered env class
  uint8_t cnt[G];
  uint8_t order[G];
  uint16_t ordcnt[G];
  int32_t nseg;
  __device__ __forceinline__ uint16_t& hd(int k, int seg) { return head[base[seg] + k]; }
  __device__ __forceinline__ int32_t& py(int w, int k, int seg) { return pay[w][base[seg] + k]; }
  __device__ __forceinline__ int cap_of(int seg) const { return base[seg + 1] - base[seg]; }
};

// A program opts into the compact acting queue with
//     static constexpr int ACTTOTAL = <entries>;
//     __host__ __device__ static int act_cap(int kind, int out_degree);   // sends per step, at most
template <class P, class = void>
struct HasActTotal : std::false_type {};
template <class P>
struct HasActTotal<P, std::void_t<decltype(P::ACTTOTAL)>> : std::true_type {};

// ... and into compact RESPONSE queues with RESPTOTAL / resp_cap (an exchange that answers a
// round with 31 messages next to 31 agents that answer with at most one).
template <class P, class = void>
struct HasRespTotal : std::false_type {};
template <class P>
struct HasRespTotal<P, std::void_t<decltype(P::RESPTOTAL)>> : std::true_type {};

template <class P, int G, bool COMPACT = HasRespTotal<P>::value>
struct RespQueueOf {
  using type = TileQueue<G, P::RESPCAP, P::PW>;
};
template <class P, int G>
struct RespQueueOf<P, G, true> {
  using type = CompactQueue<G, P::RESPTOTAL, P::PW>;
};

template <class P, int G, bool COMPACT = HasActTotal<P>::value>
struct ActQueueOf {
  using type = TileQueue<G, P::ACTCAP, P::PW>;
};
template <class P, int G>
struct ActQueueOf<P, G, true> {
  using type = CompactQueue<G, P::ACTTOTAL, P::PW>;
};

// Emission cursor of one lane: appends to the lane's own segment after the reference's send
// checks (network.py:246-254).
template <class Q>
struct Emit {
  Q* q;
  const EngineSpec* spec;
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
    if (n >= q->cap_of(slot)) {
      fault = PHX_FAULT_QUEUE_OVERFLOW;
      return;
    }
    q->hd(n, slot) = (uint16_t)((uint32_t)recv | ((uint32_t)type << 8));
    q->py(0, n, slot) = p0;
    if (sizeof(q->pay) / sizeof(q->pay[0]) > 1) q->py(sizeof(q->pay) / sizeof(q->pay[0]) > 1 ? 1 : 0, n, slot) = p1;
    ++n;
  }
};

// COLLECTIVE RESOLVE (optional part of the device-program interface, tile engine).  The queue
// restates the reference literally: every message is an entry one lane writes and another lane
// finds.  When an agent talks to MANY peers that is the wrong shape for a warp -- the sender
// lane serialises its sends (C3: a maker writes 24 Quotes, the clearing agent 31 Fills, with
// 1-7 of 32 lanes active), then every receiver searches the queue.  A program may therefore
// resolve a step's mail itself with tile collectives -- the receivers PULL the sender's payload
// with shuffles, first-come-first-served admission is a ballot + popcount rank, a batch
// aggregation is a REDUX -- when the lowered env class has the shape it was written for:
//     static constexpr bool HAS_COLLECTIVE = true;
//     static bool collective_ok(const phx_spec&);            // host: canonical env class?
//     template <int G> __device__ static void step_collective(const Ctx&, int* st, bool has_ctx,
//                     bool acts, bool has_action, const float* action, uint32_t tmask,
//                     uint32_t& fault);
// step_collective replaces [act -> Network.send -> pre_message_resolution -> resolver rounds] of
// one step for the whole tile (all lanes call it) and must leave every agent's state exactly
// as the reference's order of events would (tests: goldens of the unmodified reference, the
// sampled full-size oracle comparison, collective == queue on random tapes).  The queue path
// stays the fallback for every other topology and for message tracking.
template <class P, class = void>
struct HasCollective : std::false_type {};
template <class P>
struct HasCollective<P, std::void_t<decltype(P::HAS_COLLECTIVE)>> : std::true_type {};

template <class P>
struct EngineArgs {
  EngineSpec spec;
  int32_t T;
  int32_t qcap;         // thread-per-env engine: queue bound of this env class
  int32_t stage_out;    // thread-per-env engine: every plane segment of a full block is 16-byte
                        // aligned, so the output rows may go through shared memory + bulk stores
  int32_t collective;   // tile engine: the program resolves the mail itself (step_collective)
  int4* hdr;            // [E] step, episode, stage, -
  uint32_t* term;       // [E] PhantomEnv._terminations as a bitmask over agent slots
  uint32_t* trunc;      // [E]
  int32_t* state;       // [NWORDS][E][G]
  float* reward_cache;  // [E][G]   FSM / Stackelberg `_rewards` (by slot)
  uint32_t* reward_none;// [E]      bit slot: cached reward is None
  float* obs_cache;     // [E][G][O] FSM `_observations` (by slot)
  uint32_t* obs_cached; // [E]      bit slot: an obs is cached
  int32_t* env_state;   // [ENVW][E] env-level words of the program (nullptr if ENVW == 0)
  uint32_t* adj_env;    // [E][G]   StochasticNetwork: per-env adjacency rows (nullptr otherwise)
  const uint2* base_conn;  // [n_base] {u | v << 8, ceil(rate * 2^24)} in insertion order
  int32_t n_base;
  StepIO io;
  FaultSink faults;
  TraceSink trace;
  // thread-per-env engine, FSM env classes with a stage handler that does not resolve: the mail
  // that WAITS for a later step's resolve_network() (fsm.py:280-283 -- the resolver's queue
  // outlives the step), between launches.  nullptr = no such stage.
  int32_t* carry_n;  // [E]
  int32_t* carry;    // [E][qcap][1 + PW]  head (sender | recv << 8 | type << 16), payload words
};

// StochasticNetwork.resample_connectivity (network.py:439-448): base connection c of an env
// exists for this episode iff np.random.random() < rate, i.e. under the RNG contract
// d24(stream 5, step 0, idx c) * 2^-24 < rate  <=>  d24 < ceil(rate * 2^24) (integer compare,
// threshold precomputed on the host).  Edges are always added in both directions, so the row
// of a slot is also its column.
constexpr uint32_t RNG_STREAM_CONNECTIVITY = 5;

__device__ __forceinline__ bool base_connection_exists(uint64_t seed, uint32_t env_id,
                                                        uint32_t episode, int c, uint32_t thr) {
  return (rng_d24_hi(seed, env_id, episode, 0u, RNG_STREAM_CONNECTIVITY, (uint32_t)c) >> 8) < thr;
}

__device__ inline uint32_t resample_adj_row(uint64_t seed, const uint2* base, int n_base,
                                            uint32_t env_id, uint32_t episode, int slot) {
  uint32_t row = 0;
  for (int c = 0; c < n_base; ++c) {
    const uint2 bc = base[c];
    const int u = bc.x & 0xFF, v = (bc.x >> 8) & 0xFF;
    if (u != slot && v != slot) continue;
    if (base_connection_exists(seed, env_id, episode, c, bc.y)) row |= 1u << (u == slot ? v : u);
  }
  return row;
}

__device__ __forceinline__ uint32_t tile_mask(int G) {
  return G >= 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
}

// Shared-memory footprint of one env tile.  The acting phase and the response rounds have
// different fan-outs (a maker quotes 24 takers, but nobody answers a quote), so they get
// separately sized queues: `qa` holds the acting-phase sends (ACTCAP per agent), `qr[2]`
// alternate between the response rounds (RESPCAP per agent).
template <class P, int G>
struct TileSmem {
  typename ActQueueOf<P, G>::type qa;
  typename RespQueueOf<P, G>::type qr[2];
  int32_t views[G][P::VW > 0 ? P::VW : 1];
  int32_t first_idx[G];
  uint16_t sorted[G][P::RECVCAP];  // per receiver: its batch as (segment | entry << 8), push order
  uint8_t rcnt[G];
};

template <class P, int G>
struct BlockSmem {
  int8_t kind_tab[ENGINE_MAX_AGENTS];
  int32_t ip0_tab[ENGINE_MAX_AGENTS];
  TileSmem<P, G> tiles[ENGINE_BLOCK / G];
};

// BatchResolver(shuffle_batches=True), resolvers.py:146-151: the batch is first reduced to the
// messages whose edge still exists, then shuffled.  Contract replacement of
// np.random.shuffle (oracle/harness.py contract_shuffle): Fisher-Yates from the back,
//   for i = n-1 .. 1:  j = randint(i + 1);  swap(list[i], list[j])
// draws from stream 0x100 + receiver slot, idx = k_batch * 256 + (n - 1 - i), where k_batch
// counts the non-empty batches this receiver has shuffled in this env step.  Keyed by receiver,
// so all receiver lanes shuffle concurrently.  Returns the live count.
// (`shuffle_list` is the Fisher-Yates part on an already filtered list; the block engine calls it
// with its own list of queue positions.)
__device__ inline void shuffle_list(uint64_t seed, uint32_t env_id, uint32_t episode, int step,
                                    int slot, uint16_t* list, int live, int& k_batch) {
  if (live == 0) return;
  const uint32_t stream = 0x100u + (uint32_t)slot;
  const uint32_t idx0 = (uint32_t)k_batch * 256u;
  Philox4 blk{};
  uint32_t blk_id = 0xFFFFFFFFu;
  for (int i = live - 1; i > 0; --i) {
    const uint32_t idx = idx0 + (uint32_t)(live - 1 - i);
    if (idx / 5u != blk_id) {
      blk_id = idx / 5u;
      blk = rng_block(seed, env_id, episode, (uint32_t)step, stream, blk_id);
    }
    const uint32_t sl = idx % 5u;
    uint32_t d = rng_slot_hi(blk, 4);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (sl == (uint32_t)q) d = rng_slot_hi(blk, q);
    const int j = rng_randint(d, (uint32_t)(i + 1));
    const uint16_t t = list[i];
    list[i] = list[j];
    list[j] = t;
  }
  ++k_batch;
}
__device__ inline int shuffle_batch(const Ctx& ctx, uint16_t* list, int n, int& k_batch) {
  int live = 0;
  for (int j = 0; j < n; ++j) {
    const uint16_t ent = list[j];
    if ((ctx.in_mask >> (ent & 0xFF)) & 1u) list[live++] = ent;
  }
  shuffle_list(ctx.spec->seed, ctx.env_id, ctx.episode, ctx.step, ctx.slot, list, live, k_batch);
  return live;
}

// One resolver round (resolvers.py:137-158).
//  1. The tile partitions the current queue by receiver with a STABLE counting sort: segments
//     are visited in global push order; the lanes take one entry each, `__match_any_sync`
//     groups the entries addressed to the same receiver, a popcount of the lower peers gives
//     each entry its rank inside the group, and a per-receiver running count gives the group its
//     base -- so receiver r's list `sorted[r][0..n_r)` is its batch in push order
//     (agents.py:110-120) and the global position of its first entry is its first-arrival key
//     (dict insertion order of `messages[receiver]`, resolvers.py:126,142).
//  2. Every receiver lane handles its own batch (drops mail of done agents, applies the
//     delivery-time edge filter) and appends its responses to its segment of `qn`.
//  3. `qn`'s segment order = receivers by first-arrival key (a G-wide rank).
// Returns the number of responses pushed (tile-uniform).
template <class P, int G, bool TRACK, class QC, class QN>
__device__ __forceinline__ int engine_round(const EngineArgs<P>& a, const Ctx& ctx, int* st,
                                            bool has_ctx, QC& qc, QN& qn, TileSmem<P, G>& ts,
                                            int round, uint32_t tmask, uint32_t& fault_key,
                                            int& traced, size_t row, bool trace_lane, int& k_batch) {
  constexpr int INF = 0x7FFFFFFF;
  const bool shuffle = (ctx.spec->flags & PHX_FLAG_SHUFFLE_BATCHES) != 0;
  const int slot = ctx.slot;
  const int lane = threadIdx.x & 31;
  const uint32_t below = (1u << lane) - 1u;
  int32_t* first_idx = ts.first_idx;

  Emit<QN> resp{&qn, ctx.spec, slot, ctx.out_mask, 0, 0u};
  int first = INF;
  bool bad_type = false;
  const int nseg = qc.nseg;
  if constexpr (G >= 32) {
    // ---- 1. stable partition by receiver (pays for wide tiles: a 32-agent market step holds
    // up to 168 messages but a taker's batch is 7; measured +19 % on C3, but -50 % on 8-lane
    // tiles, whose queues are short enough to scan -- see the else branch)
    ts.rcnt[slot] = 0;
    first_idx[slot] = INF;
    __syncwarp(tmask);
    bool overflow = false;
    int pos_base = 0;
    // The segments in visiting order and their lengths, one per lane, fetched ONCE; the loop
    // then walks only the non-empty segments (a ballot) and gets (segment, length) by shuffle.
    // (Walking all 32 segments with two dependent shared-memory byte loads each was a third of
    // all stall samples of the C3 profile: in a market step 7, 24 or 1 segments hold mail.)
    const int my_oc = slot < nseg ? (int)qc.ordcnt[slot] : 0;
    const int my_seg = my_oc & 0xFF, my_cnt = my_oc >> 8;
    for (uint32_t live_segs = __ballot_sync(tmask, my_cnt > 0); live_segs; live_segs &= live_segs - 1) {
      const int si = __ffs(live_segs) - 1;
      const int seg = __shfl_sync(tmask, my_seg, si);
      const int c = __shfl_sync(tmask, my_cnt, si);
      for (int k0 = 0; k0 < c; k0 += G) {
        const int k = k0 + slot;
        const bool valid = k < c;
        const int r = valid ? (int)(qc.hd(valid ? k : 0, seg) & 0xFFu) : 0x100 + slot;
        // lanes holding an entry for the same receiver.  Built from six ballots (validity + the
        // five bits of r): __match_any_sync is one instruction but ~100 cycles of latency, and
        // this loop has nothing to overlap it with -- it was 15 % of all stall samples on C3.
        uint32_t peers = __ballot_sync(tmask, valid);
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          const bool bit = (r >> b) & 1;
          const uint32_t with_bit = __ballot_sync(tmask, bit);
          peers &= bit ? with_bit : ~with_bit;
        }
        const int rank = __popc(peers & below);
        const int base = valid ? ts.rcnt[r] : 0;
        __syncwarp(tmask);  // every peer has read the group's base before its leader bumps it
        if (valid) {
          const int idx = base + rank;
          if (idx < P::RECVCAP) ts.sorted[r][idx] = (uint16_t)(seg | (k << 8));
          else overflow = true;
          if (rank == 0) {  // lowest lane of the group = earliest entry
            if (base == 0) first_idx[r] = pos_base + k;
            ts.rcnt[r] = (uint8_t)min(base + __popc(peers), P::RECVCAP);
          }
        }
        __syncwarp(tmask);
      }
      pos_base += c;
    }
    if (overflow)
      fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) |
                                     PHX_FAULT_QUEUE_OVERFLOW);

    // ---- 2. every receiver handles its batch
    first = first_idx[slot];
    if (has_ctx) {  // done agents: mail dropped silently (resolvers.py:143-144)
      if constexpr (P::BATCHED) P::batch_begin(ctx, st);
      int n_mine = ts.rcnt[slot];
      if (shuffle) n_mine = shuffle_batch(ctx, &ts.sorted[slot][0], n_mine, k_batch);
      for (int j = 0; j < n_mine; ++j) {
        const int ent = ts.sorted[slot][j];
        const int seg = ent & 0xFF, k = ent >> 8;
        if (!((ctx.in_mask >> seg) & 1u)) continue;  // delivery-time edge filter (:146-148)
        Msg m;
        m.sender = seg;  // a segment holds the messages of one sender
        m.type = (int)(qc.hd(k, seg) >> 8);
        m.p[0] = qc.py(0, k, seg);
        m.p[1] = P::PW > 1 ? qc.py(P::PW > 1 ? 1 : 0, k, seg) : 0;
        if (!P::handle(ctx, st, m, resp)) bad_type = true;  // agents.py:140-143
      }
      if constexpr (P::BATCHED) {
        if (first != INF) P::batch_end(ctx, st, resp);
      }
    }
  } else {
    // ---- 1+2 for narrow tiles: every receiver lane scans the (short) queue in push order
    int pos = 0;
    if constexpr (P::BATCHED) {
      if (has_ctx) P::batch_begin(ctx, st);
    }
    int n_mine = 0;  // shuffle only: the batch is collected first, handled after the shuffle
    bool overflow = false;
    for (int si = 0; si < nseg; ++si) {
      const int seg = qc.order[si];
      const int c = qc.cnt[seg];
      for (int k = 0; k < c; ++k, ++pos) {
        const uint32_t hd = qc.hd(k, seg);  // recv | type << 8
        if ((int)(hd & 0xFFu) != slot) continue;
        if (first == INF) first = pos;  // first-arrival position of this receiver
        if (!has_ctx) continue;         // done agent: mail dropped (resolvers.py:143-144)
        if (!((ctx.in_mask >> seg) & 1u)) continue;  // delivery-time edge filter (:146-148)
        if (shuffle) {
          if (n_mine < P::RECVCAP) ts.sorted[slot][n_mine++] = (uint16_t)(seg | (k << 8));
          else overflow = true;
          continue;
        }
        Msg m;
        m.sender = seg;
        m.type = (int)(hd >> 8);
        m.p[0] = qc.py(0, k, seg);
        m.p[1] = P::PW > 1 ? qc.py(P::PW > 1 ? 1 : 0, k, seg) : 0;
        if (!P::handle(ctx, st, m, resp)) bad_type = true;  // agents.py:140-143
      }
    }
    if (shuffle) {
      if (overflow)
        fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) |
                                       PHX_FAULT_QUEUE_OVERFLOW);
      n_mine = shuffle_batch(ctx, &ts.sorted[slot][0], n_mine, k_batch);
      for (int j = 0; j < n_mine; ++j) {
        const int ent = ts.sorted[slot][j];
        const int seg = ent & 0xFF, k = ent >> 8;
        Msg m;
        m.sender = seg;
        m.type = (int)(qc.hd(k, seg) >> 8);
        m.p[0] = qc.py(0, k, seg);
        m.p[1] = P::PW > 1 ? qc.py(P::PW > 1 ? 1 : 0, k, seg) : 0;
        if (!P::handle(ctx, st, m, resp)) bad_type = true;
      }
    }
    if constexpr (P::BATCHED) {
      if (has_ctx && first != INF) P::batch_end(ctx, st, resp);
    }
    first_idx[slot] = first;
  }
  if (bad_type)
    fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) |
                                   PHX_FAULT_UNKNOWN_MSG_TYPE);
  if (resp.fault)
    fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | ((uint32_t)slot << 8) | resp.fault);
  qn.cnt[slot] = (uint8_t)resp.n;
  const int total_next = __reduce_add_sync(tmask, resp.n);
  __syncwarp(tmask);  // orders the shared-memory writes above (first_idx, qn) for the tile

  // ---- 3. next queue's segment order = receivers by first-arrival position
  int rank = 0, nrecv = 0;
#pragma unroll
  for (int j = 0; j < G; ++j) {
    const int fj = first_idx[j];
    nrecv += fj != INF;
    rank += fj < first;
  }
  if (first != INF) {
    qn.order[rank] = (uint8_t)slot;
    qn.ordcnt[rank] = (uint16_t)(slot | (resp.n << 8));
  }
  if (slot == 0) qn.nseg = nrecv;
  __syncwarp(tmask);
  if (TRACK && trace_lane) {
    for (int si = 0; si < nrecv; ++si) {
      const int seg = qn.order[si];
      for (int k = 0; k < qn.cnt[seg]; ++k) {
        if (traced < a.trace.cap)
          a.trace.rows[row * a.trace.cap + traced] =
              make_int4((int)(((uint32_t)qn.hd(k, seg) << 8) | (uint32_t)seg), qn.py(0, k, seg),
                        P::PW > 1 ? qn.py(P::PW > 1 ? 1 : 0, k, seg) : 0, round + 1);
        ++traced;
      }
    }
  }
  return total_next;
}

// The fused step kernel.  P is the device program of a family (see fam_*.cu for the
// interface: view / act / pre / handle / post / encode / reward / terminated / truncated /
// reset_agent).  encode and reward receive the mutable agent state: the reference's
// callbacks may have side effects (the KAT agents count their calls).
// Where a step kernel reads the lowered env class from.  SpecFromArgs: the kernel parameter
// (constant bank; one precompiled kernel serves every env class of a family).  A specialised
// build (phx_engine_host.cuh: jit_source) substitutes a compile-time constant spec, which lets
// the compiler fold every flag, stage table and mask of THIS env class.
struct SpecFromArgs {
  template <class A>
  __device__ __forceinline__ static const EngineSpec& get(const A& a) { return a.spec; }
};

template <class P, int G, bool TRACK, class SP>
__device__ __forceinline__ void engine_step_body(const EngineArgs<P>& a) {
  constexpr int TPB = ENGINE_BLOCK / G;  // env tiles per block
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockSmem<P, G>& bs = *reinterpret_cast<BlockSmem<P, G>*>(smem_raw);

  const EngineSpec& sp = SP::get(a);
  const int tb = threadIdx.x / G;
  const int slot = threadIdx.x % G;
  const int env = blockIdx.x * TPB + tb;
  const bool env_live = env < sp.E;
  const bool is_agent = env_live && slot < sp.n_agents;
  const uint32_t tmask = tile_mask(G);
  const int tile_shift = G >= 32 ? 0 : ((threadIdx.x & 31) / G * G);
  TileSmem<P, G>& ts = bs.tiles[tb];
  const int e = env_live ? env : sp.E - 1;

  if (threadIdx.x < ENGINE_MAX_AGENTS) {
    bs.kind_tab[threadIdx.x] = sp.kind[threadIdx.x];
    bs.ip0_tab[threadIdx.x] = sp.agent_iparam[threadIdx.x][0];
  }
  __syncthreads();

  const int kind = slot < sp.n_agents ? sp.kind[slot] : -1;
  const int sidx = slot < sp.n_agents ? sp.sidx[slot] : -1;
  const bool strategic = sidx >= 0;
  if constexpr (HasActTotal<P>::value) {
    // compact acting queue: segment capacities from the lowered env class, exclusive prefix sums
    // (the host has checked that they fit P::ACTTOTAL)
    int incl = slot < sp.n_agents ? min(P::ACTCAP, P::act_cap(kind, __popc(sp.adj[slot]))) : 0;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      const int v = __shfl_up_sync(tmask, incl, off, G);
      if (slot >= off) incl += v;
    }
    ts.qa.base[slot + 1] = (uint16_t)incl;
    if (slot == 0) ts.qa.base[0] = 0;
    __syncwarp(tmask);
  }
  if constexpr (HasRespTotal<P>::value) {  // the same for the two response queues
    int incl = slot < sp.n_agents ? min(P::RESPCAP, P::resp_cap(kind, __popc(sp.adj[slot]))) : 0;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      const int v = __shfl_up_sync(tmask, incl, off, G);
      if (slot >= off) incl += v;
    }
    ts.qr[0].base[slot + 1] = ts.qr[1].base[slot + 1] = (uint16_t)incl;
    if (slot == 0) ts.qr[0].base[0] = ts.qr[1].base[0] = 0;
    __syncwarp(tmask);
  }
  const uint32_t slot_bit = 1u << slot;
  const int S = sp.n_strategic, O = sp.obs_dim;

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
  // env-level words: every lane of the tile keeps the same copy
  constexpr int EW = EnvWords<P>::value;
  int envw[EW > 0 ? EW : 1], envsnap[EW > 0 ? EW : 1];
  if constexpr (EW > 0) {
#pragma unroll
    for (int w = 0; w < EW; ++w) envw[w] = a.env_state[(size_t)w * sp.E + e];
  }

  Ctx ctx;
  ctx.spec = &sp;
  ctx.slot = slot;
  ctx.kind = kind;
  ctx.env = envsnap;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.views = &ts.views[0][0];
  ctx.view_stride = P::VW > 0 ? P::VW : 1;
  ctx.kind_tab = bs.kind_tab;
  ctx.ip0_tab = bs.ip0_tab;
  ctx.out_mask = slot < sp.n_agents ? sp.adj[slot] : 0u;
  {
    uint32_t in = 0;
    for (int s = 0; s < sp.n_agents; ++s) in |= ((sp.adj[s] >> slot) & 1u) << s;
    ctx.in_mask = in;
  }
  if (a.adj_env) ctx.out_mask = ctx.in_mask = a.adj_env[(size_t)e * G + slot];

  // actions are prefetched one step ahead (a step of this kernel is far longer than the HBM
  // latency), so the acting phase never waits on the load
  float act_next[P::ACT_DIM];
  uint8_t has_next = 1;
#pragma unroll
  for (int j = 0; j < P::ACT_DIM; ++j) act_next[j] = 0.f;
  if (strategic && env_live) {
    const size_t arow = (size_t)e * S + sidx;
#pragma unroll
    for (int j = 0; j < P::ACT_DIM; ++j) act_next[j] = a.io.actions[arow * P::ACT_DIM + j];
    if (a.io.action_mask) has_next = a.io.action_mask[arow];
  }

  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * sp.E + e;  // [T,E] row of this env
    float act[P::ACT_DIM];
#pragma unroll
    for (int j = 0; j < P::ACT_DIM; ++j) act[j] = act_next[j];
    const bool has_action_now = has_next != 0;
    // next step's action is prefetched here on narrow tiles (short steps: the load needs the
    // whole step to land) and after the acting phase on 32-lane tiles (see below)
    if (G < 32 && strategic && env_live && t + 1 < a.T) {
      const size_t arow = (row + sp.E) * S + sidx;
#pragma unroll
      for (int j = 0; j < P::ACT_DIM; ++j) act_next[j] = a.io.actions[arow * P::ACT_DIM + j];
      if (a.io.action_mask) has_next = a.io.action_mask[arow];
    }
    h.x += 1;                                 // env.py:252
    ctx.step = h.x;
    ctx.episode = (uint32_t)h.y;
    ctx.stage = h.z;
    const bool was_done = ((term | trunc) & slot_bit) != 0;
    const bool has_ctx = is_agent && !was_done;  // env.py:344-348: no context for done agents
    if constexpr (EW > 0) {  // the EnvView of this step (env.py:340)
#pragma unroll
      for (int w = 0; w < EW; ++w) envsnap[w] = envw[w];
    }

    // ---- start-of-step snapshot of every agent's public state (network.py:208-222)
    if (P::VW > 0) {
      if (is_agent) P::view(ctx, st, &ts.views[slot][0]);
      __syncwarp(tmask);
    }

    // ---- acting phase (env.py:320-336; fsm.py:276-277; stackelberg.py:133-140)
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
    bool routed = false;
    if constexpr (HasCollective<P>::value) {
      if (a.collective && !TRACK) {  // tile-uniform: the program resolves this step's mail itself
        uint32_t cf = 0;
        P::template step_collective<G>(ctx, st, has_ctx, has_ctx && (acting & slot_bit) != 0,
                                       strategic && has_action_now, act, tmask, cf);
        if (cf) fault_key = min(fault_key, (0u << 16) | ((uint32_t)slot << 8) | cf);
        if (G >= 32 && strategic && env_live && t + 1 < a.T) {  // next step's action (see below)
          const size_t arow = (row + sp.E) * S + sidx;
#pragma unroll
          for (int j = 0; j < P::ACT_DIM; ++j) act_next[j] = a.io.actions[arow * P::ACT_DIM + j];
          if (a.io.action_mask) has_next = a.io.action_mask[arow];
        }
        routed = true;
      }
    }
    if (!routed) {
    Emit<decltype(ts.qa)> out{&ts.qa, &sp, slot, ctx.out_mask, 0, 0u};
    if (has_ctx && (acting & slot_bit)) P::act(ctx, st, strategic && has_action_now, act, out);
    // 32-lane tiles: issued after the acting phase, so that the loaded value is not live across
    // the program's act() call (it was spilled to local memory right after the load, which made
    // the "prefetch" wait for the load: 5 % of all stall samples on C3)
    if (G >= 32 && strategic && env_live && t + 1 < a.T) {
      const size_t arow = (row + sp.E) * S + sidx;
#pragma unroll
      for (int j = 0; j < P::ACT_DIM; ++j) act_next[j] = a.io.actions[arow * P::ACT_DIM + j];
      if (a.io.action_mask) has_next = a.io.action_mask[arow];
    }
    ts.qa.cnt[slot] = (uint8_t)out.n;
    int n_acting_segs = sp.n_agents;
    if (!sp.any_act_order) {  // agents act in slot order: segment i is the i-th visited
      ts.qa.order[slot] = (uint8_t)slot;
      ts.qa.ordcnt[slot] = (uint16_t)(slot | (out.n << 8));
      if (slot == 0) ts.qa.nseg = sp.n_agents;
    } else {  // the stage's own acting order: the segments are visited in list order
      const int phase = sp.env_kind == PHX_ENV_FSM ? ctx.stage
                        : sp.env_kind == PHX_ENV_STACKELBERG ? ((h.x & 1) == 1 ? 0 : 1) : 0;
      n_acting_segs = sp.n_act[phase];
      __syncwarp(tmask);  // the segment lengths of the other lanes
      if (slot < n_acting_segs) {
        const int o = sp.act_order[phase][slot];
        ts.qa.order[slot] = (uint8_t)o;
        ts.qa.ordcnt[slot] = (uint16_t)(o | (ts.qa.cnt[o] << 8));
      }
      if (slot == 0) ts.qa.nseg = n_acting_segs;
    }
    if (out.fault) fault_key = min(fault_key, (0u << 16) | ((uint32_t)slot << 8) | out.fault);
    int pending = __reduce_add_sync(tmask, out.n);
    __syncwarp(tmask);

    int traced = 0;
    const bool trace_lane = TRACK && env_live && slot == 0;
    if (trace_lane) {  // pushes of the acting phase, in global push order
      for (int oi = 0; oi < n_acting_segs; ++oi)
        for (int si = ts.qa.order[oi], k = 0; k < ts.qa.cnt[si]; ++k) {
          if (traced < a.trace.cap)
            a.trace.rows[row * a.trace.cap + traced] =
                make_int4((int)(((uint32_t)ts.qa.hd(k, si) << 8) | (uint32_t)si), ts.qa.py(0, k, si),
                          P::PW > 1 ? ts.qa.py(P::PW > 1 ? 1 : 0, k, si) : 0, 0);
          ++traced;
        }
    }

    // ---- pre_message_resolution for every live context, in agent order (env.py:170-173);
    // hooks only touch their own agent, so order across agents is immaterial
    if (!resolves && pending > 0) {  // the mail would wait for a later step's resolve
      fault_key = min(fault_key, (1u << 16) | (0xFFu << 8) | PHX_FAULT_UNRESOLVED_MAIL);
      pending = 0;
    }
    if (has_ctx && resolves) P::pre(ctx, st);

    // ---- BatchResolver.resolve (resolvers.py:128-163)
    int k_batch = 0;  // batches this receiver has shuffled in this step (shuffle_batches only)
    for (int round = 0; pending > 0; ++round) {
      if (sp.round_limit >= 0 && round >= sp.round_limit) {  // resolvers.py:160-163
        fault_key = min(fault_key, ((uint32_t)(round + 1) << 16) | (0xFFu << 8) | PHX_FAULT_ROUND_LIMIT);
        break;
      }
      if (round == 0)
        pending = engine_round<P, G, TRACK>(a, ctx, st, has_ctx, ts.qa, ts.qr[0], ts, round, tmask,
                                            fault_key, traced, row, trace_lane, k_batch);
      else
        pending = engine_round<P, G, TRACK>(a, ctx, st, has_ctx, ts.qr[(round - 1) & 1],
                                            ts.qr[round & 1], ts, round, tmask, fault_key, traced,
                                            row, trace_lane, k_batch);
    }
    if (trace_lane) a.trace.cnt[row] = traced;

    }

    // ---- post_message_resolution (env.py:175-178), then the env class's own override
    if (has_ctx && resolves) P::post(ctx, st);
    if constexpr (EW > 0) {
      if (resolves)  // uniform over the tile: the stage is an env-level word
        P::env_post(ctx, envw, [&](int s_, auto w_) {
          return __shfl_sync(tmask, st[decltype(w_)::value], s_, G);
        });
    }

    // ---- the stage's env handler picks the next stage (fsm.py:294-307)
    if (handled) {
      next_stage = stage_rule_pick(sp, h.z, [&](int kind, int rs, int rw, int constant) {
        int v = constant;
        if (kind == PHX_RULE_STEP) {
          v = h.x;
        } else if (kind == PHX_RULE_AGENT_WORD) {
          int mine = 0;
#pragma unroll
          for (int w = 0; w < P::NWORDS; ++w)
            if (w == rw) mine = st[w];
          v = __shfl_sync(tmask, mine, rs, G);
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
        fault_key = min(fault_key, (0xFFFEu << 16) | (0xFFu << 8) | PHX_FAULT_BAD_TRANSITION);
        next_stage = h.z;
      }
      if (!sp.stage_rewarded_none[h.z]) observing = sp.stage_acting[next_stage];
    }

    // ---- outputs for strategic agents (env.py:273-303; fsm.py:322-378;
    // stackelberg.py:149-194)
    bool obs_now = false, rew_now = false;
    float obs_val[P::OBS_DIM] = {};
    float rew_val = 0.f;
    bool t_flag = false, u_flag = false;
    if (strategic && has_ctx) {
      if (observing & slot_bit) obs_now = P::encode(ctx, st, obs_val);  // None -> false
      if (sp.env_kind == PHX_ENV_BASE) {
        if (obs_now) {  // env.py:281-284: reward only travels with an observation
          rew_val = P::reward(ctx, st);
          rew_now = true;
        }
      } else if (rewarded & slot_bit) {
        rew_val = P::reward(ctx, st);
        rew_now = true;
        rcache = rew_val;
      }
      t_flag = P::terminated(ctx, st);
      u_flag = P::truncated(ctx, st);
    }
    // tile-wide masks over slots: newly done agents, who observed, who was rewarded
    uint32_t t_slots, u_slots, obs_slots, rew_slots;
    if (G <= 8) {  // one packed reduction instead of four ballots
      const uint32_t packed = ((uint32_t)t_flag | ((uint32_t)u_flag << 8) | ((uint32_t)obs_now << 16) |
                               ((uint32_t)rew_now << 24)) << slot;
      const uint32_t all = __reduce_or_sync(tmask, packed);
      t_slots = all & 0xFFu; u_slots = (all >> 8) & 0xFFu;
      obs_slots = (all >> 16) & 0xFFu; rew_slots = all >> 24;
    } else {
      t_slots = __ballot_sync(tmask, t_flag) >> tile_shift;
      u_slots = __ballot_sync(tmask, u_flag) >> tile_shift;
      obs_slots = __ballot_sync(tmask, obs_now) >> tile_shift;
      rew_slots = __ballot_sync(tmask, rew_now) >> tile_shift;
      if (G < 32) {
        const uint32_t m = (1u << G) - 1u;
        t_slots &= m; u_slots &= m; obs_slots &= m; rew_slots &= m;
      }
    }
    term |= t_slots;
    trunc |= u_slots;
    if (cached_env) {
      rnone &= ~rew_slots;  // _rewards.update(rewards)
      if (sp.env_kind == PHX_ENV_FSM) ocached |= obs_slots;
    }
    const bool all_term = __popc(term) == S;                             // env.py:308-310
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
        float* oc = a.obs_cache + ((size_t)e * G + slot) * O;
        if (obs_now) {
#pragma unroll
          for (int j = 0; j < P::OBS_DIM; ++j)
            if (j < O) oc[j] = obs_val[j];
        }
        if (terminal) {  // fsm.py:360-375: flush the caches
          om = (ocached >> slot) & 1u;
          if (om && !obs_now) {
#pragma unroll
            for (int j = 0; j < P::OBS_DIM; ++j)
              if (j < O) obs_val[j] = oc[j];
          }
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
      if (a.io.obs && om) {
#pragma unroll
        for (int j = 0; j < P::OBS_DIM; ++j)
          if (j < O) a.io.obs[orow * O + j] = obs_val[j];
      }
      if (a.io.obs_mask) a.io.obs_mask[orow] = om;
      if (a.io.reward) a.io.reward[orow] = rm == 1 ? r_out : 0.f;
      if (a.io.reward_mask) a.io.reward_mask[orow] = rm;
      if (a.io.term) a.io.term[orow] = was_done ? 255 : (uint8_t)t_flag;
      if (a.io.trunc) a.io.trunc[orow] = was_done ? 255 : (uint8_t)u_flag;
    }
    if (env_live && slot == 0 && a.io.all_done)
      reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(all_term, all_trunc);

    // ---- PHX_FLAG_AUTO_RESET: the step that ends the episode also resets the env
    if ((sp.flags & PHX_FLAG_AUTO_RESET) && terminal) {
      h.x = 0;
      h.y += 1;
      h.z = sp.initial_stage;
      term = trunc = 0;
      ctx.step = 0;
      ctx.episode = (uint32_t)h.y;
      ctx.stage = h.z;
      if (a.adj_env)  // Network.reset of a StochasticNetwork resamples first (network.py:450-453)
        ctx.out_mask = ctx.in_mask =
            is_agent ? resample_adj_row(sp.seed, a.base_conn, a.n_base, ctx.env_id, ctx.episode, slot)
                     : 0u;
      if (is_agent) P::reset_agent(ctx, st);
      rnone = cached_env ? sp.strategic_mask : 0u;
      if constexpr (EW > 0) {  // reset() builds a fresh EnvView (fsm.py:232-236)
#pragma unroll
        for (int w = 0; w < EW; ++w) envsnap[w] = envw[w];
      }
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
        if (first_obs & slot_bit) got = P::encode(ctx, st, obs_val);
        if (a.io.obs && got) {
#pragma unroll
          for (int j = 0; j < P::OBS_DIM; ++j)
            if (j < O) a.io.obs[orow * O + j] = obs_val[j];
        }
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
      if constexpr (EW > 0) {
#pragma unroll
        for (int w = 0; w < EW; ++w) a.env_state[(size_t)w * sp.E + e] = envw[w];
      }
    }
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * G + slot] = st[w];
    if (cached_env) a.reward_cache[(size_t)e * G + slot] = rcache;
    if (a.adj_env) a.adj_env[(size_t)e * G + slot] = ctx.out_mask;
  }
  // first fault of the env in event order (phase, then agent order)
  const uint32_t fk = __reduce_min_sync(tmask, fault_key);
  if (env_live && slot == 0 && fk != 0xFFFFFFFFu) raise_fault(a.faults, e, fk & 0xFFu);
}

template <class P, int G, bool TRACK>
__global__ void __launch_bounds__(ENGINE_BLOCK) engine_step_kernel(const EngineArgs<P> a) {
  engine_step_body<P, G, TRACK, SpecFromArgs>(a);
}

// PhantomEnv.reset / FiniteStateMachineEnv.reset / StackelbergEnv.reset for masked envs.
template <class P, int G>
__device__ __forceinline__ void engine_reset_body(const EngineArgs<P>& a, const uint8_t* env_mask,
                                                  float* obs, uint8_t* obs_mask, bool agents_only) {
  constexpr int TPB = ENGINE_BLOCK / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockSmem<P, G>& bs = *reinterpret_cast<BlockSmem<P, G>*>(smem_raw);
  const EngineSpec& sp = a.spec;
  const int tb = threadIdx.x / G, slot = threadIdx.x % G;
  const int env = blockIdx.x * TPB + tb;
  const uint32_t tmask = tile_mask(G);
  const bool env_live = env < sp.E && (env_mask == nullptr || env_mask[env < sp.E ? env : 0] != 0);
  const int e = env < sp.E ? env : sp.E - 1;
  const bool is_agent = slot < sp.n_agents;
  TileSmem<P, G>& ts = bs.tiles[tb];
  if (threadIdx.x < ENGINE_MAX_AGENTS) {
    bs.kind_tab[threadIdx.x] = sp.kind[threadIdx.x];
    bs.ip0_tab[threadIdx.x] = sp.agent_iparam[threadIdx.x][0];
  }
  __syncthreads();
  int4 h = a.hdr[e];
  int st[P::NWORDS > 0 ? P::NWORDS : 1];
#pragma unroll
  for (int w = 0; w < P::NWORDS; ++w) st[w] = a.state[((size_t)w * sp.E + e) * G + slot];
  h.x = 0;
  h.y += 1;
  h.z = sp.env_kind == PHX_ENV_FSM ? sp.initial_stage : 0;
  constexpr int EW = EnvWords<P>::value;
  int envw[EW > 0 ? EW : 1];  // env-level words survive a reset unless the program says otherwise
  if constexpr (EW > 0) {
#pragma unroll
    for (int w = 0; w < EW; ++w) envw[w] = a.env_state[(size_t)w * sp.E + e];
  }
  Ctx ctx;
  ctx.spec = &sp;
  ctx.slot = slot;
  ctx.kind = is_agent ? sp.kind[slot] : -1;
  ctx.env = envw;
  ctx.step = 0;
  ctx.stage = h.z;
  ctx.env_id = sp.env_offset + (uint32_t)e;
  ctx.episode = (uint32_t)h.y;
  ctx.views = &ts.views[0][0];
  ctx.view_stride = P::VW > 0 ? P::VW : 1;
  ctx.kind_tab = bs.kind_tab;
  ctx.ip0_tab = bs.ip0_tab;
  ctx.out_mask = is_agent ? sp.adj[slot] : 0u;
  ctx.in_mask = 0;
  if (a.adj_env) {
    // StochasticNetwork.reset resamples the edges before the agents reset (network.py:450-453);
    // the constructor's own sample (add_connection, :389-391) is the draw of "episode -1"
    ctx.out_mask = is_agent ? resample_adj_row(sp.seed, a.base_conn, a.n_base, ctx.env_id,
                                               agents_only ? 0xFFFFFFFFu : ctx.episode, slot)
                            : 0u;
    if (env_live) a.adj_env[(size_t)e * G + slot] = ctx.out_mask;
  }
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
      if (obs && got) {
#pragma unroll
        for (int j = 0; j < P::OBS_DIM; ++j)
          if (j < sp.obs_dim) obs[orow * sp.obs_dim + j] = obs_val[j];
      }
      if (obs_mask) obs_mask[orow] = got;
    }
    if (slot == 0) {
      a.hdr[e] = h;
      a.term[e] = 0;
      a.trunc[e] = 0;
      if (a.carry_n != nullptr) a.carry_n[e] = 0;  // Network.reset -> resolver.reset()
      if (sp.env_kind != PHX_ENV_BASE) a.reward_none[e] = sp.strategic_mask;  // _rewards = None
    }
#pragma unroll
    for (int w = 0; w < P::NWORDS; ++w) a.state[((size_t)w * sp.E + e) * G + slot] = st[w];
  }
}

template <class P, int G>
__global__ void __launch_bounds__(ENGINE_BLOCK)
engine_reset_kernel(const EngineArgs<P> a, const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                    bool agents_only) {
  engine_reset_body<P, G>(a, env_mask, obs, obs_mask, agents_only);
}

}  // namespace phx
