// phx_engine_host.cuh -- host side of the generic queue engine: a Family implementation that
// owns the state of E envs of a device program P and launches engine_step_kernel<P, G>.
#pragma once
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "phx_engine.cuh"
#include "phx_engine1.cuh"
#include "phx_family.h"

namespace phx {

inline uint32_t low_mask(const uint32_t* words) { return words[0]; }

inline void mask_from(uint32_t& d, const uint32_t* src) { d = src[0]; }
inline void mask_from(WMask& d, const uint32_t* src) {
  for (int k = 0; k < PHX_MASK_WORDS; ++k) d.w[k] = src[k];
}
inline void mask_set(uint32_t& d, int i) { d |= 1u << i; }
inline void mask_set(WMask& d, int i) { d.w[i >> 5] |= 1u << (i & 31); }

// Lowers the generic part of phx_spec into the engines' device form (EngineSpec / WideSpec).
template <class Spec>
inline int32_t make_engine_spec(const phx_spec& s, int32_t E, uint64_t seed, int64_t env_offset,
                                Spec* out, int nwords, int envwords) {
  constexpr int MAXA = (int)(sizeof(out->kind) / sizeof(out->kind[0]));
  PHX_REQUIRE(s.n_agents <= MAXA, PHX_ERR_UNSUPPORTED,
              MAXA <= ENGINE_MAX_AGENTS
                  ? "queue engine (tile variant) supports up to 32 agents per env"
                  : "queue engine (block variant) supports up to 128 agents per env");
  Spec& d = *out;
  std::memset(&d, 0, sizeof(d));
  d.E = E;
  d.n_agents = s.n_agents;
  d.n_strategic = s.n_strategic;
  d.num_steps = s.num_steps;
  d.round_limit = s.round_limit;
  d.env_kind = s.env_kind;
  d.flags = s.flags;
  d.obs_dim = s.obs_dim;
  d.act_dim = s.act_dim;
  for (int i = 0; i < s.n_agents; ++i) {
    d.kind[i] = (int8_t)s.agent_kind[i];
    d.sidx[i] = (int8_t)s.strategic_index[i];
    mask_from(d.adj[i], s.adjacency[i]);
    if (s.strategic_index[i] >= 0) mask_set(d.strategic_mask, i);
    if (s.agent_kind[i] >= 0 && s.agent_kind[i] < 8) mask_set(d.kind_mask[s.agent_kind[i]], i);
    for (int k = 0; k < 4; ++k) d.agent_iparam[i][k] = s.agent_iparam[i][k];
    for (int k = 0; k < 4; ++k) d.agent_fparam[i][k] = s.agent_fparam[i][k];
    for (int k = 0; k < PHX_MAX_CODEC_OPS; ++k) {
      d.codec_op[i][k] = s.agent_codec_op[i][k];
      d.codec_val[i][k] = s.agent_codec_val[i][k];
    }
  }
  for (int t = 0; t < s.n_payload_types; ++t) {
    mask_from(d.sender_ok[t], s.type_sender_ok[t]);
    mask_from(d.receiver_ok[t], s.type_receiver_ok[t]);
  }
  d.n_stages = s.n_stages;
  d.initial_stage = s.env_kind == PHX_ENV_FSM ? s.initial_stage : 0;
  for (int k = 0; k < s.n_stages && k < PHX_MAX_STAGES; ++k) {
    mask_from(d.stage_acting[k], s.stages[k].acting);
    mask_from(d.stage_rewarded[k], s.stages[k].rewarded);
    d.stage_rewarded_none[k] = (uint8_t)(s.stages[k].rewarded_is_none != 0);
    const phx_stage& g = s.stages[k];
    d.stage_allowed[k] = (uint8_t)(g.next_allowed & 0xFFu);
    if (g.handler == 0) {
      PHX_REQUIRE(g.next_stage >= 0 && g.next_stage < s.n_stages, PHX_ERR_INVALID,
                  "FSM stage has an invalid next_stage");
      d.stage_next[k] = (int8_t)g.next_stage;
      continue;
    }
    // a stage with an env handler: the declarative rule (include/phx.h phx_stage.rule_*)
    PHX_REQUIRE(g.handler == 1 || g.handler == 2, PHX_ERR_UNSUPPORTED,
                "unknown FSM stage handler kind");
    phx_rule_branch one{};  // handler == 1: the single comparison as a one-branch chain
    const phx_rule_branch* br = g.rule_branch;
    int nb = g.rule_n_branches;
    if (g.handler == 1) {
      one.n_terms = 1;
      one.then = g.rule_then;
      one.term[0].lhs = g.rule_lhs;
      one.term[0].slot = g.rule_slot;
      one.term[0].word = g.rule_word;
      one.term[0].cmp = g.rule_cmp;
      one.term[0].rhs_kind = PHX_RULE_CONST;
      one.term[0].rhs = g.rule_rhs;
      br = &one;
      nb = 1;
    }
    PHX_REQUIRE(nb >= 1 && nb <= PHX_RULE_BRANCHES, PHX_ERR_INVALID,
                "FSM stage rule: 1..PHX_RULE_BRANCHES branches");
    // a RETURNED stage may be outside next_stages (a run-time FSMRuntimeError in the reference,
    // fsm.py:304-307), but it must be a stage index the kernel can hold
    PHX_REQUIRE(g.rule_else >= 0 && g.rule_else < PHX_MAX_STAGES, PHX_ERR_INVALID,
                "FSM stage rule: stage index out of range");
    auto operand_ok = [&](int kind, int slot, int word, bool rhs) {
      if (kind == PHX_RULE_STEP) return true;
      if (kind == PHX_RULE_AGENT_WORD)
        return slot >= 0 && slot < s.n_agents && word >= 0 && word < nwords;
      if (kind == PHX_RULE_ENV_WORD) return word >= 0 && word < envwords;
      return rhs && kind == PHX_RULE_CONST;
    };
    d.stage_next[k] = (int8_t)k;
    d.stage_rule[k][SR_HANDLER] = 1;
    d.stage_rule[k][SR_RESOLVES] = (int8_t)(g.rule_resolves != 0);
    d.stage_rule[k][SR_BRANCHES] = (int8_t)nb;
    d.stage_rule[k][SR_ELSE] = (int8_t)g.rule_else;
    for (int b = 0; b < nb; ++b) {
      PHX_REQUIRE(br[b].n_terms >= 1 && br[b].n_terms <= PHX_RULE_TERMS, PHX_ERR_INVALID,
                  "FSM stage rule: 1..PHX_RULE_TERMS comparisons per branch");
      PHX_REQUIRE(br[b].then >= 0 && br[b].then < PHX_MAX_STAGES, PHX_ERR_INVALID,
                  "FSM stage rule: stage index out of range");
      d.rule_branch[k][b][0] = (int8_t)br[b].n_terms;
      d.rule_branch[k][b][1] = (int8_t)br[b].then;
      for (int j = 0; j < br[b].n_terms; ++j) {
        const phx_rule_term& t = br[b].term[j];
        PHX_REQUIRE((t.cmp & ~PHX_CMP_F32) >= PHX_CMP_LT && (t.cmp & ~PHX_CMP_F32) <= PHX_CMP_GT, PHX_ERR_INVALID,
                    "FSM stage rule: bad lhs / cmp");
        // `always` only as the single unconditional rule
        const bool always = t.lhs == PHX_RULE_ALWAYS && br[b].n_terms == 1;
        PHX_REQUIRE(always || (t.lhs != PHX_RULE_ALWAYS && operand_ok(t.lhs, t.slot, t.word, false)),
                    PHX_ERR_INVALID,
                    t.lhs == PHX_RULE_AGENT_WORD ? "FSM stage rule: no such agent state word"
                    : t.lhs == PHX_RULE_ENV_WORD ? "FSM stage rule: no such env-level word"
                                                 : "FSM stage rule: bad lhs / cmp");
        PHX_REQUIRE(always || operand_ok(t.rhs_kind, t.rhs_slot, t.rhs_word, true), PHX_ERR_INVALID,
                    "FSM stage rule: bad right-hand operand");
        int8_t* o = d.rule_term[k][b][j];
        o[RT_LHS] = (int8_t)t.lhs;
        o[RT_SLOT] = (int8_t)t.slot;
        o[RT_WORD] = (int8_t)t.word;
        o[RT_CMP] = (int8_t)t.cmp;
        o[RT_RHS] = (int8_t)t.rhs_kind;
        o[RT_RHS_SLOT] = (int8_t)t.rhs_slot;
        o[RT_RHS_WORD] = (int8_t)t.rhs_word;
        d.rule_rhs[k][b][j] = t.rhs;
      }
    }
  }
  mask_from(d.leaders, s.leaders);
  mask_from(d.followers, s.followers);
  // acting order per phase (include/phx.h phx_stage.n_act_order)
  const int n_phases = s.env_kind == PHX_ENV_FSM ? s.n_stages : s.env_kind == PHX_ENV_STACKELBERG ? 2 : 1;
  for (int ph = 0; ph < n_phases && ph < PHX_MAX_STAGES; ++ph) {
    const phx_stage& g = s.stages[ph];
    const int na = s.env_kind == PHX_ENV_BASE ? 0 : g.n_act_order;
    PHX_REQUIRE(na >= 0 && na <= s.n_agents, PHX_ERR_INVALID, "n_act_order out of range");
    if (na == 0) {  // slot order
      d.n_act[ph] = (uint8_t)s.n_agents;
      for (int i = 0; i < s.n_agents; ++i) d.act_order[ph][i] = (int8_t)i;
      continue;
    }
    const uint32_t* acting = s.env_kind == PHX_ENV_FSM ? g.acting : (ph == 0 ? s.leaders : s.followers);
    uint32_t seen[PHX_MASK_WORDS] = {0, 0, 0, 0};
    for (int i = 0; i < na; ++i) {
      const int o = g.act_order[i];
      PHX_REQUIRE(o < s.n_agents && mask_bit(acting, o) && !mask_bit(seen, o), PHX_ERR_INVALID,
                  "act_order: every entry must be a distinct acting agent of the stage");
      seen[o >> 5] |= 1u << (o & 31);
      d.act_order[ph][i] = (int8_t)o;
    }
    int n_acting = 0;
    for (int i = 0; i < s.n_agents; ++i) n_acting += mask_bit(acting, i);
    PHX_REQUIRE(n_acting == na, PHX_ERR_INVALID, "act_order must list every acting agent of the stage");
    d.n_act[ph] = (uint8_t)na;
    d.any_act_order = 1;
  }
  d.seed = seed;
  d.env_offset = (uint32_t)env_offset;
  for (int k = 0; k < PHX_MAX_PARAMS; ++k) {
    d.iparams[k] = s.iparams[k];
    d.fparams[k] = (float)s.fparams[k];
  }
  for (int k = 0; k < 4; ++k) d.dparams[k] = s.fparams[k];
  return PHX_OK;
}

// ---------------------------------------------------------------- run-time specialisation
// The generic engines interpret the lowered env class (EngineSpec) at run time: every agent
// loop, mask test and per-slot table lookup is data driven, which costs an order of magnitude
// in instructions for small env classes (C4: ~1 700 per env-step).  A specialised build bakes
// the EngineSpec of ONE handle into the translation unit as a `__device__ constexpr` object;
// the same kernel body (engine1_step_body / engine_step_body) then sees compile-time agent counts, kinds, masks
// and stage tables, and the compiler unrolls and folds them.  libphx only produces the source
// text and loads the cubin; the host binding runs nvcc (phantom_b200/jit.py) and caches the
// result, so there is no compiler inside the library and no link-time dependency on the driver.
template <class P, class = void>
struct HasJit : std::false_type {};
template <class P>
struct HasJit<P, std::void_t<decltype(P::JIT_SOURCE)>> : std::true_type {};

template <class T>
inline void jit_emit(std::string& o, const T& v) {
  if constexpr (std::is_floating_point<T>::value) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), std::is_same<T, float>::value ? "%.9gf" : "%.17g", (double)v);
    std::string t(buf);
    // "1f" / "1" are not floating literals: make sure there is a '.' or an exponent
    const bool is_f = std::is_same<T, float>::value;
    std::string body = is_f ? t.substr(0, t.size() - 1) : t;
    if (body.find_first_of(".eEn") == std::string::npos) body += ".0";  // (n: nan / inf)
    if (body.find("inf") != std::string::npos || body.find("nan") != std::string::npos)
      body = "0.0";  // parameters are validated finite; never reached
    o += body + (is_f ? "f" : "");
  } else if constexpr (std::is_unsigned<T>::value) {
    o += std::to_string((unsigned long long)v) + (sizeof(T) == 8 ? "ull" : "u");
  } else {
    o += std::to_string((long long)v);
  }
}
template <class T, size_t N>
inline void jit_emit(std::string& o, const T (&a)[N]) {
  o += "{";
  for (size_t i = 0; i < N; ++i) {
    if (i) o += ", ";
    jit_emit(o, a[i]);
  }
  o += "}";
}
// EngineSpec as a C++ aggregate initialiser, fields in declaration order (phx_engine.cuh).
inline std::string jit_spec_literal(const EngineSpec& d) {
  std::string o = "{\n  ";
  auto f = [&](const auto& v) {
    jit_emit(o, v);
    o += ",\n  ";
  };
  f(d.E); f(d.n_agents); f(d.n_strategic); f(d.num_steps); f(d.round_limit); f(d.env_kind);
  f(d.flags); f(d.obs_dim); f(d.act_dim); f(d.kind); f(d.sidx); f(d.adj); f(d.sender_ok);
  f(d.receiver_ok); f(d.strategic_mask); f(d.kind_mask); f(d.n_stages); f(d.initial_stage);
  f(d.stage_acting); f(d.stage_rewarded); f(d.stage_rewarded_none); f(d.stage_next);
  f(d.leaders); f(d.followers); f(d.seed); f(d.env_offset); f(d.iparams); f(d.fparams);
  f(d.dparams); f(d.agent_iparam); f(d.agent_fparam); f(d.codec_op); f(d.codec_val);
  f(d.stage_rule); f(d.rule_branch); f(d.rule_term); f(d.rule_rhs); f(d.stage_allowed);
  f(d.any_act_order); f(d.n_act); f(d.act_order);
  o += "}";
  return o;
}

// ------------------------------------------------------------------ static message schedule
// A program opts in with two HOST functions that list its POTENTIAL sends, in emission order:
//     static void act_sends(const phx_spec&, int slot, int phase, std::vector<SendSig>& out);
//     static void handle_sends(const phx_spec&, int slot, int type, int sender,
//                              std::vector<SendSig>& out);
// `phase` = 0 under PhantomEnv, the stage index under the FSM env, 0 / 1 = leaders' / followers'
// turn under the Stackelberg env.  An actual send must be one of the listed ones, in that order
// (a subsequence); the specialised kernel checks it (PHX_FAULT_PLAN_MISMATCH).
struct SendSig {
  int recv, type;
};
template <class P, class = void>
struct HasStaticSig : std::false_type {};
template <class P>
struct HasStaticSig<P, std::void_t<decltype(&P::act_sends), decltype(&P::handle_sends)>>
    : std::true_type {};

// Network.send's checks on the static graph (network.py:246-254): 0 = pushed.
inline uint32_t plan_send_check(const phx_spec& s, int from, int to, int type) {
  if (to < 0 || to >= s.n_agents || type < 0 || type >= s.n_payload_types) return PHX_FAULT_NO_EDGE;
  if (!(s.flags & PHX_FLAG_IGNORE_CONNECTION_ERRORS) && !mask_bit(s.adjacency[from], to))
    return PHX_FAULT_NO_EDGE;
  if (!(s.flags & PHX_FLAG_NO_PAYLOAD_CHECKS) &&
      (!mask_bit(s.type_sender_ok[type], from) || !mask_bit(s.type_receiver_ok[type], to)))
    return PHX_FAULT_BAD_PAYLOAD_TYPE;
  return 0;
}

// Walks SURVEY.md A.1 rules 2-7 over the potential sends of every phase.  Returns false (and
// says why) if the env class has no static schedule under the rules stated at StaticPlan.
template <class P>
bool build_static_plan(const phx_spec& s, StaticPlan* out, std::string* why) {
  auto fail = [&](const std::string& m) {
    if (why) *why = m;
    return false;
  };
  if (s.n_agents > SPL_AGENTS) return fail("more than 8 agents");
  if (s.flags & (PHX_FLAG_STOCHASTIC_NETWORK | PHX_FLAG_SHUFFLE_BATCHES))
    return fail("per-env graphs / shuffled batches are data dependent");
  if (s.env_kind != PHX_ENV_BASE)
    for (int ph = 0; ph < PHX_MAX_STAGES; ++ph)
      if (s.stages[ph].n_act_order > 0) return fail("a stage's acting order differs from slot order");
  if (s.env_kind == PHX_ENV_FSM)
    for (int ph = 0; ph < s.n_stages; ++ph)
      if (s.stages[ph].handler != 0 && s.stages[ph].rule_resolves == 0)
        return fail("a stage handler does not resolve: mail may wait across steps");
  StaticPlan& pl = *out;
  std::memset(&pl, 0, sizeof(pl));
  pl.n_phases = s.env_kind == PHX_ENV_FSM ? s.n_stages : s.env_kind == PHX_ENV_STACKELBERG ? 2 : 1;
  if (pl.n_phases > SPL_PHASES) return fail("too many phases");
  struct M {
    int sender, recv, type, src, k;  // src: message of the previous round (-1: acting), k-th send
  };
  for (int ph = 0; ph < pl.n_phases; ++ph) {
    uint32_t acting = 0xFFFFFFFFu;
    if (s.env_kind == PHX_ENV_FSM) acting = s.stages[ph].acting[0];
    if (s.env_kind == PHX_ENV_STACKELBERG) acting = ph == 0 ? s.leaders[0] : s.followers[0];
    // round 0 in push order: acting agents in slot order, their sends in emission order
    std::vector<M> cur;
    for (int a = 0; a < s.n_agents; ++a) {
      if (!((acting >> a) & 1u)) continue;
      std::vector<SendSig> sg;
      P::act_sends(s, a, ph, sg);
      if ((int)sg.size() > SPL_MSGS) return fail("an agent may send more than 16 messages");
      for (int k = 0; k < (int)sg.size(); ++k) {
        if (plan_send_check(s, a, sg[k].recv, sg[k].type)) return fail("a potential send faults");
        cur.push_back(M{a, sg[k].recv, sg[k].type, a, k});
      }
    }
    for (int r = 0; !cur.empty(); ++r) {
      if (r >= SPL_ROUNDS) return fail("more than 4 resolver rounds");
      if ((int)cur.size() > SPL_MSGS) return fail("more than 16 messages in a round");
      // receivers in first-arrival order, batches in push order (resolvers.py:126,142)
      std::vector<int> order;
      for (const M& m : cur)
        if (std::find(order.begin(), order.end(), m.recv) == order.end()) order.push_back(m.recv);
      if (r > 0)
        for (int rc : order) {
          int snd = -1;
          for (const M& m : cur)
            if (m.recv == rc) {
              if (snd >= 0 && snd != m.sender) return fail("a response batch has two senders");
              snd = m.sender;
            }
        }
      std::vector<M> stored;  // grouped by receiver
      std::vector<int> slot_of(cur.size(), -1);
      for (int rc : order)
        for (size_t i = 0; i < cur.size(); ++i)
          if (cur[i].recv == rc) {
            slot_of[i] = (int)stored.size();
            stored.push_back(cur[i]);
          }
      pl.n_rounds[ph] = (int8_t)(r + 1);
      pl.n_msg[ph][r] = (int8_t)stored.size();
      for (size_t i = 0; i < stored.size(); ++i) {
        pl.sender[ph][r][i] = (int8_t)stored[i].sender;
        pl.recv[ph][r][i] = (int8_t)stored[i].recv;
        pl.type[ph][r][i] = (int8_t)stored[i].type;
      }
      // where each producer's sends landed
      for (size_t i = 0; i < cur.size(); ++i) {
        const M& m = cur[i];
        if (r == 0) {
          pl.act_slot[ph][m.src][m.k] = (int8_t)slot_of[i];
          pl.act_n[ph][m.src] = (int8_t)std::max<int>(pl.act_n[ph][m.src], m.k + 1);
        } else {
          if (m.k >= SPL_RESP) return fail("a handler may answer with more than 4 messages");
          pl.resp_slot[ph][r - 1][m.src][m.k] = (int8_t)slot_of[i];
          pl.resp_n[ph][r - 1][m.src] = (int8_t)std::max<int>(pl.resp_n[ph][r - 1][m.src], m.k + 1);
        }
      }
      // responses: receivers in order, each handling its batch in push order (agents.py:110-120)
      std::vector<M> next;
      for (size_t i = 0; i < stored.size(); ++i) {
        const M& m = stored[i];
        // delivery-time edge filter (resolvers.py:146-148): only with ignore_connection_errors
        if (!mask_bit(s.adjacency[m.sender], m.recv)) continue;
        std::vector<SendSig> sg;
        P::handle_sends(s, m.recv, m.type, m.sender, sg);
        for (int k = 0; k < (int)sg.size(); ++k) {
          if (plan_send_check(s, m.recv, sg[k].recv, sg[k].type))
            return fail("a potential response faults");
          next.push_back(M{m.recv, sg[k].recv, sg[k].type, (int)i, k});
        }
      }
      cur.swap(next);
    }
  }
  return true;
}

inline std::string jit_plan_literal(const StaticPlan& p) {
  std::string o = "{\n  ";
  auto f = [&](const auto& v) {
    jit_emit(o, v);
    o += ",\n  ";
  };
  f(p.n_phases); f(p.n_rounds); f(p.n_msg); f(p.sender); f(p.recv); f(p.type); f(p.resp_n);
  f(p.resp_slot); f(p.act_n); f(p.act_slot);
  o += "}";
  return o;
}

template <class P>
__global__ void engine_init_kernel(int E, int G, int4* hdr, int32_t* state, int nwords) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E) hdr[i] = make_int4(0, -1, 0, 0);  // episode becomes 0 on the first reset
  (void)state; (void)nwords; (void)G;
}

// Can mail wait for a later step's resolve?  Only under the FSM env, in a stage whose handler
// does not call resolve_network() (fsm.py:280-283).
inline bool waiting_mail_possible(const phx_spec& s) {
  if (s.env_kind != PHX_ENV_FSM) return false;
  for (int k = 0; k < s.n_stages && k < PHX_MAX_STAGES; ++k)
    if (s.stages[k].handler != 0 && s.stages[k].rule_resolves == 0) return true;
  return false;
}

// Family backed by the queue engine.  `Field` mapping: PHX_FIELD_FAMILY + w = state word w,
// int32 [E, G] (slot-major inside an env).
template <class P>
class EngineFamily : public Family {
 public:
  ~EngineFamily() override {
    cudaFree(d_state);
    cudaFree(d_rcache);
    cudaFree(d_rnone);
    cudaFree(d_ocache);
    cudaFree(d_ocached);
    cudaFree(d_adj);
    cudaFree(d_base);
    cudaFree(d_env);
    cudaFree(d_carry);
    cudaFree(d_carry_n);
    if (jit_lib) cudaLibraryUnload(jit_lib);
  }

  // Host-only part of init(): lowers the spec and picks the tiling (no CUDA call), so that the
  // specialised source of an env class can also be produced without a device (self-test hook).
  int32_t layout(const phx_spec& s) {
    int32_t rc = make_engine_spec(s, E, seed, env_offset, &espec, P::NWORDS, EnvWords<P>::value);
    if (rc != PHX_OK) return rc;
    rc = P::validate(s);
    if (rc != PHX_OK) return rc;
    G = s.n_agents <= 8 ? 8 : (s.n_agents <= 16 ? 16 : 32);
    if constexpr (HasActTotal<P>::value) {  // compact acting queue: the capacities must fit
      PHX_REQUIRE(!(s.flags & PHX_FLAG_STOCHASTIC_NETWORK), PHX_ERR_UNSUPPORTED,
                  "this family's compact acting queue is sized from a fixed graph");
      int total = 0;
      for (int i = 0; i < s.n_agents; ++i) {
        int deg = 0;
        for (int r = 0; r < s.n_agents; ++r) deg += mask_bit(s.adjacency[i], r);
        total += std::min((int)P::ACTCAP, P::act_cap(s.agent_kind[i], deg));
      }
      PHX_REQUIRE(total <= P::ACTTOTAL, PHX_ERR_UNSUPPORTED,
                  "acting-phase fan-out of this env class exceeds the family's queue");
    }
    if constexpr (HasRespTotal<P>::value) {  // compact response queues, sized from the full graph
      int total = 0;
      for (int i = 0; i < s.n_agents; ++i) {
        int deg = 0;
        for (int r = 0; r < s.n_agents; ++r) deg += mask_bit(s.adjacency[i], r);
        total += std::min((int)P::RESPCAP, P::resp_cap(s.agent_kind[i], deg));
      }
      PHX_REQUIRE(total <= P::RESPTOTAL, PHX_ERR_UNSUPPORTED,
                  "response fan-out of this env class exceeds the family's queue");
    }
    // thread-per-env variant: <= 8 agents and a program that declares its queue bound
    qcap1 = P::q1_cap(s);  // messages in flight per round, for THIS env class (0 = unsupported)
    // (shuffle_batches needs the per-receiver batch lists of the tile engine)
    const bool eligible = P::Q1CAP > 0 && P::VW <= 1 && s.n_agents <= ENGINE1_SLOTS && qcap1 > 0 &&
                          qcap1 <= P::Q1CAP && !(s.flags & PHX_FLAG_SHUFFLE_BATCHES);
    PHX_REQUIRE(s.exec_mode != PHX_EXEC_THREAD || eligible, PHX_ERR_UNSUPPORTED,
                "PHX_EXEC_THREAD needs an env class with at most 8 agents whose device program "
                "supports the thread-per-env engine");
    thread_per_env = s.exec_mode == PHX_EXEC_THREAD || (s.exec_mode == PHX_EXEC_AUTO && eligible);
    if constexpr (HasCollective<P>::value) {  // the program's collective resolve (tile engine)
      const char* off = std::getenv("PHX_COLLECTIVE");
      collective_ok_ = !(off && off[0] == '0') && P::collective_ok(s);
      if (s.env_kind != PHX_ENV_BASE)  // (the collective forms assume agents act in slot order)
        for (int ph = 0; ph < PHX_MAX_STAGES; ++ph)
          if (s.stages[ph].n_act_order > 0) collective_ok_ = false;
    }
    PHX_REQUIRE(s.obs_dim <= P::OBS_DIM, PHX_ERR_INVALID, "obs_dim exceeds the family's OBS_DIM");
    return PHX_OK;
  }

  int32_t jit_source_offline(std::string& out) override {
    const int32_t rc = layout(spec);
    return rc != PHX_OK ? rc : jit_source(out);
  }

  int32_t init(const phx_spec& s) override {
    int32_t rc = layout(s);
    if (rc != PHX_OK) return rc;
    const size_t n = (size_t)E * G;
    PHX_CUDA(cudaMalloc(&d_state, sizeof(int32_t) * n * (P::NWORDS > 0 ? P::NWORDS : 1)));
    PHX_CUDA(cudaMemset(d_state, 0, sizeof(int32_t) * n * (P::NWORDS > 0 ? P::NWORDS : 1)));
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
    if (EnvWords<P>::value > 0) {  // env-level words start at zero (e.g. avg_price = 0.0)
      PHX_CUDA(cudaMalloc(&d_env, sizeof(int32_t) * (size_t)E * EnvWords<P>::value));
      PHX_CUDA(cudaMemset(d_env, 0, sizeof(int32_t) * (size_t)E * EnvWords<P>::value));
    }
    if (s.flags & PHX_FLAG_STOCHASTIC_NETWORK) {
      // per-env adjacency rows + the base-connection table with integer thresholds
      PHX_REQUIRE(s.n_base_connections >= 0 && s.n_base_connections <= PHX_MAX_BASE_CONNECTIONS,
                  PHX_ERR_INVALID, "n_base_connections out of range");
      std::vector<uint2> base((size_t)s.n_base_connections + 1);
      for (int c = 0; c < s.n_base_connections; ++c) {
        PHX_REQUIRE(s.base_u[c] < s.n_agents && s.base_v[c] < s.n_agents, PHX_ERR_INVALID,
                    "base connection names an agent slot outside the env");
        const double r = s.base_rate[c];
        PHX_REQUIRE(r == r, PHX_ERR_INVALID, "base connection rate is NaN");
        // uniform01 < r  <=>  d24 < ceil(r * 2^24); r * 2^24 is exact in float64
        const double scaled = std::ceil(std::min(std::max(r, 0.0), 1.0) * 16777216.0);
        base[c] = make_uint2((uint32_t)s.base_u[c] | ((uint32_t)s.base_v[c] << 8), (uint32_t)scaled);
      }
      n_base = s.n_base_connections;
      PHX_CUDA(cudaMalloc(&d_base, sizeof(uint2) * base.size()));
      PHX_CUDA(cudaMemcpy(d_base, base.data(), sizeof(uint2) * base.size(), cudaMemcpyHostToDevice));
      PHX_CUDA(cudaMalloc(&d_adj, sizeof(uint32_t) * n));
      PHX_CUDA(cudaMemset(d_adj, 0, sizeof(uint32_t) * n));
    }
    // mail that waits across steps (a stage handler that does not resolve, fsm.py:280-283): kept
    // by the thread-per-env engine; the tile engines fault with PHX_FAULT_UNRESOLVED_MAIL
    if (thread_per_env && waiting_mail_possible(s)) {
      PHX_CUDA(cudaMalloc(&d_carry_n, sizeof(int32_t) * (size_t)E));
      PHX_CUDA(cudaMemset(d_carry_n, 0, sizeof(int32_t) * (size_t)E));
      PHX_CUDA(cudaMalloc(&d_carry, sizeof(int32_t) * (size_t)E * qcap1 * (1 + P::PW)));
    }
    engine_init_kernel<P><<<(E + 255) / 256, 256>>>(E, G, d_hdr, d_state, P::NWORDS);
    PHX_CUDA(cudaGetLastError());
    // constructor-time agent state (PhantomEnv.__init__ ends with agent.reset(), env.py:122-124)
    rc = launch_reset(nullptr, nullptr, nullptr, 0, /*count_episode=*/false);
    if (rc != PHX_OK) return rc;
    PHX_CUDA(cudaDeviceSynchronize());
    name = thread_per_env ? std::string("thread-per-env(G=8)")
                          : std::string("queue(G=") + std::to_string(G) +
                                (collective_ok_ && !tracking() ? ", collective)" : ")");
    return PHX_OK;
  }

  EngineArgs<P> make_args(int32_t T, const StepIO& io) const {
    EngineArgs<P> a;
    a.spec = espec;
    a.T = T;
    a.qcap = qcap1;
    a.collective = collective_ok_ && !tracking() && !thread_per_env;
    {  // output staging of the thread-per-env engine: see engine1_step_body
      const size_t S = spec.n_strategic > 0 ? spec.n_strategic : 1;
      auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
      // Default: specialised builds only.  Measured (C4, 131 072 envs x 100 steps): specialised
      // 0.594 -> 0.424 ms; the generic build is instruction bound and loses to the two block
      // barriers a step (0.878 -> 0.990 ms).  PHX_ENGINE1_STAGE = 0 | 1 forces it off / on.
      const char* ov = std::getenv("PHX_ENGINE1_STAGE");
      // (one strategic agent: the rows of consecutive envs are already close together)
      const bool want = ov ? ov[0] == '1' : (jit_kernel != nullptr && !tracking() && S >= 2);
      a.stage_out = want && ((size_t)E * S) % 16 == 0 && E % 8 == 0 &&
                    al(io.obs) && al(io.obs_mask) && al(io.reward) && al(io.reward_mask) &&
                    al(io.term) && al(io.trunc) && al(io.all_done);
    }
    a.hdr = d_hdr;
    a.term = d_term;
    a.trunc = d_trunc;
    a.state = d_state;
    a.reward_cache = d_rcache;
    a.reward_none = d_rnone;
    a.obs_cache = d_ocache;
    a.obs_cached = d_ocached;
    a.env_state = d_env;
    a.adj_env = d_adj;
    a.base_conn = d_base;
    a.n_base = n_base;
    a.io = io;
    a.faults = fault_sink();
    a.trace = trace_sink();
    a.carry_n = d_carry_n;
    a.carry = d_carry;
    return a;
  }

  template <int GG>
  int32_t launch_step(const EngineArgs<P>& a, cudaStream_t stream) {
    constexpr int TPB = ENGINE_BLOCK / GG;
    const size_t smem = sizeof(BlockSmem<P, GG>);
    const int grid = (E + TPB - 1) / TPB;
    if (jit_kernel && !tracking()) {  // the build specialised for this handle's env class
      void* args[] = {(void*)&a};
      PHX_CUDA(cudaLaunchKernel((const void*)jit_kernel, dim3(grid), dim3(ENGINE_BLOCK), args, smem,
                                stream));
      return PHX_OK;
    }
    if (tracking()) {
      PHX_CUDA(cudaFuncSetAttribute(engine_step_kernel<P, GG, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      engine_step_kernel<P, GG, true><<<grid, ENGINE_BLOCK, smem, stream>>>(a);
    } else {
      PHX_CUDA(cudaFuncSetAttribute(engine_step_kernel<P, GG, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      engine_step_kernel<P, GG, false><<<grid, ENGINE_BLOCK, smem, stream>>>(a);
    }
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  template <int GG>
  int32_t launch_reset_g(const EngineArgs<P>& a, const uint8_t* env_mask, float* obs,
                         uint8_t* obs_mask, cudaStream_t stream, bool agents_only) {
    constexpr int TPB = ENGINE_BLOCK / GG;
    const size_t smem = sizeof(BlockSmem<P, GG>);
    PHX_CUDA(cudaFuncSetAttribute(engine_reset_kernel<P, GG>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    engine_reset_kernel<P, GG><<<(E + TPB - 1) / TPB, ENGINE_BLOCK, smem, stream>>>(
        a, env_mask, obs, obs_mask, agents_only);
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  int32_t launch_reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask, cudaStream_t stream,
                       bool count_episode) {
    StepIO io{};
    EngineArgs<P> a = make_args(1, io);
    const bool ao = !count_episode;
    return G == 8 ? launch_reset_g<8>(a, env_mask, obs, obs_mask, stream, ao)
           : G == 16 ? launch_reset_g<16>(a, env_mask, obs, obs_mask, stream, ao)
                     : launch_reset_g<32>(a, env_mask, obs, obs_mask, stream, ao);
  }

  int32_t reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                cudaStream_t stream) override {
    return launch_reset(env_mask, obs, obs_mask, stream, true);
  }

  int32_t rollout(int32_t T, const StepIO& io, cudaStream_t stream) override {
    if (tracking()) {
      const int32_t rc = ensure_trace(T);
      if (rc != PHX_OK) return rc;
    }
    EngineArgs<P> a = make_args(T, io);
    if constexpr (P::Q1CAP > 0 && P::VW <= 1) {
      if (thread_per_env) {
        const Engine1Layout lay = engine1_layout<P>(spec.n_agents, spec.n_strategic, qcap1,
                                                    spec.env_kind != PHX_ENV_BASE, spec.obs_dim,
                                                    a.stage_out != 0);
        const size_t smem = sizeof(int32_t) * (size_t)lay.words;
        const int grid = (E + ENGINE1_BLOCK - 1) / ENGINE1_BLOCK;
        if (jit_kernel && !tracking()) {  // the build specialised for this handle's env class
          void* args[] = {(void*)&a};
          PHX_CUDA(cudaLaunchKernel((const void*)jit_kernel, dim3(grid), dim3(ENGINE1_BLOCK), args,
                                    smem, stream));
          return PHX_OK;
        }
        if (tracking()) {
          PHX_CUDA(cudaFuncSetAttribute(engine1_step_kernel<P, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          engine1_step_kernel<P, true><<<grid, ENGINE1_BLOCK, smem, stream>>>(a);
        } else {
          PHX_CUDA(cudaFuncSetAttribute(engine1_step_kernel<P, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          engine1_step_kernel<P, false><<<grid, ENGINE1_BLOCK, smem, stream>>>(a);
        }
        PHX_CUDA(cudaGetLastError());
        return PHX_OK;
      }
    }
    return G == 8 ? launch_step<8>(a, stream)
           : G == 16 ? launch_step<16>(a, stream)
                     : launch_step<32>(a, stream);
  }

  template <int GG>
  int generic_blocks_per_sm() {
    int nb = 0;
    const size_t smem = sizeof(BlockSmem<P, GG>);
    if (cudaFuncSetAttribute(engine_step_kernel<P, GG, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, engine_step_kernel<P, GG, false>,
                                                      ENGINE_BLOCK, smem) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return nb;
  }

  int32_t jit_source(std::string& out) override {
    if constexpr (HasJit<P>::value) {
      const bool thread_ok = P::Q1CAP > 0 && P::VW <= 1;
      PHX_REQUIRE(!thread_per_env || thread_ok, PHX_ERR_UNSUPPORTED, "no specialisation");
      const std::string prog = P::JIT_NAME;
      const std::string body =
          thread_per_env ? "engine1_step_body<" + prog + ", false, ConstSpec>(a);"
                         : "engine_step_body<" + prog + ", " + std::to_string(G) + ", false, ConstSpec>(a);";
      std::string bound = thread_per_env ? "ENGINE1_BLOCK" : "ENGINE_BLOCK";
      if (!thread_per_env) {
        // With the agent loops unrolled over a compile-time agent count the tile kernel can need
        // far more registers than the generic build (simple_market on 32-lane tiles: 173 vs 119,
        // 2 instead of 4 resident blocks per SM).  PHX_JIT_MINBLOCKS = k | auto (the generic
        // kernel's residency) adds a min-blocks launch bound; measured (profiles/
        // r01_ab_jit_minblocks.txt) it helps that env class (15.4 -> 13.0 ms, generic 9.6) but
        // costs C4 on 8-lane tiles (3.5 -> 4.2 ms), so the default stays unbounded.
        int nb = 0;
        if (const char* ov = std::getenv("PHX_JIT_MINBLOCKS")) {
          if (std::string(ov) == "auto")
            nb = G == 8 ? generic_blocks_per_sm<8>()
                 : G == 16 ? generic_blocks_per_sm<16>() : generic_blocks_per_sm<32>();
          else
            nb = std::atoi(ov);
        }
        if (nb > 0) bound += ", " + std::to_string(nb);
      }
      // static message schedule (thread-per-env engine, programs that declare their sends)
      std::string plan_text, plan_members;
      static_plan = false;
      if constexpr (HasStaticSig<P>::value) {
        StaticPlan pl;
        std::string why;
        const char* off = std::getenv("PHX_JIT_STATIC_PLAN");
        if (thread_per_env && !P::BATCHED && !(off && off[0] == '0') &&
            build_static_plan<P>(spec, &pl, &why)) {
          static_plan = true;
          plan_text = "__device__ constexpr StaticPlan kPlan = " + jit_plan_literal(pl) + ";\n";
          plan_members =
              "  static constexpr int N_PHASES = " + std::to_string(pl.n_phases) + ";\n"
              "  __device__ __forceinline__ static const StaticPlan& plan() { return kPlan; }\n";
        } else {
          plan_text = "// no static message schedule: " + why + "\n";
        }
      }
      out = std::string("// generated by libphx (phx_jit_source): the ") +
            (thread_per_env ? "thread-per-env" : "tile") + " step kernel of\n// " + prog +
            " with this handle's lowered env class as a compile-time constant\n"
            "#define PHX_JIT_TU 1\n#include \"" + P::JIT_SOURCE + "\"\n"
            "namespace phx {\nnamespace {\n__device__ constexpr EngineSpec kSpec = " +
            jit_spec_literal(espec) + ";\n" + plan_text +
            "struct ConstSpec {\n  template <class A>\n"
            "  __device__ __forceinline__ static const EngineSpec& get(const A&) { return kSpec; }\n" +
            plan_members + "};\n"
            "}  // namespace\n"
            "extern \"C\" __global__ void __launch_bounds__(" + bound + ")\n"
            "phx_jit_step(const EngineArgs<" + prog + "> a) {\n  " + body + "\n}\n}  // namespace phx\n";
      return PHX_OK;
    } else {
      return Family::jit_source(out);
    }
  }

  size_t step_smem_bytes() const {
    if constexpr (P::Q1CAP > 0 && P::VW <= 1) {
      if (thread_per_env) {
        const Engine1Layout lay = engine1_layout<P>(spec.n_agents, spec.n_strategic, qcap1,
                                                    spec.env_kind != PHX_ENV_BASE, spec.obs_dim,
                                                    /*with_stage=*/true);
        return sizeof(int32_t) * (size_t)lay.words;
      }
    }
    return G == 8 ? sizeof(BlockSmem<P, 8>) : G == 16 ? sizeof(BlockSmem<P, 16>) : sizeof(BlockSmem<P, 32>);
  }

  int32_t load_specialised(const char* cubin_path) override {
    if constexpr (HasJit<P>::value) {
      PHX_REQUIRE(cubin_path != nullptr, PHX_ERR_INVALID, "cubin path is NULL");
      cudaLibrary_t lib = nullptr;
      PHX_CUDA(cudaLibraryLoadFromFile(&lib, cubin_path, nullptr, nullptr, 0, nullptr, nullptr, 0));
      cudaKernel_t k = nullptr;
      cudaError_t err = cudaLibraryGetKernel(&k, lib, "phx_jit_step");
      if (err == cudaSuccess)
        err = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)step_smem_bytes());
      if (err != cudaSuccess) {
        cudaLibraryUnload(lib);
        PHX_CUDA(err);
      }
      if (jit_lib) cudaLibraryUnload(jit_lib);
      jit_lib = lib;
      jit_kernel = k;
      name = thread_per_env ? std::string(static_plan ? "thread-per-env(G=8, specialised, static schedule)"
                                                      : "thread-per-env(G=8, specialised)")
                            : std::string("queue(G=") + std::to_string(G) +
                                  (collective_ok_ && !tracking() ? ", collective, specialised)"
                                                                 : ", specialised)");
      return PHX_OK;
    } else {
      return Family::load_specialised(cubin_path);
    }
  }

  int32_t family_field(int32_t field, int32_t index, void** p, size_t* bytes) override {
    if (field == PHX_FIELD_ADJACENCY) {  // uint32 [E, G]: G >= n_agents slots per env
      PHX_REQUIRE(d_adj != nullptr, PHX_ERR_INVALID,
                  "PHX_FIELD_ADJACENCY needs PHX_FLAG_STOCHASTIC_NETWORK");
      *p = d_adj;
      *bytes = sizeof(uint32_t) * (size_t)E * G;
      return PHX_OK;
    }
    if (field == PHX_FIELD_ENV_STATE) {  // int32 [E]: env-level word `index`
      PHX_REQUIRE(index >= 0 && index < EnvWords<P>::value, PHX_ERR_INVALID,
                  "PHX_FIELD_ENV_STATE: this env class has no such env-level word");
      *p = d_env + (size_t)index * E;
      *bytes = sizeof(int32_t) * (size_t)E;
      return PHX_OK;
    }
    const int w = field - PHX_FIELD_FAMILY;
    if (w >= 0 && w < P::NWORDS) {
      *p = d_state + (size_t)w * E * G;
      *bytes = sizeof(int32_t) * (size_t)E * G;
      return PHX_OK;
    }
    set_error("unknown family field " + std::to_string(field));
    return PHX_ERR_INVALID;
  }

  const char* exec_name() const override { return name.c_str(); }

  EngineSpec espec{};
  int G = 8;
  bool thread_per_env = false;
  bool static_plan = false;  // the last jit_source() carried a StaticPlan
  bool collective_ok_ = false;  // the program resolves this env class's mail with tile collectives
  int qcap1 = 0;
  int32_t* d_state = nullptr;
  float* d_rcache = nullptr;
  uint32_t* d_rnone = nullptr;
  float* d_ocache = nullptr;
  uint32_t* d_ocached = nullptr;
  cudaLibrary_t jit_lib = nullptr;   // specialised build of the step kernel (load_specialised)
  cudaKernel_t jit_kernel = nullptr;
  int32_t* d_env = nullptr;   // [ENVW][E] env-level words (programs with ENVW > 0)
  int32_t* d_carry_n = nullptr;  // waiting mail (thread-per-env engine, see init)
  int32_t* d_carry = nullptr;
  uint32_t* d_adj = nullptr;  // StochasticNetwork only
  uint2* d_base = nullptr;
  int32_t n_base = 0;
  std::string name;
};

}  // namespace phx
