// fam_simple_market.cu -- device program family PHX_FAMILY_SIMPLE_MARKET: the reference's own
// second example environment, /root/reference/examples/environments/simple_market/
//   market_agents.py:33-85    BuyerAgent  (decode_action :46-63, encode_observation :65-69,
//                             compute_reward :71-74, handle_price_message :76-78, reset :80-83)
//   market_agents.py:92-129   SellerAgent (decode_action :108-112, encode_observation :114-117,
//                             compute_reward :119-122, reset :124-127, handle_order_message :129-132)
//   simple_mkt_env.py:9-58    SimpleMarketEnv: two FSM stages "Buyers" / "Sellers", custom
//                             EnvView field avg_price (:46-47) and an ENV-LEVEL
//                             post_message_resolution (:49-58): avg_price = np.mean(seller prices)
// -- the env class that exercises the engine's env-level words and hooks (phx_engine.cuh).
//
// All arithmetic of the example is float64 (python floats); it is restated in float64 with the
// reference's association and one rounding per operation (no FMA contraction), and cast to
// float32 only at the output planes.
//
// Agent kinds: 0 BuyerAgent, 1 SellerAgent (both strategic).   Payload types: 0 Price(price:
// float64, two words), 1 Order(vol: int).
// The program is a template over the largest number of sellers a buyer can remember, MAXS = 7
// (19 state words; every shipped / tutorial cast) or 15 (36 words), picked per env class.
// State words (per slot):
//   buyer   2k, 2k+1 (k < MAXS)  seller_prices[seller ordinal k]  (float64 lo / hi)
//           W_ORD (= 2 MAXS; one word for MAXS = 7, two for 15)  the dict's insertion order:
//                nibble j = ordinal of the j-th seller heard, the TOP nibble = number of sellers
//                heard (python dicts keep insertion order: it decides which of several cheapest
//                sellers random.choice can return)
//           W_REW, +1  current_reward      W_VAL, +1  type.value     (MAXS = 7: 15 and 17)
//   seller  0, 1 current_price   2, 3 current_revenue   4, 5 current_tx
// Env words: 0, 1 avg_price (float64); survives reset() like the attribute it mirrors.
// iparams: 0 n_sellers, 1 + k = slot of seller ordinal k.
// agent_iparam[slot] = {buyer / seller ordinal, index of the agent's reset draw or -1 (buyer whose
//                       type.value is a constant), -, buyer: ceil(demand_prob * 2^24)}
// agent_fparam[slot] = {low (or the constant), high} of the buyer's UniformFloatSampler / of the
//                      seller's action-space Box.
// RNG (24-bit draws): stream 6 reset draws (step 0, idx = k-th draw of the reset in agent order:
// UniformFloatSampler.sample / Box.sample), stream 7 random.choice among the cheapest sellers
// (idx = buyer ordinal), stream 8 np.random.binomial(1, demand_prob) (idx = buyer ordinal).
#include "phx_engine_host.cuh"
#ifndef PHX_JIT_TU
#include "phx_engine_wide_host.cuh"
#endif

namespace phx {
namespace {

enum { SM_BUYER = 0, SM_SELLER = 1 };
enum { SM_PRICE = 0, SM_ORDER = 1 };
constexpr int SM_STREAM_RESET = 6, SM_STREAM_CHOICE = 7, SM_STREAM_DEMAND = 8;

template <int W>
using WordC = std::integral_constant<int, W>;

template <int MAXS>
struct SimpleMarketProgramT {
  static_assert(MAXS == 7 || MAXS == 15, "order nibbles + count fill one or two words");
  // run-time specialisation (phx_jit.cuh): where this program lives and what it is called
  static constexpr const char* JIT_SOURCE = "fam_simple_market.cu";
  static constexpr const char* JIT_NAME = MAXS == 7 ? "SimpleMarketProgramT<7>" : "SimpleMarketProgramT<15>";
  // buyer state layout
  static constexpr int W_ORD = 2 * MAXS, OW = MAXS <= 7 ? 1 : 2, W_REW = W_ORD + OW,
                       W_VAL = W_REW + 2;
  static constexpr int CNT_SHIFT = OW == 1 ? 28 : 60;
  // a seller prices every neighbour in the acting phase; no handler answers a message
  static constexpr int PW = 2, NWORDS = W_VAL + 2, VW = 0, ACTCAP = 32, RESPCAP = 1, OBS_DIM = 3,
                       ACT_DIM = 1, Q1CAP = 16, ENVW = 2;
  static constexpr int RECVCAP = 32;
  // compact acting queue (phx_engine.cuh): a seller prices its neighbours, a buyer sends at most
  // one order -- (s + 1)(32 - s) entries for a 32-agent market with s sellers
  static constexpr int ACTTOTAL = MAXS == 7 ? 256 : 272;
  __host__ __device__ static int act_cap(int kind, int out_degree) {
    return kind == 1 /* SM_SELLER */ ? out_degree : 1;
  }
  // 128-lane block engine (phx_engine_wide.cuh): the same fan-outs; nobody answers a message
  static constexpr bool WIDE_OK = true;
  __host__ __device__ static int wide_act_cap(int kind, int degree, int) {
    return kind == 1 /* SM_SELLER */ ? (degree > 0 ? degree : 1) : 1;
  }
  __host__ __device__ static int wide_resp_cap(int, int, int) { return 1; }
  static constexpr bool BATCHED = false, HAS_PRE = false, HAS_POST = false;

  // thread-per-env engine: messages in flight in one round = the largest acting-phase fan-out
  static int q1_cap(const phx_spec& s) {
    int prices = 0, buyers = 0;
    for (int i = 0; i < s.n_agents; ++i) {
      if (s.agent_kind[i] == SM_SELLER) {
        for (int r = 0; r < s.n_agents; ++r) prices += mask_bit(s.adjacency[i], r);
      } else {
        ++buyers;
      }
    }
    return std::max(std::max(prices, buyers), 1);
  }

  static int32_t validate(const phx_spec& s) {
    PHX_REQUIRE(s.env_kind == PHX_ENV_FSM, PHX_ERR_UNSUPPORTED,
                "simple-market family runs under FiniteStateMachineEnv only");
    PHX_REQUIRE(s.obs_dim == 3 && s.act_dim == 1 && s.n_payload_types == 2, PHX_ERR_INVALID,
                "simple-market family: obs_dim 3, act_dim 1, 2 payload types");
    PHX_REQUIRE(s.iparams[0] >= 1 && s.iparams[0] <= MAXS, PHX_ERR_UNSUPPORTED,
                "simple-market family: 1..15 sellers");
    return PHX_OK;
  }

  __device__ static double dbl(const int* st, int w) { return __hiloint2double(st[w + 1], st[w]); }
  __device__ static void put(int* st, int w, double v) {
    st[w] = __double2loint(v);
    st[w + 1] = __double2hiint(v);
  }
  // the buyers' dict order: nibbles + count in one (MAXS = 7) or two (MAXS = 15) words
  __device__ static uint64_t order_of(const int* st) {
    uint64_t o = (uint32_t)st[W_ORD];
    if (OW > 1) o |= (uint64_t)(uint32_t)st[W_ORD + OW - 1] << 32;
    return o;
  }
  __device__ static void put_order(int* st, uint64_t o) {
    st[W_ORD] = (int)(uint32_t)o;
    if (OW > 1) st[W_ORD + OW - 1] = (int)(uint32_t)(o >> 32);
  }
  __device__ static int heard_of(uint64_t o) { return (int)(o >> CNT_SHIFT); }
  __device__ static int nth(uint64_t o, int j) { return (int)((o >> (4 * j)) & 15u); }
  // seller_prices[ordinal k] with k only known at run time: static indexing keeps st[] in registers
  __device__ static double price_of(const int* st, int k) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < MAXS; ++j)
      if (j == k) v = dbl(st, 2 * j);
    return v;
  }
  // min(self.seller_prices.values()) over the sellers heard so far; heard >= 1
  __device__ static double min_price(const int* st, int heard) {
    const uint64_t order = order_of(st);
    double best = price_of(st, nth(order, 0));
    for (int j = 1; j < heard; ++j) {
      const double p = price_of(st, nth(order, j));
      if (p < best) best = p;
    }
    return best;
  }

  template <class C, class E>
  __device__ static void act(const C& c, int* st, bool has_action, const float* action, E& out) {
    const auto& sp = *c.spec;
    if (!has_action) return;  // Agent.generate_messages default: nothing (agents.py:157-158)
    const float a0 = action[0];
    if (!(fabsf(a0) <= 1048576.0f)) {
      out.fault = PHX_FAULT_INVALID_ACTION;
      return;
    }
    if (c.kind == SM_SELLER) {  // market_agents.py:108-112
      const double price = (double)a0;
      put(st, 0, price);
      for (int r = c.next_neighbour(-1); r >= 0; r = c.next_neighbour(r))  // neighbour_ids, slot order
        out.send(r, SM_PRICE, __double2loint(price), __double2hiint(price));
      return;
    }
    // BuyerAgent.decode_action, market_agents.py:46-63 (Discrete(2) action: int(round(a)))
    const uint64_t order = order_of(st);
    const int heard = heard_of(order);
    if (heard == 0) {  // min() of an empty dict raises ValueError in the reference
      out.fault = PHX_FAULT_INVALID_ACTION;
      return;
    }
    const int vol = __float2int_rn(a0);
    if (vol == 0) return;
    const double best = min_price(st, heard);
    // min_sellers = the cheapest sellers in dict (= first-heard) order; random.choice -> contract
    int n_ties = 0;
    for (int j = 0; j < heard; ++j) n_ties += price_of(st, nth(order, j)) == best;
    const int pick = rng_randint(
        c.rand24_hi(SM_STREAM_CHOICE, (uint32_t)sp.agent_iparam[c.slot][0]), (uint32_t)n_ties);
    int seller = 0, seen = 0;
    for (int j = 0; j < heard; ++j) {
      const int k = nth(order, j);
      if (price_of(st, k) == best) {
        if (seen == pick) seller = k;
        ++seen;
      }
    }
    out.send(sp.iparams[1 + seller], SM_ORDER, vol);
    // current_reward += -action * min_price + type.value
    const double gain = __dadd_rn(__dmul_rn(-(double)vol, best), dbl(st, W_VAL));
    put(st, W_REW, __dadd_rn(dbl(st, W_REW), gain));
  }

  template <class C>
  __device__ static void view(const C&, const int*, int*) {}
  template <class C>
  __device__ static void pre(const C&, int*) {}
  template <class C>
  __device__ static void post(const C&, int*) {}

  // SimpleMarketEnv.post_message_resolution (simple_mkt_env.py:49-58): np.mean over the sellers'
  // current prices in agent order.  numpy's add.reduce takes element 0 as the initial value and
  // adds the PAIRWISE sum of the rest to it (numpy/_core/src/umath/loops_utils.h.src,
  // pairwise_sum; numpy 2.3 here, the routine is unchanged since 1.9):
  //   fewer than 8 elements  a plain left-to-right loop starting from 0.0
  //   8 .. 128 elements      eight accumulators r[j] = a[j], r[j] += a[8 i + j] over the full
  //                          groups of eight, ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)),
  //                          then the leftover elements one by one
  // (at most 14 "rest" elements here, so the accumulators only ever hold the first eight).
  template <class C, class Acc>
  __device__ static void env_post(const C& c, int* env, Acc agent_word) {
    const auto& sp = *c.spec;
    const int n = sp.iparams[0];
    double v[MAXS];
#pragma unroll
    for (int k = 0; k < MAXS; ++k) {
      v[k] = 0.0;
      if (k < n) {
        const int slot = sp.iparams[1 + k];
        v[k] = __hiloint2double(agent_word(slot, WordC<1>{}), agent_word(slot, WordC<0>{}));
      }
    }
    const int m = n - 1;  // the "rest": v[1 .. n)
    double rest = 0.0;
    if (MAXS < 9 || m < 8) {
#pragma unroll
      for (int k = 1; k < MAXS; ++k)
        if (k < n) rest = __dadd_rn(rest, v[k]);
    } else {
      if constexpr (MAXS >= 9) {
        rest = __dadd_rn(__dadd_rn(__dadd_rn(v[1], v[2]), __dadd_rn(v[3], v[4])),
                         __dadd_rn(__dadd_rn(v[5], v[6]), __dadd_rn(v[7], v[8])));
#pragma unroll
        for (int k = 9; k < MAXS; ++k)
          if (k < n) rest = __dadd_rn(rest, v[k]);
      }
    }
    const double avg = __ddiv_rn(n > 1 ? __dadd_rn(v[0], rest) : v[0], (double)n);
    env[0] = __double2loint(avg);
    env[1] = __double2hiint(avg);
  }

  template <class C, class E>
  __device__ static bool handle(const C& c, int* st, const Msg& m, E&) {
    if (c.kind == SM_BUYER) {  // market_agents.py:76-78: seller_prices[sender] = price
      if (m.type != SM_PRICE) return false;
      const int k = c.iparam0_of(m.sender);
      uint64_t order = order_of(st);
      const int heard = heard_of(order);
      bool known = false;
      for (int j = 0; j < heard; ++j) known |= nth(order, j) == k;
      if (!known) {  // a new dict key goes to the end
        const uint64_t cnt_mask = (uint64_t)15u << CNT_SHIFT;
        order = (order & ~cnt_mask & ~((uint64_t)15u << (4 * heard))) | ((uint64_t)k << (4 * heard)) |
                ((uint64_t)(heard + 1) << CNT_SHIFT);
        put_order(st, order);
      }
#pragma unroll
      for (int j = 0; j < MAXS; ++j)
        if (j == k) {
          st[2 * j] = m.p[0];
          st[2 * j + 1] = m.p[1];
        }
      return true;
    }
    if (m.type != SM_ORDER) return false;  // market_agents.py:129-132
    const double vol = (double)m.p[0];
    put(st, 2, __dadd_rn(dbl(st, 2), __dmul_rn(dbl(st, 0), vol)));
    put(st, 4, __dadd_rn(dbl(st, 4), vol));
    return true;
  }

  template <class C>
  __device__ static bool encode(const C& c, int* st, float* obs) {
    const auto& sp = *c.spec;
    if (c.kind == SM_SELLER) {  // [current_tx, env_view.avg_price]; current_tx = 0
      obs[0] = (float)dbl(st, 4);
      obs[1] = (float)__hiloint2double(c.env[1], c.env[0]);
      obs[2] = 0.f;
      put(st, 4, 0.0);
      return true;
    }
    // buyer: [min price, demand ~ binomial(1, demand_prob), type.value]
    const int heard = heard_of(order_of(st));
    // (the reference raises ValueError when no price was ever heard; the lowering only accepts
    // env classes whose sellers act first, so a quiet market shows up as +inf here)
    obs[0] = heard > 0 ? (float)min_price(st, heard) : __int_as_float(0x7f800000);
    const uint32_t d24 = c.rand24_hi(SM_STREAM_DEMAND, (uint32_t)sp.agent_iparam[c.slot][0]) >> 8;
    obs[1] = d24 < (uint32_t)sp.agent_iparam[c.slot][3] ? 1.f : 0.f;
    obs[2] = (float)dbl(st, W_VAL);
    return true;
  }

  template <class C>
  __device__ static float reward(const C& c, int* st) {
    // current_revenue / current_reward, then cleared
    double r;
    if (c.kind == SM_SELLER) {
      r = dbl(st, 2);
      put(st, 2, 0.0);
    } else {
      r = dbl(st, W_REW);
      put(st, W_REW, 0.0);
    }
    return (float)r;
  }
  template <class C>
  __device__ static bool terminated(const C&, const int*) { return false; }
  template <class C>
  __device__ static bool truncated(const C&, const int*) { return false; }

  template <class C>
  __device__ static void reset_agent(const C& c, int* st) {
    const auto& sp = *c.spec;
    const int idx = sp.agent_iparam[c.slot][1];
    double v = sp.agent_fparam[c.slot][0];
    if (idx >= 0) {  // np.random.uniform(low, high) = low + (high - low) * u, float64
      const double u = (double)(c.rand24_hi(SM_STREAM_RESET, (uint32_t)idx) >> 8) * (1.0 / 16777216.0);
      v = __dadd_rn(sp.agent_fparam[c.slot][0],
                    __dmul_rn(__dsub_rn(sp.agent_fparam[c.slot][1], sp.agent_fparam[c.slot][0]), u));
    }
    if (c.kind == SM_SELLER) {  // current_price = action_space.sample() (a float32 Box)
      put(st, 0, (double)(float)v);
      put(st, 2, 0.0);
      put(st, 4, 0.0);
    } else {  // Agent.reset: type = supertype.sample(); seller_prices = {}; current_reward = 0
      put(st, W_VAL, v);
      put_order(st, 0);
      put(st, W_REW, 0.0);
    }
  }
};

using SimpleMarketProgram = SimpleMarketProgramT<7>;

#ifndef PHX_JIT_TU
// Base = EngineFamily<P> (tile / thread engines) or WideFamily<P> (block engine)
template <class Base>
class SimpleMarketFamily final : public Base {
 public:
  int32_t init(const phx_spec& s) override {
    // seller ordinal -> slot table for the buyers' orders and the env-level mean
    phx_spec t = s;
    int k = 0;
    for (int i = 0; i < s.n_agents; ++i)
      if (s.agent_kind[i] == SM_SELLER && k < 15) t.iparams[1 + k++] = i;
    PHX_REQUIRE(k == s.iparams[0], PHX_ERR_INVALID,
                "simple-market family: iparams[0] must be the number of SellerAgents");
    this->spec = t;  // (jit_source lowers the handle's own spec again)
    return Base::init(t);
  }
};
#endif

}  // namespace

#ifndef PHX_JIT_TU  // a specialised translation unit only needs the program above
Family* make_simple_market_family(const phx_spec& s) {
  const bool wide = s.n_agents > ENGINE_MAX_AGENTS || s.exec_mode == PHX_EXEC_WIDE;
  if (s.iparams[0] <= 7) {
    if (wide) return new SimpleMarketFamily<WideFamily<SimpleMarketProgramT<7>>>();
    return new SimpleMarketFamily<EngineFamily<SimpleMarketProgramT<7>>>();
  }
  if (wide) return new SimpleMarketFamily<WideFamily<SimpleMarketProgramT<15>>>();
  return new SimpleMarketFamily<EngineFamily<SimpleMarketProgramT<15>>>();
}
#endif

}  // namespace phx
