// fam_digital_ads.cu -- device program family PHX_FAMILY_DIGITAL_ADS: the reference's third
// example environment, /root/reference/examples/environments/digital_ads_market/
// digital_ads_market.py
//   :140-193  PublisherAgent   generate_messages -> ImpressionRequest(random user) to the exchange;
//                              handle_ads: clicked ~ binomial(1, p[user][theme]) -> ImpressionResult
//   :196-363  AdvertiserAgent  handle_impression_request / _auction_result / _impression_result,
//                              pre_message_resolution, decode_action (bid = min(a * budget, left)),
//                              encode_observation (None before the first impression),
//                              compute_reward (clicks of the step), is_terminated (budget spent),
//                              reset (type.budget from an env-managed, clipped UniformFloatSampler)
//   :366-515  AdExchangeAgent  handle_impression_request: forward to every advertiser;
//                              handle_batch OVERRIDE (:429-454): every Bid of the batch -> one
//                              auction (:456-515, stable descending sort: the first of equal
//                              highest bids wins; first / second price) -> Ads to the publisher,
//                              AuctionResult to every bidder in batch order
//   :518-591  DigitalAdsEnv    stages publisher_step / advertiser_step, StochasticNetwork,
//                              ignore_connection_errors, BatchResolver(round_limit=5)
// Money is float64 in the reference (np.float64 budgets from the samplers); it is float64 here,
// one rounding per operation, float32 only in the observation plane.
//
// Agent kinds: 0 AdExchangeAgent, 1 PublisherAgent, 2 AdvertiserAgent (strategic).
// Payload types: 0 ImpressionRequest(user), 1 Bid(bid: f64), 2 AuctionResult(cost: f64),
//                3 Ads(advertiser slot, theme | user << 8), 4 ImpressionResult(clicked).
// State words:
//   advertiser 0,1 left  2,3 type.budget  4,5 bid  6 step_clicks  7 step_wins  8 current user
//              9,10 total_requests[1,2]  11,12 total_wins[1,2]  13,14 total_clicks[1,2]
//   exchange   0 user of the impression on offer;  batch scratch: 1 bids seen, 2,3 highest bid,
//              4,5 second highest, 6 winner slot, 7 + 15,16,17 bidder mask (slots 0-31, 32-127)
// iparams: 0 exchange slot, 1 publisher slot, 2 second-price auction?, 3 + (user-1)*4 + theme =
//          ceil(click probability * 2^24).
// agent_iparam[advertiser] = {index of its sampler in env._samplers or -1, theme id}
// agent_fparam[advertiser] = {low (or the constant budget), high, clip_low, clip_high}
// RNG (24-bit draws): stream 3 samplers (step 0), stream 9 the impression's user, stream 10 the click.
#include "phx_engine_host.cuh"
#ifndef PHX_JIT_TU
#include "phx_engine_wide_host.cuh"
#endif

namespace phx {
namespace {

enum { DA_EXCHANGE = 0, DA_PUBLISHER = 1, DA_ADVERTISER = 2 };
enum { DA_IMPRESSION = 0, DA_BID = 1, DA_RESULT = 2, DA_ADS = 3, DA_CLICK = 4 };
constexpr int DA_STREAM_SAMPLER = 3, DA_STREAM_USER = 9, DA_STREAM_CLICK = 10;

struct DigitalAdsProgram {
  static constexpr const char* JIT_SOURCE = "fam_digital_ads.cu";
  static constexpr const char* JIT_NAME = "DigitalAdsProgram";
  // acting phase: one message per agent; the exchange answers a round with up to 31 messages
  static constexpr int PW = 2, NWORDS = 18, VW = 0, ACTCAP = 1, RESPCAP = 32, OBS_DIM = 3,
                       ACT_DIM = 1, Q1CAP = 8;
  static constexpr int RECVCAP = 32;
  static constexpr bool BATCHED = true, HAS_PRE = true, HAS_POST = false;
  // compact response queues: only the exchange answers a round with many messages (one per
  // agent at most: the forwarded impression, or Ads + one AuctionResult per bidder); the
  // publisher answers with one ImpressionResult, an advertiser never answers
  static constexpr int RESPTOTAL = 96;
  __host__ __device__ static int resp_cap(int kind, int) { return kind == 0 /* exchange */ ? 32 : 1; }
  // the 128-lane block engine sizes its queues from the env class (phx_engine_wide.cuh): the
  // exchange answers a round with one message per agent at most, everybody else with one
  static constexpr bool WIDE_OK = true;
  __host__ __device__ static int wide_act_cap(int, int, int) { return 1; }
  __host__ __device__ static int wide_resp_cap(int kind, int, int n_agents) {
    return kind == 0 /* exchange */ ? n_agents : 1;
  }
  // word of the exchange's bidder mask that holds slot s
  __device__ static int bidder_word(int s) { return s < 32 ? 7 : 14 + (s >> 5); }

  static int q1_cap(const phx_spec& s) { return s.n_agents; }

  static int32_t validate(const phx_spec& s) {
    PHX_REQUIRE(s.env_kind == PHX_ENV_FSM, PHX_ERR_UNSUPPORTED,
                "digital-ads family runs under FiniteStateMachineEnv only");
    PHX_REQUIRE(s.obs_dim == 3 && s.act_dim == 1 && s.n_payload_types == 5, PHX_ERR_INVALID,
                "digital-ads family: obs_dim 3, act_dim 1, 5 payload types");
    int ex = 0, pub = 0;
    for (int i = 0; i < s.n_agents; ++i) {
      ex += s.agent_kind[i] == DA_EXCHANGE;
      pub += s.agent_kind[i] == DA_PUBLISHER;
    }
    PHX_REQUIRE(ex == 1 && pub == 1, PHX_ERR_UNSUPPORTED, "one exchange and one publisher per env");
    return PHX_OK;
  }

  __device__ static double dbl(const int* st, int w) { return __hiloint2double(st[w + 1], st[w]); }
  __device__ static void put(int* st, int w, double v) {
    st[w] = __double2loint(v);
    st[w + 1] = __double2hiint(v);
  }

  template <class C, class E>
  __device__ static void act(const C& c, int* st, bool has_action, const float* action, E& out) {
    const auto& sp = *c.spec;
    if (c.kind == DA_PUBLISHER) {  // generate_messages :163-164: np.random.choice([1, 2])
      const int user = 1 + rng_randint(c.rand24_hi(DA_STREAM_USER, 0u), 2u);
      out.send(sp.iparams[0], DA_IMPRESSION, user);
      return;
    }
    if (c.kind != DA_ADVERTISER || !has_action) return;
    const float a0 = action[0];
    if (!(fabsf(a0) <= 1048576.0f)) {
      out.fault = PHX_FAULT_INVALID_ACTION;
      return;
    }
    // decode_action :313-327: self.bid = min(action[0] * type.budget, left); python's min
    // keeps its first argument unless the second is smaller
    const double want = __dmul_rn((double)a0, dbl(st, 2));
    const double left = dbl(st, 0);
    const double bid = left < want ? left : want;
    put(st, 4, bid);
    if (bid > 0.0) out.send(sp.iparams[0], DA_BID, __double2loint(bid), __double2hiint(bid));
  }

  template <class C>
  __device__ static void view(const C&, const int*, int*) {}
  template <class C>
  __device__ static void pre(const C& c, int* st) {
    if (c.kind == DA_ADVERTISER) st[6] = st[7] = 0;  // :240-246, every step
  }
  template <class C>
  __device__ static void post(const C&, int*) {}

  // AdExchangeAgent.handle_batch :429-454
  template <class C>
  __device__ static void batch_begin(const C& c, int* st) {
    if (c.kind == DA_EXCHANGE) st[1] = 0;
  }

  template <class C, class E>
  __device__ static bool handle(const C& c, int* st, const Msg& m, E& out) {
    const auto& sp = *c.spec;
    if (c.kind == DA_EXCHANGE) {
      if (m.type == DA_IMPRESSION) {  // :417-427: forward to every advertiser, list order
        st[0] = m.p[0];
        for (int adv = c.next_of_kind(DA_ADVERTISER, -1); adv >= 0;
             adv = c.next_of_kind(DA_ADVERTISER, adv))
          out.send(adv, DA_IMPRESSION, m.p[0]);
        return true;
      }
      if (m.type != DA_BID) return false;
      // collected for the auction; the stable descending sort of :498-515 in one pass: a later
      // bid only displaces an earlier one when it is strictly higher
      const double bid = __hiloint2double(m.p[1], m.p[0]);
      if (st[1] == 0) {
        put(st, 2, bid);
        put(st, 4, bid);
        st[6] = m.sender;
        st[7] = st[15] = st[16] = st[17] = 0;
      } else if (bid > dbl(st, 2)) {
        put(st, 4, dbl(st, 2));
        put(st, 2, bid);
        st[6] = m.sender;
      } else if (st[1] == 1 || bid > dbl(st, 4)) {
        put(st, 4, bid);
      }
      st[1] += 1;
      {
        const int bw = bidder_word(m.sender);
        const int bit = (int)(1u << (m.sender & 31));
        if (bw == 7) st[7] |= bit;
        if (C::MASK_WORDS > 1) {
          if (bw == 15) st[15] |= bit;
          if (bw == 16) st[16] |= bit;
          if (bw == 17) st[17] |= bit;
        }
      }
      return true;
    }
    if (c.kind == DA_PUBLISHER) {  // handle_ads :167-193
      if (m.type != DA_ADS) return false;
      const int theme = m.p[1] & 0xFF, user = (m.p[1] >> 8) & 0xFF;
      const uint32_t thr = (uint32_t)sp.iparams[3 + (user - 1) * 4 + theme];
      const int clicked = (c.rand24_hi(DA_STREAM_CLICK, 0u) >> 8) < thr ? 1 : 0;
      out.send(m.p[0], DA_CLICK, clicked);
      return true;
    }
    // AdvertiserAgent
    const int user = st[8];
    if (m.type == DA_IMPRESSION) {  // :249-270
      st[8] = m.p[0];
      if (m.p[0] == 1) st[9] += 1;
      if (m.p[0] == 2) st[10] += 1;
      return true;
    }
    if (m.type == DA_RESULT) {  // :272-281
      const double cost = __hiloint2double(m.p[1], m.p[0]);
      const int won = cost != 0.0 ? 1 : 0;
      st[7] += won;
      if (user == 1) st[11] += won;
      if (user == 2) st[12] += won;
      put(st, 0, __dsub_rn(dbl(st, 0), cost));
      return true;
    }
    if (m.type == DA_CLICK) {  // :283-290
      st[6] += m.p[0];
      if (user == 1) st[13] += m.p[0];
      if (user == 2) st[14] += m.p[0];
      return true;
    }
    return false;
  }

  template <class C, class E>
  __device__ static void batch_end(const C& c, int* st, E& out) {
    if (c.kind != DA_EXCHANGE || st[1] == 0) return;
    const auto& sp = *c.spec;
    // auction :456-496: the highest bid wins; cost = it (first price) or the runner-up's bid
    const double cost = (sp.iparams[2] && st[1] > 1) ? dbl(st, 4) : dbl(st, 2);
    const int winner = st[6];
    const int theme = sp.agent_iparam[winner][1];
    out.send(sp.iparams[1], DA_ADS, winner, theme | (st[0] << 8));
#pragma unroll
    for (int w = 0; w < C::MASK_WORDS; ++w) {  // every bidder, batch (= slot) order
      for (uint32_t b = (uint32_t)(w == 0 ? st[7] : st[14 + w]); b; b &= b - 1) {
        const int adv = 32 * w + __ffs(b) - 1;
        const double charged = adv == winner ? cost : 0.0;
        out.send(adv, DA_RESULT, __double2loint(charged), __double2hiint(charged));
      }
    }
  }

  template <class C>
  __device__ static bool encode(const C& c, int* st, float* obs) {
    if (c.kind != DA_ADVERTISER || st[8] == 0) return false;  // :292-311: None before an impression
    obs[0] = (float)__ddiv_rn(dbl(st, 0), dbl(st, 2));  // budget_left
    obs[1] = (float)dbl(st, 2);                         // type.budget
    obs[2] = (float)(st[8] - 1);                        // user_id
    return true;
  }
  // (1 - 0.0) * step_clicks + (0.0 * left) / budget == float(step_clicks)  (:329-337)
  template <class C>
  __device__ static float reward(const C&, int* st) { return (float)st[6]; }
  template <class C>
  __device__ static bool terminated(const C& c, const int* st) {
    return c.kind == DA_ADVERTISER && dbl(st, 0) <= 0.0;  // :339-343
  }
  template <class C>
  __device__ static bool truncated(const C&, const int*) { return false; }

  template <class C>
  __device__ static void reset_agent(const C& c, int* st) {
#pragma unroll
    for (int w = 0; w < NWORDS; ++w) st[w] = 0;
    if (c.kind != DA_ADVERTISER) return;
    // Agent.reset: type = supertype.sample() -> the env-managed sampler's value of this episode
    // (env.py:212-216; samplers.py:142-147: np.random.uniform, then np.clip)
    const auto& sp = *c.spec;
    const int idx = sp.agent_iparam[c.slot][0];
    double budget = sp.agent_fparam[c.slot][0];
    if (idx >= 0) {
      const double u = (double)(c.rand24_hi(DA_STREAM_SAMPLER, (uint32_t)idx) >> 8) * (1.0 / 16777216.0);
      budget = __dadd_rn(sp.agent_fparam[c.slot][0],
                         __dmul_rn(__dsub_rn(sp.agent_fparam[c.slot][1], sp.agent_fparam[c.slot][0]), u));
      budget = fmin(fmax(budget, sp.agent_fparam[c.slot][2]), sp.agent_fparam[c.slot][3]);
    }
    put(st, 2, budget);
    put(st, 0, budget);  // left = type.budget
  }
};

}  // namespace

#ifndef PHX_JIT_TU
Family* make_digital_ads_family(const phx_spec& s) { return make_engine_family<DigitalAdsProgram>(s); }
#endif

}  // namespace phx
