// fam_supply_chain.cu -- device program family PHX_FAMILY_SUPPLY_CHAIN.
//
// Replaces, for a batch of E env instances, one PhantomEnv.step of
// /root/reference/examples/environments/supply_chain/supply_chain.py:
//   ShopAgent      :70-150  (decode_action :136-142, handle_order_request :104-122,
//                            handle_stock_response :98-102, pre_message_resolution :93-96,
//                            encode_observation :124-134, compute_reward :144-147, reset :149)
//   FactoryAgent   :36-45   (echo StockRequest -> StockResponse)
//   CustomerAgent  :48-67   (generate_messages: OrderRequest(randint(max_order)))
//   SupplyChainEnv :153-175 (agents [SHOP, WAREHOUSE, CUST1..N], star on SHOP)
// driven by phantom/env.py:239-303 and phantom/resolvers.py:128-163.
//
// Agent kinds (phx_spec.agent_kind):  0 = ShopAgent, 1 = FactoryAgent, 2 = CustomerAgent
// Payload types:                      0 = OrderRequest, 1 = OrderResponse,
//                                     2 = StockRequest, 3 = StockResponse
// Family parameters:                  iparams[0] = CUSTOMER_MAX_ORDER_SIZE (5)
//                                     iparams[1] = SHOP_MAX_STOCK (100)
// Family fields (phx_get_field):      PHX_FIELD_FAMILY + 0 = shop state int32 [E,4]
//                                     (stock, sales, missed_sales, delivered_stock)
//
// HBM layout: env header int4 [E] (step, episode, -, -), shop state int4 [E].  Messages,
// contexts and views never touch HBM (unless message tracking is on).
//
// Two kernels:
//  * sc_fast_kernel  -- one THREAD per env.  On the canonical topology the message graph of
//    a step is data independent, so the routing rules are evaluated once per
//    handle on the host (make_plan) into a static schedule, and the kernel executes the
//    handler bodies in that order with the whole env state in registers.
//  * ScProgram below -- the same agents as a device program of the generic queue engines
//    (phx_engine1.cuh / phx_engine.cuh): any agent order / topology, messages routed dynamically.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "phx_engine_host.cuh"
// output store hints: 1 = every output row streaming (st.global.cs), the measured best
// (tools/ab_variants.sh, profiles/r01_ab_store_hints.txt); 0 = reward only, 2 = none, 3 = write-through
#ifndef SC_OUT_HINT
#define SC_OUT_HINT 1
#endif
#include "phx_family.h"
#include "phx_rng.cuh"
#include "phx_sc_wire.h"

namespace phx {
namespace {

enum { SC_SHOP = 0, SC_FACTORY = 1, SC_CUSTOMER = 2 };
enum { SC_ORDER_REQUEST = 0, SC_ORDER_RESPONSE = 1, SC_STOCK_REQUEST = 2, SC_STOCK_RESPONSE = 3 };
constexpr int SC_STREAM_ORDER = 0;       // RNG stream of supply_chain.py:64
constexpr int SC_MAX_CUSTOMERS = 30;
constexpr float SC_MAX_ABS_ACTION = 1048576.0f;  // action contract: finite, |a| <= 2^20

// Static schedule of one step on the canonical topology (see make_plan).
struct ScPlan {
  int32_t E, num_steps, nc, max_order, max_stock;
  int32_t digits_per_word, words_per_step;  // packed draws of the customers' order sizes
  uint32_t flags;
  uint32_t fault[2];     // [shop supplied an action?] -> first fault in event order, 0 = none
  uint32_t deliver_ord;  // bit i: CUSTi's OrderRequest reaches the shop's handler
  uint32_t delivery_ok;  // StockRequest reaches the factory AND StockResponse reaches the shop
  uint32_t push_req, push_resp;  // tracking: message is pushed (passes the send checks)
  uint32_t push_ord, push_ordresp;
  uint64_t seed;
  uint32_t env_offset;
  float max_stock_f, rcp_stock, cap_f, rcp_cap;  // obs denominators and RN(1/d)
};

struct ScArgs {
  ScPlan p;
  int32_t T;
  int32_t env_begin, env_count;  // sub-range of the handle's envs stepped by this launch
  int32_t vec_actions;           // 16-byte action copies are legal (alignment, full warps)
  int4* hdr;
  int4* shop;
  StepIO io;
  FaultSink faults;
  TraceSink trace;
};

// float32(n / d) exactly as the reference computes it (float64 division, then cast):
// Markstein's FMA sequence with the correctly rounded reciprocal rcp = RN(1/d) yields the
// correctly rounded float32 quotient, and RN32(RN64(n/d)) == RN32(n/d) here because |n| < 2^24
// and d < 2^24 are exact in float32 and n/d with d | 100 has a binary expansion of period
// <= 20 < 29 bits, so the float64 rounding can never land on a float32 tie.  Checked
// exhaustively on the device by tests/test_gpu_supply_chain.py::test_ratio_exhaustive.
__device__ __forceinline__ float sc_ratio(int num, float den, float rcp) {
  return ratio_rn(num, den, rcp);  // phx_common.cuh
}

// Tuning knobs (defaults = the measured best; override with -D for A/B)
#ifndef SC_BLOCK_THREADS
#define SC_BLOCK_THREADS 64
#endif
constexpr int SC_BLOCK = SC_BLOCK_THREADS;
#ifndef SC_GROUPS
#define SC_GROUPS 3       // action ring = SC_GROUPS copy groups of four steps
#endif
constexpr int SC_NG = SC_GROUPS;
constexpr int SC_RING = 4 * SC_NG;

// The NC order sizes of one (episode, step): packed draws (phx_rng.cuh) of RNG stream
// SC_STREAM_ORDER, draw i = customer i.  NC <= kpw(max_order) here (the host checks), so a step
// consumes ONE 32-bit word and one Philox block serves four steps.
template <int NC, class Words>
__device__ __forceinline__ void sc_draw_orders(const ScPlan& p, uint32_t env_id, uint32_t episode,
                                               uint32_t step, Words& words,
                                               int (&want)[NC > 0 ? NC : 1]) {
  uint32_t x = words.take(p.seed, env_id, episode, step, SC_STREAM_ORDER);
#pragma unroll
  for (int i = 0; i < NC; ++i) want[i] = rng_next_digit(x, (uint32_t)p.max_order);
}

// One thread per env; T steps per launch with the env state in registers.
//   NC        number of customers if known at compile time (0 = runtime, up to 30)
//   TRACK     record Resolver.tracked_messages rows, one slab of `cap` rows per (step, env)
//   HAS_MASK  an action_mask plane is supplied
//   FULL_IO   all seven output planes are written; otherwise only obs, reward and all_done
//             (the four per-agent mask planes are constant for this env class)
template <int NC, bool TRACK, bool HAS_MASK, bool FULL_IO>
__global__ void __launch_bounds__(SC_BLOCK) sc_fast_kernel(const ScArgs a) {
  const ScPlan& p = a.p;
  // Threads past the last env re-run env E-1: they compute and store bit-identical values
  // (a benign duplicate), which keeps every bounds test out of the step loop.
  const int e_raw = blockIdx.x * SC_BLOCK + threadIdx.x;
  const bool real = e_raw < a.env_count;
  const int e = a.env_begin + (real ? e_raw : a.env_count - 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nc = NC > 0 ? NC : p.nc;
  const uint32_t env_id = p.env_offset + (uint32_t)e;
  const float cap_f = p.cap_f, rcp_cap = p.rcp_cap;
  const float max_stock_f = p.max_stock_f, rcp_stock = p.rcp_stock;
  const uint32_t all_customers = nc >= 32 ? 0xFFFFFFFFu : ((1u << nc) - 1u);
  // the NC > 0 instantiations without tracking are only launched when every customer's order
  // is delivered (rollout_range checks; other graphs take the runtime-N path)
  const bool all_delivered = (NC > 0 && !TRACK) ? true : p.deliver_ord == all_customers;
  const uint32_t E = (uint32_t)p.E;

  int2 h = *reinterpret_cast<const int2*>(a.hdr + e);
  int4 s = a.shop[e];
  // bit 0: a step ran without a shop action, bit 1: with one (selects p.fault[] afterwards);
  // bit 2: an action outside the contract was seen
  uint32_t seen = 0;

  // row = t * E + e indexes every [T,E,...] plane (the host guarantees T * E * 3 < 2^32)
  uint32_t row = (uint32_t)e;

  // Actions reach the kernel through a shared-memory ring filled by cp.async (LDGSTS), four
  // steps per copy group, one to two groups ahead.  HBM latency is ~0.6 us = more than a whole
  // step of this kernel (measured: a register prefetch one step ahead still stalled 31 % of all
  // warp samples on its scoreboard); the async copy has no register to wait on.
  //   vector form (a.vec_actions): ONE 16-byte LDGSTS per lane moves a warp's 32 actions of
  //     FOUR steps (lane -> step lane/8, env quad lane%8: the 32 actions of a step are 128
  //     contiguous bytes of the [T,E] plane) -- per step 1/4 of a copy instruction instead of a
  //     whole one plus its address arithmetic.  Lanes read what other lanes copied, hence the
  //     __syncwarp() after each wait and before a slot is overwritten.
  //   scalar form: every lane copies its own four actions (any alignment, partial warps).
  __shared__ __align__(16) float act_ring[SC_RING][SC_BLOCK];
  const int T = a.T;
  const bool vec = a.vec_actions != 0;
  const int cj = lane >> 3;  // vector form: the step (within a group) this lane copies
  const float* src = vec ? a.io.actions + (size_t)cj * E + (size_t)(e - lane + 4 * (lane & 7))
                         : a.io.actions + e;
  float* const dst = vec ? &act_ring[cj][warp * 32 + 4 * (lane & 7)] : &act_ring[0][threadIdx.x];
  // copies the actions of steps tg .. tg+3 into slots slot0 .. slot0+3 and commits the group
  auto fetch_group = [&](const int tg, const int slot0) {
    if (vec) {
      if (tg + cj < T) cp_async16(dst + slot0 * SC_BLOCK, src);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (tg + k < T) cp_async4(dst + (slot0 + k) * SC_BLOCK, src + (size_t)k * E);
    }
    src += (size_t)4 * E;
    cp_async_commit();
  };
#pragma unroll
  for (int g = 0; g < SC_NG; ++g) fetch_group(4 * g, 4 * g);

  // the order-size word sequence: NC > 0 consumes one word per step, in step order
  PackedWordQueue wordq;
  PackedWords words;  // runtime-N path: several words per step
  const bool auto_reset = (p.flags & PHX_FLAG_AUTO_RESET) != 0;
  // One env transition with action `act`.
  auto one_step = [&](const float act) {
    // customers' OrderRequest sizes (supply_chain.py:64): packed draws, one word per step
    int want[NC > 0 ? NC : 1];
    if (NC > 0) {
      sc_draw_orders<NC>(p, env_id, (uint32_t)h.y, (uint32_t)(h.x + 1), wordq, want);
      if (auto_reset && h.x + 1 == p.num_steps) wordq.flush();  // next step: (episode + 1, 1)
    }
    bool has = true;
    if (HAS_MASK) has = a.io.action_mask[row] != 0;
    h.x += 1;  // env.py:252
    const bool at_max = h.x == p.num_steps;  // env.py:312-318
    const bool wrap = auto_reset && at_max;

    // ---- acting phase (env.py:320-336), agent order SHOP, WAREHOUSE, CUST1..N
    // ShopAgent.decode_action: min(int(round(a)), max_stock - stock); python round() of a
    // float32 is round-half-even == cvt.rni
    const int ask = min(__float2int_rn(act), p.max_stock - s.x);
    if (HAS_MASK) seen |= has ? 2u : 1u;
    if (has && !(fabsf(act) <= SC_MAX_ABS_ACTION)) seen |= 4u;

    int cnt = 0;
    int4* trow = nullptr;
    if (TRACK && real) {
      trow = a.trace.rows + (size_t)row * a.trace.cap;  // one slab per (step, env)
      if (has && p.push_req) trow[cnt++] = trace_row(0, 1, SC_STOCK_REQUEST, ask, 0, 0);
    }

    // ---- pre_message_resolution (supply_chain.py:93-96): sales = missed_sales = 0
    // ---- round 0, receiver SHOP: handle_order_request (supply_chain.py:104-122), serially in
    // push (= customer) order, from the stock held BEFORE this step's delivery.  One order:
    //   sold = min(want, stock); missed += want - sold; stock -= sold; sales += sold
    // (covers both branches of the reference, including negative stock: want >= 0 > stock
    // sells `stock` and leaves 0).  sales telescopes to stock_before - stock_after.
    const int stock_before = s.x;
    int wanted_total = 0;
    int sold_each[NC > 0 ? NC : 1];
    (void)sold_each;
    if (NC == 0) {  // runtime-N path: draw and fill in place
      int i = 0;
      for (int q = 0; q < p.words_per_step; ++q) {
        uint32_t x = words.word(p.seed, env_id, (uint32_t)h.y,
                                (uint32_t)h.x * (uint32_t)p.words_per_step + (uint32_t)q,
                                SC_STREAM_ORDER);
        for (int r = 0; r < p.digits_per_word && i < nc; ++r, ++i) {
          const int w = rng_next_digit(x, (uint32_t)p.max_order);
          if ((p.deliver_ord >> i) & 1u) {
            const int sold = min(w, s.x);
            s.x -= sold;
            wanted_total += w;
          }
        }
      }
    }
    if (NC > 0) {
      if (all_delivered) {
#pragma unroll
        for (int i = 0; i < (NC > 0 ? NC : 1); ++i) {
          const int sold = min(want[i], s.x);
          s.x -= sold;
          wanted_total += want[i];
          if (TRACK) sold_each[i] = sold;
        }
      } else {
#pragma unroll
        for (int i = 0; i < (NC > 0 ? NC : 1); ++i) {
          int sold = 0;
          if ((p.deliver_ord >> i) & 1u) {
            sold = min(want[i], s.x);
            s.x -= sold;
            wanted_total += want[i];
          }
          if (TRACK) sold_each[i] = sold;
        }
      }
    }
    s.y = stock_before - s.x;
    s.z = wanted_total - s.y;

    if (TRACK && NC > 0 && real) {
#pragma unroll
      for (int i = 0; i < (NC > 0 ? NC : 1); ++i)
        if ((p.push_ord >> i) & 1u) trow[cnt++] = trace_row(2 + i, 0, SC_ORDER_REQUEST, want[i], 0, 0);
      // responses generated in round 0, in receiver first-arrival order: WAREHOUSE, SHOP
      if (has && p.push_resp) trow[cnt++] = trace_row(1, 0, SC_STOCK_RESPONSE, ask, 0, 1);
#pragma unroll
      for (int i = 0; i < (NC > 0 ? NC : 1); ++i)
        if ((p.push_ordresp >> i) & 1u)
          trow[cnt++] = trace_row(0, 2 + i, SC_ORDER_RESPONSE, sold_each[i], 0, 1);
      a.trace.cnt[row] = cnt;
    }

    // ---- round 1, receiver SHOP: handle_stock_response (supply_chain.py:98-102)
    if (has && p.delivery_ok) {
      s.w = ask;
      s.x = min(s.x + ask, p.max_stock);
    }

    // ---- outputs (env.py:273-303).  The shop never terminates (agents.py:307,323).
    // compute_reward (supply_chain.py:144-147): float32(sales - 0.1 * stock), the product and
    // the difference rounded in float64 by the reference.  With k = 10*sales - stock the exact
    // value is k/10; |k| < 2^24, and k/10 is either exactly representable or at least
    // 2^-24/10 (relative) away from every float32 rounding boundary, far more than the
    // float64 path's error of a few 2^-53 -- so the reference's result equals the correctly
    // rounded float32 quotient k/10 (sc_ratio).  Pinned by test_reward_identity (CPU,
    // exhaustive over k) and the golden / full-size parity tests.
    const float reward = sc_ratio(10 * s.y - s.x, 10.0f, 0.1f);

    if (wrap) {
      // Network.reset -> ShopAgent.reset: only the stock is cleared (supply_chain.py:149)
      s.x = 0;
      h.x = 0;
      h.y += 1;
    }
    const float o0 = sc_ratio(s.x, max_stock_f, rcp_stock);
    const float o1 = sc_ratio(s.y, cap_f, rcp_cap);
    const float o2 = sc_ratio(s.z, cap_f, rcp_cap);

    {  // three strided 4-byte stores per row beat a shared-memory transpose into 16-byte stores
      float* o = a.io.obs + (size_t)row * 3;
#if SC_OUT_HINT == 1  // every output streaming (evict-first): the rows are never re-read on
      // the device, and a launch writes 118 MB into a 126 MB L2
      __stcs(o, o0); __stcs(o + 1, o1); __stcs(o + 2, o2);
      st_stream(a.io.reward + row, reward);
      __stcs(reinterpret_cast<uchar2*>(a.io.all_done) + row, make_uchar2(0, at_max ? 1 : 0));
#elif SC_OUT_HINT == 2  // no hints
      o[0] = o0; o[1] = o1; o[2] = o2;
      a.io.reward[row] = reward;
      reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(0, at_max ? 1 : 0);
#elif SC_OUT_HINT == 3  // write-through
      __stwt(o, o0); __stwt(o + 1, o1); __stwt(o + 2, o2);
      __stwt(a.io.reward + row, reward);
      __stwt(reinterpret_cast<uchar2*>(a.io.all_done) + row, make_uchar2(0, at_max ? 1 : 0));
#else
      o[0] = o0; o[1] = o1; o[2] = o2;
      st_stream(a.io.reward + row, reward);
      reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(0, at_max ? 1 : 0);
#endif
      if (FULL_IO) {
        a.io.obs_mask[row] = 1;
        a.io.reward_mask[row] = 1;
        a.io.term[row] = 0;
        a.io.trunc[row] = 0;
      }
    }

    row += E;
  };

  // The loop body covers the whole ring (SC_NG groups), so every ring slot is a compile-time
  // offset.  Measured (us/launch): SC_NG = 2: 35.9, 3: 34.1, 4: 35.3, 6: 39.4 (instruction
  // cache); a one-group body with a runtime slot: 34.9 (3 groups) .. 36.6 (8 groups).
  int t0 = 0;
#pragma unroll 1
  for (; t0 + SC_RING <= T; t0 += SC_RING) {
#pragma unroll
    for (int g = 0; g < SC_NG; ++g) {
      cp_async_wait<SC_NG - 1>();  // group g has landed (the SC_NG-1 younger ones may be in flight)
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) one_step(act_ring[4 * g + j][threadIdx.x]);
      __syncwarp();  // every lane has read slots 4g .. 4g+3
      fetch_group(t0 + SC_RING + 4 * g, 4 * g);
    }
  }
  cp_async_wait<0>();  // tail (< SC_RING steps): its groups were issued above
  __syncwarp();
#pragma unroll 1
  for (int t = t0; t < T; ++t) one_step(act_ring[t - t0][threadIdx.x]);

  if (real) {
    *reinterpret_cast<int2*>(a.hdr + e) = h;
    a.shop[e] = s;
    uint32_t fault = 0;
    if (!HAS_MASK) seen |= 2u;
    // first fault in event order: a bad action is detected in decode_action, before any send
    if (seen & 4u) fault = PHX_FAULT_INVALID_ACTION;
    else if ((seen & 2u) && p.fault[1]) fault = p.fault[1];
    else if ((seen & 1u) && p.fault[0]) fault = p.fault[0];
    if (fault) raise_fault(a.faults, e, fault);
  }
}

// ---------------------------------------------------------------------------------------
// sc_fast2_kernel -- the bench kernel (round 2).  Same schedule, same results, ~40 % fewer
// instructions and a much shorter dependent chain per step than sc_fast_kernel:
//
//  * CLOSED-FORM ORDER FILL.  The shop serves its customers' orders serially
//    (handle_order_request, supply_chain.py:104-122): sold_i = min(want_i, stock); stock -=
//    sold_i.  Without message tracking only the totals are observable, and for the whole batch
//        stock_after = max(max(stock, 0) - D, 0),   D = sum_i want_i
//    (greedy fill sells min(D, stock) when stock >= 0; a negative stock -- reachable through
//    negative actions -- is "sold" to the first customer, sold_1 = min(want_1, stock) = stock,
//    which leaves 0 for the others: the reference's quirk, reproduced).  sales = stock_before -
//    stock_after and missed = D - sales as before.  10 dependent instructions become 2.
//  * ORDER TOTAL FROM ONE MULTIPLY.  The five order sizes of a step are the first five base-n
//    digits of the fraction word / 2^32 (phx_rng.cuh, packed draws), so the 5-digit number they
//    form is N = umulhi(word, n^5) -- ONE IMAD.HI -- and D = digitsum_n(N) is one byte load from a
//    n^5-entry table in shared memory (3 125 B for n = 5; built on the host).  Was: five chained
//    IMAD.WIDE + four adds.
//  * PHILOX-ALIGNED GROUPS.  Word number g = current_step of a (env, episode) lives in Philox
//    block g >> 2, so whenever every env of a warp is at a step with g % 4 == 0 the next FOUR
//    steps take their words from ONE block at compile-time positions: no word queue, no per-step
//    refill branch.  The state-independent parts of the four steps (word -> D, action -> rint)
//    are computed first and are independent of each other, the four 7-instruction state updates
//    follow, then the four output rows: the scheduler sees three batches of independent work
//    instead of one 100-instruction chain.  Steps that are not aligned (the first three after a
//    reset, the episode's last step, envs of one warp at different clocks) take the single-step
//    path -- 4 % of the bench's steps.
//  Valid for: 5 customers whose orders are all delivered, no action mask, no tracking,
//  max_order^5 <= SC2_MAX_TABLE (rollout_range checks; everything else runs sc_fast_kernel).
#ifndef SC2_INNER_LOOP
#define SC2_INNER_LOOP 0
#endif
#ifndef SC2_AR_INSTANCE
#define SC2_AR_INSTANCE 0
#endif
constexpr int SC2_MAX_TABLE = 16384;
constexpr int SC2_RING = 16;  // action ring: four copy groups of four steps

struct Sc2Args {
  ScArgs a;
  const uint8_t* dsum;  // [n^5] digit sums, device memory
  uint32_t pow5;        // n^5
  // compact wire plane of the host-buffer path (phx_sc_wire.h): one 32-bit word per env-step,
  // [T, E]; *wire_overflow is set if a value of the launch does not fit its field
  uint32_t* wire;
  uint32_t* wire_overflow;
  // programmatic dependent launch: let the NEXT launch on the stream start its blocks once this
  // one has `pdl_lead` steps left (its launch latency, block scheduling and table fill then
  // overlap this grid's tail); T + 1 = never (the implicit trigger at grid completion)
  int32_t pdl_lead;
};

template <bool FULL_IO, bool VEC, bool WIRE, int RING = SC2_RING, bool AR = false>
__global__ void __launch_bounds__(SC_BLOCK) sc_fast2_kernel(const Sc2Args args) {
  const ScArgs& a = args.a;
  const ScPlan& p = a.p;
  extern __shared__ __align__(16) unsigned char sc2_smem[];
  float (*act_ring)[SC_BLOCK] = reinterpret_cast<float (*)[SC_BLOCK]>(sc2_smem);
  const uint8_t* dsum = sc2_smem + sizeof(float) * RING * SC_BLOCK;
  // Programmatic dependent launch (PHX_PDL=1): this grid may have been started while the previous
  // launch on the stream was still stepping; it runs its prologue (the table fill below) and
  // waits at griddep_wait() before it touches anything a previous kernel may have written.
  {  // digit-sum table -> shared memory (4-byte words; the host pads the table to 16 bytes)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(args.dsum);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sc2_smem + sizeof(float) * RING * SC_BLOCK);
    for (uint32_t i = threadIdx.x; i < (args.pow5 + 3u) / 4u; i += SC_BLOCK) dst[i] = src[i];
  }
  const int e_raw = blockIdx.x * SC_BLOCK + threadIdx.x;
  const bool real = e_raw < a.env_count;
  const int e = a.env_begin + (real ? e_raw : a.env_count - 1);  // benign duplicates past the end
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t env_id = p.env_offset + (uint32_t)e;
  const float cap_f = p.cap_f, rcp_cap = p.rcp_cap;
  const float max_stock_f = p.max_stock_f, rcp_stock = p.rcp_stock;
  const uint32_t E = (uint32_t)p.E;
  const uint32_t pow5 = args.pow5;
  const int max_stock = p.max_stock, num_steps = p.num_steps;
  // AR instantiation: launched only for handles with PHX_FLAG_AUTO_RESET (the flag as a constant)
  const bool auto_reset = AR || (p.flags & PHX_FLAG_AUTO_RESET) != 0;

  griddep_wait();  // env state, actions: written by earlier work on the stream
  int2 h = *reinterpret_cast<const int2*>(a.hdr + e);
  int4 s = a.shop[e];
  bool bad_action = false;
  uint32_t row = (uint32_t)e;  // t * E + e (the host guarantees T * E * 3 < 2^32)

  // ---- action ring (cp.async, see sc_fast_kernel): copy group m = steps 4m .. 4m+3 -> slots
  // (4m .. 4m+3) % 16; group m+4 is fetched as soon as every slot of group m has been read
  const int T = a.T;
  const int cj = lane >> 3;
  const float* src = VEC ? a.io.actions + (size_t)cj * E + (size_t)(e - lane + 4 * (lane & 7))
                         : a.io.actions + e;
  float* const dst = VEC ? &act_ring[cj][warp * 32 + 4 * (lane & 7)] : &act_ring[0][threadIdx.x];
  int fetch_t = 0;  // first step of the next group to fetch
  auto fetch_group = [&]() {
    const int slot0 = fetch_t & (RING - 1);
    if (VEC) {
      if (fetch_t + cj < T) cp_async16(dst + slot0 * SC_BLOCK, src);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (fetch_t + k < T) cp_async4(dst + (slot0 + k) * SC_BLOCK, src + (size_t)k * E);
    }
    src += (size_t)4 * E;
    fetch_t += 4;
    cp_async_commit();
  };
#pragma unroll
  for (int g = 0; g < RING / 4; ++g) fetch_group();
  __syncthreads();  // the table

  // ---- the three parts of a step
  // state update: acting phase + pre hook + round 0 (orders) + round 1 (delivery); returns k =
  // 10 * sales - stock for the reward
  auto update = [&](const int r, const int D) {
    const int ask = min(r, max_stock - s.x);        // decode_action, supply_chain.py:136-142
    const int before = s.x;
    const int after = max(max(before, 0) - D, 0);   // handle_order_request x 5, closed form
    s.y = before - after;                           // sales
    s.z = D - s.y;                                  // missed_sales
    s.w = ask;                                      // handle_stock_response, :98-102 (the
    s.x = min(after + ask, max_stock);              // kernel is only launched if it is delivered)
  };
  // outputs of the step; `wrap`: the step ends the episode and the env is reset in place
  auto emit = [&](const bool at_max, const bool wrap) {
    const float reward = sc_ratio(10 * s.y - s.x, 10.0f, 0.1f);
    if (WIRE) {  // the same row as ONE word (phx_sc_wire.h), for the host-buffer path
      const bool fits = s.x >= -SCW_STOCK_BIAS && s.x < SCW_STOCK_BIAS &&
                        (uint32_t)s.y <= SCW_FIELD_MAX && (uint32_t)s.z <= SCW_FIELD_MAX;
      if (!fits) *args.wire_overflow = 1u;
      __stcs(args.wire + row, scw_pack(s.x, s.y, s.z, at_max, wrap));
    }
    if (wrap) {  // Network.reset -> ShopAgent.reset clears the stock only
      s.x = 0;
      h.x = 0;
      h.y += 1;
    }
    const float o0 = sc_ratio(s.x, max_stock_f, rcp_stock);
    const float o1 = sc_ratio(s.y, cap_f, rcp_cap);
    const float o2 = sc_ratio(s.z, cap_f, rcp_cap);
    float* o = a.io.obs + (size_t)row * 3;
    __stcs(o, o0); __stcs(o + 1, o1); __stcs(o + 2, o2);
    st_stream(a.io.reward + row, reward);
    __stcs(reinterpret_cast<uchar2*>(a.io.all_done) + row, make_uchar2(0, at_max ? 1 : 0));
    if (FULL_IO) {
      __stcs(a.io.obs_mask + row, (uint8_t)1);
      __stcs(a.io.reward_mask + row, (uint8_t)1);
      __stcs(a.io.term + row, (uint8_t)0);
      __stcs(a.io.trunc + row, (uint8_t)0);
    }
    row += E;
  };
  auto decode = [&](const float act) {
    if (!(fabsf(act) <= SC_MAX_ABS_ACTION)) bad_action = true;
    return __float2int_rn(act);  // python round() of a float32 is round-half-even == cvt.rni
  };

  PackedWords words;  // single-step path: block cache
  int t = 0;
  // NB Philox blocks = 4 * NB aligned steps: all (r, D) pairs first (independent of the state
  // and of each other: two Philox chains interleave), then the state updates and output rows
  auto fast_groups = [&](auto nbc) {
    constexpr int NB = decltype(nbc)::value;
    const uint32_t blk0 = (uint32_t)(h.x + 1) >> 2;
    int D[4 * NB], r[4 * NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const Philox4 b = rng_word_block(p.seed, env_id, (uint32_t)h.y, blk0 + q, SC_STREAM_ORDER);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        D[4 * q + k] = dsum[__umulhi(b.w[k], pow5)];
        r[4 * q + k] = decode(act_ring[(t + 4 * q + k) & (RING - 1)][threadIdx.x]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4 * NB; ++k) {
      h.x += 1;  // env.py:252
      update(r[k], D[k]);
      // (no wrap inside aligned groups, see `mine` below: with auto-reset no step of a group is
      // the episode's last one)
      emit(AR ? false : h.x == num_steps, false);
    }
    t += 4 * NB;
  };
  auto top_up = [&]() {
    // group (fetch_t / 4 - RING / 4)'s slots are free once t has passed them
    if (t >= fetch_t - (RING - 4)) {
      __syncwarp();
      fetch_group();
    }
  };
  bool triggered = false;
  while (t < T) {
    if (!triggered && t + args.pdl_lead >= T) {  // (warp-uniform: t is)
      griddep_launch_dependents();
      triggered = true;
    }
    // How many aligned 4-step groups can EVERY env of the warp run from here?  (aligned: the
    // next step's word is the first of a Philox block; no auto-reset wrap inside a group)
    const int g = h.x + 1;
    int mine = 0;
    if ((g & 3) == 0) {
      mine = (T - t) >> 2;
      if (auto_reset) mine = min(mine, g <= num_steps ? (num_steps - g) >> 2 : 0);
    }
    int n = __reduce_min_sync(0xFFFFFFFFu, mine);
    if (n > 0) {
      // two groups per trip need a ring deep enough to keep their actions a trip ahead
      for (; n >= 2 && RING >= 32; n -= 2) {
        top_up();
        top_up();
        cp_async_wait<RING / 4 - 3>();  // steps <= t + 7 have landed
        __syncwarp();
        fast_groups(std::integral_constant<int, 2>{});
      }
#if SC2_INNER_LOOP
      for (; n > 0; --n)
#else
      if (n > 0)  // (one group per vote: measured faster than an inner loop over n, 30.2 vs 32.4 us)
#endif
      {
        top_up();
        top_up();  // (after an 8-step trip the ring may be two groups behind)
        cp_async_wait<RING / 4 - 2>();
        __syncwarp();
        fast_groups(std::integral_constant<int, 1>{});
      }
    } else {
      top_up();
      top_up();
      cp_async_wait<RING / 4 - 2>();  // all but the youngest groups have landed: steps <= t + 3
      __syncwarp();
      const uint32_t x = words.word(p.seed, env_id, (uint32_t)h.y, (uint32_t)g, SC_STREAM_ORDER);
      const int D = dsum[__umulhi(x, pow5)];
      const int r = decode(act_ring[t & (RING - 1)][threadIdx.x]);
      h.x += 1;
      update(r, D);
      emit(h.x == num_steps, auto_reset && h.x == num_steps);
      t += 1;
    }
  }
  cp_async_wait<0>();

  if (real) {
    *reinterpret_cast<int2*>(a.hdr + e) = h;
    a.shop[e] = s;
    // first fault in event order: a bad action is detected in decode_action, before any send
    const uint32_t fault = bad_action ? (uint32_t)PHX_FAULT_INVALID_ACTION : p.fault[1];
    if (fault) raise_fault(a.faults, e, fault);
  }
}

// ---------------------------------------------------------------------------------------
// sc_fast3_kernel -- TIME-PARALLEL form of the same schedule.  sc_fast2_kernel
// is latency bound: 65 536 envs are 2 048 warps = 3.5 per scheduler, each walking a ~75
// instruction dependent chain per step (measured: 4 % faster than sc_fast_kernel for 25 % fewer
// instructions).  But only SEVEN of those instructions depend on the env's state:
//
//     stock' = min(max(max(stock, 0) - D, 0) + min(r, max_stock - stock), max_stock)
//
// with D (order total: Philox word -> digit sum) and r (rint of the action) functions of
// (env, step) alone, and the output rows functions of (stock', sales, missed) alone.  So a block
// of FOUR warps owns 32 envs and works in chunks of 16 steps:
//   phase A  warp w, lane e: the four steps of Philox block w of the chunk -- one Philox block,
//            four (r, D) pairs -> shared memory                          [parallel over (e, t)]
//   phase B  warp 0, lane e: the 16-step recurrence of env e, 7 instructions per step,
//            (stock, sales, missed, reward numerator) -> shared memory   [serial in t]
//   phase C  warp w, lane e: four output rows (4 correctly rounded quotients, 5 stores each)
//                                                                        [parallel over (e, t)]
// Same instruction count, four times the warps (13.8 per scheduler) and nearly all of the work
// in independent batches.  Every thread tracks the clock of its lane's env (it does not depend
// on the state), so the (episode, step) RNG coordinates and the truncation flags need no
// communication.  The step groups are aligned to the Philox blocks of the block's first env; an
// env at another clock phase just evaluates two Philox blocks for its four steps (PackedWords).
constexpr int SC3_ENVS = 32;   // envs per block: one lane per env in every warp
constexpr int SC3_CHUNK = 16;  // steps per chunk = four Philox blocks

__device__ __noinline__ Philox4 sc3_refill(uint64_t seed, uint32_t env_id, uint32_t ep, uint32_t blk) {
  return rng_word_block(seed, env_id, ep, blk, SC_STREAM_ORDER);
}

// W = warps per block (4: one step group per warp and chunk; 2: two groups per warp, for builds
// that want more registers per thread).
template <bool FULL_IO, int W, bool WIRE>
__global__ void __launch_bounds__(SC3_ENVS * W, W == 4 ? 14 : 16) sc_fast3_kernel(const Sc2Args args) {
  const ScArgs& a = args.a;
  const ScPlan& p = a.p;
  constexpr int GROUPS = 4 / W;  // step groups per thread and chunk
  __shared__ int2 in_rd[SC3_CHUNK][SC3_ENVS];   // (r, D) of every (step, env) of the chunk
  __shared__ int4 out_st[SC3_CHUNK][SC3_ENVS];  // (stock for the obs, sales, missed, 10*sales-stock)
  __shared__ uint32_t bad[SC3_ENVS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e_raw = blockIdx.x * SC3_ENVS + lane;
  const bool real = e_raw < a.env_count;
  const int e = a.env_begin + (real ? e_raw : a.env_count - 1);  // benign duplicates past the end
  const uint32_t env_id = p.env_offset + (uint32_t)e;
  const uint32_t E = (uint32_t)p.E;
  const uint32_t pow5 = args.pow5;
  const uint8_t* __restrict__ dsum = args.dsum;
  const int max_stock = p.max_stock, num_steps = p.num_steps, T = a.T;

  const int2 h0 = *reinterpret_cast<const int2*>(a.hdr + e);
  // an env wraps (auto-reset) when its clock reaches num_steps exactly; one that is already past
  // it (stepped on without auto-reset before) never does
  const bool wraps = (p.flags & PHX_FLAG_AUTO_RESET) != 0 && h0.x < num_steps;
  int4 s = make_int4(0, 0, 0, 0);
  if (warp == 0) {
    s = a.shop[e];
    bad[lane] = 0;
  }
  bool bad_action = false;
  // virtual time u = t + shift: the groups u >> 2 are the Philox blocks of the block's first env
  const int shift = (__shfl_sync(0xFFFFFFFFu, h0.x, 0) + 1) & 3;
  const int u_end = T + shift;

  // clock of step t (t steps after the launch started): steps since the episode began (= step
  // number - 1) and episode
  auto clock_at = [&](int t, int& since, int& ep) {
    since = h0.x + t;
    ep = h0.y;
    if (wraps) {
      const int q = since / num_steps;
      since -= q * num_steps;
      ep += q;
    }
  };

  for (int u0 = 0; u0 < u_end; u0 += SC3_CHUNK) {
    // ---- phase A: (r, D) of this thread's step groups
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
      const int ub = u0 + 4 * (warp + W * m);
      const int ua = max(ub, shift), uz = min(ub + 4, u_end);  // valid steps [ua, uz)
      if (ua < uz) {
        float act[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          act[k] = ub + k >= ua && ub + k < uz
                       ? __ldcs(a.io.actions + (size_t)(ub + k - shift) * E + e) : 0.f;
        int since, ep;
        clock_at(ua - shift, since, ep);
        uint32_t blk = (uint32_t)(since + 1) >> 2, blk_ep = (uint32_t)ep;
        Philox4 b = rng_word_block(p.seed, env_id, blk_ep, blk, SC_STREAM_ORDER);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ub + k >= ua && ub + k < uz) {
            const uint32_t g = (uint32_t)since + 1u;  // step number = word number
            if ((g >> 2) != blk || (uint32_t)ep != blk_ep) {  // not block aligned / episode ended
              blk = g >> 2;
              blk_ep = (uint32_t)ep;
              b = sc3_refill(p.seed, env_id, blk_ep, blk);
            }
            const uint32_t q = g & 3u;
            const uint32_t x = q == 0u ? b.w[0] : q == 1u ? b.w[1] : q == 2u ? b.w[2] : b.w[3];
            const int D = __ldg(dsum + __umulhi(x, pow5));
            if (!(fabsf(act[k]) <= SC_MAX_ABS_ACTION)) bad_action = true;
            // python round() of a float32 is round-half-even == cvt.rni
            in_rd[(ub + k) & (SC3_CHUNK - 1)][lane] = make_int2(__float2int_rn(act[k]), D);
            since += 1;
            if (wraps && since == num_steps) {
              since = 0;
              ep += 1;
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- phase B: the recurrence, one lane per env
    if (warp == 0) {
      const int c_begin = max(u0, shift), c_end = min(u0 + SC3_CHUNK, u_end);
      int since, ep;
      clock_at(c_begin - shift, since, ep);
#pragma unroll 4
      for (int u = c_begin; u < c_end; ++u) {
        const int2 rd = in_rd[u & (SC3_CHUNK - 1)][lane];
        const int ask = min(rd.x, max_stock - s.x);       // decode_action, supply_chain.py:136-142
        const int before = s.x;
        const int after = max(max(before, 0) - rd.y, 0);  // handle_order_request x 5, closed form
        s.y = before - after;                             // sales
        s.z = rd.y - s.y;                                 // missed_sales
        s.w = ask;                                        // handle_stock_response, :98-102
        s.x = min(after + ask, max_stock);
        const int k = 10 * s.y - s.x;                     // reward numerator (before a reset)
        since += 1;
        if (wraps && since == num_steps) {                // ShopAgent.reset clears the stock only
          since = 0;
          s.x = 0;
        }
        out_st[u & (SC3_CHUNK - 1)][lane] = make_int4(s.x, s.y, s.z, k);
      }
    }
    __syncthreads();
    // ---- phase C: the output rows of this thread's step groups
#pragma unroll
    for (int m = 0; m < GROUPS; ++m) {
      const int ub = u0 + 4 * (warp + W * m);
      const int ua = max(ub, shift), uz = min(ub + 4, u_end);
      if (ua < uz) {
        int since, ep;
        clock_at(ua - shift, since, ep);
        // truncation: the step whose number equals num_steps (env.py:312-318)
        const int k_max = num_steps - 1 - since;  // position of that step in the group, if any
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ub + k >= ua && ub + k < uz) {
            const int j = ub + k - ua;
            const bool at_max = wraps ? (j == k_max) : (since + j + 1 == num_steps);
            const int4 v = out_st[(ub + k) & (SC3_CHUNK - 1)][lane];
            const uint32_t row = (uint32_t)(ub + k - shift) * E + (uint32_t)e;
            const float reward = sc_ratio(v.w, 10.0f, 0.1f);
            const float o0 = sc_ratio(v.x, p.max_stock_f, p.rcp_stock);
            const float o1 = sc_ratio(v.y, p.cap_f, p.rcp_cap);
            const float o2 = sc_ratio(v.z, p.cap_f, p.rcp_cap);
            float* o = a.io.obs + (size_t)row * 3;
            __stcs(o, o0); __stcs(o + 1, o1); __stcs(o + 2, o2);
            st_stream(a.io.reward + row, reward);
            __stcs(reinterpret_cast<uchar2*>(a.io.all_done) + row, make_uchar2(0, at_max ? 1 : 0));
            if (FULL_IO) {
              __stcs(a.io.obs_mask + row, (uint8_t)1);
              __stcs(a.io.reward_mask + row, (uint8_t)1);
              __stcs(a.io.term + row, (uint8_t)0);
              __stcs(a.io.trunc + row, (uint8_t)0);
            }
            if (WIRE) {  // the same row as ONE word (phx_sc_wire.h), for the host-buffer path
              const int stock_pre = 10 * v.y - v.w;  // the stock the reward saw (before a reset)
              const bool fits = stock_pre >= -SCW_STOCK_BIAS && stock_pre < SCW_STOCK_BIAS &&
                                (uint32_t)v.y <= SCW_FIELD_MAX && (uint32_t)v.z <= SCW_FIELD_MAX;
              if (!fits) *args.wire_overflow = 1u;
              __stcs(args.wire + row, scw_pack(stock_pre, v.y, v.z, at_max, wraps && at_max));
            }
          }
        }
      }
    }
  }

  // an out-of-contract action may have been seen by any of the warps of an env
  if (bad_action) atomicOr(&bad[lane], 1u);
  __syncthreads();
  if (warp == 0 && real) {
    int since, ep;
    clock_at(T, since, ep);
    *reinterpret_cast<int2*>(a.hdr + e) = make_int2(since, ep);
    a.shop[e] = s;
    // first fault in event order: a bad action is detected in decode_action, before any send
    const uint32_t fault = bad[lane] ? (uint32_t)PHX_FAULT_INVALID_ACTION : p.fault[1];
    if (fault) raise_fault(a.faults, e, fault);
  }
}

// ---------------------------------------------------------------------------------------
// sc_fast4_kernel -- TWO LANES PER ENV, same warp, no barrier.  sc_fast2_kernel is bound by
// dependent-issue latency at 3.5 warps per scheduler (E is fixed); sc_fast3_kernel doubled and
// quadrupled the warps but paid for it with block barriers and 2.4x the instructions.  Here an
// env is owned by the lanes L and L + 16 of one warp and the work of an 8-step trip (two Philox
// blocks) is split WITHOUT divergence -- both lanes run the same instruction stream on
// different data:
//   lane p (0 / 1)   Philox block blk0 + p -> D of steps 4p .. 4p+3, rint of those four
//                    actions, packed as r * 32 + D (|r| <= 2^20, D <= 25)
//   exchange         four SHFL.BFLY (xor 16): both lanes now hold all eight (r, D) pairs
//   recurrence       both lanes run the eight 7-instruction state updates (redundant: that is
//                    the price, 56 instructions per trip) and keep the (stock, sales, missed)
//                    of THEIR four steps
//   outputs          lane p writes the rows of steps 4p .. 4p+3
// Per env-step the pair issues ~94 thread-instructions instead of 73, on twice the warps.
// Steps that are not part of an aligned trip (the first three after a reset, the last four of
// an episode) run on both lanes, lane 0 writes.
constexpr int SC4_THREADS = 64;  // 32 envs per block

template <bool FULL_IO, bool AR, int RING>
__global__ void __launch_bounds__(SC4_THREADS) sc_fast4_kernel(const Sc2Args args) {
  const ScArgs& a = args.a;
  const ScPlan& p = a.p;
  __shared__ __align__(16) float act_ring[SC4_THREADS / 32][RING][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, el = lane & 15;
  // (launched only with env_count % 32 == 0: every warp owns 16 real envs)
  const int e = a.env_begin + blockIdx.x * (SC4_THREADS / 2) + warp * 16 + el;
  const uint32_t env_id = p.env_offset + (uint32_t)e;
  const float cap_f = p.cap_f, rcp_cap = p.rcp_cap;
  const float max_stock_f = p.max_stock_f, rcp_stock = p.rcp_stock;
  const uint32_t E = (uint32_t)p.E;
  const uint32_t pow5 = args.pow5;
  const uint8_t* __restrict__ dsum = args.dsum;  // (global / L1: 3 KB, hot)
  const int max_stock = p.max_stock, num_steps = p.num_steps, T = a.T;
  const bool auto_reset = AR || (p.flags & PHX_FLAG_AUTO_RESET) != 0;

  int2 h = *reinterpret_cast<const int2*>(a.hdr + e);
  int4 s = a.shop[e];
  bool bad_action = false;

  // ---- action ring: fetch unit = 8 steps x 16 envs = 512 B = one 16-byte cp.async per lane
  // (lane -> step lane / 4, env quad lane % 4)
  const float* src = a.io.actions + (size_t)(lane >> 2) * E + (size_t)(e - el + 4 * (lane & 3));
  int fetch_t = 0;
  auto fetch_unit = [&]() {
    if (fetch_t + (lane >> 2) < T)
      cp_async16(&act_ring[warp][(fetch_t + (lane >> 2)) & (RING - 1)][4 * (lane & 3)], src);
    src += (size_t)8 * E;
    fetch_t += 8;
    cp_async_commit();
  };
#pragma unroll
  for (int g = 0; g < RING / 8; ++g) fetch_unit();
  auto top_up = [&](int t) {  // the unit (fetch_t - RING)'s slots are free once t has passed them
    if (t >= fetch_t - (RING - 8)) {
      __syncwarp();
      fetch_unit();
    }
  };

  auto update = [&](const int r, const int D) {
    const int ask = min(r, max_stock - s.x);        // decode_action, supply_chain.py:136-142
    const int before = s.x;
    const int after = max(max(before, 0) - D, 0);   // handle_order_request x 5, closed form
    s.y = before - after;                           // sales
    s.z = D - s.y;                                  // missed_sales
    s.w = ask;                                      // handle_stock_response, :98-102
    s.x = min(after + ask, max_stock);
  };
  // one output row (t, e) from (stock shown, sales, missed, reward numerator)
  auto emit = [&](const int t, const int stock, const int sales, const int missed, const int k,
                  const bool at_max) {
    const uint32_t row = (uint32_t)t * E + (uint32_t)e;
    const float reward = sc_ratio(k, 10.0f, 0.1f);
    const float o0 = sc_ratio(stock, max_stock_f, rcp_stock);
    const float o1 = sc_ratio(sales, cap_f, rcp_cap);
    const float o2 = sc_ratio(missed, cap_f, rcp_cap);
    float* o = a.io.obs + (size_t)row * 3;
    __stcs(o, o0); __stcs(o + 1, o1); __stcs(o + 2, o2);
    st_stream(a.io.reward + row, reward);
    __stcs(reinterpret_cast<uchar2*>(a.io.all_done) + row, make_uchar2(0, at_max ? 1 : 0));
    if (FULL_IO) {
      __stcs(a.io.obs_mask + row, (uint8_t)1);
      __stcs(a.io.reward_mask + row, (uint8_t)1);
      __stcs(a.io.term + row, (uint8_t)0);
      __stcs(a.io.trunc + row, (uint8_t)0);
    }
  };
  auto decode = [&](const float act) {
    if (!(fabsf(act) <= SC_MAX_ABS_ACTION)) bad_action = true;
    return __float2int_rn(act);  // python round() of a float32 is round-half-even == cvt.rni
  };

  PackedWords words;
  int t = 0;
  while (t < T) {
    // aligned 8-step trips every env of the warp can run from here (no wrap inside a trip)
    const int g = h.x + 1;
    int mine = 0;
    if ((g & 3) == 0) {
      mine = (T - t) >> 3;
      if (auto_reset) mine = min(mine, g <= num_steps ? (num_steps - g) >> 3 : 0);
    }
    int n = __reduce_min_sync(0xFFFFFFFFu, mine);
    if (n > 0) {
      for (; n > 0; --n) {
        top_up(t);
        top_up(t);
        cp_async_wait<RING / 8 - 3>();  // all but the youngest units have landed: steps <= t + 7
        __syncwarp();
        // ---- my half of the trip: one Philox block, four (r, D) pairs
        const Philox4 b = rng_word_block(p.seed, env_id, (uint32_t)h.y,
                                         ((uint32_t)(h.x + 1) >> 2) + (uint32_t)half, SC_STREAM_ORDER);
        int mine4[4], other4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int D = __ldg(dsum + __umulhi(b.w[k], pow5));
          const int r = decode(act_ring[warp][(t + 4 * half + k) & (RING - 1)][el]);
          mine4[k] = r * 32 + D;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) other4[k] = __shfl_xor_sync(0xFFFFFFFFu, mine4[k], 16);
        // ---- the recurrence over all eight steps (both lanes), keeping my four rows
        int rs[4], ry[4], rz[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int v = (j < 4) == (half == 0) ? mine4[j & 3] : other4[j & 3];
          update(v >> 5, v & 31);
          if ((j >> 2) == half) {  // (a select, not a branch: three predicated moves)
            rs[j & 3] = s.x;
            ry[j & 3] = s.y;
            rz[j & 3] = s.z;
          }
        }
        // ---- my four output rows
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int step_no = h.x + 4 * half + k + 1;
          emit(t + 4 * half + k, rs[k], ry[k], rz[k], 10 * ry[k] - rs[k],
               AR ? false : step_no == num_steps);
        }
        h.x += 8;  // env.py:252, eight times
        t += 8;
      }
    } else {
      top_up(t);
      top_up(t);
      cp_async_wait<RING / 8 - 2>();
      __syncwarp();
      const uint32_t x = words.word(p.seed, env_id, (uint32_t)h.y, (uint32_t)g, SC_STREAM_ORDER);
      const int D = __ldg(dsum + __umulhi(x, pow5));
      const int r = decode(act_ring[warp][t & (RING - 1)][el]);
      h.x += 1;
      update(r, D);
      const bool at_max = h.x == num_steps;
      const int k = 10 * s.y - s.x;  // reward numerator: before a reset clears the stock
      if (auto_reset && at_max) {    // Network.reset -> ShopAgent.reset clears the stock only
        s.x = 0;
        h.x = 0;
        h.y += 1;
      }
      if (half == 0) emit(t, s.x, s.y, s.z, k, at_max);
      t += 1;
    }
  }
  cp_async_wait<0>();

  bad_action = __shfl_xor_sync(0xFFFFFFFFu, (int)bad_action, 16) != 0 || bad_action;
  if (half == 0) {
    *reinterpret_cast<int2*>(a.hdr + e) = h;
    a.shop[e] = s;
    // first fault in event order: a bad action is detected in decode_action, before any send
    const uint32_t fault = bad_action ? (uint32_t)PHX_FAULT_INVALID_ACTION : p.fault[1];
    if (fault) raise_fault(a.faults, e, fault);
  }
}

// float32(n / den) through the kernel's own routine, for the exhaustive parity test.
__global__ void sc_ratio_selftest_kernel(int lo, int n, float den, float rcp, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sc_ratio(lo + i, den, rcp);
}

// PhantomEnv.reset (env.py:185-237) for the masked envs.
__global__ void sc_reset_kernel(ScPlan p, int4* hdr, int4* shop, uint32_t* term, uint32_t* trunc,
                                const uint8_t* env_mask, float* obs, uint8_t* obs_mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.E) return;
  if (env_mask && env_mask[e] == 0) return;
  int4 h = hdr[e];
  int4 s = shop[e];
  h.x = 0;
  h.y += 1;
  s.x = 0;  // sales / missed_sales survive the reset (supply_chain.py:149-150)
  hdr[e] = h;
  shop[e] = s;
  term[e] = 0;
  trunc[e] = 0;
  if (obs) {
    obs[e * 3 + 0] = sc_ratio(s.x, p.max_stock_f, p.rcp_stock);
    obs[e * 3 + 1] = sc_ratio(s.y, p.cap_f, p.rcp_cap);
    obs[e * 3 + 2] = sc_ratio(s.z, p.cap_f, p.rcp_cap);
  }
  if (obs_mask) obs_mask[e] = 1;
}

__global__ void sc_init_kernel(int E, int4* hdr, int4* shop) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  hdr[e] = make_int4(0, -1, 0, 0);  // episode becomes 0 on the first reset
  shop[e] = make_int4(0, 0, 0, 0);
}

#ifndef PHX_JIT_TU  // host side of the fast kernel: not part of a specialised engine unit
class SupplyChainFast final : public Family {
 public:
  ~SupplyChainFast() override {
    cudaFree(d_shop);
    cudaFree(d_scratch);
    cudaFree(d_dsum);
    if (h_wire) cudaFreeHost(h_wire);
  }

  int32_t init(const phx_spec& s) override {
    PHX_REQUIRE(s.env_kind == PHX_ENV_BASE, PHX_ERR_UNSUPPORTED,
                "supply-chain family runs under PhantomEnv (PHX_ENV_BASE) only");
    PHX_REQUIRE(s.n_payload_types == 4, PHX_ERR_INVALID,
                "supply-chain family expects 4 payload types "
                "(OrderRequest, OrderResponse, StockRequest, StockResponse)");
    PHX_REQUIRE(s.obs_dim == 3 && s.act_dim == 1, PHX_ERR_INVALID,
                "supply-chain family: obs_dim must be 3 and act_dim 1");
    PHX_REQUIRE(s.iparams[0] >= 1 && s.iparams[0] <= 255 && s.iparams[1] >= 0 &&
                    s.iparams[1] <= (1 << 20),
                PHX_ERR_INVALID, "supply-chain family: max_order / max_stock out of range");
    PHX_REQUIRE(is_canonical(s), PHX_ERR_UNSUPPORTED,
                "PHX_EXEC_FAST needs the canonical supply-chain layout "
                "[ShopAgent, FactoryAgent, CustomerAgent x N] with N <= 30, a plain Network and "
                "no shuffle_batches");
    PHX_CUDA(cudaMalloc(&d_shop, sizeof(int4) * (size_t)E));
    sc_init_kernel<<<(E + 255) / 256, 256>>>(E, d_hdr, d_shop);
    PHX_CUDA(cudaGetLastError());
    PHX_CUDA(cudaDeviceSynchronize());
    return make_plan(s);
  }

  static bool is_canonical(const phx_spec& s) {
    // the static schedule assumes one graph for all envs and push-order batches
    if (s.flags & (PHX_FLAG_STOCHASTIC_NETWORK | PHX_FLAG_SHUFFLE_BATCHES)) return false;
    if (s.n_agents < 3 || s.n_agents > 2 + SC_MAX_CUSTOMERS) return false;
    if (s.agent_kind[0] != SC_SHOP || s.agent_kind[1] != SC_FACTORY) return false;
    for (int i = 2; i < s.n_agents; ++i)
      if (s.agent_kind[i] != SC_CUSTOMER) return false;
    return s.n_strategic == 1 && s.strategic_index[0] == 0;
  }

  // network.py:246-254 evaluated on the static graph: 0 = pushed, else the fault raised.
  uint32_t send_check(const phx_spec& s, int from, int to, int type) const {
    const bool edge = mask_bit(s.adjacency[from], to);
    if (!(s.flags & PHX_FLAG_IGNORE_CONNECTION_ERRORS) && !edge) return PHX_FAULT_NO_EDGE;
    if (!(s.flags & PHX_FLAG_NO_PAYLOAD_CHECKS)) {
      if (!mask_bit(s.type_sender_ok[type], from) || !mask_bit(s.type_receiver_ok[type], to))
        return PHX_FAULT_BAD_PAYLOAD_TYPE;
    }
    return 0;
  }

  // Walks one step's event order (SURVEY.md A.1 rules 2-7) on the static graph and records
  // which messages are pushed / delivered and the first fault, once with and once without a
  // shop action.  The two walks must agree on the customer side (they do: it does not depend
  // on the shop's action), which is what makes a single static schedule valid.
  int32_t make_plan(const phx_spec& s) {
    ScPlan& p = plan;
    std::memset(&p, 0, sizeof(p));
    p.E = E;
    p.num_steps = s.num_steps;
    p.nc = s.n_agents - 2;
    p.max_order = s.iparams[0];
    p.max_stock = s.iparams[1];
    p.digits_per_word = rng_digits_per_word((uint32_t)p.max_order);
    p.words_per_step = (p.nc + p.digits_per_word - 1) / p.digits_per_word;
    p.flags = s.flags;
    p.seed = seed;
    p.env_offset = (uint32_t)env_offset;
    p.max_stock_f = (float)p.max_stock;
    p.rcp_stock = 1.0f / p.max_stock_f;  // IEEE division: correctly rounded
    p.cap_f = (float)(p.nc * p.max_order);
    p.rcp_cap = 1.0f / p.cap_f;
    for (int has = 0; has < 2; ++has) {
      uint32_t fault = 0;
      bool push_req = false, deliver_req = false;
      uint32_t push_ord = 0, deliver_ord = 0;
      // acting phase
      if (has) {
        fault = send_check(s, 0, 1, SC_STOCK_REQUEST);
        if (!fault) { push_req = true; deliver_req = mask_bit(s.adjacency[0], 1); }
      }
      for (int i = 0; i < p.nc && !fault; ++i) {
        fault = send_check(s, 2 + i, 0, SC_ORDER_REQUEST);
        if (!fault) {
          push_ord |= 1u << i;
          if (mask_bit(s.adjacency[2 + i], 0)) deliver_ord |= 1u << i;
        }
      }
      const bool any0 = push_req || push_ord;
      bool push_resp = false, deliver_resp = false;
      uint32_t push_ordresp = 0;
      if (!fault && any0 && s.round_limit == 0) fault = PHX_FAULT_ROUND_LIMIT;
      if (!fault && any0) {
        // round 0: receivers in first-arrival order = WAREHOUSE (if asked), SHOP
        if (deliver_req) {
          fault = send_check(s, 1, 0, SC_STOCK_RESPONSE);
          if (!fault) { push_resp = true; deliver_resp = mask_bit(s.adjacency[1], 0); }
        }
        for (int i = 0; i < p.nc && !fault; ++i) {
          if (!((deliver_ord >> i) & 1u)) continue;
          fault = send_check(s, 0, 2 + i, SC_ORDER_RESPONSE);
          if (!fault) push_ordresp |= 1u << i;
        }
        const bool any1 = push_resp || push_ordresp;
        if (!fault && any1 && s.round_limit == 1) fault = PHX_FAULT_ROUND_LIMIT;
        // round 1 produces no further messages (both handlers return nothing)
      }
      p.fault[has] = fault;
      if (has) {
        p.push_req = push_req;
        p.push_resp = push_resp;
        p.delivery_ok = (deliver_req && deliver_resp) ? 1u : 0u;
      }
      p.push_ord = push_ord;
      p.deliver_ord = deliver_ord;
      p.push_ordresp = push_ordresp;
    }
    if (tracking())
      PHX_REQUIRE(s.trace_capacity >= 2 * (1 + p.nc), PHX_ERR_INVALID,
                  "trace_capacity must be >= 2 * (1 + n_customers) for the supply chain");
    // sc_fast2_kernel: digit sums of the 5-digit base-max_order numbers (the order total of a step)
    pow5 = 0;
    const uint64_t n = (uint64_t)p.max_order;
    if (p.nc == 5 && p.words_per_step == 1 && n * n * n * n * n <= (uint64_t)SC2_MAX_TABLE) {
      pow5 = (uint32_t)(n * n * n * n * n);
      std::vector<uint8_t> tab((pow5 + 15u) & ~15u, 0);
      for (uint32_t v = 0; v < pow5; ++v) {
        uint32_t x = v, sum = 0;
        for (int k = 0; k < 5; ++k) { sum += x % (uint32_t)n; x /= (uint32_t)n; }
        tab[v] = (uint8_t)sum;
      }
      PHX_CUDA(cudaMalloc(&d_dsum, tab.size()));
      PHX_CUDA(cudaMemcpy(d_dsum, tab.data(), tab.size(), cudaMemcpyHostToDevice));
    }
    return PHX_OK;
  }

  int32_t reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                cudaStream_t stream) override {
    sc_reset_kernel<<<(E + 255) / 256, 256, 0, stream>>>(plan, d_hdr, d_shop, d_term, d_trunc,
                                                        env_mask, obs, obs_mask);
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  int32_t rollout(int32_t T, const StepIO& io, cudaStream_t stream) override {
    return rollout_range(T, io, 0, E, stream);
  }

  // Steps envs [env_begin, env_begin + env_count); the I/O planes keep the handle's full
  // [T, E, ...] shape (row stride E).
  int32_t rollout_range(int32_t T, const StepIO& io, int32_t env_begin, int32_t env_count,
                        cudaStream_t stream, uint32_t* wire = nullptr,
                        uint32_t* wire_overflow = nullptr) {
    PHX_REQUIRE(env_begin >= 0 && env_count >= 1 && env_begin + env_count <= E, PHX_ERR_INVALID,
                "env range out of bounds");
    ScArgs a;
    a.p = plan;
    a.T = T;
    a.env_begin = env_begin;
    a.env_count = env_count;
    // the vector action copy reads 4 consecutive envs of one step with one 16-byte cp.async; every
    // warp of every block must be a full warp of real envs (a clamped warp would compute an
    // unaligned source address), hence env_count % SC_BLOCK and not % 32
    a.vec_actions = (E % 4 == 0) && (env_begin % 4 == 0) && (env_count % SC_BLOCK == 0) &&
                    (reinterpret_cast<uintptr_t>(io.actions) % 16 == 0);
    a.hdr = d_hdr;
    a.shop = d_shop;
    a.io = io;
    a.faults = fault_sink();
    a.trace = trace_sink();
    const int grid = (env_count + SC_BLOCK - 1) / SC_BLOCK;
    const bool track = tracking();
    if (track) {
      PHX_REQUIRE(env_begin == 0 && env_count == E, PHX_ERR_INVALID,
                  "message tracking steps the whole handle");
      const int32_t rc = ensure_trace(T);
      if (rc != PHX_OK) return rc;
      a.trace = trace_sink();
    }
    PHX_REQUIRE((uint64_t)T * (uint64_t)E * 3ull < (1ull << 32), PHX_ERR_INVALID,
                "T * num_envs too large for one launch (row index is 32-bit): split the rollout");
    // Two output layouts are compiled: "lean" = obs + reward + all_done, "full" = all seven
    // planes.  Any other combination of NULLs runs "full" into scratch planes.
    const bool lean = a.io.obs && a.io.reward && a.io.all_done && !a.io.obs_mask &&
                      !a.io.reward_mask && !a.io.term && !a.io.trunc;
    if (!lean) {
      const size_t n = (size_t)T * E;
      const size_t need = n * 12 + n * 4 + n * 6;
      if (need > scratch_bytes) {
        PHX_CUDA(cudaStreamSynchronize(stream));
        if (d_scratch) PHX_CUDA(cudaFree(d_scratch));
        d_scratch = nullptr;
        scratch_bytes = 0;
        PHX_CUDA(cudaMalloc(&d_scratch, need));
        scratch_bytes = need;
      }
      uint8_t* q = (uint8_t*)d_scratch;
      if (!a.io.obs) a.io.obs = (float*)q;
      q += n * 12;
      if (!a.io.reward) a.io.reward = (float*)q;
      q += n * 4;
      if (!a.io.obs_mask) a.io.obs_mask = q;
      q += n;
      if (!a.io.reward_mask) a.io.reward_mask = q;
      q += n;
      if (!a.io.term) a.io.term = q;
      q += n;
      if (!a.io.trunc) a.io.trunc = q;
      q += n;
      if (!a.io.all_done) a.io.all_done = q;
    }
    const bool mask = a.io.action_mask != nullptr;
    const bool all_delivered_ = plan.deliver_ord == (1u << plan.nc) - 1u;
    PHX_REQUIRE(wire == nullptr || wire_ok(lean && !mask), PHX_ERR_INVALID,
                "compact wire output is not available for this launch");
    // (short launches -- phx_step is T = 1 -- stay on sc_fast_kernel: the round-2 kernels fill a
    // 3 KB table per block first, which costs a single-step launch 0.9 us: 2.95 vs 2.03 us)
    if (pow5 != 0 && !mask && !track && all_delivered_ && plan.delivery_ok && !use_v1 &&
        (T >= 4 || wire)) {
      // the round-2 kernels
      Sc2Args b;
      b.a = a;
      b.dsum = d_dsum;
      b.pow5 = pow5;
      b.pdl_lead = use_pdl ? pdl_lead : -(1 << 30);
      if (sc_variant == 4 && a.vec_actions && env_count % 32 == 0 && !wire) {  // two lanes per env
        const int grid4 = env_count / (SC4_THREADS / 2);
        const bool ar = (plan.flags & PHX_FLAG_AUTO_RESET) != 0;
        void (*k4)(const Sc2Args) =
            lean ? (ar ? (sc_ring == 32 ? sc_fast4_kernel<false, true, 32> : sc_fast4_kernel<false, true, 64>)
                       : sc_fast4_kernel<false, false, 64>)
                 : (ar ? sc_fast4_kernel<true, true, 64> : sc_fast4_kernel<true, false, 64>);
        k4<<<grid4, SC4_THREADS, 0, stream>>>(b);
        PHX_CUDA(cudaGetLastError());
        return PHX_OK;
      }
      if (sc_variant == 3 && plan.num_steps >= 4) {  // the time-parallel kernel
        const int grid3 = (env_count + SC3_ENVS - 1) / SC3_ENVS;
        const bool w2 = sc_warps == 2;
        b.wire = wire;
        b.wire_overflow = wire_overflow;
        if (wire) {  // host-buffer path: float planes (fallback) + the compact wire plane
          if (!w2) sc_fast3_kernel<false, 4, true><<<grid3, SC3_ENVS * 4, 0, stream>>>(b);
          else sc_fast3_kernel<false, 2, true><<<grid3, SC3_ENVS * 2, 0, stream>>>(b);
        } else if (lean && !w2) sc_fast3_kernel<false, 4, false><<<grid3, SC3_ENVS * 4, 0, stream>>>(b);
        else if (lean) sc_fast3_kernel<false, 2, false><<<grid3, SC3_ENVS * 2, 0, stream>>>(b);
        else if (!w2) sc_fast3_kernel<true, 4, false><<<grid3, SC3_ENVS * 4, 0, stream>>>(b);
        else sc_fast3_kernel<true, 2, false><<<grid3, SC3_ENVS * 2, 0, stream>>>(b);
        PHX_CUDA(cudaGetLastError());
        return PHX_OK;
      }
      b.wire = wire;
      b.wire_overflow = wire_overflow;
      void (*kern)(const Sc2Args) =
          wire ? (a.vec_actions ? sc_fast2_kernel<false, true, true> : sc_fast2_kernel<false, false, true>)
          : lean ? (a.vec_actions ? sc_fast2_kernel<false, true, false> : sc_fast2_kernel<false, false, false>)
                 : (a.vec_actions ? sc_fast2_kernel<true, true, false> : sc_fast2_kernel<true, false, false>);
      int ring = SC2_RING;
      if (sc_ring == 32 && lean && !wire && a.vec_actions) {  // A/B: a deeper action ring
        kern = sc_fast2_kernel<false, true, false, 32>;
        ring = 32;
      }
#if SC2_AR_INSTANCE
      else if (lean && !wire && a.vec_actions && (plan.flags & PHX_FLAG_AUTO_RESET)) {
        kern = sc_fast2_kernel<false, true, false, SC2_RING, true>;  // auto-reset as a constant
      }
#endif
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(SC_BLOCK);
      cfg.dynamicSmemBytes = sizeof(float) * ring * SC_BLOCK + ((pow5 + 15u) & ~15u);
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = use_pdl ? 1 : 0;
      PHX_CUDA(cudaLaunchKernelEx(&cfg, kern, b));
      PHX_CUDA(cudaGetLastError());
      return PHX_OK;
    }
#define SC_LAUNCH(NC_, TRACK_)                                                              \
  do {                                                                                      \
    if (mask && !lean) sc_fast_kernel<NC_, TRACK_, true, true><<<grid, SC_BLOCK, 0, stream>>>(a);  \
    else if (mask) sc_fast_kernel<NC_, TRACK_, true, false><<<grid, SC_BLOCK, 0, stream>>>(a);     \
    else if (!lean) sc_fast_kernel<NC_, TRACK_, false, true><<<grid, SC_BLOCK, 0, stream>>>(a);    \
    else sc_fast_kernel<NC_, TRACK_, false, false><<<grid, SC_BLOCK, 0, stream>>>(a);              \
  } while (0)
    const bool all_delivered = plan.deliver_ord == (1u << plan.nc) - 1u;
    // the NC = 5 instantiation draws a step's orders from one word
    if (plan.nc == 5 && plan.words_per_step == 1 && (track || all_delivered)) {
      if (track) SC_LAUNCH(5, true);
      else SC_LAUNCH(5, false);
    } else {
      PHX_REQUIRE(!track, PHX_ERR_UNSUPPORTED,
                  "message tracking on the fast path is built for 5 customers; "
                  "use PHX_EXEC_QUEUE");
      SC_LAUNCH(0, false);
    }
#undef SC_LAUNCH
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  // The compact wire word is produced by sc_fast2_kernel / sc_fast3_kernel, for the lean layout.
  bool wire_ok(bool lean_no_mask) const {
    const bool all_delivered_ = plan.deliver_ord == (1u << plan.nc) - 1u;
    return lean_no_mask && pow5 != 0 && !tracking() && all_delivered_ && plan.delivery_ok &&
           (sc_variant == 2 || (sc_variant == 3 && plan.num_steps >= 4)) &&
           plan.max_stock < SCW_STOCK_BIAS &&
           plan.nc * plan.max_order <= (int)SCW_FIELD_MAX && !no_wire;
  }

  // phx_rollout_host for the lean layout (obs + reward + all_done).  The float32 planes are
  // 18 B per env-step and the call is bound by what can be written into HOST memory: the DMA
  // engine of the device->host link on one side, the host's own cores on the other.  Both are
  // used: the first `wire_chunks` time chunks cross PCIe as ONE 32-bit word per env-step (4 B,
  // phx_sc_wire.h) and are expanded into the caller's planes by the host thread pool, while the
  // DMA engine copies the float planes of the later chunks straight into place.  A small
  // controller moves the split by one chunk per call towards the point where both finish
  // together (a 16-vCPU VM next to one GPU ends near the middle; eight ranks sharing one root
  // complex lean on the cores).  The expansion evaluates the same correctly rounded quotients as
  // the kernel, so every plane is bit-identical to the device path's
  // (tests/test_gpu_reset_and_io.py); the kernel writes the float planes of ALL rows into the
  // device staging block, and if a value did not fit its wire field (|stock| >= 2^15, sales or
  // missed outside 0..127: out-of-distribution actions) the wire rows are re-copied from there.
  static constexpr int NCH = 14;  // time chunks; boundaries in percent of T, finer at the front
  int32_t rollout_host(int32_t T, const StepIO& h) override {
    const bool lean = h.obs && h.reward && h.all_done && !h.obs_mask && !h.reward_mask &&
                      !h.term && !h.trunc && !h.action_mask;
    // Opt-in (PHX_WIRE_CHUNKS = k | auto, read when the handle is created): on the hosts measured
    // so far -- 16- and 32-vCPU VMs -- the cores write the float planes at ~30 GB/s in total,
    // slower than the link's DMA engine (~50 GB/s), so the default is the plain staged copy
    // (profiles/r02_e2e_wire_sweep.txt).
    if (wire_mode.empty() || wire_mode == "0" || !wire_ok(lean) || T < 2 * NCH ||
        (size_t)T * E < 65536 || (uint64_t)T * (uint64_t)E * 3ull >= (1ull << 32))
      return Family::rollout_host(T, h);
    static const int kFrac[NCH + 1] = {0, 2, 6, 12, 20, 28, 36, 44, 52, 60, 68, 76, 84, 92, 100};
    auto bound = [&](int c) { return ((size_t)T * kFrac[c] + 50) / 100; };
    const size_t TE = (size_t)T * E;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_act = up(TE * 4), b_obs = up(TE * 12), b_rew = up(TE * 4), b_all = up(TE * 2),
                 b_wire = up(TE * 4);
    int32_t rc = ensure_stage(b_act + b_obs + b_rew + b_all + b_wire + 256);
    if (rc != PHX_OK) return rc;
    if (TE * 4 > h_wire_bytes) {  // pinned landing buffer of the wire words
      if (h_wire) cudaFreeHost(h_wire);
      h_wire = nullptr;
      h_wire_bytes = 0;
      PHX_CUDA(cudaMallocHost(&h_wire, TE * 4));
      h_wire_bytes = TE * 4;
    }
    if (wire_chunks < 0) {  // first call: a fixed split, or a first guess for the controller
      wire_fixed = wire_mode != "auto";
      wire_chunks = wire_fixed ? std::atoi(wire_mode.c_str()) : std::min(NCH / 2, host_pool().size());
      wire_chunks = std::max(0, std::min(NCH, wire_chunks));
    }
    const int nw = wire_chunks;
    uint8_t* q = (uint8_t*)d_stage;
    float* d_act = (float*)q; q += b_act;
    float* d_obs = (float*)q; q += b_obs;
    float* d_rew = (float*)q; q += b_rew;
    uint8_t* d_all = q; q += b_all;
    uint32_t* d_wire = (uint32_t*)q; q += b_wire;
    uint32_t* d_over = (uint32_t*)q;
    PHX_CUDA(cudaMemsetAsync(d_over, 0, sizeof(uint32_t), own_stream));
    const ScWireParams wp{plan.max_stock, plan.nc * plan.max_order};
    auto planes_d2h = [&](size_t t0, size_t nt, cudaStream_t st) -> cudaError_t {
      cudaError_t e1 = cudaMemcpyAsync(h.obs + t0 * E * 3, d_obs + t0 * E * 3, nt * E * 12,
                                       cudaMemcpyDeviceToHost, st);
      if (e1 != cudaSuccess) return e1;
      e1 = cudaMemcpyAsync(h.reward + t0 * E, d_rew + t0 * E, nt * E * 4, cudaMemcpyDeviceToHost, st);
      if (e1 != cudaSuccess) return e1;
      return cudaMemcpyAsync(h.all_done + t0 * E * 2, d_all + t0 * E * 2, nt * E * 2,
                             cudaMemcpyDeviceToHost, st);
    };
    for (int c = 0; c < NCH; ++c) {
      const size_t t0 = bound(c), nt = bound(c + 1) - t0;
      if (nt == 0) continue;
      PHX_CUDA(cudaMemcpyAsync(d_act + t0 * E, h.actions + t0 * E, nt * E * 4,
                               cudaMemcpyHostToDevice, copy_in));
      PHX_CUDA(cudaEventRecord(ev_in[c], copy_in));
      PHX_CUDA(cudaStreamWaitEvent(own_stream, ev_in[c], 0));
      StepIO ic{d_act + t0 * E, nullptr, d_obs + t0 * E * 3, nullptr, d_rew + t0 * E, nullptr,
                nullptr, nullptr, d_all + t0 * E * 2};
      rc = rollout_range((int32_t)nt, ic, 0, E, own_stream, d_wire + t0 * E, d_over);
      if (rc != PHX_OK) return rc;
      PHX_CUDA(cudaEventRecord(ev_k[c], own_stream));
      PHX_CUDA(cudaStreamWaitEvent(copy_out, ev_k[c], 0));
      if (c < nw)
        PHX_CUDA(cudaMemcpyAsync(h_wire + t0 * E, d_wire + t0 * E, nt * E * 4,
                                 cudaMemcpyDeviceToHost, copy_out));
      else
        PHX_CUDA(planes_d2h(t0, nt, copy_out));
      PHX_CUDA(cudaEventRecord(ev_out[c], copy_out));
    }
    uint32_t over = 0;
    PHX_CUDA(cudaMemcpyAsync(&over, d_over, sizeof(over), cudaMemcpyDeviceToHost, copy_out));
    PHX_CUDA(cudaEventRecord(ev_out[NCH], copy_out));
    // expand the wire chunks as they land, while the DMA engine works on the plane chunks
    bool dma_done_early = false;  // everything had landed before the last expansion started
    for (int c = 0; c < nw; ++c) {
      const size_t t0 = bound(c), nt = bound(c + 1) - t0;
      if (nt == 0) continue;
      PHX_CUDA(cudaEventSynchronize(ev_out[c]));
      if (c == nw - 1) dma_done_early = cudaEventQuery(ev_out[NCH]) == cudaSuccess;
      sc_wire_expand(host_pool(), wp, h_wire + t0 * E, nt * E, h.obs + t0 * E * 3,
                     h.reward + t0 * E, h.all_done + t0 * E * 2);
    }
    const bool cpu_done_early = cudaEventQuery(ev_out[NCH]) != cudaSuccess;  // DMA still busy
    PHX_CUDA(cudaStreamSynchronize(copy_out));
    PHX_CUDA(cudaStreamSynchronize(own_stream));
    if (over && nw > 0) {  // a value did not fit the wire format: take the kernel's float planes
      PHX_CUDA(planes_d2h(0, bound(nw), copy_out));
      PHX_CUDA(cudaStreamSynchronize(copy_out));
    }
    if (!wire_fixed) {  // one chunk per call towards the balance point
      if (dma_done_early && nw > 0) wire_chunks = nw - 1;
      else if (cpu_done_early && nw < NCH) wire_chunks = nw + 1;
    }
    return PHX_OK;
  }

  int32_t family_field(int32_t field, int32_t, void** p, size_t* bytes) override {
    if (field == PHX_FIELD_FAMILY + 0) {
      *p = d_shop;
      *bytes = sizeof(int4) * (size_t)E;
      return PHX_OK;
    }
    set_error("supply-chain family: unknown field " + std::to_string(field));
    return PHX_ERR_INVALID;
  }

  const char* exec_name() const override { return "fast(thread-per-env)"; }

 private:
  ScPlan plan{};
  uint32_t* h_wire = nullptr;  // pinned landing buffer of the compact wire words (rollout_host)
  size_t h_wire_bytes = 0;
  int wire_chunks = -1;        // leading time chunks sent as wire words (rollout_host controller)
  bool wire_fixed = false;     // PHX_WIRE_CHUNKS = k pins the split, "auto" lets the controller move it
  const std::string wire_mode = std::getenv("PHX_WIRE_CHUNKS") ? std::getenv("PHX_WIRE_CHUNKS") : "";
  uint8_t* d_dsum = nullptr;  // digit-sum table of sc_fast2_kernel (nullptr: not applicable)
  uint32_t pow5 = 0;
  // PHX_SC_KERNEL = 1 | 2 | 3 picks sc_fast_kernel / sc_fast2_kernel / sc_fast3_kernel (default 2:
  // measured 30.7 us per 65 536 x 100 launch against 31.5 and 52.1, profiles/r02_ab_sc_kernels.txt)
  // where more than one is valid: A/B measurements and the cross-kernel parity tests
  const int sc_variant = std::getenv("PHX_SC_KERNEL") ? std::atoi(std::getenv("PHX_SC_KERNEL")) : 2;
  const bool use_v1 = sc_variant == 1;
  // sc_fast2_kernel is launched with programmatic stream serialization (PHX_PDL=0 turns it off):
  // a launch lets the NEXT launch on the stream start its blocks when it has PHX_PDL_LEAD (8)
  // steps left, so that the successor's launch latency, block scheduling and table fill overlap
  // this grid's tail; the successor waits at griddepcontrol.wait before it reads the env state.
  // Measured per 65 536 x 100 launch (profiles/r02_ab_sc_kernels.txt): 30.74 us without, 29.54
  // with a lead of 8 steps (4: 29.82, 12: 29.60, 20: 30.21).  A trigger at the very START of the
  // kernel was rejected earlier in the round (every queued launch becomes resident and spins).
  const bool use_pdl = !(std::getenv("PHX_PDL") != nullptr && std::getenv("PHX_PDL")[0] == '0');
  const int pdl_lead = std::getenv("PHX_PDL_LEAD") ? std::atoi(std::getenv("PHX_PDL_LEAD")) : 8;
  // PHX_NO_WIRE=1 (read when the handle is created) keeps phx_rollout_host on the float planes
  const bool no_wire = std::getenv("PHX_NO_WIRE") != nullptr && std::getenv("PHX_NO_WIRE")[0] == '1';
  const int sc_warps = std::getenv("PHX_SC_WARPS") ? std::atoi(std::getenv("PHX_SC_WARPS")) : 4;
  const int sc_ring = std::getenv("PHX_SC_RING") ? std::atoi(std::getenv("PHX_SC_RING")) : 16;
  int4* d_shop = nullptr;
  void* d_scratch = nullptr;  // planes the caller did not ask for (non-lean layouts)
  size_t scratch_bytes = 0;
};


#endif  // PHX_JIT_TU

// ---------------------------------------------------------------------------------------
// The same agents as a device program of the generic queue engine (phx_engine.cuh): any agent
// order and topology (one shop and one factory per env), messages routed dynamically.
//   state words (per slot; only the shop uses them): stock, sales, missed_sales, delivered
//   agent_iparam[slot][0] = slot of the peer the agent addresses (customer -> its shop,
//                           shop -> its factory);  [slot][1] = customer ordinal (RNG idx)
template <int SEGCAP_>
struct ScProgram {
  // run-time specialisation: where this program lives and what it is called
  static constexpr const char* JIT_SOURCE = "fam_supply_chain.cu";
  static constexpr const char* JIT_NAME = SEGCAP_ == 8 ? "ScProgram<8>" : "ScProgram<32>";
  // every agent sends at most one message in the acting phase; the shop answers every order
  static constexpr int PW = 1, NWORDS = 4, VW = 0, ACTCAP = 1, RESPCAP = SEGCAP_, OBS_DIM = 3,
                       ACT_DIM = 1;
  static constexpr int SEGCAP = SEGCAP_;
  static constexpr int Q1CAP = SEGCAP_;  // thread-per-env engine: messages in flight per round
  static constexpr int RECVCAP = SEGCAP_;  // max messages one agent receives in a round
  static constexpr bool BATCHED = false, HAS_PRE = true, HAS_POST = false;

  static int q1_cap(const phx_spec& s) { return s.n_agents; }  // <= one message per agent and round
  static int32_t validate(const phx_spec& s) {
    PHX_REQUIRE(s.env_kind == PHX_ENV_BASE, PHX_ERR_UNSUPPORTED,
                "supply-chain family runs under PhantomEnv (PHX_ENV_BASE) only");
    int shops = 0, customers = 0;
    for (int i = 0; i < s.n_agents; ++i) {
      shops += s.agent_kind[i] == SC_SHOP;
      customers += s.agent_kind[i] == SC_CUSTOMER;
      PHX_REQUIRE(s.agent_kind[i] >= SC_SHOP && s.agent_kind[i] <= SC_CUSTOMER, PHX_ERR_INVALID,
                  "unknown supply-chain agent kind");
    }
    PHX_REQUIRE(shops == 1, PHX_ERR_UNSUPPORTED, "exactly one ShopAgent per env");
    PHX_REQUIRE(customers + 1 <= SEGCAP, PHX_ERR_UNSUPPORTED, "too many customers for this queue");
    return PHX_OK;
  }

#ifndef PHX_JIT_TU
  // Static send signature (phx_engine_host.cuh build_static_plan): what act() / handle() below
  // may send, in emission order.
  static void act_sends(const phx_spec& s, int slot, int, std::vector<SendSig>& out) {
    if (s.agent_kind[slot] == SC_SHOP) out.push_back(SendSig{s.agent_iparam[slot][0], SC_STOCK_REQUEST});
    if (s.agent_kind[slot] == SC_CUSTOMER) out.push_back(SendSig{s.agent_iparam[slot][0], SC_ORDER_REQUEST});
  }
  static void handle_sends(const phx_spec& s, int slot, int type, int sender,
                           std::vector<SendSig>& out) {
    if (s.agent_kind[slot] == SC_SHOP && type == SC_ORDER_REQUEST)
      out.push_back(SendSig{sender, SC_ORDER_RESPONSE});
    if (s.agent_kind[slot] == SC_FACTORY && type == SC_STOCK_REQUEST)
      out.push_back(SendSig{sender, SC_STOCK_RESPONSE});
  }
#endif

  template <class E>
  __device__ static void act(const Ctx& c, int* st, bool has_action, const float* action, E& out) {
    const EngineSpec& sp = *c.spec;
    if (c.kind == SC_SHOP) {
      if (!has_action) return;  // Agent.generate_messages default: [] (agents.py:157-158)
      const float a0 = action[0];
      if (!(fabsf(a0) <= SC_MAX_ABS_ACTION)) {
        out.fault = PHX_FAULT_INVALID_ACTION;
        return;
      }
      const int ask = min(__float2int_rn(a0), sp.iparams[1] - st[0]);  // supply_chain.py:136-142
      out.send(sp.agent_iparam[c.slot][0], SC_STOCK_REQUEST, ask);
    } else if (c.kind == SC_CUSTOMER) {  // supply_chain.py:61-67
      const int want = c.packed_randint(SC_STREAM_ORDER, (uint32_t)sp.iparams[0],
                                        (uint32_t)sp.iparams[2], (uint32_t)sp.iparams[3],
                                        (uint32_t)sp.agent_iparam[c.slot][1]);
      out.send(sp.agent_iparam[c.slot][0], SC_ORDER_REQUEST, want);
    }
  }

  __device__ static void view(const Ctx&, const int*, int*) {}

  __device__ static void pre(const Ctx& c, int* st) {
    if (c.kind == SC_SHOP) st[1] = st[2] = 0;  // supply_chain.py:93-96
  }
  __device__ static void post(const Ctx&, int*) {}

  template <class E>
  __device__ static bool handle(const Ctx& c, int* st, const Msg& m, E& out) {
    if (c.kind == SC_SHOP) {
      if (m.type == SC_STOCK_RESPONSE) {  // supply_chain.py:98-102
        st[3] = m.p[0];
        st[0] = min(st[0] + m.p[0], c.spec->iparams[1]);
        return true;
      }
      if (m.type == SC_ORDER_REQUEST) {  // supply_chain.py:104-122
        const int sold = min(m.p[0], st[0]);
        st[2] += m.p[0] - sold;
        st[0] -= sold;
        st[1] += sold;
        out.send(m.sender, SC_ORDER_RESPONSE, sold);
        return true;
      }
      return false;
    }
    if (c.kind == SC_FACTORY) {  // supply_chain.py:40-45
      if (m.type != SC_STOCK_REQUEST) return false;
      out.send(m.sender, SC_STOCK_RESPONSE, m.p[0]);
      return true;
    }
    return m.type == SC_ORDER_RESPONSE;  // CustomerAgent.handle_order_response: no-op
  }

  __device__ static bool encode(const Ctx& c, int* st, float* obs) {
    const EngineSpec& sp = *c.spec;
    obs[0] = sc_ratio(st[0], sp.fparams[0], sp.fparams[1]);
    obs[1] = sc_ratio(st[1], sp.fparams[2], sp.fparams[3]);
    obs[2] = sc_ratio(st[2], sp.fparams[2], sp.fparams[3]);
    return true;
  }
  __device__ static float reward(const Ctx&, int* st) {
    return sc_ratio(10 * st[1] - st[0], 10.0f, 0.1f);
  }
  __device__ static bool terminated(const Ctx&, const int*) { return false; }
  __device__ static bool truncated(const Ctx&, const int*) { return false; }
  __device__ static void reset_agent(const Ctx& c, int* st) {
    if (c.kind == SC_SHOP) st[0] = 0;  // supply_chain.py:149-150
  }
};

#ifndef PHX_JIT_TU
template <int SEGCAP_>
class SupplyChainQueue final : public EngineFamily<ScProgram<SEGCAP_>> {
 public:
  int32_t init(const phx_spec& s) override {
    int customers = 0;
    for (int i = 0; i < s.n_agents; ++i) customers += s.agent_kind[i] == SC_CUSTOMER;
    PHX_REQUIRE(s.obs_dim == 3 && s.act_dim == 1 && s.n_payload_types == 4, PHX_ERR_INVALID,
                "supply-chain family: obs_dim 3, act_dim 1, 4 payload types");
    PHX_REQUIRE(s.iparams[0] >= 1 && s.iparams[0] <= 255 && s.iparams[1] >= 1 &&
                    s.iparams[1] <= (1 << 20) && customers >= 1,
                PHX_ERR_INVALID, "supply-chain family: parameters out of range");
    phx_spec t = s;  // obs denominators and their correctly rounded reciprocals
    // packed order-size draws (ScProgram::act): digits per word, words per step
    t.iparams[2] = rng_digits_per_word((uint32_t)s.iparams[0]);
    t.iparams[3] = (customers + t.iparams[2] - 1) / t.iparams[2];
    t.fparams[0] = (double)s.iparams[1];
    t.fparams[1] = (double)(1.0f / (float)s.iparams[1]);
    t.fparams[2] = (double)(customers * s.iparams[0]);
    t.fparams[3] = (double)(1.0f / (float)(customers * s.iparams[0]));
    return EngineFamily<ScProgram<SEGCAP_>>::init(t);
  }
};

#endif  // PHX_JIT_TU

}  // namespace

#ifndef PHX_JIT_TU
Family* make_supply_chain_family(const phx_spec& s) {
  const bool canonical = SupplyChainFast::is_canonical(s);
  if (s.exec_mode == PHX_EXEC_FAST || (s.exec_mode == PHX_EXEC_AUTO && canonical))
    return new SupplyChainFast();
  int customers = 0;
  for (int i = 0; i < s.n_agents; ++i) customers += s.agent_kind[i] == SC_CUSTOMER;
  if (customers + 1 <= 8) return new SupplyChainQueue<8>();
  return new SupplyChainQueue<32>();
}

int32_t selftest_ratio(int32_t device, int32_t den, int32_t lo, int32_t count, float* host_out) {
  PHX_REQUIRE(den > 0 && count > 0 && host_out != nullptr, PHX_ERR_INVALID, "bad arguments");
  PHX_CUDA(cudaSetDevice(device));
  float* d = nullptr;
  PHX_CUDA(cudaMalloc(&d, sizeof(float) * (size_t)count));
  const float den_f = (float)den;
  sc_ratio_selftest_kernel<<<(count + 255) / 256, 256>>>(lo, count, den_f, 1.0f / den_f, d);
  cudaError_t err = cudaMemcpy(host_out, d, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost);
  cudaFree(d);
  PHX_CUDA(err);
  return PHX_OK;
}

#endif  // PHX_JIT_TU

}  // namespace phx
