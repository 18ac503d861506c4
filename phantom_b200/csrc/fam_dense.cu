// fam_dense.cu -- device program family PHX_FAMILY_DENSE (BASELINE config C5): the dense-graph
// broadcast / batch-aggregation env of oracle/workloads/dense.py.
//
// Replaces, for up to 128 agents per env, the message-queue stress path of the reference:
//   env.py:320-336         acting phase: every agent sends to every neighbour (N (N-1) sends)
//   network.py:233-254     Network.send checks
//   resolvers.py:128-163   BatchResolver.resolve with round_limit = 2
//   agents.py:96-120       Agent.handle_batch, OVERRIDDEN by the agents to aggregate the batch
//                          (sum, max, first arg-max; cf. the auction of
//                          examples/environments/digital_ads_market/digital_ads_market.py:429-510)
//
// Mapping: ONE THREAD BLOCK PER ENV, thread == agent slot (the north star's layout; it pays
// here because an env has 128 agents and 16 256 messages per step).  On this family's message
// graph the queue never has to be materialised: a Signal's payload is a function of the
// sender's value and the (sender, receiver) pair, so round 0 is a PULL -- every sender
// publishes one word in shared memory (vals[sender], 512 B per env instead of a 66 KB
// mailbox), and every receiver walks its senders in slot order (= the reference's batch order
// for a receiver: global push order) reading vals[s] as a warp-wide broadcast and re-deriving
// the receiver-tailored part from a 5-entry register table.  Round 1 is a PUSH: an Ack is two
// shared-memory integer atomics on the receiver's counters (integer sums are order
// independent).  Measured: 0.62 ms per 16 384 x 8 launch against 2.15 ms for the mailbox
// version (which spent ~15 instructions per message, most of them on the 16 256 mailbox
// writes and the 128-deep Ack scan of every agent).  The receivers' first-arrival order only matters for the order in which the Acks
// are pushed, i.e. for the message trace, and is reconstructed there.
// On the benchmark's graph -- complete, every agent acting -- round 0 does not even need the
// pull: handle_batch's aggregation (sum, max, FIRST arg-max) over "everybody but me" is a
// BLOCK REDUCTION.  A Signal is value[s] + ((2 s) % 5 + (3 r) % 5) % 5, i.e. the sender's value
// plus a term that depends on the sender only through its class c = (2 s) % 5; so the max over
// senders is the max over the five classes of (class maximum + tail_r[c]), the first arg-max is
// the lowest slot among the classes' first arg-maxes, and "but me" is handled by keeping the
// best TWO senders of every class.  Eleven warp reductions (REDUX.MAX / REDUX.ADD) per warp and
// a 4-way merge per class replace the 127-message scan of every receiver (~520 instructions per
// agent and step -> ~200); any other graph / partial action set takes the pull path, in the
// same kernel, per env and step (the choice is block-uniform).  This is the north star's
// "warp-ballot / shfl reductions for BatchResolver aggregation".  A step has three block
// barriers: after the publish, after the Acks, and (shared with the next step's first one)
// before thread 0 issues the TMA stores of the staged rows.
// Outputs of a step are staged in shared memory and written with TMA bulk stores
// (cp.async.bulk.global.shared::cta, SASS UBLKCP): one env's rows are contiguous in every
// [T,E,S,...] plane.
//
// Agent kind 0 = DenseAgent (strategic).  Payload types: 0 Signal(value), 1 Ack(value).
// State words per agent: 0 signal, 1 total, 2 best, 3 best_sender, 4 acks, 5 ack_total.
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>

#include "phx_family.h"

namespace phx {
namespace {

constexpr int DN_MAX = 128;       // agents per env == threads per block
constexpr int DN_WORDS = 6;
enum { DN_SIGNAL = 0, DN_ACK = 1 };

struct DenseSpec {
  int32_t E, n, num_steps, round_limit;
  uint32_t flags;
  int32_t complete;  // the graph is complete (every agent is connected to every other agent)
  uint32_t sender_ok[2][4], receiver_ok[2][4];
};

struct DenseArgs {
  DenseSpec sp;
  int32_t T;
  int4* hdr;
  int32_t* state;        // [DN_WORDS][E][128]
  const uint32_t* adj;   // [128][4] adjacency rows (bit r of row s: edge s -> r)
  StepIO io;
  FaultSink faults;
  TraceSink trace;
};

struct alignas(16) DenseStage {  // one step's output rows of one env, 16-byte aligned planes
  float obs[DN_MAX * 3];
  float reward[DN_MAX];
  uint8_t obs_mask[DN_MAX], reward_mask[DN_MAX], term[DN_MAX], trunc[DN_MAX];
};

struct DenseSmem {
  alignas(16) int32_t vals[DN_MAX];  // round 0: the Signal value every sender published
  int32_t ack_cnt[DN_MAX];         // round 1: Acks received / their sum, by receiver
  int32_t ack_sum[DN_MAX];
  int2 ack[DN_MAX];                // trace: (Ack receiver of every agent or -1, Ack value)
  int32_t order_key[DN_MAX];       // trace only: first-arrival keys
  uint32_t sent[4];
  uint32_t any_ack;
  // reduction form of round 0 (complete graph, everybody sent): per sender class c = (2 s) % 5
  // the best and second-best sender as keys value * 128 + (127 - slot), and the sum of all values
  alignas(16) int2 wtop[5][4];   // [class][warp] (best, second best) keys of the warp's senders
  alignas(16) int32_t wsum[4];   // [warp] sum of the warp's senders' values
  DenseStage stage[2];
};

__device__ __forceinline__ int dn_tailored(int value, int s, int r) {
  return value + (7 * s + 3 * r) % 5;
}

template <bool TRACK>
__global__ void __launch_bounds__(DN_MAX) dense_step_kernel(const DenseArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DenseSmem& sm = *reinterpret_cast<DenseSmem*>(smem_raw);
  const DenseSpec& sp = a.sp;
  const int e = blockIdx.x;
  const int slot = threadIdx.x;
  const int lane = slot & 31, warp = slot >> 5;
  const int n = sp.n;
  const bool is_agent = slot < n;

  uint32_t adj[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) adj[w] = is_agent ? a.adj[slot * 4 + w] : 0u;
  int4 h = a.hdr[e];
  int st[DN_WORDS];
#pragma unroll
  for (int w = 0; w < DN_WORDS; ++w) st[w] = a.state[((size_t)w * sp.E + e) * DN_MAX + slot];
  uint32_t fault = 0;
  // Signal(value + (7 s + 3 r) % 5): (7 s + 3 r) % 5 == ((2 s) % 5 + (3 r) % 5) % 5; the sender
  // part is a compile-time constant of the unrolled sender loop, the receiver part selects one
  // of five per-lane registers
  int tailtab[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) tailtab[k] = (k + 3 * slot) % 5;
  // reduction form of round 0: this sender's class, the class populations and "all n agents"
  const int myclass = (2 * slot) % 5;
  int tails_all = 0;  // sum over ALL senders s of tailtab[(2 s) % 5]
  uint32_t full_w[4];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const int s0 = (3 * c) % 5;  // (2 s) % 5 == c  <=>  s % 5 == (3 c) % 5
    tails_all += (s0 < n ? (n - 1 - s0) / 5 + 1 : 0) * tailtab[c];
  }
#pragma unroll
  for (int w = 0; w < 4; ++w)
    full_w[w] = n >= 32 * (w + 1) ? 0xFFFFFFFFu : (n > 32 * w ? (1u << (n - 32 * w)) - 1u : 0u);
  const bool bulk = (n % 16) == 0;  // plane sizes multiples of 16 bytes
  const bool ok_send_signal = ((sp.sender_ok[DN_SIGNAL][slot >> 5] >> (slot & 31)) & 1u) ||
                              (sp.flags & PHX_FLAG_NO_PAYLOAD_CHECKS);
  const bool ok_send_ack = ((sp.sender_ok[DN_ACK][slot >> 5] >> (slot & 31)) & 1u) ||
                           (sp.flags & PHX_FLAG_NO_PAYLOAD_CHECKS);
  // loop invariants of the acting phase: does this agent have neighbours, and may all of them
  // receive a Signal (payload whitelist of the receivers, network.py:323-331)?
  const bool has_neighbours = (adj[0] | adj[1] | adj[2] | adj[3]) != 0u;
  const bool signal_recv_ok =
      (sp.flags & PHX_FLAG_NO_PAYLOAD_CHECKS) ||
      !((adj[0] & ~sp.receiver_ok[DN_SIGNAL][0]) | (adj[1] & ~sp.receiver_ok[DN_SIGNAL][1]) |
        (adj[2] & ~sp.receiver_ok[DN_SIGNAL][2]) | (adj[3] & ~sp.receiver_ok[DN_SIGNAL][3]));
  // every agent may receive an Ack (the usual case): no per-step lookup of the arg-max sender's bit
  const bool ack_recv_all_ok =
      (sp.flags & PHX_FLAG_NO_PAYLOAD_CHECKS) ||
      ((sp.receiver_ok[DN_ACK][0] & full_w[0]) == full_w[0] && (sp.receiver_ok[DN_ACK][1] & full_w[1]) == full_w[1] &&
       (sp.receiver_ok[DN_ACK][2] & full_w[2]) == full_w[2] && (sp.receiver_ok[DN_ACK][3] & full_w[3]) == full_w[3]);
  if (bulk) {  // the four mask planes of this env class are constant: staged once
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      sm.stage[b].obs_mask[slot] = 1;
      sm.stage[b].reward_mask[slot] = 1;
      sm.stage[b].term[slot] = 0;
      sm.stage[b].trunc[slot] = 0;
    }
  }
  // TMA bulk stores of one step's staged rows (issued by thread 0 one barrier after the rows
  // were written, so that no barrier exists for the stores alone)
  auto issue_stores = [&](int t) {
    const DenseStage& sg = sm.stage[t & 1];
    const size_t base = ((size_t)t * sp.E + e) * n;
    if (a.io.obs) bulk_store(a.io.obs + base * 3, sg.obs, (uint32_t)n * 12u);
    if (a.io.reward) bulk_store(a.io.reward + base, sg.reward, (uint32_t)n * 4u);
    if (a.io.obs_mask) bulk_store(a.io.obs_mask + base, sg.obs_mask, (uint32_t)n);
    if (a.io.reward_mask) bulk_store(a.io.reward_mask + base, sg.reward_mask, (uint32_t)n);
    if (a.io.term) bulk_store(a.io.term + base, sg.term, (uint32_t)n);
    if (a.io.trunc) bulk_store(a.io.trunc + base, sg.trunc, (uint32_t)n);
    bulk_commit();
  };

  for (int t = 0; t < a.T; ++t) {
    const size_t row = (size_t)t * sp.E + e;
    DenseStage& sg = sm.stage[t & 1];
    h.x += 1;  // env.py:252

    // ---- acting phase: decode_action -> Signal to every neighbour, in slot order
    bool has = false;
    float act = 0.f;
    if (is_agent) {
      has = a.io.action_mask ? a.io.action_mask[row * n + slot] != 0 : true;
      act = a.io.actions[row * n + slot];
    }
    bool sends = false;
    if (has) {
      if (!(fabsf(act) <= 1048576.0f)) fault = fault ? fault : PHX_FAULT_INVALID_ACTION;
      st[0] = max(0, min(1000, __float2int_rn(__fmul_rn(act, 1000.0f))));
      sends = has_neighbours;
      if (sends && !(ok_send_signal && signal_recv_ok)) fault = fault ? fault : PHX_FAULT_BAD_PAYLOAD_TYPE;
    }
    const uint32_t sent_w = __ballot_sync(0xFFFFFFFFu, sends);
    if (lane == 0) sm.sent[warp] = sent_w;
    if (sends) sm.vals[slot] = st[0];  // one word per sender; the receivers of the pull form read it
    // reduce form, this warp's part: the sum of its senders' values and, per sender class, its
    // best two senders as keys value * 128 + (127 - slot) (a higher key = a higher value, then
    // the LOWER slot: the first sender attaining the maximum wins, as in the reference's scan)
    if (!TRACK && sp.complete) {
      const int key = sends ? st[0] * 128 + (127 - slot) : -1;
      const int wsum = __reduce_add_sync(0xFFFFFFFFu, sends ? st[0] : 0);
      int2 mine = make_int2(-1, -1);
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int kc = myclass == c ? key : -1;
        const int t1 = __reduce_max_sync(0xFFFFFFFFu, kc);
        const int t2 = __reduce_max_sync(0xFFFFFFFFu, kc == t1 ? -1 : kc);
        if (lane == c) mine = make_int2(t1, t2);
      }
      if (lane < 5) sm.wtop[lane][warp] = mine;
      if (lane == 5) sm.wsum[warp] = wsum;
    }
    // ---- pre_message_resolution
    st[1] = 0; st[2] = 0; st[3] = -1; st[4] = 0; st[5] = 0;
    sm.ack_cnt[slot] = 0;
    sm.ack_sum[slot] = 0;
    // the staged rows of step t - 2 (same buffer as this step's) must have been read out
    if (bulk && slot == 0) bulk_wait_read<0>();
    __syncthreads();
    if (bulk && slot == 0 && t > 0) issue_stores(t - 1);

    // ---- round 0: handle_batch over this receiver's mailbox row (batch order = sender order)
    const uint32_t any_sent = sm.sent[0] | sm.sent[1] | sm.sent[2] | sm.sent[3];
    if (any_sent && sp.round_limit == 0) fault = fault ? fault : PHX_FAULT_ROUND_LIMIT;
    int ack_recv = -1;
    int first_sender = -1;
    // everybody sent over a complete graph: every receiver's batch is "all agents but me"
    const bool reduce_form = !TRACK && sp.complete && sp.round_limit != 0 && n >= 2 &&
                             sm.sent[0] == full_w[0] && sm.sent[1] == full_w[1] &&
                             sm.sent[2] == full_w[2] && sm.sent[3] == full_w[3];
    if (reduce_form) {  // block-uniform
      // merge the four warps' (best, second best) of every class into the block's: lane c of
      // each warp merges class c (five lanes busy once) and the result reaches the other lanes by
      // shuffle -- 10 SHFL instead of the same 70-instruction merge on all 128 lanes
      int g1m = -1, g2m = -1;
      if (lane < 5) {
        const int4 q01 = *reinterpret_cast<const int4*>(&sm.wtop[lane][0]);  // warps 0, 1
        const int4 q23 = *reinterpret_cast<const int4*>(&sm.wtop[lane][2]);  // warps 2, 3
        g1m = max(max(q01.x, q01.z), max(q23.x, q23.z));
        g2m = max(max(q01.x == g1m ? q01.y : q01.x, q01.z == g1m ? q01.w : q01.z),
                  max(q23.x == g1m ? q23.y : q23.x, q23.z == g1m ? q23.w : q23.z));
      }
      int g1c[5], g2c[5];
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        g1c[c] = __shfl_sync(0xFFFFFFFFu, g1m, c);
        g2c[c] = __shfl_sync(0xFFFFFFFFu, g2m, c);
      }
      if (is_agent) {
        int best = INT32_MIN, best_s = -1;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const int g1 = g1c[c], g2 = g2c[c];
          const int k = (127 - (g1 & 127)) == slot ? g2 : g1;  // "but me"
          if (k >= 0) {
            const int v = (k >> 7) + tailtab[c], sdr = 127 - (k & 127);
            if (v > best || (v == best && sdr < best_s)) {  // the FIRST sender attaining the max
              best = v;
              best_s = sdr;
            }
          }
        }
        const int4 ws = *reinterpret_cast<const int4*>(sm.wsum);
        st[1] = ws.x + ws.y + ws.z + ws.w - st[0] + tails_all - (myclass + 3 * slot) % 5;
        st[2] = best;
        st[3] = best_s;
        ack_recv = best_s;
        if (!ok_send_ack ||
            (!ack_recv_all_ok && !((sp.receiver_ok[DN_ACK][best_s >> 5] >> (best_s & 31)) & 1u)))
          fault = fault ? fault : PHX_FAULT_BAD_PAYLOAD_TYPE;
      }
    } else if (is_agent && sp.round_limit != 0) {
      int total = 0, best = INT32_MIN, best_s = -1;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t m = adj[w] & sm.sent[w];  // symmetric graph: senders with an edge to me
        if (TRACK && first_sender < 0 && m) first_sender = w * 32 + __ffs(m) - 1;
        if (m == 0u) continue;
        if (m == 0xFFFFFFFFu) {  // dense graphs: all 32 senders of this word, no per-bit tests
#pragma unroll
          for (int b4 = 0; b4 < 32; b4 += 4) {
            const int4 q = *reinterpret_cast<const int4*>(&sm.vals[w * 32 + b4]);  // broadcast
            const int qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int sdr = w * 32 + b4 + k;
              const int v = qv[k] + tailtab[(2 * sdr) % 5];
              total += v;
              if (v > best) {
                best = v;
                best_s = sdr;
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int b4 = 0; b4 < 32; b4 += 4) {
          const int4 q = *reinterpret_cast<const int4*>(&sm.vals[w * 32 + b4]);  // broadcast
          const int qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int sdr = w * 32 + b4 + k;
            if ((m >> (b4 + k)) & 1u) {
              const int v = qv[k] + tailtab[(2 * sdr) % 5];
              total += v;
              if (v > best) {  // strict: the FIRST sender attaining the max wins
                best = v;
                best_s = sdr;
              }
            }
          }
        }
      }
      if (best_s >= 0) {
        st[1] = total;
        st[2] = best;
        st[3] = best_s;
        ack_recv = best_s;
        if (!ok_send_ack || (!((sp.receiver_ok[DN_ACK][best_s >> 5] >> (best_s & 31)) & 1u) &&
                             !(sp.flags & PHX_FLAG_NO_PAYLOAD_CHECKS)))
          fault = fault ? fault : PHX_FAULT_BAD_PAYLOAD_TYPE;
      }
    }
    if (TRACK) {
      sm.ack[slot] = make_int2(ack_recv, st[2]);
      sm.order_key[slot] = first_sender < 0 ? 0x7FFFFFFF : first_sender * DN_MAX + slot;
    }
    // round 1, pushed: Ack(best) lands on its receiver's counters
    if (ack_recv >= 0 && sp.round_limit != 1) {
      atomicAdd(&sm.ack_cnt[ack_recv], 1);
      atomicAdd(&sm.ack_sum[ack_recv], st[2]);
    }
    bool any_ack;
    if (reduce_form) {  // every agent answered (n >= 2): no vote needed
      any_ack = true;
      __syncthreads();
    } else {
      const uint32_t acks_w = __ballot_sync(0xFFFFFFFFu, ack_recv >= 0);
      if (slot == 0) sm.any_ack = 0;
      __syncthreads();
      if (lane == 0 && acks_w) atomicOr(&sm.any_ack, 1u);
      __syncthreads();
      any_ack = sm.any_ack != 0;
    }
    if (any_ack && sp.round_limit == 1) fault = fault ? fault : PHX_FAULT_ROUND_LIMIT;

    // ---- round 1: the Acks (same handle_batch override): count and sum
    if (is_agent && sp.round_limit != 1 && sp.round_limit != 0) {
      st[4] = sm.ack_cnt[slot];
      st[5] = sm.ack_sum[slot];
    }

    if (TRACK && slot == 0) {  // Resolver.tracked_messages of this step, global push order
      int cnt = 0;
      int4* rows = a.trace.rows + row * a.trace.cap;  // one slab of `cap` rows per (step, env)
      for (int s = 0; s < n; ++s) {
        if (!((sm.sent[s >> 5] >> (s & 31)) & 1u)) continue;
        for (int r = 0; r < n; ++r)
          if ((a.adj[s * 4 + (r >> 5)] >> (r & 31)) & 1u) {
            if (cnt < a.trace.cap)
              rows[cnt] = trace_row(s, r, DN_SIGNAL, dn_tailored(sm.vals[s], s, r), 0, 0);
            ++cnt;
          }
      }
      // Acks are pushed in the receivers' first-arrival order: (first sender, slot)
      int last = -1;
      for (;;) {
        int best_r = -1, best_k = 0x7FFFFFFF;
        for (int r = 0; r < n; ++r) {
          const int k = sm.order_key[r];
          if (k > last && k < best_k) { best_k = k; best_r = r; }
        }
        if (best_r < 0) break;
        last = best_k;
        if (sm.ack[best_r].x >= 0) {
          if (cnt < a.trace.cap)
            rows[cnt] = trace_row(best_r, sm.ack[best_r].x, DN_ACK, sm.ack[best_r].y, 0, 1);
          ++cnt;
        }
      }
      a.trace.cnt[row] = cnt;
    }

    // ---- outputs (env.py:273-303): every agent observes and is rewarded; nobody terminates
    const bool at_max = h.x == sp.num_steps;
    // the reward belongs to the step that just ran: computed BEFORE an auto-reset clears the
    // state (only the observation is replaced by the reset observation, as in the engines)
    const float rew = (float)st[5] * (1.0f / 1024.0f);
    if ((sp.flags & PHX_FLAG_AUTO_RESET) && at_max) {
      h.x = 0;
      h.y += 1;
#pragma unroll
      for (int w = 0; w < DN_WORDS; ++w) st[w] = 0;
      st[3] = -1;
    }
    const float o0 = (float)st[1] * (1.0f / 131072.0f);  // powers of two: exact
    const float o1 = (float)st[2] * (1.0f / 1024.0f);
    const float o2 = (float)st[4] * (1.0f / 128.0f);
    if (bulk) {
      // (this buffer's previous rows were read out before the step's first barrier)
      sg.obs[slot * 3 + 0] = o0;
      sg.obs[slot * 3 + 1] = o1;
      sg.obs[slot * 3 + 2] = o2;
      sg.reward[slot] = rew;  // (the four mask planes were staged once, before the loop)
      fence_async_smem();  // issued by thread 0 after the NEXT barrier (issue_stores)
    } else if (is_agent) {
      const size_t o = row * n + slot;
      if (a.io.obs) {
        a.io.obs[o * 3 + 0] = o0; a.io.obs[o * 3 + 1] = o1; a.io.obs[o * 3 + 2] = o2;
      }
      if (a.io.reward) a.io.reward[o] = rew;
      if (a.io.obs_mask) a.io.obs_mask[o] = 1;
      if (a.io.reward_mask) a.io.reward_mask[o] = 1;
      if (a.io.term) a.io.term[o] = 0;
      if (a.io.trunc) a.io.trunc[o] = 0;
    }
    if (slot == 0 && a.io.all_done)
      reinterpret_cast<uchar2*>(a.io.all_done)[row] = make_uchar2(n == 0, at_max ? 1 : 0);
  }
  __syncthreads();  // the last step's staged rows are complete
  if (slot == 0) {
    if (bulk) {
      issue_stores(a.T - 1);
      bulk_wait<0>();
    }
    a.hdr[e] = h;
  }
#pragma unroll
  for (int w = 0; w < DN_WORDS; ++w) a.state[((size_t)w * sp.E + e) * DN_MAX + slot] = st[w];
  // first fault of the env (lowest slot)
  const uint32_t fw = __ballot_sync(0xFFFFFFFFu, fault != 0);
  __shared__ uint32_t fault_slot;
  if (slot == 0) fault_slot = 0xFFFFFFFFu;
  __syncthreads();
  if (fw && lane == (__ffs(fw) - 1)) atomicMin(&fault_slot, (uint32_t)slot);
  __syncthreads();
  if (fault && fault_slot == (uint32_t)slot) raise_fault(a.faults, e, fault);
}

__global__ void dense_reset_kernel(DenseSpec sp, int4* hdr, int32_t* state, const uint8_t* env_mask,
                                   float* obs, uint8_t* obs_mask, bool init_only) {
  const int e = blockIdx.x, slot = threadIdx.x;
  if (env_mask && env_mask[e] == 0) return;
  if (init_only) {
    if (slot == 0) hdr[e] = make_int4(0, -1, 0, 0);
  } else if (slot == 0) {
    int4 h = hdr[e];
    h.x = 0;
    h.y += 1;
    hdr[e] = h;
  }
  for (int w = 0; w < DN_WORDS; ++w)
    state[((size_t)w * sp.E + e) * DN_MAX + slot] = w == 3 ? -1 : 0;  // DenseAgent.reset
  if (!init_only && slot < sp.n) {
    const size_t o = (size_t)e * sp.n + slot;
    if (obs) { obs[o * 3] = 0.f; obs[o * 3 + 1] = 0.f; obs[o * 3 + 2] = 0.f; }
    if (obs_mask) obs_mask[o] = 1;
  }
}

class DenseFamily final : public Family {
 public:
  ~DenseFamily() override {
    cudaFree(d_state);
    cudaFree(d_adj);
  }

  int32_t init(const phx_spec& s) override {
    PHX_REQUIRE(s.env_kind == PHX_ENV_BASE, PHX_ERR_UNSUPPORTED,
                "dense family runs under PhantomEnv (PHX_ENV_BASE) only");
    PHX_REQUIRE(s.n_agents <= DN_MAX && s.n_strategic == s.n_agents, PHX_ERR_UNSUPPORTED,
                "dense family: up to 128 agents, all strategic DenseAgents");
    PHX_REQUIRE(s.obs_dim == 3 && s.act_dim == 1 && s.n_payload_types == 2, PHX_ERR_INVALID,
                "dense family: obs_dim 3, act_dim 1, 2 payload types");
    PHX_REQUIRE(!(s.flags & PHX_FLAG_IGNORE_CONNECTION_ERRORS), PHX_ERR_UNSUPPORTED,
                "dense family: ignore_connection_errors has no effect (agents only address "
                "neighbours) and is not accepted");
    PHX_REQUIRE(!(s.flags & (PHX_FLAG_STOCHASTIC_NETWORK | PHX_FLAG_SHUFFLE_BATCHES)),
                PHX_ERR_UNSUPPORTED,
                "dense family: StochasticNetwork / shuffle_batches are served by the queue engine "
                "(<= 32 agents) only");
    std::memset(&dsp, 0, sizeof(dsp));
    dsp.E = E;
    dsp.n = s.n_agents;
    dsp.num_steps = s.num_steps;
    dsp.round_limit = s.round_limit;
    dsp.flags = s.flags;
    for (int t = 0; t < 2; ++t)
      for (int w = 0; w < 4; ++w) {
        dsp.sender_ok[t][w] = s.type_sender_ok[t][w];
        dsp.receiver_ok[t][w] = s.type_receiver_ok[t][w];
      }
    uint32_t h_adj[DN_MAX * 4];
    std::memset(h_adj, 0, sizeof(h_adj));
    dsp.complete = 1;
    for (int i = 0; i < s.n_agents; ++i) {
      for (int w = 0; w < 4; ++w) h_adj[i * 4 + w] = s.adjacency[i][w];
      for (int j = 0; j < s.n_agents; ++j) {
        PHX_REQUIRE(mask_bit(s.adjacency[i], j) == mask_bit(s.adjacency[j], i), PHX_ERR_INVALID,
                    "dense family needs a symmetric graph (Network.add_connection always is)");
        if (mask_bit(s.adjacency[i], j) != (i != j)) dsp.complete = 0;
      }
    }
    PHX_CUDA(cudaMalloc(&d_adj, sizeof(h_adj)));
    PHX_CUDA(cudaMemcpy(d_adj, h_adj, sizeof(h_adj), cudaMemcpyHostToDevice));
    PHX_CUDA(cudaMalloc(&d_state, sizeof(int32_t) * DN_WORDS * (size_t)E * DN_MAX));
    dense_reset_kernel<<<E, DN_MAX>>>(dsp, d_hdr, d_state, nullptr, nullptr, nullptr, true);
    PHX_CUDA(cudaGetLastError());
    PHX_CUDA(cudaDeviceSynchronize());
    if (tracking())
      PHX_REQUIRE(s.trace_capacity >= s.n_agents * s.n_agents, PHX_ERR_INVALID,
                  "trace_capacity must be >= n_agents^2 for the dense family");
    PHX_CUDA(cudaFuncSetAttribute(dense_step_kernel<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DenseSmem)));
    PHX_CUDA(cudaFuncSetAttribute(dense_step_kernel<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DenseSmem)));
    return PHX_OK;
  }

  int32_t reset(const uint8_t* env_mask, float* obs, uint8_t* obs_mask,
                cudaStream_t stream) override {
    dense_reset_kernel<<<E, DN_MAX, 0, stream>>>(dsp, d_hdr, d_state, env_mask, obs, obs_mask, false);
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  int32_t rollout(int32_t T, const StepIO& io, cudaStream_t stream) override {
    if (tracking()) {
      const int32_t rc = ensure_trace(T);
      if (rc != PHX_OK) return rc;
    }
    DenseArgs a;
    a.sp = dsp;
    a.T = T;
    a.hdr = d_hdr;
    a.state = d_state;
    a.adj = d_adj;
    a.io = io;
    a.faults = fault_sink();
    a.trace = trace_sink();
    if (tracking()) dense_step_kernel<true><<<E, DN_MAX, sizeof(DenseSmem), stream>>>(a);
    else dense_step_kernel<false><<<E, DN_MAX, sizeof(DenseSmem), stream>>>(a);
    PHX_CUDA(cudaGetLastError());
    return PHX_OK;
  }

  int32_t family_field(int32_t field, int32_t, void** p, size_t* bytes) override {
    const int w = field - PHX_FIELD_FAMILY;
    if (w >= 0 && w < DN_WORDS) {
      *p = d_state + (size_t)w * E * DN_MAX;
      *bytes = sizeof(int32_t) * (size_t)E * DN_MAX;
      return PHX_OK;
    }
    set_error("dense family: unknown field " + std::to_string(field));
    return PHX_ERR_INVALID;
  }

  const char* exec_name() const override { return "block-per-env(G=128, pull/push)"; }

 private:
  DenseSpec dsp{};
  int32_t* d_state = nullptr;
  uint32_t* d_adj = nullptr;
};

}  // namespace

Family* make_dense_family(const phx_spec&) { return new DenseFamily(); }

}  // namespace phx
